"""GPU, >= 2 devices (skipped otherwise): the sharded step over NCCL returns exactly the single-GPU result.
Launched the way the driver launches bench.py: torch.distributed.run, one rank per GPU."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(cmd):
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("mode,precision", [("fantasy", "tf32"), ("lipschitz", "fp64")])
def test_two_gpus_agree_with_one(mode, precision):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    common = ["--workload", "c4s", "--mode", mode, "--precision", precision, "--steps", "1", "--warmup", "1",
              "--no-cpu-baseline", "--no-peaks"]
    one = _run([sys.executable, "bench.py"] + common)
    two = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                "--master-addr", "127.0.0.1", "--master-port", "29533", "bench.py", "--gpus", "2"] + common)
    for k in ["n_safe", "n_unsafe", "n_min", "pairs", "n_hit", "x_new_idx"]:
        assert one["config"][k] == two["config"][k], k
    assert two["n_gpus"] == 2
