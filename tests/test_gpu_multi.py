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


@pytest.mark.parametrize("mode,precision,orchestrator", [
    ("fantasy", "tf32x3", "library"), ("fantasy", "tf32", "torch"), ("lipschitz", "fp64", "library"), ("lipschitz", "fp64", "torch")])
def test_two_gpus_agree_with_one(mode, precision, orchestrator):
    """orchestrator = library: sbo_comm_init + sbo_*_step_sharded (collectives inside the C ABI, csrc/comm.cu);
    torch: sharded.py issues them through torch.distributed.  Both must reproduce the single-GPU step, including the
    reference-exact SafeOpt/GoOSE steps bench.py reports under `lipschitz_mode` (candidate-sharded expander)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    common = ["--workload", "c4s", "--mode", mode, "--precision", precision, "--steps", "1", "--warmup", "1",
              "--no-cpu-baseline", "--no-peaks", "--no-reference-configs"]
    one = _run([sys.executable, "bench.py"] + common)
    two = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                "--master-addr", "127.0.0.1", "--master-port", "29533", "bench.py", "--gpus", "2",
                "--orchestrator", orchestrator] + common)
    for k in ["n_safe", "n_unsafe", "n_min", "pairs", "n_hit", "x_new_idx"]:
        assert one["config"][k] == two["config"][k], k
    assert two["n_gpus"] == 2
    for kind in ("safeopt", "goose"):
        a, b = one["lipschitz_mode"][kind], two["lipschitz_mode"][kind]
        for k in ("pairs", "n_hit", "x_new_idx"):
            assert a[k] == b[k], (kind, k)
        if kind == "safeopt":        # split by candidates: the early exit and the culling do the single-GPU work, shared
            assert b["pairs_evaluated"] <= 1.1 * a["pairs_evaluated"] + 4 * 256 * 256 * 3


def test_bounds_mode_on_two_gpus_brackets_the_exact_set():
    """fantasy_refine = 3: the settled counts AND the per-candidate undecided-pair counts are all-reduced inside the library.
    The error-bound constants are maxima over the LOCAL unsafe points, so the band (and with it the certified / undecided
    split) may move by a few pairs with the sharding -- what must hold on any sharding is the bracket around the exact set."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    common = ["--workload", "c4s", "--mode", "fantasy", "--precision", "tf32", "--steps", "1", "--warmup", "1",
              "--no-cpu-baseline", "--no-peaks", "--no-reference-configs", "--no-lipschitz-steps"]
    exact = _run([sys.executable, "bench.py", "--refine", "2"] + common)["config"]
    one = _run([sys.executable, "bench.py", "--refine", "3"] + common)["config"]
    two = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", "29534", "bench.py", "--gpus", "2", "--orchestrator", "library", "--refine", "3"] + common)["config"]
    for b in (one, two):
        for k in ["n_safe", "n_unsafe", "pairs"]:
            assert exact[k] == b[k], k
        assert b["refined_safe"] == 0
        assert b["n_hit"] <= exact["n_hit"] <= b["n_hit"] + b["n_undecided"]
        assert abs(b["refined_pairs_fp64"] - exact["refined_pairs_fp64"]) <= 0.01 * exact["refined_pairs_fp64"] + 8
    assert abs(one["n_hit"] - two["n_hit"]) <= 0.01 * one["n_hit"] + 8


def test_two_contexts_on_two_devices_in_one_process(oracle):
    """the >48 KB dynamic shared-memory opt-in is per device.  One process, one context per GPU: the DMMA posterior
    (112 KB) and the tcgen05 fantasy GEMM must launch on both and agree."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import sbo_b200
    from sbo_b200 import _capi as capi, workloads
    ds, lo, hi, pts, beta = workloads.small(d=4, pts_per_dim=9, n=200, seed=11, G=4)
    res = []
    for dev in (1, 0, 1):                       # device 1 first: a process-wide flag set by device 0 would hide the bug
        eng = sbo_b200.GridEngine(dev)
        eng.set_grid(lo, hi, pts)
        st = eng.safeopt_step(ds, beta, mode="fantasy", precision="tf32", unsafe_rule=capi.UNSAFE_ANY)
        m, v = eng.posterior()
        res.append((st["n_safe"], st["x_new_idx"], st["expander"]["n_hit"], float(m.sum()), float(v.sum())))
        eng.close()
    assert res[0] == res[1] == res[2]
