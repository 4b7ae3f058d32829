"""CPU: the reference arm of bench.py (the oracle port timed on the host cores) keeps the driver's contract."""
import json
import os
import subprocess
import sys

from conftest import ROOT

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(cmd, env=None):
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run([sys.executable, "bench.py", "--impl", "reference", "--workload", "c4s", "--steps", "1", "--warmup", "0"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "expander_pair_evals_per_s" and d["unit"] == "pair-evals/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == os.cpu_count()
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--workload", "c4s", "--steps", "1",
                 "--warmup", "0"], env) == []
    # rank 0 under torchrun: OMP_NUM_THREADS=1 is imposed on the workers, the CPU arm still uses every host core
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1")
    lines = _run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--workload", "c4s", "--steps", "1",
                  "--warmup", "0"], env)
    assert len(lines) == 1 and json.loads(lines[0])["cpu_baseline"]["cores"] == os.cpu_count()


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0"], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
