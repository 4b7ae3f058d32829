#!/bin/bash
# round 2 multi-GPU session (gpurun --gpus N): 2-GPU agreement tests, then the C4 step at 1 and N GPUs with the
# collectives inside the library (default) and through torch.distributed
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
nvidia-smi -L | head -8
echo "== pytest multi"; timeout 1500 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_multi.log
B="--no-cpu-baseline --no-reference-configs --steps 3 --warmup 2 --no-peaks --c5 0"
timeout 600 python bench.py $B > gpurun_out/r02_multi_c4_1.json 2> gpurun_out/r02_multi_c4_1.err; echo "1 GPU rc=$?"
for ORCH in library torch; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --orchestrator $ORCH $B > gpurun_out/r02_multi_c4_${N}_$ORCH.json 2> gpurun_out/r02_multi_c4_${N}_$ORCH.err
  echo "$N GPUs $ORCH rc=$?"; tail -3 gpurun_out/r02_multi_c4_${N}_$ORCH.err
done
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
def load(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        return None
a = load("gpurun_out/r02_multi_c4_1.json")
for orch in ("library", "torch"):
    b = load(f"gpurun_out/r02_multi_c4_{n}_{orch}.json")
    if not a or not b:
        print(orch, "MISSING", bool(a), bool(b)); continue
    keys = ["n_safe", "n_unsafe", "n_min", "pairs", "n_hit", "x_new_idx"]
    print(orch, "AGREE" if all(a["config"][k] == b["config"][k] for k in keys) else "DIFFER", {k: (a["config"][k], b["config"][k]) for k in keys})
    print("   1 GPU: %.1f ms e2e %.1f  %s" % (a["ms_per_step"], a["e2e"]["ms_per_step"], {k: round(v, 2) for k, v in a["phase_ms"].items()}))
    print("   %s GPU: %.1f ms e2e %.1f %s  speedup %.2f" % (n, b["ms_per_step"], b["e2e"]["ms_per_step"], {k: round(v, 2) for k, v in b["phase_ms"].items()}, a["ms_per_step"] / b["ms_per_step"]))
    for kind in ("safeopt", "goose"):
        la, lb = a["lipschitz_mode"][kind], b["lipschitz_mode"][kind]
        print("   lipschitz %s: %.1f -> %.1f ms (x%.2f), pairs evaluated %.3g -> %.3g, n_hit %d/%d x_new %d/%d" % (
            kind, la["ms_per_step"], lb["ms_per_step"], la["ms_per_step"] / lb["ms_per_step"], la["pairs_evaluated"], lb["pairs_evaluated"],
            la["n_hit"], lb["n_hit"], la["x_new_idx"], lb["x_new_idx"]))
PY
echo done
