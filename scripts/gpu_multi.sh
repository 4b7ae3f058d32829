#!/bin/bash
# multi-GPU check (run with gpurun --gpus N): 1-GPU vs N-GPU results of the same step must agree
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}; shift
nvidia-smi -L | head -8
for WL in c4s c4; do
  for MODE in "fantasy tf32" "lipschitz fp64"; do
    set -- $MODE
    TAG=${WL}_$1
    timeout 600 python bench.py --workload $WL --mode $1 --precision $2 --steps 2 --warmup 1 --no-cpu-baseline --no-peaks > gpurun_out/multi_${TAG}_1.json 2> gpurun_out/multi_${TAG}_1.err
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --mode $1 --precision $2 --steps 2 --warmup 1 --no-cpu-baseline --no-peaks > gpurun_out/multi_${TAG}_$N.json 2> gpurun_out/multi_${TAG}_$N.err
    echo "rc=$? $TAG"; tail -3 gpurun_out/multi_${TAG}_$N.err
    python - "$TAG" "$N" <<'PY'
import json, sys
tag, n = sys.argv[1], sys.argv[2]
def load(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        return None
a, b = load(f"gpurun_out/multi_{tag}_1.json"), load(f"gpurun_out/multi_{tag}_{n}.json")
if not a or not b:
    print(tag, "MISSING RESULT", bool(a), bool(b)); sys.exit(0)
keys = ["n_safe", "n_unsafe", "n_min", "pairs", "x_new_idx"]
same = all(a["config"][k] == b["config"][k] for k in keys)
print(tag, "AGREE" if same else "DIFFER", {k: (a["config"][k], b["config"][k]) for k in keys})
print("   1 GPU: %.1f ms  %s" % (a["ms_per_step"], {k: round(v, 2) for k, v in a["phase_ms"].items()}))
print("   %s GPU: %.1f ms  %s  speedup %.2f" % (n, b["ms_per_step"], {k: round(v, 2) for k, v in b["phase_ms"].items()}, a["ms_per_step"] / b["ms_per_step"]))
PY
  done
done
