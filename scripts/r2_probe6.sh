#!/bin/bash
# round 2, probe 6 (1 GPU): single-pass TF32 + FP64 refinement of the (much wider) TF32 band: cost and exactness
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B="--no-cpu-baseline --no-reference-configs --no-peaks --no-lipschitz-steps --steps 3 --warmup 2 --precision tf32"
for cfg in "2 -1" "2 7" "0 -1"; do
  set -- $cfg
  echo "== tf32 refine=$1 variant=$2"
  SBO_FANTASY_REFINE=$1 SBO_FANTASY_VARIANT=$2 timeout 900 python bench.py $B > gpurun_out/r02_tf32_refine$1_v$2.json 2> gpurun_out/r02_tf32_refine$1_v$2.err; echo "rc=$?"
  python - <<PY
import json
try:
    r = json.loads(open("gpurun_out/r02_tf32_refine$1_v$2.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["config"]["n_hit"], r["config"]["x_new_idx"], r["config"]["refined_pairs_fp64"], r["config"]["refined_safe"], r["e2e"]["ms_per_step"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_tf32_refine$1_v$2.err").read()[-1500:])
PY
done
echo done
