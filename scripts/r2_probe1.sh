#!/bin/bash
# round 2, probe 1: smoke (with the fantasy step), GPU tests, C4 with/without the exact pruning, tf32 / tf32x3
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B="--no-cpu-baseline --no-reference-configs --no-lipschitz-steps --steps 3 --warmup 3"
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
for cfg in "tf32 1" "tf32 0" "tf32x3 1"; do
  set -- $cfg
  echo "== bench c4 $1 prune=$2"
  timeout 600 python bench.py $B --precision $1 --prune $2 > gpurun_out/r02_c4_$1_prune$2.json 2> gpurun_out/r02_c4_$1_prune$2.err ; echo "rc=$?"
  python - <<PY
import json
try:
    r = json.loads(open("gpurun_out/r02_c4_$1_prune$2.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value", "phase_ms", "roofline")}, r["config"]["n_hit"], r["config"]["pairs_evaluated"], r["config"]["x_new_idx"], r["e2e"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_c4_$1_prune$2.err").read()[-1500:])
PY
done
echo done
