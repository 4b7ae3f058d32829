#!/bin/bash
# round 2, probe 8 (1 GPU): one-evaluation refine epilogue + constraint-only FP64 rows: tests and the default bench
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "^C4 tf32|^C5 model tf32|FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
B="--no-cpu-baseline --no-reference-configs --no-peaks --no-lipschitz-steps --steps 5 --warmup 3"
for v in -1 5; do
echo "== default bench variant $v"
SBO_FANTASY_VARIANT=$v timeout 600 python bench.py $B > gpurun_out/r02_c4_default_v$v.json 2> gpurun_out/r02_c4_default_v$v.err; echo "rc=$?"
python - <<PY
import json
try:
    r = json.loads(open("gpurun_out/r02_c4_default_v$v.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["roofline"]["achieved"], r["roofline"]["frac"], r["config"]["n_hit"], r["config"]["refined_pairs_fp64"], r["e2e"]["ms_per_step"], r["clocks"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_c4_default_v$v.err").read()[-1500:])
PY
done
echo done
