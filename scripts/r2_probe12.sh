#!/bin/bash
# round 2, probe 12 (1 GPU): data-driven absolute error terms of the refining epilogue: tests + cost in tf32 and tf32x3
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
for prec in tf32 tf32x3; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-configs --no-peaks --no-lipschitz-steps --precision $prec > gpurun_out/r02_c4_p12_$prec.json 2> gpurun_out/r02_c4_p12_$prec.err; echo "$prec rc=$?"
python - <<PY
import json
try:
    r = json.loads(open("gpurun_out/r02_c4_p12_$prec.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["config"]["n_hit"], r["config"]["x_new_idx"], r["config"]["refined_pairs_fp64"], r["config"]["refined_safe"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_c4_p12_$prec.err").read()[-1500:])
PY
done
echo done
