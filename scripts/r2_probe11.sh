#!/bin/bash
# round 2, probe 11 (1 GPU): table-based cross-covariance kernel A/B + parity tests
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
for t in 1 0; do
SBO_POSTERIOR_TABLES=$t timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-configs --no-peaks --no-lipschitz-steps > gpurun_out/r02_c4_tables$t.json 2> gpurun_out/r02_c4_tables$t.err; echo "tables=$t rc=$?"
python - <<PY
import json
try:
    r = json.loads(open("gpurun_out/r02_c4_tables$t.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["config"]["n_hit"], r["config"]["x_new_idx"], r["config"]["n_safe"], r["config"]["n_min"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_c4_tables$t.err").read()[-1500:])
PY
done
echo done
