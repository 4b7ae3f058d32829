#!/bin/bash
# round 2, final 1-GPU check at HEAD: smoke, full GPU suite, the driver's bench lines (ours + reference arm)
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
echo "== bench (driver line)"; ( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "rc=$?"; tail -4 gpurun_out/r02_bench_final.err
python - <<'PY'
import json
r = json.loads([l for l in open("gpurun_out/r02_bench_final.json").read().splitlines() if l.startswith("{")][-1])
print({k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "gpu_launches", "value_evaluated_pairs_only")})
print({k: round(v, 2) for k, v in r["phase_ms"].items()}); print(r["roofline"]); print(r["e2e"]); print(r["cpu_baseline"]); print(r["clocks"])
PY
echo "== reference arm (driver line)"; ( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err; echo "rc=$?"; tail -4 gpurun_out/r02_bench_reference_final.err; cut -c1-700 gpurun_out/r02_bench_reference_final.json
echo done
