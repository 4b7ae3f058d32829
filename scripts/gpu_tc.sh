#!/bin/bash
# tensor-core kernel check: parity tests of the fantasy GEMM, then the C4 bench
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or fantasy_fp64 or expander_and_target or whole_steps or empty_sets" 2>&1 | tail -15
timeout 500 python bench.py --steps 3 --warmup 2 $* > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; tail -3 gpurun_out/bench_tc.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_tc.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"]); print(d["phase_ms"]); print(d["roofline"]); print(d["clocks"]); print(d.get("cpu_baseline"))
PY
