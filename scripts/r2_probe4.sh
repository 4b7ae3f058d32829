#!/bin/bash
# round 2, probe 4 (1 GPU): fused posterior (separable tables) vs the two-kernel path: parity tests + C4 timing
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest gpu (fused default)" ; timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head
B="--no-cpu-baseline --no-reference-configs --no-lipschitz-steps --no-peaks --steps 3 --warmup 2"
for f in 1 0; do
  for m in "lipschitz fp64" "fantasy tf32x3"; do
    set -- $m
    echo "== posterior_fused=$f $1"
    SBO_POSTERIOR_FUSED=$f timeout 600 python bench.py $B --mode $1 --precision $2 > gpurun_out/r02_fused${f}_$1.json 2> gpurun_out/r02_fused${f}_$1.err
    python - <<PY
import json
try:
    r = json.loads(open("gpurun_out/r02_fused${f}_$1.json").read().strip().splitlines()[-1])
    print(r["ms_per_step"], {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["config"]["n_hit"], r["config"]["x_new_idx"], r["config"]["n_safe"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_fused${f}_$1.err").read()[-1200:])
PY
  done
done
echo "== ncu posterior kernels, fused"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum --clock-control none -k regex:"k_solve_fused|k_crosscov|k_build_tables" -c 6 --csv --log-file gpurun_out/r02_ncu_posterior_fused.csv python bench.py --mode lipschitz --precision fp64 --steps 1 --warmup 0 --e2e-steps 1 --no-cpu-baseline --no-peaks --no-reference-configs --no-lipschitz-steps > gpurun_out/ncu_post.log 2>&1; echo "rc=$?"
grep -E "k_solve_fused|k_crosscov|k_build" gpurun_out/r02_ncu_posterior_fused.csv | cut -c1-60,200-420 | head -12
echo done
