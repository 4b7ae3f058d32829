#!/bin/bash
# posterior-kernel check: parity tests with the given SBO_POSTERIOR_VARIANT, then the C4 bench phases
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for V in "$@"; do
  echo "=== posterior variant $V"
  SBO_POSTERIOR_VARIANT=$V timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "posterior or model or sets_full or synthetic_small or c4_full or fantasy_fp64 or tensor_core_synthetic" 2>&1 | tail -6
  SBO_POSTERIOR_VARIANT=$V timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-peaks > gpurun_out/bench_post_v$V.json 2> gpurun_out/bench_post_v$V.err; tail -3 gpurun_out/bench_post_v$V.err
  python - $V <<'PY'
import json, sys
d=json.load(open(f"gpurun_out/bench_post_v{sys.argv[1]}.json"))
print("ms_per_step", round(d["ms_per_step"],1), {k: round(v,2) for k,v in d["phase_ms"].items()}, "x_new", d["config"]["x_new_idx"], d["clocks"])
n,dd,G,N=d["config"]["n"],d["config"]["d"],d["config"]["G"],d["config"]["N"]
fl=G*N*(n*n+n*(3*dd+6)); print("posterior TFLOP/s (solve+crosscov):", fl/((d["phase_ms"]["solve"]+d["phase_ms"]["crosscov"])*1e-3)/1e12, " solve only:", G*N*n*n/(d["phase_ms"]["solve"]*1e-3)/1e12)
PY
done
