#!/bin/bash
# compare fantasy-kernel variants on the C4 bench (after the tensor-core parity tests)
#   usage: gpu_variants.sh <pytest -k expr> <variant[:gx]> ...
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
KEXPR="$1"; shift
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$KEXPR" 2>&1 | tail -6
for VG in "$@"; do
  V=${VG%%:*}; GXV=""; [[ "$VG" == *:* ]] && GXV=${VG##*:}
  SBO_FANTASY_GX=$GXV SBO_FANTASY_VARIANT=$V timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-peaks > gpurun_out/bench_var$VG.json 2> gpurun_out/bench_var$VG.err; tail -3 gpurun_out/bench_var$VG.err
  python - $VG <<'PY'
import json, sys
d=json.load(open(f"gpurun_out/bench_var{sys.argv[1]}.json"))
print("variant", sys.argv[1], "ms_per_step", round(d["ms_per_step"],1), {k: round(v,2) for k,v in d["phase_ms"].items()}, "TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "x_new", d["config"]["x_new_idx"], d["clocks"]["sm_mhz"])
PY
done
