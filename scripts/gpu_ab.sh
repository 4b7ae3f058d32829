#!/bin/bash
# A/B two builds of the library on the C4 bench: gpu_ab.sh <pytest -k expr> <lib-or-"default">[:variant] ...
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
KEXPR="$1"; shift
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$KEXPR" 2>&1 | tail -4
for LV in "$@"; do
  L=${LV%%:*}; V=""; [[ "$LV" == *:* ]] && V=${LV##*:}
  TAG=$(basename "$L" .so)_v$V
  if [ "$L" = "default" ]; then unset SBO_B200_LIB; else export SBO_B200_LIB=$PWD/$L; fi
  SBO_FANTASY_VARIANT=$V timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-peaks > gpurun_out/bench_ab_$TAG.json 2> gpurun_out/bench_ab_$TAG.err; tail -3 gpurun_out/bench_ab_$TAG.err
  python - $TAG <<'PY'
import json, sys
d=json.load(open(f"gpurun_out/bench_ab_{sys.argv[1]}.json"))
print(sys.argv[1], "ms_per_step", round(d["ms_per_step"],1), {k: round(v,2) for k,v in d["phase_ms"].items()}, "TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "x_new", d["config"]["x_new_idx"], "n_hit", d["config"]["n_hit"], d["clocks"]["sm_mhz"])
PY
done
