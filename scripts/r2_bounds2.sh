#!/bin/bash
# round 2: bounds mode (fantasy_refine = 3) -- single-GPU bracket test against the FP64 kernel, 2-GPU agreement, C4 numbers
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L | head -2
timeout 600 python -m pytest tests/test_gpu_prune.py tests/test_gpu_multi.py -x -q -m gpu -k "bounds" -s 2>&1 | tail -15
for r in 2 3; do
  timeout 300 python bench.py --steps 3 --warmup 3 --refine $r --no-cpu-baseline --no-peaks --no-lipschitz-steps --no-reference-configs > gpurun_out/r02_bounds_c4_refine$r.json 2> gpurun_out/r02_bounds_c4_refine$r.err
  echo "refine=$r rc=$?"
  python - $r <<'PY'
import json, sys
r = json.loads([l for l in open(f"gpurun_out/r02_bounds_c4_refine{sys.argv[1]}.json").read().splitlines() if l.startswith("{")][-1])
c = r["config"]
print(r["ms_per_step"], {k: round(v, 2) for k, v in r["phase_ms"].items()}, {k: c.get(k) for k in ("n_hit", "n_undecided", "refined_pairs_fp64", "refined_safe", "x_new_idx")})
PY
done
echo done
