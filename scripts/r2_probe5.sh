#!/bin/bash
# round 2, probe 5 (1 GPU): FP64 refinement of the split-TF32 expander (tests + cost), Lipschitz kernels with shared hits
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "^C4 |^C5 |FAILED|^E  " gpurun_out/pytest_gpu.log | head -40
B="--no-cpu-baseline --no-reference-configs --no-peaks --steps 3 --warmup 2"
echo "== bench c4 tf32x3 (refine on)"
timeout 900 python bench.py $B > gpurun_out/r02_c4_x3_refine.json 2> gpurun_out/r02_c4_x3_refine.err; echo "rc=$?"
python - <<'PY'
import json
try:
    r = json.loads(open("gpurun_out/r02_c4_x3_refine.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["config"]["n_hit"], r["config"]["x_new_idx"], r["e2e"]["ms_per_step"])
    for kind in ("safeopt", "goose"):
        l = r["lipschitz_mode"][kind]
        print("lipschitz", kind, l["ms_per_step"], l["kernel_ms_rank0"], l["phase_ms_rank0"], l["pairs_evaluated"], l["n_hit"], l["x_new_idx"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_c4_x3_refine.err").read()[-1500:])
PY
echo done
