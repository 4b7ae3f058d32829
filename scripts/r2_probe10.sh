#!/bin/bash
# round 2, probe 10 (1 GPU): DMMA solve pipelined across row blocks: parity tests + posterior timings
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-configs --no-peaks > gpurun_out/r02_c4_probe10.json 2> gpurun_out/r02_c4_probe10.err; echo "rc=$?"
python - <<'PY'
import json
try:
    r = json.loads(open("gpurun_out/r02_c4_probe10.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["config"]["n_hit"], r["e2e"]["ms_per_step"])
    for kind in ("safeopt", "goose"):
        l = r["lipschitz_mode"][kind]
        print("lipschitz", kind, l["ms_per_step"], l["phase_ms_rank0"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_c4_probe10.err").read()[-1500:])
PY
echo done
