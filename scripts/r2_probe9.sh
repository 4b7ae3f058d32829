#!/bin/bash
# round 2, probe 9 (1 GPU): balanced diagonal block in the DMMA solve, single-barrier set reductions: tests + timings
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | head -20
echo "== default bench"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-configs > gpurun_out/r02_c4_default_final.json 2> gpurun_out/r02_c4_default_final.err; echo "rc=$?"
python - <<'PY'
import json
try:
    r = json.loads(open("gpurun_out/r02_c4_default_final.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["roofline"]["achieved"], r["roofline"]["frac"], r["config"]["n_hit"], r["e2e"]["ms_per_step"], r["peaks"])
    for kind in ("safeopt", "goose"):
        l = r["lipschitz_mode"][kind]
        print("lipschitz", kind, l["ms_per_step"], l["phase_ms_rank0"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_c4_default_final.err").read()[-1500:])
PY
echo "== ncu sets kernels at the C5 shard size"
timeout 900 ncu --set full --clock-control none -k regex:"k_sets_pass" -c 2 -f -o /tmp/prof_sets python scripts/c5_shard_probe.py --steps 1 > gpurun_out/ncu_sets.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_sets.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/r02_ncu_sets_c5shard.csv 2>/dev/null
cut -c1-120,300-520 gpurun_out/r02_ncu_sets_c5shard.csv | tail -3
tail -c 900 gpurun_out/ncu_sets.log
echo done
