#!/bin/bash
# round 2, probe 2 (1 GPU): GPU tests incl. the full-size fixtures, posterior chunk A/B, C5-shard probe (sets kernel at the
# size the judge asked for) with an ncu capture of the set kernels
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B="--no-cpu-baseline --no-reference-configs --no-lipschitz-steps --steps 3 --warmup 2 --no-peaks"
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5; grep -E "^C4 |^C5 |FAILED" gpurun_out/pytest_gpu.log | head -40
for mb in 48 24 96 1024; do
  echo "== posterior chunk ${mb} MB"
  SBO_POSTERIOR_CHUNK_MB=$mb timeout 600 python bench.py $B --mode lipschitz --precision fp64 > gpurun_out/r02_chunk_$mb.json 2> gpurun_out/r02_chunk_$mb.err
  python - <<PY
import json
try:
    r = json.loads(open("gpurun_out/r02_chunk_$mb.json").read().strip().splitlines()[-1])
    print($mb, r["ms_per_step"], {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["roofline"]["achieved"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_chunk_$mb.err").read()[-800:])
PY
done
echo "== c5 shard probe"; timeout 900 python scripts/c5_shard_probe.py --steps 2 > gpurun_out/r02_c5_shard_probe.json 2> gpurun_out/r02_c5_shard_probe.err; echo "rc=$?"; tail -c 1800 gpurun_out/r02_c5_shard_probe.json; tail -3 gpurun_out/r02_c5_shard_probe.err
echo "== ncu sets kernels at the C5 shard size"
timeout 900 ncu --set full --clock-control none -k regex:"k_sets_pass" -c 4 -f -o /tmp/prof_sets python scripts/c5_shard_probe.py --steps 1 > gpurun_out/ncu_sets.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_sets.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/r02_ncu_sets_c5shard.csv 2>/dev/null
cat gpurun_out/r02_ncu_sets_c5shard.csv | cut -c1-400 | head -12
echo done
