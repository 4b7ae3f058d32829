#!/bin/bash
# round 2, probe 3 (1 GPU): full GPU test suite, C4 bench in the three precisions (tf32x3 = headline after the segment
# reorder, tf32, fp64 = DMMA kernel), launch list + full ncu capture of the headline GEMM, sets kernels at the C5 shard size
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5; grep -E "^C4 |^C5 |^fp64 dmma|FAILED" gpurun_out/pytest_gpu.log | head -40
B="--no-cpu-baseline --no-reference-configs --no-lipschitz-steps --no-peaks"
for cfg in "tf32x3 3 2" "tf32 3 2" "fp64 1 1"; do
  set -- $cfg
  echo "== bench c4 $1"
  timeout 900 python bench.py $B --precision $1 --steps $2 --warmup $3 > gpurun_out/r02_c4_$1.json 2> gpurun_out/r02_c4_$1.err ; echo "rc=$?"
  python - <<PY
import json
try:
    r = json.loads(open("gpurun_out/r02_c4_$1.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, {k: r["roofline"][k] for k in ("achieved", "frac", "tensor_tflops_issued")}, r["config"]["n_hit"], r["config"]["pairs_evaluated"], r["config"]["x_new_idx"], r["e2e"]["ms_per_step"])
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_c4_$1.err").read()[-1500:])
PY
done
echo "== full default bench (CPU arm, DE context, lipschitz steps)"
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "rc=$?"; tail -c 2500 gpurun_out/r02_bench_default.json; tail -3 gpurun_out/r02_bench_default.err
echo "== reference arm"; ( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "rc=$?"; tail -c 1500 gpurun_out/r02_bench_reference.json; tail -4 gpurun_out/r02_bench_reference.err
P="--steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-peaks --no-reference-configs --no-lipschitz-steps"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_c4_fantasy_tf32x3.csv python bench.py $P > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full capture of the GEMM"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fantasy_tc2 -s 1 -c 1 -f -o gpurun_out/r02_prof_fantasy_tc2_x3 python bench.py $P > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r02_prof_fantasy_tc2_x3.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second > gpurun_out/r02_ncu_fantasy_tc2_x3_summary.csv 2>/dev/null; cut -c1-700 gpurun_out/r02_ncu_fantasy_tc2_x3_summary.csv | tail -3
echo "== ncu sets kernels at the C5 shard size"
timeout 900 ncu --set full --clock-control none -k regex:"k_sets_pass" -c 2 -f -o /tmp/prof_sets python scripts/c5_shard_probe.py --steps 1 > gpurun_out/ncu_sets.log 2>&1; echo "rc=$?"
ncu -i /tmp/prof_sets.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/r02_ncu_sets_c5shard.csv 2>/dev/null
cut -c1-500 gpurun_out/r02_ncu_sets_c5shard.csv | tail -3
rm -f gpurun_out/r02_prof_fantasy_tc2_x3.ncu-rep.tmp; du -sm gpurun_out
echo done
