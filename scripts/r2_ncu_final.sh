#!/bin/bash
# round 2, final ncu evidence at HEAD: launch list + full capture of the default GEMM + the posterior solve
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
P="--steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-peaks --no-reference-configs --no-lipschitz-steps"
python bench.py $P > gpurun_out/plain_final.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_final.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_c4_default.csv python bench.py $P > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fantasy_tc2 -s 1 -c 1 -f -o gpurun_out/r02_prof_gemm python bench.py $P > gpurun_out/ncu_full.log 2>&1; echo "gemm capture rc=$?"
ncu -i gpurun_out/r02_prof_gemm.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,launch__registers_per_thread > gpurun_out/r02_ncu_fantasy_tc2_tf32r_summary.csv 2>/dev/null
ncu -i gpurun_out/r02_prof_gemm.ncu-rep --page details 2>/dev/null | grep -iE "Duration|SM Frequency|DRAM Throughput|L2 Hit|Executed Ipc|Registers Per|highest-utilized|Tensor|Issue Slots|No Eligible|Mem Pipes" | head -20 > gpurun_out/r02_ncu_fantasy_tc2_tf32r_details.txt
cat gpurun_out/r02_ncu_fantasy_tc2_tf32r_details.txt; tail -1 gpurun_out/r02_ncu_fantasy_tc2_tf32r_summary.csv | cut -c1-500
timeout 900 ncu --set full --clock-control none -k regex:k_solve_var_dmma -s 16 -c 1 -f -o gpurun_out/r02_prof_solve python bench.py $P > gpurun_out/ncu_solve.log 2>&1; echo "solve capture rc=$?"
ncu -i gpurun_out/r02_prof_solve.ncu-rep --page details 2>/dev/null | grep -iE "Duration|SM Frequency|DRAM Throughput|L2 Hit|Executed Ipc|Registers Per|highest-utilized|Tensor|Issue Slots|No Eligible|k_solve" | head -20 > gpurun_out/r02_ncu_solve_details.txt; cat gpurun_out/r02_ncu_solve_details.txt
rm -f gpurun_out/*.ncu-rep
echo done
