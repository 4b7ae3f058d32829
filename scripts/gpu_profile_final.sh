#!/bin/bash
# Round-1 final ncu evidence (one gpurun call): after the plain runs exit 0,
#   (1) launch list + full capture of the fantasy GEMM (k_fantasy_tc2) on the default bench command,
#   (2) speed-of-light / memory / launch sections of every kernel of one Lipschitz-mode step (posterior, sets,
#       arg-reduce, pairs), exported to CSV on the box (the .ncu-rep of 44 kernels is too large to bring back).
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B="--steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-peaks --no-reference-configs --no-lipschitz-steps"
python bench.py $B > gpurun_out/plain_final_fantasy.log 2>&1 || { echo "plain fantasy run failed"; tail -20 gpurun_out/plain_final_fantasy.log; exit 1; }
python bench.py --mode lipschitz --precision fp64 $B > gpurun_out/plain_final_lipschitz.log 2>&1 || { echo "plain lipschitz run failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_final_fantasy.csv python bench.py $B > gpurun_out/ncu_launches_final.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fantasy_tc2 -s 1 -c 1 -f -o gpurun_out/prof_final_fantasy_tc2 python bench.py $B > gpurun_out/ncu_full_final_fantasy.log 2>&1
echo "fantasy capture rc=$?"
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum \
  --clock-control none -k regex:"k_sets_pass|k_argreduce|k_crosscov|k_solve_var|k_pairs_expander|k_tile_bbox" -s 40 -c 44 -f -o /tmp/prof_final_lipschitz_step python bench.py --mode lipschitz --precision fp64 $B > gpurun_out/ncu_full_final_lipschitz.log 2>&1
echo "lipschitz-step capture rc=$?"
ncu -i /tmp/prof_final_lipschitz_step.ncu-rep --page raw --csv > gpurun_out/final_lipschitz_step_raw.csv 2>/dev/null
du -sm gpurun_out
