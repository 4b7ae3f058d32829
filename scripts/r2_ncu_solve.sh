#!/bin/bash
# ncu of the DMMA posterior solve (C4, EMIT=0): pipe utilisation and warp-state statistics
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
P="--mode lipschitz --precision fp64 --steps 1 --warmup 0 --e2e-steps 1 --no-cpu-baseline --no-peaks --no-reference-configs --no-lipschitz-steps"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_var_dmma -c 1 -f -o gpurun_out/r02_prof_solve python bench.py $P > gpurun_out/ncu_solve.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r02_prof_solve.ncu-rep --page details > gpurun_out/r02_ncu_solve_details.txt 2>/dev/null
grep -E "Duration|SM Frequency|Executed Ipc|Issue Slots Busy|Pipe|pipeline|Stall|stall|Warp Cycles Per Issued|No Eligible|Eligible Warps|Active Warps|L1/TEX Hit|L2 Hit|Shared|Bank|Registers|Theoretical Occ|Achieved Occ|DRAM Throughput|Mem Busy|Max Bandwidth|FP64|Tensor" gpurun_out/r02_ncu_solve_details.txt | head -60
ncu -i gpurun_out/r02_prof_solve.ncu-rep --page raw --csv --metrics smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,smsp__average_warp_latency_issue_stalled_mio_throttle.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio,smsp__average_warp_latency_issue_stalled_lg_throttle.ratio,smsp__average_warp_latency_issue_stalled_not_selected.ratio,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum 2>/dev/null | tail -3 | cut -c1-1800
rm -f gpurun_out/r02_prof_solve.ncu-rep
