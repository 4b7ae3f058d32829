#!/bin/bash
# round 2, probe 7 (1 GPU): refined single-pass TF32 as the default: full GPU tests, smoke, default bench, ncu of the GEMM
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "^C4 |^C5 |FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-400 | head -40
echo "== default bench"
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "rc=$?"
python - <<'PY'
import json
try:
    r = json.loads(open("gpurun_out/r02_bench_default.json").read().strip().splitlines()[-1])
    print({k: r[k] for k in ("ms_per_step", "value", "dtype", "gpu_launches")}, {k: round(v, 2) for k, v in r["phase_ms"].items()})
    print(r["roofline"]); print(r["e2e"]); print(r["cpu_baseline"]); print(r["clocks"])
    print({k: r["config"][k] for k in ("n_hit", "x_new_idx", "pairs_evaluated", "refined_pairs_fp64", "refined_safe")})
    for kind in ("safeopt", "goose"):
        l = r["lipschitz_mode"][kind]
        print("lipschitz", kind, l["ms_per_step"], l["kernel_ms_rank0"], l["pairs_evaluated"])
    for k, v in r["reference_configs"].items():
        print(k, v.get("ms_per_step"), v.get("reference_shaped_de_step"))
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_bench_default.err").read()[-1500:])
PY
B="--no-cpu-baseline --no-reference-configs --no-peaks --no-lipschitz-steps --steps 3 --warmup 2"
echo "== x3 refined"; timeout 600 python bench.py $B --precision tf32x3 > gpurun_out/r02_c4_x3_refined.json 2> gpurun_out/r02_c4_x3_refined.err; python - <<'PY'
import json
r = json.loads(open("gpurun_out/r02_c4_x3_refined.json").read().strip().splitlines()[-1])
print({k: r[k] for k in ("ms_per_step", "value")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["config"]["n_hit"], r["config"]["refined_pairs_fp64"])
PY
P="--steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-peaks --no-reference-configs --no-lipschitz-steps"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_c4_default.csv python bench.py $P > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full capture of the GEMM"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fantasy_tc2 -s 1 -c 1 -f -o gpurun_out/r02_prof_fantasy_tc2_tf32r python bench.py $P > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r02_prof_fantasy_tc2_tf32r.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tc.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,launch__registers_per_thread > gpurun_out/r02_ncu_fantasy_tc2_tf32r_summary.csv 2>/dev/null; cut -c1-900 gpurun_out/r02_ncu_fantasy_tc2_tf32r_summary.csv | tail -3
ncu -i gpurun_out/r02_prof_fantasy_tc2_tf32r.ncu-rep --page details 2>/dev/null | grep -iE "tensor|Duration|DRAM Throughput|L2 Hit|Executed Ipc|Registers|SM Frequency|Pipe" | head -30 > gpurun_out/r02_ncu_fantasy_tc2_tf32r_details.txt; cat gpurun_out/r02_ncu_fantasy_tc2_tf32r_details.txt
rm -f gpurun_out/r02_prof_fantasy_tc2_x3.ncu-rep
du -sm gpurun_out
echo done
