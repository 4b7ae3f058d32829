#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, launch list and one ncu capture.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 gpurun_out/pytest_gpu.log
echo "== bench c4s" ; timeout 300 python bench.py --workload c4s --steps 2 --warmup 1 > gpurun_out/bench_c4s.json 2> gpurun_out/bench_c4s.err ; echo "rc=$?" ; tail -c 1500 gpurun_out/bench_c4s.json ; tail -5 gpurun_out/bench_c4s.err
echo "== bench c4" ; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err ; echo "rc=$?" ; tail -c 3000 gpurun_out/bench_c4.json ; tail -5 gpurun_out/bench_c4.err
if [ "$1" == "ncu" ]; then
  echo "== ncu launches (c4s)"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4s.csv \
     python bench.py --workload c4s --steps 1 --warmup 1 --no-cpu-baseline --no-peaks > gpurun_out/ncu_launches.log 2>&1 ; echo "rc=$?"
fi
echo done
