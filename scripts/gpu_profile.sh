#!/bin/bash
# One gpurun call: plain run, launch list of one bench step, then one full ncu capture of the named kernel.
#   usage: gpu_profile.sh <kernel-regex> <tag> [bench args...]
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
KREGEX="$1"; TAG="$2"; shift 2
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-peaks --no-reference-configs --no-lipschitz-steps $*"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_${TAG}.log; exit 1; }
tail -c 600 gpurun_out/plain_${TAG}.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 1 -c 1 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"; tail -5 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out/
