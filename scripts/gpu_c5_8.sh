#!/bin/bash
# 8-GPU run (gpurun --gpus 8): C4 at N=8 in the driver's form, then the north-star configuration C5 at N=8
# (one warm-up, one timed step, one end-to-end step: a C5 step is about a minute), then the CPU arm on C5.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/scale_8.json 2> gpurun_out/scale_8.err
echo "C4 N=8 rc=$?"; tail -c 300 gpurun_out/scale_8.err
timeout 560 $TR --master-port 29519 bench.py --gpus 8 --workload c5 --steps 1 --warmup 1 --e2e-steps 1 --no-peaks > gpurun_out/c5_8gpu.json 2> gpurun_out/c5_8gpu.err
echo "C5 N=8 rc=$?"; tail -c 600 gpurun_out/c5_8gpu.err
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8
timeout 120 python bench.py --impl reference --workload c5 --steps 1 --warmup 0 > gpurun_out/c5_reference.json 2> gpurun_out/c5_reference.err
echo "C5 reference rc=$?"
python - <<'PY'
import json
for f in ("scale_8", "c5_8gpu", "c5_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 1), "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d.get("phase_ms"), d.get("roofline", {}).get("achieved"), d.get("clocks"))
    except Exception as e:
        print(f, "no result", e)
PY
