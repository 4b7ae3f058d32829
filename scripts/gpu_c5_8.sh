#!/bin/bash
# 8-GPU run (gpurun --gpus 8): the north-star configuration C5 at N=8 (one warm-up, one timed step, one end-to-end
# step: a C5 step is about a minute; plus the reference-exact SafeOpt/GoOSE steps), the CPU arm on C5, then C4 at N=8
# in the driver's form.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 700 $TR --master-port 29519 bench.py --gpus 8 --workload c5 --steps 1 --warmup 1 --e2e-steps 1 --no-peaks 2> gpurun_out/c5_8gpu.err | grep "^{" > gpurun_out/c5_8gpu.json
echo "C5 N=8 rc=$?"; tail -c 400 gpurun_out/c5_8gpu.err
timeout 120 python bench.py --impl reference --workload c5 --steps 1 --warmup 0 > gpurun_out/c5_reference.json 2> gpurun_out/c5_reference.err
echo "C5 reference rc=$?"
timeout 300 $TR --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 2> gpurun_out/scale_8.err | grep "^{" > gpurun_out/scale_8.json
echo "C4 N=8 rc=$?"; tail -c 300 gpurun_out/scale_8.err
python - <<'PY'
import json
for f in ("c5_8gpu", "c5_reference", "scale_8"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 1), "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d.get("phase_ms"), d.get("roofline", {}).get("achieved"), d.get("clocks"), d.get("lipschitz_mode"))
    except Exception as e:
        print(f, "no result", e)
PY
