#!/bin/bash
# scaling check (run with gpurun --gpus 8): the driver's own launch line at N = 8, 4, 2, 1 on the default workload
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L | head -8
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  fi
  echo "N=$N rc=$?"; tail -3 gpurun_out/scale_$N.err
  python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/scale_{n}.json").read().strip().splitlines()[-1])
    print("N", n, "ms", round(d["ms_per_step"], 1), "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], {k: round(v, 2) for k, v in d["phase_ms"].items()}, "x_new", d["config"]["x_new_idx"], d["clocks"])
except Exception as e:
    print("N", n, "no result", e)
PY
done
