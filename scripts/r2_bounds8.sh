#!/bin/bash
# round 2, 8 GPUs: the C5 north-star key in bounds mode (certified expander-set members + undecided candidates)
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-8}
nvidia-smi -L | head -8
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 \
   --no-cpu-baseline --no-peaks --no-lipschitz-steps --no-reference-configs ) > gpurun_out/r02_bounds_c5_$N.json 2> gpurun_out/r02_bounds_c5_$N.err
echo "rc=$?"; tail -6 gpurun_out/r02_bounds_c5_$N.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    r = json.loads([l for l in open(f"gpurun_out/r02_bounds_c5_{n}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print({k: r[k] for k in ("ms_per_step", "value", "n_gpus")})
    print("c5:", json.dumps(r.get("c5"))[:4000])
except Exception as e:
    print("parse error", e)
PY
echo done
