#!/bin/bash
# round 2, 8-GPU session: the driver's launch line at N = 8 (C4 headline + Lipschitz steps + the C5 north-star key)
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-8}
nvidia-smi -L | head -8
( time timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r02_scale_c4_$N.json 2> gpurun_out/r02_scale_c4_$N.err
echo "rc=$?"; tail -6 gpurun_out/r02_scale_c4_$N.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    r = json.loads([l for l in open(f"gpurun_out/r02_scale_c4_{n}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print({k: r[k] for k in ("ms_per_step", "value", "n_gpus")}, {k: round(v, 2) for k, v in r["phase_ms"].items()}, r["e2e"]["ms_per_step"])
    print({k: r["config"][k] for k in ("n_safe", "n_unsafe", "n_min", "n_hit", "x_new_idx", "pairs_evaluated")})
    for kind in ("safeopt", "goose"):
        l = r["lipschitz_mode"][kind]
        print("lipschitz", kind, l["ms_per_step"], l["kernel_ms_rank0"], l["phase_ms_rank0"], l["pairs_evaluated"], l["n_hit"], l["x_new_idx"])
    print("c5:", json.dumps(r.get("c5"))[:3000])
except Exception as e:
    print("parse error", e)
PY
echo done
