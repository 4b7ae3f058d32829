#!/usr/bin/env python
"""C5 (d=6, N=16^6=2^24, n=2048, G=4) on ONE GPU as rank `--rank` of `--world`: the rank's full posterior + sets
on its 1/world shard, then the fantasy expander over its LOCAL candidates x its LOCAL unsafe points (1/world^2 of
the job's pairs).  Checks the posterior at full size against the oracle on a random sample and reports times and
device memory -- the probe that sizes the 8-GPU run.   usage: python scripts/c5_shard_probe.py [--world 8]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sbo_b200
from sbo_b200 import workloads
from oracle import gp_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--precision", default="tf32")
a = ap.parse_args()
ds, lo, hi, pts, beta = workloads.c5()
eng = sbo_b200.GridEngine(0)
eng.set_grid(lo, hi, pts)
cnt = eng.set_shard_cyclic(a.rank, a.world, 256)
out = {"workload": "C5", "world": a.world, "rank": a.rank, "local_points": cnt, "precision": a.precision}
free0, total = torch.cuda.mem_get_info(0)
for k in range(a.steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = eng.safeopt_step(ds, beta, mode="fantasy", precision=a.precision)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[f"step{k}_s"] = dt
    out[f"step{k}_phase_ms"] = eng.phase_ms()
free1, _ = torch.cuda.mem_get_info(0)
ex = st["expander"]
out.update({"n_safe": int(st["n_safe"]), "n_unsafe": int(st["n_unsafe"]), "n_min": int(st["n_min"]),
            "pairs_evaluated": int(ex["pairs_evaluated"]), "n_hit": int(ex["n_hit"]), "x_new_idx": int(st["x_new_idx"]),
            "device_mem_used_gb": (free0 - free1) / 2**30, "device_mem_total_gb": total / 2**30})
n, d = ds["X_norm"].shape
flops = ex["pairs_evaluated"] * (2.0 * n + 3 * d + 20)
out["gemm_tflops"] = flops / (out[f"step{a.steps-1}_phase_ms"]["pairs"] * 1e-3) / 1e12
# full-size parity of the posterior on a random sample of this shard (oracle: Cholesky form, FP64)
m, v = eng.posterior(keep_v=0)
rng = np.random.default_rng(0)
sel = np.sort(rng.choice(cnt, size=256, replace=False))
sb = sel // 256
gidx = (sb * a.world + (a.rank + sb + sb // a.world + sb // (a.world * a.world)) % a.world) * 256 + sel % 256
axes = O.grid_axes(lo, hi, pts)
P = np.column_stack([axes[k][(gidx // (16 ** k)) % 16] for k in range(6)])
mo, vo = O.posterior_chol(P, ds)
em = max(np.max(np.abs(m[sel, i] - mo[:, i])) / max(np.max(np.abs(mo[:, i])), ds["Y_std"][i]) for i in range(4))
ev = max(np.max(np.abs(v[sel, i] - vo[:, i])) / (O.unpack_hyper(ds["hypopt"][:, i], 6)[1] * ds["Y_std"][i] ** 2) for i in range(4))
out["posterior_err_vs_oracle"] = {"mean_rel": em, "var_rel": ev, "tol": 1e-10}
assert em <= 1e-10 and ev <= 1e-10, (em, ev)
print(json.dumps(out))
