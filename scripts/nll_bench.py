#!/usr/bin/env python
"""Population NLL (hyper-parameter fit, GP_Safe.py:169-224): GPU batch vs the NumPy port, per population."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sbo_b200
from oracle import gp_oracle as O
eng = sbo_b200.GridEngine(0)
out = []
for n, d, P in [(14, 2, 60), (35, 2, 60), (512, 4, 90), (2048, 6, 120)]:
    rng = np.random.default_rng(n)
    Xn = rng.normal(size=(n, d)); y = np.sin(Xn[:, 0]); y = (y - y.mean()) / y.std()
    H = np.column_stack([rng.uniform(-1.5, 1.5, size=(P, d + 1)), rng.uniform(-5., -2., size=P)])
    eng.nll_batch(Xn, y, H)
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps): g = eng.nll_batch(Xn, y, H)
    tg = (time.perf_counter() - t0) / reps
    pc = min(P, 8 if n >= 512 else P)
    t0 = time.perf_counter(); c = np.array([O.negative_loglikelihood(h, Xn, y[:, None]) for h in H[:pc]])
    tc = (time.perf_counter() - t0) * P / pc
    out.append({"n": n, "d": d, "P": P, "gpu_ms_per_population": tg * 1e3, "cpu_port_ms_per_population": tc * 1e3,
                "cores": os.cpu_count(), "max_rel_err": float(np.max(np.abs(g[:pc] - c) / np.maximum(1, np.abs(c))))})
    print(out[-1], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "nll_bench.json"), "w"))
