"""Pair-constraint values produced by the REFERENCE'S OWN SOURCE for the model states of ref_*.npz.

Run HERE (build container): python tests/golden/make_reference_pairs.py  ->  tests/golden/ref_pairs.npz
Loads the model state the reference's own fit produced (tests/golden/ref_c1_benoit.npz, ref_c3_wor.npz) back into the
reference's unmodified `models/SafeOpt.BO` (over the NumPy jax stand-in, refshim/README.md) and records, for random
(x, z) pairs of the box, the values of the constraint lambdas the reference hands to SciPy in `Expander()` / `Target()`
(SafeOpt.py:99-111, GoOSE.py:93-101): `lcb(x, i)` (safe-set test of x), `lcb_constraint_min(z)` (the MAX of the
constraint lcbs, <= 0 for an admissible z) and `Lipschitz_continuity_constraint([x; z], index, L)`.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")

import jax.numpy as jnp  # noqa: E402  (the stand-in)
from models import SafeOpt  # noqa: E402  (the reference's own module)

out = {}
rng = np.random.default_rng(4242)
for name, n in (("ref_c1_benoit", 9), ("ref_c1_benoit", 14), ("ref_c3_wor", 20), ("ref_c3_wor", 35)):
    r = np.load(os.path.join(HERE, name + ".npz"))
    G = r["Y"].shape[1]
    bo = SafeOpt.BO([None] * G, jnp.array(r["bound"]), float(r["beta"]))
    bo.X, bo.Y, bo.kernel = jnp.array(r["X"][:n]), jnp.array(r["Y"][:n]), 'RBF'
    bo.n_point, bo.nx_dim, bo.ny_dim, bo.multi_hyper, bo.var_out = n, 2, G, 5, True
    for k in ("X_mean", "X_std", "Y_mean", "Y_std", "X_norm", "Y_norm", "hypopt"):
        setattr(bo, k, jnp.array(r[f"{k}_{n}"]))
    bo.invKopt = [jnp.array(k) for k in r[f"invKopt_{n}"]]
    bo.update_inference_dataset()
    lo, hi = r["bound"][:, 0], r["bound"][:, 1]
    m = 96
    x = lo + (hi - lo) * rng.random((m, 2))
    # z near x so that both signs of the Lipschitz constraint occur
    z = np.clip(x + (hi - lo) * rng.normal(scale=0.08, size=(m, 2)), lo, hi)
    L = r[f"L_{n}"]
    key = f"{name}_{n}"
    out[key + "_x"], out[key + "_z"], out[key + "_L"] = x, z, L
    out[key + "_lcb_x"] = np.array([[float(bo.lcb(jnp.array(p), i)) for i in range(G)] for p in x])
    out[key + "_lcbmax_z"] = np.array([float(bo.lcb_constraint_min(jnp.array(p))) for p in z])
    out[key + "_lip"] = np.array([[float(bo.Lipschitz_continuity_constraint(jnp.array(np.concatenate([a, b])), idx, L[G - 1]))
                                   for idx in range(1, G)] for a, b in zip(x, z)])
    print(key, "lip >= 0:", int((out[key + "_lip"] >= 0).sum()), "of", out[key + "_lip"].size,
          "admissible z:", int((out[key + "_lcbmax_z"] <= 0).sum()))
np.savez_compressed(os.path.join(HERE, "ref_pairs.npz"), **out)
