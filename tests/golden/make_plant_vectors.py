"""Plant outputs produced by the REFERENCE'S OWN problems/*.py (run over the NumPy jax stand-in, see refshim/README.md).

Run HERE (build container): python tests/golden/make_plant_vectors.py  ->  tests/golden/ref_plants.npz
The plants are black-box host functions restated in NumPy under
safe-bayesian-optimization_b200/problems/ as fixtures; this pins those restatements at noise = 0."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")

import jax.numpy as jnp  # noqa: E402  (the stand-in)
from problems import Benoit_Problem, WilliamOttoReactor_Problem  # noqa: E402  (the reference's own modules)

rng = np.random.default_rng(77)
ub = np.column_stack([rng.uniform(-.6, 1.5, 32), rng.uniform(-1., 1., 32)])
uw = np.column_stack([rng.uniform(4., 7., 32), rng.uniform(70., 100., 32)])
R = WilliamOttoReactor_Problem.WilliamOttoReactor()
out = {
    "benoit_u": ub,
    "benoit_f1": np.array([float(Benoit_Problem.Benoit_System_1(jnp.array(u))) for u in ub]),
    "benoit_f2": np.array([float(Benoit_Problem.Benoit_System_2(jnp.array(u))) for u in ub]),
    "benoit_con1": np.array([float(Benoit_Problem.con1_system(jnp.array(u))) for u in ub]),
    "benoit_con1_tight": np.array([float(Benoit_Problem.con1_system_tight(jnp.array(u))) for u in ub]),
    "wor_u": uw,
    "wor_obj": np.array([float(R.get_objective(jnp.array(u), 0.)) for u in uw]),
    "wor_con1": np.array([float(np.ravel(R.get_constraint1(jnp.array(u), 0.))[0]) for u in uw]),
    "wor_con2": np.array([float(np.ravel(R.get_constraint2(jnp.array(u), 0.))[0]) for u in uw]),
}
np.savez_compressed(os.path.join(HERE, "ref_plants.npz"), **out)
print({k: v.shape for k, v in out.items()})
