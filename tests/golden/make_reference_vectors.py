"""Golden vectors produced by the REFERENCE'S OWN SOURCE (not by the oracle).

Run HERE (the build container, /root/reference mounted):
    python tests/golden/make_reference_vectors.py [--fast]

How: `tests/golden/refshim/` puts NumPy-backed stand-ins for `jax`, `sobol_seq`, matplotlib, imageio and IPython
on sys.path (none of them is installed; see refshim/README.md), then imports the reference's unmodified
`models/GP_Safe.py`, `models/SafeOpt.py`, `models/GoOSE.py` and `problems/*.py` from /root/reference and drives
them exactly as the reference's scripts do:

* `BO(plant_system, bound, b)` + `GP_initialization(X, Y, 'RBF', multi_hyper=5, var_out=True)`
  (test/test_SafeOpt.py:21-33,255-284) on the reference's recorded trajectories (data/*.npz) -> the reference's own
  DE hyper-fit (`GP_Safe.py:194-234`; SciPy's DE is unseeded there, so the global NumPy RNG is seeded here),
  its `hypopt`, `invKopt` and normalisation constants;
* `GP_inference` / `mean` / `lcb` / `ucb` at the points the reference's scripts print at, at random points of the
  box and at 4096 random nodes of the 400x400 plot grid;
* the plot mask `vmap(GP_m.lcb, in_axes=(0, None))(points, 1) > 0.` over the full 400x400 meshgrid, built with the
  statements of `create_data_for_plot` (test/test_SafeOpt.py:324-338), stored packed;
* `Minimizer()`, `Expander()`, `maximize_infnorm_mean_grad(i)` (SafeOpt.py:53-124) and GoOSE's
  `minimize_obj_lcb()`, `Target()`, `explore_safeset()` (GoOSE.py:63-119) -- SciPy DE results, stochastic in the
  reference, deterministic here through the seeded global RNG.  These are continuous-domain optima: the grid path
  is compared to them by containment up to the grid resolution (tests/test_reference_vectors.py).

Outputs: tests/golden/ref_c1_benoit.npz, tests/golden/ref_c3_wor.npz.
"""
import contextlib
import io
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

from make_golden import run_of  # noqa: E402  (stub unpickler for the recorded trajectories)

warnings.filterwarnings("ignore")

import jax.numpy as jnp  # noqa: E402  (the shim)
from jax import vmap  # noqa: E402
from models import GoOSE, SafeOpt  # noqa: E402  (the reference's own modules)
from problems import Benoit_Problem, WilliamOttoReactor_Problem  # noqa: E402

FAST = "--fast" in sys.argv


def quiet(f, *a):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a)


def plot_mask(GP_m, lo, hi, n_side=400):
    """test/test_SafeOpt.py:324-338, statement by statement (bounds are the case's)."""
    x_0 = jnp.linspace(lo[0], hi[0], n_side)
    x_1 = jnp.linspace(lo[1], hi[1], n_side)
    X_0, X_1 = jnp.meshgrid(x_0, x_1)
    X_0_flat = X_0.ravel()
    X_1_flat = X_1.ravel()
    points = jnp.column_stack((X_0_flat, X_1_flat))
    lcb_vmap = vmap(GP_m.lcb, in_axes=(0, None))
    lcb1 = np.asarray(lcb_vmap(points, 1))
    mask_safe = lcb1.reshape(X_0.shape) > 0.
    return np.asarray(points), lcb1, mask_safe


def case(name, plant_system, X, Y, sizes, bound, b, test_points, seed0):
    out = {"X": X, "Y": Y, "sizes": np.array(sizes), "bound": np.array(bound), "beta": np.array(b),
           "test_points": np.array(test_points)}
    lo, hi = np.array(bound)[:, 0], np.array(bound)[:, 1]
    rng = np.random.default_rng(seed0)
    box_pts = lo + (hi - lo) * rng.random((64, 2))
    for n in sizes:
        t0 = time.time()
        np.random.seed(seed0 + n)                           # SciPy DE (unseeded in the reference) draws from here
        GP_s = SafeOpt.BO(plant_system, jnp.array(bound), b)
        GP_s.GP_initialization(jnp.array(X[:n]), jnp.array(Y[:n]), 'RBF', multi_hyper=5, var_out=True)
        ds = GP_s.inference_datasets
        G = GP_s.n_fun
        out[f"hypopt_{n}"] = np.asarray(ds["hypopt"])
        out[f"invKopt_{n}"] = np.stack([np.asarray(k) for k in ds["invKopt"]])
        for k in ("X_mean", "X_std", "Y_mean", "Y_std", "X_norm", "Y_norm"):
            out[f"{k}_{n}"] = np.asarray(ds[k])
        # single-point inference, exactly the call the BO wrappers make (SafeOpt.py:29-45)
        pts = np.vstack([np.array(test_points), box_pts, X[:n]])
        mv = [GP_s.GP_inference_jit(jnp.array(p), ds) for p in pts]
        out[f"pts_{n}"] = pts
        out[f"mean_{n}"] = np.stack([np.asarray(m) for m, _ in mv])
        out[f"var_{n}"] = np.stack([np.asarray(v) for _, v in mv])
        out[f"lcb_{n}"] = np.array([[float(GP_s.lcb(jnp.array(p), i)) for i in range(G)] for p in pts])
        out[f"ucb_{n}"] = np.array([[float(GP_s.ucb(jnp.array(p), i)) for i in range(G)] for p in pts])
        out[f"gradinf_{n}"] = np.array([[float(GP_s.infnorm_mean_grad(jnp.array(p), i)) for i in range(G)]
                                        for p in pts[:8]])
        # plot grid (test_SafeOpt.py:324-338)
        side = 100 if FAST else 400
        points, lcb1, mask = plot_mask(GP_s, lo, hi, side)
        out[f"mask_bits_{n}"] = np.packbits(mask.ravel(), bitorder="little")
        out[f"mask_side_{n}"] = np.array(side)
        sel = rng.choice(points.shape[0], size=4096, replace=False)
        sel.sort()
        mvg = [GP_s.GP_inference_jit(jnp.array(points[p]), ds) for p in sel]
        out[f"grid_idx_{n}"] = sel
        out[f"grid_mean_{n}"] = np.stack([np.asarray(m) for m, _ in mvg])
        out[f"grid_var_{n}"] = np.stack([np.asarray(v) for _, v in mvg])
        out[f"grid_lcb1_{n}"] = lcb1[sel]
        t1 = time.time()
        # SafeOpt acquisition (SafeOpt.py:53-124); one maximize_infnorm_mean_grad per constraint index
        np.random.seed(seed0 + 100 + n)
        xm, sm = quiet(GP_s.Minimizer)
        np.random.seed(seed0 + 200 + n)
        Ls = [float(quiet(GP_s.maximize_infnorm_mean_grad, i)) for i in range(G)]
        np.random.seed(seed0 + 300 + n)
        xe, se = quiet(GP_s.Expander)
        out[f"safeopt_minimizer_x_{n}"] = np.asarray(xm); out[f"safeopt_minimizer_std_{n}"] = np.array(float(sm))
        out[f"safeopt_expander_x_{n}"] = np.asarray(xe); out[f"safeopt_expander_std_{n}"] = np.array(float(se))
        out[f"L_{n}"] = np.array(Ls)
        t2 = time.time()
        # GoOSE acquisition (GoOSE.py:63-119) on the same model state
        GP_g = GoOSE.BO(plant_system, jnp.array(bound), b)
        for k in ("X", "Y", "kernel", "n_point", "nx_dim", "ny_dim", "multi_hyper", "var_out", "X_norm", "Y_norm",
                  "X_mean", "X_std", "Y_mean", "Y_std", "hypopt", "invKopt"):
            setattr(GP_g, k, getattr(GP_s, k))
        GP_g.update_inference_dataset()
        np.random.seed(seed0 + 400 + n)
        xs, ls = quiet(GP_g.minimize_obj_lcb)
        np.random.seed(seed0 + 500 + n)
        zt, lt = quiet(GP_g.Target)
        np.random.seed(seed0 + 600 + n)
        xn = quiet(GP_g.explore_safeset, np.asarray(zt))
        out[f"goose_safe_min_x_{n}"] = np.asarray(xs); out[f"goose_safe_min_lcb_{n}"] = np.array(float(ls))
        out[f"goose_target_z_{n}"] = np.asarray(zt); out[f"goose_target_lcb_{n}"] = np.array(float(lt))
        out[f"goose_explore_x_{n}"] = np.asarray(xn)
        print(f"{name} n={n}: fit+grid {t1 - t0:.1f}s safeopt {t2 - t1:.1f}s goose {time.time() - t2:.1f}s | "
              f"|mask|={int(mask.sum())} min=({np.round(xm, 4)}, {float(sm):.5f}) exp=({np.round(xe, 4)}, {float(se):.5f}) "
              f"L={np.round(Ls, 4)} goose: safe_min={float(ls):.5f} target={np.round(zt, 4)} lcb={float(lt):.5f}",
              flush=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    Xb, Yb = run_of(os.path.join(REF, "data", "data_SafeOpt_Benoit.npz"))
    tps = [[1.45698204, -0.76514894], [1.19497006, -0.74191489], [0.9, -0.6], [10.0, 10.0]]
    case("ref_c1_benoit", [Benoit_Problem.Benoit_System_1, Benoit_Problem.con1_system_tight], Xb, Yb,
         [4, 9, 14], [[-.6, 1.5], [-1., 1.]], 3., tps, 20260000)
    Xw, Yw = run_of(os.path.join(REF, "data", "data_multi_SafeOpt_WilliamOttoReactor.npz"))
    Reactor = WilliamOttoReactor_Problem.WilliamOttoReactor()
    tpw = [[6.9, 83.0], [5.5, 80.0], [4.5, 75.0], [7.0, 100.0]]
    case("ref_c3_wor", [Reactor.get_objective, Reactor.get_constraint1, Reactor.get_constraint2], Xw, Yw,
         [5, 20, 35], [[4., 7.], [70., 100.]], 2., tpw, 20260500)
