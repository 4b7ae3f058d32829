"""Reference-SHAPED acquisition step on the CPU  --  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

What one SafeOpt / GoOSE acquisition step costs in the reference: SciPy differential evolution with
``NonlinearConstraint`` lambdas over a SINGLE-POINT ``GP_inference`` (models/SafeOpt.py:47-124,
models/GoOSE.py:63-119).  The reference cannot run on the GPU box (JAX is not installed), so this is a restatement
on top of the oracle's single-point inference (``gp_oracle.gp_inference`` = GP_Safe.py:310-352) and analytic mean
gradient (the reference differentiates with ``jax.grad``).  Same objective lambdas, same constraints, same SciPy
defaults (popsize 15, maxiter 1000, polish=False); the only liberties are a seed (the reference is unseeded) and an
optional ``maxiter`` cap so that bench.py can bound the time (a capped run is reported as a lower bound).
bench.py times it next to the GPU step as context ("B-DE" in BASELINE.md); nothing in the product imports it.
"""
from __future__ import annotations

import time

import numpy as np
from scipy.optimize import NonlinearConstraint, differential_evolution

from . import gp_oracle as O


class DEStep:
    def __init__(self, ds, bound, beta, seed=0, maxiter=1000):
        self.ds, self.bound, self.b = ds, np.asarray(bound, dtype=float), float(beta)
        self.n_fun = ds["Y_norm"].shape[1]
        self.d = self.bound.shape[0]
        self.rng = np.random.default_rng(seed)
        self.maxiter = int(maxiter)
        self.n_inference = 0
        self.capped = False

    # ---- SafeOpt.py:29-45 -------------------------------------------------------------------
    def _inf(self, x):
        self.n_inference += 1
        return O.gp_inference(x, self.ds)

    def ucb(self, x, i):
        m, v = self._inf(x)
        return m[i] + self.b * np.sqrt(v[i])

    def lcb(self, x, i):
        m, v = self._inf(x)
        return m[i] - self.b * np.sqrt(v[i])

    def lcb_constraint_min(self, x):                     # SafeOpt.py:73-77 (returns the max)
        return max(self.lcb(x, i) for i in range(1, self.n_fun))

    def _de(self, fun, bound, cons=()):
        r = differential_evolution(fun, bound, constraints=cons, polish=False, maxiter=self.maxiter,
                                   seed=int(self.rng.integers(1 << 31)))
        if r.nit >= self.maxiter:
            self.capped = True
        return r

    def safe_cons(self, sl=slice(None)):
        return [NonlinearConstraint(lambda x, i=i: self.lcb(x[sl], i), 0.0, np.inf) for i in range(1, self.n_fun)]

    # ---- SafeOpt.py:47-66 -------------------------------------------------------------------
    def minimizer(self):
        cons = self.safe_cons()
        min_ucb = self._de(lambda x: self.ucb(x, 0), self.bound, cons).fun
        cons.append(NonlinearConstraint(lambda x: min_ucb - self.lcb(x, 0), 0.0, np.inf))
        r = self._de(lambda x: -self._inf(x)[1][0], self.bound, cons)
        return r.x, float(np.sqrt(max(-r.fun, 0.0)))

    # ---- SafeOpt.py:68-83 -------------------------------------------------------------------
    def maximize_infnorm_mean_grad(self, i):
        return -self._de(lambda x: -float(np.max(np.abs(O.mean_grad(x.reshape(1, -1), self.ds, i)))), self.bound).fun

    def _pair_cons(self, index):
        d = self.d
        cons = self.safe_cons(slice(0, d))
        cons.append(NonlinearConstraint(lambda x: self.lcb_constraint_min(x[d:]), -np.inf, 0.0))
        L = self.maximize_infnorm_mean_grad(self.n_fun - 1)          # SafeOpt.py:110: the leaked loop variable
        cons.append(NonlinearConstraint(
            lambda x, index=index: self.ucb(x[:d], index) - L * np.linalg.norm(x[:d] - x[d:] + 1e-8), 0.0, np.inf))
        return cons

    # ---- SafeOpt.py:90-124 ------------------------------------------------------------------
    def expander(self):
        d = self.d
        bound2 = np.vstack((self.bound, self.bound))
        best = (None, -np.inf)
        for index in range(1, self.n_fun):
            r = self._de(lambda x: -self._inf(x[:d])[1][0], bound2, self._pair_cons(index))
            std = float(np.sqrt(max(-r.fun, 0.0)))
            if std > best[1]:
                best = (r.x[:d], std)
        return best

    # ---- GoOSE.py:63-119 --------------------------------------------------------------------
    def minimize_obj_lcb(self):
        r = self._de(lambda x: self.lcb(x, 0), self.bound, self.safe_cons())
        return r.x, r.fun

    def target(self):
        d = self.d
        bound2 = np.vstack((self.bound, self.bound))
        best = (None, np.inf)
        for index in range(1, self.n_fun):
            r = self._de(lambda x: self.lcb(x[d:], 0), bound2, self._pair_cons(index))
            if r.fun < best[1]:
                best = (r.x[d:], r.fun)
        return best

    def explore_safeset(self, target):
        return self._de(lambda x: float(np.linalg.norm(x - target)), self.bound, self.safe_cons()).x


def time_step(ds, bound, beta, kind, seed=0, maxiter=1000):
    """Wall seconds of one reference-shaped step: SafeOpt = Minimizer() + Expander() (test/test_SafeOpt.py:144-158),
    GoOSE = minimize_obj_lcb() + Target() + explore_safeset() (test/test_GoOSE.py:151-162)."""
    st = DEStep(ds, bound, beta, seed=seed, maxiter=maxiter)
    t0 = time.perf_counter()
    if kind == "safeopt":
        xm, sm = st.minimizer()
        xe, se = st.expander()
        x_new = xm if sm > se else xe
    else:
        xs, ls = st.minimize_obj_lcb()
        zt, lt = st.target()
        x_new = xs if (ls <= lt or zt is None) else st.explore_safeset(zt)
    return {"seconds": time.perf_counter() - t0, "n_inference": st.n_inference, "x_new": [float(v) for v in np.atleast_1d(x_new)],
            "maxiter": maxiter, "capped": st.capped}
