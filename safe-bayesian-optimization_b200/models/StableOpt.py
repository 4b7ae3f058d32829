"""Mirror of the inner/outer optimisations of the reference's ``models/StableOpt.py`` (class ``BO``) on the grid
pipeline.

The reference models the plant outputs over the joint input ``x = (x_c, d)`` -- controlled inputs and disturbance --
with the zero-prior-mean GP of ``models/GP_Robust.py`` and answers
  Maximise_d(fun, xc, i)   max_d fun_i(xc, d)      5 SLSQP starts + DE fallback per call        (StableOpt.py:96-115)
  Minimise_d(fun, xc, i)   min_d fun_i(xc, d)                                                   (:117-136)
  Minimize_Maximise(fun)   min over {xc : min_d lcb_i(xc,d) >= 0 for all i} of max_d fun_0      (:138-154; a DE whose
                           every individual runs 5 + 5(G-1) SLSQP solves)
Here the joint space is one meshgrid (controlled axes first = fastest): ONE posterior pass over it, one column
reduction over the disturbance axes per x_c (``sbo_stable_minmax``) and a masked arg-min.  Same constructor
``BO(plant_system, bound, bound_d, b)`` and the same return values; results are grid points.
Additive keyword: ``grid_points_per_dim`` (default 60 per controlled axis, 40 per disturbance axis).
"""
from __future__ import annotations

import numpy as np

from .GP_Safe import GP
from ._boot import package


class BO(GP):
    def __init__(self, plant_system, bound, bound_d, b, grid_points_per_dim=None, device=0):
        GP.__init__(self, plant_system, device=device)
        self.bound = np.asarray(bound, dtype=np.float64)
        self.bound_d = np.asarray(bound_d, dtype=np.float64)
        self.nxc_dim = self.bound.shape[0]
        self.nd_dim = self.bound_d.shape[0]
        self.b = b
        self.GP_inference_jit = self.GP_inference
        if grid_points_per_dim is None:
            grid_points_per_dim = [60] * self.nxc_dim + [40] * self.nd_dim
        elif np.isscalar(grid_points_per_dim):
            grid_points_per_dim = [int(grid_points_per_dim)] * (self.nxc_dim + self.nd_dim)
        self.grid_shape = tuple(int(p) for p in grid_points_per_dim)
        self._grid_set = False
        self._post = False
        self.engine.set_option("prior_mean_zero", 1)          # GP_Robust.py:322-323: zero prior mean for every output

    def _on_model_changed(self):
        self._post = False

    def _ensure_post(self):
        self._ensure_uploaded(self.inference_datasets)
        if not self._grid_set:
            lo = np.concatenate([self.bound[:, 0], self.bound_d[:, 0]])
            hi = np.concatenate([self.bound[:, 1], self.bound_d[:, 1]])
            self.engine.set_grid(lo, hi, list(self.grid_shape))
            self._grid_set = True
            self._post = False
        if not self._post:
            self.engine.posterior(with_grad=False, keep_v=0, fetch=False)
            self._post = True

    def _xc(self, idx):
        if idx < 0:
            return np.full(self.nxc_dim, np.nan)
        return self.engine.point_coords(idx)[: self.nxc_dim]      # x_c axes are the fastest: index idx has d-index 0

    # ------------------------------------------------------------------ StableOpt.py:64-94
    def _check(self, xc, d):
        if np.ndim(xc) != 1 or np.ndim(d) != 1:
            raise ValueError("xc or d needs to be in 1d")

    def mean(self, xc, d, i):
        self._check(xc, d)
        return self.GP_inference_jit(np.concatenate((xc, d)), self.inference_datasets)[0][i]

    def ucb(self, xc, d, i):
        self._check(xc, d)
        m, v = self.GP_inference_jit(np.concatenate((xc, d)), self.inference_datasets)
        return m[i] + self.b * np.sqrt(v[i])

    def lcb(self, xc, d, i):
        self._check(xc, d)
        m, v = self.GP_inference_jit(np.concatenate((xc, d)), self.inference_datasets)
        return m[i] - self.b * np.sqrt(v[i])

    # ------------------------------------------------------------------ StableOpt.py:96-136: inner problems at any xc
    def _d_grid(self):
        axes = [np.linspace(self.bound_d[k, 0], self.bound_d[k, 1], self.grid_shape[self.nxc_dim + k]) for k in range(self.nd_dim)]
        mesh = np.meshgrid(*axes[::-1], indexing="ij")
        return np.column_stack([mesh[self.nd_dim - 1 - k].ravel() for k in range(self.nd_dim)])

    def _fun_over_d(self, fun, xc, i):
        name = getattr(fun, "__name__", None)
        if name not in ("ucb", "lcb", "mean"):
            raise ValueError("fun needs to be either self.ucb, lcb or mean")
        D = self._d_grid()
        pts = np.hstack([np.tile(np.asarray(xc, dtype=np.float64), (D.shape[0], 1)), D])
        m, v = self.GP_inference_batch(pts)                         # one batched device call over the disturbance grid
        if name == "mean":
            return m[:, i]
        s = self.b * np.sqrt(v[:, i])
        return m[:, i] + s if name == "ucb" else m[:, i] - s

    def Maximise_d(self, fun, xc, i):
        return float(np.max(self._fun_over_d(fun, xc, i)))

    def Minimise_d(self, fun, xc, i):
        return float(np.min(self._fun_over_d(fun, xc, i)))

    # ------------------------------------------------------------------ StableOpt.py:138-154
    def Minimize_Maximise(self, fun):
        name = getattr(fun, "__name__", None)
        if name not in ("ucb", "lcb", "mean"):
            raise ValueError("fun needs to be either self.ucb, lcb or mean")
        self._ensure_post()
        idx, val, self.n_robust_safe = self.engine.stable_minmax(self.nxc_dim, name, self.b)
        return self._xc(idx), (val if idx >= 0 else np.inf)
