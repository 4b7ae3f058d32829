"""Drop-in mirror of the reference's ``models/GP_TR.py`` (class ``BO``): trust-region BO.

``minimize_obj_lcb(r, x_0)`` (GP_TR.py:43-51) is the grid pipeline's masked arg-min with ONE extra mask: the safe
set of the step intersected with the trust-region ball ``||x - x_0|| <= r``.  The ball is index geometry of the
meshgrid (built on the host from the same numpy.linspace axes the device uses) and is handed to the device as a user
bitmask (``sbo_set_user_mask``); the lcb_0 values and the reduction are the device's.  ``update_TR``
(GP_TR.py:56-91) is host control flow over two posterior means.  Same constructor and return values as the
reference; an empty intersection returns ``(nan, +inf)``.
"""
from __future__ import annotations

import numpy as np

from .SafeOpt import BO as _GridBO


class BO(_GridBO):
    def __init__(self, plant_system, bound, b, TR_parameters, grid_points_per_dim=None, device=0):
        _GridBO.__init__(self, plant_system, bound, b, grid_points_per_dim=grid_points_per_dim, device=device)
        self.TR_parameters = TR_parameters

    # ------------------------------------------------------------------ grid geometry
    def _ball_mask(self, x_0, r):
        """bool (N,) over the grid, x_0 fastest (test/test_SafeOpt.py:324-334 order): ||x - x_0||_2 <= r."""
        self._ensure_grid()
        axes = [np.linspace(self.bound[k, 0], self.bound[k, 1], self.grid_shape[k]) for k in range(self.bound.shape[0])]
        x_0 = np.asarray(x_0, dtype=np.float64)
        d2 = np.zeros(self.grid_shape[::-1])                     # slowest axis first
        nd = len(axes)
        for k, ax in enumerate(axes):
            shape = [1] * nd
            shape[nd - 1 - k] = ax.shape[0]
            d2 = d2 + ((ax - x_0[k]) ** 2).reshape(shape)
        return (np.sqrt(d2) <= r).ravel()

    # ------------------------------------------------------------------ reference API
    def minimize_obj_lcb(self, r, x_0):
        """GP_TR.py:43-51 -- min lcb_0 over {lcb_i >= 0, i >= 1} within the ball (x_0, r) -> (x, value)."""
        self._ensure_step()
        capi = self._capi()
        self.engine.user_mask_ball(x_0, r, capi.MASK_SAFE)          # S AND ball, on the device (no mask round trip)
        idx, val = self.engine.argreduce(capi.ARGMIN_LCB0, capi.MASK_USER)
        if idx < 0:
            return self._x(-1), np.inf
        return self._x(idx), val

    def TR_constraint(self, x, x_0, r):
        """GP_TR.py:53-54."""
        return r - np.linalg.norm(np.asarray(x, dtype=np.float64) - np.asarray(x_0, dtype=np.float64) + 1e-8)

    def update_TR(self, x_initial, x_new, radius, plant_oldoutput, plant_newoutput):
        """GP_TR.py:56-91 -- accept / reject the step and resize the trust region -> (centre, radius)."""
        p = self.TR_parameters
        rejected = (x_initial, radius * p['radius_red'])
        if any(plant_newoutput[i] < 0. for i in range(1, self.n_fun)):          # :72-76 a plant constraint is violated
            return rejected
        f_old, f_new = plant_oldoutput[0], plant_newoutput[0]
        m_old = self.GP_inference_jit(x_initial, self.inference_datasets)[0][0]
        m_new = self.GP_inference_jit(x_new, self.inference_datasets)[0][0]
        rho = (f_new - f_old) / (m_new - m_old)                                  # :78 actual / predicted change
        if f_old < f_new:                                                        # :80-83 the plant objective went up
            return rejected
        if rho < p['rho_lb']:
            return rejected
        if rho < p['rho_ub']:
            return x_new, radius
        return x_new, min(radius * p['radius_inc'], p['radius_max'])
