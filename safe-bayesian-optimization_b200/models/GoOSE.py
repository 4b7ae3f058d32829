"""Drop-in mirror of the reference's ``models/GoOSE.py`` (class ``BO``), restated on a dense grid.

``minimize_obj_lcb`` / ``Target`` / ``explore_safeset`` (GoOSE.py:63-119) each become one masked
arg-reduction (plus, for ``Target``, the Lipschitz pair kernel with the reduction over the safe side)
of the CUDA grid pipeline.  Same constructor and return values as the reference.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import NonlinearConstraint

from .SafeOpt import BO as _GridBO


class BO(_GridBO):
    def __init__(self, plant_system, bound, b, grid_points_per_dim=None, unsafe_rule='all', device=0):
        _GridBO.__init__(self, plant_system, bound, b, grid_points_per_dim=grid_points_per_dim,
                         expander_mode='lipschitz', precision='fp64', unsafe_rule=unsafe_rule, device=device)
        # GoOSE.py:22-25 -- kept for API compatibility (SciPy constraint objects over lcb_i >= 0)
        self.safe_set_cons = [NonlinearConstraint(lambda x, i=i: self.lcb(x, i), 0., np.inf)
                              for i in range(1, self.n_fun)]

    def minimize_obj_lcb(self):
        """GoOSE.py:63-67 -- min over the safe set of lcb_0 -> (x, value)."""
        s = self._ensure_step()["sets"]
        return self._x(s["min_lcb0_idx"]), s["min_lcb0"]

    def Target(self):
        """GoOSE.py:80-114 -- lowest-lcb_0 unsafe point reachable from the safe set -> (z, lcb_0).
        Empty target set -> (nan, +inf) so the driver's ``min_safe_lcb <= target_lcb`` picks the safe minimiser."""
        st = self._ensure_step()
        if "target" not in st:
            L = np.full(self.n_fun, st["L"][self.n_fun - 1])       # GoOSE.py:100: leaked i = n_fun-1
            st["target"] = self.engine.goose_target(self.b, L)
        tg = st["target"]
        if tg["best_idx"] < 0:
            return self._x(-1), np.inf
        return self._x(tg["best_idx"]), tg["best_value"]

    def explore_safeset(self, target):
        """GoOSE.py:116-119 -- the safe point nearest to ``target``."""
        self._ensure_step()
        capi = self._capi()
        idx, _ = self.engine.argreduce(capi.ARGMIN_DIST, capi.MASK_SAFE, 0, np.asarray(target, dtype=np.float64))
        return self._x(idx)
