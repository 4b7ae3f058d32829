"""NumPy/SciPy restatement of the Williams-Otto reactor plant (reference
problems/WilliamOttoReactor_Problem.py:19-93): steady state of six mass balances by fsolve.
Host-side black box used as a fixture; not part of the hot path."""
import numpy as np
from scipy.optimize import fsolve


class WilliamOttoReactor:
    def __init__(self, measure_disturbance=False):
        self.rng = np.random.default_rng(42)
        self.measure_disturbance = measure_disturbance
        self._z = 0.0

    def noise_generator(self):
        self._z = float(np.clip(self.rng.normal(), -2.05, 2.05))

    def odecallback(self, w, x, normal_noise):
        xa, xb, xc, xp, xe, xg = w
        Fa = 1.8275
        Fb, Tr = x
        Fb = Fb + normal_noise
        Fr = Fa + Fb
        Vr = 2105.2
        k1 = 1.6599e6 * np.exp(-6666.7 / (Tr + 273))
        k2 = 7.2177e8 * np.exp(-8333.3 / (Tr + 273))
        k3 = 2.6745e12 * np.exp(-11111 / (Tr + 273))
        return [(Fa - Fr * xa - Vr * xa * xb * k1) / Vr,
                (Fb - Fr * xb - Vr * xa * xb * k1 - Vr * xb * xc * k2) / Vr,
                -Fr * xc / Vr + 2 * xa * xb * k1 - 2 * xb * xc * k2 - xc * xp * k3,
                -Fr * xp / Vr + xb * xc * k2 - 0.5 * xp * xc * k3,
                -Fr * xe / Vr + 2 * xb * xc * k2,
                -Fr * xg / Vr + 1.5 * xp * xc * k3]

    def _solve(self, u, noise):
        nn = self._z * np.sqrt(noise)
        sol = fsolve(func=lambda w: self.odecallback(w, u, nn), x0=np.full(6, 0.1))
        return sol, nn

    def get_objective(self, u, noise=0.):
        (xa, xb, xc, xp, xe, xg), nn = self._solve(u, noise)
        Fa, Fb = 1.8275, u[0] + nn
        fx = 1043.38 * xp * (Fa + Fb) + 20.92 * xe * (Fa + Fb) - 79.23 * Fa - 118.34 * Fb
        return (-fx, nn) if self.measure_disturbance else -fx

    def get_constraint1(self, u, noise=0.):
        sol, nn = self._solve(u, noise)
        g = float(0.12 - sol[0])
        return (g, nn) if self.measure_disturbance else g

    def get_constraint2(self, u, noise=0.):
        sol, nn = self._solve(u, noise)
        g = float(0.08 - sol[5])
        return (g, nn) if self.measure_disturbance else g
