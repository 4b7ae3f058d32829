"""NumPy/SciPy restatement of the Williams-Otto reactor plant (reference
problems/WilliamOttoReactor_Problem.py:19-93): steady state of six component mass balances found by fsolve.
Host-side black box used as a fixture; not part of the hot path.

Inputs u = (F_B, T_R); species order (A, B, C, P, E, G); reactions A+B->C (k1), B+C->P+E (k2), C+P->G (k3)."""
import numpy as np
from scipy.optimize import fsolve

F_A = 1.8275                    # feed of A
V_R = 2105.2                    # reactor hold-up
ARRHENIUS = np.array([[1.6599e6, 6666.7], [7.2177e8, 8333.3], [2.6745e12, 11111.]])   # (k0, E/R) of k1..k3
PRICES = {"P": 1043.38, "E": 20.92, "A": 79.23, "B": 118.34}


def rate_constants(T_R):
    return ARRHENIUS[:, 0] * np.exp(-ARRHENIUS[:, 1] / (T_R + 273))


def mass_balance_residuals(w, F_B, T_R):
    """d(mass fraction)/dt of the six species; zero at the steady state."""
    xa, xb, xc, xp, xe, xg = w
    k1, k2, k3 = rate_constants(T_R)
    F_R = F_A + F_B
    r1, r2, r3 = k1 * xa * xb, k2 * xb * xc, k3 * xc * xp
    out = -F_R / V_R
    return [F_A / V_R + out * xa - r1,
            F_B / V_R + out * xb - r1 - r2,
            out * xc + 2 * r1 - 2 * r2 - r3,
            out * xp + r2 - 0.5 * r3,
            out * xe + 2 * r2,
            out * xg + 1.5 * r3]


class WilliamOttoReactor:
    def __init__(self, measure_disturbance=False):
        self.rng = np.random.default_rng(42)          # reference: jax.random.PRNGKey(42)
        self.measure_disturbance = measure_disturbance
        # the reference draws jax.random.normal(PRNGKey(42)'s subkey) at construction (:13-15): with noise > 0 the first
        # evaluations are disturbed even before noise_generator() is called.  Same behaviour, NumPy generator.
        self._z = float(np.clip(self.rng.normal(), -2.05, 2.05))

    def noise_generator(self):
        """Draw the next disturbance sample (clipped to +-2.05 as in the reference, :48)."""
        self._z = float(np.clip(self.rng.normal(), -2.05, 2.05))

    def odecallback(self, w, x, normal_noise):
        """Reference name (:19-44): residuals at state w for inputs x with the feed disturbed by normal_noise."""
        return mass_balance_residuals(w, x[0] + normal_noise, x[1])

    def steady_state(self, u, noise):
        shift = self._z * np.sqrt(noise)
        w = fsolve(mass_balance_residuals, np.full(6, 0.1), args=(u[0] + shift, u[1]))
        return w, shift

    def _ret(self, value, shift):
        return (value, shift) if self.measure_disturbance else value

    def get_objective(self, u, noise=0.):
        """Negative profit (minimised): product revenue minus feed cost (:46-62)."""
        w, shift = self.steady_state(u, noise)
        F_B = u[0] + shift
        profit = (PRICES["P"] * w[3] + PRICES["E"] * w[4]) * (F_A + F_B) - PRICES["A"] * F_A - PRICES["B"] * F_B
        return self._ret(-profit, shift)

    def get_constraint1(self, u, noise=0.):
        """0.12 - x_A >= 0 (:64-75)."""
        w, shift = self.steady_state(u, noise)
        return self._ret(float(0.12 - w[0]), shift)

    def get_constraint2(self, u, noise=0.):
        """0.08 - x_G >= 0 (:77-90)."""
        w, shift = self.steady_state(u, noise)
        return self._ret(float(0.08 - w[5]), shift)
