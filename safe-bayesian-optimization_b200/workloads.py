"""Synthetic workloads of SURVEY.md section 8(d) / BASELINE.md: C4 (d=4, N=32^4=2^20, n=512, G=4) and
C5 (d=6, N=16^6=2^24, n=2048, G=4).  Deterministic (seeded) inputs; hyper-parameters are fixed inside
the reference's bounds (models/GP_Safe.py:205-206) so no fit is needed.

Retuned (allowed by SURVEY.md 8d, frozen here and recorded in BASELINE.md/DESIGN.md): with the survey's
training ball of radius 0.6 the safe set covers only 4 % (C4) / 0.1 % (C5) of the grid because the
reference's constraint prior mean m0 = -2*Ybar/Ystd (GP_Safe.py:331) makes everything away from the data
unsafe.  The training ball radius and the constraint radii below put the safe set at ~10 % of the grid.
"""
import numpy as np


def synthetic(d, pts_per_dim, n, seed, G=4, beta=2.0, r_train=1.1, radii=(0.9, 1.0, 1.1, 1.2, 1.3, 1.4, 1.5)):
    """Benoit generalised to d dims: f0 = sum x_k^2 + sum x_k x_{k+1} (reference Benoit_Problem.py:16),
    constraints g_j = r_j^2 - |x - a_j|^2, j = 1..G-1, a_j = +-0.2 e_{j-1}.  Training inputs uniform in the
    ball of radius r_train around x_c = (0.3, ...), clipped to the box [-1,1]^d.
    Returns (inference_datasets dict without invKopt, lo, hi, pts, beta)."""
    rng = np.random.default_rng(seed)
    xc = np.full(d, 0.3)
    xi = rng.normal(size=(n, d))
    xi /= np.linalg.norm(xi, axis=1, keepdims=True)
    X = np.clip(xc + r_train * rng.uniform(size=(n, 1)) ** (1.0 / d) * xi, -1.0, 1.0)
    Y = np.empty((n, G))
    Y[:, 0] = np.sum(X * X, axis=1) + np.sum(X[:, :-1] * X[:, 1:], axis=1)
    for j in range(1, G):
        a = np.zeros(d)
        a[(j - 1) % d] = 0.2 if j % 2 else -0.2
        Y[:, j] = radii[j - 1] ** 2 - np.sum((X - a) ** 2, axis=1)
    X_mean, X_std = X.mean(axis=0), X.std(axis=0)
    Y_mean, Y_std = Y.mean(axis=0), Y.std(axis=0)
    hyp = np.zeros((d + 2, G))
    hyp[d + 1, :] = -3.0                     # 1/2 log ell = 0, 1/2 log sf2 = 0, 1/2 log sn2 = -3
    ds = {"X_mean": X_mean, "X_std": X_std, "Y_mean": Y_mean, "Y_std": Y_std,
          "X_norm": (X - X_mean) / X_std, "Y_norm": (Y - Y_mean) / Y_std, "invKopt": None, "hypopt": hyp}
    lo, hi = -np.ones(d), np.ones(d)
    return ds, lo, hi, [int(pts_per_dim)] * d, beta


def c4(pts_per_dim=32, n=512):
    return synthetic(d=4, pts_per_dim=pts_per_dim, n=n, seed=1234, r_train=1.1, radii=(0.9, 1.0, 1.1))


def c5(pts_per_dim=16, n=2048):
    return synthetic(d=6, pts_per_dim=pts_per_dim, n=n, seed=5678, r_train=1.4, radii=(1.3, 1.4, 1.5))


def small(d=3, pts_per_dim=12, n=40, seed=7, G=3):
    """A seconds-scale case of the same recipe for parity tests."""
    return synthetic(d=d, pts_per_dim=pts_per_dim, n=n, seed=seed, G=G, r_train=0.9)
