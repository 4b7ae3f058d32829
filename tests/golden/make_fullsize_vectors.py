"""Full-size parity fixtures for the fantasy expander ("oracle-check the benchmarked configs at full size").

Run once in the build container (minutes of CPU, ~12 GB of RAM); the outputs are small and committed:

  c4_full_sampled.npz   C4 (d=4, 32^4 grid, n=512, G=4, the bench workload): for 160 candidates x in S -- evenly spaced
                        over S in grid order, the first/last 8 (ragged ends) and the 16 smallest / largest variances (the
                        two ends of the key order the GPU sorts by) -- the EXACT FP64 newly-safe count g(x) against ALL
                        757 532 unsafe z, plus, per candidate, how many pairs lie within 1e-4 and within the single-pass
                        TF32 band of the threshold (the tolerance the test may use).
  c5_points_sampled.npz C5 model (d=6, n=2048, G=4) on 49 152 points drawn (seeded) from its 16^6 grid and handed to the
                        library as EXPLICIT points: FP64 counts for every safe point of the sample.  K = 2048 (64 K blocks),
                        12-float records, ~60 work items per cluster.

Oracle functions used: gp_oracle.posterior_chol / chol_factors / bounds / safe_mask / unsafe_mask and the arithmetic of
gp_oracle.fantasy_counts restated blockwise (same operations; V rows only for the points that take part).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import sbo_b200  # noqa: E402,F401
from sbo_b200 import workloads  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402


def v_rows(points, ds, i, L, block=32768):
    """V = L^-1 k(X, points) as rows (N, n) for GP i (FP64, triangular solve like gp_oracle.posterior_chol)."""
    import scipy.linalg as sla
    d = points.shape[1]
    ell, sf2, _ = O.unpack_hyper(ds["hypopt"][:, i], d)
    out = np.empty((points.shape[0], ds["X_norm"].shape[0]))
    for s in range(0, points.shape[0], block):
        xn = (points[s:s + block] - ds["X_mean"]) / ds["X_std"]
        k = sf2 * np.exp(-0.5 * O.sq_dist_direct(ds["X_norm"], xn, ell))
        out[s:s + block] = sla.solve_triangular(L, k, lower=True).T
    return out


def exact_counts(P, ds, beta, mean, var, xs, zs, tf32_tol=1e-3):
    """g(x) for the candidates xs against the unsafe points zs + near-threshold pair counts (normalised margins)."""
    G = ds["Y_norm"].shape[1]
    d = P.shape[1]
    fac = O.chol_factors(ds)
    mu_n, var_n = mean / ds["Y_std"], var / ds["Y_std"] ** 2
    xn = (P - ds["X_mean"]) / ds["X_std"]
    ok = np.ones((zs.size, xs.size), dtype=bool)
    marg = np.full((zs.size, xs.size), np.inf)
    gain = np.zeros(xs.size)
    sf2max = 0.0
    for i in range(1, G):
        ell, sf2, sn2 = O.unpack_hyper(ds["hypopt"][:, i], d)
        sn2 = sn2 + O.EPS_F32
        sf2max = max(sf2max, sf2)
        Vz = v_rows(P[zs], ds, i, fac[i][0])
        Vx = v_rows(P[xs], ds, i, fac[i][0])
        for s in range(0, zs.size, 65536):
            zb = zs[s:s + 65536]
            kzx = sf2 * np.exp(-0.5 * O.sq_dist_direct(xn[zb], xn[xs], ell))
            c = kzx - Vz[s:s + 65536] @ Vx.T
            den = var_n[xs, i] + sn2
            mu_p = mu_n[zb, i][:, None] + c * (beta * np.sqrt(var_n[xs, i]) / den)[None, :]
            s2_p = var_n[zb, i][:, None] - c * c / den[None, :]
            m = mu_p - beta * np.sqrt(np.maximum(s2_p, 0.0))
            ok[s:s + 65536] &= m >= 0.0
            marg[s:s + 65536] = np.minimum(marg[s:s + 65536], m)
        gain = np.maximum(gain, beta * np.sqrt(var_n[xs, i]) / (var_n[xs, i] + sn2))
        del Vz, Vx
    counts = ok.sum(axis=0).astype(np.int64)
    near4 = (np.abs(marg) <= 1e-4 * sf2max).sum(axis=0).astype(np.int64)
    near_tf32 = (np.abs(marg) <= tf32_tol * sf2max * (1.0 + gain)[None, :]).sum(axis=0).astype(np.int64)
    return counts, near4, near_tf32, gain


def make_c4():
    ds, lo, hi, pts, beta = workloads.c4()
    P = O.make_grid(lo, hi, pts)
    t0 = time.time()
    mean, var = O.posterior_chol(P, ds)
    lcb, _ = O.bounds(mean, var, beta)
    S, Z = O.safe_mask(lcb), O.unsafe_mask(lcb)
    xs_all, zs = np.flatnonzero(S), np.flatnonzero(Z)
    order = np.argsort(var[xs_all, 1], kind="stable")
    pick = set(xs_all[np.linspace(0, xs_all.size - 1, 112).astype(int)].tolist())
    pick |= set(xs_all[:8].tolist()) | set(xs_all[-8:].tolist())
    pick |= set(xs_all[order[:16]].tolist()) | set(xs_all[order[-16:]].tolist())
    xs = np.array(sorted(pick), dtype=np.int64)
    counts, near4, near_tf32, gain = exact_counts(P, ds, beta, mean, var, xs, zs)
    np.savez_compressed(os.path.join(HERE, "c4_full_sampled.npz"), x_idx=xs, counts=counts, near_1e4=near4,
                        near_tf32=near_tf32, gain=gain, n_safe=int(S.sum()), n_unsafe=int(Z.sum()), beta=beta)
    print("c4:", xs.size, "candidates,", int(S.sum()), "safe,", int(Z.sum()), "unsafe,", round(time.time() - t0), "s;",
          "counts", counts.min(), counts.max(), "near_1e4 max", near4.max(), "near_tf32 max", near_tf32.max())


def c5_sample_points(n_points=49152, seed=2024):
    ds, lo, hi, pts, beta = workloads.c5()
    rng = np.random.default_rng(seed)
    N = int(np.prod(pts))
    idx = np.sort(rng.choice(N, size=n_points, replace=False))
    axes = O.grid_axes(lo, hi, pts)
    d = len(pts)
    sub = np.unravel_index(idx, tuple(int(p) for p in pts[::-1]))
    P = np.column_stack([axes[k][sub[d - 1 - k]] for k in range(d)])
    return ds, beta, idx, P


def make_c5():
    ds, beta, idx, P = c5_sample_points()
    t0 = time.time()
    mean, var = O.posterior_chol(P, ds)
    lcb, _ = O.bounds(mean, var, beta)
    S, Z = O.safe_mask(lcb), O.unsafe_mask(lcb)
    xs, zs = np.flatnonzero(S), np.flatnonzero(Z)
    counts, near4, near_tf32, gain = exact_counts(P, ds, beta, mean, var, xs, zs)
    np.savez_compressed(os.path.join(HERE, "c5_points_sampled.npz"), grid_idx=idx, x_local=xs, counts=counts,
                        near_1e4=near4, near_tf32=near_tf32, gain=gain, n_safe=int(S.sum()), n_unsafe=int(Z.sum()),
                        beta=beta)
    print("c5 sample:", P.shape[0], "points,", xs.size, "safe,", zs.size, "unsafe,", round(time.time() - t0), "s;",
          "counts", counts.min(), counts.max(), "near_1e4 max", near4.max())


if __name__ == "__main__":
    which = sys.argv[1:] or ["c4", "c5"]
    if "c5" in which:
        make_c5()
    if "c4" in which:
        make_c4()
