"""CPU arm of bench.py  --  TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product package).

The reference's own arithmetic for the grid step, timed on the host cores on a BOUNDED sample and split into the
two cost classes that scale differently, so that each is extrapolated by its own ratio:

  per-point work   the reference's inverse-form posterior  k^T invK (Y - m0),  k^T invK k  (models/GP_Safe.py:310-352,
                   vmapped over the grid as test/test_SafeOpt.py:324-338 does), the bounds and sets
                   (SafeOpt.py:34-66) and -- fantasy mode only -- the whitened rows V = L^-1 k the pair stage consumes.
                   Cost is linear in the number of grid points  ->  scaled by N / N_sample.
  per-pair work    the expander pair test for a block of candidates x unsafe points, GEMM-shaped in FP64:
                   fantasy mode:   acc = V_z V_x^T (DGEMM, K = n), k(z,x) from the dot-form distance
                                   (GP_Safe.py:112-119: -2 A.B^T + |A|^2 + |B|^2), rank-1 update, threshold;
                   Lipschitz mode: dot-form distance in raw space, ucb - L*sqrt(.) >= 0 (SafeOpt.py:85-88).
                   Cost is linear in the number of (x, z, constraint) triples  ->  scaled by pairs / pairs_sample.

Dense kernels are OpenBLAS DGEMMs through NumPy; the element-wise epilogues run in oracle/pair_epilogue.c (fused,
OpenMP, all host threads; built by __graft_entry__.build()) because NumPy's are single-threaded multi-pass -- with it
the pair test is DGEMM-bound.  If that library is not built the same epilogue runs on torch CPU tensors.  Everything
is FP64.
"""
from __future__ import annotations

import time

import numpy as np

from . import gp_oracle as O


import ctypes as C
import os

_EPI = None


def _epilogue_lib():
    """oracle/_build/libpair_epilogue.so (gcc -fopenmp; oracle/Makefile), or None when it has not been built."""
    global _EPI
    if _EPI is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libpair_epilogue.so")
        _EPI = False
        if os.path.exists(path):
            try:
                lib = C.CDLL(path)
                D, U8 = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
                lib.fantasy_epilogue.argtypes = [C.c_int64, C.c_int64, D, D, D, D, D, D, C.c_double, C.c_double, U8]
                lib.lipschitz_epilogue.argtypes = [C.c_int64, C.c_int64, D, D, C.c_double, U8]
                lib.fantasy_epilogue.restype = lib.lipschitz_epilogue.restype = None
                lib.epilogue_set_threads.argtypes = [C.c_int]
                lib.epilogue_set_threads(os.cpu_count() or 1)        # every host core (torchrun sets OMP_NUM_THREADS=1)
                _EPI = lib
            except OSError:
                _EPI = False
    return _EPI or None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _torch():
    import torch
    return torch


def model_state(ds):
    """What the reference rebuilds on every add_sample (GP_Safe.py:226-232): K, inv(K) per GP; plus the Cholesky factor
    the fantasy rows need.  Returns (ds with invKopt, [L^-1 per GP], seconds)."""
    t0 = time.perf_counter()
    G = ds["Y_norm"].shape[1]
    d = ds["X_norm"].shape[1]
    out = dict(ds)
    Ks = [O.build_K(ds["X_norm"], ds["hypopt"][:, i]) for i in range(G)]
    out["invKopt"] = [np.linalg.inv(K) for K in Ks]                                  # GP_Safe.py:232
    Ws = [np.linalg.inv(np.linalg.cholesky(K)) for K in Ks]                          # L^-1 (fantasy rows only)
    return out, Ws, time.perf_counter() - t0


def per_point(points, dso, Ws, beta, fantasy):
    """Inverse-form posterior + bounds + sets (+ V rows) at `points`.  Returns dict with arrays and seconds/flops."""
    N, d = points.shape
    G = dso["Y_norm"].shape[1]
    n = dso["X_norm"].shape[0]
    t0 = time.perf_counter()
    mean, var = O.posterior_inv(points, dso)                                         # GP_Safe.py:310-352
    lcb, ucb = O.bounds(mean, var, beta)                                             # SafeOpt.py:34-45
    S, Z = O.safe_mask(lcb), O.unsafe_mask(lcb)                                      # SafeOpt.py:58-59,73-77,109
    if S.any():
        O.minimizer(var, lcb, ucb, S)                                                # SafeOpt.py:53-66
    t_post = time.perf_counter() - t0
    flops = G * N * (2.0 * n * n + n * (3 * d + 6))                                  # reference form: 2 n^2 per point
    V = None
    t_v = 0.0
    if fantasy:
        t0 = time.perf_counter()
        xn = (points - dso["X_mean"]) / dso["X_std"]
        V = []
        for i in range(1, G):
            ell, sf2, _ = O.unpack_hyper(dso["hypopt"][:, i], d)
            k = O.cov_mat(dso["X_norm"], xn, ell, sf2)                               # (n, N) dot-form, GP_Safe.py:146-167
            V.append(np.ascontiguousarray((Ws[i] @ k).T))                            # (N, n)
        t_v = time.perf_counter() - t0
        flops += (G - 1) * N * (2.0 * n * n + n * (3 * d + 2))
    return {"mean": mean, "var": var, "lcb": lcb, "ucb": ucb, "S": S, "Z": Z, "V": V, "seconds": t_post + t_v,
            "flops": flops}


def pairs_fantasy(points, dso, beta, pp, xs, zs, block=2048):
    """Fantasy pair test for candidates xs x unsafe zs (index arrays into `points`), all constraints, FP64,
    GEMM-shaped blocks.  Returns (counts per x, seconds, flops)."""
    torch = _torch()
    G = dso["Y_norm"].shape[1]
    d = points.shape[1]
    n = dso["X_norm"].shape[0]
    xn = (points - dso["X_mean"]) / dso["X_std"]
    mu_n = pp["mean"] / dso["Y_std"]
    var_n = pp["var"] / dso["Y_std"] ** 2
    counts = np.zeros(xs.size, dtype=np.int64)
    epi = _epilogue_lib()
    t0 = time.perf_counter()
    for xb0 in range(0, xs.size, block):
        xb = xs[xb0:xb0 + block]
        for zb0 in range(0, zs.size, block):
            zb = zs[zb0:zb0 + block]
            ok = np.ones((zb.size, xb.size), dtype=np.uint8) if epi else torch.ones((zb.size, xb.size), dtype=torch.bool)
            for i in range(1, G):
                ell, sf2, sn2 = O.unpack_hyper(dso["hypopt"][:, i], d)
                sn2 = sn2 + O.EPS_F32
                A = xn[zb] / np.sqrt(ell)
                B = xn[xb] / np.sqrt(ell)
                dist = (A * A).sum(1)[:, None] + (B * B).sum(1)[None, :] - 2.0 * (A @ B.T)     # GP_Safe.py:119
                acc = pp["V"][i - 1][zb] @ pp["V"][i - 1][xb].T                                # DGEMM, K = n
                den = var_n[xb, i] + sn2
                if epi:
                    ax = np.ascontiguousarray(beta * np.sqrt(var_n[xb, i]) / den)
                    bx = np.ascontiguousarray(1.0 / den)
                    mz, sz = np.ascontiguousarray(mu_n[zb, i]), np.ascontiguousarray(var_n[zb, i])
                    epi.fantasy_epilogue(zb.size, xb.size, _dp(dist), _dp(acc), _dp(mz), _dp(sz), _dp(ax), _dp(bx),
                                         float(sf2), float(beta), ok.ctypes.data_as(C.POINTER(C.c_uint8)))
                    continue
                c = torch.from_numpy(dist).mul_(-0.5).exp_().mul_(sf2).sub_(torch.from_numpy(acc))
                a = torch.from_numpy(beta * np.sqrt(var_n[xb, i]) / den)[None, :]
                b = torch.from_numpy(1.0 / den)[None, :]
                mu = torch.from_numpy(mu_n[zb, i])[:, None] + c * a
                s2 = (torch.from_numpy(var_n[zb, i])[:, None] - c.mul_(c).mul_(b)).clamp_(min=0.0)
                ok &= mu.sub_(s2.sqrt_().mul_(beta)) >= 0.0
            counts[xb0:xb0 + xb.size] += ok.sum(axis=0, dtype=np.int64) if epi else ok.sum(dim=0).numpy()
    secs = time.perf_counter() - t0
    pairs = xs.size * zs.size * (G - 1)
    return counts, secs, pairs * (2.0 * n + 3 * d + 20)


def pairs_lipschitz(points, pp, L, xs, zs, G, block=4096):
    """Lipschitz pair test  ucb_idx(x) - L*||x - z + 1e-8|| >= 0  (SafeOpt.py:85-88) for xs x zs, every constraint,
    dot-form distance (GEMM-shaped, K = d).  Returns (hit flags (G-1, |xs|), seconds, flops)."""
    torch = _torch()
    d = points.shape[1]
    hit = np.zeros((G - 1, xs.size), dtype=np.uint8)
    epi = _epilogue_lib()
    t0 = time.perf_counter()
    for xb0 in range(0, xs.size, block):
        xb = xs[xb0:xb0 + block]
        X = points[xb] + O.PAIR_OFFSET
        for zb0 in range(0, zs.size, block):
            Zp = points[zs[zb0:zb0 + block]]
            d2 = np.ascontiguousarray((X * X).sum(1)[:, None] + (Zp * Zp).sum(1)[None, :] - 2.0 * (X @ Zp.T))
            if epi:
                for i in range(1, G):
                    h = np.ascontiguousarray(hit[i - 1, xb0:xb0 + xb.size])
                    u = np.ascontiguousarray(pp["ucb"][xb, i])
                    epi.lipschitz_epilogue(xb.size, Zp.shape[0], _dp(d2), _dp(u), float(L[i]), h.ctypes.data_as(C.POINTER(C.c_uint8)))
                    hit[i - 1, xb0:xb0 + xb.size] = h
                continue
            dist = torch.from_numpy(d2).clamp_(min=0.0).sqrt_()
            for i in range(1, G):
                r = torch.from_numpy(pp["ucb"][xb, i] / L[i])[:, None]
                hit[i - 1, xb0:xb0 + xb.size] |= (dist <= r).any(dim=1).numpy().astype(np.uint8)
    secs = time.perf_counter() - t0
    pairs = xs.size * zs.size * (G - 1)
    return hit.astype(bool), secs, pairs * (3.0 * d + 6)


def sample_step(ds, lo, hi, pts, beta, mode, n_points, n_x, n_z, seed=0):
    """One bounded CPU step.  `n_points` random grid points (seeded, without replacement; per-point work does not
    depend on which points), then the pair test for up to n_x safe x n_z unsafe of them."""
    G = ds["Y_norm"].shape[1]
    d = len(pts)
    N = int(np.prod(pts))
    rng = np.random.default_rng(seed)
    dso, Ws, t_model = model_state(ds)
    idx = np.sort(rng.choice(N, size=min(n_points, N), replace=False))
    axes = O.grid_axes(lo, hi, pts)
    sub = np.unravel_index(idx, tuple(int(p) for p in pts[::-1]))          # slowest axis first
    P = np.column_stack([axes[k][sub[d - 1 - k]] for k in range(d)])
    fantasy = mode == "fantasy"
    pp = per_point(P, dso, Ws, beta, fantasy)
    xs, zs = np.flatnonzero(pp["S"])[:n_x], np.flatnonzero(pp["Z"])[:n_z]
    t_pairs, f_pairs, pairs = 0.0, 0.0, 0
    if xs.size and zs.size:
        if fantasy:
            _, t_pairs, f_pairs = pairs_fantasy(P, dso, beta, pp, xs, zs)
        else:
            Lg = [0.0] + [max(O.lipschitz_constant(P[: min(4096, P.shape[0])], dso, G - 1), 1e-12)] * (G - 1)
            _, t_pairs, f_pairs = pairs_lipschitz(P, pp, Lg, xs, zs, G)
        pairs = int(xs.size) * int(zs.size) * (G - 1)
    return {"t_model": t_model, "t_points": pp["seconds"], "n_points": int(idx.size), "point_flops": pp["flops"],
            "t_pairs": t_pairs, "pairs": pairs, "pair_flops": f_pairs,
            "safe_frac": float(pp["S"].mean()), "unsafe_frac": float(pp["Z"].mean())}


def extrapolate(s, N, pairs_full):
    """Full-step seconds: model once + per-point work scaled by the POINT ratio + per-pair work by the PAIR ratio."""
    t = s["t_model"] + s["t_points"] * (N / s["n_points"])
    if s["pairs"]:
        t += s["t_pairs"] * (pairs_full / s["pairs"])
    return t


def describe(s, N, pairs_full, mode):
    gp = s["point_flops"] / s["t_points"] / 1e9 if s["t_points"] > 0 else 0.0
    gq = s["pair_flops"] / s["t_pairs"] / 1e9 if s["t_pairs"] > 0 else 0.0
    return (f"{s['n_points']} random grid points: inverse-form posterior + sets{' + V rows' if mode == 'fantasy' else ''} "
            f"({gp:.0f} GFLOP/s FP64), scaled by the point ratio {N / s['n_points']:.1f}; {s['pairs']:.3g} pair-evals "
            f"({mode}, FP64 GEMM-shaped blocks, {gq:.0f} GFLOP/s, {s['pairs'] / max(s['t_pairs'], 1e-9):.3g} pair-evals/s), "
            f"scaled by the pair ratio {pairs_full / max(s['pairs'], 1):.3g}; model state {s['t_model'] * 1e3:.0f} ms once")
