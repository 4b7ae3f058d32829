"""CPU oracle for the SafeOpt/GoOSE grid hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy FP64 restatement of the arithmetic of the reference
(dleeim/Safe-Bayesian-Optimization).  It is the *checker* for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under
``safe-bayesian-optimization_b200/`` imports it; the product path has no CPU
fallback.

PARITY PINNED against outputs of the reference's own source (with one stated gap).  JAX is not installed here, so
``tests/golden/make_reference_vectors.py`` imports the UNMODIFIED ``/root/reference/models/{GP_Safe,SafeOpt,
GoOSE}.py`` over a NumPy-backed ``jax`` stand-in (``tests/golden/refshim/``) and records what the reference computes:
its DE hyper-fit, ``invKopt``, normalisation, single-point ``GP_inference``/``lcb``/``ucb``/``infnorm_mean_grad``,
the 400x400 plot mask of ``create_data_for_plot`` and the DE-based ``Minimizer/Expander/Target/explore_safeset``
(committed as ``tests/golden/ref_*.npz``).  ``tests/test_reference_vectors.py`` checks this restatement against them:
model state bit-for-bit (normalisation, K, inv(K)), posterior to 1e-9 scale-relative, plot mask identical, DE
optima by containment up to the grid resolution.  The gap: the arithmetic backend of those vectors is NumPy/LAPACK
FP64, not XLA-CPU, so XLA's rounding is not pinned.  The fantasy expander (``fantasy_*``) is not in the reference
at all (north_star addition): it stays UNPINNED and is validated by re-running the pinned inference on the
augmented (n+1)-point data set.  Every function cites the reference lines it follows; internal consistency checks
(inverse-form vs Cholesky-form posterior, analytic gradient vs central differences, the properties the reference's
scripts print) are in ``tests/test_oracle.py``.

Conventions
-----------
* ``ds`` is the reference's ``inference_datasets`` dict (GP_Safe.py:16-23,236-245):
  X_mean,X_std (d,), Y_mean,Y_std (G,), X_norm (n,d), Y_norm (n,G),
  invKopt list of G (n,n), hypopt (d+2,G) with rows [0:d]=1/2 log ell,
  [d]=1/2 log sf2, [d+1]=1/2 log sn2.
* grids are meshgrids flattened with x_0 the fastest axis
  (test/test_SafeOpt.py:324-334: meshgrid 'xy' + ravel => p = r*400 + c).
* every arg-reduction breaks ties towards the lowest grid index.
"""
from __future__ import annotations

import numpy as np

EPS_F32 = float(np.finfo(np.float32).eps)  # GP_Safe.py:229 jnp.finfo(jnp.float32).eps
PAIR_OFFSET = 1e-8                         # SafeOpt.py:87 / GoOSE.py:71  "+1e-8"


# ----------------------------------------------------------------------------
# grid  (test/test_SafeOpt.py:324-334, test/test_GoOSE.py:192-202)
# ----------------------------------------------------------------------------
def grid_axes(lo, hi, pts):
    """Per-axis linspace, exactly numpy's: lo + i*step, last point = hi."""
    return [np.linspace(float(l), float(h), int(m)) for l, h, m in zip(lo, hi, pts)]


def make_grid(lo, hi, pts):
    """(N,d) points of the meshgrid, x_0 fastest (index p = sum_k i_k * prod_{j<k} pts_j)."""
    axes = grid_axes(lo, hi, pts)
    d = len(axes)
    mesh = np.meshgrid(*axes[::-1], indexing="ij")  # slowest axis first
    cols = [mesh[d - 1 - k].ravel() for k in range(d)]
    return np.column_stack(cols)


# ----------------------------------------------------------------------------
# kernel / model  (models/GP_Safe.py)
# ----------------------------------------------------------------------------
def squared_seuclidean(X, Y, V):
    """GP_Safe.py:98-120 -- dot form  -2 A.B^T + |A|^2 + |B|^2  with A = X * V**-0.5."""
    V_sqrt_inv = V ** -0.5
    Xa = X * V_sqrt_inv
    Ya = Y * V_sqrt_inv
    return -2.0 * np.dot(Xa, Ya.T) + np.sum(Xa ** 2, axis=1)[:, None] + np.sum(Ya ** 2, axis=1)


def cov_mat(X_norm, Y_norm, ell, sf2):
    """GP_Safe.py:122-143 / 146-167 -- sf2 * exp(-0.5 * dist), RBF only."""
    if ell.shape[0] != X_norm.shape[1]:
        raise ValueError("ERROR W and X_norm dimension should be same")
    return sf2 * np.exp(-0.5 * squared_seuclidean(X_norm, Y_norm, ell))


def unpack_hyper(hyp_col, d):
    """GP_Safe.py:338 -- ell, sf2, sn2 = exp(2*hyper[:d]), exp(2*hyper[d]), exp(2*hyper[d+1])."""
    return np.exp(2.0 * hyp_col[:d]), float(np.exp(2.0 * hyp_col[d])), float(np.exp(2.0 * hyp_col[d + 1]))


def build_K(X_norm, hyp_col):
    """GP_Safe.py:226-231 -- Kopt = Cov + (sn2 + eps_f32) * I."""
    n, d = X_norm.shape
    ell, sf2, sn2 = unpack_hyper(hyp_col, d)
    return cov_mat(X_norm, X_norm, ell, sf2) + (sn2 + EPS_F32) * np.eye(n)


def normalize(X, Y):
    """GP_Safe.py:84-96 -- z-score with population std (ddof=0)."""
    X_mean, X_std = np.mean(X, axis=0), np.std(X, axis=0)
    Y_mean, Y_std = np.mean(Y, axis=0), np.std(Y, axis=0)
    return X_mean, X_std, Y_mean, Y_std, (X - X_mean) / X_std, (Y - Y_mean) / Y_std


def make_inference_datasets(X, Y, hypopt, with_inverse=True):
    """GP_Safe.py:236-245 -- the model state consumed by GP_inference."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    hypopt = np.asarray(hypopt, dtype=np.float64)
    X_mean, X_std, Y_mean, Y_std, X_norm, Y_norm = normalize(X, Y)
    invK = []
    if with_inverse:
        for i in range(Y.shape[1]):
            invK.append(np.linalg.inv(build_K(X_norm, hypopt[:, i])))  # GP_Safe.py:232
    return {"X_mean": X_mean, "X_std": X_std, "Y_mean": Y_mean, "Y_std": Y_std,
            "X_norm": X_norm, "Y_norm": Y_norm, "invKopt": invK, "hypopt": hypopt}


def prior_mean(ds):
    """GP_Safe.py:331-332 -- m0 = -2*Y_mean/Y_std, objective prior = 0."""
    m0 = (-2.0 * ds["Y_mean"]) / ds["Y_std"]
    m0 = m0.copy()
    m0[0] = 0.0
    return m0


def negative_loglikelihood(hyper, X, Y):
    """GP_Safe.py:169-192 (Y is one (n,1) column)."""
    n, d = X.shape
    W = np.exp(2.0 * hyper[:d])
    sf2 = np.exp(2.0 * hyper[d])
    sn2 = np.exp(2.0 * hyper[d + 1])
    K = cov_mat(X, X, W, sf2) + (sn2 + 1e-8) * np.eye(n)
    K = (K + K.T) * 0.5
    L = np.linalg.cholesky(K)
    logdetK = 2.0 * np.sum(np.log(np.diag(L)))
    import scipy.linalg as sla
    invLY = sla.solve_triangular(L, Y, lower=True)
    alpha = sla.solve_triangular(L.T, invLY, lower=False)
    return float(np.dot(Y.T, alpha)[0][0] + logdetK)


def fit_hyper(X_norm, Y_norm, seed=0, maxiter=1000):
    """GP_Safe.py:194-234 -- per-output differential evolution over
    [-1.5,1.5]^(d+1) x [-5,-2]; *seeded* here (the reference is unseeded)."""
    from scipy.optimize import differential_evolution
    n, d = X_norm.shape
    G = Y_norm.shape[1]
    bounds = [(-1.5, 1.5)] * (d + 1) + [(-5.0, -2.0)]
    hyp = np.zeros((d + 2, G))
    for i in range(G):
        def nll(h):
            try:
                return negative_loglikelihood(h, X_norm, Y_norm[:, i:i + 1])
            except np.linalg.LinAlgError:
                return 1e30
        res = differential_evolution(nll, bounds=bounds, seed=seed + i, maxiter=maxiter)
        hyp[:, i] = res.x
    return hyp


# ----------------------------------------------------------------------------
# posterior  (GP_Safe.py:310-352), batched over points
# ----------------------------------------------------------------------------
def posterior_inv(points, ds, block=65536):
    """Form (a), reference-exact: dot-form distance, explicit inv(K),
    mean = m0 + k^T invK (Y-m0), var = max(0, sf2 - k^T invK k).  Returns (N,G),(N,G)."""
    points = np.atleast_2d(np.asarray(points, dtype=np.float64))
    N, d = points.shape
    G = ds["Y_norm"].shape[1]
    m0 = prior_mean(ds)
    mean = np.empty((N, G))
    var = np.empty((N, G))
    for s in range(0, N, block):
        xn = (points[s:s + block] - ds["X_mean"]) / ds["X_std"]          # :326
        for i in range(G):
            ell, sf2, _ = unpack_hyper(ds["hypopt"][:, i], d)            # :338
            k = cov_mat(ds["X_norm"], xn, ell, sf2)                       # (n,B) :341
            kiK = k.T @ ds["invKopt"][i]                                  # (B,n)
            mu = m0[i] + kiK @ (ds["Y_norm"][:, i] - m0[i])               # :342
            s2 = np.maximum(0.0, sf2 - np.einsum("bn,nb->b", kiK, k))     # :343
            mean[s:s + block, i] = mu * ds["Y_std"][i] + ds["Y_mean"][i]  # :346
            var[s:s + block, i] = s2 * ds["Y_std"][i] ** 2                # :347
    return mean, var


def chol_factors(ds):
    """Per-GP (L, alpha): K = L L^T, alpha = K^-1 (Y - m0)."""
    import scipy.linalg as sla
    G = ds["Y_norm"].shape[1]
    m0 = prior_mean(ds)
    out = []
    for i in range(G):
        K = build_K(ds["X_norm"], ds["hypopt"][:, i])
        L = np.linalg.cholesky(K)
        r = ds["Y_norm"][:, i] - m0[i]
        alpha = sla.solve_triangular(L.T, sla.solve_triangular(L, r, lower=True), lower=False)
        out.append((L, alpha))
    return out


def sq_dist_direct(A, B, ell):
    """sum_k (A_jk - B_pk)^2 / ell_k  without the dot-form cancellation. (n,B)."""
    diff = A[:, None, :] - B[None, :, :]
    return np.einsum("jpk,k->jp", diff * diff, 1.0 / ell)


def posterior_chol(points, ds, block=16384, return_V=False, factors=None):
    """Form (b): direct distance, triangular solve v = L^-1 k,
    mean = m0 + k.alpha, var = max(0, sf2 - |v|^2).  Same un-normalisation."""
    import scipy.linalg as sla
    points = np.atleast_2d(np.asarray(points, dtype=np.float64))
    N, d = points.shape
    G = ds["Y_norm"].shape[1]
    m0 = prior_mean(ds)
    fac = factors if factors is not None else chol_factors(ds)
    mean = np.empty((N, G))
    var = np.empty((N, G))
    Vs = [np.empty((N, ds["X_norm"].shape[0])) for _ in range(G)] if return_V else None
    for s in range(0, N, block):
        xn = (points[s:s + block] - ds["X_mean"]) / ds["X_std"]
        for i in range(G):
            ell, sf2, _ = unpack_hyper(ds["hypopt"][:, i], d)
            L, alpha = fac[i]
            k = sf2 * np.exp(-0.5 * sq_dist_direct(ds["X_norm"], xn, ell))   # (n,B)
            v = sla.solve_triangular(L, k, lower=True)
            mu = m0[i] + k.T @ alpha
            s2 = np.maximum(0.0, sf2 - np.sum(v * v, axis=0))
            mean[s:s + block, i] = mu * ds["Y_std"][i] + ds["Y_mean"][i]
            var[s:s + block, i] = s2 * ds["Y_std"][i] ** 2
            if return_V:
                Vs[i][s:s + block] = v.T
    if return_V:
        return mean, var, Vs
    return mean, var


def gp_inference(x, ds):
    """Single point, reference signature: returns (mean (G,), var (G,)). GP_Safe.py:310-350."""
    m, v = posterior_inv(np.asarray(x, dtype=np.float64).reshape(1, -1), ds)
    return m[0], v[0]


# ----------------------------------------------------------------------------
# bounds and sets  (models/SafeOpt.py:29-66, models/GoOSE.py:22-31,40-67)
# ----------------------------------------------------------------------------
def bounds(mean, var, beta):
    """SafeOpt.py:34-45 -- ucb = mean + b*sqrt(var), lcb = mean - b*sqrt(var).  (N,G) each."""
    s = beta * np.sqrt(var)
    return mean - s, mean + s


def safe_mask(lcb, strict=False):
    """SafeOpt.py:58-59 / GoOSE.py:22-25: S = {lcb_i >= 0 for all constraints i>=1}.
    strict=True is the plot-mask variant ``lcb > 0.`` (test_SafeOpt.py:337-338)."""
    c = lcb[:, 1:]
    return np.all(c > 0.0, axis=1) if strict else np.all(c >= 0.0, axis=1)


def unsafe_mask(lcb, rule="all"):
    """SafeOpt.py:73-77,109: ``lcb_constraint_min`` returns the MAX of the
    constraint lcbs and is constrained <= 0, so the reference's Z is
    {z : lcb_i(z) <= 0 for ALL constraints}  (rule='all', default = reference).
    rule='any' is the complement-style set {z : some lcb_i(z) < 0} = not S."""
    c = lcb[:, 1:]
    if rule == "all":
        return np.max(c, axis=1) <= 0.0
    if rule == "any":
        return np.any(c < 0.0, axis=1)
    raise ValueError(rule)


def masked_argmin(values, mask):
    """Lowest-index arg-min over mask; (-1, +inf) when the mask is empty."""
    if not np.any(mask):
        return -1, np.inf
    v = np.where(mask, values, np.inf)
    i = int(np.argmin(v))
    return i, float(v[i])


def masked_argmax(values, mask):
    """Lowest-index arg-max over mask; (-1, -inf) when the mask is empty."""
    if not np.any(mask):
        return -1, -np.inf
    v = np.where(mask, values, -np.inf)
    i = int(np.argmax(v))
    return i, float(v[i])


def minimize_obj_ucb(ucb, S):
    """SafeOpt.py:47-51 -- min over S of ucb_0."""
    return masked_argmin(ucb[:, 0], S)


def minimize_obj_lcb(lcb, S):
    """GoOSE.py:63-67 -- min over S of lcb_0."""
    return masked_argmin(lcb[:, 0], S)


def minimizer_set(lcb, ucb, S):
    """SafeOpt.py:53-62 -- M = {x in S : lcb_0(x) <= min_S ucb_0}."""
    _, min_ucb = minimize_obj_ucb(ucb, S)
    return S & (lcb[:, 0] <= min_ucb), min_ucb


def minimizer(var, lcb, ucb, S):
    """SafeOpt.py:53-66 -- argmax var_0 over M; returns (idx, std, M, min_ucb)."""
    M, min_ucb = minimizer_set(lcb, ucb, S)
    idx, v = masked_argmax(var[:, 0], M)
    return idx, (float(np.sqrt(v)) if idx >= 0 else 0.0), M, min_ucb


# ----------------------------------------------------------------------------
# Lipschitz constant  (SafeOpt.py:68-83, GoOSE.py:58-61,74-78)
# ----------------------------------------------------------------------------
def mean_grad(points, ds, i, block=32768):
    """Analytic d mu_i / d x  (autodiff of GP_inference in the reference):
    dmu/dx_k = (Ystd_i/Xstd_k) * sum_j a_j k_j * (-(xn_k - Xn_jk)/ell_k),  a = invK (Y - m0)."""
    points = np.atleast_2d(np.asarray(points, dtype=np.float64))
    N, d = points.shape
    m0 = prior_mean(ds)
    ell, sf2, _ = unpack_hyper(ds["hypopt"][:, i], d)
    a = ds["invKopt"][i] @ (ds["Y_norm"][:, i] - m0[i])
    out = np.empty((N, d))
    Xn = ds["X_norm"]
    for s in range(0, N, block):
        xn = (points[s:s + block] - ds["X_mean"]) / ds["X_std"]
        k = sf2 * np.exp(-0.5 * sq_dist_direct(Xn, xn, ell))            # (n,B)
        w = k * a[:, None]
        for kk in range(d):
            diff = xn[None, :, kk] - Xn[:, None, kk]                      # (n,B)
            out[s:s + block, kk] = -(w * diff).sum(axis=0) / ell[kk] * ds["Y_std"][i] / ds["X_std"][kk]
    return out


def lipschitz_constant(points, ds, i):
    """SafeOpt.py:79-83 restated on the grid: L_i = max_p || grad mu_i(p) ||_inf."""
    g = mean_grad(points, ds, i)
    return float(np.max(np.abs(g)))


# ----------------------------------------------------------------------------
# pair tests  (SafeOpt.py:85-124, GoOSE.py:69-119)
# ----------------------------------------------------------------------------
def pair_reach(xs, ucb_x, zs, L, block=2048):
    """Boolean (|xs|,|zs|) generator in blocks:  ucb(x) - L*||x - z + 1e-8||_2 >= 0
    (SafeOpt.py:85-88; raw x-space, offset added per component before the norm)."""
    for s in range(0, xs.shape[0], block):
        xb = xs[s:s + block]
        diff = xb[:, None, :] - zs[None, :, :] + PAIR_OFFSET
        dist = np.sqrt(np.sum(diff * diff, axis=2))
        yield s, (ucb_x[s:s + block, None] - L * dist) >= 0.0


def expander_lipschitz(points, S, Z, ucb, var, L_per_idx, block=1024):
    """SafeOpt.py:90-124 restated on the grid, all-pairs brute force.
    For each constraint idx: G_idx = {x in S : exists z in Z reachable}; pick
    argmax var_0 over G_idx; across idx keep the first largest std.
    L_per_idx[idx] is the Lipschitz constant used for constraint idx (the
    reference uses L_{G-1} for every idx, SafeOpt.py:110).
    Returns dict(best_idx, best_std, per_idx=[(idx_point, std)], masks (G-1,N) bool)."""
    N, G = ucb.shape
    xs_idx = np.flatnonzero(S)
    zs = points[Z]
    masks = np.zeros((G - 1, N), dtype=bool)
    per = []
    for idx in range(1, G):
        hit = np.zeros(xs_idx.shape[0], dtype=bool)
        if zs.shape[0] and xs_idx.shape[0]:
            for s, r in pair_reach(points[xs_idx], ucb[xs_idx, idx], zs, L_per_idx[idx], block):
                hit[s:s + r.shape[0]] = r.any(axis=1)
        masks[idx - 1, xs_idx[hit]] = True
        i, v = masked_argmax(var[:, 0], masks[idx - 1])
        per.append((i, float(np.sqrt(v)) if i >= 0 else 0.0))
    best_idx, best_std = -1, 0.0
    for i, s in per:                      # SafeOpt.py:120-122: max() keeps the first maximum
        if i >= 0 and (best_idx < 0 or s > best_std):
            best_idx, best_std = i, s
    return {"best_idx": best_idx, "best_std": best_std, "per_idx": per, "masks": masks}


def goose_target(points, S, Z, ucb, lcb, L_per_idx, block=1024):
    """GoOSE.py:80-114 restated on the grid.  For each constraint idx:
    O_idx = {z in Z : exists x in S with ucb_idx(x) - L||x-z+1e-8|| >= 0};
    pick argmin lcb_0 over O_idx; across idx keep the first smallest."""
    N, G = ucb.shape
    xs_idx = np.flatnonzero(S)
    zs_idx = np.flatnonzero(Z)
    masks = np.zeros((G - 1, N), dtype=bool)
    per = []
    for idx in range(1, G):
        hit = np.zeros(zs_idx.shape[0], dtype=bool)
        if zs_idx.shape[0] and xs_idx.shape[0]:
            for s, r in pair_reach(points[xs_idx], ucb[xs_idx, idx], points[zs_idx], L_per_idx[idx], block):
                hit |= r.any(axis=0)
        masks[idx - 1, zs_idx[hit]] = True
        i, v = masked_argmin(lcb[:, 0], masks[idx - 1])
        per.append((i, v))
    best_idx, best_val = -1, np.inf
    for i, v in per:                      # GoOSE.py:110-112: min() keeps the first minimum
        if i >= 0 and (best_idx < 0 or v < best_val):
            best_idx, best_val = i, v
    return {"best_idx": best_idx, "best_lcb": best_val, "per_idx": per, "masks": masks}


def explore_safeset(points, S, target):
    """GoOSE.py:116-119 -- nearest safe point to the target (Euclidean, raw space)."""
    diff = points - np.asarray(target)[None, :]
    d2 = np.sum(diff * diff, axis=1)
    i, v = masked_argmin(d2, S)
    return i, (float(np.sqrt(v)) if i >= 0 else np.inf)


# ----------------------------------------------------------------------------
# fantasy expander (north_star; NOT in the reference -- SURVEY.md section 8 row a12)
# ----------------------------------------------------------------------------
def fantasy_terms(points, ds, beta, factors=None):
    """Per point and GP, in *normalised* units: mu_n (N,G) [= raw mean / Y_std],
    var_n (N,G) clamped, V list of (N,n), xn (N,d)."""
    fac = factors if factors is not None else chol_factors(ds)
    mean, var, Vs = posterior_chol(points, ds, return_V=True, factors=fac)
    mu_n = mean / ds["Y_std"]            # = mu~ + Ybar/Ystd ; lcb_raw >= 0  <=>  mu_n - beta*sqrt(var_n) >= 0
    var_n = var / ds["Y_std"] ** 2
    xn = (points - ds["X_mean"]) / ds["X_std"]
    return mu_n, var_n, Vs, xn


def round_tf32(a):
    """Round float32 values to TF32 (10-bit mantissa) like PTX cvt.rna.tf32.f32: nearest, ties away from zero."""
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    b = (b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return b.view(np.float32)


def fantasy_counts(points, ds, beta, S, Z, block=256, dtype=np.float64):
    """For every candidate x in S and every z in Z: fantasise the observation
    y_i = ucb_i(x) for each constraint GP i>=1 (noise sn2_i + eps_f32, the diagonal
    actually added to K, GP_Safe.py:229-231) and test whether z becomes safe:
        c_i   = k_i(z,x) - v_z.v_x                      (posterior covariance)
        mu'_i = mu_i(z) + c_i * beta*sigma_i(x) / (sigma_i^2(x) + sn2_i)
        s2'_i = sigma_i^2(z) - c_i^2 / (sigma_i^2(x) + sn2_i)
        newly safe  <=>  mu'_i - beta*sqrt(max(s2'_i,0)) >= 0  for all i>=1.
    Returns counts g(x) (int64, length N, zero outside S).  ``dtype`` selects the
    precision of the V.V^T contraction only ("tf32" emulates the tensor-core operands: V rounded to TF32)."""
    N, d = points.shape
    G = ds["Y_norm"].shape[1]
    mu_n, var_n, Vs, xn = fantasy_terms(points, ds, beta)
    xs = np.flatnonzero(S)
    zs = np.flatnonzero(Z)
    counts = np.zeros(N, dtype=np.int64)
    if xs.size == 0 or zs.size == 0:
        return counts
    hyp = ds["hypopt"]
    for s in range(0, xs.size, block):
        xb = xs[s:s + block]
        ok = np.ones((zs.size, xb.size), dtype=bool)
        for i in range(1, G):
            ell, sf2, sn2 = unpack_hyper(hyp[:, i], d)
            sn2 = sn2 + EPS_F32
            kzx = sf2 * np.exp(-0.5 * sq_dist_direct(xn[zs], xn[xb], ell))           # (|Z|,B)
            if dtype == "tf32":      # operands rounded to TF32, exact products, wide accumulation
                acc = round_tf32(Vs[i][zs]).astype(np.float64) @ round_tf32(Vs[i][xb]).astype(np.float64).T
            else:
                acc = (Vs[i][zs].astype(dtype) @ Vs[i][xb].astype(dtype).T).astype(np.float64)
            c = kzx - acc
            denom = var_n[xb, i] + sn2
            mu_p = mu_n[zs, i][:, None] + c * (beta * np.sqrt(var_n[xb, i]) / denom)[None, :]
            s2_p = var_n[zs, i][:, None] - c * c / denom[None, :]
            ok &= (mu_p - beta * np.sqrt(np.maximum(s2_p, 0.0))) >= 0.0
        counts[xb] = ok.sum(axis=0)
    return counts


def fantasy_margin(points, ds, beta, S, Z, block=256):
    """Like fantasy_counts but returns, per (z,x) pair, min_i (mu' - beta*sigma') in
    normalised units -- used by the tests to find pairs within tolerance of 0."""
    N, d = points.shape
    G = ds["Y_norm"].shape[1]
    mu_n, var_n, Vs, xn = fantasy_terms(points, ds, beta)
    xs = np.flatnonzero(S)
    zs = np.flatnonzero(Z)
    out = np.full((zs.size, xs.size), np.inf)
    for i in range(1, G):
        ell, sf2, sn2 = unpack_hyper(ds["hypopt"][:, i], d)
        sn2 = sn2 + EPS_F32
        kzx = sf2 * np.exp(-0.5 * sq_dist_direct(xn[zs], xn[xs], ell))
        c = kzx - Vs[i][zs] @ Vs[i][xs].T
        denom = var_n[xs, i] + sn2
        mu_p = mu_n[zs, i][:, None] + c * (beta * np.sqrt(var_n[xs, i]) / denom)[None, :]
        s2_p = var_n[zs, i][:, None] - c * c / denom[None, :]
        out = np.minimum(out, mu_p - beta * np.sqrt(np.maximum(s2_p, 0.0)))
    return out


def fantasy_by_augmentation(x, zs, ds, beta):
    """Oracle of the oracle: add the fantasy observation (x, ucb_i(x)) to the data
    set of every constraint GP at FIXED normalisation and hyper-parameters and re-run
    the inverse-form inference at the points zs.  Returns lcb' (|zs|,G) raw units
    (column 0 = objective, untouched)."""
    x = np.asarray(x, dtype=np.float64).reshape(1, -1)
    zs = np.atleast_2d(zs)
    d = x.shape[1]
    G = ds["Y_norm"].shape[1]
    m0 = prior_mean(ds)
    mean_x, var_x = posterior_inv(x, ds)
    lcb_out = np.empty((zs.shape[0], G))
    xn = (x - ds["X_mean"]) / ds["X_std"]
    zn = (zs - ds["X_mean"]) / ds["X_std"]
    for i in range(G):
        ell, sf2, sn2 = unpack_hyper(ds["hypopt"][:, i], d)
        if i == 0:
            mz, vz = posterior_inv(zs, ds)
            lcb_out[:, 0] = mz[:, 0] - beta * np.sqrt(vz[:, 0])
            continue
        Xa = np.vstack([ds["X_norm"], xn])
        y_f = (mean_x[0, i] + beta * np.sqrt(var_x[0, i]) - ds["Y_mean"][i]) / ds["Y_std"][i]
        Ya = np.concatenate([ds["Y_norm"][:, i], [y_f]])
        Ka = cov_mat(Xa, Xa, ell, sf2) + (sn2 + EPS_F32) * np.eye(Xa.shape[0])
        iKa = np.linalg.inv(Ka)
        k = cov_mat(Xa, zn, ell, sf2)
        mu = m0[i] + k.T @ iKa @ (Ya - m0[i])
        s2 = np.maximum(0.0, sf2 - np.einsum("bn,nm,mb->b", k.T, iKa, k))
        mu = mu * ds["Y_std"][i] + ds["Y_mean"][i]
        s2 = s2 * ds["Y_std"][i] ** 2
        lcb_out[:, i] = mu - beta * np.sqrt(s2)
    return lcb_out


def expander_fantasy(points, ds, beta, S, Z, var, dtype=np.float64):
    """Fantasy-mode expander: G = {x in S : g(x) > 0}; best = argmax var_0 over G."""
    counts = fantasy_counts(points, ds, beta, S, Z, dtype=dtype)
    mask = counts > 0
    i, v = masked_argmax(var[:, 0], mask)
    return {"best_idx": i, "best_std": float(np.sqrt(v)) if i >= 0 else 0.0, "counts": counts, "mask": mask}


# ----------------------------------------------------------------------------
# whole steps (drivers' decision rules: test/test_SafeOpt.py:144-158, test/test_GoOSE.py:151-162)
# ----------------------------------------------------------------------------
def safeopt_step(points, ds, beta, mode="lipschitz", unsafe_rule="all", L_override=None, form="inv"):
    G = ds["Y_norm"].shape[1]
    mean, var = (posterior_inv if form == "inv" else posterior_chol)(points, ds)
    lcb, ucb = bounds(mean, var, beta)
    S = safe_mask(lcb)
    Z = unsafe_mask(lcb, unsafe_rule)
    m_idx, m_std, M, min_ucb = minimizer(var, lcb, ucb, S)
    out = {"mean": mean, "var": var, "S": S, "Z": Z, "M": M, "min_ucb0": min_ucb,
           "minimizer_idx": m_idx, "minimizer_std": m_std}
    if mode == "lipschitz":
        if L_override is None:
            Ls = [0.0] + [lipschitz_constant(points, ds, i) for i in range(1, G)]
            Lref = [Ls[G - 1]] * G            # SafeOpt.py:110 -- leaked loop variable i = n_fun-1
        else:
            Ls = list(L_override)
            Lref = list(L_override)
        ex = expander_lipschitz(points, S, Z, ucb, var, Lref)
        out.update({"L": Ls, "expander_masks": ex["masks"], "expander_per_idx": ex["per_idx"]})
    else:
        ex = expander_fantasy(points, ds, beta, S, Z, var)
        out.update({"counts": ex["counts"], "expander_masks": ex["mask"][None, :]})
    out.update({"expander_idx": ex["best_idx"], "expander_std": ex["best_std"]})
    # test/test_SafeOpt.py:153-158
    out["x_new_idx"] = m_idx if m_std > ex["best_std"] else ex["best_idx"]
    return out


def goose_step(points, ds, beta, unsafe_rule="all", L_override=None, form="inv"):
    G = ds["Y_norm"].shape[1]
    mean, var = (posterior_inv if form == "inv" else posterior_chol)(points, ds)
    lcb, ucb = bounds(mean, var, beta)
    S = safe_mask(lcb)
    Z = unsafe_mask(lcb, unsafe_rule)
    s_idx, s_lcb = minimize_obj_lcb(lcb, S)
    if L_override is None:
        Ls = [0.0] + [lipschitz_constant(points, ds, i) for i in range(1, G)]
        Lref = [Ls[G - 1]] * G                # GoOSE.py:100
    else:
        Ls = list(L_override)
        Lref = list(L_override)
    tg = goose_target(points, S, Z, ucb, lcb, Lref)
    out = {"mean": mean, "var": var, "S": S, "Z": Z, "L": Ls, "safe_min_idx": s_idx, "safe_min_lcb": s_lcb,
           "target_idx": tg["best_idx"], "target_lcb": tg["best_lcb"], "target_masks": tg["masks"],
           "target_per_idx": tg["per_idx"]}
    # test/test_GoOSE.py:158-162
    if s_lcb <= tg["best_lcb"] or tg["best_idx"] < 0:
        out["x_new_idx"] = s_idx
        out["explore_idx"] = -1
    else:
        e_idx, _ = explore_safeset(points, S, points[tg["best_idx"]])
        out["x_new_idx"] = e_idx
        out["explore_idx"] = e_idx
    return out
