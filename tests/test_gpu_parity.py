"""GPU parity tests: the CUDA grid pipeline (called through the C ABI) vs the NumPy oracle on the same
seeded inputs, and vs the committed golden fixtures at the reference's full 400x400 grid.

Tolerances (stated scale-relatively):
  posterior   |d mean| <= TOL * max|mean| (per GP),  |d var| <= TOL * sf2 * Ystd^2
              TOL = 1e-10 vs the Cholesky-form oracle, 2e-8 vs the reference's inverse form
              (the two FP64 formulations themselves differ by ~4e-10 at cond(K) ~ 1e7)
  sets        identical except points whose bound lies within SET_TOL*scale of the threshold
  arg-reductions  identical grid index, or (near-ties) an index whose oracle score is within
              1e-9 relative of the optimum
"""
import numpy as np
import pytest

from conftest import golden_ds

pytestmark = pytest.mark.gpu

TOL_CHOL = 1e-10
TOL_INV = 2e-8
SET_TOL = 1e-8
GOLD_REL = 2e-5     # committed golden scalars (inverse-form oracle) vs the trsm-form GPU path


def _capi():
    from sbo_b200 import _capi
    return _capi


def post_err(oracle, ds, m, v, mo, vo):
    d = ds["X_norm"].shape[1]
    em, ev = 0.0, 0.0
    for i in range(m.shape[1]):
        _, sf2, _ = oracle.unpack_hyper(ds["hypopt"][:, i], d)
        em = max(em, np.max(np.abs(m[:, i] - mo[:, i])) / max(np.max(np.abs(mo[:, i])), ds["Y_std"][i]))
        ev = max(ev, np.max(np.abs(v[:, i] - vo[:, i])) / (sf2 * ds["Y_std"][i] ** 2))
    return em, ev


def check_mask(got, want, margin, scale, what):
    """got/want bool masks; margin = |bound - threshold| per point.  Mismatches only where ambiguous."""
    bad = got != want
    if bad.any():
        assert np.all(margin[bad] <= SET_TOL * scale), f"{what}: {bad.sum()} mismatches beyond tolerance"
    return int(bad.sum())


# ------------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [4, 14, 35, 70, 200])
def test_model_factorisation(engine, oracle, c3, n):
    rng = np.random.default_rng(n)
    if n <= 35:
        X, Y, hyp = c3["X"][:n], c3["Y"][:n], c3[f"hyp_{35 if n > 20 else (20 if n > 5 else 5)}"]
    else:
        X = rng.uniform([4, 70], [7, 100], size=(n, 2))
        Y = np.column_stack([np.sin(X[:, 0]) * X[:, 1], 0.1 - 0.01 * X[:, 0], 0.05 + 0.001 * X[:, 1]])
        hyp = c3["hyp_35"]
    ds = oracle.make_inference_datasets(X, Y, hyp)
    engine.set_model(ds)
    L, W, alpha = engine.get_model()
    fac = oracle.chol_factors(ds)
    for i in range(3):
        Lo, ao = fac[i]
        K = oracle.build_K(ds["X_norm"], ds["hypopt"][:, i])
        assert np.max(np.abs(L[i] @ L[i].T - K)) <= 1e-12 * np.max(np.abs(K))
        assert np.max(np.abs(L[i] - Lo)) <= 1e-10 * np.max(np.abs(Lo))
        assert np.max(np.abs(W[i] @ Lo - np.eye(n))) <= 1e-9
        assert np.max(np.abs(alpha[i] - ao)) <= 1e-8 * max(1.0, np.max(np.abs(ao)))


def test_not_positive_definite_is_an_error(engine, oracle, c1):
    import sbo_b200
    ds = golden_ds(oracle, c1, 4)
    ds = dict(ds)
    ds["X_norm"] = np.vstack([ds["X_norm"], ds["X_norm"][:1]])          # duplicated input ...
    ds["Y_norm"] = np.vstack([ds["Y_norm"], ds["Y_norm"][:1]])
    hyp = ds["hypopt"].copy()
    hyp[3, :] = -30.0                                                   # ... and (almost) no noise
    hyp[2, :] = 6.0
    ds["hypopt"] = hyp
    try:
        engine.set_model(ds)                                            # eps_f32*I may or may not rescue it
    except sbo_b200.SboError as e:
        assert "positive definite" in str(e)
    engine.set_model(golden_ds(oracle, c1, 4))                          # the context stays usable


# ------------------------------------------------------------------------------------------------
# posterior on the reference's grids (C1 Benoit, C3 WOR; 400 x 400)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n", [("c1", 4), ("c1", 9), ("c1", 14), ("c3", 5), ("c3", 20), ("c3", 35)])
def test_posterior_full_grid(engine, oracle, request, name, n):
    gold = request.getfixturevalue(name)
    ds = golden_ds(oracle, gold, n)
    engine.set_model(ds)
    engine.set_grid(gold["lo"], gold["hi"], [400, 400])
    m, v = engine.posterior()
    assert m.shape == (160000, ds["Y_norm"].shape[1]) and np.all(v >= 0)
    pts = oracle.make_grid(gold["lo"], gold["hi"], [400, 400])
    mb, vb = oracle.posterior_chol(pts, ds)
    em, ev = post_err(oracle, ds, m, v, mb, vb)
    assert em <= TOL_CHOL and ev <= TOL_CHOL, (em, ev)
    ma, va = oracle.posterior_inv(pts, ds)
    em, ev = post_err(oracle, ds, m, v, ma, va)
    assert em <= TOL_INV and ev <= TOL_INV, (em, ev)


@pytest.mark.parametrize("name,sizes", [("c1", [4, 9, 14]), ("c3", [5, 20, 35])])
def test_point_posterior_golden(engine, oracle, request, name, sizes):
    # the points the reference's scripts print at (test_SafeOpt.py:47, test_GP_Safe.py:25,28,44)
    gold = request.getfixturevalue(name)
    for n in sizes:
        ds = golden_ds(oracle, gold, n)
        engine.set_model(ds)
        m, v = engine.point_posterior(gold["test_points"])
        em, ev = post_err(oracle, ds, m, v, gold[f"tp_mean_{n}"], gold[f"tp_var_{n}"])
        assert em <= TOL_INV and ev <= TOL_INV, (n, em, ev)


def test_explicit_points_and_ragged_sizes(engine, oracle, c3):
    ds = golden_ds(oracle, c3, 20)
    engine.set_model(ds)
    rng = np.random.default_rng(3)
    for N in [1, 31, 33, 64, 1000, 4097]:
        pts = rng.uniform(c3["lo"], c3["hi"], size=(N, 2))
        engine.set_points(pts)
        m, v = engine.posterior()
        mb, vb = oracle.posterior_chol(pts, ds)
        em, ev = post_err(oracle, ds, m, v, mb, vb)
        assert em <= TOL_CHOL and ev <= TOL_CHOL, (N, em, ev)
        mp, vp = engine.point_posterior(pts[: min(N, 7)])
        assert np.max(np.abs(mp - m[: min(N, 7)])) == 0.0 and np.max(np.abs(vp - v[: min(N, 7)])) == 0.0


def test_properties_reference_prints(engine, oracle, c1):
    ds = golden_ds(oracle, c1, 9)
    engine.set_model(ds)
    m, v = engine.point_posterior(c1["X"][:9])          # interpolation, test_GP_Safe.py:33-41
    assert np.max(np.abs(m - c1["Y"][:9])) < 0.05 and np.max(v) < 1e-2
    m, v = engine.point_posterior(np.array([[10.0, 10.0]]))   # constraint prior, test_GP_Safe.py:43-47
    assert m[0, 1] == pytest.approx(-ds["Y_mean"][1], rel=1e-6)


def test_mean_gradient_and_lipschitz(engine, oracle, c3):
    ds = golden_ds(oracle, c3, 20)
    engine.set_model(ds)
    x = np.array([[6.0, 85.0], [5.1, 78.0], [4.2, 99.0]])
    for i in range(3):
        g = engine.point_mean_grad(x, i)
        go = oracle.mean_grad(x, ds, i)
        assert np.max(np.abs(g - go)) <= 1e-9 * max(1.0, np.max(np.abs(go)))
    engine.set_grid(c3["lo"], c3["hi"], [150, 150])
    engine.posterior(with_grad=True, fetch=False)
    L = engine.lipschitz()
    pts = oracle.make_grid(c3["lo"], c3["hi"], [150, 150])
    for i in range(3):
        assert L[i] == pytest.approx(oracle.lipschitz_constant(pts, ds, i), rel=1e-9)


# ------------------------------------------------------------------------------------------------
# sets + arg-reductions on the full reference grid
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,beta", [("c1", 4, 3.0), ("c1", 9, 3.0), ("c1", 14, 3.0),
                                          ("c3", 5, 2.0), ("c3", 20, 2.0), ("c3", 35, 2.0)])
@pytest.mark.parametrize("rule", ["all", "any"])
def test_sets_full_grid(engine, oracle, request, name, n, beta, rule):
    capi = _capi()
    gold = request.getfixturevalue(name)
    ds = golden_ds(oracle, gold, n)
    engine.set_model(ds)
    engine.set_grid(gold["lo"], gold["hi"], [400, 400])
    m, v = engine.posterior()
    s = engine.sets(beta, capi.UNSAFE_ALL if rule == "all" else capi.UNSAFE_ANY)
    # oracle sets from the oracle's own posterior (inverse form = the reference's)
    pts = oracle.make_grid(gold["lo"], gold["hi"], [400, 400])
    mo, vo = oracle.posterior_inv(pts, ds)
    lcb, ucb = oracle.bounds(mo, vo, beta)
    S, Z = oracle.safe_mask(lcb), oracle.unsafe_mask(lcb, rule)
    scale = float(np.max(ds["Y_std"][1:]))
    cmargin = np.min(np.abs(lcb[:, 1:]), axis=1)
    nS = check_mask(engine.mask(capi.MASK_SAFE), S, cmargin, scale, "S")
    nZ = check_mask(engine.mask(capi.MASK_UNSAFE), Z, cmargin, scale, "Z")
    assert abs(s["n_safe"] - S.sum()) <= nS and abs(s["n_unsafe"] - Z.sum()) <= nZ
    # same sets from the GPU's own posterior must be bit-identical (bounds are FMA-free on both sides)
    lg, ug = oracle.bounds(m, v, beta)
    assert np.array_equal(engine.mask(capi.MASK_SAFE), oracle.safe_mask(lg))
    assert np.array_equal(engine.mask(capi.MASK_UNSAFE), oracle.unsafe_mask(lg, rule))
    i_u, v_u = oracle.minimize_obj_ucb(ug, oracle.safe_mask(lg))
    i_l, v_l = oracle.minimize_obj_lcb(lg, oracle.safe_mask(lg))
    assert (s["min_ucb0_idx"], s["min_lcb0_idx"]) == (i_u, i_l)
    if i_u >= 0:
        assert s["min_ucb0"] == v_u and s["min_lcb0"] == v_l
    mi, mstd, M, _ = oracle.minimizer(v, lg, ug, oracle.safe_mask(lg))
    assert np.array_equal(engine.mask(capi.MASK_MIN), M) and s["n_min"] == M.sum() and s["minimizer_idx"] == mi
    # golden summary (oracle inverse form at the full grid, rule 'all'): sizes within the ambiguous count
    if rule == "all":
        g = gold[f"summary_{n}"]
        assert abs(s["n_safe"] - g[0]) <= nS + 2 and abs(s["n_unsafe"] - g[2]) <= nZ + 2
        sc = gold[f"scalars_{n}"]
        assert s["min_ucb0"] == pytest.approx(sc[0], rel=GOLD_REL, abs=1e-9)
        assert np.sqrt(s["minimizer_var"]) == pytest.approx(sc[1], rel=GOLD_REL)
        # chosen minimiser: identical index unless the oracle's optimum is a near-tie
        if s["minimizer_idx"] != g[3]:
            assert vo[s["minimizer_idx"], 0] >= vo[g[3], 0] * (1 - 1e-9)


def test_sets_strict_and_plot_mask(engine, oracle, c1):
    capi = _capi()
    ds = golden_ds(oracle, c1, 9)
    engine.set_model(ds)
    engine.set_grid(c1["lo"], c1["hi"], [400, 400])
    m, v = engine.posterior()
    engine.sets(3.0, capi.UNSAFE_ALL, strict=True)
    want = (m[:, 1] - 3.0 * np.sqrt(v[:, 1])) > 0.0          # test/test_SafeOpt.py:337-338
    assert np.array_equal(engine.mask(capi.MASK_SAFE), want)


def test_empty_sets_and_no_constraints(engine, oracle, c1):
    capi = _capi()
    ds = golden_ds(oracle, c1, 4)
    engine.set_model(ds)
    engine.set_grid([5.0, 5.0], [6.0, 6.0], [9, 7])             # far from the data: S is empty
    engine.posterior(with_grad=True, fetch=False)
    s = engine.sets(3.0)
    assert s["n_safe"] == 0 and s["min_ucb0_idx"] == -1 and s["minimizer_idx"] == -1 and s["n_min"] == 0
    ex = engine.expander(3.0, np.array([1.0, 1.0]))
    assert ex["best_idx"] == -1 and ex["n_x"] == 0
    tg = engine.goose_target(3.0, np.array([1.0, 1.0]))
    assert tg["best_idx"] == -1
    st = engine.safeopt_step(ds, 3.0)
    assert st["x_new_idx"] == -1
    # G = 1: objective only -> every point is safe, there is no unsafe set
    ds1 = oracle.make_inference_datasets(c1["X"][:9], c1["Y"][:9, :1], c1["hyp_9"][:, :1])
    engine.set_model(ds1)
    engine.set_grid(c1["lo"], c1["hi"], [50, 30])
    engine.posterior(fetch=False)
    s = engine.sets(3.0)
    assert s["n_safe"] == 1500 and s["n_unsafe"] == 0 and s["minimizer_idx"] >= 0


# ------------------------------------------------------------------------------------------------
# Lipschitz-mode pair kernels vs the oracle's all-pairs brute force (reduced grids: the oracle is O(|S||Z|))
# ------------------------------------------------------------------------------------------------
def _pair_inputs(engine, oracle, gold, n, beta, pts_per_dim, rule="all"):
    capi = _capi()
    ds = golden_ds(oracle, gold, n)
    engine.set_model(ds)
    engine.set_grid(gold["lo"], gold["hi"], pts_per_dim)
    m, v = engine.posterior(with_grad=True)
    s = engine.sets(beta, capi.UNSAFE_ALL if rule == "all" else capi.UNSAFE_ANY)
    pts = oracle.make_grid(gold["lo"], gold["hi"], pts_per_dim)
    lcb, ucb = oracle.bounds(m, v, beta)
    return ds, pts, m, v, lcb, ucb, oracle.safe_mask(lcb), oracle.unsafe_mask(lcb, rule), s


@pytest.mark.parametrize("name,n,beta,grid", [("c1", 4, 3.0, [90, 70]), ("c1", 9, 3.0, [101, 77]), ("c1", 14, 3.0, [64, 64]),
                                               ("c3", 5, 2.0, [80, 80]), ("c3", 20, 2.0, [75, 90]), ("c3", 35, 2.0, [60, 60])])
def test_expander_and_target_lipschitz(engine, oracle, request, name, n, beta, grid):
    capi = _capi()
    gold = request.getfixturevalue(name)
    ds, pts, m, v, lcb, ucb, S, Z, s = _pair_inputs(engine, oracle, gold, n, beta, grid)
    G = m.shape[1]
    Lg = engine.lipschitz()
    # two regimes: the reference's L (max-gradient; nearly every safe point qualifies) and a 5x larger,
    # discriminating one
    for mult in [1.0, 5.0]:
        L = np.full(G, Lg[G - 1] * mult)
        ex = engine.expander(beta, L)
        exo = oracle.expander_lipschitz(pts, S, Z, ucb, v, L)
        for c in range(G - 1):
            got = engine.mask(capi.MASK_EXPANDER, c)
            bad = got != exo["masks"][c]
            assert bad.sum() <= 2, (mult, c, bad.sum())        # only exact-threshold rounding may differ
        assert ex["n_x"] == S.sum() and ex["n_z"] == Z.sum()
        assert ex["pairs_algorithmic"] == S.sum() * Z.sum() * (G - 1)
        if exo["best_idx"] >= 0 and not any((engine.mask(capi.MASK_EXPANDER, c) != exo["masks"][c]).any() for c in range(G - 1)):
            assert ex["best_idx"] == exo["best_idx"]
            assert np.sqrt(ex["best_value"]) == pytest.approx(exo["best_std"], rel=1e-12)
            assert ex["per_idx"] == [p for p, _ in exo["per_idx"]]
        tg = engine.goose_target(beta, L)
        tgo = oracle.goose_target(pts, S, Z, ucb, lcb, L)
        same = True
        for c in range(G - 1):
            bad = engine.mask(capi.MASK_TARGET, c) != tgo["masks"][c]
            assert bad.sum() <= 2, (mult, c, bad.sum())
            same = same and not bad.any()
        if same:
            assert tg["best_idx"] == tgo["best_idx"]
            if tgo["best_idx"] >= 0:
                assert tg["best_value"] == tgo["best_lcb"]
                e_idx, dist = engine.argreduce(capi.ARGMIN_DIST, capi.MASK_SAFE, 0, pts[tgo["best_idx"]])
                eo, do = oracle.explore_safeset(pts, S, pts[tgo["best_idx"]])
                assert e_idx == eo and dist == pytest.approx(do, rel=1e-14)


@pytest.mark.parametrize("name,n,beta,grid,mult", [("c1", 9, 3.0, [200, 160], 1.0), ("c1", 14, 3.0, [160, 200], 6.0),
                                                   ("c3", 20, 2.0, [150, 170], 1.0), ("c3", 35, 2.0, [190, 130], 20.0)])
def test_tile_culling_is_exact(engine, oracle, request, name, n, beta, grid, mult):
    """The bounding-box culling of the Lipschitz pair kernels skips tiles only: every mask, optimum and index is
    bit-identical to the un-culled all-pairs run, and fewer (never more) pairs are evaluated."""
    capi = _capi()
    gold = request.getfixturevalue(name)
    ds = golden_ds(oracle, gold, n)
    engine.set_model(ds)
    engine.set_grid(gold["lo"], gold["hi"], grid)
    engine.posterior(with_grad=True, fetch=False)
    engine.sets(beta, capi.UNSAFE_ALL)
    G = ds["Y_norm"].shape[1]
    L = np.full(G, engine.lipschitz()[G - 1] * mult)        # larger L = smaller radii = more culling
    res = {}
    for cull in (0, 1):
        engine.set_option("pair_cull", cull)
        ex = engine.expander(beta, L)
        em = [engine.mask(capi.MASK_EXPANDER, c) for c in range(G - 1)]
        tg = engine.goose_target(beta, L)
        tm = [engine.mask(capi.MASK_TARGET, c) for c in range(G - 1)]
        res[cull] = (ex, em, tg, tm)
    engine.set_option("pair_cull", 1)
    (ex0, em0, tg0, tm0), (ex1, em1, tg1, tm1) = res[0], res[1]
    for c in range(G - 1):
        assert np.array_equal(em0[c], em1[c]) and np.array_equal(tm0[c], tm1[c])
    for k in ("best_idx", "best_value", "per_idx", "per_value", "n_hit", "pairs_algorithmic"):
        assert ex0[k] == ex1[k] and tg0[k] == tg1[k], k
    assert ex1["pairs_evaluated"] <= ex0["pairs_evaluated"] and tg1["pairs_evaluated"] <= tg0["pairs_evaluated"]
    print(f"culling {name} n={n} x{mult}: expander {ex0['pairs_evaluated']} -> {ex1['pairs_evaluated']} pair-evals, "
          f"target {tg0['pairs_evaluated']} -> {tg1['pairs_evaluated']}; hits {ex1['n_hit']}/{tg1['n_hit']}")


@pytest.mark.parametrize("name,n,beta", [("c1", 9, 3.0), ("c3", 20, 2.0)])
def test_whole_steps_match_oracle(engine, oracle, request, name, n, beta):
    gold = request.getfixturevalue(name)
    ds = golden_ds(oracle, gold, n)
    grid = [72, 56]
    engine.set_grid(gold["lo"], gold["hi"], grid)
    pts = oracle.make_grid(gold["lo"], gold["hi"], grid)
    st = engine.safeopt_step(ds, beta)
    so = oracle.safeopt_step(pts, ds, beta, form="chol")
    assert st["n_safe"] == so["S"].sum() and st["n_min"] == so["M"].sum()
    assert st["minimizer_idx"] == so["minimizer_idx"] and st["expander_idx"] == so["expander_idx"]
    assert st["x_new_idx"] == so["x_new_idx"]
    assert st["L"][-1] == pytest.approx(so["L"][-1], rel=1e-9)
    gs = engine.goose_step(ds, beta)
    go = oracle.goose_step(pts, ds, beta, form="chol")
    assert gs["min_lcb0_idx"] == go["safe_min_idx"] and gs["target_idx"] == go["target_idx"]
    assert gs["x_new_idx"] == go["x_new_idx"]


def test_full_grid_step_vs_golden_summary(engine, oracle, c1, c3):
    # full 400x400 Lipschitz steps against the committed oracle summaries (make_golden.py)
    for gold, n, beta in [(c1, 9, 3.0), (c1, 14, 3.0), (c3, 20, 2.0)]:
        ds = golden_ds(oracle, gold, n)
        engine.set_grid(gold["lo"], gold["hi"], [400, 400])
        st = engine.safeopt_step(ds, beta)
        g, sc = gold[f"summary_{n}"], gold[f"scalars_{n}"]
        assert abs(st["n_safe"] - g[0]) <= 3 and abs(st["n_min"] - g[1]) <= 3
        assert st["L"][-1] == pytest.approx(sc[3], rel=1e-7)
        assert st["expander_std"] == pytest.approx(sc[2], rel=GOLD_REL)
        assert st["minimizer_std"] == pytest.approx(sc[1], rel=GOLD_REL)
        assert st["x_new_idx"] == g[5] or st["expander_std"] == pytest.approx(sc[2], rel=1e-9)
        gs = engine.goose_step(ds, beta)
        assert gs["min_lcb0"] == pytest.approx(sc[4], rel=GOLD_REL, abs=1e-9)
        assert gs["target_lcb"] == pytest.approx(sc[5], rel=GOLD_REL, abs=1e-9)
        assert gs["min_lcb0_idx"] == g[6] and gs["target_idx"] == g[7] and gs["x_new_idx"] == g[8]


# ------------------------------------------------------------------------------------------------
# fantasy expander, FP64 kernel vs oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,beta,grid,rule", [("c1", 9, 3.0, [40, 36], "all"), ("c3", 20, 2.0, [33, 47], "any"),
                                                    ("c3", 35, 2.0, [30, 30], "all")])
def test_fantasy_fp64_counts(engine, oracle, request, name, n, beta, grid, rule):
    capi = _capi()
    gold = request.getfixturevalue(name)
    ds = golden_ds(oracle, gold, n)
    engine.set_model(ds)
    engine.set_grid(gold["lo"], gold["hi"], grid)
    m, v = engine.posterior(keep_v=1)
    engine.sets(beta, capi.UNSAFE_ALL if rule == "all" else capi.UNSAFE_ANY)
    ex = engine.expander(beta, None, capi.MODE_FANTASY, capi.PREC_FP64, want_counts=True)
    pts = oracle.make_grid(gold["lo"], gold["hi"], grid)
    lcb, _ = oracle.bounds(m, v, beta)
    S, Z = oracle.safe_mask(lcb), oracle.unsafe_mask(lcb, rule)
    want = oracle.fantasy_counts(pts, ds, beta, S, Z)
    margin = oracle.fantasy_margin(pts, ds, beta, S, Z)                 # (|Z|,|S|)
    amb = (np.abs(margin) <= 1e-9).sum(axis=0)                          # pairs within tolerance of the threshold
    diff = np.abs(ex["counts"][S].astype(np.int64) - want[S])
    assert np.all(diff <= amb), (diff.max(), amb.max())
    assert (ex["counts"][~S] == 0).all()
    assert ex["pairs_algorithmic"] == S.sum() * Z.sum() * (m.shape[1] - 1)
    if not diff.any():
        eo = oracle.expander_fantasy(pts, ds, beta, S, Z, v)
        assert ex["best_idx"] == eo["best_idx"] and ex["n_hit"] == eo["mask"].sum()
        assert np.array_equal(engine.mask(capi.MASK_EXPANDER, 0), eo["mask"])


# ------------------------------------------------------------------------------------------------
# higher-dimensional synthetic recipe (C4/C5 family) at a size the oracle finishes in seconds,
# and size-independent properties at the full C4 size
# ------------------------------------------------------------------------------------------------
def test_synthetic_small_all_stages(engine, oracle):
    from sbo_b200 import workloads
    capi = _capi()
    ds, lo, hi, pts_per_dim, beta = workloads.small(d=3, pts_per_dim=14, n=70, seed=7, G=3)
    dso = dict(ds)
    dso["invKopt"] = [np.linalg.inv(oracle.build_K(ds["X_norm"], ds["hypopt"][:, i])) for i in range(3)]
    engine.set_model(ds)
    engine.set_grid(lo, hi, pts_per_dim)
    m, v = engine.posterior(with_grad=True, keep_v=1)
    pts = oracle.make_grid(lo, hi, pts_per_dim)
    mb, vb = oracle.posterior_chol(pts, dso)
    em, ev = post_err(oracle, dso, m, v, mb, vb)
    assert em <= TOL_CHOL and ev <= TOL_CHOL
    s = engine.sets(beta, capi.UNSAFE_ANY)
    lcb, ucb = oracle.bounds(m, v, beta)
    S, Z = oracle.safe_mask(lcb), oracle.unsafe_mask(lcb, "any")
    assert s["n_safe"] == S.sum() > 0 and s["n_unsafe"] == Z.sum() > 0
    L = np.full(3, 4.0 * engine.lipschitz()[2])
    ex = engine.expander(beta, L)
    exo = oracle.expander_lipschitz(pts, S, Z, ucb, v, L)
    for c in range(2):
        assert (engine.mask(capi.MASK_EXPANDER, c) != exo["masks"][c]).sum() <= 2
    fz = engine.expander(beta, None, capi.MODE_FANTASY, capi.PREC_FP64, want_counts=True)
    want = oracle.fantasy_counts(pts, dso, beta, S, Z)
    amb = (np.abs(oracle.fantasy_margin(pts, dso, beta, S, Z)) <= 1e-9).sum(axis=0)
    assert np.all(np.abs(fz["counts"][S].astype(np.int64) - want[S]) <= amb)


def test_c4_full_size_properties(engine, oracle):
    """BASELINE config C4 at full size (N = 2^20, n = 512, G = 4): properties that need no O(N n^2) oracle."""
    from sbo_b200 import workloads
    capi = _capi()
    ds, lo, hi, pts_per_dim, beta = workloads.c4()
    engine.set_model(ds)
    engine.set_grid(lo, hi, pts_per_dim)
    m, v = engine.posterior(with_grad=True)
    assert m.shape == (1 << 20, 4) and np.all(np.isfinite(m)) and np.all(v >= 0)
    sf2 = 1.0
    assert np.all(v <= sf2 * ds["Y_std"] ** 2 * (1 + 1e-12))            # posterior variance <= prior variance
    # a random sample of grid points against the oracle (Cholesky form)
    rng = np.random.default_rng(11)
    idx = np.sort(rng.choice(1 << 20, size=3000, replace=False))
    pts = np.stack([engine.point_coords(i) for i in idx])
    dso = dict(ds)
    mb, vb = oracle.posterior_chol(pts, dso)
    em, ev = post_err(oracle, dso, m[idx], v[idx], mb, vb)
    assert em <= TOL_CHOL and ev <= TOL_CHOL, (em, ev)
    # sets: popcounts agree with the masks, M subset of S, Z disjoint from S; idempotence
    s1 = engine.sets(beta)
    S, Z, M = engine.mask(capi.MASK_SAFE), engine.mask(capi.MASK_UNSAFE), engine.mask(capi.MASK_MIN)
    assert s1["n_safe"] == S.sum() and s1["n_unsafe"] == Z.sum() and s1["n_min"] == M.sum()
    assert not (M & ~S).any() and not (S & Z).any()
    assert 0.03 < S.mean() < 0.4
    s2 = engine.sets(beta)
    assert s1 == s2
    # far-from-data prior: the constraint mean tends to -Y_mean (GP_Safe.py:331)
    far, _ = engine.point_posterior(np.full((1, 4), 50.0))
    assert np.allclose(far[0, 1:], -ds["Y_mean"][1:], rtol=1e-9)
    # interpolation at the training inputs
    Xraw = ds["X_norm"] * ds["X_std"] + ds["X_mean"]
    Yraw = ds["Y_norm"] * ds["Y_std"] + ds["Y_mean"]
    mt, vt = engine.point_posterior(Xraw)
    assert np.max(np.abs(mt - Yraw) / ds["Y_std"]) < 0.2 and np.max(vt / ds["Y_std"] ** 2) < 0.05


@pytest.mark.parametrize("world,block", [(2, 32), (3, 64), (8, 32), (8, 256)])
def test_rotated_block_cyclic_shards_tile_the_grid(engine, oracle, c3, world, block):
    """sbo_set_shard_cyclic: the ranks' shards are disjoint, cover the grid, and local point p is the global point
    the header documents (posterior of the shard == rows of the full-grid posterior)."""
    ds = golden_ds(oracle, c3, 20)
    engine.set_model(ds)
    grid = [50, 37]                                    # 1850 points: ragged last block, partial last super-block
    engine.set_grid(c3["lo"], c3["hi"], grid)
    m_full, v_full = engine.posterior()
    N = m_full.shape[0]
    nblk = (N + block - 1) // block
    seen = np.zeros(N, dtype=int)
    for rank in range(world):
        sb = np.arange((nblk + world - 1) // world)
        gb = sb * world + (rank + sb + sb // world + sb // (world * world)) % world
        gb = gb[gb < nblk]
        gidx = np.concatenate([np.arange(b * block, min(N, (b + 1) * block)) for b in gb])
        cnt = engine.set_shard_cyclic(rank, world, block)
        assert cnt == gidx.size
        m, v = engine.posterior()
        np.testing.assert_array_equal(m, m_full[gidx])
        np.testing.assert_array_equal(v, v_full[gidx])
        seen[gidx] += 1
        # arg-reductions report GLOBAL indices
        s = engine.sets_pass1(2.0, 0)
        if s["min_lcb0_idx"] >= 0:
            assert s["min_lcb0_idx"] in set(gidx.tolist())
    assert (seen == 1).all()
    engine.set_grid(c3["lo"], c3["hi"], grid)           # back to the unsharded default


# ------------------------------------------------------------------------------------------------
# fantasy expander, TF32 tcgen05/TMEM kernel.  Three checks, tolerances stated on the FP64 margin
#   m(z,x) = min_i (mu'_i - beta*sigma'_i)   (normalised units)  of a (z,x) pair:
#   (1) implementation: vs the oracle evaluated with the SAME TF32-rounded operands (exact products, wide
#       accumulation, FP64 epilogue).  Only FP32 accumulation order / FP32 epilogue rounding may differ:
#       decisions may differ where |m| <= IMPL_TOL * sf2 * (1 + gain(x)).
#   (2) TF32 mode vs FP64: operand rounding (2^-11 relative) perturbs v_x.v_z by up to ~1e-3*sf2 and the rank-1
#       update multiplies that by gain(x) = beta*sigma(x)/(sigma^2(x)+sn2): decisions may differ where
#       |m| <= TF32_TOL * sf2 * (1 + gain(x)).  (For ill-conditioned fits, sn2 ~ 5e-5, gain reaches ~200 and the
#       single-pass TF32 mode is not usable -- that is what TF32X3 is for.)
#   (3) TF32X3 mode (split operands, three passes) vs FP64: |m| <= IMPL_TOL * sf2 * (1 + gain(x)).
# ------------------------------------------------------------------------------------------------
TF32_TOL = 1e-3
IMPL_TOL = 2e-5


def _fantasy_tc_case(engine, oracle, ds, lo, hi, grid, beta, rule, variant, precision):
    capi = _capi()
    prec, keep_v = capi.PRECISIONS[precision]
    engine.set_option("fantasy_variant", variant)
    engine.set_model(ds)
    engine.set_grid(lo, hi, grid)
    m, v = engine.posterior(keep_v=keep_v)
    engine.sets(beta, capi.UNSAFE_ALL if rule == "all" else capi.UNSAFE_ANY)
    # defaults: exact pruning on, FP64 refinement of the pairs inside the tensor-core error bound on
    ex_ref = engine.expander(beta, None, capi.MODE_FANTASY, prec, want_counts=True)
    try:
        engine.set_option("fantasy_refine", 0)                                        # tensor-core decisions as they are
        ex = engine.expander(beta, None, capi.MODE_FANTASY, prec, want_counts=True)
        engine.set_option("fantasy_prune", 0)                                         # ... and every pair through the GEMM
        ex_all = engine.expander(beta, None, capi.MODE_FANTASY, prec, want_counts=True)
    finally:
        engine.set_option("fantasy_prune", 1)
        engine.set_option("fantasy_refine", 2)
        engine.set_option("fantasy_variant", -1)
    pts = oracle.make_grid(lo, hi, grid)
    lcb, _ = oracle.bounds(m, v, beta)
    S, Z = oracle.safe_mask(lcb), oracle.unsafe_mask(lcb, rule)
    got = ex["counts"][S].astype(np.int64)
    assert (ex["counts"][~S] == 0).all()
    assert ex["n_x"] == S.sum() and ex["n_z"] == Z.sum()
    margin = np.abs(oracle.fantasy_margin(pts, ds, beta, S, Z))            # (|Z|,|S|)
    G, d = m.shape[1], pts.shape[1]
    gain = np.zeros(S.sum())
    sf2max = 0.0
    for i in range(1, G):
        _, sf2, sn2 = oracle.unpack_hyper(ds["hypopt"][:, i], d)
        sf2max = max(sf2max, sf2)
        vn = v[S, i] / ds["Y_std"][i] ** 2
        gain = np.maximum(gain, beta * np.sqrt(vn) / (vn + sn2 + oracle.EPS_F32))
    scale = sf2max * (1.0 + gain)[None, :]
    w64 = oracle.fantasy_counts(pts, ds, beta, S, Z)[S]
    out = {"newly_safe_fp64": int(w64.sum())}
    # refined (the default): the counts are the FP64 counts; only a pair within FP64 rounding of the threshold may differ
    exact = (margin <= 1e-12 * scale).sum(axis=0)
    dr = np.abs(ex_ref["counts"][S].astype(np.int64) - w64)
    assert np.all(dr <= exact), (precision, int(dr.max()), int((dr > exact).sum()))
    assert ex_ref["n_ambiguous"] >= ex_ref["n_refined_safe"] >= 0
    out.update(refined=int(ex_ref["n_ambiguous"]), refined_safe=int(ex_ref["n_refined_safe"]))
    assert np.all(got <= ex_all["counts"][S]) and ex["pairs_evaluated"] <= ex_all["pairs_evaluated"]
    if precision == "tf32":
        wtf = oracle.fantasy_counts(pts, ds, beta, S, Z, dtype="tf32")[S]
        amb_impl = (margin <= IMPL_TOL * scale).sum(axis=0)
        # (1) is a statement about the un-pruned kernel: the pruning removes TF32 false positives the bound refutes
        d1 = np.abs(ex_all["counts"][S].astype(np.int64) - wtf)
        # the TF32-operand oracle decides on ITS margin; allow its own near-threshold pairs as well
        assert np.all(d1 <= amb_impl + (np.abs(wtf - w64) > 0) * 2 + 2), (int(d1.max()), int(amb_impl.max()))
        amb = (margin <= TF32_TOL * scale).sum(axis=0)
        d2 = np.abs(got - w64)
        assert np.all(d2 <= amb), (int(d2.max()), int(amb.max()), int((d2 > amb).sum()))
        out.update(diff_vs_tf32_oracle=int(d1.sum()), diff_vs_fp64=int(d2.sum()), ambiguous=int(amb.sum()))
    else:
        amb = (margin <= IMPL_TOL * scale).sum(axis=0)
        d2 = np.abs(got - w64)
        assert np.all(d2 <= amb), (int(d2.max()), int(amb.max()), int((d2 > amb).sum()))
        out.update(diff_vs_fp64=int(d2.sum()), ambiguous=int(amb.sum()))
    return out


@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 5, 7])
@pytest.mark.parametrize("name,n,beta,grid,rule", [("c1", 9, 3.0, [48, 40], "all"), ("c3", 20, 2.0, [45, 61], "any"),
                                                    ("c3", 35, 2.0, [70, 50], "all")])
def test_fantasy_tensor_core_counts(engine, oracle, request, name, n, beta, grid, rule, variant, precision):
    gold = request.getfixturevalue(name)
    ds = golden_ds(oracle, gold, n)
    r = _fantasy_tc_case(engine, oracle, ds, gold["lo"], gold["hi"], grid, beta, rule, variant, precision)
    print(f"fantasy {precision} v{variant} {name} n={n}: {r}")


@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 5, 7])
def test_fantasy_tensor_core_synthetic(engine, oracle, variant, precision):
    from sbo_b200 import workloads
    for (d, ppd, n, G) in [(3, 14, 70, 3), (4, 9, 200, 4), (6, 5, 130, 3)]:
        ds, lo, hi, pts_per_dim, beta = workloads.small(d=d, pts_per_dim=ppd, n=n, seed=7 + d, G=G)
        r = _fantasy_tc_case(engine, oracle, dict(ds), lo, hi, pts_per_dim, beta, "any", variant, precision)
        print(f"fantasy {precision} v{variant} synthetic d={d} n={n}: {r}")
        if precision == "tf32x3":      # unrefined split mode
            assert r["diff_vs_fp64"] <= max(2, r["newly_safe_fp64"] // 1000)


def test_user_mask_argreductions(engine, oracle, c3):
    """sbo_set_user_mask + sbo_argreduce over an arbitrary caller-supplied mask (all four reductions), lowest index on
    ties, -1 for an empty mask."""
    capi = _capi()
    ds = golden_ds(oracle, c3, 20)
    beta = 2.0
    grid = [67, 45]                                           # 3015 points: not a multiple of 32
    engine.set_model(ds)
    engine.set_grid(c3["lo"], c3["hi"], grid)
    m, v = engine.posterior()
    engine.sets(beta, capi.UNSAFE_ALL)                        # installs beta for the lcb/ucb reductions
    pts = oracle.make_grid(c3["lo"], c3["hi"], grid)
    lcb, ucb = oracle.bounds(m, v, beta)
    rng = np.random.default_rng(5)
    for mask in (rng.random(pts.shape[0]) < 0.3, np.arange(pts.shape[0]) % 97 == 5, np.zeros(pts.shape[0], bool)):
        engine.set_user_mask(mask)
        assert np.array_equal(engine.mask(capi.MASK_USER), mask)
        for kind, vals, want_max in ((capi.ARGMAX_VAR0, v[:, 0], True), (capi.ARGMIN_LCB0, lcb[:, 0], False),
                                     (capi.ARGMIN_UCB0, ucb[:, 0], False)):
            idx, val = engine.argreduce(kind, capi.MASK_USER)
            io, vo = (oracle.masked_argmax if want_max else oracle.masked_argmin)(vals, mask)
            assert idx == io
            if io >= 0:
                assert val == pytest.approx(vo, rel=1e-14)
        t = np.array([5.5, 81.0])
        idx, dist = engine.argreduce(capi.ARGMIN_DIST, capi.MASK_USER, 0, t)
        io, do = oracle.explore_safeset(pts, mask, t)
        assert idx == io
        if io >= 0:
            assert dist == pytest.approx(do, rel=1e-14)
