"""Pins the oracle (CPU) and the CUDA path (GPU) against vectors produced by the REFERENCE'S OWN SOURCE.

tests/golden/ref_c1_benoit.npz and ref_c3_wor.npz were written by tests/golden/make_reference_vectors.py, which
imports the unmodified /root/reference/models/{GP_Safe,SafeOpt,GoOSE}.py over a NumPy-backed `jax` stand-in
(tests/golden/refshim/README.md) and drives them as the reference's scripts do: the reference's own DE hyper-fit,
`GP_inference`, `lcb/ucb`, `infnorm_mean_grad`, the 400x400 plot mask of `create_data_for_plot`
(test/test_SafeOpt.py:324-338) and the DE-based `Minimizer/Expander/Target/...`.

Tolerances
  * posterior vs the reference's single-point `GP_inference`: |d mean| <= 1e-9 * max(max|mean|, Y_std),
    |d var| <= 1e-9 * sf2 * Y_std^2 for the oracle's inverse form (same formula; only BLAS batching differs) and
    2e-8 for the trsm/Cholesky form (oracle Cholesky form and the CUDA path) -- cond(K) reaches ~1e7.
  * plot mask: identical except points with |lcb_1| <= 1e-8 * Y_std_1.
  * DE results are continuous-domain optima found by a stochastic search, so the grid path is compared by
    containment: a grid optimum may beat the DE value (DE stuck in a local optimum) but may fall short of it only
    by DE_SLACK (grid resolution), and the DE argument must be feasible under the grid path's own bounds.
"""
import numpy as np
import pytest

from conftest import load_golden

DE_SLACK = 0.05
CPU_SIDE = 160
CPU_SLACK = DE_SLACK * 400 / CPU_SIDE      # the slack is a grid-resolution allowance
FD_RTOL, FD_ATOL = 1e-5, 1e-7      # the stand-in's jax.grad is a 4th-order central difference, not autodiff
CASES = [("ref_c1_benoit", [4, 9, 14]), ("ref_c3_wor", [5, 20, 35])]
CASE_SIZES = [(name, n) for name, sizes in CASES for n in sizes]


@pytest.fixture(scope="module")
def ref():
    return {name: load_golden(name) for name, _ in CASES}


def ref_ds(r, n, with_inverse=True):
    ds = {k: r[f"{k}_{n}"] for k in ("X_mean", "X_std", "Y_mean", "Y_std", "X_norm", "Y_norm", "hypopt")}
    if with_inverse:
        ds["invKopt"] = list(r[f"invKopt_{n}"])
    return ds


def post_err(oracle, ds, m, v, mo, vo):
    d = ds["X_norm"].shape[1]
    em = ev = 0.0
    for i in range(mo.shape[1]):
        _, sf2, _ = oracle.unpack_hyper(ds["hypopt"][:, i], d)
        em = max(em, np.max(np.abs(m[:, i] - mo[:, i])) / max(np.max(np.abs(mo[:, i])), ds["Y_std"][i]))
        ev = max(ev, np.max(np.abs(v[:, i] - vo[:, i])) / (sf2 * ds["Y_std"][i] ** 2))
    return em, ev


def unpack_mask(r, n):
    side = int(r[f"mask_side_{n}"])
    return side, np.unpackbits(r[f"mask_bits_{n}"], bitorder="little")[: side * side].astype(bool)


# ------------------------------------------------------------------------------------------------
# CPU: the oracle vs the reference's own outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n", CASE_SIZES)
def test_oracle_model_state_matches_reference(oracle, ref, name, n):
    r = ref[name]
    X, Y = r["X"][:n], r["Y"][:n]
    xm, xs, ym, ys, Xn, Yn = oracle.normalize(X, Y)                       # GP_Safe.py:84-96
    for got, key in ((xm, "X_mean"), (xs, "X_std"), (ym, "Y_mean"), (ys, "Y_std"), (Xn, "X_norm"), (Yn, "Y_norm")):
        np.testing.assert_allclose(got, r[f"{key}_{n}"], rtol=1e-13, atol=1e-13)
    ds = oracle.make_inference_datasets(X, Y, r[f"hypopt_{n}"])          # GP_Safe.py:226-245
    for i in range(Y.shape[1]):
        iK = r[f"invKopt_{n}"][i]
        assert np.max(np.abs(ds["invKopt"][i] - iK)) <= 1e-9 * np.max(np.abs(iK))
        K = oracle.build_K(ds["X_norm"], ds["hypopt"][:, i])
        assert np.max(np.abs(K @ iK - np.eye(n))) <= 1e-6


@pytest.mark.parametrize("name,n", CASE_SIZES)
def test_oracle_posterior_matches_reference(oracle, ref, name, n):
    r = ref[name]
    ds = ref_ds(r, n)
    beta = float(r["beta"])
    for pts, mk, vk in ((r[f"pts_{n}"], f"mean_{n}", f"var_{n}"),):
        m, v = oracle.posterior_inv(pts, ds)
        em, ev = post_err(oracle, ds, m, v, r[mk], r[vk])
        assert em <= 1e-9 and ev <= 1e-9, (em, ev)
        mc, vc = oracle.posterior_chol(pts, ds)
        em, ev = post_err(oracle, ds, mc, vc, r[mk], r[vk])
        assert em <= 2e-8 and ev <= 2e-8, (em, ev)
        lcb, ucb = oracle.bounds(m, v, beta)                               # SafeOpt.py:34-45
        scale = np.maximum(np.max(np.abs(r[mk]), axis=0), ds["Y_std"])
        assert np.all(np.abs(lcb - r[f"lcb_{n}"]) <= 1e-6 * scale)         # sqrt amplifies d var near var = 0
        assert np.all(np.abs(ucb - r[f"ucb_{n}"]) <= 1e-6 * scale)
    # single-point entry (GP_Safe.py:310-352) at the points the reference's scripts print at
    for p, mo, vo in zip(r["test_points"], r[f"mean_{n}"], r[f"var_{n}"]):
        m1, v1 = oracle.gp_inference(p, ds)
        np.testing.assert_allclose(m1, mo, rtol=1e-9, atol=1e-9 * np.max(ds["Y_std"]))
        np.testing.assert_allclose(v1, vo, rtol=1e-6, atol=1e-9 * np.max(ds["Y_std"]) ** 2)


@pytest.mark.parametrize("name,n", CASE_SIZES)
def test_oracle_gradient_matches_reference(oracle, ref, name, n):
    # SafeOpt.py:68-71 (jax.grad in the reference; 4th-order central differences in the stand-in)
    r = ref[name]
    ds = ref_ds(r, n)
    pts = r[f"pts_{n}"][:8]
    for i in range(ds["Y_norm"].shape[1]):
        g = np.max(np.abs(oracle.mean_grad(pts, ds, i)), axis=1)
        want = r[f"gradinf_{n}"][:, i]
        assert np.all(np.abs(g - want) <= FD_RTOL * np.abs(want) + FD_ATOL * ds["Y_std"][i] / np.min(ds["X_std"])), (g, want)


@pytest.mark.parametrize("name,n", CASE_SIZES)
def test_oracle_plot_mask_and_grid_nodes_match_reference(oracle, ref, name, n):
    r = ref[name]
    ds = ref_ds(r, n)
    side, want = unpack_mask(r, n)
    P = oracle.make_grid(r["bound"][:, 0], r["bound"][:, 1], [side, side])
    m, v = oracle.posterior_inv(P, ds)
    sel = r[f"grid_idx_{n}"]
    em, ev = post_err(oracle, ds, m[sel], v[sel], r[f"grid_mean_{n}"], r[f"grid_var_{n}"])
    assert em <= 1e-9 and ev <= 1e-9, (em, ev)
    lcb, _ = oracle.bounds(m, v, float(r["beta"]))
    got = oracle.safe_mask(lcb[:, :2], strict=True)                        # test_SafeOpt.py:337-338: lcb_1 > 0.
    bad = got != want
    assert np.all(np.abs(lcb[bad, 1]) <= 1e-8 * ds["Y_std"][1]), int(bad.sum())
    assert bad.sum() <= 4


def _feasible(oracle, ds, beta, x, tol=1e-7):
    m, v = oracle.posterior_inv(np.asarray(x)[None, :], ds)
    lcb, ucb = oracle.bounds(m, v, beta)
    return bool(np.all(lcb[0, 1:] >= -tol * ds["Y_std"][1:])), lcb[0], ucb[0], v[0]


def _check_safeopt_containment(oracle, r, n, ds, got, slack=DE_SLACK):
    """got: dict with min_ucb0, minimizer_std, expander_std, L (per GP), produced on the 400x400 grid."""
    beta = float(r["beta"])
    G = ds["Y_norm"].shape[1]
    # Minimizer (SafeOpt.py:53-66)
    x_ref, s_ref = r[f"safeopt_minimizer_x_{n}"], float(r[f"safeopt_minimizer_std_{n}"])
    ok, lcb, ucb, var = _feasible(oracle, ds, beta, x_ref)
    assert ok, "the reference's minimiser is not in the safe set under the oracle's bounds"
    assert np.sqrt(var[0]) == pytest.approx(s_ref, rel=1e-6)
    assert lcb[0] <= got["min_ucb0"] + slack * ds["Y_std"][0]           # x_ref is in M up to the grid's min ucb_0
    assert got["minimizer_std"] >= s_ref * (1.0 - slack)
    # Lipschitz constant of the last constraint (the one SafeOpt.py:110 uses) and of every other GP
    for i in range(1, G):
        assert got["L"][i] == pytest.approx(float(r[f"L_{n}"][i]), rel=0.02)
    # Expander (SafeOpt.py:90-124)
    x_ref, s_ref = r[f"safeopt_expander_x_{n}"], float(r[f"safeopt_expander_std_{n}"])
    ok, lcb, ucb, var = _feasible(oracle, ds, beta, x_ref)
    assert ok
    assert np.sqrt(var[0]) == pytest.approx(s_ref, rel=1e-6)
    assert got["expander_std"] >= s_ref * (1.0 - slack)


def _check_goose_containment(oracle, r, n, ds, got, slack=DE_SLACK):
    beta = float(r["beta"])
    scale = ds["Y_std"][0]
    # minimize_obj_lcb (GoOSE.py:63-67): a minimum over S; DE minimises over the continuum
    ok, lcb, _, _ = _feasible(oracle, ds, beta, r[f"goose_safe_min_x_{n}"])
    assert ok and lcb[0] == pytest.approx(float(r[f"goose_safe_min_lcb_{n}"]), rel=1e-6, abs=1e-9)
    assert got["safe_min_lcb"] <= float(r[f"goose_safe_min_lcb_{n}"]) + slack * scale
    # Target (GoOSE.py:80-114): z is NOT safe (max_i lcb_i(z) <= 0) and has the lowest lcb_0 among reachable z
    z_ref = r[f"goose_target_z_{n}"]
    m, v = oracle.posterior_inv(z_ref[None, :], ds)
    lz, _ = oracle.bounds(m, v, beta)
    assert np.max(lz[0, 1:]) <= 1e-7 * np.max(ds["Y_std"][1:])
    assert lz[0, 0] == pytest.approx(float(r[f"goose_target_lcb_{n}"]), rel=1e-6, abs=1e-9)
    assert got["target_lcb"] <= float(r[f"goose_target_lcb_{n}"]) + slack * scale
    # explore_safeset (GoOSE.py:116-119): the safe point nearest to the target
    ok, _, _, _ = _feasible(oracle, ds, beta, r[f"goose_explore_x_{n}"])
    assert ok


@pytest.mark.parametrize("name,n", CASE_SIZES)
def test_oracle_grid_step_contains_reference_de_results(oracle, ref, name, n):
    r = ref[name]
    ds = ref_ds(r, n)
    beta = float(r["beta"])
    # the CPU oracle does the all-pairs Lipschitz test in NumPy: 160x160 keeps it to seconds (the GPU test below runs
    # the reference's full 400x400 grid)
    P = oracle.make_grid(r["bound"][:, 0], r["bound"][:, 1], [CPU_SIDE, CPU_SIDE])
    st = oracle.safeopt_step(P, ds, beta)
    G = ds["Y_norm"].shape[1]
    L = [0.0] + [oracle.lipschitz_constant(P, ds, i) for i in range(1, G)]
    _check_safeopt_containment(oracle, r, n, ds, {"min_ucb0": st["min_ucb0"], "minimizer_std": st["minimizer_std"],
                                                   "expander_std": st["expander_std"], "L": L}, slack=CPU_SLACK)
    gs = oracle.goose_step(P, ds, beta)
    _check_goose_containment(oracle, r, n, ds, {"safe_min_lcb": gs["safe_min_lcb"], "target_lcb": gs["target_lcb"]},
                             slack=CPU_SLACK)


@pytest.mark.parametrize("name,n,seed", [("ref_c1_benoit", 4, 20260004), ("ref_c1_benoit", 14, 20260014),
                                         ("ref_c3_wor", 20, 20260520)])
def test_host_mirror_fit_reproduces_reference_fit(ref, name, n, seed):
    """The drop-in GP class keeps the hyper-parameter fit on the host (GP_Safe.py:194-234): same SciPy DE call, so
    under the RNG seed the generator used it lands on the reference's own optimum (NLL rounding differs slightly)."""
    import sbo_b200  # noqa: F401
    from sbo_b200.models.GP_Safe import GP
    r = ref[name]
    gp = GP([None] * r["Y"].shape[1])
    np.random.seed(seed)
    gp.GP_initialization(r["X"][:n], r["Y"][:n], 'RBF', multi_hyper=5, var_out=True)
    assert np.max(np.abs(gp.hypopt - r[f"hypopt_{n}"])) <= 2e-3
    for k in ("X_mean", "X_std", "Y_mean", "Y_std", "X_norm", "Y_norm"):
        np.testing.assert_allclose(gp.inference_datasets[k], r[f"{k}_{n}"], rtol=1e-13, atol=1e-13)
    with pytest.raises(ValueError):
        gp.GP_initialization(r["X"][:n], r["Y"][:n], 'Matern', multi_hyper=5)        # GP_Safe.py:136-137


# ------------------------------------------------------------------------------------------------
# GPU: the CUDA path (through the C ABI) vs the reference's own outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,n", CASE_SIZES)
def test_cuda_posterior_matches_reference(engine, oracle, ref, name, n):
    r = ref[name]
    ds = ref_ds(r, n, with_inverse=False)          # the device factorises K itself; invKopt is never uploaded
    engine.set_model(ds)
    m, v = engine.point_posterior(r[f"pts_{n}"])
    em, ev = post_err(oracle, ds, m, v, r[f"mean_{n}"], r[f"var_{n}"])
    assert em <= 2e-8 and ev <= 2e-8, (em, ev)
    for i in range(ds["Y_norm"].shape[1]):
        g = np.max(np.abs(engine.point_mean_grad(r[f"pts_{n}"][:8], i)), axis=1)
        want = r[f"gradinf_{n}"][:, i]
        assert np.all(np.abs(g - want) <= FD_RTOL * np.abs(want) + FD_ATOL * ds["Y_std"][i] / np.min(ds["X_std"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", CASE_SIZES)
def test_cuda_plot_mask_and_step_vs_reference(engine, oracle, ref, name, n):
    from sbo_b200 import _capi as capi
    r = ref[name]
    ds = ref_ds(r, n, with_inverse=False)
    beta = float(r["beta"])
    side, want = unpack_mask(r, n)
    engine.set_model(ds)
    engine.set_grid(r["bound"][:, 0], r["bound"][:, 1], [side, side])
    m, v = engine.posterior(with_grad=True)
    sel = r[f"grid_idx_{n}"]
    em, ev = post_err(oracle, ds, m[sel], v[sel], r[f"grid_mean_{n}"], r[f"grid_var_{n}"])
    assert em <= 2e-8 and ev <= 2e-8, (em, ev)
    # the plot mask (test_SafeOpt.py:337-338) is the strict safe set of constraint 1; with G = 2 that is S itself
    lcb1 = m[:, 1] - beta * np.sqrt(v[:, 1])
    if ds["Y_norm"].shape[1] == 2:
        engine.sets(beta, capi.UNSAFE_ALL, strict=True)
        got = engine.mask(capi.MASK_SAFE)
    else:
        got = lcb1 > 0.0
    bad = got != want
    assert np.all(np.abs(lcb1[bad]) <= 1e-7 * ds["Y_std"][1]), int(bad.sum())
    assert bad.sum() <= 8
    # whole steps on the reference's 400x400 grid vs its DE results (containment)
    st = engine.safeopt_step(ds, beta)
    _check_safeopt_containment(oracle, r, n, dict(ds, invKopt=list(r[f"invKopt_{n}"])),
                               {"min_ucb0": st["min_ucb0"], "minimizer_std": st["minimizer_std"],
                                "expander_std": st["expander_std"], "L": st["L"]})
    gs = engine.goose_step(ds, beta)
    _check_goose_containment(oracle, r, n, dict(ds, invKopt=list(r[f"invKopt_{n}"])),
                             {"safe_min_lcb": gs["min_lcb0"], "target_lcb": gs["target_lcb"]})


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("ref_c1_benoit", 9), ("ref_c3_wor", 20)])
def test_dropin_classes_vs_reference(oracle, ref, name, n):
    """The reference's call sequence (test/test_SafeOpt.py:21-33,144-158; test/test_GoOSE.py:151-162) through the
    drop-in classes, on the model state the reference's own fit produced."""
    import sbo_b200  # noqa: F401
    from sbo_b200.models import GoOSE, SafeOpt
    r = ref[name]
    G = r["Y"].shape[1]
    plant = [lambda x, noise=0, j=j: 0.0 for j in range(G)]
    beta = float(r["beta"])
    ds = ref_ds(r, n)
    bo = SafeOpt.BO(plant, r["bound"], beta)
    bo.GP_initialization(r["X"][:n], r["Y"][:n], 'RBF', multi_hyper=5, var_out=True, hypopt=r[f"hypopt_{n}"])
    for p, mo, vo, lo_, uo in zip(r[f"pts_{n}"][:6], r[f"mean_{n}"], r[f"var_{n}"], r[f"lcb_{n}"], r[f"ucb_{n}"]):
        m, v = bo.GP_inference(p, bo.inference_datasets)
        scale = np.maximum(np.abs(mo), ds["Y_std"])
        assert np.all(np.abs(m - mo) <= 2e-8 * scale)
        for i in range(G):
            assert abs(bo.lcb(p, i) - lo_[i]) <= 1e-5 * scale[i] and abs(bo.ucb(p, i) - uo[i]) <= 1e-5 * scale[i]
    x_min, std_min = bo.Minimizer()
    x_exp, std_exp = bo.Expander()
    _, min_ucb = bo.minimize_obj_ucb(None)
    L = [0.0] + [bo.maximize_infnorm_mean_grad(i) for i in range(1, G)]
    _check_safeopt_containment(oracle, r, n, ds, {"min_ucb0": min_ucb, "minimizer_std": std_min, "expander_std": std_exp, "L": L})
    go = GoOSE.BO(plant, r["bound"], beta)
    go.GP_initialization(r["X"][:n], r["Y"][:n], 'RBF', multi_hyper=5, var_out=True, hypopt=r[f"hypopt_{n}"])
    _, lcb_min = go.minimize_obj_lcb()
    z, lcb_t = go.Target()
    x_e = go.explore_safeset(z)
    _check_goose_containment(oracle, r, n, ds, {"safe_min_lcb": lcb_min, "target_lcb": lcb_t})
    assert all(bo.lcb(x_e, i) >= 0 for i in range(1, G))                      # the exploration point is safe


# ------------------------------------------------------------------------------------------------
# CPU: the pair constraints of Expander()/Target() (SafeOpt.py:73-77,85-88,99-111; GoOSE.py:93-101)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n", [("ref_c1_benoit", 9), ("ref_c1_benoit", 14), ("ref_c3_wor", 20), ("ref_c3_wor", 35)])
def test_oracle_pair_constraints_match_reference(oracle, ref, name, n):
    """tests/golden/ref_pairs.npz: values of the reference's own constraint lambdas at random (x, z) pairs.
    The oracle's safe / unsafe masks and its Lipschitz reach predicate must take the same decisions (pairs whose
    reference value is within 1e-9 of the threshold are exempt)."""
    rp = load_golden("ref_pairs")
    r = ref[name]
    ds = ref_ds(r, n)
    beta = float(r["beta"])
    key = f"{name}_{n}"
    x, z, L = rp[key + "_x"], rp[key + "_z"], rp[key + "_L"]
    G = ds["Y_norm"].shape[1]
    mx, vx = oracle.posterior_inv(x, ds)
    lcb_x, ucb_x = oracle.bounds(mx, vx, beta)
    scale = np.maximum(np.max(np.abs(mx), axis=0), ds["Y_std"])
    assert np.all(np.abs(lcb_x - rp[key + "_lcb_x"]) <= 1e-6 * scale)
    clear = np.all(np.abs(rp[key + "_lcb_x"][:, 1:]) > 1e-9 * scale[1:], axis=1)
    assert np.array_equal(oracle.safe_mask(lcb_x)[clear], np.all(rp[key + "_lcb_x"][:, 1:] >= 0, axis=1)[clear])
    mz, vz = oracle.posterior_inv(z, ds)
    lcb_z, _ = oracle.bounds(mz, vz, beta)
    ref_max = rp[key + "_lcbmax_z"]                                   # lcb_constraint_min returns the MAX (sic)
    assert np.all(np.abs(np.max(lcb_z[:, 1:], axis=1) - ref_max) <= 1e-6 * np.max(scale[1:]))
    clear = np.abs(ref_max) > 1e-9 * np.max(scale[1:])
    assert np.array_equal(oracle.unsafe_mask(lcb_z, "all")[clear], (ref_max <= 0)[clear])
    # Lipschitz reach predicate, pair (x_j, z_j), constraint index idx, L of the LAST constraint (SafeOpt.py:110)
    for c in range(G - 1):
        reach = np.concatenate([blk for _, blk in oracle.pair_reach(x, ucb_x[:, c + 1], z, L[G - 1])], axis=0)
        got = np.diag(reach)
        want = rp[key + "_lip"][:, c]
        clear = np.abs(want) > 1e-9 * scale[c + 1]
        assert np.array_equal(got[clear], (want >= 0)[clear]), (c, int((got[clear] != (want >= 0)[clear]).sum()))
        # and the value itself: ucb_idx(x) - L * ||x - z + 1e-8||
        val = ucb_x[:, c + 1] - L[G - 1] * np.sqrt(np.sum((x - z + oracle.PAIR_OFFSET) ** 2, axis=1))
        assert np.all(np.abs(val - want) <= 1e-6 * scale[c + 1])
