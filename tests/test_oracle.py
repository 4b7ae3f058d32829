"""CPU tests of the oracle itself: golden fixtures, internal consistency and the properties the
reference's scripts print (last bullet)."""
import numpy as np
import pytest

from conftest import golden_ds


@pytest.mark.parametrize("name,sizes", [("c1", [4, 9, 14]), ("c3", [5, 20, 35])])
def test_golden_point_posterior(oracle, request, name, sizes):
    gold = request.getfixturevalue(name)
    for n in sizes:
        ds = golden_ds(oracle, gold, n)
        m, v = oracle.posterior_inv(gold["test_points"], ds)
        np.testing.assert_allclose(m, gold[f"tp_mean_{n}"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(v, gold[f"tp_var_{n}"], rtol=1e-7, atol=1e-10)


def test_inv_form_vs_chol_form(oracle, c1, c3):
    for gold, n in [(c1, 14), (c3, 35)]:
        ds = golden_ds(oracle, gold, n)
        pts = oracle.make_grid(gold["lo"], gold["hi"], [60, 60])
        ma, va = oracle.posterior_inv(pts, ds)
        mb, vb = oracle.posterior_chol(pts, ds)
        G = ma.shape[1]
        for i in range(G):
            _, sf2, _ = oracle.unpack_hyper(ds["hypopt"][:, i], 2)
            assert np.max(np.abs(ma[:, i] - mb[:, i])) <= 2e-8 * max(np.max(np.abs(ma[:, i])), ds["Y_std"][i])
            assert np.max(np.abs(va[:, i] - vb[:, i])) <= 2e-8 * sf2 * ds["Y_std"][i] ** 2


def test_interpolation_at_training_point(oracle, c1):
    # test/test_GP_Safe.py:33-41: posterior at a training input ~ observation, small variance
    ds = golden_ds(oracle, c1, 9)
    m, v = oracle.posterior_inv(c1["X"][:9], ds)
    assert np.max(np.abs(m - c1["Y"][:9])) < 0.05
    assert np.all(v >= 0) and np.max(v) < 1e-2


def test_constraint_prior_far_from_data(oracle, c1):
    # test/test_GP_Safe.py:43-47, GP_Safe.py:331: far away the constraint mean -> -Y_mean < 0
    ds = golden_ds(oracle, c1, 9)
    m, v = oracle.gp_inference(np.array([10.0, 10.0]), ds)
    assert m[1] == pytest.approx(-ds["Y_mean"][1], rel=1e-6)
    assert m[0] == pytest.approx(ds["Y_mean"][0], rel=1e-6)


def test_gradient_vs_central_differences(oracle, c3):
    ds = golden_ds(oracle, c3, 20)
    x = np.array([[6.0, 85.0], [5.1, 78.0]])
    for i in range(3):
        g = oracle.mean_grad(x, ds, i)
        for k in range(2):
            h = 1e-5 * (1.0 if k == 0 else 10.0)
            e = np.zeros(2)
            e[k] = h
            fd = (oracle.posterior_inv(x + e, ds)[0][:, i] - oracle.posterior_inv(x - e, ds)[0][:, i]) / (2 * h)
            np.testing.assert_allclose(g[:, k], fd, rtol=2e-5, atol=1e-7)


def test_grid_order_x0_fastest(oracle):
    pts = oracle.make_grid([-0.6, -1.0], [1.5, 1.0], [400, 400])
    x0, x1 = np.linspace(-0.6, 1.5, 400), np.linspace(-1.0, 1.0, 400)
    X0, X1 = np.meshgrid(x0, x1)
    assert np.array_equal(pts, np.column_stack((X0.ravel(), X1.ravel())))   # test_SafeOpt.py:324-334
    p3 = oracle.make_grid([0, 0, 0], [1, 2, 3], [3, 4, 5])
    assert p3.shape == (60, 3) and p3[1, 0] == 0.5 and p3[1, 1] == 0 and p3[3, 1] == pytest.approx(2 / 3)


def test_rank1_fantasy_vs_augmented_inference(oracle, c3):
    ds = golden_ds(oracle, c3, 20)
    beta = 2.0
    pts = oracle.make_grid(c3["lo"], c3["hi"], [14, 14])
    mean, var = oracle.posterior_inv(pts, ds)
    lcb, _ = oracle.bounds(mean, var, beta)
    S, Z = oracle.safe_mask(lcb), oracle.unsafe_mask(lcb, "any")
    xs, zs = np.flatnonzero(S), np.flatnonzero(Z)
    assert xs.size and zs.size
    margin = oracle.fantasy_margin(pts, ds, beta, S, Z)            # (|Z|,|S|) normalised units
    for j in [0, xs.size // 2, xs.size - 1]:
        lcb_aug = oracle.fantasy_by_augmentation(pts[xs[j]], pts[zs], ds, beta)[:, 1:]
        got = margin[:, j]
        want = np.min(lcb_aug / ds["Y_std"][1:], axis=1)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)


def test_steps_on_small_grid(oracle, c1):
    ds = golden_ds(oracle, c1, 9)
    pts = oracle.make_grid(c1["lo"], c1["hi"], [40, 40])
    st = oracle.safeopt_step(pts, ds, 3.0)
    assert st["S"].sum() > 0 and st["M"].sum() > 0 and (st["M"] & ~st["S"]).sum() == 0
    assert st["expander_masks"].shape == (1, 1600) and (st["expander_masks"][0] & ~st["S"]).sum() == 0
    assert st["x_new_idx"] in (st["minimizer_idx"], st["expander_idx"])
    gs = oracle.goose_step(pts, ds, 3.0)
    assert (gs["target_masks"][0] & ~gs["Z"]).sum() == 0
    assert gs["x_new_idx"] >= 0 and st["S"][gs["x_new_idx"]]
    fz = oracle.safeopt_step(pts, ds, 3.0, mode="fantasy")
    assert fz["counts"].shape == (1600,) and (fz["counts"][~fz["S"]] == 0).all()


def test_empty_sets(oracle, c1):
    ds = golden_ds(oracle, c1, 4)
    pts = oracle.make_grid([5.0, 5.0], [6.0, 6.0], [8, 8])      # far from the data: nothing is safe
    st = oracle.safeopt_step(pts, ds, 3.0)
    assert st["S"].sum() == 0 and st["minimizer_idx"] == -1 and st["expander_idx"] == -1 and st["x_new_idx"] == -1
    gs = oracle.goose_step(pts, ds, 3.0)
    assert gs["safe_min_idx"] == -1 and gs["target_idx"] == -1


# ------------------------------------------------------------------------------------------------
# size-independent properties of the set definitions
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,beta", [("c1", 9, 3.0), ("c3", 20, 2.0), ("c3", 35, 2.0)])
def test_set_algebra_and_monotonicity(oracle, request, name, n, beta):
    gold = request.getfixturevalue(name)
    ds = golden_ds(oracle, gold, n)
    pts = oracle.make_grid(gold["lo"], gold["hi"], [48, 40])
    m, v = oracle.posterior_inv(pts, ds)
    G = m.shape[1]
    lcb, ucb = oracle.bounds(m, v, beta)
    S, Z_all, Z_any = oracle.safe_mask(lcb), oracle.unsafe_mask(lcb, "all"), oracle.unsafe_mask(lcb, "any")
    assert not np.any(S & Z_any) and np.array_equal(Z_any, ~S)            # 'any' is the complement of S
    assert np.all(Z_all <= (Z_any | np.all(lcb[:, 1:] == 0.0, axis=1)))    # 'all' is contained in it (up to lcb == 0)
    _, _, M, min_ucb = oracle.minimizer(v, lcb, ucb, S)
    assert np.all(M <= S) and (not S.any() or np.all(lcb[M, 0] <= min_ucb))
    # a larger beta can only shrink the safe set (lcb decreases pointwise)
    lcb2, ucb2 = oracle.bounds(m, v, beta * 1.5)
    assert np.all(oracle.safe_mask(lcb2) <= S)
    assert np.all(lcb2 <= lcb + 1e-15) and np.all(ucb2 >= ucb - 1e-15)
    if not (S.any() and Z_all.any()):
        return
    # a larger Lipschitz constant shrinks every reach radius: expander and target sets can only shrink
    L1 = np.full(G, max(1e-3, oracle.lipschitz_constant(pts, ds, G - 1)))
    e1 = oracle.expander_lipschitz(pts, S, Z_all, ucb, v, L1)
    e2 = oracle.expander_lipschitz(pts, S, Z_all, ucb, v, L1 * 3.0)
    t1 = oracle.goose_target(pts, S, Z_all, ucb, lcb, L1)
    t2 = oracle.goose_target(pts, S, Z_all, ucb, lcb, L1 * 3.0)
    for c in range(G - 1):
        assert np.all(e2["masks"][c] <= e1["masks"][c]) and np.all(e1["masks"][c] <= S)
        assert np.all(t2["masks"][c] <= t1["masks"][c]) and np.all(t1["masks"][c] <= Z_all)
    # an expander exists for constraint c iff a target exists for it (both say: some (x, z) pair passes the test)
    for c in range(G - 1):
        assert e1["masks"][c].any() == t1["masks"][c].any()
    # the chosen expander is the most uncertain member of the union, lowest index on ties
    if e1["best_idx"] >= 0:
        union = np.any(e1["masks"], axis=0)
        assert v[e1["best_idx"], 0] == v[union, 0].max()
        assert e1["best_idx"] == np.flatnonzero(union & (v[:, 0] == v[union, 0].max()))[0]


def test_arg_reductions_break_ties_towards_the_lowest_index(oracle):
    vals = np.array([3.0, 1.0, 1.0, 5.0, 5.0, 1.0])
    mask = np.array([True, False, True, True, True, True])
    assert oracle.masked_argmin(vals, mask) == (2, 1.0)
    assert oracle.masked_argmax(vals, mask) == (3, 5.0)
    assert oracle.masked_argmin(vals, np.zeros(6, bool))[0] == -1 and oracle.masked_argmax(vals, np.zeros(6, bool))[0] == -1
