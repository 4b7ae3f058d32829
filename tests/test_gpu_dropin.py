"""GPU: the drop-in classes (reference names and signatures) drive the CUDA grid pipeline and return what the
reference's drivers expect (test/test_SafeOpt.py:135-186, test/test_GoOSE.py:142-190), checked against the
oracle restated on the same grid."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_ds

pytestmark = pytest.mark.gpu


def _models():
    # reference-style import: the package directory on sys.path, then `from models import SafeOpt`
    pkg = os.path.join(ROOT, "safe-bayesian-optimization_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from models import SafeOpt, GoOSE, GP_Safe
    from utils import utils_SafeOpt, utils_GoOSE
    from problems import Benoit_Problem
    return SafeOpt, GoOSE, GP_Safe, utils_SafeOpt, utils_GoOSE, Benoit_Problem


def test_gp_inference_single_point(oracle, c1):
    SafeOpt, GoOSE, GP_Safe, _, _, B = _models()
    gp = GP_Safe.GP([B.Benoit_System_1, B.con1_system_tight])
    gp.GP_initialization(c1["X"][:9], c1["Y"][:9], 'RBF', multi_hyper=5, var_out=True, hypopt=c1["hyp_9"])
    x = np.array([1.45698204, -0.76514894])                       # test/test_SafeOpt.py:47
    mean, var = gp.GP_inference(x, gp.inference_datasets)
    assert mean.shape == (2,) and var.shape == (2,) and mean.dtype == np.float64
    mo, vo = oracle.gp_inference(x, golden_ds(oracle, c1, 9))
    np.testing.assert_allclose(mean, mo, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, vo, rtol=1e-5, atol=1e-10)
    gp.var_out = False
    assert gp.GP_inference(x, gp.inference_datasets) == pytest.approx(mo[0], rel=1e-8)
    # add_sample re-normalises and the next inference sees the new model (GP_Safe.py:283-304)
    gp.var_out = True
    gp.add_sample(c1["X"][9], c1["Y"][9], hypopt=c1["hyp_9"])
    assert gp.n_point == 10 and gp.X_norm.shape == (10, 2) and len(gp.invKopt) == 2
    m2, v2 = gp.GP_inference(c1["X"][9], gp.inference_datasets)
    assert abs(m2[0] - c1["Y"][9, 0]) < 0.05 and v2[0] < 1e-2


@pytest.mark.parametrize("n", [4, 9, 14])
def test_safeopt_bo_api(oracle, c1, n):
    SafeOpt, _, _, utils_SafeOpt, _, B = _models()
    plant = [B.Benoit_System_1, B.con1_system_tight]
    bound = np.array([[-.6, 1.5], [-1., 1.]])
    bo = SafeOpt.BO(plant, bound, 3., grid_points_per_dim=120)
    bo.GP_initialization(c1["X"][:n], c1["Y"][:n], 'RBF', multi_hyper=5, var_out=True, hypopt=c1[f"hyp_{n}"])
    ds = golden_ds(oracle, c1, n)
    pts = oracle.make_grid(bound[:, 0], bound[:, 1], [120, 120])
    so = oracle.safeopt_step(pts, ds, 3.0, form="chol")
    x_min, std_min = bo.Minimizer()
    x_exp, std_exp = bo.Expander()
    assert x_min.shape == (2,) and x_exp.shape == (2,)
    np.testing.assert_array_equal(x_min, pts[so["minimizer_idx"]])
    assert std_min == pytest.approx(so["minimizer_std"], rel=1e-9)
    np.testing.assert_array_equal(x_exp, pts[so["expander_idx"]])
    assert std_exp == pytest.approx(so["expander_std"], rel=1e-9)
    x_u, f_u = bo.minimize_obj_ucb(None)
    assert f_u == pytest.approx(so["min_ucb0"], rel=1e-9)
    assert bo.maximize_infnorm_mean_grad(1) == pytest.approx(so["L"][1], rel=1e-9)
    # scalar helpers behave like the reference's (SafeOpt.py:29-45,73-77,85-88)
    x = np.array([0.9, -0.6])
    mo, vo = oracle.gp_inference(x, ds)
    assert bo.mean(x, 0) == pytest.approx(mo[0], rel=1e-8)
    assert bo.ucb(x, 1) == pytest.approx(mo[1] + 3.0 * np.sqrt(vo[1]), rel=1e-6, abs=1e-9)
    assert bo.lcb(x, 1) == pytest.approx(mo[1] - 3.0 * np.sqrt(vo[1]), rel=1e-6, abs=1e-9)
    assert bo.lcb_constraint_min(x) == pytest.approx(bo.lcb(x, 1))
    g = oracle.mean_grad(x[None, :], ds, 1)[0]
    assert bo.infnorm_mean_grad(x, 1) == pytest.approx(np.max(np.abs(g)), rel=1e-8)
    x2 = np.array([0.9, -0.6, 0.7, -0.2])
    want = bo.ucb(x2[:2], 1) - 2.0 * np.linalg.norm(x2[:2] - x2[2:] + 1e-8)
    assert bo.Lipschitz_continuity_constraint(x2, 1, 2.0) == pytest.approx(want)
    out = bo.calculate_plant_outputs(x)
    assert out.shape == (2,) and out[0] == pytest.approx(B.Benoit_System_1(x))
    # driver decision (test/test_SafeOpt.py:153-158) and the plot-mask producer (test_SafeOpt.py:324-345)
    x_new = x_min if std_min > std_exp else x_exp
    np.testing.assert_array_equal(x_new, pts[so["x_new_idx"]])
    X_0, X_1, mask_safe, obj = utils_SafeOpt.create_data_for_plot(bo, plant, n_grid=400)
    assert X_0.shape == (400, 400) and mask_safe.shape == (400, 400) and mask_safe.dtype == bool
    p400 = oracle.make_grid(bound[:, 0], bound[:, 1], [400, 400])
    m400, v400 = oracle.posterior_chol(p400, ds)
    want_mask = ((m400[:, 1] - 3.0 * np.sqrt(v400[:, 1])) > 0.).reshape(400, 400)
    assert (mask_safe != want_mask).sum() <= 2
    assert obj[3, 7] == pytest.approx(B.Benoit_System_1(np.array([X_0[3, 7], X_1[3, 7]])))
    # the grid is restored after the plot call
    x_min2, std_min2 = bo.Minimizer()
    np.testing.assert_array_equal(x_min2, x_min)


def test_goose_bo_api(oracle, c1):
    _, GoOSE, _, _, _, B = _models()
    plant = [B.Benoit_System_1, B.con1_system_tight]
    bound = np.array([[-.6, 1.5], [-1., 1.]])
    bo = GoOSE.BO(plant, bound, 3., grid_points_per_dim=100)
    bo.GP_initialization(c1["X"][:9], c1["Y"][:9], 'RBF', multi_hyper=5, var_out=True, hypopt=c1["hyp_9"])
    ds = golden_ds(oracle, c1, 9)
    pts = oracle.make_grid(bound[:, 0], bound[:, 1], [100, 100])
    go = oracle.goose_step(pts, ds, 3.0, form="chol")
    assert len(bo.safe_set_cons) == 1
    x_safe_min, min_safe_lcb = bo.minimize_obj_lcb()
    x_target, target_lcb = bo.Target()
    np.testing.assert_array_equal(x_safe_min, pts[go["safe_min_idx"]])
    assert min_safe_lcb == pytest.approx(go["safe_min_lcb"], rel=1e-9)
    np.testing.assert_array_equal(x_target, pts[go["target_idx"]])
    assert target_lcb == pytest.approx(go["target_lcb"], rel=1e-9)
    if min_safe_lcb <= target_lcb:                                   # test/test_GoOSE.py:158-162
        x_new = x_safe_min
    else:
        x_new = bo.explore_safeset(x_target)
    np.testing.assert_array_equal(x_new, pts[go["x_new_idx"]])
    # a full SafeOpt-style loop of 3 iterations with add_sample (fixed hypers) runs end to end
    for it in range(3):
        x_safe_min, min_safe_lcb = bo.minimize_obj_lcb()
        x_target, target_lcb = bo.Target()
        x_new = x_safe_min if min_safe_lcb <= target_lcb else bo.explore_safeset(x_target)
        y_new = bo.calculate_plant_outputs(x_new)
        assert y_new[1] >= -0.05                                     # the queried point is (nearly) safe
        bo.add_sample(x_new, y_new, hypopt=c1["hyp_9"])
    assert bo.n_point == 12


def test_wor_three_gps_fantasy_mode(oracle, c3):
    SafeOpt, _, _, _, _, _ = _models()
    from problems import WilliamOttoReactor_Problem
    wo = WilliamOttoReactor_Problem.WilliamOttoReactor()
    plant = [wo.get_objective, wo.get_constraint1, wo.get_constraint2]
    bound = np.array([[4., 7.], [70., 100.]])
    ds = golden_ds(oracle, c3, 20)
    pts = oracle.make_grid(bound[:, 0], bound[:, 1], [60, 50])
    for precision in ["fp64", "tf32x3"]:
        bo = SafeOpt.BO(plant, bound, 2., grid_points_per_dim=[60, 50], expander_mode='fantasy', precision=precision,
                        unsafe_rule='any')
        bo.GP_initialization(c3["X"][:20], c3["Y"][:20], 'RBF', multi_hyper=5, var_out=True, hypopt=c3["hyp_20"])
        x_exp, std_exp = bo.Expander()
        so = oracle.safeopt_step(pts, ds, 2.0, mode="fantasy", unsafe_rule="any", form="chol")
        if precision == "fp64":
            np.testing.assert_array_equal(x_exp, pts[so["expander_idx"]])
            assert np.array_equal(bo.safe_mask('expander'), so["expander_masks"][0])
        assert std_exp == pytest.approx(so["expander_std"], rel=1e-6)
        assert np.array_equal(bo.safe_mask('safe'), so["S"])


def test_safeopt_and_goose_runs_on_benoit_end_to_end(tmp_path):
    """BASELINE.json configs[0]/[1] as the reference's drivers run them (test/test_SafeOpt.py:21-33,135-186;
    test/test_GoOSE.py:142-190): 4 initial samples around (1.4,-0.8), hyper-fit on the host, every acquisition on the
    GPU grid pipeline.  Checks what the reference's GIFs show: the runs stay safe and walk towards the constrained
    optimum f = 0.145249 at (0.368,-0.393)."""
    from sbo_b200 import drivers
    from sbo_b200.models import GoOSE, SafeOpt
    from sbo_b200.problems import Benoit_Problem as P
    from sbo_b200.utils import utils_SafeOpt
    plant = [P.Benoit_System_1, P.con1_system_tight]
    bound = np.array([[-.6, 1.5], [-1., 1.]])
    for algo in ("safeopt", "goose"):
        bo = (SafeOpt.BO if algo == "safeopt" else GoOSE.BO)(plant, bound, 3.)
        bo.key = np.random.default_rng(3)
        bo.hyper_seed = 11
        X, Y = bo.Data_sampling(4, np.array([1.4, -.8]), 0.3, 0.)
        bo.GP_initialization(X, Y, 'RBF', multi_hyper=5, var_out=True)
        f0 = Y[:, 0].min()
        frames = []

        def on_it(i, GP_m, x_new, y, info):
            if i == 0:
                X_0, X_1, mask, obj = utils_SafeOpt.create_data_for_plot(GP_m, plant, n_grid=100)
                frames.append((mask.shape, int(mask.sum())))
        data = (drivers.run_safeopt(bo, n_iteration=10, on_iteration=on_it) if algo == "safeopt"
                else drivers.run_goose(bo, n_iteration=10, on_iteration=on_it))
        con = np.array(data["con"])
        obj = np.array(data["obj"])
        assert frames and frames[0][0] == (100, 100) and frames[0][1] > 0
        assert np.all(con >= -1e-9), (algo, con)                      # every query satisfied the true constraint
        assert obj.min() < f0 and obj.min() < 0.6, (algo, obj)       # moved from f ~ 1.4 towards the optimum 0.145
        assert bo.n_point == 4 + len(data["i"])
    runs = drivers.run_multiple(lambda: SafeOpt.BO(plant, bound, 2.), [1.4, -.8], 0.3, 4, 2, 3, 0.005,
                                path=tmp_path / "multi.npz", seeds=[5, 6])
    back = drivers.load_runs(tmp_path / "multi.npz")
    assert set(back) == {"0", "1"} and back["0"]["sampled_x"].shape == (4, 2)
    assert back["1"]["observed_output"].shape[1] == 2 and 1 <= back["1"]["observed_x"].shape[0] <= 3
    assert runs["0"]["observed_x"].shape == back["0"]["observed_x"].shape


@pytest.mark.parametrize("n,d,P", [(4, 2, 60), (14, 2, 60), (35, 2, 33), (200, 4, 16), (600, 6, 5)])
def test_batched_nll_matches_oracle(oracle, c3, n, d, P):
    """sbo_nll_batch vs GP.negative_loglikelihood (GP_Safe.py:169-192) for a population inside the fit's bounds
    (GP_Safe.py:205-206), including sizes that are not multiples of the 32-wide factorisation block."""
    import sbo_b200
    rng = np.random.default_rng(100 + n)
    if d == 2 and n <= 35:
        X, Y = c3["X"][:n], c3["Y"][:n]
        Xn, Yn = oracle.normalize(X, Y)[4:]
        y = Yn[:, 1]
    else:
        Xn = rng.normal(size=(n, d))
        y = np.sin(Xn[:, 0]) + 0.1 * rng.normal(size=n)
        y = (y - y.mean()) / y.std()
    H = np.column_stack([rng.uniform(-1.5, 1.5, size=(P, d + 1)), rng.uniform(-5., -2., size=P)])
    eng = sbo_b200.GridEngine(0)
    got = eng.nll_batch(Xn, y, H)
    want = np.array([oracle.negative_loglikelihood(h, Xn, y[:, None]) for h in H])
    eng.close()
    assert np.all(np.abs(got - want) <= 1e-9 * np.maximum(1.0, np.abs(want))), np.max(np.abs(got - want))


def test_fit_on_device_reaches_the_host_optimum(c1):
    """The DE fit with population NLLs on the GPU lands on (at least) the host fit's optimum."""
    from sbo_b200.models.GP_Safe import GP
    X, Y = c1["X"][:14], c1["Y"][:14]
    host = GP([None, None]); host.hyper_seed = 5
    host.GP_initialization(X, Y, 'RBF', multi_hyper=5)
    dev = GP([None, None]); dev.hyper_seed = 5; dev.fit_on_device = True
    dev.GP_initialization(X, Y, 'RBF', multi_hyper=5)
    for i in range(2):
        a = host.negative_loglikelihood(host.hypopt[:, i], host.X_norm, host.Y_norm[:, i:i + 1])
        b = host.negative_loglikelihood(dev.hypopt[:, i], host.X_norm, host.Y_norm[:, i:i + 1])
        assert b <= a + 1e-3 * max(1.0, abs(a)), (i, a, b)
    m, v = dev.GP_inference(X[3], dev.inference_datasets)
    assert np.all(np.abs(m - Y[3]) < 0.05)


def test_trust_region_minimize_obj_lcb(oracle, c1):
    """models/GP_TR.BO.minimize_obj_lcb(r, x_0) (reference GP_TR.py:43-51): argmin lcb_0 over the safe set within the
    ball -- the grid pipeline with one extra (user) mask -- vs the oracle on the same grid."""
    from sbo_b200.models import GP_TR
    from sbo_b200.problems import Benoit_Problem as P
    TRp = {'radius': 0.5, 'radius_max': 1, 'radius_red': 0.8, 'radius_inc': 1.1, 'rho_lb': 0.2, 'rho_ub': 0.8}
    bound = np.array([[-.6, 1.5], [-1., 1.]])
    n, beta, side = 9, 3.0, 120
    bo = GP_TR.BO([P.Benoit_System_1, P.con1_system_tight], bound, beta, TRp, grid_points_per_dim=side)
    bo.GP_initialization(c1["X"][:n], c1["Y"][:n], 'RBF', multi_hyper=5, var_out=True, hypopt=c1[f"hyp_{n}"])
    ds = golden_ds(oracle, c1, n)
    pts = oracle.make_grid(bound[:, 0], bound[:, 1], [side, side])
    m, v = bo.grid_posterior()
    lcb, _ = oracle.bounds(m, v, beta)
    S = oracle.safe_mask(lcb)
    for x_0, r in (([1.4, -0.8], 0.3), ([0.5, -0.3], 0.25), ([-0.5, 0.9], 0.2)):
        dist = np.linalg.norm(pts - np.asarray(x_0), axis=1)
        assert not np.any(np.abs(dist - r) < 1e-12)              # no grid point sits on the sphere: the ball is unambiguous
        ball = dist <= r
        io, vo = oracle.masked_argmin(lcb[:, 0], S & ball)
        x, val = bo.minimize_obj_lcb(r, np.asarray(x_0))
        if io < 0:
            assert np.isnan(x).all() and val == np.inf
        else:
            np.testing.assert_allclose(x, pts[io], rtol=0, atol=1e-14)
            assert val == pytest.approx(vo, rel=1e-14) and bo.TR_constraint(x, np.asarray(x_0), r) >= -1e-7
    # one trust-region iteration as the reference's driver does (test/test_GP_TR.py:50-57)
    x_old, r_old = np.array([1.4, -0.8]), 0.3
    y_old = bo.calculate_plant_outputs(x_old)
    x_min, _ = bo.minimize_obj_lcb(r_old, x_old)
    centre, radius = bo.update_TR(x_old, x_min, r_old, y_old, bo.calculate_plant_outputs(x_min))
    assert radius in (pytest.approx(r_old * 0.8), pytest.approx(r_old), pytest.approx(min(r_old * 1.1, 1)))
    assert np.array_equal(centre, x_old) or np.array_equal(centre, x_min)


def test_append_sample_matches_factorisation_from_scratch(engine, oracle):
    """sbo_append_sample (rank-1 update at fixed hyper-parameters and normalisation) against
    sbo_set_model of the grown data set: factor, W, alpha and the grid posterior; n crosses a 128-row padding boundary."""
    from sbo_b200 import workloads
    ds, lo, hi, pts, beta = workloads.small(d=3, pts_per_dim=10, n=132, seed=3, G=3)
    n0 = 126
    base = dict(ds)
    base["X_norm"], base["Y_norm"] = ds["X_norm"][:n0], ds["Y_norm"][:n0]
    engine.set_grid(lo, hi, pts)
    engine.set_model(base)
    for j in range(n0, 132):
        engine.append_sample(ds["X_norm"][j], ds["Y_norm"][j])
    La, Wa, aa = engine.get_model()
    ma, va = engine.posterior()
    engine.set_model(ds)
    Lb, Wb, ab = engine.get_model()
    mb, vb = engine.posterior()
    assert np.max(np.abs(La - Lb)) <= 1e-11 * np.max(np.abs(Lb))
    assert np.max(np.abs(Wa - Wb)) <= 1e-9 * np.max(np.abs(Wb))
    assert np.max(np.abs(aa - ab)) <= 1e-8 * np.max(np.abs(ab))
    assert np.max(np.abs(ma - mb)) <= 1e-9 * np.max(np.abs(mb))
    assert np.max(np.abs(va - vb)) <= 1e-9 * np.max(ds["Y_std"]) ** 2


def test_stableopt_minmax_on_the_grid(oracle):
    """StableOpt's Minimize_Maximise / Maximise_d / Minimise_d (models/StableOpt.py:96-154) on the joint (x_c, d) grid
    against the oracle posterior with the zero prior mean of GP_Robust.py:322-323."""
    from sbo_b200 import workloads
    from sbo_b200.models import StableOpt
    ds, lo, hi, pts, beta = workloads.small(d=3, pts_per_dim=10, n=60, seed=5, G=3)
    bo = StableOpt.BO([None] * 3, np.column_stack([lo[:2], hi[:2]]), np.column_stack([lo[2:], hi[2:]]), beta,
                      grid_points_per_dim=[14, 13, 9])
    bo.X_mean, bo.X_std, bo.Y_mean, bo.Y_std = ds["X_mean"], ds["X_std"], ds["Y_mean"], ds["Y_std"]
    bo.X_norm, bo.Y_norm, bo.hypopt, bo.invKopt = ds["X_norm"], ds["Y_norm"], ds["hypopt"], None
    bo.n_point, bo.nx_dim, bo.ny_dim, bo.var_out, bo.kernel = 60, 3, 3, True, "RBF"
    bo.update_inference_dataset()
    P = oracle.make_grid(lo, hi, [14, 13, 9])
    zero = dict(ds)
    orig = oracle.prior_mean
    try:
        oracle.prior_mean = lambda d_: np.zeros(d_["Y_norm"].shape[1])          # GP_Robust: zero prior mean
        mean, var = oracle.posterior_chol(P, zero)
    finally:
        oracle.prior_mean = orig
    lcb, ucb = oracle.bounds(mean, var, beta)
    Nxc = 14 * 13
    worst = lcb[:, 1:].reshape(9, Nxc, 2).min(axis=0)                           # min over d (slowest axis) per x_c
    R = (worst >= 0.0).all(axis=1)
    for name, arr in (("ucb", ucb[:, 0]), ("mean", mean[:, 0]), ("lcb", lcb[:, 0])):
        score = arr.reshape(9, Nxc).max(axis=0)
        x, val = bo.Minimize_Maximise(getattr(bo, name))
        assert bo.n_robust_safe == R.sum()
        if R.any():
            io, vo = oracle.masked_argmin(score, R)
            assert np.allclose(x, P[io, :2]) and val == pytest.approx(vo, rel=1e-9)
        else:
            assert np.isnan(x).all() and val == np.inf
    xc = np.array([0.11, -0.23])
    D = np.linspace(lo[2], hi[2], 9)
    pts = np.column_stack([np.tile(xc, (9, 1)), D])
    try:
        oracle.prior_mean = lambda d_: np.zeros(d_["Y_norm"].shape[1])
        mo, vo = oracle.posterior_chol(pts, zero)
    finally:
        oracle.prior_mean = orig
    assert bo.Maximise_d(bo.ucb, xc, 1) == pytest.approx(np.max(mo[:, 1] + beta * np.sqrt(vo[:, 1])), rel=1e-9)
    assert bo.Minimise_d(bo.lcb, xc, 2) == pytest.approx(np.min(mo[:, 2] - beta * np.sqrt(vo[:, 2])), rel=1e-9)
    bo.engine.set_option("prior_mean_zero", 0)
    bo.engine.close()


def test_lipschitz_needs_a_posterior_with_gradients(engine):
    """sbo_lipschitz after sbo_posterior(with_grad=0) must fail instead of returning zeros (L = 0 would make every
    safe point 'reach' every unsafe point)."""
    from sbo_b200 import workloads
    ds, lo, hi, pts, beta = workloads.small(d=3, pts_per_dim=8, n=30, seed=2, G=3)
    engine.set_model(ds)
    engine.set_grid(lo, hi, pts)
    engine.posterior(with_grad=False, fetch=False)
    with pytest.raises(ValueError):
        engine.lipschitz()
    engine.posterior(with_grad=True, fetch=False)
    assert np.all(engine.lipschitz()[1:] > 0)


def test_library_communicator_single_rank_steps_match_the_plain_steps():
    """csrc/comm.cu on one GPU: a 1-rank NCCL communicator (sbo_comm_unique_id / sbo_comm_init) and the whole sharded steps
    (sbo_safeopt_step_sharded / sbo_goose_step_sharded) must return what GridEngine.safeopt_step / goose_step return.
    (The 2-GPU agreement tests need a 2-GPU box; this one runs wherever the GPU suite runs.)"""
    import sbo_b200
    from sbo_b200 import _capi as capi, workloads
    ds, lo, hi, pts, beta = workloads.small(d=4, pts_per_dim=9, n=200, seed=11, G=4)
    eng = sbo_b200.GridEngine(0)
    try:
        eng.set_grid(lo, hi, pts)
        eng.set_shard_cyclic(0, 1, 256)
        eng.comm_init(0, 1, eng.comm_unique_id())
        for mode, prec in (("fantasy", "tf32"), ("fantasy", "fp64"), ("lipschitz", "fp64")):
            a = eng.safeopt_step(ds, beta, mode=mode, precision=prec, unsafe_rule=capi.UNSAFE_ANY)
            b = eng.safeopt_step_sharded(ds, beta, mode=mode, precision=prec, unsafe_rule=capi.UNSAFE_ANY)
            for k in ("n_safe", "n_unsafe", "n_min", "min_ucb0_idx", "minimizer_idx", "expander_idx", "x_new_idx"):
                assert a[k] == b[k], (mode, prec, k)
            assert a["expander"]["n_hit"] == b["expander"]["n_hit"] and a["expander"]["n_z"] == b["expander"]["n_z"]
            assert a["min_ucb0"] == b["min_ucb0"] and a["minimizer_var"] == b["minimizer_var"]
        a = eng.goose_step(ds, beta, unsafe_rule=capi.UNSAFE_ANY)
        b = eng.goose_step_sharded(ds, beta, unsafe_rule=capi.UNSAFE_ANY)
        for k in ("n_safe", "n_unsafe", "min_lcb0_idx", "target_idx", "x_new_idx", "explore_idx"):
            assert a[k] == b[k], k
        assert a["target_lcb"] == b["target_lcb"]
    finally:
        eng.close()
