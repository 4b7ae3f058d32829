"""CPU: host-side logic of the drop-in classes that does not touch the GPU."""
import numpy as np
import pytest

import sbo_b200
from sbo_b200 import workloads
from sbo_b200.engine import unpack_bits
from sbo_b200.models import GP_Safe
from sbo_b200.problems import Benoit_Problem, WilliamOttoReactor_Problem


def test_normalisation_and_dataset_contract(oracle, c1):
    gp = GP_Safe.GP([Benoit_Problem.Benoit_System_1, Benoit_Problem.con1_system_tight])
    gp.GP_initialization(c1["X"][:9], c1["Y"][:9], 'RBF', multi_hyper=5, var_out=True, hypopt=c1["hyp_9"])
    ds = gp.inference_datasets
    ref = oracle.make_inference_datasets(c1["X"][:9], c1["Y"][:9], c1["hyp_9"])
    for k in ["X_mean", "X_std", "Y_mean", "Y_std", "X_norm", "Y_norm", "hypopt"]:
        np.testing.assert_allclose(ds[k], ref[k], rtol=1e-13, atol=1e-13)
    for a, b in zip(ds["invKopt"], ref["invKopt"]):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-6)
    assert gp.n_point == 9 and gp.nx_dim == 2 and gp.ny_dim == 2 and gp.n_fun == 2


def test_errors_like_reference(c1):
    gp = GP_Safe.GP([Benoit_Problem.Benoit_System_1, Benoit_Problem.con1_system_tight])
    with pytest.raises(ValueError):
        gp.GP_initialization(c1["X"][:4], c1["Y"][:4], 'Matern', multi_hyper=1)
    gp.GP_initialization(c1["X"][:4], c1["Y"][:4], 'RBF', multi_hyper=1, hypopt=c1["hyp_4"])
    with pytest.raises(ValueError):
        gp.Cov_mat('RBF', gp.X_norm, gp.X_norm, np.ones(3), 1.0)      # GP_Safe.py:134-135


def test_hyper_fit_seeded(c1, oracle):
    gp = GP_Safe.GP([Benoit_Problem.Benoit_System_1, Benoit_Problem.con1_system_tight])
    gp.hyper_seed = 1004
    gp.GP_initialization(c1["X"][:4], c1["Y"][:4], 'RBF', multi_hyper=1)
    assert gp.hypopt.shape == (4, 2)
    nll_fit = gp.negative_loglikelihood(gp.hypopt[:, 0], gp.X_norm, gp.Y_norm[:, :1])
    nll_gold = gp.negative_loglikelihood(c1["hyp_4"][:, 0], gp.X_norm, gp.Y_norm[:, :1])
    assert nll_fit <= nll_gold + 1e-3
    assert nll_fit == pytest.approx(oracle.negative_loglikelihood(gp.hypopt[:, 0], gp.X_norm, gp.Y_norm[:, :1]), rel=1e-9)


def test_data_sampling_in_ball():
    gp = GP_Safe.GP([Benoit_Problem.Benoit_System_1, Benoit_Problem.con1_system_tight])
    X, Y = gp.Data_sampling(50, np.array([1.4, -0.8]), 0.3)
    assert X.shape == (50, 2) and Y.shape == (50, 2)
    assert np.all(np.linalg.norm(X - np.array([1.4, -0.8]), axis=1) <= 0.3 + 1e-12)
    assert Y[3, 0] == pytest.approx(Benoit_Problem.Benoit_System_1(X[3]))


def test_bit_unpack():
    w = np.array([0b1011, 1 << 31], dtype=np.uint32)
    b = unpack_bits(w, 64)
    assert b[:4].tolist() == [True, True, False, True] and b[63] and b.sum() == 4


def test_plants():
    assert Benoit_Problem.con1_system_tight(np.array([0.36845785, -0.39299271])) == pytest.approx(0.0, abs=1e-6)
    assert Benoit_Problem.Benoit_System_1(np.array([0.36845785, -0.39299271])) == pytest.approx(0.145249, abs=5e-3)   # test/test_GoOSE.py:182 tolerance
    wo = WilliamOttoReactor_Problem.WilliamOttoReactor()
    u = np.array([6.09187167, 79.46867795])          # first sampled point of the reference's WOR run '0'
    # the recorded outputs (16.41, 0.0218, 0.0344) were taken with noise = 0.01 on Fb
    # (test/test_SafeOpt.py:266), so only the neighbourhood is checked, plus the steady-state residual
    assert abs(wo.get_objective(u) - 16.4139329) < 30.0
    assert wo.get_constraint1(u) == pytest.approx(2.18462911e-02, abs=5e-3)
    assert wo.get_constraint2(u) == pytest.approx(3.44493646e-02, abs=5e-3)
    sol, _ = wo.steady_state(u, 0.0)
    assert np.max(np.abs(wo.odecallback(sol, u, 0.0))) < 1e-10


def test_workloads_deterministic():
    a = workloads.c4(pts_per_dim=8, n=32)
    b = workloads.c4(pts_per_dim=8, n=32)
    assert np.array_equal(a[0]["X_norm"], b[0]["X_norm"]) and a[3] == [8, 8, 8, 8]
    assert a[0]["hypopt"].shape == (6, 4)


# ------------------------------------------------------------------------------------------------
# drivers.py: the reference's loops (test/test_SafeOpt.py:135-253, test/test_GoOSE.py:142-190) over a scripted BO
# ------------------------------------------------------------------------------------------------
class _ScriptedBO:
    """Duck-typed BO whose acquisition answers are scripted: checks control flow only (no GPU)."""
    nx_dim, n_fun = 2, 2

    def __init__(self, script):
        self.script, self.k, self.added = script, 0, []
        self.plant_system = [lambda x, noise=0: float(x[0] ** 2 + x[1] ** 2), lambda x, noise=0: float(1.0 - x[0])]

    def _cur(self):
        return self.script[min(self.k, len(self.script) - 1)]

    def Minimizer(self):
        return np.array(self._cur()["min"][0]), self._cur()["min"][1]

    def Expander(self):
        return np.array(self._cur()["exp"][0]), self._cur()["exp"][1]

    def ucb(self, x, j):
        return self._cur().get("ucb", 1.0)

    def minimize_obj_lcb(self):
        return np.array(self._cur()["safe"][0]), self._cur()["safe"][1]

    def Target(self):
        return np.array(self._cur()["tgt"][0]), self._cur()["tgt"][1]

    def explore_safeset(self, t):
        return np.asarray(t) * 0.5

    def calculate_plant_outputs(self, x, noise=0):
        return np.array([f(x, noise) for f in self.plant_system])

    def add_sample(self, x, y, hypopt=None):
        self.added.append((np.asarray(x), np.asarray(y)))
        self.k += 1


def test_safeopt_driver_decision_rule_and_early_stop():
    from sbo_b200 import drivers
    bo = _ScriptedBO([{"min": ([0.1, 0.2], 0.5), "exp": ([0.3, 0.4], 0.2)},          # std_min > std_exp -> minimiser
                      {"min": ([0.1, 0.2], 0.1), "exp": ([0.3, 0.4], 0.2)},          # else -> expander
                      {"min": ([0.5, 0.5], 0.001), "exp": ([0.6, 0.6], 0.002)},      # both < 0.01 -> stop after adding
                      {"min": ([9, 9], 1.0), "exp": ([9, 9], 1.0)}])
    data = drivers.run_safeopt(bo, n_iteration=10)
    assert data["i"] == [0, 1, 2] and len(bo.added) == 3                               # test_SafeOpt.py:175-179
    assert [i["chose"] for i in data["info"]] == ["minimizer", "expander", "expander"]
    assert data["x_0"][:2] == [0.1, 0.3] and data["obj"][0] == pytest.approx(0.05)
    # multi-run guard (test_SafeOpt.py:228-240): an expander with every constraint ucb < 0 is not taken
    bo = _ScriptedBO([{"min": ([0.1, 0.2], 0.1), "exp": ([0.3, 0.4], 0.2), "ucb": -1.0}])
    x, info = drivers.safeopt_iteration(bo, require_lipschitz_ucb=True)
    assert info["chose"] == "minimizer" and np.allclose(x, [0.1, 0.2])
    # empty expander set: (nan, 0.0) never beats the minimiser
    bo = _ScriptedBO([{"min": ([0.1, 0.2], 0.05), "exp": ([np.nan, np.nan], 0.0)}])
    assert drivers.safeopt_iteration(bo)[1]["chose"] == "minimizer"


def test_goose_driver_decision_rule():
    from sbo_b200 import drivers
    bo = _ScriptedBO([{"safe": ([0.2, 0.2], 0.10), "tgt": ([0.8, 0.8], 0.30)},        # min_safe_lcb <= target_lcb
                      {"safe": ([0.2, 0.2], 0.50), "tgt": ([0.8, 0.4], 0.30)},        # else explore towards the target
                      {"safe": ([0.2, 0.2], 0.50), "tgt": ([np.nan, np.nan], np.inf)}])   # empty target set
    data = drivers.run_goose(bo, n_iteration=3, f_opt=None)
    assert [i["chose"] for i in data["info"]] == ["safe_minimum", "explore", "safe_minimum"]
    assert np.isnan(data["info"][0]["x_target"]).all()                                # test_GoOSE.py:160
    assert (data["x_0"][1], data["x_1"][1]) == (0.4, 0.2)
    stop = drivers.run_goose(_ScriptedBO([{"safe": ([0.3, 0.2], 0.1), "tgt": ([1, 1], 9.)}]), n_iteration=5,
                             f_opt=0.13, tol=0.005)
    assert stop["i"] == [0]                                                           # test_GoOSE.py:182


def test_result_file_layout_matches_reference_consumers(tmp_path):
    from sbo_b200 import drivers
    data = {"0": {"sampled_x": np.zeros((4, 2)), "sampled_output": np.ones((4, 2)),
                  "observed_x": np.zeros((3, 2)), "observed_output": np.arange(6.).reshape(3, 2)},
            "1": {"sampled_x": np.zeros((4, 2)), "sampled_output": np.ones((4, 2)),
                  "observed_x": np.zeros((2, 2)), "observed_output": np.arange(4.).reshape(2, 2)}}
    p = tmp_path / "runs.npz"
    drivers.save_runs(p, data)
    raw = np.load(p, allow_pickle=True)
    # the access pattern of the reference's utils/utils_solve_Benoit.py:16-31
    n_start = len(raw.items())
    shapes = [np.array(raw[f"{i}"].item()["observed_output"]).shape for i in range(n_start)]
    assert shapes == [(3, 2), (2, 2)]
    back = drivers.load_runs(p)
    np.testing.assert_array_equal(back["1"]["observed_output"], data["1"]["observed_output"])


def test_plants_match_the_references_own_outputs():
    """problems/*.py (NumPy restatements used as fixtures) vs outputs of the reference's own plants at noise = 0
    (tests/golden/make_plant_vectors.py ran /root/reference/problems/*.py over the NumPy jax stand-in)."""
    from conftest import load_golden
    r = load_golden("ref_plants")
    for u, f1, f2, c1, c1t in zip(r["benoit_u"], r["benoit_f1"], r["benoit_f2"], r["benoit_con1"], r["benoit_con1_tight"]):
        assert Benoit_Problem.Benoit_System_1(u) == pytest.approx(f1, rel=1e-14, abs=1e-15)
        assert Benoit_Problem.Benoit_System_2(u) == pytest.approx(f2, rel=1e-14, abs=1e-15)
        assert Benoit_Problem.con1_system(u) == pytest.approx(c1, rel=1e-14, abs=1e-15)
        assert Benoit_Problem.con1_system_tight(u) == pytest.approx(c1t, rel=1e-14, abs=1e-15)
    wo = WilliamOttoReactor_Problem.WilliamOttoReactor()
    for u, f, g1, g2 in zip(r["wor_u"], r["wor_obj"], r["wor_con1"], r["wor_con2"]):
        assert wo.get_objective(u) == pytest.approx(f, rel=1e-10)
        assert wo.get_constraint1(u) == pytest.approx(g1, abs=1e-12)
        assert wo.get_constraint2(u) == pytest.approx(g2, abs=1e-12)


def test_plot_helpers_run_against_a_recording_matplotlib(tmp_path, monkeypatch):
    """matplotlib / imageio are not installed in this image (they are imported lazily): run the helpers against
    recording stand-ins to check the call sequence and the reference signatures (utils_SafeOpt.py:14-84)."""
    import sys
    import types
    from unittest import mock
    plt = mock.MagicMock(name="pyplot")
    top, bottom, fig2 = mock.MagicMock(), mock.MagicMock(), mock.MagicMock()
    plt.subplots.return_value = (fig2, (top, bottom))
    mpl = types.ModuleType("matplotlib")
    mpl.use = mock.MagicMock()
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    iio = mock.MagicMock(name="imageio.v2")
    pkg = types.ModuleType("imageio")
    pkg.v2 = iio
    monkeypatch.setitem(sys.modules, "imageio", pkg)
    monkeypatch.setitem(sys.modules, "imageio.v2", iio)
    from sbo_b200.utils import utils_GoOSE, utils_SafeOpt
    x0, x1 = np.meshgrid(np.linspace(-.6, 1.5, 5), np.linspace(-1, 1, 4))
    mask = x0 > 0.3
    X = np.array([[1.4, -0.8], [1.3, -0.7]])
    bound = np.array([[-.6, 1.5], [-1., 1.]])
    data = {"x_0": [1.2, 1.0], "x_1": [-0.7, -0.6], "x_target_0": 0.5, "x_target_1": -0.4}
    utils_SafeOpt.create_frame(utils_SafeOpt.plot_safe_region_Benoit(X, x0, x1, mask, x0 ** 2 + x1 ** 2, bound, data),
                               str(tmp_path / "f0.png"))
    ax = plt.figure.return_value.gca.return_value
    assert ax.contourf.called and ax.contour.called and ax.plot.call_count >= 4
    plt.gcf.return_value.savefig.assert_called_with(str(tmp_path / "f0.png"))
    n_before = ax.plot.call_count
    utils_GoOSE.plot_safe_region_Benoit(X, x0, x1, mask, x0 ** 2 + x1 ** 2, bound, data)
    assert ax.plot.call_count >= n_before + 6                      # the frame again + target marker + dashed line
    utils_SafeOpt.plant_outputs_drawing([0, 1], [1.0, 0.5], [0.2, 0.1], "out.png", output_dir=str(tmp_path))
    assert top.plot.called and bottom.axhline.called
    fig2.savefig.assert_called_with(str(tmp_path / "out.png"))
    f = tmp_path / "frame.png"
    f.write_bytes(b"x")
    utils_SafeOpt.create_GIF(700, [str(f)], "run.gif", output_dir=str(tmp_path))
    assert iio.mimsave.called and not f.exists()


def test_trust_region_update_rule():
    """models/GP_TR.update_TR (reference GP_TR.py:56-91) is host control flow: check every branch with a scripted
    posterior mean."""
    from sbo_b200.models import GP_TR
    P = {'radius': 0.5, 'radius_max': 1, 'radius_red': 0.8, 'radius_inc': 1.1, 'rho_lb': 0.2, 'rho_ub': 0.8}
    bo = GP_TR.BO([None, None], np.array([[-.6, 1.5], [-1., 1.]]), 3., P)
    means = {}
    bo.GP_inference_jit = lambda x, ds: (np.array([means[tuple(x)], 0.0]), np.zeros(2))
    x0, x1 = (1.4, -0.8), (1.2, -0.7)
    means[x0], means[x1] = 2.0, 1.0                                       # predicted change -1
    assert bo.update_TR(x0, x1, 0.5, [2.0, 0.3], [1.5, -0.1]) == (x0, 0.4)            # constraint violated -> shrink
    assert bo.update_TR(x0, x1, 0.5, [2.0, 0.3], [2.1, 0.1]) == (x0, 0.4)             # objective went up -> shrink
    assert bo.update_TR(x0, x1, 0.5, [2.0, 0.3], [1.9, 0.1]) == (x0, 0.4)             # rho = 0.1 < rho_lb
    assert bo.update_TR(x0, x1, 0.5, [2.0, 0.3], [1.5, 0.1]) == (x1, 0.5)             # rho = 0.5 -> accept, keep r
    assert bo.update_TR(x0, x1, 0.5, [2.0, 0.3], [1.1, 0.1]) == (x1, pytest.approx(0.55))   # rho = 0.9 -> grow
    assert bo.update_TR(x0, x1, 0.95, [2.0, 0.3], [1.1, 0.1]) == (x1, 1)              # capped at radius_max
    assert bo.TR_constraint(np.array([1.0, 0.0]), np.array([1.0, 0.3]), 0.5) == pytest.approx(0.2, abs=1e-7)
    # ball mask geometry: x_0 fastest, numpy.linspace axes
    bo.grid_points_per_dim = [5, 3]
    bo._grid_set = True
    bo.grid_shape = (5, 3)
    m = bo._ball_mask([1.5, 1.0], 0.6).reshape(3, 5)
    assert m[2, 4] and m[2, 3] and not m[2, 2] and not m[1, 4] and m.sum() == 2


def test_drivers_stop_on_an_empty_safe_set_instead_of_sampling_nan():
    """with no safe grid point the acquisition functions return NaN coordinates; the drivers must not evaluate the
    plant there (it would poison the normalisation and the factor of every later iteration)."""
    import sbo_b200  # noqa: F401
    from sbo_b200 import drivers

    class Empty:
        n_fun, nx_dim = 2, 2
        def Minimizer(self): return np.full(2, np.nan), 0.0
        def Expander(self): return np.full(2, np.nan), 0.0
        def minimize_obj_lcb(self): return np.full(2, np.nan), np.inf
        def Target(self): return np.full(2, np.nan), np.inf
        def explore_safeset(self, t): return np.full(2, np.nan)

    with pytest.raises(drivers.EmptySafeSet):
        drivers.safeopt_iteration(Empty())
    with pytest.raises(drivers.EmptySafeSet):
        drivers.goose_iteration(Empty())

    class OnlyExpander(Empty):
        def Expander(self): return np.array([0.1, 0.2]), 0.0
    x, info = drivers.safeopt_iteration(OnlyExpander())        # std tie -> "expander" branch is NaN-free
    assert np.allclose(x, [0.1, 0.2])
