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
    sol, _ = wo._solve(u, 0.0)
    assert np.max(np.abs(wo.odecallback(sol, u, 0.0))) < 1e-10


def test_workloads_deterministic():
    a = workloads.c4(pts_per_dim=8, n=32)
    b = workloads.c4(pts_per_dim=8, n=32)
    assert np.array_equal(a[0]["X_norm"], b[0]["X_norm"]) and a[3] == [8, 8, 8, 8]
    assert a[0]["hypopt"].shape == (6, 4)
