"""CPU: the measurement infrastructure of bench.py's CPU arm (oracle/cpu_arm.py, oracle/pair_epilogue.c, oracle/de_step.py)
computes what the oracle computes -- a baseline that is timed must also be right."""
import numpy as np
import pytest


def _small():
    import sbo_b200  # noqa: F401
    from sbo_b200 import workloads
    return workloads.small(d=3, pts_per_dim=9, n=40, seed=7, G=3)


@pytest.mark.parametrize("use_c_epilogue", [True, False])
def test_cpu_arm_pair_stages_match_the_oracle(oracle, use_c_epilogue, monkeypatch):
    from oracle import cpu_arm
    if not use_c_epilogue:
        monkeypatch.setattr(cpu_arm, "_EPI", False)            # torch fall-back of the element-wise tails
    elif cpu_arm._epilogue_lib() is None:
        pytest.skip("oracle/_build/libpair_epilogue.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    ds, lo, hi, pts, beta = _small()
    P = oracle.make_grid(lo, hi, pts)
    dso, Ws, _ = cpu_arm.model_state(ds)
    pp = cpu_arm.per_point(P, dso, Ws, beta, fantasy=True)
    mean, var = oracle.posterior_chol(P, ds)
    assert np.max(np.abs(pp["mean"] - mean)) <= 1e-8 * np.max(np.abs(mean))
    lcb, ucb = oracle.bounds(mean, var, beta)
    S, Z = oracle.safe_mask(lcb), oracle.unsafe_mask(lcb)
    assert np.array_equal(pp["S"], S) and np.array_equal(pp["Z"], Z)
    xs, zs = np.flatnonzero(S), np.flatnonzero(Z)
    counts, secs, flops = cpu_arm.pairs_fantasy(P, dso, beta, pp, xs, zs, block=64)
    want = oracle.fantasy_counts(P, ds, beta, S, Z)[S]
    margin = np.abs(oracle.fantasy_margin(P, ds, beta, S, Z))
    assert np.all(np.abs(counts - want) <= (margin <= 1e-7).sum(axis=0))      # inverse-form vs Cholesky-form posterior inputs
    L = [0.0] + [oracle.lipschitz_constant(P, dso, 2)] * 2
    hit, _, _ = cpu_arm.pairs_lipschitz(P, pp, L, xs, zs, 3, block=64)
    ex = oracle.expander_lipschitz(P, S, Z, ucb, var, L)
    for c in range(2):
        got = np.zeros(P.shape[0], bool)
        got[xs[hit[c]]] = True
        diff = got != ex["masks"][c]
        assert diff.sum() <= 2                                               # dot-form distance vs the reference's +1e-8 offset form


def test_sample_step_and_extrapolation():
    from oracle import cpu_arm
    ds, lo, hi, pts, beta = _small()
    s = cpu_arm.sample_step(ds, lo, hi, pts, beta, "fantasy", n_points=300, n_x=64, n_z=128, seed=1)
    assert s["n_points"] == 300 and s["t_points"] > 0 and 0 <= s["safe_frac"] <= 1
    N = int(np.prod(pts))
    t = cpu_arm.extrapolate(s, N, 10 * max(s["pairs"], 1))
    assert t >= s["t_model"] + s["t_points"] * N / 300 - 1e-12
    assert "point ratio" in cpu_arm.describe(s, N, 10 * max(s["pairs"], 1), "fantasy")


def test_reference_shaped_de_step_runs_and_stays_feasible(oracle, c1):
    """oracle/de_step.py restates SafeOpt.Minimizer/Expander (models/SafeOpt.py:47-124): its optimum must be safe under the
    oracle's bounds (the DE is seeded and capped here: this is a smoke test of the timed context step, not a parity test)."""
    from conftest import golden_ds
    from oracle import de_step
    ds = golden_ds(oracle, c1, 9)
    bound = np.column_stack([c1["lo"], c1["hi"]])
    r = de_step.time_step(ds, bound, 3.0, "safeopt", seed=0, maxiter=8)
    x = np.array(r["x_new"])
    assert r["seconds"] > 0 and r["n_inference"] > 100 and x.shape == (2,)
    assert np.all(x >= bound[:, 0] - 1e-9) and np.all(x <= bound[:, 1] + 1e-9)
    g = de_step.time_step(ds, bound, 3.0, "goose", seed=0, maxiter=5)
    assert len(g["x_new"]) == 2
