"""GPU: the exact pruning of the fantasy expander (option fantasy_prune, default on; csrc/pairs.cu k_key_z / k_key_x) must
not change any FP64 count: |cov(z,x)| <= sigma_z sigma_x bounds every updated lcb, so tile pairs whose keys cannot meet are
skipped.  In the tensor-core modes the only pairs it can remove are false positives of the lower precision (pairs the
bound proves unsafe), so pruned counts are never larger."""
import numpy as np
import pytest

from conftest import golden_ds

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp64", "tf32x3"])
def test_prune_is_exact(engine, oracle, c3, precision):
    from sbo_b200 import _capi as capi, workloads
    prec, keep_v = capi.PRECISIONS[precision]
    cases = [(golden_ds(oracle, c3, 20), c3["lo"], c3["hi"], [40, 44], 2.0, capi.UNSAFE_ANY)]
    ds, lo, hi, ppd, beta = workloads.small(d=4, pts_per_dim=9, n=200, seed=11, G=4)
    cases.append((ds, lo, hi, ppd, beta, capi.UNSAFE_ALL))
    cases.append((ds, lo, hi, ppd, beta, capi.UNSAFE_ANY))
    for ds, lo, hi, grid, beta, rule in cases:
        engine.set_model(ds)
        engine.set_grid(lo, hi, grid)
        engine.posterior(keep_v=keep_v, fetch=False)
        engine.sets(beta, rule)
        try:
            engine.set_option("fantasy_prune", 0)
            full = engine.expander(beta, None, capi.MODE_FANTASY, prec, want_counts=True)
            engine.set_option("fantasy_prune", 1)
            pruned = engine.expander(beta, None, capi.MODE_FANTASY, prec, want_counts=True)
        finally:
            engine.set_option("fantasy_prune", 1)
        if precision == "fp64":
            assert np.array_equal(full["counts"], pruned["counts"])
            assert full["best_idx"] == pruned["best_idx"] and full["n_hit"] == pruned["n_hit"]
        else:
            assert np.all(pruned["counts"] <= full["counts"])
            assert int((full["counts"] - pruned["counts"]).sum()) <= max(2, int(1e-3 * full["counts"].sum()))
        assert full["n_z"] == pruned["n_z"] and full["pairs_algorithmic"] == pruned["pairs_algorithmic"]
        assert pruned["pairs_evaluated"] <= full["pairs_evaluated"]
        print(f"prune {precision}: pairs evaluated {pruned['pairs_evaluated']} of {full['pairs_evaluated']}, newly-safe total {int(full['counts'].sum())}")


def test_fp64_tensor_core_kernel_matches_the_simt_kernel(engine):
    """FP64 fantasy expander: the DMMA (mma.sync.m8n8k4.f64) tile kernel against the SIMT reference kernel, with and
    without the exact pruning, ragged tile edges included (d = 2, 3, 4, 6; n not a multiple of the K chunk)."""
    from sbo_b200 import _capi as capi, workloads
    prec, keep_v = capi.PRECISIONS["fp64"]
    for (d, ppd, n, G) in [(2, 41, 37, 2), (3, 14, 70, 3), (4, 9, 200, 4), (6, 5, 130, 3)]:
        ds, lo, hi, pts, beta = workloads.small(d=d, pts_per_dim=ppd, n=n, seed=20 + d, G=G)
        engine.set_model(ds)
        engine.set_grid(lo, hi, pts)
        engine.posterior(keep_v=keep_v, fetch=False)
        engine.sets(beta, capi.UNSAFE_ANY)
        res = {}
        try:
            for variant in (0, 1):
                for prune in (0, 1):
                    engine.set_option("fantasy_f64_variant", variant)
                    engine.set_option("fantasy_prune", prune)
                    res[(variant, prune)] = engine.expander(beta, None, capi.MODE_FANTASY, prec, want_counts=True)
        finally:
            engine.set_option("fantasy_f64_variant", 1)
            engine.set_option("fantasy_prune", 1)
        ref = res[(0, 0)]
        for k, r in res.items():
            # the two kernels sum the dot product in different orders: a pair exactly on the threshold may flip
            dd = np.abs(r["counts"].astype(np.int64) - ref["counts"])
            assert dd.sum() <= 2, (d, k, int(dd.sum()))
            assert r["n_x"] == ref["n_x"] and r["n_z"] == ref["n_z"]
        print(f"fp64 dmma d={d} n={n}: newly-safe {int(ref['counts'].sum())}, pairs evaluated {res[(1, 1)]['pairs_evaluated']} of {ref['pairs_algorithmic']}")


def test_refinement_list_overflow_reruns_with_an_exact_size(engine, oracle):
    """The ambiguous-pair list of the refining tensor-core epilogue is sized by a heuristic; when it overflows the library
    sizes it exactly and runs the GEMM once more.  Forced here with a 16-entry list: the counts must still be FP64's."""
    from sbo_b200 import _capi as capi, workloads
    ds, lo, hi, pts, beta = workloads.small(d=4, pts_per_dim=9, n=200, seed=11, G=4)
    engine.set_model(ds)
    engine.set_grid(lo, hi, pts)
    res = {}
    for prec in ("fp64", "tf32"):
        p, kv = capi.PRECISIONS[prec]
        engine.posterior(keep_v=kv, fetch=False)
        engine.sets(beta, capi.UNSAFE_ANY)
        try:
            engine.set_option("fantasy_refine_cap", 16 if prec == "tf32" else 0)
            res[prec] = engine.expander(beta, None, capi.MODE_FANTASY, p, want_counts=True)
        finally:
            engine.set_option("fantasy_refine_cap", 0)
    assert res["tf32"]["n_ambiguous"] > 16, "the case must overflow a 16-entry list to test anything"
    assert np.abs(res["tf32"]["counts"].astype(np.int64) - res["fp64"]["counts"]).sum() <= 1
    assert res["tf32"]["best_idx"] == res["fp64"]["best_idx"]


@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
def test_bounds_mode_brackets_the_fp64_expander_set(engine, precision):
    """fantasy_refine = 3 (bounds mode, for shards whose ambiguous pairs do not fit a list): no FP64 pass; the counts are the
    pairs the error bound SETTLES as newly safe, candidates with none of those but some pair inside the bound are reported
    as undecided (count -1).  Against the FP64 kernel: settled counts never exceed the FP64 counts, every certified member
    is an FP64 member, every FP64 member is certified or undecided, and x_new is the FP64 one whenever the best undecided
    candidate cannot beat it."""
    from sbo_b200 import _capi as capi, workloads
    for (d, ppd, n, G, rule) in [(4, 9, 200, 4, capi.UNSAFE_ANY), (3, 14, 70, 3, capi.UNSAFE_ALL), (6, 5, 130, 3, capi.UNSAFE_ANY)]:
        ds, lo, hi, pts, beta = workloads.small(d=d, pts_per_dim=ppd, n=n, seed=20 + d, G=G)
        engine.set_model(ds)
        engine.set_grid(lo, hi, pts)
        engine.posterior(keep_v=capi.PRECISIONS["fp64"][1], fetch=False)
        engine.sets(beta, rule)
        exact = engine.expander(beta, None, capi.MODE_FANTASY, capi.PREC_FP64, want_counts=True)
        prec, keep_v = capi.PRECISIONS[precision]
        engine.posterior(keep_v=keep_v, fetch=False)
        engine.sets(beta, rule)
        try:
            engine.set_option("fantasy_refine", 3)
            lo_b = engine.expander(beta, None, capi.MODE_FANTASY, prec, want_counts=True)
            engine.set_option("fantasy_refine", 2)
            ref = engine.expander(beta, None, capi.MODE_FANTASY, prec, want_counts=True)
        finally:
            engine.set_option("fantasy_refine", 2)
        assert np.array_equal(ref["counts"], exact["counts"])
        c, e = lo_b["counts"], exact["counts"]
        und = c < 0
        assert np.all(c[~und] <= e[~und])
        assert np.all(e[c > 0] > 0)                                 # certified members are FP64 members
        assert np.all((c > 0) | und | (e == 0))                     # FP64 members are certified or undecided
        assert lo_b["n_hit"] == int((c > 0).sum()) and lo_b["n_undecided"] == int(und.sum())
        assert lo_b["n_hit"] <= exact["n_hit"] <= lo_b["n_hit"] + lo_b["n_undecided"]
        assert lo_b["n_ambiguous"] == ref["n_ambiguous"] and lo_b["n_refined_safe"] == 0
        if lo_b["n_undecided"] == 0 or lo_b["undecided_best_value"] < lo_b["best_value"]:
            assert lo_b["best_idx"] == exact["best_idx"]
        if lo_b["n_undecided"]:
            assert und[lo_b["undecided_best_idx"]]
        print(f"bounds {precision} d={d}: certified {lo_b['n_hit']}, undecided {lo_b['n_undecided']}, fp64 {exact['n_hit']}, "
              f"ambiguous pairs {lo_b['n_ambiguous']}")
