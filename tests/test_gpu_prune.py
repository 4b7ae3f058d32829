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
