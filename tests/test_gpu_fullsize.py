"""GPU, full size: the BENCHMARKED configurations against the FP64 oracle (VERDICT r1 "oracle-check the benchmarked configs").

* C4 (the bench workload: 32^4 grid, n = 512, G = 4): exact FP64 newly-safe counts of 160 candidates against ALL 757 532
  unsafe points (tests/golden/c4_full_sampled.npz, produced by tests/golden/make_fullsize_vectors.py with the oracle).
* C5's model (d = 6, n = 2048, 64 K blocks, 12-float records) on 49 152 of its grid points as explicit points
  (tests/golden/c5_points_sampled.npz): every safe point's count.
* C4, reference-exact Lipschitz expander: the exact Euclidean distance transform of the unsafe set is the oracle
  (SURVEY.md section 7 step 1): x is an expander for constraint c iff dist(x, Z) <= ucb_c(x)/L.

Stated tolerances (normalised margin m = min_c (mu'_c - beta sigma'_c), sf2 = 1 on these models):
  tf32, tf32x3   (default: FP64 refinement of every pair inside the tensor-core error bound) the FP64 counts, exactly
  -norefine      option fantasy_refine = 0, the tensor-core value decides:
                 tf32x3  a decision may differ only where |m| <= 1e-4 * sf2 (north_star's TF32-mode tolerance)
                 tf32    only where |m| <= 1e-3 * sf2 * (1 + gain(x))  (operand rounding 2^-11, amplified by the update gain)
  fp64    GPU FP64 kernel: only where |m| <= 1e-4 * sf2 (expected: identical)
"""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _capi():
    from sbo_b200 import _capi
    return _capi


def _check(name, got, exact, near):
    d = np.abs(got.astype(np.int64) - exact)
    bad = d > near
    print(f"{name}: {int((d > 0).sum())} of {d.size} sampled candidates differ, sum|diff| = {int(d.sum())}, "
          f"near-threshold pairs allowed = {int(near.sum())}, newly-safe pairs (FP64) = {int(exact.sum())}")
    assert not bad.any(), (name, np.flatnonzero(bad)[:10], d[bad][:10], near[bad][:10])


def test_c4_fantasy_counts_vs_full_size_oracle(engine):
    capi = _capi()
    from sbo_b200 import workloads
    f = np.load(os.path.join(GOLD, "c4_full_sampled.npz"))
    ds, lo, hi, pts, beta = workloads.c4()
    assert float(f["beta"]) == beta
    engine.set_model(ds)
    engine.set_grid(lo, hi, pts)
    xi = f["x_idx"]
    counts = {}
    for prec in ("tf32", "tf32-norefine", "tf32x3", "tf32x3-norefine", "fp64"):
        p, kv = capi.PRECISIONS[prec.split("-")[0]]
        engine.set_option("fantasy_refine", 0 if prec.endswith("norefine") else 2)
        engine.posterior(keep_v=kv, fetch=False)
        s = engine.sets(beta, capi.UNSAFE_ALL)
        assert s["n_safe"] == int(f["n_safe"]) and s["n_unsafe"] == int(f["n_unsafe"])
        ex = engine.expander(beta, None, capi.MODE_FANTASY, p, want_counts=True)
        engine.set_option("fantasy_refine", 2)
        counts[prec] = ex["counts"].astype(np.int64)
        print(f"C4 {prec}: n_hit {ex['n_hit']}, pairs evaluated {ex['pairs_evaluated']} of {ex['pairs_algorithmic']}, best {ex['best_idx']}, "
              f"refined {ex['n_ambiguous']} pairs ({ex['n_refined_safe']} safe)")
        near = {"tf32-norefine": f["near_tf32"], "tf32x3-norefine": f["near_1e4"]}.get(prec, np.zeros_like(f["near_1e4"]))
        _check(f"C4 {prec}", counts[prec][xi], f["counts"], near)
        engine.release(3)
    # all 116 645 candidates: the split-TF32 mode against the GPU FP64 kernel (itself pinned to the oracle above)
    d = np.abs(counts["tf32x3"] - counts["fp64"])
    print(f"C4 tf32x3 vs fp64 over all candidates: {int((d > 0).sum())} candidates differ, sum|diff| = {int(d.sum())} of "
          f"{int(counts['fp64'].sum())} newly-safe pairs; expander set sizes {int((counts['tf32x3'] > 0).sum())} / {int((counts['fp64'] > 0).sum())}")
    assert d.sum() <= 2, "the refined split-TF32 counts must be the FP64 counts"
    dt = np.abs(counts["tf32"] - counts["fp64"])
    print(f"C4 tf32 (refined, the bench default) vs fp64 over all candidates: {int((dt > 0).sum())} candidates differ, sum|diff| = {int(dt.sum())}")
    assert dt.sum() <= 2, "the refined single-pass TF32 counts must be the FP64 counts"
    dn = np.abs(counts["tf32x3-norefine"] - counts["fp64"])
    print(f"C4 tf32x3 without the refinement vs fp64: {int((dn > 0).sum())} candidates differ, sum|diff| = {int(dn.sum())}")
    # unrefined, the FP32 accumulator of the tensor core bounds the split mode (error grows with K)
    assert dn.sum() <= 5e-4 * counts["fp64"].sum() + 8
    d1 = np.abs(counts["tf32-norefine"] - counts["fp64"])
    print(f"C4 tf32 (single pass, no refinement) vs fp64: {int((d1 > 0).sum())} candidates differ, sum|diff| = {int(d1.sum())}")


def test_c5_model_fantasy_counts_on_sampled_points(engine):
    capi = _capi()
    spec = importlib.util.spec_from_file_location("make_fullsize_vectors", os.path.join(GOLD, "make_fullsize_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    f = np.load(os.path.join(GOLD, "c5_points_sampled.npz"))
    ds, beta, idx, P = mod.c5_sample_points()
    assert np.array_equal(idx, f["grid_idx"])
    engine.set_model(ds)
    engine.set_points(P)
    xl = f["x_local"]
    for prec in ("tf32", "tf32-norefine", "tf32x3", "tf32x3-norefine", "fp64"):
        p, kv = capi.PRECISIONS[prec.split("-")[0]]
        engine.set_option("fantasy_refine", 0 if prec.endswith("norefine") else 2)
        engine.posterior(keep_v=kv, fetch=False)
        s = engine.sets(beta, capi.UNSAFE_ALL)
        assert s["n_safe"] == int(f["n_safe"]) and s["n_unsafe"] == int(f["n_unsafe"])
        ex = engine.expander(beta, None, capi.MODE_FANTASY, p, want_counts=True)
        engine.set_option("fantasy_refine", 2)
        print(f"C5 model {prec}: n_hit {ex['n_hit']}, pairs evaluated {ex['pairs_evaluated']} of {ex['pairs_algorithmic']}, "
              f"refined {ex['n_ambiguous']} pairs ({ex['n_refined_safe']} safe)")
        assert (ex["counts"][np.setdiff1d(np.arange(P.shape[0]), xl)] == 0).all()
        near = {"tf32-norefine": f["near_tf32"], "tf32x3-norefine": f["near_1e4"]}.get(prec, np.zeros_like(f["near_1e4"]))
        _check(f"C5 model {prec}", ex["counts"][xl], f["counts"], near)
        engine.release(3)


def test_c4_lipschitz_expander_vs_distance_transform(engine, oracle):
    from scipy import ndimage
    capi = _capi()
    from sbo_b200 import workloads
    ds, lo, hi, pts, beta = workloads.c4()
    engine.set_model(ds)
    engine.set_grid(lo, hi, pts)
    mean, var = engine.posterior(with_grad=True)
    st = engine.safeopt_step(ds, beta, mode="lipschitz", upload=False)
    G = mean.shape[1]
    L = st["L"][G - 1]
    lcb, ucb = oracle.bounds(mean, var, beta)
    S, Z = oracle.safe_mask(lcb), oracle.unsafe_mask(lcb)
    assert st["n_safe"] == S.sum() and st["n_unsafe"] == Z.sum()
    shape = tuple(int(p) for p in pts[::-1])                          # slowest axis first (x_0 fastest)
    step = [(h - l) / (p - 1) for l, h, p in zip(lo, hi, pts)][::-1]
    dist = ndimage.distance_transform_edt(~Z.reshape(shape), sampling=step).ravel()     # exact distance to the nearest z
    union = np.zeros(S.size, dtype=bool)
    for c in range(1, G):
        got = engine.mask(capi.MASK_EXPANDER, c - 1)
        r = ucb[:, c] / L
        want = S & (r >= 0) & (dist <= r)
        amb = S & (np.abs(dist - r) <= 1e-6)                          # the reference adds 1e-8 per component before the norm
        diff = (got != want) & ~amb
        print(f"C4 Lipschitz expander, constraint {c}: |G_c| = {int(got.sum())}, oracle {int(want.sum())}, ambiguous {int(amb.sum())}")
        assert not diff.any(), np.flatnonzero(diff)[:10]
        union |= got
    assert st["expander"]["n_hit"] == union.sum()
    i, v = oracle.masked_argmax(var[:, 0], union)
    assert st["expander_idx"] == i
