"""GPU debug: TF32 tcgen05 fantasy kernel vs (a) the FP64 oracle, (b) the oracle with TF32-rounded operands."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sbo_b200
from sbo_b200 import _capi as capi, workloads
from oracle import gp_oracle as O

def case(eng, ds, lo, hi, grid, beta, rule, variant, name):
    eng.set_option("fantasy_variant", variant)
    eng.set_model(ds); eng.set_grid(lo, hi, grid)
    m, v = eng.posterior(keep_v=2)
    eng.sets(beta, capi.UNSAFE_ALL if rule == "all" else capi.UNSAFE_ANY)
    ex = eng.expander(beta, None, capi.MODE_FANTASY, capi.PREC_TF32, want_counts=True)
    eng.posterior(keep_v=1, fetch=False); eng.sets(beta, capi.UNSAFE_ALL if rule == "all" else capi.UNSAFE_ANY)
    ex64 = eng.expander(beta, None, capi.MODE_FANTASY, capi.PREC_FP64, want_counts=True)
    pts = O.make_grid(lo, hi, grid)
    lcb, _ = O.bounds(m, v, beta)
    S, Z = O.safe_mask(lcb), O.unsafe_mask(lcb, rule)
    w64 = O.fantasy_counts(pts, ds, beta, S, Z)
    wtf = O.fantasy_counts(pts, ds, beta, S, Z, dtype="tf32")
    g = ex["counts"][S].astype(np.int64)
    print(f"{name} v{variant}: |S|={S.sum()} |Z|={Z.sum()} newly-safe total fp64={w64[S].sum()} tf32emul={wtf[S].sum()} gpu_tc={g.sum()} gpu_f64={ex64['counts'][S].sum()}")
    print(f"   sum|gpu_tc - tf32emul| = {np.abs(g - wtf[S]).sum()}   sum|gpu_tc - fp64| = {np.abs(g - w64[S]).sum()}   sum|tf32emul - fp64| = {np.abs(wtf[S]-w64[S]).sum()}  sum|gpu_f64-fp64|={np.abs(ex64['counts'][S]-w64[S]).sum()}")
    hy = ds["hypopt"]; d = pts.shape[1]
    print("   sf2", np.exp(2*hy[d]).round(3), "sn2", np.exp(2*hy[d+1]))

if __name__ == "__main__":
    eng = sbo_b200.GridEngine(0)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from conftest import load_golden, golden_ds
    c1, c3 = load_golden("c1_benoit"), load_golden("c3_wor")
    for variant in (0, 1):
        case(eng, golden_ds(O, c1, 9), c1["lo"], c1["hi"], [48, 40], 3.0, "all", variant, "c1-9")
        case(eng, golden_ds(O, c3, 20), c3["lo"], c3["hi"], [45, 61], 2.0, "any", variant, "c3-20")
        case(eng, golden_ds(O, c3, 35), c3["lo"], c3["hi"], [70, 50], 2.0, "all", variant, "c3-35")
        ds, lo, hi, ppd, beta = workloads.small(d=4, pts_per_dim=9, n=200, seed=11, G=4)
        case(eng, ds, lo, hi, ppd, beta, "any", variant, "syn-d4")
