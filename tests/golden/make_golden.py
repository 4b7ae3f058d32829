"""Generate the committed golden fixtures under tests/golden/.

Run HERE (the build container), where /root/reference is mounted:
    python tests/golden/make_golden.py

Inputs  : the reference's recorded trajectories  /root/reference/data/*.npz
          (pickled jax arrays -> read with a stub Unpickler, jax is not installed).
Outputs : tests/golden/c1_benoit.npz, tests/golden/c3_wor.npz -- realistic (X, Y)
          training sets at the sizes SURVEY.md section 8(d) names, hyper-parameters
          fitted by the oracle's *seeded* restatement of GP_Safe.py:194-234, and the
          oracle's own outputs (posterior at the points the reference's scripts print
          at, set sizes and chosen grid indices on the reference's 400x400 grid).

The reference holds no known answers for this path (zero assertions, unseeded DE), so these are oracle outputs
kept as regression vectors.  The vectors that PIN the oracle are produced by the reference's own source:
make_reference_vectors.py / make_reference_pairs.py / make_plant_vectors.py (ref_*.npz).
"""
import os
import pickle
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as O  # noqa: E402

REF = "/root/reference/data"


class _JaxStub(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("jax") and name == "_reconstruct_array":
            def rec(fun, args, arr_state, aval_state):
                a = fun(*args)
                a.__setstate__(arr_state)
                return a
            return rec
        return super().find_class(module, name)


def load_ref_npz(path):
    out = {}
    with zipfile.ZipFile(path) as z:
        for nm in z.namelist():
            with z.open(nm) as f:
                ver = np.lib.format.read_magic(f)
                if ver == (1, 0):
                    np.lib.format.read_array_header_1_0(f)
                else:
                    np.lib.format.read_array_header_2_0(f)
                out[nm[:-4]] = _JaxStub(f).load()
    return out


def run_of(path, key="0"):
    d = load_ref_npz(path)[key]
    d = d.item() if hasattr(d, "item") and not isinstance(d, dict) else d
    X = np.vstack([np.asarray(d["sampled_x"], float), np.asarray(d["observed_x"], float)])
    Y = np.vstack([np.asarray(d["sampled_output"], float), np.asarray(d["observed_output"], float)])
    return X, Y


def build(name, X, Y, sizes, lo, hi, beta, test_points):
    out = {"X": X, "Y": Y, "sizes": np.array(sizes), "lo": np.array(lo), "hi": np.array(hi),
           "beta": np.array(beta), "test_points": np.array(test_points)}
    pts = O.make_grid(lo, hi, [400, 400])
    for n in sizes:
        Xn, Yn = X[:n], Y[:n]
        _, _, _, _, X_norm, Y_norm = O.normalize(Xn, Yn)
        hyp = O.fit_hyper(X_norm, Y_norm, seed=1000 + n)
        ds = O.make_inference_datasets(Xn, Yn, hyp)
        m, v = O.posterior_inv(np.array(test_points), ds)
        st = O.safeopt_step(pts, ds, beta)
        gs = O.goose_step(pts, ds, beta)
        out[f"hyp_{n}"] = hyp
        out[f"tp_mean_{n}"] = m
        out[f"tp_var_{n}"] = v
        out[f"summary_{n}"] = np.array([st["S"].sum(), st["M"].sum(), st["Z"].sum(),
                                         st["minimizer_idx"], st["expander_idx"], st["x_new_idx"],
                                         gs["safe_min_idx"], gs["target_idx"], gs["x_new_idx"]], dtype=np.int64)
        out[f"scalars_{n}"] = np.array([st["min_ucb0"], st["minimizer_std"], st["expander_std"],
                                         st["L"][-1], gs["safe_min_lcb"], gs["target_lcb"]])
        print(name, n, "hyp", hyp.T.round(3).tolist(), "summary", out[f"summary_{n}"].tolist(),
              out[f"scalars_{n}"].round(5).tolist())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    Xb, Yb = run_of(os.path.join(REF, "data_SafeOpt_Benoit.npz"))
    tps = [[1.45698204, -0.76514894], [1.19497006, -0.74191489], [0.9, -0.6], [10.0, 10.0]]
    build("c1_benoit", Xb, Yb, [4, 9, 14], [-0.6, -1.0], [1.5, 1.0], 3.0, tps)
    Xw, Yw = run_of(os.path.join(REF, "data_multi_SafeOpt_WilliamOttoReactor.npz"))
    tpw = [[6.9, 83.0], [5.5, 80.0], [4.5, 75.0], [7.0, 100.0]]
    build("c3_wor", Xw, Yw, [5, 20, 35], [4.0, 70.0], [7.0, 100.0], 2.0, tpw)
