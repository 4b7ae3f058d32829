"""CPU, world_size = 2, gloo: the multi-GPU orchestration of sbo_b200.sharded (collectives, candidate
all-gather, offsets, deterministic reductions) driven by a NumPy stand-in for GridEngine that is built from
the oracle.  The kernels themselves are covered by the -m gpu tests; here the question is whether the sharded
step returns the same answer as the single-process oracle step."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleEngine:
    """Implements the slice of the GridEngine interface that sharded.py uses, with the oracle's arithmetic,
    on the block-cyclic shard of `rank`."""

    def __init__(self, O, rank, world, lo, hi, pts, block=32):
        self.O, self.rank, self.world = O, rank, world
        self.points_all = O.make_grid(lo, hi, pts)
        N = self.points_all.shape[0]
        nblk = (N + block - 1) // block
        sb = np.arange((nblk + world - 1) // world)            # rotated block-cyclic, as sbo_set_shard_cyclic
        mine = sb * world + (rank + sb + sb // world + sb // (world * world)) % world
        mine = mine[mine < nblk]
        self.gidx = np.concatenate([np.arange(b * block, min(N, (b + 1) * block)) for b in mine])
        self.points = self.points_all[self.gidx]
        self.device = 0
        self.count = self.gidx.size
        self.block, self.N = block, N
        self.seg, self.gz = None, None

    @staticmethod
    def shard_gidx(rank, world, N, block):
        nblk = (N + block - 1) // block
        sb = np.arange((nblk + world - 1) // world)
        mine = sb * world + (rank + sb + sb // world + sb // (world * world)) % world
        mine = mine[mine < nblk]
        return np.concatenate([np.arange(b * block, min(N, (b + 1) * block)) for b in mine])

    # ---- sharded Lipschitz expander: segments, mask export, all-gathered unsafe set (sbo_pairs_set_segments,
    # sbo_mask_export_dev, sbo_pairs_set_global_unsafe_dev)
    def pairs_set_segments(self, n_per_rank, rank):
        self.seg = (np.asarray(n_per_rank, dtype=np.int64), int(rank))

    def mask_export(self, kind, dst, which=0):
        assert kind == 2                                        # MASK_UNSAFE
        bits = np.zeros(dst.numel() * 32, dtype=np.uint8)
        bits[: self.count] = self.Z
        dst.copy_(torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int32).copy()))

    def pairs_set_global_unsafe(self, gathered, words_per_rank, nranks):
        w = gathered.numpy().view(np.uint32).reshape(nranks, words_per_rank)
        Zg = np.zeros(self.N, dtype=bool)
        for r in range(nranks):
            g = self.shard_gidx(r, nranks, self.N, self.block)
            bits = np.unpackbits(w[r].view(np.uint8), bitorder="little")[: g.size].astype(bool)
            Zg[g] = bits
        self.gz = np.flatnonzero(Zg)

    def set_model(self, ds):
        self.ds = ds
        self.G = ds["Y_norm"].shape[1]
        self.d = ds["X_norm"].shape[1]

    def posterior(self, with_grad=False, keep_v=0, fetch=True):
        self.mean, self.var, self.V = self.O.posterior_chol(self.points, self.ds, return_V=True)
        self.with_grad = with_grad

    def lipschitz(self):
        return np.array([self.O.lipschitz_constant(self.points, self.ds, i) for i in range(self.G)])

    def _g(self, i):
        return int(self.gidx[i]) if i >= 0 else -1

    def sets_pass1(self, beta, rule):
        self.beta = beta
        self.lcb, self.ucb = self.O.bounds(self.mean, self.var, beta)
        self.S = self.O.safe_mask(self.lcb)
        self.Z = self.O.unsafe_mask(self.lcb, "all" if rule == 0 else "any")
        iu, vu = self.O.minimize_obj_ucb(self.ucb, self.S)
        il, vl = self.O.minimize_obj_lcb(self.lcb, self.S)
        return {"min_ucb0": vu, "min_ucb0_idx": self._g(iu), "min_lcb0": vl, "min_lcb0_idx": self._g(il),
                "n_safe": int(self.S.sum()), "n_unsafe": int(self.Z.sum())}

    def sets_pass2(self, min_ucb0):
        self.M = self.S & (self.lcb[:, 0] <= min_ucb0)
        i, v = self.O.masked_argmax(self.var[:, 0], self.M)
        return {"minimizer_var": v, "minimizer_idx": self._g(i), "n_min": int(self.M.sum())}

    def point_coords(self, idx):
        return self.points_all[idx]

    def argreduce(self, kind, mask_kind, which, target):
        i, dval = self.O.explore_safeset(self.points, self.S, target)
        return self._g(i), dval

    # ---- staged pair driver
    def pairs_prepare(self, mode, prec, beta, L):
        self.mode, self.L = mode, (None if L is None else np.asarray(L))
        self.xs, self.zs = np.flatnonzero(self.S), np.flatnonzero(self.Z)
        nc, n = self.G - 1, self.ds["X_norm"].shape[0]
        self.seg, self.gz = None, None
        return {"n_x_local": self.xs.size, "n_z_local": self.zs.size, "row_doubles": 2 * self.d + 3 * nc + 1,
                "vrow_bytes": nc * n * 8 if mode == 1 else 0}

    def pairs_export(self, rows, vrows):
        O, ds, d, nc = self.O, self.ds, self.d, self.G - 1
        x = self.points[self.xs]
        xn = (x - ds["X_mean"]) / ds["X_std"]
        r = np.zeros((self.xs.size, 2 * d + 3 * nc + 1))
        r[:, :d] = x
        r[:, -1] = self.gidx[self.xs]
        r[:, d + nc:2 * d + nc] = xn
        for c in range(nc):
            _, _, sn2 = O.unpack_hyper(ds["hypopt"][:, c + 1], d)
            vn = self.var[self.xs, c + 1] / ds["Y_std"][c + 1] ** 2
            den = vn + sn2 + O.EPS_F32
            r[:, d + c] = self.ucb[self.xs, c + 1]
            r[:, 2 * d + nc + c] = self.beta * np.sqrt(vn) / den
            r[:, 2 * d + 2 * nc + c] = 1.0 / den
        rows[: self.xs.size] = torch.from_numpy(r)
        if vrows is not None and self.xs.size:
            v = np.stack([self.V[c + 1][self.xs] for c in range(nc)], axis=1)          # (nx, nc, n)
            vrows[: self.xs.size] = torch.from_numpy(np.ascontiguousarray(v).view(np.uint8).reshape(self.xs.size, -1))

    def pairs_import(self, n_total, rows, vrows):
        self.rows = rows.numpy()[:n_total].copy()       # the buffers hold one placeholder row when n_total == 0
        self.n_total = n_total
        if vrows is not None:
            nc, n = self.G - 1, self.ds["X_norm"].shape[0]
            self.Vx = vrows.numpy()[:n_total].copy().view(np.float64).reshape(n_total, nc, n)

    def pairs_run(self, goose, result):
        O, ds, d, nc = self.O, self.ds, self.d, self.G - 1
        z = self.points[self.zs]
        res = result.numpy()
        by_cand = self.mode == 0 and not goose and self.gz is not None
        if self.n_total == 0 or nc == 0 or (self.zs.size == 0 and not by_cand):
            return
        x = self.rows[:, :d]
        if by_cand:
            # this rank's share of the candidate tiles (grid order, 256 per tile, dealt round-robin) x ALL unsafe points
            world, rank = self.seg[0].size, self.seg[1]
            order = np.argsort(self.rows[:, -1], kind="stable")
            mine = np.concatenate([order[t:t + 256] for t in range(rank * 256, self.n_total, world * 256)] or [np.zeros(0, int)])
            zg = self.points_all[self.gz]
            if mine.size and zg.shape[0]:
                diff = x[mine][:, None, :] - zg[None, :, :] + O.PAIR_OFFSET
                dist_ = np.sqrt(np.sum(diff * diff, axis=2))
                for c in range(nc):
                    reach = (self.rows[mine, d + c][:, None] - self.L[c + 1] * dist_) >= 0.0
                    res[c * self.n_total + mine] = reach.any(axis=1)
        elif self.mode == 0:
            diff = x[:, None, :] - z[None, :, :] + O.PAIR_OFFSET
            dist_ = np.sqrt(np.sum(diff * diff, axis=2))
            for c in range(nc):
                reach = (self.rows[:, d + c][:, None] - self.L[c + 1] * dist_) >= 0.0
                if goose:
                    res[c * self.zs.size:(c + 1) * self.zs.size] = reach.any(axis=0)
                else:
                    res[c * self.n_total:(c + 1) * self.n_total] = reach.any(axis=1)
        else:
            zn = (z - ds["X_mean"]) / ds["X_std"]
            xn = self.rows[:, d + nc:2 * d + nc]
            ok = np.ones((self.zs.size, self.n_total), dtype=bool)
            for c in range(nc):
                ell, sf2, _ = O.unpack_hyper(ds["hypopt"][:, c + 1], d)
                kzx = sf2 * np.exp(-0.5 * O.sq_dist_direct(zn, xn, ell))
                cov = kzx - self.V[c + 1][self.zs] @ self.Vx[:, c, :].T
                mu = (self.mean[self.zs, c + 1] / ds["Y_std"][c + 1])[:, None] + cov * self.rows[:, 2 * d + nc + c][None, :]
                s2 = (self.var[self.zs, c + 1] / ds["Y_std"][c + 1] ** 2)[:, None] - cov * cov * self.rows[:, 2 * d + 2 * nc + c][None, :]
                ok &= (mu - self.beta * np.sqrt(np.maximum(s2, 0.0))) >= 0.0
            res[: self.n_total] = ok.sum(axis=0)

    def pairs_finish(self, goose, offset, result, want_counts=False):
        O, nc = self.O, self.G - 1
        res = result.numpy()
        per_v, per_i, hit_any = [], [], np.zeros(self.count, dtype=bool)
        nmask = 1 if self.mode == 1 else nc
        for c in range(nmask):
            mask = np.zeros(self.count, dtype=bool)
            if goose:
                mask[self.zs] = res[c * self.zs.size:(c + 1) * self.zs.size].astype(bool)
                i, v = O.masked_argmin(self.lcb[:, 0], mask)
            else:
                if self.mode == 1:
                    mask[self.xs] = res[offset:offset + self.xs.size] > 0
                else:
                    mask[self.xs] = res[c * self.n_total + offset: c * self.n_total + offset + self.xs.size].astype(bool)
                i, v = O.masked_argmax(self.var[:, 0], mask)
            hit_any |= mask
            per_v.append(v)
            per_i.append(self._g(i))
        return {"per_value": per_v, "per_idx": per_i, "n_hit": int(hit_any.sum()), "pairs_evaluated": 0}


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import gp_oracle as O
        import sbo_b200  # noqa: F401
        from sbo_b200 import sharded
        from conftest import load_golden, golden_ds
        gold = load_golden(case["gold"])
        ds = golden_ds(O, gold, case["n"])
        eng = OracleEngine(O, rank, world, gold["lo"], gold["hi"], case["grid"])
        dev = torch.device("cpu")
        st = sharded.safeopt_step(eng, ds, case["beta"], mode=case["mode"], precision="fp64", device=dev)
        gs = sharded.goose_step(eng, ds, case["beta"], device=dev) if case["mode"] == "lipschitz" else None
        keys = ["n_safe", "n_unsafe", "n_min", "minimizer_idx", "expander_idx", "x_new_idx", "min_ucb0", "minimizer_std", "expander_std"]
        res = {k: st[k] for k in keys}
        res["n_hit"] = st["expander"]["n_hit"]
        if gs is not None:
            res.update({"g_" + k: gs[k] for k in ["min_lcb0_idx", "target_idx", "x_new_idx", "target_lcb"]})
        q.put((rank, res))
    except Exception as e:      # surface the failure instead of letting the parent time out
        import traceback
        q.put((rank, {"error": traceback.format_exc()}))
        raise
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("case", [
    {"gold": "c1_benoit", "n": 9, "beta": 3.0, "grid": [36, 30], "mode": "lipschitz"},
    {"gold": "c3_wor", "n": 20, "beta": 2.0, "grid": [31, 37], "mode": "lipschitz"},
    {"gold": "c3_wor", "n": 20, "beta": 2.0, "grid": [22, 19], "mode": "fantasy"},
    # three ranks, ragged last block and a partial last super-block of the rotated block-cyclic shards
    {"gold": "c3_wor", "n": 35, "beta": 2.0, "grid": [29, 23], "mode": "lipschitz", "world": 3},
    # empty safe set on every rank (huge beta): no candidates are exchanged, every optimum is "none"
    {"gold": "c1_benoit", "n": 4, "beta": 500.0, "grid": [20, 17], "mode": "lipschitz"},
    {"gold": "c1_benoit", "n": 4, "beta": 500.0, "grid": [20, 17], "mode": "fantasy"},
])
def test_sharded_step_matches_single_process_oracle(oracle, case):
    world = case.get("world", 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for r in results.values():
        assert "error" not in r, r.get("error")
    for rk in range(1, world):
        assert results[0] == results[rk]                 # every rank returns the same global answer
    from conftest import load_golden, golden_ds
    gold = load_golden(case["gold"])
    ds = golden_ds(oracle, gold, case["n"])
    pts = oracle.make_grid(gold["lo"], gold["hi"], case["grid"])
    so = oracle.safeopt_step(pts, ds, case["beta"], mode=case["mode"], form="chol")
    r = results[0]
    assert r["n_safe"] == so["S"].sum() and r["n_unsafe"] == so["Z"].sum() and r["n_min"] == so["M"].sum()
    assert r["minimizer_idx"] == so["minimizer_idx"] and r["expander_idx"] == so["expander_idx"]
    assert r["x_new_idx"] == so["x_new_idx"]
    assert r["min_ucb0"] == pytest.approx(so["min_ucb0"], rel=1e-12)
    assert r["n_hit"] == np.any(so["expander_masks"], axis=0).sum()
    if case["mode"] == "lipschitz":
        go = oracle.goose_step(pts, ds, case["beta"], form="chol")
        assert r["g_min_lcb0_idx"] == go["safe_min_idx"] and r["g_target_idx"] == go["target_idx"]
        assert r["g_x_new_idx"] == go["x_new_idx"]


def test_reduce_arg_and_first_best():
    import sbo_b200  # noqa: F401
    from sbo_b200 import sharded
    assert sharded.first_best([1.0, 3.0, 3.0], [5, 9, 2], maximize=True) == (9, 3.0)      # first maximum wins
    assert sharded.first_best([2.0, 1.0, 1.0], [5, 9, 2], maximize=False) == (9, 1.0)
    assert sharded.first_best([np.inf, np.inf], [-1, -1], maximize=False) == (-1, np.inf)
