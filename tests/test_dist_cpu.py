"""CPU, world_size 2 over gloo: the multi-rank selection protocol.  The CUDA phases are
replaced by a numpy double with the kernels' exact semantics (radix keys, 11/11/10-bit digit
histograms, min-above), so what is under test is the host orchestration in api.order_stats /
percentile_device: which words are all-reduced, with which op, in which order -- and that every
rank ends with bit-identical order statistics equal to the single-process result."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import strainer_oracle as O


def f2key(v):
    b = np.asarray(v, np.float32).view(np.uint32).copy()
    nan = (b & 0x7FFFFFFF) > 0x7F800000
    b[b == 0x80000000] = 0
    k = np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)
    k[nan] = 0xFFFFFFFF
    return k


def key2f(k):
    k = np.uint32(k)
    if k == 0xFFFFFFFF:
        return np.float32(np.nan)
    b = (k & np.uint32(0x7FFFFFFF)) if (k & np.uint32(0x80000000)) else ~k
    return np.array([b], np.uint32).view(np.float32)[0]


class FakeSelectOps:
    """numpy restatement of sel::begin/hist<PASS>/step/finish (csrc/select.cu): four 8-bit passes, the
    last one also tracks the smallest key above the selected 24-bit bucket."""
    NAN, MINABOVE, PREFIX, KREM, SELKEY, NEEDNEXT, NEXTIN = 256, 257, 260, 262, 264, 265, 266

    def _u(self, ws):
        return ws.numpy().view(np.uint32)

    def begin(self, ws, k):
        u = self._u(ws)
        u[:] = 0
        u[self.MINABOVE] = 0xFFFFFFFF
        u[self.NEXTIN] = 0xFFFFFFFF
        u[self.KREM:self.KREM + 2].view(np.uint64)[0] = k

    def hist(self, values, ws, p):
        u = self._u(ws)
        key = f2key(values.numpy())
        prefix = u[self.PREFIX]
        shift = 24 - 8 * p
        if p == 0:
            sel = np.ones(key.shape, bool)
            u[self.NAN] += np.uint32((key == 0xFFFFFFFF).sum())
        else:
            hi = key >> np.uint32(shift + 8)
            sel = hi == prefix
            if p == 3:
                above = key[hi > prefix]
                if above.size:
                    u[self.MINABOVE] = min(u[self.MINABOVE], above.min())
        dig = (key >> np.uint32(shift)) & 0xFF
        u[:256] += np.bincount(dig[sel], minlength=256).astype(np.uint32)

    def step(self, ws, p):
        u = self._u(ws)
        k = int(u[self.KREM:self.KREM + 2].view(np.uint64)[0])
        h = u[:256].astype(np.uint64)
        c = np.cumsum(h)
        b = int(np.searchsorted(c, k, side="right"))
        before = int(c[b - 1]) if b else 0
        u[self.PREFIX] = b if p == 0 else ((int(u[self.PREFIX]) << 8) | b) & 0xFFFFFFFF
        u[self.KREM:self.KREM + 2].view(np.uint64)[0] = k - before
        if p == 3:
            u[self.SELKEY] = u[self.PREFIX]
            u[self.NEEDNEXT] = 1 if (k - before + 1 >= int(h[b])) else 0
            nz = np.nonzero(h[b + 1:])[0]
            u[self.NEXTIN] = ((int(u[self.PREFIX]) & ~0xFF) | (b + 1 + int(nz[0]))) if nz.size else 0xFFFFFFFF
        u[:256] = 0

    def finish(self, ws, out2):
        u = self._u(ws)
        if u[self.NAN]:
            out2[:] = float("nan")
            return
        a = key2f(u[self.SELKEY])
        b = a
        if u[self.NEEDNEXT]:
            if u[self.NEXTIN] != 0xFFFFFFFF:
                b = key2f(u[self.NEXTIN])
            elif u[self.MINABOVE] != 0xFFFFFFFF:
                b = key2f(u[self.MINABOVE])
        out2[0], out2[1] = float(a), float(b)

    def lerp(self, stats2, gamma, kind):
        a, b = np.float32(stats2[0].item()), np.float32(stats2[1].item())
        return torch.tensor([O.np_lerp(a, b, np.float32(gamma))], dtype=torch.float32)


class FakeGmmOps:
    """numpy restatement of csrc/gmm.cu (accumulate / update kernels) on one rank's shard: per-rank partial sums
    in ``sums`` (all-reduced by api.gmm_fit_device), parameters in a 16-double state."""
    W, MU, VAR, LB, ITER, CONV, PHASE, KLEFT = 0, 2, 4, 6, 7, 8, 9, 10

    def __init__(self):
        self.st = np.zeros(16)
        self.sums = torch.zeros(8, dtype=torch.float64)

    def prepare(self, losses):
        return torch.as_tensor(np.asarray(losses, np.float32)).reshape(-1)

    def begin(self, centers, kmeans_iters):
        self.st[:] = 0
        self.st[self.MU:self.MU + 2] = centers.numpy().astype(np.float64)
        self.st[self.KLEFT] = kmeans_iters

    def accumulate(self, v):
        st, x = self.st, v.numpy().astype(np.float64)
        out = np.zeros(8)
        if st[self.PHASE] == 0:
            c1 = np.abs(x - st[self.MU + 1]) < np.abs(x - st[self.MU])
            out[0:3] = [(~c1).sum(), x[~c1].sum(), (x[~c1] ** 2).sum()]
            out[3:6] = [c1.sum(), x[c1].sum(), (x[c1] ** 2).sum()]
        elif st[self.PHASE] == 1:
            w, mu, var = st[0:2], st[2:4], st[4:6]
            lp = np.log(w) - 0.5 * (np.log(2 * np.pi) + np.log(var)) - (x[:, None] - mu) ** 2 * (0.5 / var)
            mx = lp.max(1)
            lse = mx + np.log(np.exp(lp - mx[:, None]).sum(1))
            r = np.exp(lp - lse[:, None])
            out[0:3] = [r[:, 0].sum(), (r[:, 0] * x).sum(), (r[:, 0] * x * x).sum()]
            out[3:6] = [r[:, 1].sum(), (r[:, 1] * x).sum(), (r[:, 1] * x * x).sum()]
            out[6] = lse.sum()
        else:
            return                                    # converged: the device kernel returns without touching sums
        self.sums[:] = torch.from_numpy(out)

    def update(self, n_total, reg_covar, tol, max_iter):
        st, s = self.st, self.sums.numpy()
        eps10 = 10 * np.finfo(np.float64).eps
        if st[self.PHASE] == 2:
            return

        def params():
            nk0, nk1 = s[0] + eps10, s[3] + eps10
            mu0, mu1 = s[1] / nk0, s[4] / nk1
            st[self.MU:self.MU + 2] = [mu0, mu1]
            st[self.VAR:self.VAR + 2] = [max(s[2] / nk0 - mu0 * mu0, 0) + reg_covar, max(s[5] / nk1 - mu1 * mu1, 0) + reg_covar]
            st[self.W:self.W + 2] = [nk0 / n_total, nk1 / n_total]
        if st[self.PHASE] == 0:
            m0 = s[1] / s[0] if s[0] > 0 else st[self.MU]
            m1 = s[4] / s[3] if s[3] > 0 else st[self.MU + 1]
            moved = (m0 != st[self.MU]) or (m1 != st[self.MU + 1])
            st[self.KLEFT] -= 1
            if moved and st[self.KLEFT] > 0:
                st[self.MU:self.MU + 2] = [m0, m1]
                return
            params()
            st[self.LB], st[self.ITER], st[self.PHASE] = -np.inf, 0, 1
            return
        lb = s[6] / n_total
        params()
        change = lb - st[self.LB]
        st[self.LB] = lb
        st[self.ITER] += 1
        if abs(change) < tol:
            st[self.CONV], st[self.PHASE] = 1, 2
        elif st[self.ITER] >= max_iter:
            st[self.PHASE] = 2

    def state(self):
        return self.st.copy()


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import strainer_b200 as sb
    rng = np.random.default_rng(123)
    n = 20000
    v = O.synth_losses(n, seed=77)
    m = rng.random(n)
    v[m < 0.3] = np.float32(0.25)
    v[(m > 0.3) & (m < 0.33)] = np.float32(-0.0)
    bounds = [0, 8192, n]                         # chunk-aligned, unequal shards
    shard = torch.from_numpy(v[bounds[rank]:bounds[rank + 1]].copy())
    s = np.sort(v)
    ok = True
    for k in (0, 1234, 5999, 6000, n - 2, n - 1):
        got = sb.order_stats(shard, k, dist.group.WORLD, FakeSelectOps()).numpy()
        want = np.array([s[k], s[min(k + 1, n - 1)]], np.float32)
        ok &= bool(np.array_equal(got, want))
    for q in (90.0, (1 - 0.8) * 100, 30.0, 100.0, 0.0):
        thr = sb.percentile_device(shard, q, dist.group.WORLD, n, FakeSelectOps()).numpy()[0]
        ok &= bool(thr == np.percentile(v, q))
        # kept indices: local compaction + global base; concatenation by rank == np.where order
        local = np.where(shard.numpy() < thr)[0] + bounds[rank]
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        ok &= bool(np.array_equal(np.concatenate(gathered), np.where(v < thr)[0]))
    vn = shard.clone()
    if rank == 1:
        vn[5] = float("nan")                       # a NaN on ONE rank poisons the global threshold on ALL ranks
    ok &= bool(np.isnan(sb.percentile_device(vn, 90.0, dist.group.WORLD, n, FakeSelectOps()).numpy()[0]))
    # sharded GMM EM: global order statistics for the start + an 8-double all-reduce per iteration; every rank must
    # end with the fit of the whole vector (float64 restatement in the oracle)
    vg = O.synth_losses(n, seed=78)
    g = sb.gmm_fit_device(vg[bounds[rank]:bounds[rank + 1]].copy(), group=dist.group.WORLD, n_global=n, ops=FakeGmmOps(),
                          select_ops=FakeSelectOps())
    want = O.gmm_fit_deterministic(vg)
    ok &= g["n_iter"] == want["n_iter"] and g["converged"] == want["converged"]
    for key in ("weights", "means", "stds"):
        ok &= bool(np.allclose(g[key], want[key], rtol=1e-10, atol=1e-13))
    open(os.path.join(tmp, f"ok{rank}"), "w").write(str(ok))
    dist.destroy_process_group()


def test_two_rank_select_protocol_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["True", "True"]


def test_fake_ops_match_single_process():
    """the numpy double itself against np.sort (so the gloo test checks the protocol, not the double)"""
    import strainer_b200 as sb
    rng = np.random.default_rng(5)
    v = rng.standard_normal(5000).astype(np.float32)
    v[::7] = 0.5
    s = np.sort(v)
    for k in (0, 17, 2500, 4998, 4999):
        got = sb.order_stats(torch.from_numpy(v), k, None, FakeSelectOps()).numpy()
        assert np.array_equal(got, [s[k], s[min(k + 1, 4999)]])
