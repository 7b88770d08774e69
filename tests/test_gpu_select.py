"""GPU parity: selection / compaction kernels vs the oracle (bit exact; integer and index work)."""
import numpy as np
import pytest
import torch

from oracle import strainer_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import strainer_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return strainer_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def adversarial(n, rng):
    """ties, -0.0, 100.0 clamps, denormals"""
    v = O.synth_losses(n, seed=int(rng.integers(1 << 30)))
    m = rng.random(n)
    v[m < 0.3] = np.float32(0.25)          # 30 % exact ties
    v[(m >= 0.3) & (m < 0.35)] = np.float32(-0.0)
    v[(m >= 0.35) & (m < 0.4)] = np.float32(0.0)
    v[(m >= 0.4) & (m < 0.45)] = np.float32(100.0)
    v[(m >= 0.45) & (m < 0.5)] = np.float32(5.9604645e-08)
    v[(m >= 0.5) & (m < 0.52)] = np.float32(-1.5)
    return v


def test_radix_select_order_stats(sb):
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 31, 1000, 4097, 70001, 1_000_003):
        for trial in range(3):
            v = adversarial(n, rng) if trial else rng.standard_normal(n).astype(np.float32)
            s = np.sort(v)
            for k in {0, n // 2, max(n - 2, 0), n - 1, int(rng.integers(0, n))}:
                got = sb.order_stats(dev(v), k).cpu().numpy()
                want = np.array([s[k], s[min(k + 1, n - 1)]], np.float32)
                assert np.array_equal(got, want), (n, k, got, want)  # -0.0 == 0.0 under array_equal


def test_onepass_select_large(sb):
    """n >= 2^21 takes the sampled one-pass path (sg_select_kth): pivots from a sample, one streaming
    read, radix passes over the candidates; the fallbacks (ties overflowing the candidate buffer, sorted
    input, extreme ranks, NaN) must stay exact."""
    rng = np.random.default_rng(21)
    n = (1 << 22) + 12345
    cases = {
        "lognormal": O.synth_losses(n, seed=3),
        "normal": rng.standard_normal(n).astype(np.float32),
        "adversarial_ties": adversarial(n, rng),
        "all_equal": np.full(n, 0.75, np.float32),
        "sorted": np.sort(rng.standard_normal(n).astype(np.float32)),
        "reverse_sorted": np.sort(rng.standard_normal(n).astype(np.float32))[::-1].copy(),
        "two_values": (rng.random(n) < 0.5).astype(np.float32),
    }
    for name, v in cases.items():
        s = np.sort(v)
        d = dev(v)
        for k in (0, 1, n // 10, n // 2, (9 * n) // 10, n - 2, n - 1, int(rng.integers(0, n))):
            got = sb.order_stats(d, k).cpu().numpy()
            want = np.array([s[k], s[min(k + 1, n - 1)]], np.float32)
            assert np.array_equal(got, want), (name, k, got, want)
    # unaligned view, percentile on top
    v = O.synth_losses(n + 1, seed=4)
    d = dev(v)[1:]
    assert sb.percentile_device(d, 90.0).cpu().numpy()[0] == np.percentile(v[1:], 90.0)
    # NaN anywhere -> NaN
    v = O.synth_losses(n, seed=5)
    v[n // 3] = np.nan
    assert np.isnan(sb.order_stats(dev(v), n // 2).cpu().numpy()).all()


def test_radix_select_nan(sb):
    v = np.arange(100, dtype=np.float32)
    v[17] = np.nan
    assert np.isnan(sb.order_stats(dev(v), 50).cpu().numpy()).all()
    assert np.isnan(sb.percentile_device(dev(v), 90.0).cpu().numpy()[0])


def test_percentile_bit_exact_vs_numpy(sb):
    rng = np.random.default_rng(12)
    for _ in range(60):
        n = int(rng.integers(1, 20000))
        q = [float(rng.uniform(0, 100)), (1 - 0.8) * 100, (1 - 0.2) * 100, 90.0, 75, 25, 0.0, 100.0][int(rng.integers(0, 8))]
        v = adversarial(n, rng) if rng.random() < 0.5 else rng.standard_normal(n).astype(np.float32)
        got = sb.percentile_device(dev(v), q).cpu().numpy()[0]
        want = np.percentile(v, q)
        assert got.dtype == want.dtype == np.float32
        assert got == want, (n, q, got, want)
    v = O.synth_losses(1 << 20, seed=5)
    for q in (90.0, (1 - 0.8) * 100):
        assert sb.percentile_device(dev(v), q).cpu().numpy()[0] == np.percentile(v, q)
    # np.float64 q: numpy interpolates in float64
    got = sb.percentile_device(dev(v), np.float64(33.3))
    assert float(got[0]) == np.percentile(v, np.float64(33.3))


def test_quantile_bit_exact_vs_torch(sb):
    rng = np.random.default_rng(13)
    for n in (2, 64, 128, 256, 512, 1000, 2048, 2049, 4096, 65536):
        for q in (0.1, 0.9, 0.8731, float(rng.uniform(0, 1)), 0.0, 1.0):
            v = rng.standard_normal(n).astype(np.float32)
            if n == 128:
                v = np.round(v, 1)
            got = sb.quantile_device(dev(v), q).cpu().numpy()[0]
            want = torch.quantile(torch.from_numpy(v), q).numpy()
            assert got == want, (n, q, got, want)


@pytest.mark.parametrize("n", [0, 1, 5, 4096, 4097, 100_000, 1_000_003])
def test_compact_indices_vs_np_where(sb, n):
    rng = np.random.default_rng(14 + n)
    v = adversarial(n, rng) if n else np.zeros(0, np.float32)
    if n > 10:
        v[7] = np.nan
    for cmp_code, fn in ((0, np.less), (1, np.less_equal), (2, np.greater_equal), (3, np.greater)):
        for thr in (np.float32(0.25), np.float32(-0.0), np.float32(1e9), np.float32(-1e9)):
            idx, count, mask = sb.compact_indices(dev(v) if n else torch.zeros(0, device="cuda"), float(thr), cmp_code, 0, True)
            want = np.where(fn(v, thr))[0]
            c = int(count.item())
            assert c == len(want)
            assert np.array_equal(idx[:c].cpu().numpy(), want)
            assert np.array_equal(mask.cpu().numpy().astype(bool), fn(v, thr))
            idx2, count2, _ = sb.compact_indices(dev(v) if n else torch.zeros(0, device="cuda"), float(thr), cmp_code | 4, 1000)
            assert np.array_equal(idx2[:int(count2.item())].cpu().numpy(), np.where(~fn(v, thr))[0] + 1000)


def test_partition_rows_vs_boolean_index(sb):
    rng = np.random.default_rng(15)
    for n, shape in ((128, (3, 64, 64)), (7, (4,)), (5000, (8,)), (1, (3, 64, 64))):
        x = torch.from_numpy(rng.standard_normal((n,) + shape).astype(np.float32))
        m = torch.from_numpy(rng.random(n) < 0.9)
        kept, dropped, counts = sb.partition_rows(x.cuda(), m.cuda())
        c = counts.cpu().numpy()
        assert c[0] == int(m.sum()) and c[1] == n - int(m.sum())
        assert torch.equal(kept[:c[0]].cpu(), x[m]) and torch.equal(dropped[:c[1]].cpu(), x[~m])


def test_select_below_percentile_golden(sb, golden):
    """the selection half of refine_dataset_by_loss on the reference's own losses -> its own indices"""
    losses = golden["g1_losses"]
    for tag in "abc":
        q = (1 - float(golden[f"g1{tag}_ratio"])) * 100
        idx, thr = sb.select_below_percentile(dev(losses), q)
        assert thr == golden[f"g1{tag}_threshold"] and thr.dtype == np.float32
        assert np.array_equal(idx, golden[f"g1{tag}_indices"])


def test_full_size_properties(sb):
    """N = 2**20 (config 5): properties that do not need the oracle at full size."""
    n = 1 << 20
    rng = np.random.default_rng(16)
    v = adversarial(n, rng)
    d = dev(v)
    idx, thr = sb.select_below_percentile(d, 90.0)
    assert thr == np.percentile(v, 90.0)
    assert np.all(np.diff(idx) > 0)                       # ascending, unique
    assert np.all(v[idx] < thr) and len(idx) == int((v < thr).sum())
    # idempotence: straining the kept set with q=100 removes only the maximum ties
    kept = v[idx]
    idx2, thr2 = sb.select_below_percentile(dev(kept), 100.0)
    assert thr2 == kept.max() and len(idx2) == int((kept < kept.max()).sum())
    # all-equal losses: nothing survives (reference fallback case)
    idx3, thr3 = sb.select_below_percentile(torch.full((n,), 0.5, device="cuda"), 90.0)
    assert len(idx3) == 0 and thr3 == np.float32(0.5)


def test_sharded_select_emulated(sb):
    """Histogram phases on 4 shards with the all-reduce emulated by summation == single pass."""
    lib = sb._lib.load()
    L = sb._lib
    rng = np.random.default_rng(17)
    v = adversarial(50_000, rng)
    n = v.size
    k = 12345
    shards = [dev(s) for s in np.array_split(v, 4)]
    wss = [torch.empty(L.SG_SELECT_WS_WORDS, dtype=torch.int32, device="cuda") for _ in shards]
    st = L.P(torch.cuda.current_stream().cuda_stream)
    p = lambda t: L.P(t.data_ptr())
    for ws in wss:
        L.check(lib.sg_select_begin(p(ws), k, st))
    NW = L.SG_SELECT_WS_NANCOUNT + 1
    MA = L.SG_SELECT_WS_MINABOVE
    for ps in range(L.SG_SELECT_NUM_PASSES):
        for ws, sh in zip(wss, shards):
            L.check(lib.sg_select_hist(p(sh), sh.numel(), p(ws), ps, st))
        tot = sum(ws[:NW].clone() for ws in wss)
        if ps == L.SG_SELECT_NUM_PASSES - 1:
            mins = torch.stack([ws[MA] ^ -2147483648 for ws in wss]).min() ^ -2147483648
        for ws in wss:
            ws[:NW] = tot
            if ps == L.SG_SELECT_NUM_PASSES - 1:
                ws[MA] = mins
            L.check(lib.sg_select_step(p(ws), ps, st))
    outs = []
    for ws in wss:
        o = torch.empty(2, device="cuda")
        L.check(lib.sg_select_finish(p(ws), p(o), st))
        outs.append(o.cpu().numpy())
    s = np.sort(v)
    for o in outs:
        assert np.array_equal(o, s[k:k + 2])


def test_device_sort(sb):
    rng = np.random.default_rng(41)
    for n in (1, 2, 33, 2048, 2049, 100_003, 1 << 20):
        v = adversarial(n, rng) if n > 2 else rng.standard_normal(n).astype(np.float32)
        if n > 100:
            v[11] = np.nan
        out, order = sb.sort_values(v, return_order=True)
        want_order = np.argsort(v, kind="stable")
        got = out.cpu().numpy()
        assert np.array_equal(got, np.sort(v), equal_nan=True)
        # stable: equal keys (incl. -0.0 / +0.0, which share a key) keep their input order
        o = order.cpu().numpy()
        assert np.array_equal(np.sort(o), np.arange(n))
        assert np.array_equal(v[o], v[want_order], equal_nan=True)
        same = v[o][1:] == v[o][:-1]
        assert np.all(o[1:][same] > o[:-1][same])


def test_device_sort_variants(sb):
    """keys only / order only, several waves of tiles (look-back across finished tiles), degenerate digit distributions."""
    L = sb._lib
    lib = L.init(0)
    st = L.P(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(43)
    n = (1 << 22) + 12345          # 1028 tiles: more than one wave of 296 CTAs
    cases = {
        "lognormal": np.exp(rng.standard_normal(n)).astype(np.float32),
        "all_equal": np.full(n, 0.25, np.float32),
        "ascending": np.arange(n, dtype=np.float32) - 1000.0,
        "descending": -np.arange(n, dtype=np.float32),
        "few_values": rng.choice(np.array([-1.5, 0.0, -0.0, 3.0, np.inf, -np.inf, np.nan], np.float32), n),
    }
    ws = torch.empty(lib.sg_sort_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    for name, v in cases.items():
        vd = torch.from_numpy(v).cuda()
        want_order = np.argsort(v, kind="stable")
        keys = torch.empty(n, dtype=torch.float32, device="cuda")
        L.check(lib.sg_sort_f32(L.P(vd.data_ptr()), n, L.P(keys.data_ptr()), L.P(0), L.P(ws.data_ptr()), st), name)
        assert np.array_equal(keys.cpu().numpy(), np.sort(v), equal_nan=True), name
        order = torch.empty(n, dtype=torch.int32, device="cuda")
        L.check(lib.sg_sort_f32(L.P(vd.data_ptr()), n, L.P(0), L.P(order.data_ptr()), L.P(ws.data_ptr()), st), name)
        o = order.cpu().numpy()
        if name == "few_values":   # -0.0 and +0.0 share a key: compare through the values and check stability
            assert np.array_equal(v[o], v[want_order], equal_nan=True)
            same = (v[o][1:] == v[o][:-1]) | (np.isnan(v[o][1:]) & np.isnan(v[o][:-1]))
            assert np.all(o[1:][same] > o[:-1][same])
        else:
            assert np.array_equal(o, want_order), name
        both_k = torch.empty(n, dtype=torch.float32, device="cuda")
        both_o = torch.empty(n, dtype=torch.int32, device="cuda")
        L.check(lib.sg_sort_f32(L.P(vd.data_ptr()), n, L.P(both_k.data_ptr()), L.P(both_o.data_ptr()), L.P(ws.data_ptr()), st), name)
        assert torch.equal(both_k.view(torch.int32), keys.view(torch.int32)) and torch.equal(both_o, order), name


def test_dbscan1d_vs_sklearn(sb):
    rng = np.random.default_rng(42)
    for n, eps, ms in ((50, 0.05, 3), (2000, 0.01, 3), (2000, 0.002, 5), (300, 0.5, 3), (5, 0.1, 3), (20000, 0.0005, 4)):
        v = O.synth_losses(n, seed=int(rng.integers(1 << 30)))
        ratio, noise = sb.dbscan1d_clean_ratio(v, eps, ms, return_noise=True)
        want = O.dbscan1d_noise_sklearn(v, eps, ms)
        assert np.array_equal(noise.cpu().numpy(), want), (n, eps, ms)
        assert ratio == np.sum(~want) / n
    v = np.round(O.synth_losses(3000, seed=3), 2)      # heavy ties, distances exactly == eps
    ratio, noise = sb.dbscan1d_clean_ratio(v, 0.01, 3, return_noise=True)
    assert np.array_equal(noise.cpu().numpy(), O.dbscan1d_noise_sklearn(v, 0.01, 3))
    # many tiles, windows from 1 to tens of thousands of points, ties, huge magnitudes (scikit-learn rejects inf / NaN
    # inputs); oracle: the sort-based numpy form, itself checked against scikit-learn in tests/test_oracle_golden.py
    big = np.concatenate([O.synth_losses(300_000, seed=9), np.full(5000, 0.5, np.float32),
                          np.array([3e38, 3e38, -3e38, 1e30, -1e30], np.float32)])
    rng.shuffle(big)
    for eps, ms in ((1e-6, 3), (1e-4, 50), (0.0, 2), (3.0, 100)):
        ratio, noise = sb.dbscan1d_clean_ratio(big, eps, ms, return_noise=True)
        want = O.dbscan1d_noise(big, eps, ms)
        assert np.array_equal(noise.cpu().numpy(), want), (eps, ms)
        assert ratio == np.sum(~want) / big.size
        assert sb.dbscan1d_clean_ratio(big, eps, ms) == ratio      # keys-only sort path (no per-sample flags)
    # z-score + 1-D DBSCAN straining (config 3): clean_ratio -> torch.quantile(max_z, ratio), <=
    mz = O.zscore_max_torch(torch.from_numpy(O.synth_features(4096)))
    r = sb.dbscan1d_clean_ratio(mz, 0.05, 3)
    assert r == O.dbscan1d_clean_ratio(mz.numpy(), 0.05, 3)


def test_gmm_em_device(sb):
    """deterministic 2-component EM on the device vs its float64 restatement (tight) and vs a
    seeded scikit-learn fit of the reference call (same optimum on a bimodal loss vector)."""
    from sklearn.mixture import GaussianMixture
    for seed, n in ((1, 20000), (2, 65536), (3, 1000)):
        v = O.synth_losses(n, seed=seed)
        got = sb.gmm_fit_device(v)
        want = O.gmm_fit_deterministic(v)
        assert got["n_iter"] == want["n_iter"] and got["converged"] == want["converged"]
        for k in ("weights", "means", "stds"):
            assert np.allclose(got[k], want[k], rtol=1e-9, atol=1e-12), (k, got[k], want[k])
        np.random.seed(0)
        g = GaussianMixture(n_components=2, max_iter=10, tol=1e-2, reg_covar=5e-4).fit(v.reshape(-1, 1))
        sm, ss = g.means_.flatten().astype(np.float64), np.sqrt(g.covariances_.flatten()).astype(np.float64)
        o = np.argsort(sm)
        o2 = np.argsort(got["means"])
        assert np.allclose(got["means"][o2], sm[o], rtol=5e-2), (got["means"], sm)
        assert np.allclose(got["stds"][o2], ss[o], rtol=1e-1), (got["stds"], ss)
        t_dev = sb.get_gmm_threshold(v, fit="device")
        t_ref = O.gmm_intersection(sm.astype(np.float32), ss.astype(np.float32))
        assert abs(t_dev - t_ref) <= 5e-2 * abs(t_ref), (t_dev, t_ref)
    ds = torch.utils.data.TensorDataset(torch.zeros(1000, 1))
    clean, noisy = sb.divide_dataset(v, ds, fit="device")
    thr = sb.get_gmm_threshold(v, fit="device")
    wc, wn = O.divide_by_threshold(v, thr)
    assert np.array_equal(np.asarray(clean.indices), wc) and np.array_equal(np.asarray(noisy.indices), wn)


def test_strain_rows_edges(sb):
    """fused in-batch selection (sg_strain_rows): smallest / largest batch, ties at the threshold, NaN score"""
    rng = np.random.default_rng(9)
    for n in (1, 2, 3, 64, 777, 2048):
        rows = torch.from_numpy(rng.standard_normal((n, 3, 4, 4)).astype(np.float32)).cuda()
        s = rng.standard_normal(n).astype(np.float32)
        if n >= 64:
            s[: n // 3] = np.float32(0.25)                     # ties
        for q in (0.1, 0.5, 0.0, 1.0):
            kept, dropped, mask, thr = sb.strain_scores(rows, torch.from_numpy(s).cuda(), q)
            wk, wd, wm, wt = O.strain_scores(rows.cpu(), torch.from_numpy(s), q)
            assert float(thr) == float(wt) and torch.equal(mask.cpu(), wm), (n, q)
            assert torch.equal(kept.cpu(), wk) and torch.equal(dropped.cpu(), wd)
    s = rng.standard_normal(100).astype(np.float32)
    s[7] = np.nan
    rows = torch.zeros(100, 4, device="cuda")
    kept, dropped, mask, thr = sb.strain_scores(rows, torch.from_numpy(s).cuda(), 0.1)
    assert np.isnan(float(thr)) and kept.shape[0] == 0 and dropped.shape[0] == 100      # s >= nan is False everywhere
