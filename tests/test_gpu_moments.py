"""GPU parity: z-score, elbow histogram, detect_outliers variants, moments."""
import numpy as np
import pytest
import torch

from oracle import strainer_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import strainer_b200
    assert torch.cuda.is_available()
    return strainer_b200


@pytest.fixture(scope="module")
def feats():
    return O.synth_features(4096)


def test_zscore_max(sb, feats, golden):
    got = sb.zscore_max(torch.from_numpy(feats)).cpu().numpy()
    want = golden["g4_maxz"]
    assert np.allclose(got, want, rtol=2e-5, atol=1e-6), np.abs(got - want).max()   # fp32 reductions differ in the last bits
    got0 = sb.zscore_max(torch.from_numpy(feats), ddof=0, eps_add=1e-7).cpu().numpy()
    assert np.allclose(got0, golden["g7_maxz_np"], rtol=2e-5, atol=1e-6)
    f2 = feats.copy()
    f2[:, 5] = 1.0  # constant column -> 0/0 = NaN -> every row NaN
    assert np.isnan(sb.zscore_max(torch.from_numpy(f2)).cpu().numpy()).all()
    # feature widths other than the reference's 512 (vectorised and scalar column paths, ragged row counts)
    rng = np.random.default_rng(5)
    for n, d in ((1000, 64), (777, 100), (300, 128), (513, 1024), (129, 2048), (64, 7)):
        f = (rng.standard_normal((n, d)) * rng.uniform(0.5, 3, d) + rng.uniform(-2, 2, d)).astype(np.float32)
        got = sb.zscore_max(torch.from_numpy(f)).cpu().numpy()
        want = O.zscore_max_torch(torch.from_numpy(f)).numpy()
        assert np.allclose(got, want, rtol=2e-5, atol=1e-6), (n, d, np.abs(got - want).max())


def test_find_elbow_threshold_bit_exact(sb, golden):
    """given identical z-scores the histogram, density and threshold are bit exact"""
    thr, centers, hist = sb.find_elbow_threshold(golden["g4_maxz"])
    assert thr == golden["g4_threshold"]
    assert np.array_equal(centers, golden["g4_centers"]) and np.array_equal(hist, golden["g4_hist"])
    rng = np.random.default_rng(21)
    for n, scale in ((1, 1.0), (17, 1e-3), (100_003, 37.5), (5, 0.0)):
        z = (rng.standard_normal(n) * scale).astype(np.float32)
        t2, c2, h2 = sb.find_elbow_threshold(z)
        t1, c1, h1 = O.find_elbow_threshold(z)
        assert t1 == t2 and np.array_equal(c1, c2) and np.array_equal(h1, h2, equal_nan=True)
    with pytest.raises(ValueError):
        sb.find_elbow_threshold(np.array([1.0, np.nan], np.float32))
    # values on / one ulp around the fp32 linspace edges, and a range far from zero (coarse ulps): the
    # edge-corrected binning must still agree with np.histogram count for count
    for lo, hi in ((0.0, 7.3), (-3.0, 11.0), (1000.0, 1001.0), (-2.5e4, -2.4e4), (1e-3, 1.5e-3)):
        edges = np.linspace(np.float32(lo), np.float32(hi), 101, dtype=np.float32)
        pts = np.concatenate([edges, np.nextafter(edges, np.float32(np.inf)), np.nextafter(edges, np.float32(-np.inf)),
                              rng.uniform(lo, hi, 200_000).astype(np.float32)])
        pts = pts[(pts >= np.float32(lo)) & (pts <= np.float32(hi))].astype(np.float32)
        rng.shuffle(pts)
        t2, c2, h2 = sb.find_elbow_threshold(pts)
        t1, c1, h1 = O.find_elbow_threshold(pts)
        assert np.array_equal(h1, h2) and t1 == t2 and np.array_equal(c1, c2), (lo, hi)


def test_detect_outliers_variants(sb, feats, golden):
    ds = torch.utils.data.TensorDataset(torch.from_numpy(feats), torch.zeros(len(feats)))
    ident = torch.nn.Identity()
    mz = golden["g4_maxz"]

    def near(thr, tol=1e-4):
        return np.abs(mz - thr) <= tol * np.abs(thr)

    got = sb.detect_outliers(ds, ident)
    assert got.dtype == bool
    bad = got != golden["g5_elbow_inlier"]
    assert not (bad & ~near(float(golden["g4_threshold"]))).any()
    got = sb.detect_outliers(ds, ident, 5)
    assert not ((got != golden["g5_user5_inlier"]) & ~near(5.0)).any()
    got = sb.detect_outliers(ds, ident, threshold=5.0)
    assert isinstance(got, torch.Tensor) and got.dtype == torch.bool
    assert not ((got.numpy() != golden["g5_fixed5_inlier"]) & ~near(5.0)).any()
    for tag in "ab":
        r = float(golden[f"g5_ratio{tag}"])
        got = sb.detect_outliers(ds, ident, clean_ratio=r).numpy()
        thr = float(torch.quantile(torch.from_numpy(mz), r))
        assert not ((got != golden[f"g5_ratio{tag}_inlier"]) & ~near(thr)).any()
        assert abs(int(got.sum()) - int(golden[f"g5_ratio{tag}_inlier"].sum())) <= 2
    assert np.allclose(sb.compute_z_scores(ds, ident), golden["g7_maxz_np"], rtol=2e-5, atol=1e-6)


def test_moments_threshold(sb):
    lib = sb._lib.load()
    L = sb._lib
    rng = np.random.default_rng(22)
    for n in (2, 4096, 4097, 100_003):
        e = rng.random(n).astype(np.float32)
        d = torch.from_numpy(e).cuda()
        chunks = (n + L.SG_MOMENT_CHUNK - 1) // L.SG_MOMENT_CHUNK
        part = torch.empty(2 * chunks, dtype=torch.float64, device="cuda")
        stats = torch.empty(2, dtype=torch.float64, device="cuda")
        thr = torch.empty(1, dtype=torch.float32, device="cuda")
        st = L.P(torch.cuda.current_stream().cuda_stream)
        L.check(lib.sg_chunk_moments(L.P(d.data_ptr()), n, L.P(part.data_ptr()), st))
        L.check(lib.sg_moments_finish(L.P(part.data_ptr()), chunks, n, 2.0, L.P(stats.data_ptr()), L.P(thr.data_ptr()), st))
        te = torch.from_numpy(e)
        want = (te.mean() + 2.0 * te.std()).item()
        assert abs(thr.item() - want) <= 2e-6 * abs(want)          # fp32 tolerance: 2e-6 relative
        s = stats.cpu().numpy()
        assert abs(s[0] - e.astype(np.float64).mean()) < 1e-12 and abs(s[1] - e.astype(np.float64).std(ddof=1)) < 1e-10
        # sharding invariance: partials of chunk-aligned shards are the same numbers
        if n > 8192:
            half = 4096 * (chunks // 2)
            p2 = torch.empty(2 * chunks, dtype=torch.float64, device="cuda")
            L.check(lib.sg_chunk_moments(L.P(d.data_ptr()), half, L.P(p2.data_ptr()), st))
            L.check(lib.sg_chunk_moments(L.P(d[half:].data_ptr()), n - half, L.P(p2[2 * (chunks // 2):].data_ptr()), st))
            assert torch.equal(part, p2)


def test_autoencoder_straining_golden(sb, golden):
    """AE reconstruction straining vs the reference's own detect_outliers_autoencoder outputs"""
    x = torch.from_numpy(O.synth_images(0, 160))
    torch.manual_seed(O.SEED)
    ae = O.AutoEncoder()
    err = sb.ae_errors(ae, x, "cuda").cpu().numpy()
    want = golden["g6_errors"]
    assert (np.abs(err - want) / np.maximum(want, 1e-6)).max() <= 1e-3, np.abs(err - want).max()   # fp32: 1e-3 rel (measured ~1e-6)
    ds = torch.utils.data.TensorDataset(x, torch.zeros(160, dtype=torch.long))
    for thr_k, key in ((2.0, "g6_inlier"), (0.5, "g6_inlier_t05")):
        got = sb.detect_outliers_autoencoder(ae, ds, "cuda", thr_k) if thr_k != 2.0 else sb.detect_outliers_autoencoder(ae, ds, "cuda")
        assert isinstance(got, torch.Tensor) and got.dtype == torch.bool and got.device.type == "cpu"
        te = torch.from_numpy(want)
        t = (te.mean() + thr_k * te.std()).item()
        near = np.abs(want - t) <= 1e-3 * abs(t)
        assert not ((got.numpy() != golden[key]) & ~near).any()


def test_autoencoder_bf16_conv_mode(sb, golden):
    """bf16 conv mode (BASELINE config 4, stated separately from the fp32 bar): 7x7 layers on tcgen05 with bf16
    operands and bf16 activations.  Reconstruction errors within 2e-2 relative of the reference's fp32 values;
    inlier masks may differ only for samples within that tolerance of the threshold.  Batch sizes around the
    tile sizes (1 image = 1 tile of L3 = 2 tiles of L4) incl. a ragged last chunk."""
    torch.manual_seed(O.SEED)
    ae = O.AutoEncoder()
    x = torch.from_numpy(O.synth_images(0, 160))
    want = golden["g6_errors"]
    err = sb.ae_errors(ae, x, "cuda", conv_mode="bf16").cpu().numpy()
    rel = np.abs(err - want) / np.maximum(want, 1e-6)
    assert rel.max() <= 2e-2, rel.max()
    for n, chunk in ((1, 2048), (3, 2), (149, 64), (160, 2048)):
        e2 = sb.ae_errors(ae, x[:n], "cuda", chunk=chunk, conv_mode="bf16").cpu().numpy()
        assert np.array_equal(e2, err[:n]), (n, chunk)     # independent of batch size / chunking (fixed-order sums)
    ds = torch.utils.data.TensorDataset(x, torch.zeros(160, dtype=torch.long))
    got = sb.detect_outliers_autoencoder(ae, ds, "cuda", 2.0, conv_mode="bf16")
    te = torch.from_numpy(want)
    t = (te.mean() + 2.0 * te.std()).item()
    near = np.abs(want - t) <= 2e-2 * abs(t)
    assert not ((got.numpy() != golden["g6_inlier"]) & ~near).any()
    # trained-looking weights: larger, non-default parameters exercise every tap of the 7x7 layers
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in ae.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.5 / np.sqrt(max(p[0].numel(), 1))))
    ref = O.ae_errors(ae, x[:48]).numpy()
    e3 = sb.ae_errors(ae, x[:48], "cuda", conv_mode="bf16").cpu().numpy()
    assert (np.abs(e3 - ref) / np.maximum(ref, 1e-6)).max() <= 2e-2
    e4 = sb.ae_errors(ae, x[:48], "cuda").cpu().numpy()                      # the default: 'auto' (fp16 pass + recovery)
    assert (np.abs(e4 - ref) / np.maximum(ref, 1e-6)).max() <= 1e-3, (np.abs(e4 - ref) / np.maximum(ref, 1e-6)).max()
    e6 = sb.ae_errors(ae, x[:48], "cuda", conv_mode="fp16").cpu().numpy()    # fp16 mode: one tensor pass, the fp32 bar
    assert (np.abs(e6 - ref) / np.maximum(ref, 1e-6)).max() <= 1e-3, (np.abs(e6 - ref) / np.maximum(ref, 1e-6)).max()
    e5 = sb.ae_errors(ae, x[:48], "cuda", conv_mode="fp32").cpu().numpy()    # fp32-parity arithmetic (bf16 hi/lo split)
    assert (np.abs(e5 - ref) / np.maximum(ref, 1e-6)).max() <= 1e-5
    for n, chunk in ((1, 2048), (5, 2), (48, 17)):
        e6 = sb.ae_errors(ae, x[:n], "cuda", chunk=chunk).cpu().numpy()
        assert np.array_equal(e6, e4[:n]), (n, chunk)


def test_dbscan_nd_clean_ratio_vs_sklearn(sb):
    """StandardScaler -> DBSCAN -> mean(labels != -1) on [N, 512] features via thresholded
    pairwise-distance GEMMs on tcgen05; counts of core / non-noise points equal scikit-learn's."""
    from sklearn.cluster import DBSCAN
    from sklearn.preprocessing import StandardScaler
    rng = np.random.default_rng(17)
    for n, d in ((700, 512), (3000, 512), (1025, 64), (2500, 128)):
        # a few tight clusters + scattered points: core, border and noise points all occur
        centers = rng.standard_normal((6, d)).astype(np.float32) * 3
        lab = rng.integers(0, 6, n)
        f = centers[lab] + rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.2, 1.2, (n, 1)).astype(np.float32)
        f[: n // 10] = rng.standard_normal((n // 10, d)).astype(np.float32) * 6          # outliers
        z = StandardScaler().fit_transform(f)
        dist = np.sqrt(np.maximum((z * z).sum(1)[:, None] + (z * z).sum(1)[None, :] - 2 * z.astype(np.float64) @ z.T.astype(np.float64), 0))
        for frac in (0.02, 0.1, 0.4):
            eps = float(np.quantile(dist[np.triu_indices(n, 1)], frac))
            if np.abs(dist - eps).min() < 1e-4 * eps:
                eps *= 1.0003                                             # keep every pair off the decision boundary
            m = DBSCAN(eps=eps, min_samples=3).fit(z)
            want_core, want_clean = len(m.core_sample_indices_), int((m.labels_ != -1).sum())
            ratio, core, clean = sb.dbscan_clean_ratio(torch.from_numpy(f), eps, 3, return_counts=True)
            assert (core, clean) == (want_core, want_clean), (n, d, frac, core, clean, want_core, want_clean)
            assert ratio == want_clean / n
    ds = torch.utils.data.TensorDataset(torch.from_numpy(f), torch.zeros(n))
    r_dev = sb.estimate_ratio_dbscan(ds, eps=eps, min_samples=3, feature_extractor=torch.nn.Identity())
    assert r_dev == O.estimate_ratio_dbscan_features(f, eps, 3)


def test_dbscan_nd_edges(sb):
    """degenerate neighbourhoods and tiny inputs of the tensor-core DBSCAN"""
    rng = np.random.default_rng(3)
    f = rng.standard_normal((37, 64)).astype(np.float32)
    assert sb.dbscan_clean_ratio(torch.from_numpy(f), 1e-3, 3, return_counts=True) == (0.0, 0, 0)          # everyone alone
    assert sb.dbscan_clean_ratio(torch.from_numpy(f), 1e3, 3, return_counts=True) == (1.0, 37, 37)        # one big cluster
    assert sb.dbscan_clean_ratio(torch.from_numpy(f), 1e-3, 1, return_counts=True) == (1.0, 37, 37)       # min_samples 1: self
    g = np.repeat(rng.standard_normal((5, 128)).astype(np.float32), 3, axis=0)                              # 5 exact triplets
    g = np.concatenate([g, rng.standard_normal((4, 128)).astype(np.float32) * 5])
    ratio, core, clean = sb.dbscan_clean_ratio(torch.from_numpy(g), 1e-2, 3, return_counts=True)
    assert (core, clean) == (15, 15) and ratio == 15 / 19
    with pytest.raises(RuntimeError):
        sb.dbscan_clean_ratio(torch.from_numpy(rng.standard_normal((10, 100)).astype(np.float32)), 1.0, 3)   # d % 64 != 0


def test_autoencoder_tensor_core_forms_vs_cuda_core(sb):
    """every conv layer of the single-segment modes runs on tcgen05 (stride-2 TMA boxes, parity-class accumulators, the two
    7x7 layers in row-tap form with the column taps summed by shuffles).  With non-trivial weights every tap matters: the
    tensor-core modes must agree with the oracle, also on ragged image counts
    (last CTA wave partly empty) and independently of the chunking."""
    torch.manual_seed(11)
    ae = O.AutoEncoder()
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for p in ae.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.7 / np.sqrt(max(p[0].numel(), 1))))
    x = torch.from_numpy(O.synth_images(300, 333))          # odd count: more images than SMs, ragged last wave
    ref = O.ae_errors(ae, x).numpy()
    for mode, tol in (("bf16", 2e-2), ("fp16", 1e-3), ("fp32", 1e-3)):
        e = sb.ae_errors(ae, x, "cuda", conv_mode=mode).cpu().numpy()
        rel = (np.abs(e - ref) / np.maximum(ref, 1e-6)).max()
        assert rel <= tol, (mode, rel)
        for n, chunk in ((1, 2048), (37, 5), (333, 100)):
            e2 = sb.ae_errors(ae, x[:n], "cuda", chunk=chunk, conv_mode=mode).cpu().numpy()
            assert np.array_equal(e2, e[:n]), (mode, n, chunk)


def test_workspaces_are_1024_aligned_whatever_the_allocator_returns(sb):
    """torch's caching allocator aligns small blocks to 512 bytes only; the C ABI wants 1024 on its TMA workspaces.  Shift
    the small pool by 512-byte blocks and request fresh scratch buffers: every one must come back 1024-aligned and the
    small 512-d DBSCAN call (whose workspace is a small block) must work at any pool offset."""
    from strainer_gan_b200 import api
    rng = np.random.default_rng(3)
    f = torch.from_numpy(rng.standard_normal((300, 64)).astype(np.float32))
    keep = []
    for i in range(6):
        keep.append(torch.empty(512, dtype=torch.uint8, device="cuda"))      # moves the next small block by 512 bytes
        api._Scratch._cache.clear()
        t = api._aligned_empty(1000 + 512 * i, torch.device("cuda", 0))
        assert t.data_ptr() % 1024 == 0 and t.numel() >= 1000 + 512 * i
        r = sb.dbscan_clean_ratio(f, 9.0, 3)
        assert 0.0 <= r <= 1.0
        for (dev, key), buf in api._Scratch._cache.items():
            assert buf.data_ptr() % 1024 == 0, key
