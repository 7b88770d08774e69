"""CPU: the oracle restatements vs (1) fixtures produced by executing the reference's own code
(tests/golden/make_golden.py) and (2) the library calls the reference makes."""
import numpy as np
import pytest
import torch

from oracle import strainer_oracle as O

N1 = 160


@pytest.fixture(scope="module")
def imgs():
    return torch.from_numpy(O.synth_images(0, N1))


@pytest.fixture(scope="module")
def netD():
    return O.make_discriminator(O.SEED)


def test_synth_is_counter_based():
    a = O.synth_images(0, 16)
    b = O.synth_images(8, 8)
    assert np.array_equal(a[8:], b)
    assert a.dtype == np.float32 and a.min() >= -1 and a.max() < 1
    noisy = O.synth_is_noisy(O.SEED, np.arange(20000))
    assert 0.17 < noisy.mean() < 0.23


def test_refine_dataset_by_loss_golden(golden, imgs, netD):
    losses = O.score_losses(netD, imgs, 64)
    assert np.array_equal(losses, golden["g1_losses"])
    for tag in "abc":
        idx, thr, _ = O.refine_dataset_by_loss(imgs, netD, float(golden[f"g1{tag}_ratio"]))
        assert thr.dtype == np.float32
        assert np.array_equal(np.asarray(thr), golden[f"g1{tag}_threshold"])
        assert np.array_equal(idx, golden[f"g1{tag}_indices"])


def test_refine_fallback_degenerate(golden, netD):
    idx, thr, _ = O.refine_dataset_by_loss(torch.zeros(8, 3, 64, 64), netD, 0.2)
    assert np.array_equal(np.asarray(idx), golden["g1z_indices"])
    assert np.array_equal(np.asarray(thr), golden["g1z_threshold"])


def test_evaluate_dataset_golden(golden, imgs, netD):
    assert np.array_equal(O.evaluate_dataset(netD, imgs), golden["g2_eval_losses"])


def test_gmm_divide_golden(golden):
    lo = golden["g2_losses"]
    assert np.array_equal(lo, O.synth_losses(5000))
    np.random.seed(1234)
    thr = O.get_gmm_threshold(lo.copy())
    clean, noisy = O.divide_by_threshold(lo, thr)
    assert np.array_equal(clean, golden["g2_clean_idx"])
    assert np.array_equal(noisy, golden["g2_noisy_idx"])
    assert thr == golden["g3_gmm"]


def test_ensemble_golden(golden):
    lo = golden["g2_losses"]
    assert O.get_percentile_threshold(lo) == golden["g3_p75"]
    assert str(np.asarray(O.get_percentile_threshold(lo)).dtype) == str(golden["g3_p75_dtype"])
    assert O.get_iqr_threshold(lo) == golden["g3_iqr"]
    np.random.seed(1234)
    thr = O.get_ensemble_threshold(lo.copy())
    assert thr == golden["g3_ensemble"]
    assert np.array_equal(O.divide_by_threshold(lo, thr)[0], golden["g3_clean_idx"])


def test_elbow_and_detect_outliers_golden(golden):
    feats = torch.from_numpy(O.synth_features(4096))
    mz = O.zscore_max_torch(feats).numpy()
    assert np.array_equal(mz, golden["g4_maxz"])
    thr, centers, hist = O.find_elbow_threshold(mz)
    assert thr == golden["g4_threshold"]
    assert np.array_equal(centers, golden["g4_centers"]) and np.array_equal(hist, golden["g4_hist"])
    counts, edges = O.np_histogram_f32(mz, 100)
    thr2, c2, h2 = O.elbow_from_hist(counts, edges)
    assert thr2 == thr and np.array_equal(c2, centers) and np.array_equal(h2, hist)
    assert np.array_equal(O.detect_outliers_elbow(feats)[0], golden["g5_elbow_inlier"])
    assert np.array_equal(O.detect_outliers_elbow(feats, 5)[0], golden["g5_user5_inlier"])
    assert np.array_equal(O.detect_outliers_fixed(feats).numpy(), golden["g5_fixed5_inlier"])
    assert np.array_equal(O.detect_outliers_fixed(feats, 4.5).numpy(), golden["g5_fixed45_inlier"])
    for tag in "ab":
        got = O.detect_outliers_ratio(feats, float(golden[f"g5_ratio{tag}"])).numpy()
        assert np.array_equal(got, golden[f"g5_ratio{tag}_inlier"])
    assert np.array_equal(O.zscore_max_numpy(feats.numpy()), golden["g7_maxz_np"])


def test_autoencoder_golden(golden, imgs):
    torch.manual_seed(O.SEED)
    ae = O.AutoEncoder()
    inl, thr, err = O.detect_outliers_autoencoder(ae, imgs)
    assert np.array_equal(err.numpy(), golden["g6_errors"])
    assert np.array_equal(inl.numpy(), golden["g6_inlier"])
    assert np.array_equal(O.detect_outliers_autoencoder(ae, imgs, 0.5)[0].numpy(), golden["g6_inlier_t05"])


@pytest.mark.parametrize("B", [64, 128])
def test_inbatch_strain_golden(golden, imgs, B):
    d = O.make_discriminator(O.SEED)  # train mode
    fr, ff, mask, thr, scores = O.strain_batch(d, imgs[:B])
    assert np.array_equal(scores.numpy(), golden[f"g8_{B}_scores"])
    assert np.array_equal(thr.numpy(), golden[f"g8_{B}_threshold"])
    assert np.array_equal(mask.numpy(), golden[f"g8_{B}_mask"])
    assert fr.shape[0] == golden[f"g8_{B}_nreal"] and ff.shape[0] == golden[f"g8_{B}_nfake"]
    assert ff.shape[0] == {64: 7, 128: 13}[B]
    assert np.array_equal(d.main[3].running_mean.numpy(), golden[f"g8_{B}_bn1_mean"])
    assert np.array_equal(d.main[9].running_var.numpy(), golden[f"g8_{B}_bn3_var"])
    assert int(d.main[3].num_batches_tracked) == golden[f"g8_{B}_nbt"] == 1


@pytest.mark.parametrize("B", [256, 512])
def test_inbatch_strain_golden_configs_3_4(golden2, B):
    """BASELINE configs 3 / 4 batch sizes: the oracle's in-batch block equals the reference's own code (train-mode BN)"""
    d = O.make_discriminator(O.SEED)  # train mode
    x = torch.from_numpy(O.synth_images(0, B))
    fr, ff, mask, thr, scores = O.strain_batch(d, x)
    assert np.array_equal(scores.numpy(), golden2[f"g8_{B}_scores"])
    assert np.array_equal(thr.numpy(), golden2[f"g8_{B}_threshold"])
    assert np.array_equal(mask.numpy(), golden2[f"g8_{B}_mask"])
    assert ff.shape[0] == golden2[f"g8_{B}_nfake"] == {256: 26, 512: 52}[B]  # s = 26 / 52
    for li, name in ((3, "bn1"), (6, "bn2"), (9, "bn3")):
        assert np.array_equal(d.main[li].running_mean.numpy(), golden2[f"g8_{B}_{name}_mean"])
        assert np.array_equal(d.main[li].running_var.numpy(), golden2[f"g8_{B}_{name}_var"])


def test_concat_block_golden(golden2):
    """strained -> fake concat + generator-loss gradient ("# 상위 10% 제거해서 fake image에 concate.py:265-273, 282-284")"""
    B = 64
    d = O.make_discriminator(O.SEED).eval()
    x = torch.from_numpy(O.synth_images(1000, B))
    with torch.no_grad():
        fr, ff, mask, thr, scores = O.strain_batch(d, x)
    assert np.array_equal(mask.numpy(), golden2["g9_mask"]) and ff.shape[0] == golden2["g9_nfake"]
    g = torch.Generator().manual_seed(1234)
    gz = torch.tanh(torch.randn(B - ff.shape[0], 3, 64, 64, generator=g)).requires_grad_(True)
    fake = O.concat_fake(gz, ff)
    assert fake.shape[0] == golden2["g9_label_len"] == B
    assert np.array_equal(fake.detach().double().sum(dim=(1, 2, 3)).numpy(), golden2["g9_fake_rowsum"])
    out = d(fake).view(-1)
    err = torch.nn.BCELoss()(out, torch.full((B,), 1.0))
    err.backward()
    assert np.array_equal(out.detach().numpy(), golden2["g9_output"])
    assert np.array_equal(gz.grad[:, :, ::16, ::16].numpy(), golden2["g9_grad_sample"])


# ---- bit-level restatements vs the library calls themselves --------------------------------
def test_np_percentile_restatement():
    rng = np.random.default_rng(7)
    for _ in range(400):
        n = int(rng.integers(1, 3000))
        q = [float(rng.uniform(0, 100)), (1 - 0.8) * 100, 90.0, 75, 25, 0.0, 100.0][int(rng.integers(0, 7))]
        v = rng.standard_normal(n).astype(np.float32)
        if rng.random() < 0.3:
            v = np.round(v, 1)  # ties
        want = np.percentile(v, q)
        got = O.np_percentile_f32(v, q)
        assert want.dtype == got.dtype == np.float32
        assert want == got or (np.isnan(want) and np.isnan(got)), (n, q, want, got)
    v = rng.standard_normal(1_000_003).astype(np.float32)
    for q in (90.0, (1 - 0.8) * 100, 33.3333):
        assert np.percentile(v, q) == O.np_percentile_f32(v, q)
    v[5] = np.nan
    assert np.isnan(O.np_percentile_f32(v, 90.0)) and np.isnan(np.percentile(v, 90.0))


def test_torch_quantile_restatement():
    rng = np.random.default_rng(8)
    bad = 0
    for _ in range(600):
        n = int(rng.integers(2, 1500))
        q = [0.1, 0.9, 0.8731, float(rng.uniform(0, 1))][int(rng.integers(0, 4))]
        v = rng.standard_normal(n).astype(np.float32)
        want = torch.quantile(torch.from_numpy(v), q).numpy()
        got = O.torch_quantile_f32(v, q)
        bad += int(want != got)
    assert bad == 0


def test_np_histogram_restatement():
    rng = np.random.default_rng(9)
    for n in (1, 2, 17, 1000, 70001):
        for scale in (1.0, 1e-3, 37.5):
            v = (rng.standard_normal(n) * scale).astype(np.float32)
            want, we = np.histogram(v, bins=100)
            got, ge = O.np_histogram_f32(v, 100)
            assert np.array_equal(want, got) and np.array_equal(we, ge) and we.dtype == ge.dtype


def test_dbscan1d_restatement_vs_sklearn():
    rng = np.random.default_rng(10)
    for n, eps, ms in ((50, 0.05, 3), (2000, 0.01, 3), (2000, 0.002, 5), (300, 0.5, 3), (5, 0.1, 3)):
        v = O.synth_losses(n, seed=int(rng.integers(1 << 30)))
        want = O.dbscan1d_noise_sklearn(v, eps, ms)
        got = O.dbscan1d_noise(v, eps, ms)
        assert np.array_equal(want, got), (n, eps, ms, want.sum(), got.sum())
    v = np.round(O.synth_losses(3000, seed=3), 2)  # heavy ties, distances exactly == eps
    assert np.array_equal(O.dbscan1d_noise_sklearn(v, 0.01, 3), O.dbscan1d_noise(v, 0.01, 3))


def test_train_step_restatement_vs_reference_golden(golden3):
    """O.train_step_through_d == the reference Discriminator driven through the literal D step / G step
    ("#strainer gan.py:586-615", fixture tests/golden/make_golden_v3.py), CPU fp32."""
    g = golden3
    B = g["g10_out_real"].shape[0]
    real = torch.from_numpy(O.synth_images(2000, B))
    fake = torch.tanh(torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(4321)))
    netD = O.make_discriminator(O.SEED).train()
    r = O.train_step_through_d(netD, real, fake)
    for k in ("out_real", "out_fake", "out_g"):
        assert np.allclose(r[k].numpy(), g["g10_" + k], rtol=1e-5, atol=1e-7), k
    assert abs(float(r["errD"]) - float(g["g10_errD"])) <= 1e-5 and abs(float(r["errG"]) - float(g["g10_errG"])) <= 1e-5
    for n, gr in zip(O.D64_PARAM_NAMES, r["d_grads"]):
        want = g[f"g10_d_{n}_sample"]
        got = (gr.reshape(-1)[::61] if gr.numel() > 4096 else gr.reshape(-1)).numpy()
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), n
        assert abs(float(gr.double().norm()) - float(g[f"g10_d_{n}_norm"])) <= 1e-4 * float(g[f"g10_d_{n}_norm"]), n
    want = g["g10_dfake_sample"]
    assert np.abs(r["dfake"][:, :, ::8, ::8].numpy() - want).max() <= 1e-4 * np.abs(want).max()
    bns = [m for m in netD.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for i, bn in enumerate(bns):
        assert np.allclose(bn.running_mean.numpy(), g[f"g10_bn{i + 2}_mean"], rtol=1e-5, atol=1e-7)
        assert np.allclose(bn.running_var.numpy(), g[f"g10_bn{i + 2}_var"], rtol=1e-5, atol=1e-7)
    assert int(bns[0].num_batches_tracked) == int(g["g10_nbt"]) == 3
