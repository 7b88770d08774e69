"""GPU parity of the round-2 behaviours: the 'auto' conv mode (fp16 pass + fp32-parity recovery of overflowed chunks) as the
drop-in default, the fused strained->fake concat with its autograd, the in-batch block at the batch sizes of BASELINE
configs 3 / 4, dataset streaming, back-to-back host scoring, the per-device library state, the reference's call sites of
the four ``detect_outliers`` definitions."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import strainer_oracle as O

pytestmark = pytest.mark.gpu
ATOL = 1e-6


@pytest.fixture(scope="module")
def sb():
    import strainer_b200
    assert torch.cuda.is_available()
    return strainer_b200


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), ATOL)


def _overflowing_discriminator(x_probe):
    """A D whose act1 is ~3e4 on ordinary images (fp16 holds it) and beyond 65504 on 4x brighter ones, with BN2's running
    statistics rescaled so that every later activation -- and the losses -- are those of the unscaled network: only the
    intermediate overflows, exactly the case conv_mode='auto' has to recover."""
    d = O.make_discriminator(5).eval()
    with torch.no_grad():
        a1 = torch.nn.functional.leaky_relu(d.main[0](x_probe), 0.2)
        s = float(3.0e4 / a1.abs().max())
        d.main[0].weight.mul_(s)                 # act1 *= s (LeakyReLU is positively homogeneous)
        d.main[3].running_mean.mul_(s)           # conv2 output *= s: BN2 divides it out again
        d.main[3].running_var.mul_(s * s)        # (eps stays: the oracle scores this same modified network)
    return d


def test_auto_mode_recovers_overflowed_chunks(sb):
    """default call, no conv_mode: chunk 0 scores in fp16, chunk 1 overflows fp16 and is re-scored in fp32-parity
    arithmetic on the device; every loss within 1e-3 of the oracle, the mask exact away from the threshold, no raise"""
    n, cb = 128, 64
    x = torch.from_numpy(O.synth_images(40, n)).clone()
    d = _overflowing_discriminator(x[:cb])
    x[cb:] *= 4.0
    widx, wthr, wloss = O.refine_dataset_by_loss(x, d, 0.1)
    wloss = wloss.reshape(-1)
    # the fp16 mode alone must see the overflow (otherwise this test exercises nothing)
    with pytest.raises(RuntimeError, match="fp16"):
        sb.D64Scorer(d, "cuda", "fp16", max_batch=cb).score(x.cuda(), ("loss",))
    sc = sb.D64Scorer(d, "cuda", "auto", max_batch=cb)
    loss = sc.score(x.cuda(), ("loss",))["loss"].cpu().numpy()
    assert sc.fallback_chunks == 1
    assert np.isfinite(loss).all() and _rel(loss, wloss).max() <= 1e-3, _rel(loss, wloss).max()
    # host-resident input: the overflowed chunk is uploaded again for the recovery
    loss_h = sc.score(x, ("loss",))["loss"].cpu().numpy()
    assert np.array_equal(loss_h, loss) and sc.fallback_chunks == 2
    # and through the reference-facing call with its defaults
    ds = torch.utils.data.TensorDataset(x, torch.zeros(n, dtype=torch.long))
    sub, thr = sb.refine_dataset_by_loss(ds, d, "cuda", 0.1)
    assert abs(thr - wthr) <= 1e-3 * abs(wthr)
    got = np.zeros(n, bool)
    got[np.asarray(sub.indices)] = True
    want = np.zeros(n, bool)
    want[widx] = True
    near = np.abs(wloss - wthr) <= 1e-3 * abs(wthr)
    assert not ((got != want) & ~near).any()


def test_auto_mode_strain_batch_overflow_train_bn(sb):
    """in-batch block, train-mode BN, a batch that overflows fp16: the running statistics are updated exactly ONCE (by
    the fp32-parity pass) and the scores match the reference's train-mode forward"""
    B = 64
    x = torch.from_numpy(O.synth_images(90, B)).clone()
    d = _overflowing_discriminator(x)
    x *= 4.0
    d.train()
    import copy
    d_ref = copy.deepcopy(d)
    with torch.no_grad():
        want_scores = d_ref(x).view(-1)
    wt = torch.quantile(want_scores, 0.1)
    d = d.cuda()
    fr, ff, mask, t = sb.strain_batch(d, x.cuda())
    assert sb.get_scorer(d, "cuda", "auto", max_batch=512).fallback_chunks == 1
    assert abs(t.item() - wt.item()) <= 1e-3 * abs(wt.item())
    near = (want_scores - wt).abs() <= 1e-3 * wt.abs()
    assert not ((mask.cpu() != (want_scores >= wt)) & ~near).any()
    for li in (3, 6, 9):
        assert np.allclose(d.main[li].running_mean.cpu().numpy(), d_ref.main[li].running_mean.numpy(), rtol=1e-3, atol=1e-4)
        assert np.allclose(d.main[li].running_var.cpu().numpy(), d_ref.main[li].running_var.numpy(), rtol=1e-3, atol=1e-5)
        assert int(d.main[li].num_batches_tracked) == 1


def test_default_mode_meets_fp32_bar_at_scale(sb):
    """an UNMODIFIED call (no conv_mode) against the oracle on 4096 samples: losses within 1e-3, mask exact away from
    the threshold"""
    n = 4096
    netD = O.make_discriminator(O.SEED)
    x = torch.from_numpy(O.synth_images(20000, n))
    widx, wthr, wloss = O.refine_dataset_by_loss(x, netD, 0.1)
    wloss = wloss.reshape(-1)
    ds = torch.utils.data.TensorDataset(x, torch.zeros(n, dtype=torch.long))
    sub, thr = sb.refine_dataset_by_loss(ds, netD, "cuda", 0.1)
    loss = sb.evaluate_dataset(netD, ds, "cuda")
    assert _rel(loss, wloss).max() <= 1e-3
    assert abs(thr - wthr) <= 1e-3 * abs(wthr)
    got = np.zeros(n, bool)
    got[np.asarray(sub.indices)] = True
    want = np.zeros(n, bool)
    want[widx] = True
    near = np.abs(wloss - wthr) <= 1e-3 * abs(wthr)
    assert not ((got != want) & ~near).any()
    assert sb.get_scorer(netD, "cuda").fallback_chunks == 0


@pytest.mark.parametrize("B", [256, 512])
@pytest.mark.parametrize("mode", ["auto", "fp32"])
def test_strain_batch_train_mode_bn_configs_3_4(sb, golden2, B, mode):
    """BASELINE configs 3 / 4 (B = 256 / 512): train-mode BN scoring against fixtures from the reference's own block"""
    x = torch.from_numpy(O.synth_images(0, B))
    d = O.make_discriminator(O.SEED).cuda()          # train mode
    fr, ff, mask, thr = sb.strain_batch(d, x.cuda(), conv_mode=mode)
    scores_w = torch.from_numpy(golden2[f"g8_{B}_scores"])
    thr_w = float(golden2[f"g8_{B}_threshold"])
    near = (scores_w - thr_w).abs() <= 1e-3 * abs(thr_w)
    assert not ((mask.cpu().numpy() != golden2[f"g8_{B}_mask"]) & ~near.numpy()).any()
    assert abs(thr.item() - thr_w) <= 1e-3 * abs(thr_w)
    assert abs(ff.shape[0] - int(golden2[f"g8_{B}_nfake"])) <= 1 and fr.shape[0] + ff.shape[0] == B
    m = mask.cpu()
    assert torch.equal(fr.cpu(), x[m]) and torch.equal(ff.cpu(), x[~m])
    for li, name in ((3, "bn1"), (6, "bn2"), (9, "bn3")):
        for t, key in ((d.main[li].running_mean, "mean"), (d.main[li].running_var, "var")):
            w = golden2[f"g8_{B}_{name}_{key}"]
            atol = 1e-5 if mode == "fp32" else 1e-4      # fp16 operands: statistics of fp16-rounded conv outputs
            assert np.allclose(t.cpu().numpy(), w, rtol=1e-3, atol=atol), (name, key, np.abs(t.cpu().numpy() - w).max())
        assert int(d.main[li].num_batches_tracked) == 1


def test_concat_fake_literal_block(sb, golden2):
    """K12 against the literal block "# 상위 10% 제거해서 fake image에 concate.py:265-273, 282-284": values, label lengths
    and the gradient reaching the generator rows; fixtures from the reference's D, live torch ops beside them"""
    B = 64
    d = O.make_discriminator(O.SEED).eval().cuda()
    real = torch.from_numpy(O.synth_images(1000, B)).cuda()
    filtered_real, filtered_fake, mask, thr = sb.strain_scores(real, torch.from_numpy(golden2["g9_scores"]).cuda())
    assert np.array_equal(mask.cpu().numpy(), golden2["g9_mask"])
    b_size_fake = filtered_fake.size(0)
    assert b_size_fake == int(golden2["g9_nfake"])
    g = torch.Generator().manual_seed(1234)
    gz_host = torch.tanh(torch.randn(B - b_size_fake, 3, 64, 64, generator=g))
    gz = gz_host.cuda().requires_grad_(True)           # stands in for netG(noise)
    gz_ref = gz_host.cuda().requires_grad_(True)
    strained_copy = filtered_fake.clone()
    # --- the library: strained rows already sit behind the generator rows; one launch copies the generator rows
    fake = sb.concat_fake(gz, filtered_fake)
    assert fake.data_ptr() + (B - b_size_fake) * 3 * 64 * 64 * 4 == filtered_fake.data_ptr()    # no copy of the strained rows
    # --- the literal block
    fake_ref = torch.cat([gz_ref, strained_copy], dim=0)
    label_fake = torch.full((fake.size(0),), 0.0, dtype=torch.float, device="cuda")
    assert label_fake.numel() == int(golden2["g9_label_len"]) == fake_ref.size(0)
    assert torch.equal(fake.detach(), fake_ref.detach())
    assert np.array_equal(fake.detach().double().sum(dim=(1, 2, 3)).cpu().numpy(), golden2["g9_fake_rowsum"])
    label_g = torch.full((fake.size(0),), 1.0, dtype=torch.float, device="cuda")
    crit = nn.BCELoss()
    out_d = d(fake.detach()).view(-1)                       # D step on fake.detach() (":271")
    assert out_d.shape[0] == B and not fake.detach().requires_grad and fake.requires_grad
    errG = crit(d(fake).view(-1), label_g)
    errG.backward()
    crit(d(fake_ref).view(-1), label_g).backward()
    # the two backward passes run the same cuDNN dgrad kernels on identical data, but those kernels are not
    # bit-reproducible run to run (split-K atomics): compare closely here, exactly on an injected gradient below
    assert float((gz.grad - gz_ref.grad).norm()) <= 1e-3 * float(gz_ref.grad.norm())
    gz3 = gz_host.cuda().requires_grad_(True)
    up = torch.randn(B, 3, 64, 64, device="cuda")
    sb.concat_fake(gz3, filtered_fake).backward(up)
    assert torch.equal(gz3.grad, up[:B - b_size_fake])      # bit-identical to autograd through torch.cat: grad[:B-s]
    # against the reference's CPU fp32 gradient: D's backward here is torch autograd on the GPU (TF32 convolutions by
    # default), so the bar is on the vector, not per element
    got, want = gz.grad[:, :, ::16, ::16].cpu().numpy().astype(np.float64), golden2["g9_grad_sample"].astype(np.float64)
    assert np.linalg.norm(got - want) <= 2e-2 * np.linalg.norm(want)
    assert abs(errG.item() - float(golden2["g9_errG"])) <= 1e-3 * abs(float(golden2["g9_errG"]))
    # a strained tensor that is NOT the pre-placed tail goes through the same kernel into a fresh buffer
    other = torch.randn(5, 3, 64, 64, device="cuda")
    gz2 = gz_host.cuda().requires_grad_(True)
    cat2 = sb.concat_fake(gz2, other)
    assert torch.equal(cat2.detach(), torch.cat([gz2.detach(), other]))
    (cat2 * 2).sum().backward()
    assert torch.equal(gz2.grad, torch.full_like(gz2, 2.0))
    with pytest.raises(ValueError):
        sb.concat_fake(gz2, torch.zeros(3, 1, 64, 64, device="cuda"))


class _GetItemOnly(torch.utils.data.Dataset):
    """a map-style dataset WITHOUT a resident tensor (what an ImageFolder with transforms is)"""

    def __init__(self, x):
        self.x = x

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i].clone(), 0


def test_streamed_dataset_equals_resident(sb, monkeypatch):
    n = 300
    netD = O.make_discriminator(O.SEED)
    x = torch.from_numpy(O.synth_images(7000, n))
    want = sb.evaluate_dataset(netD, torch.utils.data.TensorDataset(x, torch.zeros(n)), "cuda")
    import strainer_gan_b200.api as api
    monkeypatch.setattr(api, "STREAM_BATCH", 128)          # 3 loader batches, the last one ragged
    got = sb.evaluate_dataset(netD, _GetItemOnly(x), "cuda")
    assert np.array_equal(got, want)
    sub, thr = sb.refine_dataset_by_loss(_GetItemOnly(x), netD, "cuda", 0.2)
    sub_w, thr_w = sb.refine_dataset_by_loss(torch.utils.data.TensorDataset(x, torch.zeros(n)), netD, "cuda", 0.2)
    assert thr == thr_w and np.array_equal(np.asarray(sub.indices), np.asarray(sub_w.indices))
    img, _ = sub[3]
    assert torch.equal(img, x[sub.indices[3]])


@pytest.mark.parametrize("mode", ["auto", "fp32"])
def test_back_to_back_host_scoring(sb, mode):
    """two host-streamed scorings enqueued back to back reuse the scorer's staging buffers: the second call's copies
    must wait for the first call's kernels (regression: losses of the first call were computed from the second's pixels)"""
    n = 600
    netD = O.make_discriminator(O.SEED)
    xa = torch.from_numpy(O.synth_images(100, n))
    xb = torch.from_numpy(O.synth_images(9000, n)).pin_memory()
    sc = sb.D64Scorer(netD.eval(), "cuda", mode, max_batch=128)
    wa = sc.score(xa.cuda(), ("loss",))["loss"].clone()
    wb = sc.score(xb.cuda(), ("loss",))["loss"].clone()
    for _ in range(3):
        la = sb.evaluate_dataset(netD, torch.utils.data.TensorDataset(xa, torch.zeros(n)), "cuda", conv_mode=mode,
                                 return_device=True)
        lb = sb.evaluate_dataset(netD, torch.utils.data.TensorDataset(xb, torch.zeros(n)), "cuda", conv_mode=mode,
                                 return_device=True)
        a1 = sc.score(xa, ("loss",))["loss"]
        b1 = sc.score(xb, ("loss",))["loss"]
        assert torch.equal(a1, wa) and torch.equal(b1, wb)
        assert torch.equal(la, wa) and torch.equal(lb, wb)


def test_scorer_cache_follows_module_lifetime(sb):
    import gc
    import strainer_gan_b200.api as api
    d = O.make_discriminator(3).eval()
    x = torch.from_numpy(O.synth_images(0, 8)).cuda()
    sb.get_scorer(d, "cuda", "auto", max_batch=64).score(x)
    key = (id(d), torch.cuda.current_device(), "auto", 64)
    assert key in api._SCORERS
    del d
    gc.collect()
    assert key not in api._SCORERS


def test_reference_call_sites_of_detect_outliers(sb, golden):
    """the reference defines four functions called ``detect_outliers``; each call site, written as upstream writes it,
    reaches its own rule ("#strainer gan.py:331-360", "#z_score.py:276-294", "#z_score + 엘보우 threshold.py:306-330",
    "# z_score + DBSCAN.py:305-326")"""
    feats = torch.from_numpy(O.synth_features(4096))
    fds = torch.utils.data.TensorDataset(feats, torch.zeros(feats.shape[0]))
    ident = nn.Identity()
    mz = golden["g4_maxz"]

    def same(got, want, thr, tol=1e-4):      # z-scores agree to 2e-5: only samples AT the threshold may differ
        near = np.abs(mz - thr) <= tol * abs(thr)
        return not ((np.asarray(got) != want) & ~near).any()

    # "#strainer gan.py:331-360": detect_outliers(dataset, feature_extractor) / (..., user_threshold)
    from strainer_b200 import detect_outliers
    got = detect_outliers(fds, ident)
    assert isinstance(got, np.ndarray) and same(got, golden["g5_elbow_inlier"], float(golden["g4_threshold"]))
    assert same(detect_outliers(fds, ident, 5.0), golden["g5_user5_inlier"], 5.0)
    # "#z_score.py:276-294": detect_outliers(dataset, feature_extractor, threshold=5.0); 4.5 passed positionally
    from strainer_b200 import detect_outliers_fixed as detect_outliers_z
    got = detect_outliers_z(fds, ident)
    assert isinstance(got, torch.Tensor) and got.dtype == torch.bool and same(got.numpy(), golden["g5_fixed5_inlier"], 5.0)
    assert same(detect_outliers_z(fds, ident, 4.5).numpy(), golden["g5_fixed45_inlier"], 4.5)
    # "#z_score + 엘보우 threshold.py:306-330": detect_outliers(dataset, feature_extractor)
    from strainer_b200 import detect_outliers_elbow
    assert same(detect_outliers_elbow(fds, ident), golden["g5_elbow_inlier"], float(golden["g4_threshold"]))
    # "# z_score + DBSCAN.py:351": detect_outliers(dataset, feature_extractor, clean_ratio) -- POSITIONAL third argument
    from strainer_b200 import detect_outliers_ratio as detect_outliers_r
    for tag in ("a", "b"):
        cr = float(golden[f"g5_ratio{tag}"])
        thr = float(torch.quantile(torch.from_numpy(mz), cr))
        got = detect_outliers_r(fds, ident, cr)
        assert isinstance(got, torch.Tensor) and same(got.numpy(), golden[f"g5_ratio{tag}_inlier"], thr)
    with pytest.raises(TypeError):
        detect_outliers(fds, ident, 5.0, clean_ratio=0.9)


def test_estimate_ratio_dbscan_1d_standardize(sb):
    """the 1-D branch standardises with the library's moment kernels (no eager torch reductions) and equals sklearn"""
    from sklearn.cluster import DBSCAN
    from sklearn.preprocessing import StandardScaler
    rng = np.random.default_rng(11)
    v = np.concatenate([rng.lognormal(-1.2, 0.5, 3000), rng.lognormal(0.7, 0.4, 700), [25.0, 31.0]]).astype(np.float32)
    got = sb.estimate_ratio_dbscan(torch.from_numpy(v), eps=0.05, min_samples=3)
    z = StandardScaler().fit_transform(v.reshape(-1, 1).astype(np.float64))
    labels = DBSCAN(eps=0.05, min_samples=3).fit_predict(z)
    assert got == np.sum(labels != -1) / len(labels)


def test_device_argument_other_than_current(sb):
    """device='cuda:1' while torch's current device is 0: runs on device 1, leaves the current device alone, and a
    later call on device 0 still works (per-device library state)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    netD = O.make_discriminator(O.SEED)
    x = torch.from_numpy(O.synth_images(0, 96))
    ds = torch.utils.data.TensorDataset(x, torch.zeros(96))
    torch.cuda.set_device(0)
    sub0, thr0 = sb.refine_dataset_by_loss(ds, netD, "cuda:0", 0.2)
    sub1, thr1 = sb.refine_dataset_by_loss(ds, netD, "cuda:1", 0.2)
    assert torch.cuda.current_device() == 0
    assert thr0 == thr1 and np.array_equal(np.asarray(sub0.indices), np.asarray(sub1.indices))
    l1 = sb.evaluate_dataset(netD, ds, torch.device("cuda", 1), return_device=True)
    assert l1.device.index == 1
    sub0b, thr0b = sb.refine_dataset_by_loss(ds, netD, "cuda:0", 0.2)
    assert thr0b == thr0
    idx, _ = sb.select_below_percentile(l1, 80.0)
    assert np.array_equal(idx, np.asarray(sub1.indices))


def test_mlp_tensor_core_chain_dataset_scale(sb):
    """config 1 at dataset scale with the reference's MLP D: the tcgen05 GEMM chain (fp16 operands) against torch fp32 on
    the CPU -- losses within 1e-3, the refine mask exact away from the threshold; ragged tails (M and K) included"""
    torch.manual_seed(3)
    d = O.MLPDiscriminator().eval()
    with torch.no_grad():
        for p in d.parameters():              # logits spread over [-1.1, 0.55]: with fp16 operands the logit error is
            p.mul_(2.0)                       # ~5e-4 absolute, which IS the relative loss error of a confident sample
    n = 5000                                  # 2 chunks of 4096 (the second: 904 rows = 7 M tiles + a 8-row tail)
    x = torch.from_numpy(O.synth_images28(100, n)).reshape(n, 784)
    with torch.no_grad():
        want_p = d(x).reshape(-1)
    want = O.bce_vs_ones(want_p).numpy()
    sc = sb.MLPScorer(d, "cuda", max_batch=4096, mode="fp16")
    out = sc.score(x.cuda(), ("loss", "prob", "logit"))
    loss = out["loss"].cpu().numpy()
    rel = np.abs(loss - want) / np.maximum(np.abs(want), ATOL)
    assert rel.max() <= 1e-3, rel.max()
    assert (out["prob"].cpu() - want_p).abs().max().item() <= 5e-4
    # host rows, the auto mode (tensor cores from 1024 rows on) and the reference-facing call
    ds = torch.utils.data.TensorDataset(x.reshape(n, 1, 28, 28), torch.zeros(n))
    ev = sb.evaluate_dataset(d, ds, "cuda")
    assert np.array_equal(ev, loss)
    sub, thr = sb.refine_dataset_by_loss(ds, d, "cuda", 0.2)
    wthr = np.percentile(want, 80.0)
    assert abs(thr - wthr) <= 1e-3 * abs(wthr)
    got = np.zeros(n, bool)
    got[np.asarray(sub.indices)] = True
    near = np.abs(want - wthr) <= 1e-3 * abs(wthr)
    assert not ((got != (want < wthr)) & ~near).any()
    # small batches keep the fp32 kernels in auto mode (bit-close to torch)
    small = sb.get_mlp_scorer(d, "cuda", 512).score(x[:64].cuda(), ("prob",))["prob"]
    assert (small.cpu() - want_p[:64]).abs().max().item() <= 1e-5


@pytest.mark.parametrize("n", [1, 64, 300, 4096 + 77])
def test_dcgan28_conv_discriminator(sb, n):
    """BASELINE config 1 with the repo-defined DCGAN-28 conv D (SURVEY 8d C1 option ii): its oracle is the same net in
    torch.nn on the CPU; conv 2 runs as a tcgen05 GEMM over im2col rows written by the conv-1 kernel"""
    d = O.make_discriminator28(O.SEED).eval()
    with torch.no_grad():
        d.main[5].weight.mul_(4.0)            # spread the logits
    x = torch.from_numpy(O.synth_images28(0, n))
    with torch.no_grad():
        want_p = d(x).reshape(-1)
    want = O.bce_vs_ones(want_p).numpy()
    sc = sb.D28Scorer(d, "cuda", max_batch=4096)
    out = sc.score(x.cuda(), ("loss", "prob"))
    rel = np.abs(out["loss"].cpu().numpy() - want) / np.maximum(np.abs(want), ATOL)
    assert rel.max() <= 1e-3, rel.max()
    assert (out["prob"].cpu() - want_p).abs().max().item() <= 5e-4
    if n >= 64:
        # the in-batch block of config 1: batch 64, top-10 % removal -> 7 strained reals
        fr_w, ff_w, mask_w, thr_w = O.strain_scores(x[:64], want_p[:64])
        fr, ff, mask, thr = sb.strain_batch(d, x[:64].cuda())
        near = (want_p[:64] - thr_w).abs() <= 1e-3 * thr_w.abs()
        assert not ((mask.cpu() != mask_w) & ~near).any()
        assert ff.shape[0] == 7 and fr.shape[0] == 57
        with pytest.raises(NotImplementedError):
            sb.strain_batch(d.train(), x[:64].cuda())
        d.eval()
    if n > 4096:
        ds = torch.utils.data.TensorDataset(x, torch.zeros(n))
        sub, thr = sb.refine_dataset_by_loss(ds, d, "cuda", 0.1)
        wthr = np.percentile(want, 90.0)
        assert abs(thr - wthr) <= 1e-3 * abs(wthr)
        got = np.zeros(n, bool)
        got[np.asarray(sub.indices)] = True
        near = np.abs(want - wthr) <= 1e-3 * abs(wthr)
        assert not ((got != (want < wthr)) & ~near).any()


def test_select_step_peer_single_rank(sb):
    """the NVLink peer-memory all-reduce fused into the select step, exercised with ONE rank (its own buffer is the only
    peer; the 2/4/8-rank identity runs are tools/multi_gpu_check.py and bench.py's multi_gpu_parity): same order
    statistics as the single-device select over many calls (both parity slots, sequence numbers)"""
    import ctypes
    L = sb._lib
    lib = L.init(torch.cuda.current_device())
    buf = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(64)
    L.check(lib.sg_peer_alloc(1, ctypes.byref(buf), handle), "sg_peer_alloc")
    try:
        class OneRank:
            table = torch.tensor([buf.value], dtype=torch.int64, device="cuda")
            rank, nranks, seq = 0, 1, 0

            def next_seq(self):
                self.seq += 1
                return self.seq
        comm = OneRank()
        rng = np.random.default_rng(9)
        for n in (1, 7, 1000, 70001):
            v = rng.standard_normal(n).astype(np.float32)
            if n > 100:
                v[::7] = np.round(v[::7], 1)            # ties
            d = torch.from_numpy(v).cuda()
            s = np.sort(v)
            for k in sorted({0, n // 10, n // 2, n - 1}):
                got = sb.order_stats(d, k, comm=comm).cpu().numpy()
                assert got[0] == s[k] and got[1] == s[min(k + 1, n - 1)], (n, k, got)
        v[5] = np.nan
        assert np.isnan(sb.order_stats(torch.from_numpy(v).cuda(), 3, comm=comm).cpu().numpy()).all()
        assert comm.seq > 40
    finally:
        torch.cuda.synchronize()
        L.check(lib.sg_peer_free(buf), "sg_peer_free")


def test_autoencoder_linear_halo_forms_reconstruction_and_alignment_contract(sb):
    """The decoder's last two layers run in linear-halo form (one contiguous bulk copy per channel group, filter shifts as
    descriptor offsets, zero halo written by the producing layer).  Through the C ABI: the reconstruction equals the oracle's
    AutoEncoder forward ("#autoencoder.py:269-291"), the per-sample errors equal the mean squared difference of that
    reconstruction ("#autoencoder.py:315-316") -- for image counts around the tile sizes, twice in the same workspace (the
    halo must survive reuse with another batch size) -- and an input that is not 16-byte aligned is refused with SG_EINVAL
    instead of reaching a kernel."""
    from strainer_gan_b200 import api
    dev = torch.device("cuda", 0)
    lib = api._lib_for(dev)
    torch.manual_seed(3)
    ae = O.AutoEncoder().eval()
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for p in ae.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.7 / np.sqrt(max(p[0].numel(), 1))))
    params = api._ae_params(ae, dev)
    arr = (api.L.P * 12)(*[t.data_ptr() for t in params])
    for mode, tol in ((api.L.SG_CONV_FP16, 2e-3), (api.L.SG_CONV_BF16, 2e-2)):
        ws = api._aligned_empty(lib.sg_ae_tc_workspace_bytes(150, mode), dev)
        ws.fill_(0x7f)                                     # NaN patterns wherever a kernel forgets to write a halo
        for n in (150, 1, 9, 2, 150):
            x = torch.from_numpy(O.synth_images(40, n))
            with torch.no_grad():
                want = ae(x).numpy()
            xd = x.to(dev)
            err = torch.empty(n, device=dev)
            rec = torch.full((n, 3, 64, 64), float("nan"), device=dev)
            api.L.check(lib.sg_ae_score_tc(api._p(xd), n, arr, api._p(ws), mode, api._p(err), api._p(rec), api._stream()), "ae")
            api.L.check(lib.sg_ae_bf16_check(api._p(ws), api._stream()), "check")
            r = rec.cpu().numpy()
            assert np.isfinite(r).all()
            assert np.abs(r - want).max() <= tol, (mode, n, np.abs(r - want).max())
            mse = ((r.astype(np.float64) - x.numpy()) ** 2).reshape(n, -1).mean(1)
            e = err.cpu().numpy()
            assert (np.abs(e - mse) / mse).max() <= 1e-5, (mode, n)
        # the same images 4 bytes further: refused, nothing launched
        buf = torch.empty(2 * 12288 + 4, device=dev)
        xu = buf[1:1 + 2 * 12288].view(2, 3, 64, 64)
        assert xu.data_ptr() % 16 == 4
        rc = lib.sg_ae_score_tc(api._p(xu), 2, arr, api._p(ws), mode, api._p(err), api.L.P(0), api._stream())
        assert rc == -1, rc          # SG_EINVAL
        torch.cuda.synchronize()
