"""uint8 datasets: the reference's transform chain ToTensor() + Normalize(mean, std) ("#strainer gan.py:85-91,
113-116") computed on the device (sg_u8_normalize) must be bit-identical to the host transform, so that straining a
U8ImageDataset gives the same thresholds and kept-index lists as straining the fp32 tensors a DataLoader yields."""
import numpy as np
import pytest
import torch

from oracle import strainer_oracle as O


@pytest.fixture(scope="module")
def sb():
    import strainer_b200
    return strainer_b200


def host_transform(px_nchw_u8, mean, std):
    """ToTensor + Normalize: the oracle's restatement (numpy fp32 arithmetic, one rounding per step)."""
    return O.to_tensor_normalize(px_nchw_u8, mean, std)


def test_dataset_getitem_equals_torchvision(sb):
    """CPU: what U8ImageDataset yields == torchvision's own ToTensor + Normalize on the PIL image (the reference's
    transform), and == the numpy restatement used by the GPU tests."""
    tv = pytest.importorskip("torchvision.transforms")
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    hwc = rng.integers(0, 256, size=(5, 64, 64, 3), dtype=np.uint8)
    hwc[0, :4, :64, 0] = np.arange(256, dtype=np.uint8).reshape(4, 64)   # every pixel value at least once
    for mean, std in (((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        tf = tv.Compose([tv.ToTensor(), tv.Normalize(mean, std)])
        want = torch.stack([tf(Image.fromarray(im)) for im in hwc])
        ds_hwc = sb.U8ImageDataset(torch.from_numpy(hwc), None, mean, std, layout="NHWC")
        ds_chw = sb.U8ImageDataset(torch.from_numpy(hwc).permute(0, 3, 1, 2).contiguous(), torch.arange(5), mean, std)
        for i in range(5):
            assert torch.equal(ds_hwc[i][0], want[i])
            assert torch.equal(ds_chw[i][0], want[i])
            assert int(ds_chw[i][1]) == i
        assert np.array_equal(host_transform(hwc.transpose(0, 3, 1, 2), mean, std), want.numpy())
    # a DataLoader over it batches like any map-style dataset (what unmodified reference code would do)
    xb, yb = next(iter(torch.utils.data.DataLoader(ds_chw, batch_size=4, shuffle=False)))
    assert torch.equal(xb, want[:4]) and yb.tolist() == [0, 1, 2, 3]


def test_u8_images_host_logic(sb):
    """CPU: slicing / fancy indexing / Subset resolution of a uint8 dataset keep pixels, layout and normalisation; the
    resolver hands Subset-of-Subset chains of uint8 datasets to the scorer as uint8 rows in the right order."""
    rng = np.random.default_rng(5)
    px = torch.from_numpy(rng.integers(0, 256, size=(20, 64, 64, 3), dtype=np.uint8))
    imgs = sb.U8Images(px, (0.4, 0.5, 0.6), (0.2, 0.3, 0.4), layout="NHWC")
    assert imgs.shape == (20, 3, 64, 64) and len(imgs) == 20 and not imgs.is_cuda
    part = imgs[3:9]
    assert isinstance(part, sb.U8Images) and part.shape == (6, 3, 64, 64) and part.mean == imgs.mean and part.layout == "NHWC"
    assert torch.equal(part.pixels, px[3:9])
    pick = imgs[np.array([7, 2, 2, 19])]
    assert torch.equal(pick.pixels, px[[7, 2, 2, 19]])
    assert torch.equal(pick.host_f32(1), imgs.host_f32(2))
    assert imgs[4].shape == (1, 3, 64, 64) and torch.equal(imgs[4].pixels[0], px[4])
    ds = sb.U8ImageDataset(imgs, torch.arange(20))
    sub = torch.utils.data.Subset(torch.utils.data.Subset(ds, [1, 5, 9, 13]), [3, 0])
    from strainer_gan_b200.api import _dataset_images
    got = _dataset_images(sub)
    assert isinstance(got, sb.U8Images) and torch.equal(got.pixels, px[[13, 1]])
    assert torch.equal(sub[0][0], imgs.host_f32(13)) and int(sub[0][1]) == 13
    with pytest.raises(ValueError):
        sb.U8Images(px.float())
    with pytest.raises(ValueError):
        sb.U8Images(px, (0.5,), (0.5,), layout="NHWC")
    with pytest.raises(ValueError):
        sb.U8Images(px, layout="CHWN")
    # without a GPU every compute entry point fails loudly (no CPU fallback), including the uint8 path
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            imgs.to_f32()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(7, 3, 64, 64), (3, 1, 28, 28), (2, 4, 5, 3), (1, 3, 16, 16), (0, 3, 64, 64)])
@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
def test_u8_normalize_bit_exact(sb, shape, layout):
    rng = np.random.default_rng(11)
    n, c, h, w = shape
    px = rng.integers(0, 256, size=shape, dtype=np.uint8)
    k = min(256, px.size)
    px.reshape(-1)[:k] = np.arange(k, dtype=np.uint8)   # every pixel value
    mean = [0.5, 0.456, 0.0, 0.25][:c]
    std = [0.5, 0.224, 1.0, 3.0][:c]
    want = host_transform(px, mean, std)
    src = px if layout == "NCHW" else np.ascontiguousarray(px.transpose(0, 2, 3, 1))
    for dev in ("cpu", "cuda"):
        imgs = sb.U8Images(torch.from_numpy(src).to(dev), mean, std, layout)
        assert imgs.shape == shape
        got = imgs.to_f32("cuda").cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want)
    # unaligned views take the generic kernel
    if n >= 2 and layout == "NCHW":
        flat = torch.zeros(px.size + 1, dtype=torch.uint8, device="cuda")
        flat[1:] = torch.from_numpy(px).reshape(-1).cuda()
        imgs = sb.U8Images(flat[1:].view(shape), mean, std)
        assert imgs.pixels.data_ptr() % 16 != 0
        assert np.array_equal(imgs.to_f32().cpu().numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("where", ["pageable", "pinned", "cuda"])
def test_refine_u8_dataset_equals_fp32_dataset(sb, where):
    """Same thresholds / kept indices / losses (bit for bit) whether the dataset is fp32 or uint8 + device normalise; and
    the oracle run on what the uint8 dataset's __getitem__ yields agrees within the fp32-mode tolerance."""
    n = 300
    netD = O.make_discriminator(O.SEED)
    f = O.synth_images(0, n)
    px = torch.from_numpy(np.clip(np.rint((f + 1.0) * 127.5), 0, 255).astype(np.uint8))
    if where == "pinned":
        px = px.pin_memory()
    elif where == "cuda":
        px = px.cuda()
    ds8 = sb.U8ImageDataset(px, torch.zeros(n, dtype=torch.long))
    x32 = torch.from_numpy(host_transform(px.cpu().numpy(), (0.5,) * 3, (0.5,) * 3))
    assert torch.equal(ds8[5][0], x32[5])
    ds32 = torch.utils.data.TensorDataset(x32, torch.zeros(n, dtype=torch.long))
    sc = sb.get_scorer(netD, "cuda", "fp32", max_batch=128)   # several chunks + a ragged tail through the double buffers
    l8 = sc.score(ds8.images, ("loss",))["loss"].cpu()
    l32 = sc.score(x32, ("loss",))["loss"].cpu()
    assert torch.equal(l8, l32)
    sub8, thr8 = sb.refine_dataset_by_loss(ds8, netD, "cuda", 0.2)
    sub32, thr32 = sb.refine_dataset_by_loss(ds32, netD, "cuda", 0.2)
    assert thr8 == thr32 and np.array_equal(np.asarray(sub8.indices), np.asarray(sub32.indices))
    assert sub8.dataset is ds8
    widx, wthr, wloss = O.refine_dataset_by_loss(x32, netD, 0.2)
    assert abs(thr8 - wthr) <= 1e-3 * abs(wthr)
    got = np.zeros(n, bool)
    got[np.asarray(sub8.indices)] = True
    want = np.zeros(n, bool)
    want[widx] = True
    near = np.abs(wloss.reshape(-1) - wthr) <= 1e-3 * abs(wthr)
    assert not ((got != want) & ~near).any()
    # Subset of a uint8 dataset (the second straining pass of "# final.py:444") stays uint8
    sub_again, _ = sb.refine_dataset_by_loss(sub8, netD, "cuda", 0.2)
    ref_again, _ = sb.refine_dataset_by_loss(torch.utils.data.Subset(ds32, sub32.indices), netD, "cuda", 0.2)
    assert np.array_equal(np.asarray(sub_again.indices), np.asarray(ref_again.indices))


@pytest.mark.gpu
def test_autoencoder_u8_dataset(sb):
    n = 70
    torch.manual_seed(O.SEED)
    ae = O.AutoEncoder()
    f = O.synth_images(1000, n)
    px = torch.from_numpy(np.clip(np.rint((f + 1.0) * 127.5), 0, 255).astype(np.uint8))
    ds8 = sb.U8ImageDataset(px)
    x32 = torch.from_numpy(host_transform(px.numpy(), (0.5,) * 3, (0.5,) * 3))
    m8 = sb.detect_outliers_autoencoder(ae, ds8, "cuda")
    m32 = sb.detect_outliers_autoencoder(ae, torch.utils.data.TensorDataset(x32, torch.zeros(n)), "cuda")
    assert torch.equal(m8, m32)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
def test_resident_subset_u8(sb, layout):
    """Device-resident epochs over uint8 rows: same kept set as the fp32 resident dataset, batches = gather + normalise."""
    n = 200
    netD = O.make_discriminator(O.SEED)
    px = np.clip(np.rint((O.synth_images(50, n) + 1.0) * 127.5), 0, 255).astype(np.uint8)
    x32 = torch.from_numpy(host_transform(px, (0.5,) * 3, (0.5,) * 3)).cuda()
    src = px if layout == "NCHW" else np.ascontiguousarray(px.transpose(0, 2, 3, 1))
    u8 = sb.U8Images(torch.from_numpy(src).cuda(), layout=layout)
    s8 = sb.ResidentSubset.refine(u8, netD, 0.2)
    s32 = sb.ResidentSubset.refine(x32, netD, 0.2)
    assert torch.equal(s8.indices, s32.indices) and torch.equal(s8.threshold, s32.threshold)
    b8 = list(s8.batches(64, shuffle=False))
    b32 = list(s32.batches(64, shuffle=False))
    assert len(b8) == len(b32) and all(torch.equal(a, b) for a, b in zip(b8, b32))
    g8, g32 = torch.Generator(device="cuda").manual_seed(3), torch.Generator(device="cuda").manual_seed(3)
    for a, b in zip(s8.batches(32, generator=g8), s32.batches(32, generator=g32)):
        assert a.dtype == torch.float32 and torch.equal(a, b)
