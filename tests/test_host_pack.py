"""Host-packed PCIe copy (csrc/host_pack.cpp + sg_f16_expand): a fp32 HOST dataset may cross PCIe as fp16 because conv1
of the default conv mode rounds its input to fp16 anyway ("#strainer gan.py:364-375": the DataLoader's fp32 batches).
CPU: the host conversion is the IEEE round-to-nearest-even conversion in every instruction-set form (numpy's astype is
the checker).  GPU: scores, thresholds and kept indices are bit-identical with and without the packing, for pinned and
pageable sources, mixed shares and ragged tails; an fp16 overflow still falls back to the fp32 source."""
import numpy as np
import pytest
import torch

from oracle import strainer_oracle as O


@pytest.fixture(scope="module")
def sb():
    import strainer_b200
    return strainer_b200


def _edge_bits():
    return np.array([0, 0x80000000, 0x7f800000, 0xff800000, 0x7fc00000, 0x477fe000, 0x477fefff, 0x477ff000, 0x477ff001,
                     0x38800000, 0x387fffff, 0x33000000, 0x33000001, 0x32ffffff, 0x33800000, 0x337fffff, 0x38000000,
                     0x3f800000, 0x3f801000, 0x3f802000, 0x3f803000, 0x3f801001, 0x00000001, 0x007fffff], dtype=np.uint32)


@pytest.mark.parametrize("isa", [0, 1, 2, 3])
def test_host_f32_to_f16_is_ieee_rne(sb, isa):
    from strainer_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(11 + isa)
    bits = rng.integers(0, 2 ** 32, size=400_003, dtype=np.uint32)      # every exponent, both signs, NaNs, subnormals
    edge = _edge_bits()
    bits[:edge.size] = edge
    x = np.concatenate([bits.view(np.float32), rng.uniform(-1, 1, 100_000).astype(np.float32)])
    with np.errstate(all="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    nan = np.isnan(x)
    for threads in (1, 3):
        out = np.full(x.size + 5, 0xABCD, dtype=np.uint16)
        rc = lib.sg_host_f32_to_f16(x.ctypes.data, x.size, out.ctypes.data, threads, isa)
        if rc != 0:
            assert isa in (2, 3), L.last_error()        # an instruction set this CPU does not have is refused, not emulated
            pytest.skip(L.last_error())
        got = out[:x.size]
        assert np.array_equal(got[~nan], want[~nan])
        assert np.all((got[nan] & 0x7c00) == 0x7c00) and np.all((got[nan] & 0x03ff) != 0)
        assert np.all(out[x.size:] == 0xABCD)           # nothing written past the end
    # unaligned source / destination, empty input, bad arguments
    out = np.zeros(x.size + 8, dtype=np.uint16)
    assert lib.sg_host_f32_to_f16(x.ctypes.data + 4, x.size - 1, out.ctypes.data + 2, 2, isa) == 0
    assert np.array_equal(out[1:x.size][~nan[1:]], want[1:][~nan[1:]])
    assert lib.sg_host_f32_to_f16(None, 0, None, 1, isa) == 0
    assert lib.sg_host_f32_to_f16(None, 5, out.ctypes.data, 1, isa) != 0
    assert lib.sg_host_f32_to_f16(x.ctypes.data, 5, out.ctypes.data, 1, 9) != 0
    assert lib.sg_host_threads() >= 1


def test_host_to_f16_tensor_wrapper(sb):
    from strainer_b200 import api
    x = torch.from_numpy(O.synth_images(0, 33))
    out = torch.empty(x.shape, dtype=torch.float16)
    api._host_to_f16(x, out)
    assert torch.equal(out, x.to(torch.float16))
    with pytest.raises(ValueError):
        api._host_to_f16(x.double(), out)
    with pytest.raises(ValueError):
        api._host_to_f16(x[:, :, ::2], out[:, :, ::2])


def test_pack_tuner_state_machine(sb):
    """The share search: all rows first, 0.1 less per call while that is > 2 % faster, then locked on the best."""
    from strainer_b200 import api
    _PackTuner = api._PackTuner
    t = _PackTuner()
    assert t.share == 1.0 and not t.locked
    t.report(1.0, 100.0)
    assert t.share == 0.9 and not t.locked
    t.report(0.5, 500.0)                  # a call that ran with another share (explicit host_pack) is ignored
    assert t.share == 0.9
    t.report(0.9, 110.0)
    assert t.share == 0.8
    t.report(0.8, 111.0)                  # < 2 % better: back to the best, locked
    assert t.locked and t.share == 0.9
    t.report(0.9, 1.0)
    assert t.locked and t.share == 0.9
    t = _PackTuner()
    rate = 100.0
    while not t.locked:                   # a box whose host threads are too slow ends at the plain copy
        share = t.share
        rate *= 1.1
        t.report(share, rate)
    assert t.share == 0.0 and t.best_share == 0.0


@pytest.mark.gpu
def test_f16_expand_exact(sb):
    from strainer_b200 import api, _lib as L
    dev = torch.device("cuda", 0)
    lib = api._lib_for(dev)
    bits = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16)
    h = bits.view(torch.float16).repeat(3)[:-5].contiguous()             # every fp16 value; length not a multiple of 8
    for off in (0, 1):                                                   # aligned and misaligned (element fallback) views
        src = h.to(dev)[off:]
        out = torch.full((src.numel() + 4,), -7.0, dtype=torch.float32, device=dev)
        L.check(lib.sg_f16_expand(src.data_ptr(), src.numel(), out.data_ptr(), api._stream()), "sg_f16_expand")
        want = src.float()
        got = out[:src.numel()]
        assert torch.equal(got.view(torch.int32), want.view(torch.int32))   # bits, so NaN payload classes and -0 count
        assert bool((out[src.numel():] == -7.0).all())


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [True, False])
def test_packed_copy_scores_bit_identical(sb, pinned):
    """fp16-on-host + expand == fp32 copy, bit for bit, at every share (0, mixed, 1), with a ragged last chunk."""
    dev = torch.device("cuda", 0)
    n = 4096 + 300
    x = torch.from_numpy(O.synth_images(0, n))
    if pinned:
        x = x.pin_memory()
    netD = O.make_discriminator(O.SEED)
    sc = sb.D64Scorer(netD, dev, "auto", max_batch=2048)
    sc.host_pack = False
    base = sc.score(x, ("loss", "logit"))
    assert sc.last_pack_fraction == 0.0
    base = {k: v.clone() for k, v in base.items()}
    b0 = sc.h2d_bytes
    for share in (1.0, 0.37, "auto"):
        sc.host_pack = share
        before = sc.h2d_bytes
        got = sc.score(x, ("loss", "logit"))
        for k in base:
            assert torch.equal(got[k].view(torch.int32), base[k].view(torch.int32)), (share, k)
        moved = sc.h2d_bytes - before
        if share == 1.0:
            assert sc.last_pack_fraction == 1.0 and moved == n * 12288 * 2
        elif share == 0.37:
            assert n * 12288 * 2 < moved < n * 12288 * 4
        else:
            assert 0.0 <= sc.last_pack_fraction <= 1.0 and (not pinned) <= sc.last_pack_fraction
    assert b0 == n * 12288 * 4
    # the resident path and the reference-facing function agree with it too
    res = sc.score(x.to(dev), ("loss",))["loss"]
    assert torch.equal(res, base["loss"])
    sc.host_pack = "auto"
    ds = torch.utils.data.TensorDataset(x, torch.zeros(n, dtype=torch.long))
    sub, thr = sb.refine_dataset_by_loss(ds, netD, dev, 0.1)
    idx, thr2 = sb.select_below_percentile(base["loss"], 90.0)
    assert thr == thr2 and np.array_equal(np.asarray(sub.indices), idx)


@pytest.mark.gpu
def test_packed_copy_overflow_falls_back_to_fp32_source(sb):
    """A chunk that overflows fp16 is re-scored from the ORIGINAL fp32 rows (the packed copy of an out-of-range pixel is
    inf): the losses equal the unpacked auto mode's."""
    dev = torch.device("cuda", 0)
    n = 4096
    x = torch.from_numpy(O.synth_images(0, n)).clone()
    x[5] *= 3.0e5                      # beyond 65504 after the fp16 rounding
    x = x.pin_memory()
    netD = O.make_discriminator(O.SEED)
    sc = sb.D64Scorer(netD, dev, "auto", max_batch=2048)
    sc.host_pack = False
    want = sc.score(x, ("loss",))["loss"].clone()
    f0 = sc.fallback_chunks
    sc.host_pack = 1.0
    got = sc.score(x, ("loss",))["loss"]
    assert sc.fallback_chunks == f0 + 1
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))
