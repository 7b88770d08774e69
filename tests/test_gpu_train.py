"""D64 training step on the library's kernels (csrc/d64_train.cu, strainer-gan_b200/train.py) against the reference's own
autograd arithmetic: the golden fixture produced by the reference Discriminator on the CPU (fp32), and the oracle's
restatement of "#strainer gan.py:586-615" at the batch sizes of the training configs.

TOLERANCES.  The path computes in fp16 operands with fp32 accumulation (11-bit significands, the arithmetic class of the
TF32 convolutions torch itself uses by default on this GPU).  Forward outputs: within 5e-3 relative per sample and 1e-3 in
relative L2.  Gradients: the D step's BatchNorm backward subtracts the batch-mean components of a gradient whose samples all
push the same way (every label equal), so operand rounding of ~3e-4 is amplified to ~1e-2 in the early layers for ANY
11-bit-operand implementation -- torch's default TF32 autograd differs from fp32 autograd by 0.5-2 % on the same problem
(tools/train_check.py prints both).  The bar here: relative L2 error against fp32 autograd <= 5e-2 per tensor, cosine
similarity >= 0.999, and (on the GPU, same inputs) no more than 2.5x torch-TF32's own deviation.
"""
import copy

import numpy as np
import pytest
import torch
from torch import nn

from oracle import strainer_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import strainer_b200
    assert torch.cuda.is_available()
    return strainer_b200


def _rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm())


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm()))


def _step(D, real, fake, **g_kw):
    """the literal step of O.train_step_through_d with D = the accelerated module"""
    crit = nn.BCELoss()
    fake = fake.detach().clone().requires_grad_(True)
    D.zero_grad()
    out_real = D(real).view(-1)
    crit(out_real, torch.ones_like(out_real)).backward()
    out_fake = D(fake.detach()).view(-1)
    crit(out_fake, torch.zeros_like(out_fake)).backward()
    d_grads = [p.grad.detach().clone() for p in O.d64_params(D)]
    out_g = D(fake, **g_kw).view(-1)
    errG = crit(out_g, torch.ones_like(out_g))
    errG.backward()
    return dict(out_real=out_real.detach(), out_fake=out_fake.detach(), out_g=out_g.detach(), d_grads=d_grads,
                dfake=fake.grad.detach(), errG=errG.detach())


def test_train_step_vs_reference_golden(sb, golden3):
    g = golden3
    B = g["g10_out_real"].shape[0]
    real = torch.from_numpy(O.synth_images(2000, B)).cuda()
    fake = torch.tanh(torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(4321))).cuda()
    netD = O.make_discriminator(O.SEED).cuda().train()
    D = sb.accelerate_discriminator(netD, max_batch=B)
    r = _step(D, real, fake)
    D.check()
    for k in ("out_real", "out_fake", "out_g"):
        got, want = r[k].cpu().numpy(), g["g10_" + k]
        assert np.abs(got - want).max() <= 5e-3 * np.abs(want).max(), k
        assert np.linalg.norm(got - want) <= 1e-3 * np.linalg.norm(want), k
    assert abs(float(r["errG"]) - float(g["g10_errG"])) <= 2e-3 * abs(float(g["g10_errG"]))
    for n, gr in zip(O.D64_PARAM_NAMES, r["d_grads"]):
        want = torch.from_numpy(g[f"g10_d_{n}_sample"])
        got = (gr.reshape(-1)[::61] if gr.numel() > 4096 else gr.reshape(-1)).cpu()
        assert _rel(got, want) <= 5e-2 and _cos(got, want) >= 0.999, (n, _rel(got, want))
        assert abs(float(gr.double().norm()) - float(g[f"g10_d_{n}_norm"])) <= 2e-2 * float(g[f"g10_d_{n}_norm"]), n
    want = torch.from_numpy(g["g10_dfake_sample"])
    got = r["dfake"][:, :, ::8, ::8].cpu()
    assert _rel(got, want) <= 5e-2 and _cos(got, want) >= 0.999
    assert abs(float(r["dfake"].double().norm()) - float(g["g10_dfake_norm"])) <= 2e-2 * float(g["g10_dfake_norm"])
    # BatchNorm side effects: three training-mode forward passes
    bns = [m for m in netD.modules() if isinstance(m, nn.BatchNorm2d)]
    for i, bn in enumerate(bns):
        assert np.allclose(bn.running_mean.cpu().numpy(), g[f"g10_bn{i + 2}_mean"], rtol=2e-3, atol=2e-4)
        assert np.allclose(bn.running_var.cpu().numpy(), g[f"g10_bn{i + 2}_var"], rtol=2e-3, atol=1e-5)
    assert int(bns[0].num_batches_tracked) == int(g["g10_nbt"])


@pytest.mark.parametrize("B", [128, 115, 6])
def test_train_step_vs_oracle_and_tf32(sb, B):
    """batch 128 (BASELINE configs 2-4's training batch), 115 (the strained real batch: ragged tiles), 6 (one partial tile)"""
    real_h = torch.from_numpy(O.synth_images(50, B))
    fake_h = torch.tanh(torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(B)))
    want = O.train_step_through_d(O.make_discriminator(O.SEED).train(), real_h, fake_h)          # CPU fp32
    netD = O.make_discriminator(O.SEED).cuda().train()
    D = sb.accelerate_discriminator(netD, max_batch=128)
    r = _step(D, real_h.cuda(), fake_h.cuda(), param_grads=False)
    D.check()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        tf = O.train_step_through_d(O.make_discriminator(O.SEED).cuda().train(), real_h.cuda(), fake_h.cuda())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    for k in ("out_real", "out_fake", "out_g"):
        assert torch.allclose(r[k].cpu(), want[k], rtol=5e-3, atol=1e-6), k
    for n, got, w, t in zip(O.D64_PARAM_NAMES, r["d_grads"], want["d_grads"], tf["d_grads"]):
        e, et = _rel(got, w), _rel(t, w)
        assert e <= 5e-2 and _cos(got, w) >= 0.999, (n, e)
        assert e <= max(2.5 * et, 2e-3), (n, e, et)
    e, et = _rel(r["dfake"], want["dfake"]), _rel(tf["dfake"], want["dfake"])
    assert e <= 5e-2 and e <= max(2.5 * et, 2e-3), (e, et)
    # param_grads=False (the G step) left the D-step gradients untouched
    for p, gr in zip(O.d64_params(D), r["d_grads"]):
        assert torch.equal(p.grad, gr)


def test_trainable_d64_module_contract(sb):
    netD = O.make_discriminator(O.SEED).cuda().train()
    D = sb.accelerate_discriminator(netD)
    assert [id(p) for p in D.parameters()] == [id(p) for p in netD.parameters()]      # the optimiser's tensors
    x = torch.from_numpy(O.synth_images(0, 16)).cuda()
    v0 = [bn.running_mean._version for bn in netD.modules() if isinstance(bn, nn.BatchNorm2d)]
    with torch.no_grad():
        p = D(x)
    assert p.shape == (16, 1, 1, 1) and not p.requires_grad
    v1 = [bn.running_mean._version for bn in netD.modules() if isinstance(bn, nn.BatchNorm2d)]
    assert all(b > a for a, b in zip(v0, v1))
    # two forwards before their backwards: each keeps its own saved tensors
    pa, pb = D(x), D(x.flip(0))
    (pa.sum() + pb.sum()).backward()
    ga = [p.grad.clone() for p in O.d64_params(D)]
    D.zero_grad()
    D(x).sum().backward()
    D(x.flip(0)).sum().backward()
    for a, b in zip(ga, [p.grad for p in O.d64_params(D)]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7)
    # eval mode: the wrapped module itself
    D.eval()
    with torch.no_grad():
        assert torch.equal(D(x), netD(x))
    D.train()
    with pytest.raises(RuntimeError):
        D(x.cpu())
    with pytest.raises(ValueError):
        D(x[:1])
    D.check()


def test_identical_weight_updates_over_iterations(sb):
    """five optimiser iterations of the D step on both implementations from the same start: the parameters stay together.
    SGD with momentum, not the reference's Adam: Adam's first steps are lr * sign(g) per element, which turns a 1 % gradient
    difference on a near-zero element into a full-size step difference and measures nothing about the gradients."""
    torch.manual_seed(0)
    a = O.make_discriminator(O.SEED).cuda().train()
    b = copy.deepcopy(a)
    start = [p.detach().clone() for p in a.parameters()]
    Db = sb.accelerate_discriminator(b)
    oa = torch.optim.SGD(a.parameters(), lr=2e-3, momentum=0.5)
    ob = torch.optim.SGD(b.parameters(), lr=2e-3, momentum=0.5)
    crit = nn.BCELoss()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for it in range(5):
            real = torch.from_numpy(O.synth_images(64 * it, 64)).cuda()
            fake = torch.tanh(torch.randn(64, 3, 64, 64, device="cuda"))
            for D, opt in ((a, oa), (Db, ob)):
                D.zero_grad()
                o = D(real).view(-1)
                crit(o, torch.ones_like(o)).backward()
                o = D(fake).view(-1)
                crit(o, torch.zeros_like(o)).backward()
                opt.step()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    Db.check()
    for p0, pa, pb in zip(start, a.parameters(), b.parameters()):
        ua, ub = (pa - p0).detach().double().flatten(), (pb - p0).detach().double().flatten()
        assert float((ua - ub).norm() / ua.norm()) <= 0.1, float((ua - ub).norm() / ua.norm())
        assert float((ua @ ub) / (ua.norm() * ub.norm())) >= 0.995
    o1, o2 = a(real).view(-1), Db(real).view(-1)
    assert torch.allclose(o1, o2, rtol=2e-2, atol=1e-4)


@pytest.mark.parametrize("B", [64, 37])
def test_train_step_fp32_parity_precision(sb, B):
    """precision='fp32' (split operands x = hi + lo, three tensor passes per GEMM): the D step and the G step's pass through
    D against the oracle run in FLOAT64 on the CPU (the truth).  Outputs within 1e-4 (measured 1e-5).  Gradients: hi + lo
    carry 22 significant bits, pre-activations agree with the truth to ~1e-6 -- but LeakyReLU's derivative is discontinuous:
    an element whose pre-activation is within that distance of 0 takes the other slope, and ONE such flip among the 2 M
    elements of a layer changes that layer's gradient by ~7e-4 of its norm.  With a handful of flips per pass (fp32 autograd
    has them too, 7x fewer: its own conv1 gradient is 3e-4 .. 5e-4 from the truth here; in layer 4, 0.5 M elements, one flip
    is 1.4e-3) the bar is 3e-3 on every tensor with cosine >= 0.99999; tensors without a flip agree to 1e-6 .. 3e-4."""
    real_h = torch.from_numpy(O.synth_images(900, B))
    fake_h = torch.tanh(torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(1000 + B)))
    want = O.train_step_through_d(O.make_discriminator(O.SEED).train().double(), real_h.double(), fake_h.double())
    f32 = O.train_step_through_d(O.make_discriminator(O.SEED).train(), real_h, fake_h)
    netD = O.make_discriminator(O.SEED).cuda().train()
    D = sb.accelerate_discriminator(netD, max_batch=128, precision="fp32")
    r = _step(D, real_h.cuda(), fake_h.cuda())
    D.check()
    for k in ("out_real", "out_fake", "out_g"):
        assert torch.allclose(r[k].cpu().double(), want[k], rtol=1e-4, atol=1e-7), k
    for n, got, w, f in zip(O.D64_PARAM_NAMES, r["d_grads"], want["d_grads"], f32["d_grads"]):
        e = _rel(got, w)
        assert e <= 3e-3 and _cos(got, w) >= 0.99999, (n, e, _rel(f, w))
    e = _rel(r["dfake"], want["dfake"])
    assert e <= 3e-3 and _cos(r["dfake"], want["dfake"]) >= 0.99999, (e, _rel(f32["dfake"], want["dfake"]))
    assert int([m for m in netD.modules() if isinstance(m, nn.BatchNorm2d)][0].num_batches_tracked) == 3


def test_fp16_overflow_is_reported_and_leaves_running_stats(sb):
    """an activation beyond fp16's range (conv1 weights scaled up): the pass yields non-finite values, check() reports it,
    and the BatchNorm running statistics are NOT committed (bn_commit_kernel); the split-operand precision has the same
    range and reports the same; a following ordinary batch works"""
    d = O.make_discriminator(O.SEED).cuda().train()
    with torch.no_grad():
        d.main[0].weight.mul_(4.0e6)
    before = [t.clone() for bn in d.modules() if isinstance(bn, nn.BatchNorm2d) for t in (bn.running_mean, bn.running_var)]
    x = torch.from_numpy(O.synth_images(0, 32)).cuda()
    for prec in ("fp16", "fp32"):
        D = sb.accelerate_discriminator(d, precision=prec)
        with torch.no_grad():
            D(x)
        with pytest.raises(RuntimeError, match="non-finite"):
            D.check()
        after = [t for bn in d.modules() if isinstance(bn, nn.BatchNorm2d) for t in (bn.running_mean, bn.running_var)]
        assert all(torch.equal(a, b) for a, b in zip(before, after))
        D.check()                                # the words are cleared by the report
    with torch.no_grad():
        d.main[0].weight.mul_(1.0 / 4.0e6)
    D = sb.accelerate_discriminator(d)
    with torch.no_grad():
        p = D(x)
    D.check()
    assert torch.isfinite(p).all() and not torch.equal(before[0], d.main[3].running_mean)
