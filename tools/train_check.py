"""Stage-by-stage check of the D64 training path (csrc/d64_train.cu) against torch autograd in fp32 (TF32 off) on the GPU:
saved activations after every layer, every parameter gradient, the input gradient, BatchNorm running statistics; then the
timing of forward + backward against autograd.  One JSON object on stdout.

  python tools/train_check.py [--batch 128] [--no-time]
"""
import argparse
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch import nn  # noqa: E402

import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402
from strainer_gan_b200 import _lib as L  # noqa: E402
from strainer_gan_b200 import api as A  # noqa: E402
from strainer_gan_b200 import train as T  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def maxrel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def gpu_time(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--no-time", action="store_true")
    ap.add_argument("--precision", default="fp16")
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B = a.batch
    out = {"batch": B}

    ref = O.make_discriminator(O.SEED).to(dev).train()
    mine = copy.deepcopy(ref)
    x = sb.synth_images(0, B, O.SEED, dev)

    # ---- reference forward with hooks ------------------------------------------------------------------------------------
    feats = {}
    mods = [m for m in ref.modules() if isinstance(m, (nn.Conv2d, nn.BatchNorm2d, nn.LeakyReLU))]
    for i, m in enumerate(mods):
        m.register_forward_hook(lambda mod, inp, o, i=i: feats.__setitem__(i, o.detach().clone()))
    xr = x.clone().requires_grad_(True)
    pr = ref(xr).view(-1)
    target = torch.ones_like(pr)
    loss_r = nn.functional.binary_cross_entropy(pr, target)
    loss_r.backward()
    # module order: conv1 lrelu | conv2 bn2 lrelu | conv3 bn3 lrelu | conv4 bn4 lrelu | conv5
    kinds = [type(m).__name__ for m in mods]
    out["ref_modules"] = kinds

    # ---- stage check through the C ABI -----------------------------------------------------------------------------------
    lib = A._lib_for(dev)
    prec = 1 if a.precision == "fp32" else 0
    out["precision"] = a.precision
    ws = T._Workspace(dev, lib, B, prec)
    wrapped = sb.TrainableD64(mine, max_batch=B, precision=a.precision)
    params = wrapped._params()
    stats = wrapped._running_stats()
    prob = torch.empty(B, device=dev)
    logit = torch.empty(B, device=dev)
    packed, _ = wrapped._packed_for(dev, lib, params)
    L.check(lib.sg_d64_train_forward(L.P(x.data_ptr()), B, B, prec, L.P(packed.data_ptr()), T._ptr_array(params[5:]), T._ptr_array(stats), 0.1, 1e-5,
                                     L.P(ws.buf.data_ptr()), L.P(prob.data_ptr()), L.P(logit.data_ptr()), A._stream()), "fwd")
    torch.cuda.synchronize()

    def read(what, c, s):
        t = torch.empty(B, c, s, s, device=dev)
        L.check(lib.sg_d64_train_read(L.P(ws.buf.data_ptr()), B, B, prec, what, L.P(t.data_ptr()), A._stream()), "read")
        torch.cuda.synchronize()
        return t
    stages = {}
    stages["act1"] = rel(read(1, 64, 32), feats[1])
    stages["raw2"] = rel(read(2, 128, 16), feats[2])
    stages["act2"] = rel(read(5, 128, 16), feats[4])
    stages["raw3"] = rel(read(3, 256, 8), feats[5])
    stages["act3"] = rel(read(6, 256, 8), feats[7])
    stages["raw4"] = rel(read(4, 512, 4), feats[8])
    stages["act4"] = rel(read(7, 512, 4), feats[10])
    stages["prob"] = rel(prob, pr.detach())
    stages["prob_max_rel"] = float(((prob - pr.detach()).abs() / pr.detach().abs()).max())
    out["forward_rel_l2"] = stages
    rbn = [m for m in ref.modules() if isinstance(m, nn.BatchNorm2d)]
    mbn = [m for m in mine.modules() if isinstance(m, nn.BatchNorm2d)]
    out["running_stats_rel_l2"] = [[rel(m.running_mean, r.running_mean), rel(m.running_var, r.running_var)] for m, r in zip(mbn, rbn)]

    gprob = (-(target / pr.detach()) / B).contiguous()      # d BCE(mean) / d prob for target 1
    grads = [torch.zeros_like(p) for p in params]
    gx = torch.zeros_like(x)
    L.check(lib.sg_d64_train_backward(L.P(gprob.data_ptr()), B, B, prec, L.P(packed.data_ptr()), L.P(ws.buf.data_ptr()), T._ptr_array(grads), L.P(gx.data_ptr()),
                                      A._stream()), "bwd")
    torch.cuda.synchronize()
    rc = lib.sg_d64_train_check(L.P(ws.buf.data_ptr()), A._stream())
    out["check_rc"] = int(rc)
    if rc != 0:
        out["check_error"] = L.last_error()
    rparams = [c.weight for c in ref.modules() if isinstance(c, nn.Conv2d)] + [t for bn in rbn for t in (bn.weight, bn.bias)]
    names = ["w1", "w2", "w3", "w4", "w5", "g2", "b2", "g3", "b3", "g4", "b4"]
    out["grad_rel_l2"] = {n: rel(g, p.grad) for n, g, p in zip(names, grads, rparams)}
    out["grad_max_rel"] = {n: maxrel(g, p.grad) for n, g, p in zip(names, grads, rparams)}
    out["grad_x_rel_l2"] = rel(gx, xr.grad)
    out["grad_x_max_rel"] = maxrel(gx, xr.grad)
    out["grad_norms_ref"] = {n: float(p.grad.norm()) for n, p in zip(names, rparams)}
    out["grad_norms_mine"] = {n: float(g.norm()) for n, g in zip(names, grads)}

    # ---- float64 autograd on the GPU as the truth: what fp32 autograd itself, and this path, differ from it by ---------------
    r64 = copy.deepcopy(ref).double()
    for p in r64.parameters():
        p.grad = None
    x64 = x.double().clone().requires_grad_(True)
    nn.functional.binary_cross_entropy(r64(x64).view(-1), target.double()).backward()
    p64 = [c.weight for c in r64.modules() if isinstance(c, nn.Conv2d)] + \
        [t for bn in r64.modules() if isinstance(bn, nn.BatchNorm2d) for t in (bn.weight, bn.bias)]
    out["vs_float64"] = {"this_path": {n: rel(g, p.grad) for n, g, p in zip(names, grads, p64)},
                         "this_path_grad_x": rel(gx, x64.grad),
                         "torch_fp32_no_tf32": {n: rel(p.grad, q.grad) for n, p, q in zip(names, rparams, p64)},
                         "torch_fp32_no_tf32_grad_x": rel(xr.grad, x64.grad)}

    # ---- what torch's own default (TF32 convolutions) differs by, for scale ----------------------------------------------
    torch.backends.cudnn.allow_tf32 = True
    tf = copy.deepcopy(ref)
    for p in tf.parameters():
        p.grad = None
    xt = x.clone().requires_grad_(True)
    nn.functional.binary_cross_entropy(tf(xt).view(-1), target).backward()
    tparams = [c.weight for c in tf.modules() if isinstance(c, nn.Conv2d)] + \
        [t for bn in tf.modules() if isinstance(bn, nn.BatchNorm2d) for t in (bn.weight, bn.bias)]
    out["torch_tf32_grad_rel_l2"] = {n: rel(t.grad, p.grad) for n, t, p in zip(names, tparams, rparams)}
    out["torch_tf32_grad_x_rel_l2"] = rel(xt.grad, xr.grad)

    # ---- the nn.Module wrapper through autograd --------------------------------------------------------------------------
    mine2 = copy.deepcopy(ref)
    for p in mine2.parameters():
        p.grad = None
    for m, r in zip([m for m in mine2.modules() if isinstance(m, nn.BatchNorm2d)], rbn):
        pass
    w2 = sb.accelerate_discriminator(mine2, max_batch=B, precision=a.precision)
    xm = x.clone().requires_grad_(True)
    po = w2(xm)
    out["wrapper_shape"] = list(po.shape)
    nn.functional.binary_cross_entropy(po.view(-1), target).backward()
    mparams = w2._params()
    out["wrapper_grad_rel_l2"] = {n: rel(m.grad, p.grad) for n, m, p in zip(names, mparams, rparams)}
    out["wrapper_grad_x_rel_l2"] = rel(xm.grad, xr.grad)
    w2.check()

    if not a.no_time:
        def mine_step():
            for p in mparams:
                p.grad = None
            nn.functional.binary_cross_entropy(w2(x).view(-1), target).backward()

        def mine_gstep():
            xg = x.detach().requires_grad_(True)
            nn.functional.binary_cross_entropy(w2(xg, param_grads=False).view(-1), target).backward()

        def ref_step():
            for p in tf.parameters():
                p.grad = None
            nn.functional.binary_cross_entropy(tf(x).view(-1), target).backward()

        def ref_gstep():
            xg = x.detach().requires_grad_(True)
            nn.functional.binary_cross_entropy(tf(xg).view(-1), target).backward()
        out["ms"] = {"b200_d_step": gpu_time(mine_step) * 1e3, "b200_g_step_through_d": gpu_time(mine_gstep) * 1e3,
                     "torch_tf32_d_step": gpu_time(ref_step) * 1e3, "torch_tf32_g_step_through_d": gpu_time(ref_gstep) * 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
