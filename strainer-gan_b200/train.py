"""The discriminator of the TRAINING step on the library's kernels.

``TrainableD64(netD)`` wraps the reference's 64x64 DCGAN ``Discriminator`` ("#strainer gan.py:230-256") without copying
its parameters: optimisers built on ``netD.parameters()`` keep working, ``wrapped(x)`` returns what ``netD(x)`` returns
(``[B, 1, 1, 1]`` probabilities, BatchNorm in training mode with its running statistics updated in place) and
``err.backward()`` fills the same ``.grad`` fields -- but forward and backward run as ONE C call each
(``sg_d64_train_forward`` / ``sg_d64_train_backward``: tcgen05 implicit-GEMM fprop / dgrad / wgrad, fused BatchNorm and
LeakyReLU passes) instead of autograd's ~150 cuDNN / elementwise launches.  The loop it replaces:

    netD.zero_grad(); output = netD(real).view(-1); errD_real = criterion(output, label); errD_real.backward()   # :586-592
    fake = netG(noise); output = netD(fake.detach()).view(-1); errD_fake = ...; errD_fake.backward()             # :595-603
    optimizerD.step(); netG.zero_grad(); output = netD(fake).view(-1); errG = ...; errG.backward()               # :605-615

There is no fallback: without the CUDA library (or on a module that is not the reference Discriminator) it raises."""
import ctypes

import torch
from torch import nn

from . import _lib as L
from . import api as A


class _Workspace:
    """One forward -> backward pair's saved tensors (a caller-owned buffer of the C ABI, borders zeroed once)."""

    def __init__(self, device, lib, capacity: int, precision: int = 0):
        self.capacity = int(capacity)
        self.precision = int(precision)
        self.buf = A._aligned_empty(lib.sg_d64_train_workspace_bytes(self.capacity, self.precision), device)
        L.check(lib.sg_d64_train_workspace_init(L.P(self.buf.data_ptr()), self.capacity, self.precision, A._stream()),
                "sg_d64_train_workspace_init")


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _D64TrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, param_grads, x, *params):
        dev = x.device
        lib = A._lib_for(dev)
        b = x.shape[0]
        ws = owner._take(dev, lib, b)
        prob = torch.empty(b, device=dev, dtype=torch.float32)
        stats = owner._running_stats()
        packed, versions = owner._packed_for(dev, lib, params)
        L.check(lib.sg_d64_train_forward(L.P(x.data_ptr()), b, ws.capacity, ws.precision, L.P(packed.data_ptr()),
                                         _ptr_array(params[5:]), _ptr_array(stats) if stats else None, owner.momentum, owner.eps,
                                         L.P(ws.buf.data_ptr()), L.P(prob.data_ptr()), None, A._stream()),
                "sg_d64_train_forward")
        if stats:
            A._bump_versions(stats)
            owner._count_batch()
        ctx.owner, ctx.ws, ctx.batch, ctx.param_grads = owner, ws, b, param_grads
        ctx.packed, ctx.versions, ctx.weights = packed, versions, params[:5]
        ctx.shapes = [p.shape for p in params]
        ctx.x_shape = x.shape
        return prob.view(b, 1, 1, 1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        ws = ctx.ws
        if ws is None:
            raise RuntimeError("strainer_b200: the saved tensors of this discriminator pass are gone (a second backward through "
                               "the same output, e.g. retain_graph=True, is not supported)")
        if [w._version for w in ctx.weights] != ctx.versions or ctx.owner._packed_versions != ctx.versions:
            raise RuntimeError("strainer_b200: a discriminator weight was modified in place between this forward and its backward "
                               "(autograd raises in the same situation)")
        dev = grad_out.device
        lib = A._lib_for(dev)
        g = grad_out.reshape(-1).to(torch.float32).contiguous()
        need_x = ctx.needs_input_grad[2]
        need_p = ctx.param_grads and any(ctx.needs_input_grad[3:])
        grads = [torch.empty(s, device=dev, dtype=torch.float32) for s in ctx.shapes] if need_p else None
        gx = torch.empty(ctx.x_shape, device=dev, dtype=torch.float32) if need_x else None
        L.check(lib.sg_d64_train_backward(L.P(g.data_ptr()), ctx.batch, ws.capacity, ws.precision, L.P(ctx.packed.data_ptr()),
                                          L.P(ws.buf.data_ptr()),
                                          _ptr_array(grads) if grads else None, L.P(gx.data_ptr()) if need_x else None,
                                          A._stream()), "sg_d64_train_backward")
        ctx.owner._give_back(ws)
        ctx.ws = None
        out = [None, None, gx]
        out += grads if grads else [None] * len(ctx.shapes)
        return tuple(out)


class TrainableD64(nn.Module):
    """``netD`` of the reference with forward + backward on the library's kernels while ``training`` is set; in eval mode
    the wrapped module itself runs.  ``forward(x, param_grads=False)`` skips the parameter gradients of this call (the G
    step: the reference computes and then discards them with the next ``netD.zero_grad()``).
    ``precision``: 'fp16' (default: fp16 operands, fp32 accumulation -- the arithmetic class of torch's own default TF32
    convolutions, gradients within 1-3 % of fp32 autograd like TF32's) or 'fp32' (split operands, three tensor passes:
    gradients within 1e-3 of fp32 autograd)."""

    def __init__(self, discriminator: nn.Module, max_batch: int = 128, precision: str = "fp16"):
        super().__init__()
        if precision not in ("fp16", "fp32"):
            raise ValueError("precision must be 'fp16' or 'fp32'")
        self.precision = 1 if precision == "fp32" else 0
        self.netD = discriminator
        convs, bns = A._d64_modules(discriminator)
        if len({bn.eps for bn in bns}) != 1 or len({bn.momentum for bn in bns}) != 1 or bns[0].momentum is None:
            raise NotImplementedError("strainer_b200: the three BatchNorm2d layers must share eps and a fixed momentum")
        if not all(bn.affine and bn.track_running_stats for bn in bns):
            raise NotImplementedError("strainer_b200: BatchNorm2d must be affine with running statistics (the reference's)")
        self._convs, self._bns = convs, bns
        self.eps, self.momentum = float(bns[0].eps), float(bns[0].momentum)
        self.max_batch = int(max_batch)
        self._free = []
        self._packed = None
        self._packed_versions = None

    # -- workspaces: one per forward whose backward is still to come --------------------------------------------------
    def _take(self, device, lib, batch):
        if batch > self.max_batch:
            self.max_batch = int(batch)
        for i, ws in enumerate(self._free):
            if ws.capacity >= batch and ws.buf.device == device:
                return self._free.pop(i)
        return _Workspace(device, lib, self.max_batch, self.precision)

    def _give_back(self, ws):
        if len(self._free) < 4:
            self._free.append(ws)

    def _packed_for(self, device, lib, params):
        """fp16 operand forms of the conv weights, repacked when a weight's version counter moved (an optimiser step)"""
        versions = [w._version for w in params[:5]]
        if self._packed is None or self._packed.device != device:
            self._packed = A._aligned_empty(lib.sg_d64_train_packed_bytes(self.precision), device)
            self._packed_versions = None
        if versions != self._packed_versions:
            L.check(lib.sg_d64_train_pack(_ptr_array(params[:5]), self.precision, L.P(self._packed.data_ptr()), A._stream()),
                    "sg_d64_train_pack")
            self._packed_versions = versions
        return self._packed, versions

    def _running_stats(self):
        return [t for bn in self._bns for t in (bn.running_mean, bn.running_var)]

    def _count_batch(self):
        with torch.no_grad():
            torch._foreach_add_([bn.num_batches_tracked for bn in self._bns], 1)

    def _params(self):
        return [c.weight for c in self._convs] + [t for bn in self._bns for t in (bn.weight, bn.bias)]

    def score_train(self, x: torch.Tensor, prob: torch.Tensor) -> torch.Tensor:
        """``netD(x)`` in training mode without a graph (the in-batch strain block, "# 상위 10% 제거해서 fake image에
        concate.py:244-245"): batch-statistics BatchNorm, running statistics committed on the device only if every logit is
        finite.  Writes ``prob[B]``; returns the workspace's two status words (device int32[2], sticky)."""
        dev = x.device
        lib = A._lib_for(dev)
        b = x.shape[0]
        ws = self._take(dev, lib, b)
        params = self._params()
        stats = self._running_stats()
        packed, _ = self._packed_for(dev, lib, params)
        L.check(lib.sg_d64_train_forward(L.P(x.data_ptr()), b, ws.capacity, ws.precision, L.P(packed.data_ptr()),
                                         _ptr_array(params[5:]), _ptr_array(stats), self.momentum, self.eps, L.P(ws.buf.data_ptr()),
                                         L.P(prob.data_ptr()), None, A._stream()), "sg_d64_train_forward")
        self._give_back(ws)
        return ws.buf[:8].view(torch.int32)

    def committed(self):
        """module-side effects of one committed training-mode forward: version counters, ``num_batches_tracked``"""
        A._bump_versions(self._running_stats())
        self._count_batch()

    def check(self):
        """Synchronises and raises if a kernel timed out or an fp16 value left its range since the last check."""
        for ws in self._free:
            lib = A._lib_for(ws.buf.device)
            L.check(lib.sg_d64_train_check(L.P(ws.buf.data_ptr()), A._stream()), "sg_d64_train_check")

    def forward(self, x, param_grads: bool = True):
        if not self.training:
            return self.netD(x)
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4 and tuple(x.shape[1:]) == (3, 64, 64)):
            raise RuntimeError("strainer_b200: TrainableD64 takes CUDA tensors [B, 3, 64, 64]; it has no CPU fallback")
        if x.shape[0] < 2:
            raise ValueError("BatchNorm in training mode needs more than one sample")
        params = self._params()
        if any(p.device != x.device or p.dtype != torch.float32 or not p.is_contiguous() for p in params):
            raise RuntimeError("strainer_b200: parameters must be contiguous fp32 tensors on the input's device")
        xc = x.to(torch.float32).contiguous()
        if not torch.is_grad_enabled() or not (xc.requires_grad or any(p.requires_grad for p in params)):
            # no graph: the workspace goes straight back to the pool
            prob = torch.empty(xc.shape[0], device=xc.device, dtype=torch.float32)
            self.score_train(xc, prob)
            self.committed()
            return prob.view(-1, 1, 1, 1)
        return _D64TrainFn.apply(self, bool(param_grads), xc, *params)


def trainer_for(discriminator: nn.Module, max_batch: int = 128, precision: str = "fp16") -> TrainableD64:
    """The module's one TrainableD64 per precision (kept on the module, outside nn.Module's registry): the in-batch strain
    block and the training step share its packed weights and workspaces."""
    if isinstance(discriminator, TrainableD64):
        return discriminator
    key = "_sg_trainer_" + precision
    t = discriminator.__dict__.get(key)
    if t is None:
        t = TrainableD64(discriminator, max_batch, precision)
        object.__setattr__(discriminator, key, t)
    return t


def accelerate_discriminator(discriminator: nn.Module, max_batch: int = 128, precision: str = "fp16") -> TrainableD64:
    """``netD = accelerate_discriminator(netD)`` after the optimiser was built: same parameters, same outputs, the
    training-step forward / backward on tcgen05.  ``precision='fp32'``: split-operand arithmetic (3 tensor passes)."""
    t = trainer_for(discriminator, max_batch, precision)
    t.max_batch = max(t.max_batch, int(max_batch))
    return t
