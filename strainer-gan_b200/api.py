"""Host-side mirror of the reference's straining functions.

Same names, positional parameters, defaults and return types as the functions / inline blocks of
the reference scripts (cited per function), with the compute moved to libstrainer_b200.so:
tcgen05 implicit-GEMM discriminator scoring, radix select, single-pass compaction.  PyTorch is used
only for device memory, streams and torch.distributed.  There is NO CPU fallback: without an
sm_100 GPU or without the built library every entry point raises.
"""
from __future__ import annotations

import ctypes
import math
import os
import time
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L

_c = L  # short alias


# ----------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------
def _dev(device=None) -> torch.device:
    if type(device) is torch.device and device.type == "cuda" and device.index is not None:
        return device      # the device of a tensor that already lives there (the per-batch paths call this twice)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    device = torch.device(device) if device is not None else None
    if device is None or device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("strainer_b200 needs a CUDA (sm_100a) device; it has no CPU fallback "
                           f"(asked for device={device!r})")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class _DevLib:
    """The C ABI bound to ONE device: every call runs with that device current (the library launches on, and keeps its
    state per, the current CUDA device), whatever torch's current device is; the caller's device is restored."""

    def __init__(self, lib, index: int):
        self._lib = lib
        self._index = index

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        index = self._index

        def call(*args):
            if torch.cuda.current_device() == index:
                return fn(*args)
            with torch.cuda.device(index):
                return fn(*args)
        setattr(self, name, call)
        return call


_DEVLIBS: dict = {}
_ACTIVE = [None]    # device index of the library calls being assembled (set by _lib_for, read by _stream)


def _dev_of(x) -> torch.device:
    """The device an input already lives on (CUDA tensors / U8Images), else torch's current CUDA device."""
    px = getattr(x, "pixels", x)
    if isinstance(px, torch.Tensor) and px.is_cuda:
        return px.device
    return _dev()


def _lib_for(device: torch.device):
    idx = device.index
    d = _DEVLIBS.get(idx)
    if d is None:
        d = _DevLib(L.init(idx), idx)
        _DEVLIBS[idx] = d
    _ACTIVE[0] = idx
    return d


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """torch's current stream ON THE DEVICE of the call being assembled (not of torch's current device).  The raw query
    costs < 1 us; building a torch.cuda.Stream object per C call cost 6 us of the in-batch block's ~60 us of host time."""
    if _RAW_STREAM is not None:
        return L.P(_RAW_STREAM(_ACTIVE[0]))
    return L.P(torch.cuda.current_stream(_ACTIVE[0]).cuda_stream)


def _p(t):
    return L.P(t.data_ptr()) if t is not None else L.P(0)


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    return t.to(device=device, dtype=torch.float32).contiguous()


def _aligned_empty(nbytes: int, device, align: int = 1024) -> torch.Tensor:
    """uint8 device buffer whose address is a multiple of ``align``: the caching allocator only guarantees 512 bytes (small
    blocks), the C ABI asks for 1024 (TMA / swizzle atoms) on its workspaces."""
    n = (int(nbytes) + align - 1) // align * align
    raw = torch.empty(n + align, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % align
    return raw[off:off + n]


class _Scratch:
    """Per-device cache of workspaces (caller-owned buffers of the C ABI)."""
    _cache: dict = {}

    @classmethod
    def get(cls, device, key, nbytes, dtype=torch.uint8):
        k = (device.index, key)
        t = cls._cache.get(k)
        n = (int(nbytes) + 1023) // 1024 * 1024
        if t is None or t.numel() < n:
            t = _aligned_empty(n, device)
            cls._cache[k] = t
        return t


# ----------------------------------------------------------------------------------------------
# order statistics / thresholds on device
# ----------------------------------------------------------------------------------------------
def _np_percentile_plan(n: int, q):
    """(k_prev, k_next, gamma) exactly as numpy 2.x ``_quantile`` (method 'linear') derives them for
    a float32 array: q/100 in the array dtype for python scalars, virtual index (n-1)*q."""
    qq = np.true_divide(q, np.float32(100))
    v = np.asanyarray((n - 1) * qq)
    if not (0.0 <= float(qq) <= 1.0):
        raise ValueError("Percentiles must be in the range [0, 100]")
    prev = np.floor(v)
    if v >= n - 1:
        return n - 1, n - 1, v.dtype.type(v - (n - 1)), v.dtype
    if v < 0:
        return 0, 0, v.dtype.type(v), v.dtype
    return int(prev), int(prev) + 1, v.dtype.type(v - prev), v.dtype


def _torch_quantile_plan(n: int, q: float):
    """(below, above, weight) as ATen ``quantile_compute``: rank = q*(n-1) in fp32."""
    if not 0.0 <= q <= 1.0:
        raise RuntimeError("quantile() q values must be in the range [0, 1]")
    rank = np.float32(q) * np.float32(n - 1)
    below = np.floor(rank)
    return int(below), int(np.ceil(rank)), np.float32(rank - below)


class _SelectOps:
    """The five radix-select phases of the C ABI on one device (tests substitute a numpy double to
    exercise the multi-rank protocol on CPU/gloo)."""

    def __init__(self, device):
        self.device = device

    @property
    def lib(self):
        return _lib_for(self.device)

    def begin(self, ws, k):
        L.check(self.lib.sg_select_begin(_p(ws), k, _stream()), "sg_select_begin")

    def hist(self, values, ws, p):
        L.check(self.lib.sg_select_hist(_p(values), values.numel(), _p(ws), p, _stream()), "sg_select_hist")

    def step(self, ws, p):
        L.check(self.lib.sg_select_step(_p(ws), p, _stream()), "sg_select_step")

    def finish(self, ws, out2):
        L.check(self.lib.sg_select_finish(_p(ws), _p(out2), _stream()), "sg_select_finish")


class PeerComm:
    """NVLink peer-memory communicator of one process group (one process per GPU, one node): every rank owns a small
    buffer that all ranks map through CUDA IPC; ``sg_select_step_peer`` all-reduces the radix histograms through it inside
    the select's own kernels (one-shot tagged 64-bit stores), so a sharded strain needs no collective launch at all.
    Built once per group (``PeerComm.for_group``); every rank must build it and use it in the same order."""
    _cache: dict = {}

    def __init__(self, group, device):
        import torch.distributed as dist
        self.group = group
        self.device = _dev(device)
        self.rank = dist.get_rank(group)
        self.nranks = dist.get_world_size(group)
        lib = _lib_for(self.device)
        buf = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        L.check(lib.sg_peer_alloc(self.nranks, ctypes.byref(buf), handle), "sg_peer_alloc")
        self.buffer = buf.value
        handles = [None] * self.nranks
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        ptrs = []
        self._opened = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self.buffer)
                continue
            p = ctypes.c_void_p()
            L.check(lib.sg_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)), "sg_peer_open")
            ptrs.append(p.value)
            self._opened.append(p.value)
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.seq = 0
        dist.barrier(group=group)          # every buffer is zeroed and mapped before the first tagged store

    def next_seq(self) -> int:
        self.seq = self.seq % 0xFFFFFFF0 + 1       # never 0: a zeroed slot holds no valid word
        return self.seq

    @classmethod
    def for_group(cls, group, device):
        """The group's communicator, or None when it cannot exist (non-NCCL backend, a single rank, SG_NO_PEER set, or
        the GPUs cannot map each other's memory): callers then all-reduce through torch.distributed."""
        import torch.distributed as dist
        if group is None or os.environ.get("SG_NO_PEER") or dist.get_world_size(group) < 2 or dist.get_backend(group) != "nccl":
            return None
        key = (id(group), _dev(device).index)
        if key not in cls._cache:
            ok = torch.ones(1, dtype=torch.int32, device=device)
            comm = None
            try:
                comm = cls(group, device)
            except Exception as e:      # every rank must agree on the path: vote below
                ok.zero_()
                import warnings
                warnings.warn(f"strainer_b200: NVLink peer buffers unavailable ({e}); using NCCL all-reduces")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            cls._cache[key] = comm if int(ok.item()) == 1 else None
        return cls._cache[key]


def order_stats(values: torch.Tensor, k: int, group=None, ops=None, comm=None) -> torch.Tensor:
    """Device tensor [x_(k), x_(k+1)] of a 1-D fp32 CUDA tensor (radix select, no sort).
    With ``group`` (torch.distributed), ``values`` is this rank's shard and k the GLOBAL rank: the
    integer digit histograms (+ NaN count) are all-reduced with SUM after each of the 4 passes and the
    smallest key above the selected bucket with MIN, so every rank derives bit-identical statistics.  With ``comm``
    (a ``PeerComm``) those reductions run inside the select's own step kernels over NVLink peer memory -- no collective
    launch; without it through torch.distributed (NCCL / gloo)."""
    device = values.device
    n = values.numel()
    out2 = torch.empty(2, dtype=torch.float32, device=device)
    if comm is not None:
        lib = _lib_for(device)
        ws = _Scratch.get(device, "select_peer", L.SG_SELECT_WS_WORDS * 4).view(torch.int32)
        L.check(lib.sg_select_begin(_p(ws), k, _stream()), "sg_select_begin")
        for p in range(L.SG_SELECT_NUM_PASSES):
            L.check(lib.sg_select_hist(_p(values), n, _p(ws), p, _stream()), "sg_select_hist")
            L.check(lib.sg_select_step_peer(_p(ws), p, _p(comm.table), comm.rank, comm.nranks, comm.next_seq(), _stream()),
                    "sg_select_step_peer")
        L.check(lib.sg_select_finish(_p(ws), _p(out2), _stream()), "sg_select_finish")
        return out2
    if group is None and ops is None:
        # single device: one streaming read for large n (sampled pivots), fused radix passes otherwise
        lib = _lib_for(device)
        nbytes = int(lib.sg_select_workspace_bytes(n))
        ws = _Scratch.get(device, "select", nbytes)
        L.check(lib.sg_select_kth(_p(values), n, k, _p(ws), nbytes, _p(out2), _stream()), "sg_select_kth")
        return out2
    ws = torch.empty(L.SG_SELECT_WS_WORDS, dtype=torch.int32, device=device)
    import torch.distributed as dist
    ops = ops or _SelectOps(device)
    ops.begin(ws, k)
    for p in range(L.SG_SELECT_NUM_PASSES):
        ops.hist(values, ws, p)
        if group is not None:
            dist.all_reduce(ws[:L.SG_SELECT_WS_NANCOUNT + 1], op=dist.ReduceOp.SUM, group=group)
            if p == L.SG_SELECT_NUM_PASSES - 1:
                m = ws[L.SG_SELECT_WS_MINABOVE:L.SG_SELECT_WS_MINABOVE + 1]
                m ^= -2147483648  # uint32 key order -> int32 order for the MIN reduction
                dist.all_reduce(m, op=dist.ReduceOp.MIN, group=group)
                m ^= -2147483648
        ops.step(ws, p)
    ops.finish(ws, out2)
    return out2


def _lerp_dev(stats2: torch.Tensor, weight, kind: int) -> torch.Tensor:
    thr = torch.empty(1, dtype=torch.float32, device=stats2.device)
    lib = _lib_for(stats2.device)
    L.check(lib.sg_lerp_threshold(_p(stats2), float(weight), kind, _p(thr), _stream()), "sg_lerp_threshold")
    return thr


def percentile_device(values: torch.Tensor, q, group=None, n_global=None, ops=None, comm=None) -> torch.Tensor:
    """``np.percentile(values_f32, q)`` evaluated on the GPU; returns a 1-element fp32 device tensor
    holding bit-for-bit numpy's result for python-scalar q (numpy evaluates q and the lerp in fp32)."""
    if group is not None and n_global is None:
        n_global = _group_total(values.numel(), group, values.device)   # never rank the global vector by a shard length
    n = values.numel() if n_global is None else int(n_global)
    k0, k1, gamma, gdt = _np_percentile_plan(n, q)
    stats2 = order_stats(values, k0, group, ops, comm)
    if gdt != np.float32:
        # np.float64 q: numpy interpolates in float64 -> host finish on the two order statistics
        a, b = stats2.cpu().numpy().astype(np.float64)
        d = b - a
        r = a + d * gamma if gamma < 0.5 else b - d * (1 - gamma)
        return torch.tensor([r], dtype=torch.float64)
    if k1 == k0:
        stats2 = torch.stack([stats2[0], stats2[0]])
    if ops is not None and hasattr(ops, "lerp"):
        return ops.lerp(stats2, gamma, L.SG_LERP_NUMPY)
    return _lerp_dev(stats2, gamma, L.SG_LERP_NUMPY)


def quantile_device(values: torch.Tensor, q: float) -> torch.Tensor:
    """``torch.quantile(values, q)`` for a 1-D fp32 CUDA tensor; 1-element device tensor."""
    n = values.numel()
    k0, k1, w = _torch_quantile_plan(n, float(q))
    device = values.device
    lib = _lib_for(device)
    if n <= 2048:
        stats2 = torch.empty(2, dtype=torch.float32, device=device)
        L.check(lib.sg_segment_order_stats(_p(values), 1, n, k0, k1, _p(stats2), _stream()), "sg_segment_order_stats")
    else:
        stats2 = order_stats(values, k0)
        if k1 == k0:
            stats2 = torch.stack([stats2[0], stats2[0]])
    return _lerp_dev(stats2, w, L.SG_LERP_TORCH)


def _f32_threshold(thr, cmp: int) -> np.float32:
    """Directed rounding of a (possibly float64) threshold so that the fp32 device comparison
    ``v CMP thr32`` equals numpy's promoted comparison ``v_f32 CMP thr`` for every fp32 v."""
    t = np.float64(thr)
    f = np.float32(t)
    if np.isnan(t) or np.float64(f) == t:
        return f
    lo = f if np.float64(f) < t else np.nextafter(f, np.float32(-np.inf))
    hi = f if np.float64(f) > t else np.nextafter(f, np.float32(np.inf))
    return np.float32(hi) if (cmp & 3) in (L.SG_LT, L.SG_GE) else np.float32(lo)


def compact_indices(values: torch.Tensor, thr, cmp: int = L.SG_LT, index_base: int = 0, want_mask: bool = False):
    """Ascending int64 indices i (+index_base) with values[i] CMP thr, as ``np.where(...)[0]``.
    ``thr`` is a 1-element fp32 device tensor or a host scalar.  Returns (idx_buffer, count_dev, mask)
    -- all on device, no synchronisation; idx_buffer[:count] is valid."""
    device = values.device
    lib = _lib_for(device)
    n = values.numel()
    if not isinstance(thr, torch.Tensor):
        thr = torch.tensor([_f32_threshold(thr, cmp)], dtype=torch.float32).to(device)
    elif thr.dtype != torch.float32 or thr.device != device:
        thr = torch.tensor([_f32_threshold(float(thr.reshape(-1)[0]), cmp)], dtype=torch.float32).to(device)
    idx = torch.empty(max(n, 1), dtype=torch.int64, device=device)
    count = torch.empty(1, dtype=torch.int64, device=device)
    mask = torch.empty(max(n, 1), dtype=torch.uint8, device=device) if want_mask else None
    ws = _Scratch.get(device, "compact", lib.sg_compact_workspace_bytes(n))
    L.check(lib.sg_compact_indices(_p(values), n, _p(thr), cmp, index_base, _p(idx), _p(count), _p(mask), _p(ws),
                                   _stream()), "sg_compact_indices")
    return idx, count, (mask[:n] if want_mask else None)


def partition_rows(rows: torch.Tensor, mask: torch.Tensor, kept_out=None, dropped_out=None):
    """Stable two-way row partition ``(rows[mask], rows[~mask])`` in ONE pass.  ``kept_out`` /
    ``dropped_out`` may be pre-sized destination views (e.g. the tail of the fake batch).
    Returns (kept_buffer, dropped_buffer, counts_dev[2])."""
    device = rows.device
    lib = _lib_for(device)
    n = rows.shape[0]
    rows = rows.contiguous()
    row_bytes = rows[0].numel() * rows.element_size() if n else 16
    m8 = mask.to(torch.uint8).contiguous() if mask.dtype != torch.uint8 else mask.contiguous()
    kept = kept_out if kept_out is not None else torch.empty_like(rows)
    dropped = dropped_out if dropped_out is not None else torch.empty_like(rows)
    counts = torch.empty(2, dtype=torch.int64, device=device)
    ws = _Scratch.get(device, "compact", lib.sg_compact_workspace_bytes(n))
    L.check(lib.sg_compact_rows(_p(rows), n, row_bytes, _p(m8), _p(kept), _p(dropped), _p(counts), _p(ws), _stream()),
            "sg_compact_rows")
    return kept, dropped, counts


# ----------------------------------------------------------------------------------------------
# D64 scoring
# ----------------------------------------------------------------------------------------------
_MODES = {"bf16": L.SG_CONV_BF16, "fp32": L.SG_CONV_BF16X3, "bf16x3": L.SG_CONV_BF16X3, "fp16": L.SG_CONV_FP16,
          "auto": L.SG_CONV_FP16}
_FP16_OVERFLOW = 0x46503136     # status word 1 of the head kernel (csrc/d64.cu: kFp16OverflowMagic)
_FALLBACK_CHUNK = 1024          # samples per launch group of the fp32-parity re-score of an overflowed chunk


def _d64_modules(discriminator: nn.Module):
    """The five Conv2d and three BatchNorm2d of the reference Discriminator ("#strainer gan.py:230-256");
    anything else is rejected (no fallback)."""
    mod = discriminator.module if isinstance(discriminator, nn.DataParallel) else discriminator
    convs = [m for m in mod.modules() if isinstance(m, nn.Conv2d)]
    bns = [m for m in mod.modules() if isinstance(m, nn.BatchNorm2d)]
    shapes = [tuple(c.weight.shape) for c in convs]
    want = [(64, 3, 4, 4), (128, 64, 4, 4), (256, 128, 4, 4), (512, 256, 4, 4), (1, 512, 4, 4)]
    ok = shapes == want and len(bns) == 3 and all(c.bias is None for c in convs)
    ok = ok and [c.stride for c in convs] == [(2, 2)] * 4 + [(1, 1)] and [c.padding for c in convs] == [(1, 1)] * 4 + [(0, 0)]
    if not ok:
        raise NotImplementedError(
            "strainer_b200 scores the reference's 64x64 DCGAN Discriminator (nc=3, ndf=64) only; got conv shapes "
            f"{shapes} with {len(bns)} BatchNorm2d layers")
    return convs, bns


def _raise_status(st0: int):
    raise RuntimeError(f"strainer_b200: the conv pipeline timed out waiting on an mbarrier (role code {st0}: 1x producer, "
                       "2x mma, 3x accumulator, 4x epilogue); the losses of this call are invalid")


class _PackTuner:
    """Picks the host-packed share of a pinned fp32 source per (device, host thread count) by timing the first few
    ``score`` calls of at least four chunks: all rows packed first, then 0.1 less per call while that is more than 2 %
    faster (on the 16-core B200 box: 1.0 with 16 threads, 0.9 with 12, 0.8 with 8; tools/host_pack_sweep.py), then the
    best share is kept.  A share below 0.3 ends at the plain fp32 copy."""
    _by_key: dict = {}

    def __init__(self):
        self.share = 1.0
        self.best_rate = 0.0
        self.best_share = 1.0
        self.locked = False

    @classmethod
    def get(cls, device) -> "_PackTuner":
        key = (device.index, _host_threads())
        t = cls._by_key.get(key)
        if t is None:
            t = cls._by_key[key] = cls()
        return t

    def report(self, share: float, samples_per_s: float):
        if self.locked or share != self.share:
            return
        if samples_per_s > self.best_rate * 1.02:
            self.best_rate, self.best_share = samples_per_s, share
            nxt = round(share - 0.1, 2)
            if share <= 0.0:
                self.locked = True
            else:
                self.share = nxt if nxt >= 0.3 else 0.0
        else:
            self.share = self.best_share
            self.locked = True


def _host_threads() -> int:
    """Host threads of this process for the fp16 staging conversion: its CPU affinity, shared between the ranks of a
    torchrun launch on this node."""
    cpus = int(L.load().sg_host_threads())
    local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(1, cpus // max(local, 1))


_PACK_EXECUTOR = None


def _pack_executor():
    """One background thread that drives the host conversion of the NEXT chunk while the caller queues this one."""
    global _PACK_EXECUTOR
    if _PACK_EXECUTOR is None:
        from concurrent.futures import ThreadPoolExecutor
        _PACK_EXECUTOR = ThreadPoolExecutor(max_workers=1, thread_name_prefix="sg-host-pack")
    return _PACK_EXECUTOR


def _host_to_f16(src: torch.Tensor, dst: torch.Tensor, isa: int = 0):
    """dst (host fp16, contiguous) = src (host fp32, contiguous) rounded to nearest even, on the library's host threads."""
    if src.is_cuda or dst.is_cuda or src.dtype != torch.float32 or dst.dtype != torch.float16:
        raise ValueError("_host_to_f16: host fp32 source and host fp16 destination")
    if not (src.is_contiguous() and dst.is_contiguous()) or src.numel() != dst.numel():
        raise ValueError("_host_to_f16: contiguous tensors of equal size")
    L.check(L.load().sg_host_f32_to_f16(src.data_ptr(), src.numel(), dst.data_ptr(), _host_threads(), isa),
            "sg_host_f32_to_f16")


class D64Scorer:
    """Packs a reference ``Discriminator``'s weights for the tcgen05 kernels and scores batches:
    eval-mode BN folded into the conv epilogues, sigmoid and BCE-vs-1 fused into the head.

    mode 'auto' (the default of every entry point): fp16 operands and activations with fp32 accumulation -- one tensor
    pass, losses within 1e-3 of fp32 (measured 2e-4) -- and, for exactly the chunks whose head reported a non-finite
    logit (an activation beyond fp16's 65504), a second scoring of that chunk in the fp32-parity mode; nothing raises.
    mode 'fp16': the same kernels without the recovery (an overflow raises).  mode 'fp32' / 'bf16x3': fp32-parity
    arithmetic (bf16 hi/lo split, 3 tensor-core passes, ~1e-5 relative on the losses).  mode 'bf16': bf16 operands /
    fp32 accumulate (2e-2 bar, stated separately)."""

    def __init__(self, discriminator: nn.Module, device=None, mode: str = "auto", max_batch: int = 4096):
        self.device = _dev(device)
        if mode not in _MODES:
            raise ValueError(f"mode must be one of {sorted(_MODES)}")
        self.mode_name = mode
        self.mode = _MODES[mode]
        self.max_batch = int(max_batch)
        lib = self.lib
        self.packed = _aligned_empty(lib.sg_d64_packed_bytes(self.mode), self.device)
        self.ws = _aligned_empty(lib.sg_d64_workspace_bytes(self.max_batch, self.mode), self.device)
        self.ws[:1024].zero_()      # sticky status words (pipeline time-out, fp16 overflow) of calls without a status slice
        self._sig = None
        self._fallback = None       # 'auto': fp32-parity scorer, built when the first chunk overflows
        self._stage = {}            # cached staging buffers of the host-streaming path
        self.fallback_chunks = 0    # chunks (or batches) re-scored since construction
        self.host_pack = "auto"     # fp32 HOST sources: share of each chunk rounded to fp16 on the host before its PCIe copy
        self.h2d_bytes = 0          # bytes this scorer has copied host -> device (bench.py reports the difference per step)
        self.last_pack_fraction = 0.0
        self.repack(discriminator)

    @property
    def lib(self):
        return _lib_for(self.device)

    @staticmethod
    def _signature(convs, bns):
        ts = [c.weight for c in convs]
        for b in bns:
            ts += [b.weight, b.bias, b.running_mean, b.running_var]
        return tuple((t.data_ptr(), t._version) for t in ts)

    def repack(self, discriminator: nn.Module, force: bool = False):
        # fast path: same module object, same parameter / buffer tensors, same versions -> nothing to do
        cached = getattr(self, "_mods", None)
        if cached is not None and cached[0]() is discriminator and not force:
            convs, bns = cached[1], cached[2]
        else:
            convs, bns = _d64_modules(discriminator)
            self._mods = (weakref.ref(discriminator), convs, bns)
        sig = self._signature(convs, bns)
        if sig == self._sig and not force:
            return
        ptrs = tuple(p_ for p_, _ in sig)
        if ptrs == getattr(self, "_ptrs", None) and getattr(self, "_resident", False):
            # the usual training-loop case: the same device tensors, updated in place by the optimiser -> the cached
            # ctypes argument list is still valid, only the pack kernel has to run again
            cargs, eps_v, args = self._cargs, self._eps, self._keep
        else:
            with torch.no_grad():
                srcs = [c.weight for c in convs]
                for b in bns:
                    srcs += [b.weight, b.bias, b.running_mean, b.running_var]
                args = [_f32c(t.detach(), self.device) for t in srcs]
            eps = {float(b.eps) for b in bns}
            if len(eps) != 1:
                raise NotImplementedError("BatchNorm layers with different eps")
            eps_v = eps.pop()
            cargs = [_p(a) for a in args]
            # cacheable only if no conversion copy was made (the arguments alias the module's own storage)
            self._resident = all(a.data_ptr() == t.data_ptr() for a, t in zip(args, srcs))
            self._ptrs, self._cargs, self._eps = ptrs, cargs, eps_v
        L.check(self.lib.sg_d64_pack(*cargs, eps_v, self.mode, _p(self.packed), _stream()), "sg_d64_pack")
        self._keep = args  # stream-ordered: keep alive until the pack kernels ran
        self._sig = sig

    def _module(self):
        m = self._mods[0]()
        if m is None:
            raise RuntimeError("the discriminator this scorer was built for no longer exists")
        return m

    def fallback(self) -> "D64Scorer":
        """The fp32-parity scorer that re-scores what overflowed fp16 in 'auto' mode (built on first use)."""
        if self._fallback is None:
            self._fallback = D64Scorer(self._module(), self.device, "fp32", max_batch=min(self.max_batch, _FALLBACK_CHUNK))
        else:
            self._fallback.repack(self._module())
        return self._fallback

    def score_into(self, x: torch.Tensor, logit=None, prob=None, loss=None, status=None):
        """x: fp32 CUDA [b,3,64,64], b <= max_batch; writes into the given device slices.  ``status``: a zeroed device
        int32[2] that receives this launch group's sticky status words (None: the workspace's own, see ``check``)."""
        b = x.shape[0]
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        L.check(self.lib.sg_d64_score_status(_p(x), b, _p(self.packed), _p(self.ws), self.mode, _p(logit), _p(prob), _p(loss),
                                             _p(status), _stream()), "sg_d64_score_status")

    def score_train_into(self, discriminator: nn.Module, x: torch.Tensor, logit=None, prob=None, loss=None, status=None):
        """``netD(x)`` for a netD in TRAIN mode (under no_grad): BatchNorm uses the statistics of this batch
        and its running_mean / running_var / num_batches_tracked are updated exactly as
        nn.BatchNorm2d does.  In the fp16-operand modes the running statistics are committed on
        the device only if no activation overflowed (the caller then scores the batch again in another mode)."""
        b = x.shape[0]
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        self.repack(discriminator)
        bns = self._mods[2]          # cached by repack()
        stats, back = [], []
        for bn in bns:
            if not bn.track_running_stats or bn.running_mean is None:
                stats += [None, None]
                continue
            for t in (bn.running_mean, bn.running_var):
                if t.is_cuda and t.device == self.device and t.dtype == torch.float32 and t.is_contiguous():
                    stats.append(t)
                else:
                    d = t.detach().to(self.device, torch.float32).contiguous()
                    stats.append(d)
                    back.append((t, d))
        moms = {bn.momentum for bn in bns}
        if len(moms) != 1 or None in moms:
            raise NotImplementedError("BatchNorm layers with different / cumulative momentum")
        L.check(self.lib.sg_d64_score_train_status(_p(x), b, _p(self.packed), _p(self.ws), self.mode, *[_p(t) for t in stats],
                                                   float(moms.pop()), float(bns[0].eps), _p(logit), _p(prob), _p(loss),
                                                   _p(status), _stream()), "sg_d64_score_train_status")
        self._train_back = back
        self._sig = None  # running stats changed under the packed eval-mode fold

    def commit_train_side_effects(self):
        """The module-side effects of a train-mode forward that are not written by the kernels: copies back running
        statistics that do not live on this device, bumps ``num_batches_tracked`` and the version counters of the
        running statistics (so that every cached fold of this module is rebuilt).  Called once per scored batch."""
        bns = self._mods[2]
        with torch.no_grad():
            for t, d in getattr(self, "_train_back", ()):
                t.copy_(d)
            self._train_back = []
            nbt = [bn.num_batches_tracked for bn in bns if bn.track_running_stats and bn.num_batches_tracked is not None]
            if nbt:
                torch._foreach_add_(nbt, 1)      # one launch for the three counters
        _bump_versions([t for bn in bns for t in (bn.running_mean, bn.running_var) if t is not None])

    def run_layer(self, x, layer: int, logit=None, prob=None, loss=None):
        """One stage (1..5) of score_into on this scorer's workspace (benchmark / tests)."""
        L.check(self.lib.sg_d64_run_layer(_p(x), x.shape[0], _p(self.packed), _p(self.ws), self.mode, layer, _p(logit),
                                          _p(prob), _p(loss), _stream()), "sg_d64_run_layer")

    def check(self):
        """Synchronises and raises if a launch WITHOUT a status slice recorded a pipeline time-out or an fp16 overflow
        since the last check (the words are sticky; this clears them)."""
        L.check(self.lib.sg_d64_check(_p(self.ws), _stream()), "sg_d64_check")

    def read_activation(self, batch: int, layer: int) -> torch.Tensor:
        c, s = {1: (64, 32), 2: (128, 16), 3: (256, 8), 4: (512, 4)}[layer]
        out = torch.empty(batch, c, s, s, dtype=torch.float32, device=self.device)
        L.check(self.lib.sg_d64_read_activation(_p(self.ws), batch, self.mode, layer, _p(out), _stream()),
                "sg_d64_read_activation")
        return out

    # -- dataset-scale scoring ---------------------------------------------------------------------------------------
    def _staging(self, key, shape, dtype, pinned=False):
        t = self._stage.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype).pin_memory() if pinned else torch.empty(shape, dtype=dtype, device=self.device)
            self._stage[key] = t
        return t

    def _host_pack_fraction(self, src: torch.Tensor, pinned_src: bool) -> float:
        """Share of each chunk of a fp32 HOST source that is rounded to fp16 on the host before its PCIe copy
        (``host_pack``: 'auto', a float in [0, 1], or False).  Only where conv1 rounds the input to fp16 anyway (modes
        'auto' / 'fp16') and only for inputs of at least 4 096 images.  'auto': a pageable source is packed entirely
        (the conversion reads it in place, where the raw copy would need a staging memcpy first); for a pinned source
        the share comes from ``_PackTuner`` (the host threads and the PCIe copies compete for host memory bandwidth, so
        the best share is found by timing whole calls, not from the two rates in isolation)."""
        hp = self.host_pack
        if hp is False or hp is None or self.mode_name not in ("auto", "fp16") or src.dtype != torch.float32:
            return 0.0
        if not isinstance(hp, str):
            return float(min(max(float(hp), 0.0), 1.0))
        if src.shape[0] < 4096:
            return 0.0
        if not pinned_src:
            return 1.0
        return _PackTuner.get(self.device).share

    def score(self, images, want=("loss",)):
        """Scores N images: a fp32 tensor [N,3,64,64] or a ``U8Images`` (uint8 pixels + Normalize, converted on the
        device), CUDA resident or on the host (streamed through double buffers with the H2D copies overlapped with
        compute; pinned source tensors are copied from directly).  Returns a dict of fp32 device tensors [N].
        Ends with ONE host read of the per-chunk status words: a pipeline time-out raises in every mode, an fp16
        overflow raises in mode 'fp16' and re-scores the affected chunks in fp32-parity arithmetic in mode 'auto'."""
        n = images.shape[0]
        outs = {k: torch.empty(n, dtype=torch.float32, device=self.device) for k in want}
        u8 = images if isinstance(images, U8Images) else None
        if tuple(images.shape[1:]) != (3, 64, 64) or (u8 is None and images.dtype != torch.float32):
            raise ValueError("images must be float32 [N,3,64,64] or U8Images of 3x64x64 pixels")
        cb = self.max_batch
        nchunks = (n + cb - 1) // cb
        if nchunks == 0:
            return outs
        lib = self.lib            # also makes this scorer's device the one _stream() refers to
        status = torch.zeros((nchunks, 2), dtype=torch.int32, device=self.device)

        def sl(name, i, b):
            return outs[name][i:i + b] if name in outs else None

        def score_chunk(x, ci, i, b):
            self.score_into(x, sl("logit", i, b), sl("prob", i, b), sl("loss", i, b), status[ci])

        src = (u8.pixels if u8 is not None else images).contiguous()
        tuned = None
        if src.is_cuda:
            f32 = self._staging("f32_0", (cb, 3, 64, 64), torch.float32) if u8 is not None else None
            for ci in range(nchunks):
                i = ci * cb
                b = min(cb, n - i)
                score_chunk(u8.normalize_into(src[i:i + b], f32[:b]) if u8 is not None else src[i:i + b], ci, i, b)
        else:
            # host path: 2 device buffers (+ 2 pinned staging buffers for a pageable source), copy stream ahead of compute.
            # The buffers are cached on the scorer and only ever touched in this order: the copy stream first waits for
            # everything already queued on the main stream (an earlier call's kernels may still be reading them).
            pinned_src = src.is_pinned()
            main = torch.cuda.current_stream(self.device)
            copy_stream = _copy_stream(self.device)
            copy_stream.wait_stream(main)
            row = tuple(src.shape[1:])
            dev = [self._staging(f"dev_{k}", (cb,) + row, src.dtype) for k in range(2)]
            f32 = [self._staging(f"f32_{k}", (cb, 3, 64, 64), torch.float32) for k in range(2)] if u8 is not None else None
            # host-packed copy (csrc/host_pack.cpp): the last `pk` rows of every chunk are rounded to fp16 by the host
            # threads -- the rounding conv1 applies anyway -- while the first rows cross PCIe as fp32; the share balances
            # the two (measured once per process and device).  Bit-identical scores, about half the PCIe bytes.
            frac = self._host_pack_fraction(src, pinned_src) if u8 is None else 0.0
            self.last_pack_fraction = frac
            if u8 is None and self.host_pack == "auto" and pinned_src and nchunks >= 4 and \
                    self.mode_name in ("auto", "fp16") and src.dtype == torch.float32:
                tuned = (_PackTuner.get(self.device), frac, time.perf_counter())
            pack_rows = (lambda b_: b_ if frac >= 1.0 else (int(b_ * frac) // 16) * 16) if frac > 0.0 else (lambda b_: 0)
            pin16 = dev16 = None
            if frac > 0.0:
                if "pin16_0" not in self._stage:
                    tuned = None               # this call pins its staging buffers (~0.1 s): not a timing to tune on
                # three pinned fp16 buffers: the conversion of chunk c + 1 runs on the host threads (behind a one-thread
                # executor, the C call releases the GIL) while this thread queues the copies and kernels of chunk c
                pin16 = [self._staging(f"pin16_{k}", (cb,) + row, torch.float16, pinned=True) for k in range(3)]
                dev16 = [self._staging(f"dev16_{k}", (cb,) + row, torch.float16) for k in range(2)]
            need_pin = not pinned_src and frac < 1.0
            pin = [self._staging(f"pin_{k}", (cb,) + row, src.dtype, pinned=True) for k in range(2)] if need_pin else None
            copied = [torch.cuda.Event() for _ in range(2)]
            copied16 = [torch.cuda.Event() for _ in range(3)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            row_elems = int(np.prod(row))

            def split(ci):
                i = ci * cb
                b = min(cb, n - i)
                pk = pack_rows(b)
                return i, b, b - pk, pk

            def start_pack(ci):
                i, b, raw, pk = split(ci)
                if not pk:
                    return None
                if ci >= 3:
                    copied16[ci % 3].synchronize()     # the H2D copy that last read this pinned buffer has finished
                return _pack_executor().submit(_host_to_f16, src[i + raw:i + b], pin16[ci % 3][:pk])

            pending = start_pack(0)
            for ci in range(nchunks):
                s_ = ci & 1
                i, b, raw, pk = split(ci)
                if pending is not None:
                    pending.result()
                pending = start_pack(ci + 1) if ci + 1 < nchunks else None
                if ci >= 2 and need_pin:
                    copied[s_].synchronize()  # the H2D copy that last read pin[s_] has finished
                with torch.cuda.stream(copy_stream):
                    if ci >= 2:
                        copy_stream.wait_event(consumed[s_])
                    if pk:
                        dev16[s_][:pk].copy_(pin16[ci % 3][:pk], non_blocking=True)
                        copied16[ci % 3].record(copy_stream)
                        self.h2d_bytes += pk * row_elems * 2
                    if raw:
                        part = src[i:i + raw]
                        if not pinned_src:
                            pin[s_][:raw].copy_(part)
                            part = pin[s_][:raw]
                        dev[s_][:raw].copy_(part, non_blocking=True)
                        self.h2d_bytes += raw * row_elems * src.element_size()
                    copied[s_].record(copy_stream)
                main.wait_event(copied[s_])
                if u8 is not None:
                    # the uint8 staging buffer is free again as soon as the conversion has run
                    x = u8.normalize_into(dev[s_][:b], f32[s_][:b])
                    consumed[s_].record(main)
                    score_chunk(x, ci, i, b)
                else:
                    if pk:
                        L.check(lib.sg_f16_expand(_p(dev16[s_]), pk * row_elems, _p(dev[s_][raw:]), _stream()), "sg_f16_expand")
                    score_chunk(dev[s_][:b], ci, i, b)
                    consumed[s_].record(main)
            if need_pin or frac > 0.0:
                copy_stream.synchronize()   # the cached pinned buffers may be reused by the next call
        del lib
        self._resolve(status, images, outs, cb)
        if tuned is not None:
            tuned[0].report(tuned[1], n / max(time.perf_counter() - tuned[2], 1e-9))   # _resolve has synchronised
        return outs

    def _resolve(self, status: torch.Tensor, images, outs, cb: int):
        """One host read of the status words of a ``score`` call (a stream synchronisation; every caller reads results
        back right after)."""
        st = status.cpu().numpy()
        if st[:, 0].any():
            _raise_status(int(st[:, 0][st[:, 0] != 0][0]))
        bad = np.nonzero(st[:, 1] == _FP16_OVERFLOW)[0]
        if bad.size == 0:
            return
        if self.mode_name != "auto":
            raise RuntimeError("strainer_b200: non-finite logit in the fp16 conv mode: an activation exceeded 65504 (or the "
                               "input is not finite); use conv_mode='auto' (re-scores such chunks in fp32-parity "
                               "arithmetic), 'fp32' or 'bf16'")
        fb = self.fallback()
        n = images.shape[0]
        for ci in bad:
            i0, i1 = int(ci) * cb, min(n, (int(ci) + 1) * cb)
            part = images[i0:i1]
            if not isinstance(part, U8Images):
                part = part.to(self.device)
            sub = fb.score(part, tuple(outs.keys()))
            for k, v in outs.items():
                v[i0:i1].copy_(sub[k])
            self.fallback_chunks += 1


def _bump_versions(tensors):
    """Tensors that a kernel wrote through raw pointers get their ``_version`` bumped, so that every cache keyed on it
    (the BN folds of all scorers of this module) is rebuilt."""
    ts = [t for t in tensors if isinstance(t, torch.Tensor)]
    if not ts:
        return
    try:    # no kernel launch: the counters are host-side objects
        torch._C._autograd._unsafe_set_version_counter(ts, [t._version + 1 for t in ts])
    except Exception:
        with torch.no_grad():
            torch._foreach_mul_(ts, 1.0)


MLP_TC_MIN_BATCH = 1024      # below this the chain is launch bound and the fp32 small-batch kernels win


class _SmallNetScorer:
    """Shared ``score`` loop of the 28x28 scorers: chunks of ``max_batch`` rows, host rows copied per chunk, one status read
    at the end; a chunk that produced a non-finite logit in the fp16 tensor-core form is scored again in fp32."""

    def _score_chunk_tc(self, x, b, logit, prob, loss, status):
        raise NotImplementedError

    def _score_chunk_fp32(self, x, b, logit, prob, loss):
        raise NotImplementedError

    def score(self, images, want=("loss",)):
        n = images.shape[0]
        outs = {k: torch.empty(n, dtype=torch.float32, device=self.device) for k in want}
        if n == 0:
            return outs
        if isinstance(images, U8Images):
            images = images.to_f32(self.device)
        rows = images.reshape(n, -1)
        if rows.shape[1] != 784:
            raise ValueError("the 28x28 scorers take [N,1,28,28] or [N,784] inputs")
        cb = self.max_batch
        nchunks = (n + cb - 1) // cb
        status = torch.zeros((nchunks, 2), dtype=torch.int32, device=self.device)
        tc = []
        for ci in range(nchunks):
            i = ci * cb
            x = _f32c(rows[i:i + cb], self.device)
            b = x.shape[0]
            o = [outs[k][i:i + b] if k in outs else None for k in ("logit", "prob", "loss")]
            if self.use_tc(b):
                self._score_chunk_tc(x, b, *o, status[ci])
                tc.append(ci)
            else:
                self._score_chunk_fp32(x, b, *o)
        if tc:
            st = status.cpu().numpy()
            if st[:, 0].any():
                _raise_status(int(st[:, 0][st[:, 0] != 0][0]))
            for ci in np.nonzero(st[:, 1] != 0)[0]:
                if self.mode_name == "fp16":
                    raise RuntimeError("strainer_b200: non-finite logit in the fp16 tensor-core form; use conv_mode='auto' or 'fp32'")
                i = int(ci) * cb
                x = _f32c(rows[i:i + cb], self.device)
                b = x.shape[0]
                self._score_chunk_fp32(x, b, *[outs[k][i:i + b] if k in outs else None for k in ("logit", "prob", "loss")])
                self.fallback_chunks += 1
        return outs

    def use_tc(self, b: int) -> bool:
        return self.mode_name in ("fp16", "bf16") or (self.mode_name == "auto" and b >= MLP_TC_MIN_BATCH)


class MLPScorer(_SmallNetScorer):
    """Scores the reference's 28x28 MLP Discriminator ("Untitled-2.py:79-94", "# 1,2,8.py:110-128").
    mode 'auto' (default): batches of >= MLP_TC_MIN_BATCH rows run as a tcgen05 GEMM chain (fp16 operands, fp32
    accumulation, the 1e-3 bar) with an fp32 re-score of a chunk whose logits overflowed; smaller batches and mode 'fp32'
    use the fp32 CUDA-core kernels (launch bound at the reference's B = 64); 'fp16': the tensor-core chain at any size."""

    def __init__(self, discriminator: nn.Module, device=None, max_batch: int = 4096, mode: str = "auto"):
        self.device = _dev(device)
        if mode not in ("auto", "fp32", "fp16"):
            raise ValueError("MLP scorer modes: 'auto', 'fp32', 'fp16'")
        self.mode_name = mode
        self.max_batch = int(max_batch)
        self.fallback_chunks = 0
        self.ws = torch.empty(self.lib.sg_mlp_workspace_bytes(self.max_batch), dtype=torch.uint8, device=self.device)
        self.ws_tc = None
        self.packed = None
        self._packed_sig = None
        self.refresh(discriminator)

    @property
    def lib(self):
        return _lib_for(self.device)

    @staticmethod
    def linears(discriminator):
        lins = [m for m in discriminator.modules() if isinstance(m, nn.Linear)]
        shapes = [tuple(m.weight.shape) for m in lins]
        if shapes != [(1024, 784), (512, 1024), (256, 512), (1, 256)] or any(m.bias is None for m in lins):
            raise NotImplementedError(f"strainer_b200 scores the reference 784-1024-512-256-1 MLP only; got {shapes}")
        return lins

    def refresh(self, discriminator, force: bool = False):
        lins = self.linears(discriminator)
        sig = tuple((t.data_ptr(), t._version) for m in lins for t in (m.weight, m.bias))
        if sig == getattr(self, "_sig", None) and not force:
            return
        self._sig = sig
        self.params = []
        for m in lins:
            self.params += [_f32c(m.weight.detach(), self.device), _f32c(m.bias.detach(), self.device)]
        self.arr = (L.P * 8)(*[t.data_ptr() for t in self.params])

    def _ensure_tc(self):
        lib = self.lib
        if self.ws_tc is None:
            self.ws_tc = _aligned_empty(lib.sg_mlp_tc_workspace_bytes(self.max_batch), self.device)
            self.ws_tc[:1024].zero_()
            self.packed = _aligned_empty(lib.sg_mlp_tc_packed_bytes(), self.device)
        if self._packed_sig != self._sig:
            L.check(lib.sg_mlp_tc_pack(self.arr, _p(self.packed), _stream()), "sg_mlp_tc_pack")
            self._packed_sig = self._sig

    def _score_chunk_tc(self, x, b, logit, prob, loss, status):
        self._ensure_tc()
        L.check(self.lib.sg_mlp_score_tc(_p(x), b, self.arr, _p(self.packed), _p(self.ws_tc), _p(logit), _p(prob), _p(loss),
                                         _p(status), _stream()), "sg_mlp_score_tc")

    def _score_chunk_fp32(self, x, b, logit, prob, loss):
        L.check(self.lib.sg_mlp_score(_p(x), b, self.arr, _p(self.ws), _p(logit), _p(prob), _p(loss), _stream()),
                "sg_mlp_score")

    def score_into(self, x, logit=None, prob=None, loss=None):
        """One batch (<= max_batch rows, fp32 CUDA [b,784]) through the fp32 kernels (the in-batch strain block)."""
        b = x.shape[0]
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        self._score_chunk_fp32(x, b, logit, prob, loss)


def _d28_modules(discriminator: nn.Module):
    convs = [m for m in discriminator.modules() if isinstance(m, nn.Conv2d)]
    bns = [m for m in discriminator.modules() if isinstance(m, nn.BatchNorm2d)]
    shapes = [tuple(c.weight.shape) for c in convs]
    ok = shapes == [(64, 1, 4, 4), (128, 64, 4, 4), (1, 128, 7, 7)] and len(bns) == 1 and all(c.bias is None for c in convs)
    ok = ok and [c.stride for c in convs] == [(2, 2), (2, 2), (1, 1)] and [c.padding for c in convs] == [(1, 1), (1, 1), (0, 0)]
    return (convs, bns) if ok else None


class D28Scorer(_SmallNetScorer):
    """Scores the DCGAN-28 conv discriminator of BASELINE config 1 (SURVEY 8d C1 option ii; ``oracle.Discriminator28``):
    conv 1 on the CUDA cores writing the im2col rows of conv 2, conv 2 + folded eval-mode BN + LeakyReLU as one tcgen05
    GEMM in fp16, conv 3 + sigmoid + BCE as a dot per image.  Eval-mode BatchNorm only."""

    def __init__(self, discriminator: nn.Module, device=None, max_batch: int = 4096, mode: str = "auto"):
        self.device = _dev(device)
        if mode not in ("auto", "fp16"):
            raise ValueError("DCGAN-28 scorer modes: 'auto', 'fp16'")
        self.mode_name = mode
        self.max_batch = int(max_batch)
        self.fallback_chunks = 0
        lib = self.lib
        self.packed = _aligned_empty(lib.sg_d28_packed_bytes(), self.device)
        self.ws = _aligned_empty(lib.sg_d28_workspace_bytes(self.max_batch), self.device)
        self.ws[:1024].zero_()
        self._sig = None
        self.refresh(discriminator)

    @property
    def lib(self):
        return _lib_for(self.device)

    def use_tc(self, b: int) -> bool:
        return True

    def refresh(self, discriminator, force: bool = False):
        mods = _d28_modules(discriminator)
        if mods is None:
            raise NotImplementedError("not the DCGAN-28 discriminator (Conv 1->64 k4s2p1, Conv 64->128 k4s2p1 + BN, Conv 128->1 k7)")
        convs, bns = mods
        bn = bns[0]
        ts = [c.weight for c in convs] + [bn.weight, bn.bias, bn.running_mean, bn.running_var]
        sig = tuple((t.data_ptr(), t._version) for t in ts)
        if sig == self._sig and not force:
            return
        self._sig = sig
        self.params = [_f32c(t.detach(), self.device) for t in ts]
        w1, w2, w3, g, bt, m, v = self.params
        L.check(self.lib.sg_d28_pack(_p(w2), _p(w3), _p(g), _p(bt), _p(m), _p(v), float(bn.eps), _p(self.packed), _stream()),
                "sg_d28_pack")

    def _score_chunk_tc(self, x, b, logit, prob, loss, status):
        L.check(self.lib.sg_d28_score(_p(x), b, _p(self.params[0]), _p(self.packed), _p(self.ws), _p(logit), _p(prob), _p(loss),
                                      _p(status), _stream()), "sg_d28_score")

    def _score_chunk_fp32(self, x, b, logit, prob, loss):
        raise RuntimeError("strainer_b200: non-finite logit in the DCGAN-28 fp16 tensor-core form (an activation beyond 65504); "
                           "this scorer has no fp32 form")

    def score_into(self, x, logit=None, prob=None, loss=None, status=None):
        b = x.shape[0]
        if b > self.max_batch:
            raise ValueError(f"batch {b} exceeds max_batch {self.max_batch}")
        self._score_chunk_tc(x, b, logit, prob, loss, status)


_MLP_SCORERS: dict = {}


def _cache_put(cache: dict, key, module: nn.Module, scorer):
    """Scorer caches hold packed weights and multi-GB workspaces: an entry lives exactly as long as its module (the
    scorer itself only keeps a weak reference to it)."""
    cache[key] = scorer
    weakref.finalize(module, cache.pop, key, None)


def get_mlp_scorer(discriminator: nn.Module, device=None, max_batch: int = 4096, mode: str = "auto") -> "MLPScorer":
    """Per-module cache: the weights are uploaded again only when a parameter tensor changed."""
    device = _dev(device)
    key = (id(discriminator), device.index, max_batch, mode)
    sc = _MLP_SCORERS.get(key)
    if sc is None:
        sc = MLPScorer(discriminator, device, max_batch, mode)
        _cache_put(_MLP_SCORERS, key, discriminator, sc)
    else:
        sc.refresh(discriminator)
    return sc


def get_d28_scorer(discriminator: nn.Module, device=None, max_batch: int = 4096, mode: str = "auto") -> "D28Scorer":
    device = _dev(device)
    key = ("d28", id(discriminator), device.index, max_batch, mode)
    sc = _MLP_SCORERS.get(key)
    if sc is None:
        sc = D28Scorer(discriminator, device, max_batch, mode)
        _cache_put(_MLP_SCORERS, key, discriminator, sc)
    else:
        sc.refresh(discriminator)
    return sc


_KINDS = weakref.WeakKeyDictionary()      # module -> 'mlp' | 'd28' | 'd64': the structure walk costs ~50 us, a batch call has ~200


def _kind(netD) -> str:
    k = _KINDS.get(netD)
    if k is None:
        if any(isinstance(m, nn.Linear) for m in netD.modules()) and not any(isinstance(m, nn.Conv2d) for m in netD.modules()):
            k = "mlp"
        else:
            k = "d28" if _d28_modules(netD) is not None else "d64"
        _KINDS[netD] = k
    return k


def _is_d28(netD) -> bool:
    return _kind(netD) == "d28"


def scorer_for(discriminator: nn.Module, device=None, conv_mode: str = "auto", max_batch: int = 4096):
    """The scorer of whichever discriminator of the path this is: the 64x64 DCGAN D (tcgen05 implicit-GEMM convs), the
    28x28 MLP D or the DCGAN-28 conv D (tcgen05 GEMM chains).  All expose ``score(images, want)``."""
    if _is_mlp(discriminator):
        return get_mlp_scorer(discriminator, device, max_batch, conv_mode if conv_mode in ("auto", "fp32", "fp16") else "auto")
    if _is_d28(discriminator):
        return get_d28_scorer(discriminator, device, max_batch, "auto")
    return get_scorer(discriminator, device, conv_mode, max_batch)


def _is_mlp(netD) -> bool:
    return _kind(netD) == "mlp"


_COPY_STREAMS: dict = {}


def _copy_stream(device):
    s = _COPY_STREAMS.get(device.index)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _COPY_STREAMS[device.index] = s
    return s


_SCORERS: dict = {}
DATASET_CHUNK = 8192     # samples per scoring launch group at dataset scale (tools/chunk_sweep.py; 2 GB of activations)


def _chunk_for(n: int) -> int:
    """Launch-group size for a dataset of n samples: DATASET_CHUNK at scale, the next power of two (>= 512) below it, so
    that a small dataset does not allocate the 2 GB workspace of the large one."""
    c = 512
    while c < min(int(n), DATASET_CHUNK):
        c *= 2
    return c


def get_scorer(discriminator: nn.Module, device=None, mode: str = "auto", max_batch: int = 4096) -> D64Scorer:
    device = _dev(device)
    key = (id(discriminator), device.index, mode, max_batch)
    sc = _SCORERS.get(key)
    if sc is None:
        sc = D64Scorer(discriminator, device, mode, max_batch)
        _cache_put(_SCORERS, key, discriminator, sc)
    else:
        sc.repack(discriminator)
    return sc


def clear_scorer_caches():
    """Drops every cached scorer (packed weights, workspaces, staging buffers) and scratch workspace."""
    _SCORERS.clear()
    _MLP_SCORERS.clear()
    _Scratch._cache.clear()


# ----------------------------------------------------------------------------------------------
# uint8 datasets: ToTensor + Normalize on the device
# ----------------------------------------------------------------------------------------------
class U8Images:
    """A batch of uint8 images plus the ``Normalize(mean, std)`` that follows ``ToTensor`` in the reference's
    transform ("#strainer gan.py:89-90").  ``pixels``: uint8 ``[N,C,H,W]`` (layout 'NCHW') or ``[N,H,W,C]`` ('NHWC',
    what ``np.asarray(PIL image)`` gives), host (pinned or not) or CUDA.  ``to_f32`` produces, on the device, exactly the
    fp32 NCHW tensor the host transform would (``sg_u8_normalize``: correctly rounded fp32 division / subtraction)."""

    def __init__(self, pixels: torch.Tensor, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), layout: str = "NCHW"):
        if not isinstance(pixels, torch.Tensor):
            pixels = torch.as_tensor(np.asarray(pixels))
        if pixels.dtype != torch.uint8 or pixels.dim() != 4:
            raise ValueError("pixels must be a uint8 tensor [N,C,H,W] or [N,H,W,C]")
        if layout not in ("NCHW", "NHWC"):
            raise ValueError("layout must be 'NCHW' or 'NHWC'")
        self.pixels = pixels.contiguous()
        self.layout = layout
        n, a, b, c = self.pixels.shape
        self.chw = (a, b, c) if layout == "NCHW" else (c, a, b)
        ch = self.chw[0]
        mean = [float(m) for m in (mean if hasattr(mean, "__len__") else [mean] * ch)]
        std = [float(v) for v in (std if hasattr(std, "__len__") else [std] * ch)]
        if not 1 <= ch <= 4 or len(mean) != ch or len(std) != ch:
            raise ValueError("1..4 channels with one mean / std per channel")
        self.mean, self.std = tuple(mean), tuple(std)
        self._cmean = (L.c_float * ch)(*mean)
        self._cstd = (L.c_float * ch)(*std)

    @property
    def shape(self):
        return (self.pixels.shape[0],) + self.chw

    @property
    def is_cuda(self):
        return self.pixels.is_cuda

    def __len__(self):
        return self.pixels.shape[0]

    def _like(self, pixels):
        return U8Images(pixels, self.mean, self.std, self.layout)

    def __getitem__(self, key):
        if isinstance(key, slice):
            return self._like(self.pixels[key])
        idx = torch.as_tensor(key, dtype=torch.long, device=self.pixels.device).reshape(-1)   # an int selects one image
        return self._like(self.pixels.index_select(0, idx))

    def normalize_into(self, src_dev: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """src_dev: device uint8 rows in this object's layout -> out: fp32 [b,C,H,W] (stream ordered)."""
        b = src_dev.shape[0]
        c, h, w = self.chw
        lay = L.SG_LAYOUT_NCHW if self.layout == "NCHW" else L.SG_LAYOUT_NHWC
        L.check(_lib_for(out.device).sg_u8_normalize(_p(src_dev), b, c, h * w, lay, self._cmean, self._cstd, _p(out),
                                                     _stream()), "sg_u8_normalize")
        return out

    def to_f32(self, device=None) -> torch.Tensor:
        device = _dev(device if device is not None else (self.pixels.device if self.is_cuda else None))
        out = torch.empty(self.shape, dtype=torch.float32, device=device)
        return self.normalize_into(self.pixels.to(device), out)

    def host_f32(self, i=None) -> torch.Tensor:
        """The host transform itself (ToTensor + Normalize as torchvision computes them): what ``__getitem__`` of the
        dataset yields, so that unmodified reference code sees the same values the device path computes."""
        px = self.pixels if i is None else self.pixels[i]
        px = px.cpu()
        if self.layout == "NHWC":
            px = px.permute(2, 0, 1) if px.dim() == 3 else px.permute(0, 3, 1, 2)
        x = px.contiguous().to(torch.float32).div(255)
        shape = (-1, 1, 1)
        mean = torch.as_tensor(self.mean, dtype=torch.float32).view(shape)
        std = torch.as_tensor(self.std, dtype=torch.float32).view(shape)
        return x.sub_(mean).div_(std)


class U8ImageDataset(torch.utils.data.Dataset):
    """Map-style dataset over uint8 pixels whose ``__getitem__`` yields ``(ToTensor+Normalize(image), label)`` exactly as
    the reference's ``ImageFolder(..., transform=Compose([..., ToTensor(), Normalize(mean, std)]))`` does
    ("#strainer gan.py:85-91"): unmodified reference code can iterate it through a ``DataLoader``, while the functions of
    this module recognise it, keep it uint8 over PCIe / in HBM and normalise on the device (bit-identical values)."""

    def __init__(self, pixels, labels=None, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), layout: str = "NCHW"):
        self.images = pixels if isinstance(pixels, U8Images) else U8Images(pixels, mean, std, layout)
        self.labels = labels

    def __len__(self):
        return len(self.images)

    def __getitem__(self, i):
        label = 0 if self.labels is None else self.labels[i]
        return self.images.host_f32(int(i)), label


def _device_f32_chunks(images, device, chunk: int):
    """Yields (start, fp32 device tensor [b,C,H,W]) over a float tensor or a U8Images (host or CUDA)."""
    n = images.shape[0]
    if isinstance(images, U8Images):
        buf = torch.empty((min(chunk, max(n, 1)),) + images.chw, dtype=torch.float32, device=device)
        for i in range(0, n, chunk):
            src = images.pixels[i:i + chunk].to(device, non_blocking=True)
            yield i, images.normalize_into(src, buf[:src.shape[0]])
        return
    for i in range(0, n, chunk):
        yield i, _f32c(images[i:i + chunk], device)


def _resident_images(dataset):
    """The image tensor behind a dataset that exposes one (tensor, U8Images, U8ImageDataset, TensorDataset, Subset of
    those), else None."""
    from torch.utils.data import Subset, TensorDataset
    if isinstance(dataset, (torch.Tensor, U8Images)):
        return dataset
    if isinstance(dataset, U8ImageDataset):
        return dataset.images
    if isinstance(dataset, TensorDataset):
        return dataset.tensors[0]
    if isinstance(dataset, Subset):
        base = _resident_images(dataset.dataset)
        if base is None:
            return None
        idx = np.asarray(dataset.indices).reshape(-1)
        if isinstance(base, U8Images):
            return base[idx]
        return base.index_select(0, torch.as_tensor(idx, dtype=torch.long, device=base.device))
    return None


STREAM_BATCH = 4096      # samples per DataLoader batch when a dataset has to be streamed through __getitem__
STREAM_WORKERS = 2       # the reference's own loaders use workers=2 ("#strainer gan.py:50")


def _iter_image_chunks(dataset, chunk: int = None):
    """Image batches of a map-style dataset WITHOUT a resident tensor (e.g. an ImageFolder with transforms), through a
    DataLoader like the reference's ("#strainer gan.py:366": batch_size 64, shuffle False), only with a larger batch:
    never more than one batch of decoded images is held on the host."""
    loader = torch.utils.data.DataLoader(dataset, batch_size=chunk or STREAM_BATCH, shuffle=False,
                                         num_workers=STREAM_WORKERS if len(dataset) > 4 * (chunk or STREAM_BATCH) else 0)
    for batch in loader:
        yield batch[0] if isinstance(batch, (list, tuple)) else batch


def _dataset_images(dataset):
    """Resident image tensor of a dataset if it exposes one, else the dataset materialised through its own
    ``__getitem__`` (small datasets / feature rows only: the scoring entry points stream instead, ``_score_dataset``)."""
    imgs = _resident_images(dataset)
    if imgs is not None:
        return imgs
    return torch.cat(list(_iter_image_chunks(dataset)), dim=0)


def _score_dataset(scorer, dataset, want=("loss",)):
    """``scorer.score`` over a dataset: in one call when it exposes a resident tensor, else streamed batch by batch
    through a DataLoader (bounded host memory; the losses are assembled on the device)."""
    imgs = _resident_images(dataset)
    if imgs is not None:
        return scorer.score(imgs, want)
    parts = [scorer.score(x.contiguous().float(), want) for x in _iter_image_chunks(dataset)]
    if not parts:
        return {k: torch.empty(0, dtype=torch.float32, device=scorer.device) for k in want}
    return {k: torch.cat([p_[k] for p_ in parts]) for k in want}


# ----------------------------------------------------------------------------------------------
# the reference's function boundary
# ----------------------------------------------------------------------------------------------
def evaluate_dataset(netD, dataset, device, *, conv_mode: str = "auto", return_device: bool = False):
    """``evaluate_dataset`` ("#clean 분포와 ... .py:272-287", "# 종합 loss.py:315-330"): per-sample
    BCE(D(x), 1) with eval-mode BN (sticky ``netD.eval()``); returns np.ndarray (N,) float32."""
    device = _dev(device)
    netD.eval()
    losses = _score_dataset(scorer_for(netD, device, conv_mode, _chunk_for(len(dataset))), dataset, ("loss",))["loss"]
    return losses if return_device else losses.cpu().numpy()


def refine_dataset_by_loss(dataset, discriminator, device, loss_ratio=0.2, *, conv_mode: str = "auto"):
    """``refine_dataset_by_loss`` ("#strainer gan.py:364-392"): score every sample, threshold at the
    (1-loss_ratio)*100 percentile, keep ``loss < threshold`` in ascending index order.
    Returns (torch.utils.data.Subset, np.float32 threshold).  conv_mode 'auto' (default): one fp16 tensor pass held to
    the fp32 bar (1e-3), chunks that overflow fp16 re-scored in fp32-parity arithmetic; 'fp32' / 'bf16' / 'fp16'."""
    device = _dev(device)
    discriminator.eval()  # sticky, as in the reference
    n = len(dataset)
    losses = _score_dataset(scorer_for(discriminator, device, conv_mode, _chunk_for(n)), dataset, ("loss",))["loss"]
    clean_indices, threshold = select_below_percentile(losses, (1 - loss_ratio) * 100)
    if len(clean_indices) == 0:
        # reference fallback on its (N,1,1)-shaped loss array: argsort along the last axis (len 1) -> zeros
        clean_indices = np.zeros((max(n // 2, 1), 1, 1), dtype=np.intp)
    return torch.utils.data.Subset(dataset, clean_indices), threshold


def _group_total(n_local: int, group, device) -> int:
    """Global element count of a sharded vector: the sum of the shard lengths over ``group`` (one 8-byte all-reduce)."""
    import torch.distributed as dist
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=device if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def _select_check(device, key="select"):
    """After a synchronisation point: raises if the select kernels of the last call on ``device`` gave up (a grid barrier
    of the cooperative kernel, or a peer that never arrived in the NVLink all-reduce)."""
    ws = _Scratch._cache.get((device.index, key))
    if ws is not None:
        L.check(_lib_for(device).sg_select_check(_p(ws), _stream()), "sg_select_check")


def select_below_percentile(losses: torch.Tensor, q, group=None, index_base: int = 0, n_global=None, comm="auto"):
    """threshold = np.percentile(losses, q); indices = np.where(losses < threshold)[0] -- on device.
    One host sync at the end (count + threshold).  With ``group`` the losses are this rank's shard of a global vector
    of ``n_global`` elements (summed over the group when not given).  Returns (np.int64 indices, np.float32 threshold)."""
    if group is not None and n_global is None:
        n_global = _group_total(losses.numel(), group, losses.device)
    if isinstance(comm, str):      # 'auto': the group's NVLink peer communicator when there is one, else NCCL all-reduces
        comm = PeerComm.for_group(group, losses.device) if (group is not None and losses.is_cuda) else None
    thr = percentile_device(losses, q, group, n_global, comm=comm)
    idx, count, _ = compact_indices(losses, thr, L.SG_LT, index_base)
    thr_h = thr.cpu().numpy()[0]
    c = int(count.item())
    if group is None:
        _select_check(losses.device)
    elif comm is not None:
        _select_check(losses.device, "select_peer")
    return idx[:c].cpu().numpy(), thr_h


def strain_shard(images, discriminator, loss_ratio=0.2, *, group=None, index_base: int = 0,
                 n_global=None, conv_mode: str = "auto", device=None, comm="auto"):
    """Data-parallel ``refine_dataset_by_loss`` ("#strainer gan.py:364-392") for one rank of a
    sharded dataset: this rank scores ``images`` (global indices index_base ...), the threshold is the
    GLOBAL percentile (only the radix histograms cross NVLink: inside the select kernels over peer memory when the group
    has a ``PeerComm``, else as NCCL all-reduces; ``comm=None`` forces the latter), and the returned kept indices are global.
    Concatenating the per-rank index arrays in rank order reproduces the single-GPU / reference
    ``np.where`` order.  ``n_global``: total sample count over the group (all-reduced when not given).
    Returns (np.int64 kept_global_indices, np.float32 threshold, losses_dev)."""
    device = _dev(device if device is not None else _dev_of(images))
    discriminator.eval()
    losses = scorer_for(discriminator, device, conv_mode, _chunk_for(images.shape[0])).score(images, ("loss",))["loss"]
    idx, thr = select_below_percentile(losses, (1 - loss_ratio) * 100, group, index_base, n_global, comm)
    return idx, thr, losses


class ResidentSubset:
    """Device-resident replacement for ``Subset`` + a fresh ``DataLoader`` per epoch ("# final.py:444",
    SURVEY 8f item 1): the strain result stays a kept-index tensor on the GPU and every batch is one
    vectorised row gather from the resident image tensor -- no host index list, no per-sample
    ``__getitem__``, no H2D copies.

        sub = ResidentSubset.refine(images_dev, netD, loss_ratio=0.2)     # = refine_dataset_by_loss, on device
        for epoch in ...:
            for real in sub.batches(128, shuffle=True, generator=g):       # [b,3,64,64] device tensors
                ...

    ``indices`` are ascending (the reference's ``np.where`` order); ``batches(shuffle=False)`` therefore yields
    exactly ``DataLoader(Subset(dataset, clean_indices), batch_size, shuffle=False)``'s image batches."""

    def __init__(self, images, indices: torch.Tensor, threshold=None):
        # a U8Images keeps the resident rows uint8 (12 288 B per sample); batches are normalised after the gather
        self.u8 = images if isinstance(images, U8Images) else None
        rows = images.pixels if self.u8 is not None else images
        if not rows.is_cuda:
            raise ValueError("ResidentSubset keeps the dataset in HBM: pass a CUDA tensor (or a U8Images of one)")
        self.images = rows.contiguous()
        self.indices = indices.to(device=rows.device, dtype=torch.int64).contiguous()
        self.threshold = threshold

    def __len__(self):
        return int(self.indices.numel())

    @classmethod
    def refine(cls, images, discriminator, loss_ratio=0.2, *, conv_mode: str = "auto"):
        """``refine_dataset_by_loss`` ("#strainer gan.py:364-392") without leaving the device."""
        device = _dev(images.pixels.device if isinstance(images, U8Images) else images.device)
        discriminator.eval()
        losses = get_scorer(discriminator, device, conv_mode, _chunk_for(images.shape[0])).score(images, ("loss",))["loss"]
        thr = percentile_device(losses, (1 - loss_ratio) * 100)
        idx, count, _ = compact_indices(losses, thr, L.SG_LT, 0)
        c = int(count.item())
        if c == 0:   # the reference's degenerate fallback keeps index 0, n // 2 times
            return cls(images, torch.zeros(max(images.shape[0] // 2, 1), dtype=torch.int64, device=device), thr)
        return cls(images, idx[:c], thr)

    def batches(self, batch_size: int, shuffle: bool = True, generator=None, drop_last: bool = False):
        device = self.images.device
        lib = _lib_for(device)
        n = len(self)
        order = self.indices
        if shuffle:
            perm = torch.randperm(n, device=device, generator=generator)
            order = self.indices.index_select(0, perm)
        row_bytes = self.images[0].numel() * self.images.element_size()
        for i in range(0, n, batch_size):
            idx = order[i:i + batch_size]
            if drop_last and idx.numel() < batch_size:
                break
            out = torch.empty((idx.numel(),) + tuple(self.images.shape[1:]), dtype=self.images.dtype, device=device)
            L.check(lib.sg_gather_rows(_p(self.images), row_bytes, _p(idx), idx.numel(), L.P(0), _p(out), _stream()),
                    "sg_gather_rows")
            if self.u8 is not None:
                out = self.u8.normalize_into(out, torch.empty((idx.numel(),) + self.u8.chw, dtype=torch.float32, device=device))
            yield out


def get_percentile_threshold(losses, percentile=75):
    """"# 종합 loss.py:287-288" on a device (or host) loss vector."""
    lt = _f32c(losses, _dev_of(losses)).reshape(-1)
    return percentile_device(lt, percentile).cpu().numpy()[0]


def get_iqr_threshold(losses):
    """"# 종합 loss.py:290-294": Q3 + 1.5 * (Q3 - Q1)."""
    lt = _f32c(losses, _dev_of(losses)).reshape(-1)
    q1 = percentile_device(lt, 25).cpu().numpy()[0]
    q3 = percentile_device(lt, 75).cpu().numpy()[0]
    return q3 + 1.5 * (q3 - q1)


def _gmm_intersection(means, stds):
    ci = np.argmin(means)
    ni = 1 - ci
    a = 1 / (2 * stds[ci] ** 2) - 1 / (2 * stds[ni] ** 2)
    b = means[ni] / (stds[ni] ** 2) - means[ci] / (stds[ci] ** 2)
    c = means[ci] ** 2 / (2 * stds[ci] ** 2) - means[ni] ** 2 / (2 * stds[ni] ** 2) - np.log(stds[ni] / stds[ci])
    return (-b + np.sqrt(b ** 2 - 4 * a * c)) / (2 * a)


class _GmmOps:
    """The three phases of the device EM on one GPU (tests substitute a numpy double to exercise the multi-rank
    protocol on CPU/gloo).  ``sums`` is the 8-double tensor that is all-reduced between accumulate and update."""

    def __init__(self, device):
        self.device = device
        self.ws = torch.empty(self.lib.sg_gmm1d_workspace_bytes(), dtype=torch.uint8, device=device)
        self.sums = self.ws[16 * 8:24 * 8].view(torch.float64)

    @property
    def lib(self):
        return _lib_for(self.device)

    def prepare(self, losses):
        return _f32c(losses, self.device).reshape(-1)

    def begin(self, centers, kmeans_iters):
        L.check(self.lib.sg_gmm1d_begin(_p(centers), kmeans_iters, _p(self.ws), _stream()), "sg_gmm1d_begin")

    def accumulate(self, v):
        L.check(self.lib.sg_gmm1d_accumulate(_p(v), v.numel(), _p(self.ws), _stream()), "sg_gmm1d_accumulate")

    def update(self, n_total, reg_covar, tol, max_iter):
        L.check(self.lib.sg_gmm1d_update(n_total, float(reg_covar), float(tol), int(max_iter), _p(self.ws), _stream()),
                "sg_gmm1d_update")

    def state(self):
        return self.ws[:16 * 8].view(torch.float64).cpu().numpy()


def gmm_fit_device(losses, max_iter: int = 10, tol: float = 1e-2, reg_covar: float = 5e-4, *, group=None,
                   n_global=None, kmeans_iters: int = 30, ops=None, select_ops=None):
    """2-component 1-D Gaussian-mixture EM on the GPU: sklearn's EM equations
    (``GaussianMixture(n_components=2, max_iter, tol, reg_covar)``) with a DETERMINISTIC initialisation -- Lloyd
    iterations from the 25 % / 75 % order statistics instead of a k-means run seeded by the global numpy RNG.
    With ``group`` the losses are this rank's shard: 8 partial sums are all-reduced per iteration, every rank
    obtains the identical fit.  Returns dict(weights, means, stds, n_iter, converged) (float64 numpy)."""
    ops = ops or _GmmOps(_dev_of(losses))
    v = ops.prepare(losses)
    n = v.numel()
    if group is not None and n_global is None:
        n_global = _group_total(n, group, v.device)
    n_tot = int(n_global) if n_global is not None else n
    c0 = order_stats(v, (n_tot - 1) // 4, group, select_ops)[0:1]
    c1 = order_stats(v, (3 * (n_tot - 1)) // 4, group, select_ops)[0:1]
    ops.begin(torch.cat([c0, c1]), kmeans_iters)
    for _ in range(kmeans_iters + max_iter):
        ops.accumulate(v)
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(ops.sums, op=dist.ReduceOp.SUM, group=group)
        ops.update(n_tot, reg_covar, tol, max_iter)
    st = ops.state()
    return {"weights": st[0:2].copy(), "means": st[2:4].copy(), "stds": np.sqrt(st[4:6]), "n_iter": int(st[7]),
            "converged": bool(st[8]), "lower_bound": float(st[6])}


def get_gmm_threshold(losses, *, fit: str = "sklearn"):
    """"# 종합 loss.py:270-285".  fit='sklearn' (default, the reference's own call: the EM fit is seeded by the
    global np.random state exactly like upstream); fit='device': the deterministic GPU EM of ``gmm_fit_device``."""
    if fit == "device":
        g = gmm_fit_device(losses)
        return _gmm_intersection(g["means"], g["stds"])
    from sklearn.mixture import GaussianMixture
    lo = losses.cpu().numpy() if isinstance(losses, torch.Tensor) else np.asarray(losses)
    gmm = GaussianMixture(n_components=2, max_iter=10, tol=1e-2, reg_covar=5e-4)
    gmm.fit(lo.reshape(-1, 1))
    return _gmm_intersection(gmm.means_.flatten(), np.sqrt(gmm.covariances_.flatten()))


def get_ensemble_threshold(losses):
    """"# 종합 loss.py:296-301"."""
    return np.median([get_gmm_threshold(losses), get_percentile_threshold(losses), get_iqr_threshold(losses)])


def _divide(losses, dataset, threshold):
    lt = _f32c(losses, _dev_of(losses)).reshape(-1)
    n = lt.numel()
    idx, count, _ = compact_indices(lt, threshold, L.SG_LT, 0)
    nidx, ncount, _ = compact_indices(lt, threshold, L.SG_LT | L.SG_NOT, 0)  # ~(loss < thr): NaNs are noisy
    c, nc_ = int(count.item()), int(ncount.item())
    assert c + nc_ == n
    sub = torch.utils.data.Subset
    return sub(dataset, idx[:c].cpu().numpy()), sub(dataset, nidx[:nc_].cpu().numpy())


def divide_dataset(losses, dataset, *, fit: str = "sklearn"):
    """GMM form, "#clean 분포와 ... .py:289-316": clean = losses < intersection threshold."""
    return _divide(losses, dataset, get_gmm_threshold(losses, fit=fit))


def divide_dataset_ensemble(losses, dataset):
    """Ensemble form, "# 종합 loss.py:303-312"."""
    return _divide(losses, dataset, get_ensemble_threshold(losses))


# ---- feature z-score / elbow -------------------------------------------------------------------
def zscore_max(features, ddof: int = 1, eps_add: float = 0.0) -> torch.Tensor:
    """max_j |(x_ij - mean_j) / (std_j + eps_add)| per row of a [N, D] feature matrix
    ("#z_score.py:286-291" with ddof=1; "# 1,2,8.py:164-168" with ddof=0, eps_add=1e-7). Device tensor."""
    device = _dev_of(features)
    lib = _lib_for(device)
    x = _f32c(features, device)
    n, d = x.shape
    mean = torch.empty(d, dtype=torch.float32, device=device)
    denom = torch.empty(d, dtype=torch.float32, device=device)
    ws = _Scratch.get(device, "colmom", lib.sg_col_moments_workspace_bytes(n, d))
    L.check(lib.sg_col_moments(_p(x), n, d, ddof, float(np.float32(eps_add)), _p(mean), _p(denom), _p(ws), _stream()),
            "sg_col_moments")
    out = torch.empty(n, dtype=torch.float32, device=device)
    L.check(lib.sg_row_max_absz(_p(x), n, d, _p(mean), _p(denom), _p(out), _stream()), "sg_row_max_absz")
    return out


def find_elbow_threshold(z_scores, bins=100):
    """``find_elbow_threshold`` ("#strainer gan.py:291-309").  min/max and the 100-bin histogram run
    on the device with numpy's exact bin arithmetic; the 100-element tail is numpy on the host."""
    device = _dev_of(z_scores)
    lib = _lib_for(device)
    z = _f32c(z_scores, device).reshape(-1)
    n = z.numel()
    mm = torch.empty(8, dtype=torch.float32, device=device)
    L.check(lib.sg_minmax(_p(z), n, _p(mm), _stream()), "sg_minmax")
    first, last = mm[:2].cpu().numpy()
    if not (np.isfinite(first) and np.isfinite(last)):
        raise ValueError(f"autodetected range of [{first}, {last}] is not finite")
    if first == last:
        first, last = first - 0.5, last + 0.5
    bin_edges = np.linspace(first, last, bins + 1, endpoint=True, dtype=np.result_type(first, last, np.float32))
    counts = torch.zeros(bins, dtype=torch.int64, device=device)
    edges_d = torch.from_numpy(bin_edges.astype(np.float32)).to(device)
    L.check(lib.sg_hist_uniform(_p(z), n, _p(edges_d), bins, _p(counts), _stream()), "sg_hist_uniform")
    cnt = counts.cpu().numpy()
    db = np.array(np.diff(bin_edges), float)
    hist = cnt / db / cnt.sum()
    bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2
    peak_index = np.argmax(hist)
    target_index = np.argmin(np.abs(hist[peak_index:] - 0.01))
    threshold = (bin_centers[peak_index] + bin_centers[peak_index:][target_index]) / 2
    return threshold, bin_centers, hist


def _features_of(dataset, feature_extractor, device):
    imgs = _dataset_images(dataset)
    if feature_extractor is None or isinstance(feature_extractor, nn.Identity):
        return imgs.to_f32(device) if isinstance(imgs, U8Images) else _f32c(imgs, device)
    feats = []
    with torch.no_grad():
        for _, x in _device_f32_chunks(imgs, device, 64):
            feats.append(feature_extractor(x).float())
    return torch.cat(feats, dim=0)


def detect_outliers(dataset, feature_extractor, user_threshold=None, *, threshold=None, clean_ratio=None):
    """``detect_outliers`` of "#strainer gan.py:331-360" (the canonical script): inlier = max|z| < threshold with the
    threshold given (``user_threshold``, third positional argument as upstream) or found by ``find_elbow_threshold``;
    returns a numpy bool array.  The reference defines three MORE functions of this name in other scripts, whose third
    positional argument means something else; they are exported under their own names so that an unmodified call site
    can never be routed to the wrong rule silently:
      ``detect_outliers_fixed(dataset, fe, threshold=5.0)``   "#z_score.py:276-294"  (strict <, torch bool)
      ``detect_outliers_elbow(dataset, fe)``                  "#z_score + 엘보우 threshold.py:306-330"
      ``detect_outliers_ratio(dataset, fe, clean_ratio)``     "# z_score + DBSCAN.py:305-326" (torch.quantile, <=)
    (a script imports the one it defines: ``from strainer_b200 import detect_outliers_ratio as detect_outliers``).
    The keywords ``threshold=`` / ``clean_ratio=`` select the same variants from this entry point.
    Feature extraction itself (pretrained ResNet18) is out of scope: pass the module, or nn.Identity()
    over a dataset of feature rows."""
    if sum(v is not None for v in (user_threshold, threshold, clean_ratio)) > 1:
        raise TypeError("detect_outliers: give at most one of user_threshold, threshold=, clean_ratio=")
    device = _dev()
    mz = zscore_max(_features_of(dataset, feature_extractor, device))
    if clean_ratio is not None:
        thr = quantile_device(mz, float(clean_ratio))
        _, _, mask = compact_indices(mz, thr, L.SG_LE, 0, want_mask=True)
        return mask.bool().cpu()
    if threshold is not None:
        _, _, mask = compact_indices(mz, float(threshold), L.SG_LT, 0, want_mask=True)
        return mask.bool().cpu()
    thr = find_elbow_threshold(mz)[0] if user_threshold is None else user_threshold
    _, _, mask = compact_indices(mz, thr, L.SG_LT, 0, want_mask=True)
    return mask.bool().cpu().numpy()


def detect_outliers_fixed(dataset, feature_extractor, threshold=5.0):
    """``detect_outliers(dataset, feature_extractor, threshold=5.0)`` of "#z_score.py:276-294": max|z| < threshold
    (strict), CPU torch.BoolTensor."""
    return detect_outliers(dataset, feature_extractor, threshold=float(threshold))


def detect_outliers_elbow(dataset, feature_extractor):
    """``detect_outliers(dataset, feature_extractor)`` of "#z_score + 엘보우 threshold.py:306-330": elbow threshold, numpy bool."""
    return detect_outliers(dataset, feature_extractor)


def detect_outliers_ratio(dataset, feature_extractor, clean_ratio):
    """``detect_outliers(dataset, feature_extractor, clean_ratio)`` of "# z_score + DBSCAN.py:305-326":
    threshold = torch.quantile(max|z|, clean_ratio), inlier = max|z| <= threshold, CPU torch.BoolTensor."""
    return detect_outliers(dataset, feature_extractor, clean_ratio=float(clean_ratio))


def compute_z_scores(dataset, feature_extractor):
    """``compute_z_scores`` ("# 1,2,8.py:154-170"): np.std (ddof 0) + 1e-7; returns np.ndarray (N,)."""
    device = _dev()
    return zscore_max(_features_of(dataset, feature_extractor, device), ddof=0, eps_add=1e-7).cpu().numpy()


# ---- device sort / 1-D DBSCAN ----------------------------------------------------------------
def sort_values(values, return_order: bool = False):
    """Ascending stable device radix sort of a 1-D fp32 vector (NaN last)."""
    device = _dev_of(values)
    lib = _lib_for(device)
    v = _f32c(values, device).reshape(-1)
    n = v.numel()
    out = torch.empty(n, dtype=torch.float32, device=device)
    order = torch.empty(n, dtype=torch.int32, device=device) if return_order else None
    ws = _Scratch.get(device, "sort", lib.sg_sort_workspace_bytes(n))
    L.check(lib.sg_sort_f32(_p(v), n, _p(out), _p(order), _p(ws), _stream()), "sg_sort_f32")
    return (out, order) if return_order else out


def dbscan1d_clean_ratio(values, eps, min_samples=3, return_noise: bool = False):
    """Fraction of non-noise points of ``sklearn.cluster.DBSCAN(eps, min_samples)`` applied to a 1-D
    value vector (the north_star's 1-D variant of ``estimate_ratio_dbscan``,
    "# z_score + DBSCAN.py:291-299"); exact: sort + neighbour counts, no O(N^2) neighbour search."""
    device = _dev_of(values)
    lib = _lib_for(device)
    v = _f32c(values, device).reshape(-1)
    n = v.numel()
    counts = torch.zeros(1, dtype=torch.int64, device=device)
    noise = torch.empty(n, dtype=torch.uint8, device=device) if return_noise else None
    ws = _Scratch.get(device, "dbscan", lib.sg_dbscan1d_workspace_bytes(n))
    L.check(lib.sg_dbscan1d(_p(v), n, float(eps), int(min_samples), _p(counts), _p(noise), _p(ws), _stream()), "sg_dbscan1d")
    ratio = counts.item() / n
    return (ratio, noise.bool()) if return_noise else ratio


def dbscan_clean_ratio(features, eps, min_samples=3, return_counts: bool = False):
    """``mean(DBSCAN(eps, min_samples).fit_predict(StandardScaler().fit_transform(features)) != -1)``
    ("# z_score + DBSCAN.py:291-299") for an [N, d] feature matrix, d a multiple of 64, on the GPU: column
    standardisation + two thresholded pairwise-distance GEMMs on tcgen05 (core points, then points within eps of
    a core point).  Nothing of size N^2 is stored."""
    device = _dev_of(features)
    lib = _lib_for(device)
    f = _f32c(features, device)
    n, d = f.shape
    mean = torch.empty(d, dtype=torch.float32, device=device)
    den = torch.empty(d, dtype=torch.float32, device=device)
    zws = _Scratch.get(device, "colmom", lib.sg_col_moments_workspace_bytes(n, d))
    L.check(lib.sg_col_moments(_p(f), n, d, 0, 0.0, _p(mean), _p(den), _p(zws), _stream()), "sg_col_moments")
    den = torch.where(den == 0, torch.ones_like(den), den)     # StandardScaler: zero-variance columns are left unscaled
    counts = torch.empty(2, dtype=torch.int64, device=device)
    ws = _Scratch.get(device, "dbscan_nd", lib.sg_dbscan_nd_workspace_bytes(n, d))
    L.check(lib.sg_dbscan_nd(_p(f), n, d, _p(mean), _p(den), float(eps), int(min_samples), _p(counts), _p(ws), _stream()),
            "sg_dbscan_nd")
    L.check(lib.sg_dbscan_nd_check(_p(ws), _stream()), "sg_dbscan_nd_check")
    c = counts.cpu().numpy()
    ratio = int(c[1]) / n
    return (ratio, int(c[0]), int(c[1])) if return_counts else ratio


def estimate_ratio_dbscan(dataset, eps=20, min_samples=3, feature_extractor=None, *, neighbors: str = "device"):
    """``estimate_ratio_dbscan`` ("# z_score + DBSCAN.py:272-301"): StandardScaler -> DBSCAN -> fraction of
    non-noise points of the (512-d ResNet18) features.  neighbors='device' (default when the feature width is a
    multiple of 64): ``dbscan_clean_ratio`` on the tensor cores; 'sklearn': the reference's own host call.
    One-dimensional inputs (a score / loss / max-z vector) use the device sort path ``dbscan1d_clean_ratio``."""
    device = _dev()
    feats = _features_of(dataset, feature_extractor, device)
    if feats.dim() == 1 or feats.shape[1] == 1:
        # StandardScaler on one column: (x - mean) / std (ddof 0) from the library's fixed-order fp64 chunk moments
        f = feats.reshape(-1).contiguous()
        lib = _lib_for(device)
        n = f.numel()
        chunks = (n + L.SG_MOMENT_CHUNK - 1) // L.SG_MOMENT_CHUNK
        part = torch.empty(2 * max(chunks, 1), dtype=torch.float64, device=device)
        stats = torch.empty(2, dtype=torch.float64, device=device)
        L.check(lib.sg_chunk_moments(_p(f), n, _p(part), _stream()), "sg_chunk_moments")
        L.check(lib.sg_moments_finish(_p(part), chunks, n, 0.0, _p(stats), L.P(0), _stream()), "sg_moments_finish")
        z = torch.empty(n, dtype=torch.float32, device=device)
        L.check(lib.sg_standardize(_p(f), n, _p(stats), 0, _p(z), _stream()), "sg_standardize")
        return dbscan1d_clean_ratio(z, eps, min_samples)
    if neighbors == "device" and feats.shape[1] % 64 == 0:
        return dbscan_clean_ratio(feats, eps, min_samples)
    from sklearn.cluster import DBSCAN
    from sklearn.preprocessing import StandardScaler
    labels = DBSCAN(eps=eps, min_samples=min_samples).fit_predict(StandardScaler().fit_transform(feats.cpu().numpy()))
    return np.sum(labels != -1) / len(labels)


# ---- auto-encoder straining ---------------------------------------------------------------------
def _ae_params(autoencoder: nn.Module, device):
    mods = [m for m in autoencoder.modules() if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d))]
    want = [(nn.Conv2d, (16, 3, 3, 3)), (nn.Conv2d, (32, 16, 3, 3)), (nn.Conv2d, (64, 32, 7, 7)),
            (nn.ConvTranspose2d, (64, 32, 7, 7)), (nn.ConvTranspose2d, (32, 16, 3, 3)), (nn.ConvTranspose2d, (16, 3, 3, 3))]
    ok = len(mods) == 6 and all(type(m) is t and tuple(m.weight.shape) == sh and m.bias is not None
                                for m, (t, sh) in zip(mods, want))
    if not ok:
        raise NotImplementedError("strainer_b200 scores the reference AutoEncoder (\"#autoencoder.py:269-291\") only")
    ps = []
    for m in mods:
        ps += [_f32c(m.weight.detach(), device), _f32c(m.bias.detach(), device)]
    return ps


AE_CHUNK = 8192      # images per launch group of the auto-encoder pipeline (1.1 GB of 16-bit activations)


def ae_errors(autoencoder: nn.Module, images: torch.Tensor, device=None, chunk: int = AE_CHUNK, *,
              conv_mode: str = "auto") -> torch.Tensor:
    """Per-sample reconstruction MSE of the reference AutoEncoder on the GPU; fp32 device tensor [N].
    conv_mode 'auto' (default): fp16 operands and activations with fp32 accumulation -- one tensor pass, errors within
    the 1e-3 fp32 bar (measured 2e-6) -- and chunks whose errors come out non-finite (an activation beyond fp16's range)
    scored again in fp32-parity arithmetic; 'fp16': the same without the recovery; 'fp32': fp32-parity arithmetic on the
    tensor cores (bf16 hi/lo split of activations and 7x7 weights, three GEMM segments; ~1e-7 relative); 'bf16' (BASELINE
    config 4): bf16 operands and activations; 'fp32_cuda': plain fp32 on the CUDA cores, experiment builds only."""
    device = _dev(device)
    lib = _lib_for(device)
    if conv_mode not in ("auto", "fp32", "fp16", "bf16", "fp32_cuda"):
        raise ValueError("conv_mode must be 'auto', 'fp32', 'fp16', 'bf16' or 'fp32_cuda'")
    params = _ae_params(autoencoder, device)
    arr = (L.P * 12)(*[t.data_ptr() for t in params])
    n = images.shape[0]
    err = torch.empty(n, dtype=torch.float32, device=device)
    cb = min(chunk, max(n, 1))
    if conv_mode == "fp32_cuda":
        if not hasattr(L.load(), "sg_ae_score"):
            raise RuntimeError("conv_mode='fp32_cuda' (the plain fp32 CUDA-core pipeline) is an experiment-build variant: "
                               "rebuild with SG_AB_VARIANTS=1 python strainer-gan_b200/build.py --force")
        ws = _Scratch.get(device, "ae", lib.sg_ae_workspace_bytes(cb))
        for i, x in _device_f32_chunks(images, device, chunk):
            L.check(lib.sg_ae_score(_p(x), x.shape[0], arr, _p(ws), _p(err[i:i + chunk]), L.P(0), _stream()), "sg_ae_score")
        return err
    mode = {"bf16": L.SG_CONV_BF16, "fp16": L.SG_CONV_FP16, "auto": L.SG_CONV_FP16}.get(conv_mode, L.SG_CONV_BF16X3)
    ws = _Scratch.get(device, ("ae_tc", mode), lib.sg_ae_tc_workspace_bytes(cb, mode))
    L.check(lib.sg_ae_pack_tc(arr, _p(ws), mode, _stream()), "sg_ae_pack_tc")      # once; one forward per chunk below
    nchunks = (n + chunk - 1) // chunk
    mm = torch.empty((max(nchunks, 1), 8), dtype=torch.float32, device=device) if conv_mode == "auto" else None
    for ci, (i, x) in enumerate(_device_f32_chunks(images, device, chunk)):
        e = err[i:i + x.shape[0]]
        L.check(lib.sg_ae_forward_tc(_p(x), x.shape[0], arr, _p(ws), mode, _p(e), L.P(0), _stream()), "sg_ae_forward_tc")
        if mm is not None:     # min / max of the chunk's errors (NaN if any NaN): the overflow probe of the auto mode
            L.check(lib.sg_minmax(_p(e), x.shape[0], _p(mm[ci]), _stream()), "sg_minmax")
    L.check(lib.sg_ae_bf16_check(_p(ws), _stream()), "sg_ae_bf16_check")             # synchronises: pipeline time-outs
    if mm is not None and n:
        bad = np.nonzero(~np.isfinite(mm[:nchunks, :2].cpu().numpy()).all(axis=1))[0]
        for ci in bad:
            i0, i1 = int(ci) * chunk, min(n, (int(ci) + 1) * chunk)
            err[i0:i1].copy_(ae_errors(autoencoder, images[i0:i1], device, min(chunk, 2048), conv_mode="fp32"))
    return err


def mean_plus_k_std(values: torch.Tensor, k: float) -> torch.Tensor:
    """1-element fp32 device tensor mean + k * std (unbiased), "#autoencoder.py:320"; fixed-order fp64 sums."""
    device = values.device
    lib = _lib_for(device)
    n = values.numel()
    chunks = (n + L.SG_MOMENT_CHUNK - 1) // L.SG_MOMENT_CHUNK
    part = torch.empty(2 * max(chunks, 1), dtype=torch.float64, device=device)
    thr = torch.empty(1, dtype=torch.float32, device=device)
    L.check(lib.sg_chunk_moments(_p(values), n, _p(part), _stream()), "sg_chunk_moments")
    L.check(lib.sg_moments_finish(_p(part), chunks, n, float(k), L.P(0), _p(thr), _stream()), "sg_moments_finish")
    return thr


def detect_outliers_autoencoder(autoencoder, dataset, device, threshold=2.0, *, conv_mode: str = "auto"):
    """``detect_outliers_autoencoder`` ("#autoencoder.py:307-322"): per-sample reconstruction MSE,
    inlier = error < mean + threshold * std (unbiased).  Returns a CPU torch.BoolTensor [N] like the
    reference (whose errors are ``.cpu()``'d)."""
    device = _dev(device)
    autoencoder.eval()
    err = ae_errors(autoencoder, _dataset_images(dataset), device, conv_mode=conv_mode)
    thr = mean_plus_k_std(err, threshold)
    _, _, mask = compact_indices(err, thr, L.SG_LT, 0, want_mask=True)
    return mask.bool().cpu()


# ---- in-batch strain + concat --------------------------------------------------------------------
def _alias(buf: torch.Tensor, row0: int, rows: int) -> torch.Tensor:
    """rows [row0, row0 + rows) of ``buf`` as a tensor of its own (same memory, no autograd / view relation)."""
    out = torch.empty(0, dtype=buf.dtype, device=buf.device)
    shape = (rows,) + tuple(buf.shape[1:])
    return out.set_(buf.untyped_storage(), buf.storage_offset() + row0 * buf.stride(0), shape, buf.stride())


def strain_scores(real: torch.Tensor, real_scores: torch.Tensor, q: float = 0.1, *, _extra_status=None):
    """Selection half of the in-batch block ("# 상위 10% 제거해서 fake image에 concate.py:246-249"):
    threshold = torch.quantile(scores, q); mask = scores >= threshold; real[mask], real[~mask] in one
    pass.  For one batch (<= 2048 rows) the strained rows ``real[~mask]`` are written straight to the TAIL of a
    pre-sized ``[B, ...]`` buffer -- behind the ``B - s`` rows the generator output will occupy -- so that
    ``concat_fake(fake, filtered_fake)`` (":268") only has to copy the generator rows.
    Returns (filtered_real, filtered_fake, mask, threshold)."""
    device = real.device
    scores = real_scores.reshape(-1).to(torch.float32).contiguous()
    n = scores.numel()
    if 1 <= n <= 2048 and n == real.shape[0]:
        # one batch: sort + torch.quantile's lerp + mask + stable partition ranks in ONE CTA, then the row mover
        lib = _lib_for(device)
        k0, k1, w = _torch_quantile_plan(n, float(q))
        rows = real.contiguous()
        row_bytes = rows[0].numel() * rows.element_size()
        if row_bytes % 16:
            raise ValueError("row size must be a multiple of 16 bytes")
        kept, concat = torch.empty_like(rows), torch.empty_like(rows)
        mask = torch.empty(n, dtype=torch.uint8, device=device)
        thr = torch.empty(1, dtype=torch.float32, device=device)
        counts = torch.empty(2, dtype=torch.int64, device=device)
        ws = _Scratch.get(device, "strain", 8 * n)
        L.check(lib.sg_strain_rows_concat(_p(scores), n, k0, k1, float(w), L.SG_LERP_TORCH, L.SG_GE, _p(rows), row_bytes,
                                          _p(kept), _p(concat), _p(mask), _p(thr), _p(counts), _p(ws), _stream()),
                "sg_strain_rows_concat")
        nk, nd = _read_counts(counts, _extra_status)
        filtered_fake = _alias(concat, nk, nd)
        filtered_fake._sg_concat_buffer = concat      # concat_fake recognises the pre-placed tail
        return kept[:nk], filtered_fake, mask.view(torch.bool), thr[0]
    thr = quantile_device(scores, q)
    _, _, mask = compact_indices(scores, thr, L.SG_GE, 0, want_mask=True)
    kept, dropped, counts = partition_rows(real, mask)
    nk, nd = _read_counts(counts, _extra_status)
    return kept[:nk], dropped[:nd], mask.bool(), thr[0]


_PINNED: dict = {}


def _read_counts(counts: torch.Tensor, extra_status=None):
    """The one synchronisation of the strain block (the reference's boolean indexing has two): the output
    shapes depend on the counts.  Pinned staging + a stream sync instead of a pageable ``.cpu()``.  ``extra_status``:
    a device int32[2] (a scorer's status words) read back in the same synchronisation into ``_PINNED_STATUS``."""
    key = counts.device.index
    h = _PINNED.get(key)
    if h is None:
        h = (torch.empty(2, dtype=torch.int64).pin_memory(), torch.zeros(2, dtype=torch.int32).pin_memory())
        _PINNED[key] = h
    h[0].copy_(counts, non_blocking=True)
    if extra_status is not None:
        h[1].copy_(extra_status, non_blocking=True)
    torch.cuda.current_stream(counts.device).synchronize()
    return int(h[0][0]), int(h[0][1])


_STATUS_WORDS: dict = {}


def _status_words(device) -> torch.Tensor:
    """A per-device int32[2] for the status words of one in-batch scoring call: zeroed when created and again only after a
    call that found it set (the kernels write it on an error only), which saves a memset launch per batch."""
    t = _STATUS_WORDS.get(device.index)
    if t is None:
        t = torch.zeros(2, dtype=torch.int32, device=device)
        _STATUS_WORDS[device.index] = t
    return t


def _last_status(device) -> tuple:
    h = _PINNED[device.index][1]
    return int(h[0]), int(h[1])


def strain_batch(netD, real: torch.Tensor, q: float = 0.1, *, conv_mode: str = "auto"):
    """The in-batch strain block ("# 상위 10% 제거해서 fake image에 concate.py:243-251"):
    real_scores = netD(real) under no_grad, threshold = torch.quantile(scores, q), mask = scores >= thr.
    netD is used AS IS: in train mode BatchNorm normalises with the batch statistics and its running
    statistics are updated; in eval mode (the state every script that ran a
    dataset-scale strain is in, quirk 1) the folded running statistics are used.
    Returns (filtered_real, filtered_fake, mask, threshold); ``filtered_fake`` already sits behind the generator rows of
    the batch ``concat_fake`` assembles (":268").  conv_mode as in ``refine_dataset_by_loss``: 'auto' scores in fp16
    and, if the batch overflowed fp16, once more in fp32-parity arithmetic (the status words travel with the counts)."""
    device = _dev(real.device)
    b = real.shape[0]
    prob = torch.empty(b, dtype=torch.float32, device=device)
    if _is_mlp(netD):
        # 28x28 path: MLP discriminator on flattened images ("# 1,2,8.py:110-128", "Untitled-2.py:79-94")
        if netD.training and any(isinstance(m, nn.Dropout) and m.p > 0 for m in netD.modules()):
            raise NotImplementedError("train-mode Dropout draws from torch's RNG stream and cannot be mirrored; "
                                      "score with netD.eval() (Dropout = identity)")
        sc = get_mlp_scorer(netD, device, max_batch=max(b, 512))
        sc.score_into(_f32c(real.reshape(b, -1), device), None, prob, None)
        return strain_scores(real, prob, q)
    if _is_d28(netD):
        # 28x28 conv path (BASELINE config 1): DCGAN-28 on a tcgen05 GEMM
        if netD.training:
            raise NotImplementedError("the DCGAN-28 scorer folds eval-mode BatchNorm; call netD.eval() before straining "
                                      "(train-mode batch statistics are implemented for the 64x64 discriminator only)")
        sc = get_d28_scorer(netD, device, max_batch=max(b, 512))
        status = torch.zeros(2, dtype=torch.int32, device=device)
        sc.score_into(_f32c(real, device).contiguous(), None, prob, None, status)
        out = strain_scores(real, prob, q, _extra_status=status)
        st0, st1 = _last_status(device)
        if st0:
            _raise_status(st0)
        if st1:
            raise RuntimeError("strainer_b200: non-finite logit in the DCGAN-28 fp16 tensor-core form")
        return out
    sc = get_scorer(netD, device, conv_mode, max_batch=max(b, 512))
    x = real.contiguous()
    train = netD.training
    # train-mode BatchNorm in the fp16 class: the training step's forward (csrc/d64_train.cu: one C call, programmatic
    # dependent launches, weights packed once per optimiser step and shared with the D step that follows)
    trainer = None
    if train and sc.mode_name in ("auto", "fp16") and 2 <= b <= 4096 and x.dtype == torch.float32 and x.is_cuda:
        from .train import trainer_for
        trainer = trainer_for(netD, max(b, 128))
        if not all(t.device == x.device and t.is_contiguous() and t.dtype == torch.float32
                   for t in trainer._params() + trainer._running_stats()):
            trainer = None       # a module living elsewhere (or in another dtype): the scorer below converts and copies back

    def run(scorer, status):
        if train:
            scorer.score_train_into(netD, x, None, prob, None, status)
        else:
            scorer.score_into(x, None, prob, None, status)
        return strain_scores(real, prob, q, _extra_status=status)

    if trainer is not None:
        status = trainer.score_train(x, prob)
        out = strain_scores(real, prob, q, _extra_status=status)
    else:
        status = _status_words(device)      # zero unless the previous call found an error (and cleared it below)
        out = run(sc, status)
    st0, st1 = _last_status(device)
    if st0:
        status.zero_()
        _raise_status(st0)
    if trainer is not None and st1 != _FP16_OVERFLOW:
        trainer.committed()
        return out
    if st1 == _FP16_OVERFLOW:
        status.zero_()          # sticky words (the training workspace's, or the per-device pair): reported here
        status = _status_words(device)
        if sc.mode_name != "auto":
            raise RuntimeError("strainer_b200: non-finite logit in the fp16 conv mode (an activation exceeded 65504); use "
                               "conv_mode='auto', 'fp32' or 'bf16'")
        # the fp16 pass left the running statistics untouched (bn_commit_kernel); score the batch again in fp32 parity
        sc.fallback_chunks += 1
        fb = sc.fallback()
        if fb.max_batch < b:
            fb = sc._fallback = D64Scorer(netD, device, "fp32", max_batch=b)
        status.zero_()
        out = run(fb, status)
        st0, _ = _last_status(device)
        if st0:
            _raise_status(st0)
        sc = fb
    if train:
        sc.commit_train_side_effects()
    return out


class _ConcatFake(torch.autograd.Function):
    """torch.cat([fake, strained], 0) of "# 상위 10% 제거해서 fake image에 concate.py:268" as ONE row-copy launch into a
    pre-sized buffer; the gradient of the result flows to the generator rows only (the strained reals are data)."""

    @staticmethod
    def forward(ctx, fake, strained):
        n1, n2 = fake.shape[0], strained.shape[0]
        ctx.n1 = n1
        device = fake.device
        lib = _lib_for(device)
        f = fake.contiguous()
        row_bytes = (f[0].numel() if n1 else strained[0].numel()) * f.element_size()
        buf = getattr(strained, "_sg_concat_buffer", None)
        in_place = (buf is not None and buf.shape[0] == n1 + n2 and buf.dtype == f.dtype and buf.device == device
                    and tuple(buf.shape[1:]) == tuple(f.shape[1:])
                    and strained.data_ptr() == buf.data_ptr() + n1 * row_bytes)
        if in_place:
            out = _alias(buf, 0, n1 + n2)
            src_b = strained
        else:
            out = torch.empty((n1 + n2,) + tuple(f.shape[1:]), dtype=f.dtype, device=device)
            src_b = strained.contiguous()
        if row_bytes % 16:
            raise ValueError("row size must be a multiple of 16 bytes")
        L.check(lib.sg_concat_rows(_p(f), n1, _p(src_b), n2, row_bytes, _p(out), _stream()), "sg_concat_rows")
        return out

    @staticmethod
    def backward(ctx, grad):
        return grad[:ctx.n1], None


def concat_fake(fake: torch.Tensor, strained: torch.Tensor) -> torch.Tensor:
    """``torch.cat([fake, filtered_fake], dim=0)`` (":268") for the ``filtered_fake`` of ``strain_batch`` /
    ``strain_scores``: those rows were already written behind the ``fake.size(0)`` generator rows of a pre-sized
    ``[B, ...]`` buffer, so only the generator rows are copied (one launch) and the buffer is returned; the backward
    hands ``grad[:fake.size(0)]`` to the generator.  Any other ``strained`` tensor is concatenated by the same kernel
    into a fresh buffer.  Note: the pre-sized buffer is consumed -- call this once per ``strain_batch`` result."""
    if fake.dim() != strained.dim() or tuple(fake.shape[1:]) != tuple(strained.shape[1:]) or fake.dtype != strained.dtype:
        raise ValueError(f"concat_fake: incompatible shapes {tuple(fake.shape)} / {tuple(strained.shape)} or dtypes")
    return _ConcatFake.apply(fake, strained)


def sample_pool(pool: torch.Tensor, b: int, indices: torch.Tensor | None = None) -> torch.Tensor:
    """``potential_fake_data[torch.randperm(P)[:b]]`` ("# strainer gan + concate.py:623-624"): the
    permutation is drawn on the host exactly like the reference (same RNG stream), the 48 KiB-row
    gather runs as one vectorised kernel."""
    device = _dev(pool.device)
    lib = _lib_for(device)
    if indices is None:
        indices = torch.randperm(pool.size(0))[:b]
    idx = indices.to(device=device, dtype=torch.int64)
    pool = pool.contiguous()
    out = torch.empty((idx.numel(),) + tuple(pool.shape[1:]), dtype=pool.dtype, device=device)
    row_bytes = pool[0].numel() * pool.element_size()
    L.check(lib.sg_gather_rows(_p(pool), row_bytes, _p(idx), idx.numel(), L.P(0), _p(out), _stream()), "sg_gather_rows")
    return out


def synth_images(start: int, count: int, seed: int = 999, device=None) -> torch.Tensor:
    """Counter-based synthetic fp32 [count,3,64,64] images generated on the device."""
    device = _dev(device)
    lib = _lib_for(device)
    out = torch.empty((count, 3, 64, 64), dtype=torch.float32, device=device)
    L.check(lib.sg_synth_images(_p(out), start, count, seed, _stream()), "sg_synth_images")
    return out
