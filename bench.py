#!/usr/bin/env python
"""Benchmark of the straining hot path (BASELINE.json metric: strained samples/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config 5 of BASELINE.json, weak scaling): every GPU holds a shard of SHARD = 131072 synthetic
64x64x3 fp32 samples (2^20 samples at 8 GPUs); ONE STEP = one full dataset-scale strain through the repo's public call
with its DEFAULTS (``strain_shard(images, netD, 0.1, group=...)`` = the data-parallel ``refine_dataset_by_loss``):
  D scoring of every sample (tcgen05 convs + fused sigmoid/BCE head, conv_mode 'auto')  ->  global top-10 % cutoff
  (np.percentile(losses, 90) by radix select; integer histograms all-reduced over NCCL when N > 1)
  ->  ascending kept-index compaction (np.where(loss < thr)), kept indices read back to the host.
`value` is the whole-job samples/s with the shard resident in HBM; `e2e` is the same call on a HOST dataset (H2D of
every image and D2H of the kept indices inside the timed region; at N > 1 ONE host dataset split by rank, global
threshold).  Under torchrun one process per GPU; timing = CUDA events, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHARD = 131072          # samples per GPU (6.4 GB fp32): inputs are far larger than the 126 MB L2
CHUNK = 8192            # samples per scoring launch group (the library's DATASET_CHUNK)
LOSS_RATIO = 0.1        # "remove top 10 % loss"
PARITY_TOTAL = 98304    # samples of the multi-GPU parity check (single-GPU strain of the concatenated shards, rank 0)
FLOP_CONV = 2 * (256 * 128 * 1024 + 64 * 256 * 2048 + 16 * 512 * 4096)   # L2..L4 = 201.3 MFLOP / sample
FLOP_ALL = FLOP_CONV + 2 * (1024 * 64 * 48 + 8192)                      # 207.6 MFLOP / sample


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].startswith("Active") for r in self.rows)]
        # under load = the upper half of the samples (the sampler also sees the idle edges)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": (max(mx) if mx else None),
                "reasons": reasons, "samples": len(sm)}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle port of refine_dataset_by_loss,
    "#strainer gan.py:364-392": torch-CPU D forward + np.percentile + np.where) on the host cores.  The reference is 21
    flat scripts that cannot be installed or imported (they load datasets at import time), so the port is what runs."""
    if rank != 0:
        return
    import torch
    from oracle import strainer_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = int(os.environ.get("SG_BENCH_REF_SAMPLES", "4096"))   # bounded sample of the workload per step
    x = torch.from_numpy(O.synth_images(0, sample))
    netD = O.make_discriminator(O.SEED)
    for _ in range(max(args.warmup, 1)):
        O.refine_dataset_by_loss(x[:256], netD, LOSS_RATIO)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        idx, thr, _ = O.refine_dataset_by_loss(x, netD, LOSS_RATIO)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    line = {"impl": "reference", "metric": "strained_samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 dataset-scale D64 scoring + np.percentile(90) + np.where, CPU oracle port of "
                                   "refine_dataset_by_loss", "samples_per_step": sample, "loss_ratio": LOSS_RATIO},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} samples/step, torch CPU {torch.get_num_threads()} threads"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def numa_cpus_of_gpu(torch, index):
    """(numa node, cpu set) of the host memory closest to GPU `index`, from sysfs; (None, None) if unknown."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return node, cpus
    except Exception:
        return None, None


def conv_traffic_from_profile(mode_is_16bit, chunk):
    """dram bytes (read + write) of the L2 + L3 + L4 launches of one 8192-sample chunk, read from the newest committed
    ncu --set full summary (tools/ncu_traffic.py writes profiles/*_conv_traffic.json); None when there is none."""
    import glob
    if not mode_is_16bit or chunk != 8192:
        return None, None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_conv_traffic.json")))
    if not files:
        return None, None
    try:
        d = json.load(open(files[-1]))
        return float(d["l2_l3_l4_dram_bytes"]), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


def main():
    global CHUNK
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fp16", "bf16", "fp32"],
                    help="conv arithmetic of the headline; 'auto' = the library default (no conv_mode argument is passed)")
    ap.add_argument("--shard", type=int, default=SHARD)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the DCGAN train iters/sec leg")
    ap.add_argument("--no-numa", action="store_true", help="do not bind this rank to the NUMA node of its GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import strainer_b200 as sb
    from oracle import strainer_oracle as O
    CHUNK = sb.DATASET_CHUNK

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # pinned host buffers are first-touched by this process: run it on the cores next to its GPU (NUMA-local memory)
    numa_node, numa_cpus = (None, None) if args.no_numa else numa_cpus_of_gpu(torch, local)
    if numa_cpus:
        try:
            os.sched_setaffinity(0, numa_cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
        except Exception:
            numa_node = None
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    pk = peaks()
    shard = args.shard
    n_global = shard * world
    base = rank * shard
    kw = {} if args.mode == "auto" else {"conv_mode": args.mode}     # the headline passes NO conv_mode argument
    lib = sb._lib.load()

    netD = O.make_discriminator(O.SEED)   # reference architecture, weights_init + perturbed BN stats
    netD.eval()
    images = sb.synth_images(base, shard, O.SEED, device)      # resident shard, generated on device
    q = (1 - LOSS_RATIO) * 100

    def strain_step(**mode_kw):
        # the public data-parallel call, defaults only: D scoring -> global percentile -> global kept indices (host)
        return sb.strain_shard(images, netD, LOSS_RATIO, group=group, index_base=base, n_global=n_global, **mode_kw)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.sg_launch_count()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        launches = lib.sg_launch_count() - l0
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if group is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, out, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, (kept_idx, thr, losses), launches = timed(lambda: strain_step(**kw), args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    value = n_global / (ms_step * 1e-3)
    kept = int(len(kept_idx))
    scorer = sb.get_scorer(netD, device, args.mode, CHUNK)
    fallback_chunks = scorer.fallback_chunks
    nchunks = (shard + CHUNK - 1) // CHUNK

    # ---- the same step in the three-pass fp32-parity arithmetic (bf16 hi/lo split), stated next to the headline ------
    ms32, _, _ = timed(lambda: strain_step(conv_mode="fp32"), max(2, args.steps // 3), 3)
    fp32_parity_value = n_global / (ms32 * 1e-3)
    other_modes = {"fp32": {"value": fp32_parity_value, "ms_per_step": ms32}}
    for other in ("bf16",):
        ms2, _, _ = timed(lambda: strain_step(conv_mode=other), max(2, args.steps // 3), 3)
        other_modes[other] = {"value": n_global / (ms2 * 1e-3), "ms_per_step": ms2}
    sb.clear_scorer_caches()           # the fp32 workspaces (4 GB) are not needed below
    scorer = sb.D64Scorer(netD, device, args.mode, max_batch=CHUNK)
    loss_buf = torch.empty(shard, dtype=torch.float32, device=device)

    # ---- per-kernel times INSIDE the real step: one full pass over the shard, an event after every launch; the first two
    #      chunks (clock ramp) are dropped.  These are sustained-clock times: frac is taken against the sustained peak.
    def in_step_layer_times():
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(nchunks)]
        passes = max(3, int(round(600.0 / max(ms_step, 1.0))))   # ~0.6 s of the same launches: the board settles at its
        for rep in range(passes):                                # power-capped clocks; the LAST pass is the one measured
            for ci in range(nchunks):
                xs = images[ci * CHUNK:(ci + 1) * CHUNK]
                ev[ci][0].record()
                for layer in range(1, 6):
                    scorer.run_layer(xs, layer, None, None, loss_buf[ci * CHUNK:ci * CHUNK + xs.shape[0]])
                    ev[ci][layer].record()
        torch.cuda.synchronize()
        use = range(2, nchunks) if nchunks > 4 else range(nchunks)
        t = np.array([[ev[ci][l].elapsed_time(ev[ci][l + 1]) for l in range(5)] for ci in use])
        # median over the chunks: one preempted launch (seen once: a single 3.5 ms conv4 among fourteen of 0.40 ms) must not
        # move the per-kernel figure; the whole-step `value` is measured separately and keeps every such event
        return np.median(t, axis=0), float(np.median([ev[ci][0].elapsed_time(ev[ci][5]) for ci in use]))

    # ---- the same launches timed alone in a short loop (burst clocks): only comparable with the BURST peak
    def short_loop_layer_times(reps):
        xs = images[:CHUNK]
        for _ in range(2):
            for layer in range(1, 6):
                scorer.run_layer(xs, layer, None, None, loss_buf[:CHUNK])
        torch.cuda.synchronize()
        time.sleep(0.5)                 # let the board cool off the previous leg
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(reps)]
        for r in range(reps):
            ev[r][0].record()
            for layer in range(1, 6):
                scorer.run_layer(images[(r % nchunks) * CHUNK:(r % nchunks) * CHUNK + CHUNK], layer, None, None, loss_buf[:CHUNK])
                ev[r][layer].record()
        torch.cuda.synchronize()
        return np.array([[ev[r][l].elapsed_time(ev[r][l + 1]) for l in range(5)] for r in range(reps)]).mean(axis=0)

    lt, chunk_ms = in_step_layer_times()
    lt_short = short_loop_layer_times(8)
    scorer.check()
    conv_ms = float(lt[1] + lt[2] + lt[3])
    conv_ms_short = float(lt_short[1] + lt_short[2] + lt_short[3])
    nseg = 3 if args.mode == "fp32" else 1
    achieved_tf = FLOP_CONV * CHUNK / (conv_ms * 1e-3) / 1e12
    achieved_tf_short = FLOP_CONV * CHUNK / (conv_ms_short * 1e-3) / 1e12
    traffic, traffic_src = conv_traffic_from_profile(args.mode != "fp32", CHUNK)
    roofline = {"bound": "tensor",
                "kernel": "conv2_swap2_kernel + 2 x conv_pair2_kernel (the L2 + L3 + L4 implicit-GEMM launches of one 8192-sample chunk)",
                "achieved": achieved_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved_tf / pk["tf_sust"],
                "how": "CUDA events after every launch of a full pass over the shard (same launches and order as the timed "
                       "step), median over chunks 2.. : sustained clocks, hence the SUSTAINED cuBLAS bf16 peak",
                "peak_source": f"bf16_tflops_sustained of {pk['src']}",
                "short_loop": {"achieved": achieved_tf_short, "peak": pk["tf_burst"], "frac": achieved_tf_short / pk["tf_burst"],
                               "how": "8 chunks timed right after a pause (burst clocks) against the BURST cuBLAS peak"},
                "whole_step_frac_of_sustained_peak": value / world / (pk["tf_sust"] * 1e12 / FLOP_ALL),
                "traffic": traffic, "traffic_source": traffic_src,
                "traffic_unit": "dram bytes read + written by the three launches of one chunk (ncu --set full)",
                "algorithmic_bytes": {"act1_read": CHUNK * 131072, "act2_write_read": 2 * CHUNK * 65536,
                                      "act3_write_read": 2 * CHUNK * 32768, "act4_write": CHUNK * 16384},
                "algorithmic_flops_per_sample": FLOP_CONV, "samples_per_launch_group": CHUNK,
                "issued_tensor_flops_factor": nseg}
    names = ["conv1", "conv2", "conv3", "conv4", "head"]
    kernels = {"chunk": CHUNK, "in_step_ms": {k: float(v) for k, v in zip(names, lt)},
               "in_step_chunk_ms": chunk_ms, "in_step_sum_ms": float(lt.sum()),
               "short_loop_ms": {k: float(v) for k, v in zip(names, lt_short)},
               "in_step_tflops": {k: float(f * CHUNK / (v * 1e-3) / 1e12) for k, f, v in zip(
                   ["conv2", "conv3", "conv4"], [2 * 256 * 128 * 1024, 2 * 64 * 256 * 2048, 2 * 16 * 512 * 4096], lt[1:4])},
               "conv1_hbm_gbs": float(CHUNK * (49152 + 131072) / (lt[0] * 1e-3) / 1e9)}

    # selection + compaction kernels alone (HBM-bound in theory; 0.5 MB here => launch/L2 bound)
    # the step's own path: with a group, the radix histograms are reduced over NVLink peer memory inside the select's
    # kernels (sb.PeerComm) when the GPUs can map each other, else by NCCL all-reduces; the NCCL form is timed beside it
    peer = sb.PeerComm.for_group(group, device) if group is not None else None

    def time_select(comm):
        def sel_only():
            t = sb.percentile_device(losses, q, group, n_global, comm=comm)
            return sb.compact_indices(losses, t, 0, base)
        for _ in range(5):
            sel_only()
        torch.cuda.synchronize()
        if group is not None:
            dist.barrier()
        evs = []
        for _ in range(30):
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            sel_only()
            b_.record()
            evs.append((a_, b_))
        torch.cuda.synchronize()
        return float(np.median([a_.elapsed_time(b_) for a_, b_ in evs]))   # median: robust to a host hiccup in the enqueue loop
    ms_sel = time_select(peer)
    ms_sel_nccl = time_select(None) if peer is not None else None

    # ---- multi-GPU parity (outside every timed region): the sharded strain of PARITY_TOTAL samples against rank 0's
    #      single-GPU strain of the concatenated shards -- threshold bits and kept-index list must be identical ---------
    parity = None
    if world > 1:
        per = (PARITY_TOTAL // world) // 4096 * 4096
        pidx, pthr, _ = sb.strain_shard(images[:per], netD, LOSS_RATIO, group=group, index_base=rank * per,
                                        n_global=per * world, **kw)
        gathered = [None] * world
        dist.all_gather_object(gathered, (np.asarray(pidx), np.float32(pthr).tobytes()))
        if rank == 0:
            whole = torch.cat([sb.synth_images(r * shard, per, O.SEED, device) for r in range(world)])
            # global index of sample j of rank r in the parity run is r*per + j
            sidx, sthr, _ = sb.strain_shard(whole, netD, LOSS_RATIO, **kw)
            multi = np.concatenate([g[0] for g in gathered])
            parity = {"n": per * world, "threshold_equal": all(g[1] == np.float32(sthr).tobytes() for g in gathered),
                      "kept_indices_equal": bool(np.array_equal(multi, sidx)), "kept": int(len(sidx)),
                      "threshold": float(sthr)}
            del whole
        dist.barrier()

    # ---- H2D micro-benchmark: one cudaMemcpyAsync of a chunk from pinned memory, all ranks at once ----------------------
    h2d = None
    if not args.no_e2e:
        hb = torch.empty((CHUNK, 3, 64, 64), dtype=torch.float32).pin_memory()
        hb.normal_()
        db = torch.empty_like(hb, device=device)
        for _ in range(2):
            db.copy_(hb, non_blocking=True)
        torch.cuda.synchronize()
        if group is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(12):
            db.copy_(hb, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = torch.tensor([12 * hb.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9], device=device)
        allg = [torch.zeros_like(gbs) for _ in range(world)]
        if group is not None:
            dist.all_gather(allg, gbs)
        else:
            allg = [gbs]
        per_gpu = [float(t.item()) for t in allg]
        h2d = {"gbs_per_gpu": per_gpu, "gbs_aggregate": float(sum(per_gpu)), "bytes_per_copy": hb.numel() * 4,
               "numa_node": numa_node, "numa_bound": bool(numa_cpus),
               "how": "12 x cudaMemcpyAsync of one 8192-image chunk (403 MB) from pinned host memory, every rank at the same "
                      "time; the pinned buffer is first-touched on the cores of the GPU's NUMA node"}
        del hb, db

    # ---- end to end through the reference-facing call, host dataset -----------------------------------------
    e2e = None
    e2e_u8 = None
    if not args.no_e2e:
        def e2e_leg(fn, ne):
            for _ in range(3):
                out = fn()
            torch.cuda.synchronize()
            if group is not None:
                dist.barrier()
            es = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(es):
                out = fn()
            torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / es], device=device)
            if group is not None:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return ne * world / dt.item(), out

        import psutil
        ne = shard
        while ne > 8192 and ne * 49152 * max(world, 1) > 0.5 * psutil.virtual_memory().available:
            ne //= 2
        if group is not None:       # the ranks read "available" at different moments: agree on the smallest shard
            ne_t = torch.tensor([ne], dtype=torch.int64, device=device)
            dist.all_reduce(ne_t, op=dist.ReduceOp.MIN)
            ne = int(ne_t.item())
        host = torch.empty((ne, 3, 64, 64), dtype=torch.float32).pin_memory()
        for i in range(0, ne, CHUNK):
            host[i:i + CHUNK].copy_(images[i:i + CHUNK])
        # the scorer the API calls below resolve to (cached per module / device / mode / chunk): its h2d_bytes counter is
        # the bytes it really copied; host_pack = False on it gives the plain fp32 copy for comparison
        e2e_scorer = sb.scorer_for(netD, device, args.mode, sb.api._chunk_for(ne))
        if world == 1:
            ds = torch.utils.data.TensorDataset(host, torch.zeros(ne, dtype=torch.long))
            run_e2e = lambda: sb.refine_dataset_by_loss(ds, netD, device, LOSS_RATIO, **kw)   # noqa: E731
            api = "refine_dataset_by_loss(TensorDataset(host pinned fp32), netD, device, 0.1)"
            note = "the PCIe copies of the dataset are inside the timed region"
        else:
            # ONE host dataset of ne*world samples split by rank: rank r holds samples [r*ne, (r+1)*ne) of it; the
            # threshold is the GLOBAL percentile, the kept indices are global (strain_shard = sharded refine_dataset_by_loss)
            run_e2e = lambda: sb.strain_shard(host, netD, LOSS_RATIO, group=group, index_base=rank * ne,   # noqa: E731
                                              n_global=ne * world, device=device, **kw)
            api = "strain_shard(host pinned fp32 shard, netD, 0.1, group=WORLD, index_base=rank*n, n_global=N*n)"
            note = ("one host dataset split by rank, global threshold (radix histograms reduced over NVLink peer memory, NCCL as the fallback) and global kept indices; "
                    "the PCIe copies of the shard are inside the timed region")

        def fp32_leg(host_pack):
            hp0 = getattr(e2e_scorer, "host_pack", None)
            e2e_scorer.host_pack = host_pack
            tuner = sb.api._PackTuner.get(device)
            if group is not None:
                dist.barrier()                          # the ranks pinned their host shards at different speeds: start together
            for _ in range(8):                          # untimed: pins the staging buffers, then lets the tuner settle its share
                run_e2e()
                done = host_pack is False or tuner.locked
                if group is not None:
                    # every run_e2e() holds a collective (the global select): all ranks must make the SAME number of calls.
                    # The tuners time their own rank and lock after different numbers of calls (an 8-GPU run deadlocked here:
                    # some ranks were already in e2e_leg's barrier while others still tuned), so the ranks agree on "done".
                    flag = torch.tensor([1 if done else 0], dtype=torch.int32, device=device)
                    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                    done = bool(flag.item())
                if done:
                    break
            b0 = e2e_scorer.h2d_bytes
            v_, out_ = e2e_leg(run_e2e, ne)
            es_ = 3 + max(3, min(args.steps, 10))       # e2e_leg: 3 warm-up + the timed calls
            moved = (e2e_scorer.h2d_bytes - b0) // es_
            e2e_scorer.host_pack = hp0
            kept_ = out_[0].indices if world == 1 else out_[0]
            return v_, int(moved), int(len(kept_)) * 8 + 4, float(e2e_scorer.last_pack_fraction), kept_

        v, moved, d2h, share, kept_a = fp32_leg("auto")       # the library default
        v_raw, moved_raw, _d, _s, kept_b = fp32_leg(False)    # every byte as fp32
        same_kept = bool(len(kept_a) == len(kept_b) and np.array_equal(np.asarray(kept_a), np.asarray(kept_b)))
        e2e = {"value": v, "unit": "samples/s", "h2d_bytes_per_step": moved, "d2h_bytes_per_step": d2h,
               "samples_per_gpu": ne, "api": api, "note": note, "h2d_gbs_per_gpu_implied": v / world * (moved / ne) / 1e9,
               "host_pack": {"share_of_rows_sent_as_fp16": share, "host_threads_per_rank": int(sb.api._host_threads()),
                             "what": "library default for fp32 HOST datasets in the fp16 conv mode: the host threads round that "
                                     "share of every chunk to fp16 (the rounding conv1 applies to its input anyway) before the "
                                     "PCIe copy, sg_f16_expand widens it on the device; scores bit-identical to the fp32 copy "
                                     "(tests/test_host_pack.py); the share is measured once (conversion rate vs PCIe rate)",
                             "kept_indices_equal_to_fp32_copy": same_kept},
               "fp32_copy": {"value": v_raw, "h2d_bytes_per_step": moved_raw,
                             "note": "the same call with scorer.host_pack = False: 49152 B/sample over PCIe"}}
        if h2d is not None:
            # what bounds the plain leg: the rate at which this box feeds its GPUs from pinned host memory when all ranks copy
            # at once (max-over-ranks timing: the slowest rank sets the step)
            slow = min(h2d["gbs_per_gpu"])
            e2e["fp32_copy"]["bound"] = {"by": "host-to-device copies (box limit, measured by h2d_microbench with every rank copying at once)",
                                         "slowest_rank_h2d_gbs": slow, "aggregate_h2d_gbs": h2d["gbs_aggregate"],
                                         "frac_of_h2d_limit": (v_raw / world * 49152 / 1e9) / slow}
            e2e["bound"] = {"by": "the host threads' fp32 -> fp16 conversion (host memory bandwidth: source read + staging write + DMA read) beside the PCIe copies", "slowest_rank_h2d_gbs": slow,
                            "pcie_busy_frac": (v / world * (moved / ne) / 1e9) / slow}
        del host
        # the same call on a uint8 host dataset (the pixels the reference's ImageFolder decodes, ToTensor + Normalize
        # applied on the device, bit-identical to the host transform): 12288 B/sample over PCIe
        ne8 = shard
        host8 = torch.empty((ne8, 3, 64, 64), dtype=torch.uint8).pin_memory()
        for i in range(0, ne8, CHUNK):
            host8[i:i + CHUNK].copy_(((images[i:i + CHUNK] + 1.0) * 127.5).round_().clamp_(0, 255).to(torch.uint8))
        u8 = sb.U8ImageDataset(host8, None, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5))
        if world == 1:
            v8, (sub8, _t) = e2e_leg(lambda: sb.refine_dataset_by_loss(u8, netD, device, LOSS_RATIO, **kw), ne8)
            d2h8 = int(len(sub8.indices)) * 8 + 4
        else:
            v8, (idx8, _t, _l) = e2e_leg(lambda: sb.strain_shard(u8.images, netD, LOSS_RATIO, group=group, index_base=rank * ne8,
                                                                 n_global=ne8 * world, device=device, **kw), ne8)
            d2h8 = int(len(idx8)) * 8 + 4
        e2e_u8 = {"value": v8, "unit": "samples/s", "h2d_bytes_per_step": ne8 * 12288, "d2h_bytes_per_step": d2h8,
                  "samples_per_gpu": ne8,
                  "api": "the same call on U8ImageDataset(host pinned uint8, Normalize(.5,.5))",
                  "note": "same strain on the uint8 pixels of the same images quantised to 8 bits; ToTensor + Normalize run on "
                          "the device (sg_u8_normalize), 12288 B/sample over PCIe inside the timed region"}
        del host8, u8

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port of the reference on the host cores ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if numa_cpus:
            try:
                os.sched_setaffinity(0, range(os.cpu_count() or 1))     # the CPU leg may use every core again
            except Exception:
                pass
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ns = 16384
        xs = torch.from_numpy(O.synth_images(0, ns))
        O.refine_dataset_by_loss(xs[:256], netD, LOSS_RATIO)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            O.refine_dataset_by_loss(xs, netD, LOSS_RATIO)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        cpu = {"value": ns / best, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"first {ns} samples of the same stream, oracle refine_dataset_by_loss (torch CPU fp32, bs 64, no "
                         f"DataLoader), best of 3 passes ({3 * best:.1f} s of CPU work)"}
        # parity of every conv mode against the oracle's losses on those samples (the oracle as the checker)
        widx, wthr, wloss = O.refine_dataset_by_loss(xs, netD, LOSS_RATIO)
        wloss = wloss.reshape(-1)
        parity_modes = {"samples": ns, "oracle_threshold": float(wthr)}
        for m in ("auto", "bf16", "fp32"):
            im, tm, lm = sb.strain_shard(images[:ns], netD, LOSS_RATIO, conv_mode=m)
            lm_h = lm.cpu().numpy()
            rel = np.abs(lm_h - wloss) / np.maximum(np.abs(wloss), 1e-6)
            got = np.zeros(ns, bool)
            got[im] = True
            want = np.zeros(ns, bool)
            want[widx] = True
            near = np.abs(wloss - wthr) <= 1e-3 * abs(wthr)
            parity_modes[m] = {"max_rel_loss_err": float(rel.max()), "threshold": float(tm),
                               "mask_disagreements": int((got != want).sum()),
                               "mask_disagreements_outside_1e-3_of_threshold": int(((got != want) & ~near).sum())}
        cpu["parity_vs_oracle"] = parity_modes

    # ---- "existing Blackwell kernels" bar: the reference's own torch ops on this GPU (eager, cuDNN) ----------
    eager = None
    if rank == 0 and world == 1 and not args.no_train:
        import copy
        netDg = copy.deepcopy(netD).to(device).eval()

        def eager_scores(bs, nsamp):
            with torch.no_grad():
                out = []
                for i in range(0, nsamp, bs):
                    p_ = netDg(images[i:i + bs]).view(-1)
                    out.append(torch.nn.functional.binary_cross_entropy(p_, torch.ones_like(p_), reduction="none"))
                return torch.cat(out)
        eager = {"unit": "samples/s", "note": "torch eager D forward + BCE on cuda (scoring only, no select/compact), fp32 weights, "
                                              "torch defaults (cuDNN, TF32 convs allowed)"}
        for bs, nsamp in ((64, 8192), (4096, 32768)):
            eager_scores(bs, min(nsamp, 4 * bs))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eager_scores(bs, nsamp)
            torch.cuda.synchronize()
            eager[f"batch_{bs}"] = nsamp / (time.perf_counter() - t0)
        del netDg

    # ---- second headline metric of BASELINE.json: DCGAN 64x64 train iters/sec (rank 0, N = 1) ------------------------
    # The generator and the optimisers stay torch autograd in every arm; the arms differ in the in-batch strain block
    # ("# 상위 10% 제거해서 fake image에 concate.py:243-273") and in who runs the discriminator's forward / backward
    # (*_d_on_tcgen05 = sb.accelerate_discriminator, csrc/d64_train.cu).
    train = None
    if rank == 0 and world == 1 and not args.no_train:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import config_bench
        del images, losses, loss_buf
        sb.clear_scorer_caches()
        torch.cuda.empty_cache()
        train = {"batch": 128, "unit": "iters/s",
                 "plain_dcgan_no_strain": config_bench.train_iters(device, 40, "none"),
                 "reference_eager_strain_block_on_gpu": config_bench.train_iters(device, 40, "torch"),
                 "b200_strain_batch_concat_fake": config_bench.train_iters(device, 40, "b200"),
                 "plain_dcgan_d_on_tcgen05": config_bench.train_iters(device, 40, "none_train"),
                 "b200_strain_and_d_on_tcgen05": config_bench.train_iters(device, 40, "b200_train"),
                 "note": "generator forward / backward and Adam = torch autograd in all arms; *_d_on_tcgen05: the discriminator's "
                         "forward + backward of the D and G steps run on this repo's kernels (sb.accelerate_discriminator)"}

    # ---- the other BASELINE.json configs on this box (rank 0, N = 1): parity-test cases, reported beside the headline -----
    other_configs = None
    if rank == 0 and world == 1 and not args.no_train:
        other_configs = {}
        # config 4: auto-encoder reconstruction-error scoring, 65 536 resident images in 8 192-image launches
        torch.manual_seed(O.SEED)
        ae = O.AutoEncoder().eval()
        na = 65536
        xa = sb.synth_images(0, na, O.SEED, device)
        ref_e = O.ae_errors(ae, xa[:64].cpu()).numpy()
        c4 = {"unit": "samples/s", "images": na, "flop_per_sample": 46.6e6,
              "what": "sb.ae_errors(ae, images_on_device, chunk=8192): six tcgen05 conv layers + tanh + per-sample MSE"}
        for m in ("auto", "bf16"):
            e_ = sb.ae_errors(ae, xa, device, chunk=8192, conv_mode=m)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(5):
                e_ = sb.ae_errors(ae, xa, device, chunk=8192, conv_mode=m)
            ev1.record()
            torch.cuda.synchronize()
            ms_ = ev0.elapsed_time(ev1) / 5
            rel_ = float((np.abs(e_[:64].cpu().numpy() - ref_e) / np.maximum(ref_e, 1e-6)).max())
            c4[m] = {"value": na / (ms_ * 1e-3), "ms": ms_, "tflops": na * 46.6e6 / (ms_ * 1e-3) / 1e12,
                     "max_rel_err_vs_oracle_64": rel_}
        other_configs["config4_autoencoder"] = c4
        del xa
        # config 2: the in-batch strain block at B = 128 (eval-mode and train-mode BatchNorm)
        real_b = sb.synth_images(0, 128, O.SEED, device)
        c2 = {"unit": "us per batch", "batch": 128, "what": "sb.strain_batch(netD, real, 0.1): scoring + quantile + mask + "
              "partition, incl. the host read of the kept count"}
        for label, train_mode in (("eval_bn", False), ("train_bn", True)):
            nd_ = O.make_discriminator(O.SEED).to(device)
            nd_.train(train_mode)
            for _ in range(10):
                sb.strain_batch(nd_, real_b, LOSS_RATIO)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(100):
                sb.strain_batch(nd_, real_b, LOSS_RATIO)
            torch.cuda.synchronize()
            c2[label] = (time.perf_counter() - t0) / 100 * 1e6
        other_configs["config2_strain_block"] = c2

    if rank == 0:
        line = {"metric": "strained_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"auto": "fp16 (fp32 accumulate; overflowed chunks re-scored in bf16x3)", "fp16": "fp16", "bf16": "bf16",
                          "fp32": "bf16x3 (fp32-parity split)"}[args.mode],
                "data": "synthetic",
                "config": {"workload": "C5 dataset-scale D64 scoring + global top-10% radix select + index compaction",
                           "api": "strain_shard(images, netD, 0.1, group, index_base, n_global) -- library defaults, no conv_mode",
                           "samples_per_gpu": shard, "samples_total": n_global, "chunk": CHUNK, "loss_ratio": LOSS_RATIO,
                           "conv_mode": args.mode, "l2": "inputs (6.4 GB/GPU) larger than L2; no flush needed",
                           "kept": kept, "threshold": float(thr), "fp16_chunks_rescored": fallback_chunks},
                "clocks": clocks, "e2e": e2e, "e2e_u8_dataset": e2e_u8, "gpu_launches": int(launches),
                "gpu_launches_note": "sg_launch_count() difference over the timed region on rank 0 (every launch site of the library)",
                "fp32_parity_value": fp32_parity_value,
                "roofline": roofline, "cpu_baseline": cpu, "multi_gpu_parity": parity, "h2d_microbench": h2d,
                "kernels": kernels, "select_compact_ms": ms_sel,
                "select_compact": {"ms": ms_sel, "reduction": ("nvlink peer memory inside the select kernels (PeerComm)" if peer is not None
                                                               else ("nccl all-reduces" if group is not None else "single device")),
                                   "ms_with_nccl_all_reduces": ms_sel_nccl}, "train_iters_per_sec": train, "torch_eager_gpu": eager,
                "other_modes": other_modes, "other_configs": other_configs}
        print(json.dumps(line), flush=True)
    if group is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
