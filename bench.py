#!/usr/bin/env python
"""Benchmark of the straining hot path (BASELINE.json metric: strained samples/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config 5 of BASELINE.json, weak scaling): every GPU holds a shard of SHARD = 131072 synthetic
64x64x3 fp32 samples (2^20 samples at 8 GPUs); ONE STEP = one full dataset-scale strain =
  D scoring of every sample (tcgen05 convs + fused sigmoid/BCE head)  ->  global top-10 % cutoff
  (np.percentile(losses, 90) by radix select; integer histograms all-reduced over NCCL when N > 1)
  ->  ascending kept-index compaction (np.where(loss < thr)).
`value` is the whole-job samples/s with the shard resident in HBM; `e2e` is the same strain through the
reference-facing call (refine_dataset_by_loss on a HOST dataset: H2D of every image and D2H of the kept
indices inside the timed region).  Under torchrun one process per GPU; timing = CUDA events, max over ranks.
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
CHUNK = 8192            # samples per scoring launch group
E2E_SAMPLES = 32768     # per-GPU host dataset for the end-to-end leg (1.6 GB pinned)
LOSS_RATIO = 0.1        # "remove top 10 % loss"
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
    "#strainer gan.py:364-392": torch-CPU D forward + np.percentile + np.where) on the host cores."""
    if rank != 0:
        return
    import numpy as np
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


def main():
    global CHUNK
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="conv arithmetic: fp16 (one tensor pass, meets the 1e-3 fp32 bar), bf16 (one pass, 2e-2 bar), "
                         "fp32 (bf16 hi/lo split, three passes)")
    ap.add_argument("--shard", type=int, default=SHARD)
    ap.add_argument("--chunk", type=int, default=CHUNK, help="samples per scoring launch group")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the DCGAN train iters/sec leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    CHUNK = args.chunk

    import numpy as np
    import torch
    import torch.distributed as dist
    import strainer_b200 as sb
    from oracle import strainer_oracle as O

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    pk = peaks()
    shard = args.shard
    n_global = shard * world
    base = rank * shard

    netD = O.make_discriminator(O.SEED)   # reference architecture, weights_init + perturbed BN stats
    netD.eval()
    images = sb.synth_images(base, shard, O.SEED, device)      # resident shard, generated on device
    scorer = sb.D64Scorer(netD, device, args.mode, max_batch=CHUNK)
    losses = torch.empty(shard, dtype=torch.float32, device=device)
    q = (1 - LOSS_RATIO) * 100

    def strain_step(sc):
        for i in range(0, shard, CHUNK):
            b = min(CHUNK, shard - i)
            sc.score_into(images[i:i + b], None, None, losses[i:i + b])
        thr = sb.percentile_device(losses, q, group, n_global)
        idx, count, _ = sb.compact_indices(losses, thr, 0, base)
        return thr, idx, count

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if group is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, out

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, (thr, idx, count) = timed(lambda: strain_step(scorer), args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    scorer.check()
    value = n_global / (ms_step * 1e-3)
    kept = int(count.item())
    nchunks = (shard + CHUNK - 1) // CHUNK
    # kernels per step: 5 per scoring chunk | select: begin + cooperative radix phases (N = 1) or begin + 4 x (hist, step)
    # + finish (N > 1, all-reduce between) | lerp | index compaction
    launches_per_step = nchunks * 5 + (2 if world == 1 else 10) + 1 + 1

    # ---- per-kernel timing of the conv kernels inside a long loop (sustained clocks) ----------------------
    def layer_times(sc, reps):
        xs = images[:CHUNK]
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(reps)]
        for w in range(2):
            for layer in range(1, 6):
                sc.run_layer(xs, layer, None, None, losses[:CHUNK])
        torch.cuda.synchronize()
        for r in range(reps):
            ev[r][0].record()
            for layer in range(1, 6):
                sc.run_layer(images[(r % nchunks) * CHUNK:(r % nchunks) * CHUNK + CHUNK], layer, None, None, losses[:CHUNK])
                ev[r][layer].record()
        torch.cuda.synchronize()
        t = np.array([[ev[r][l].elapsed_time(ev[r][l + 1]) for l in range(5)] for r in range(reps)])
        return t.mean(axis=0)   # ms per launch of [conv1, conv2, conv3, conv4, head] at CHUNK samples

    lt = layer_times(scorer, 24)
    conv_ms = float(lt[1] + lt[2] + lt[3])
    nseg = 3 if args.mode == "fp32" else 1
    achieved_tf = FLOP_CONV * CHUNK / (conv_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "conv2_swap2_kernel + 2 x conv_pair2_kernel (L2+L3+L4 implicit-GEMM launches of one chunk)",
                "achieved": achieved_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved_tf / pk["tf_sust"],
                "peak_source": f"bf16_tflops_sustained of {pk['src']} (cuBLAS bf16 back to back for 4 s: the convs are timed inside a "
                               "long loop under the same sw_power_cap)",
                "frac_of_burst_peak": achieved_tf / pk["tf_burst"], "burst_peak": pk["tf_burst"],
                # dram__bytes_read.sum + dram__bytes_write.sum of the three launches at 8192 samples, bf16 mode, from the
                # ncu --set full captures profiles/r1d_conv_kernels_full.txt (bf16) / r1f_conv_kernels_fp16_full.txt (L2 1.074+0.506,
                # L3 0.538+0.235, L4 0.273+0.107 GB);
                # algorithmic: act1 1.074 + act2 0.537 read, act2 0.537 + act3 0.268 + act4 0.134 written = 2.55 GB
                "traffic": 2.734e9 if (args.mode in ("bf16", "fp16") and CHUNK == 8192) else None,
                "traffic_unit": "bytes per launch group (ncu, profiles/r1d_conv_kernels_full.txt, r1f_conv_kernels_fp16_full.txt)",
                "algorithmic_flops_per_sample": FLOP_CONV, "samples_per_launch_group": CHUNK,
                "issued_tensor_flops_factor": nseg}
    kernels = {"chunk": CHUNK, "ms": {k: float(v) for k, v in zip(["conv1", "conv2", "conv3", "conv4", "head"], lt)},
               "tflops": {k: float(f * CHUNK / (v * 1e-3) / 1e12) for k, f, v in zip(
                   ["conv2", "conv3", "conv4"], [2 * 256 * 128 * 1024, 2 * 64 * 256 * 2048, 2 * 16 * 512 * 4096], lt[1:4])}}

    # selection + compaction kernels alone (HBM-bound in theory; 0.5 MB here => launch/L2 bound)
    def sel_only():
        t = sb.percentile_device(losses, q, group, n_global)
        return sb.compact_indices(losses, t, 0, base)
    for _ in range(5):
        sel_only()
    torch.cuda.synchronize()
    evs = []
    for _ in range(30):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        sel_only()
        b_.record()
        evs.append((a_, b_))
    torch.cuda.synchronize()
    ms_sel = float(np.median([a_.elapsed_time(b_) for a_, b_ in evs]))   # median: robust to a host hiccup in the enqueue loop

    # ---- secondary: the other conv modes on the same step ---------------------------------------------------
    other_modes = {}
    for other in [m for m in ("fp16", "bf16", "fp32") if m != args.mode]:
        sc2 = sb.D64Scorer(netD, device, other, max_batch=CHUNK)
        ms2, _ = timed(lambda: strain_step(sc2), max(2, args.steps // 3), 3)
        sc2.check()
        other_modes[other] = {"value": n_global / (ms2 * 1e-3), "ms_per_step": ms2}
        del sc2

    # ---- end to end through the reference-facing call, host dataset -----------------------------------------
    e2e = None
    e2e_u8 = None
    if not args.no_e2e:
        def e2e_leg(ds, ne):
            def e2e_step():
                return sb.refine_dataset_by_loss(ds, netD, device, LOSS_RATIO, conv_mode=args.mode)
            for _ in range(3):
                sub, _t = e2e_step()
            torch.cuda.synchronize()
            if group is not None:
                dist.barrier()
            es = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(es):
                sub, _t = e2e_step()
            torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / es], device=device)
            if group is not None:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return ne * world / dt.item(), int(len(sub.indices)) * 8 + 12

        ne = min(E2E_SAMPLES, shard)
        host = torch.empty((ne, 3, 64, 64), dtype=torch.float32).pin_memory()
        host.copy_(images[:ne])
        v, d2h = e2e_leg(torch.utils.data.TensorDataset(host, torch.zeros(ne, dtype=torch.long)), ne)
        e2e = {"value": v, "unit": "samples/s", "h2d_bytes_per_step": ne * 49152,
               "d2h_bytes_per_step": d2h, "samples_per_gpu": ne,
               "api": "refine_dataset_by_loss(TensorDataset(host pinned fp32), netD, device, 0.1)",
               "note": "each rank strains its own host dataset (replicas); PCIe H2D of 49152 B/sample is inside the timed region"}
        del host
        # the same call on a uint8 host dataset (the pixels the reference's ImageFolder decodes, ToTensor + Normalize
        # applied on the device, bit-identical to the host transform): 12288 B/sample over PCIe
        ne8 = shard
        host8 = torch.empty((ne8, 3, 64, 64), dtype=torch.uint8).pin_memory()
        for i in range(0, ne8, CHUNK):
            host8[i:i + CHUNK].copy_(((images[i:i + CHUNK] + 1.0) * 127.5).round_().clamp_(0, 255).to(torch.uint8))
        v8, d2h8 = e2e_leg(sb.U8ImageDataset(host8, None, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), ne8)
        e2e_u8 = {"value": v8, "unit": "samples/s", "h2d_bytes_per_step": ne8 * 12288, "d2h_bytes_per_step": d2h8,
                  "samples_per_gpu": ne8,
                  "api": "refine_dataset_by_loss(U8ImageDataset(host pinned uint8, Normalize(.5,.5)), netD, device, 0.1)",
                  "note": "same strain on the uint8 pixels of the same images quantised to 8 bits; ToTensor + Normalize run on "
                          "the device (sg_u8_normalize), 12288 B/sample over PCIe inside the timed region"}
        del host8

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port of the reference on the host cores ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
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
               "sample": f"first {ns} samples of the same stream, oracle refine_dataset_by_loss (torch CPU fp32, bs 64), "
                         f"best of 3 passes ({3 * best:.1f} s of CPU work)"}
        # parity of every conv mode against the oracle's losses on those samples (the oracle as the checker)
        widx, wthr, wloss = O.refine_dataset_by_loss(xs, netD, LOSS_RATIO)
        wloss = wloss.reshape(-1)
        parity = {"samples": ns, "oracle_threshold": float(wthr)}
        for m in ("fp16", "bf16", "fp32"):
            scm = scorer if m == args.mode else sb.D64Scorer(netD, device, m, max_batch=CHUNK)
            lm = scm.score(images[:ns], ("loss",))["loss"]
            tm = sb.percentile_device(lm, q)
            im, cm, _ = sb.compact_indices(lm, tm, 0, 0)
            lm_h = lm.cpu().numpy()
            rel = np.abs(lm_h - wloss) / np.maximum(np.abs(wloss), 1e-6)
            got = np.zeros(ns, bool)
            got[im[:int(cm.item())].cpu().numpy()] = True
            want = np.zeros(ns, bool)
            want[widx] = True
            near = np.abs(wloss - wthr) <= 1e-3 * abs(wthr)
            parity[m] = {"max_rel_loss_err": float(rel.max()), "threshold": float(tm.item()),
                         "mask_disagreements": int((got != want).sum()),
                         "mask_disagreements_outside_1e-3_of_threshold": int(((got != want) & ~near).sum())}
        cpu["parity_vs_oracle"] = parity

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
    # The G/D update is outside the straining path and stays torch autograd in every arm; the arms
    # differ in the in-batch strain block only ("# 상위 10% 제거해서 fake image에 concate.py:243-273").
    train = None
    if rank == 0 and world == 1 and not args.no_train:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import config_bench
        del images, losses
        torch.cuda.empty_cache()
        train = {"batch": 128, "unit": "iters/s",
                 "plain_dcgan_no_strain": config_bench.train_iters(device, 40, "none"),
                 "reference_eager_strain_block_on_gpu": config_bench.train_iters(device, 40, "torch"),
                 "b200_strain_batch_concat_fake": config_bench.train_iters(device, 40, "b200"),
                 "note": "G/D forward + backward + Adam = torch autograd in all arms; only the strain block differs"}

    if rank == 0:
        line = {"metric": "strained_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": {"fp16": "fp16", "bf16": "bf16", "fp32": "bf16x3 (fp32-parity split)"}[args.mode],
                "data": "synthetic",
                "config": {"workload": "C5 dataset-scale D64 scoring + global top-10% radix select + index compaction",
                           "samples_per_gpu": shard, "samples_total": n_global, "chunk": CHUNK, "loss_ratio": LOSS_RATIO,
                           "conv_mode": args.mode, "l2": "inputs (6.4 GB/GPU) larger than L2; no flush needed",
                           "kept": kept, "threshold": float(thr.item())},
                "clocks": clocks, "e2e": e2e, "e2e_u8_dataset": e2e_u8, "gpu_launches": launches_per_step * args.steps,
                "roofline": roofline, "cpu_baseline": cpu,
                "frac_of_conv_roofline": value / world / (pk["tf_sust"] * 1e12 / FLOP_ALL),
                "kernels": kernels, "select_compact_ms": ms_sel, "train_iters_per_sec": train, "torch_eager_gpu": eager,
                "other_modes": other_modes}
        print(json.dumps(line), flush=True)
    if group is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
