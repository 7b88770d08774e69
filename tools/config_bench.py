"""Per-config timings of the in-batch straining paths (BASELINE.json configs 1-4) and the second headline
metric, DCGAN 64x64 train iters/sec with the strain block in the loop.  One JSON object on stdout.

  python tools/config_bench.py [--iters 30] [--no-cpu]

Every GPU number is the repo's public API (strain_batch / concat_fake / detect_outliers_* ...) on cuda:0,
timed with CUDA events after warm-up; "cpu" is the oracle restatement of the reference block on the host
cores (torch CPU, all threads) at the same shapes.  The G/D update of the training loop is NOT part of the
straining path: it stays torch autograd on the GPU in both arms of the iters/sec figure;
what changes between the arms is the strain block (reference: eager torch ops + torch.quantile + boolean
indexing; here: strain_batch + concat_fake on the CUDA kernels).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402


def gpu_time(fn, iters, warm=5):
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


def cpu_time(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters


class Generator(nn.Module):
    """DCGAN generator of the reference scripts ("#strainer gan.py:195-225": nz 100, ngf 64, 3x64x64)."""

    def __init__(self, nz=100, ngf=64, nc=3):
        super().__init__()
        self.main = nn.Sequential(
            nn.ConvTranspose2d(nz, ngf * 8, 4, 1, 0, bias=False), nn.BatchNorm2d(ngf * 8), nn.ReLU(True),
            nn.ConvTranspose2d(ngf * 8, ngf * 4, 4, 2, 1, bias=False), nn.BatchNorm2d(ngf * 4), nn.ReLU(True),
            nn.ConvTranspose2d(ngf * 4, ngf * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ngf * 2), nn.ReLU(True),
            nn.ConvTranspose2d(ngf * 2, ngf, 4, 2, 1, bias=False), nn.BatchNorm2d(ngf), nn.ReLU(True),
            nn.ConvTranspose2d(ngf, nc, 4, 2, 1, bias=False), nn.Tanh())

    def forward(self, x):
        return self.main(x)


def train_iters(dev, iters, mode):
    """One DCGAN iteration of "# 상위 10% 제거해서 fake image에 concate.py:236-300" at B = 128:
    strain block -> D step on (filtered real, fake + strained) -> G step.  mode: 'none' (plain DCGAN,
    "#%basic.py:233-305"), 'torch' (the reference's eager strain block on the GPU), 'b200' (this repo's strain block),
    'b200_train' (strain block + the discriminator's forward / backward of the D and G steps on this repo's kernels,
    sb.accelerate_discriminator), 'none_train' (plain DCGAN with the accelerated discriminator)."""
    torch.manual_seed(999)
    B, nz = 128, 100
    netD = O.make_discriminator(O.SEED).to(dev).train()
    netG = Generator().to(dev)
    netG.apply(O.weights_init)
    optD = torch.optim.Adam(netD.parameters(), lr=2e-4, betas=(0.5, 0.999))
    optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.5, 0.999))
    crit = nn.BCELoss()
    data = sb.synth_images(0, B * 8, O.SEED, dev)
    state = {"i": 0}
    accel = mode in ("b200_train", "none_train")
    D = sb.accelerate_discriminator(netD, max_batch=B) if accel else netD
    kw = {"param_grads": False} if accel else {}

    def step():
        real = data[(state["i"] % 8) * B:(state["i"] % 8 + 1) * B]
        state["i"] += 1
        if mode in ("none", "none_train"):
            freal, ffake = real, real[:0]
        elif mode == "torch":
            with torch.no_grad():
                s = netD(real).view(-1)
                thr = torch.quantile(s, 0.1)
                m = s >= thr
                freal, ffake = real[m], real[~m]
        else:
            freal, ffake, _, _ = sb.strain_batch(netD, real, 0.1)          # library defaults
        netD.zero_grad()
        out = D(freal).view(-1)
        errD_real = crit(out, torch.ones_like(out))
        errD_real.backward()
        noise = torch.randn(freal.shape[0], nz, 1, 1, device=dev)
        fake = netG(noise)
        # ":268": fake = cat([fake, filtered_fake]); D sees fake.detach(), the G step the concatenated batch itself
        if mode in ("b200", "b200_train"):
            fake = sb.concat_fake(fake, ffake)
        elif mode == "torch":
            fake = torch.cat([fake, ffake], dim=0)
        out = D(fake.detach()).view(-1)
        errD_fake = crit(out, torch.zeros_like(out))
        errD_fake.backward()
        optD.step()
        netG.zero_grad()
        out = D(fake, **kw).view(-1)      # the G step discards D's parameter gradients (next netD.zero_grad())
        errG = crit(out, torch.ones_like(out))
        errG.backward()
        optG.step()

    t = gpu_time(step, iters, warm=5)
    return 1.0 / t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {"cpu_cores": cores, "iters": a.iters, "configs": {}}
    rng = np.random.default_rng(7)

    # ---- C1: 28x28, batch 64, MLP discriminator, top-10 % by in-batch quantile -------------------------------
    mlp = O.MLPDiscriminator().eval()
    x28 = torch.from_numpy(rng.uniform(-1, 1, (64, 1, 28, 28)).astype(np.float32))
    x28d = x28.to(dev)
    mlpd = O.MLPDiscriminator().eval()
    mlpd.load_state_dict(mlp.state_dict())
    t = gpu_time(lambda: sb.strain_batch(mlpd, x28d, 0.1), a.iters)
    c1 = {"batch": 64, "gpu_us_per_batch": t * 1e6, "gpu_samples_per_s": 64 / t}
    msc = sb.get_mlp_scorer(mlpd, dev, max_batch=512)
    pr = torch.empty(64, device=dev)
    xf = x28d.reshape(64, -1)
    c1["gpu_us_mlp_scoring_only"] = gpu_time(lambda: msc.score_into(xf, None, pr, None), 200, 20) * 1e6
    c1["gpu_us_strain_scores_only"] = gpu_time(lambda: sb.strain_scores(x28d, pr, 0.1), 200, 20) * 1e6
    c1["gpu_us_per_batch_200_iters"] = gpu_time(lambda: sb.strain_batch(mlpd, x28d, 0.1), 200, 20) * 1e6
    if not a.no_cpu:
        tc = cpu_time(lambda: O.strain_batch(mlp, x28.reshape(64, -1), 0.1), 20, 3)
        c1.update(cpu_us_per_batch=tc * 1e6, cpu_samples_per_s=64 / tc)
    out["configs"]["C1_mlp28_b64_top10"] = c1

    # ---- C2: 64x64 RGB, batch 128: in-batch strain + concat; feature z-score + elbow on [65536,512] -----------
    netD = O.make_discriminator(O.SEED).eval()
    for bs, name in ((128, "C2_d64_b128_strain_concat"), (256, "C3_d64_b256_strain_concat"), (512, "C4_d64_b512_strain_concat")):
        real = sb.synth_images(0, bs, O.SEED, dev)
        fake = torch.randn(bs, 3, 64, 64, device=dev)

        def blk():
            fr, ff, _, _ = sb.strain_batch(netD, real, 0.1, conv_mode="bf16")
            return sb.concat_fake(fake[:fr.shape[0]], ff)
        t = gpu_time(blk, a.iters)
        c = {"batch": bs, "gpu_us_per_batch": t * 1e6, "gpu_samples_per_s": bs / t, "conv_mode": "bf16"}
        t32 = gpu_time(lambda: sb.strain_batch(netD, real, 0.1, conv_mode="fp32"), a.iters)
        c["gpu_us_per_batch_fp32_mode"] = t32 * 1e6
        if not a.no_cpu:
            rc, fc = real.cpu(), fake.cpu()

            def cblk():
                fr, ff, _, _, _ = O.strain_batch(netD, rc, 0.1)
                return O.concat_fake(fc[:fr.shape[0]], ff)
            tc = cpu_time(cblk, 3, 1)
            c.update(cpu_us_per_batch=tc * 1e6, cpu_samples_per_s=bs / tc)
        out["configs"][name] = c

    nfeat = 65536
    feats = torch.from_numpy(O.synth_features(nfeat))
    fd = feats.to(dev)

    def elbow():
        z = sb.zscore_max(fd)
        return sb.find_elbow_threshold(z)
    t = gpu_time(elbow, a.iters)
    c = {"rows": nfeat, "gpu_ms": t * 1e3, "gpu_rows_per_s": nfeat / t}
    if not a.no_cpu:
        tc = cpu_time(lambda: O.detect_outliers_elbow(feats), 3, 1)
        c.update(cpu_ms=tc * 1e3, cpu_rows_per_s=nfeat / tc)
    out["configs"]["C2_zscore_elbow_65536x512"] = c

    # ---- C3: z-score + 1-D DBSCAN clean ratio + quantile(max_z, ratio) ---------------------------------------
    def dbs():
        z = sb.zscore_max(fd)
        r = sb.dbscan1d_clean_ratio(z, 0.05, 3)
        return sb.quantile_device(z, float(r))
    t = gpu_time(dbs, a.iters)
    c = {"rows": nfeat, "gpu_ms": t * 1e3, "gpu_rows_per_s": nfeat / t}
    if not a.no_cpu:
        def cdbs():
            z = O.zscore_max_torch(feats).numpy()
            r = O.dbscan1d_clean_ratio(z, 0.05, 3)
            return np.quantile(z, r)
        tc = cpu_time(cdbs, 2, 1)
        c.update(cpu_ms=tc * 1e3, cpu_rows_per_s=nfeat / tc)
    out["configs"]["C3_zscore_dbscan1d_65536x512"] = c

    # 512-d DBSCAN clean ratio (the reference's estimate_ratio_dbscan) on the tensor cores
    for nd in (8192, 65536):
        fdd = fd[:nd].contiguous()
        t = gpu_time(lambda: sb.dbscan_clean_ratio(fdd, 28.0, 3), 3, 1)
        c = {"rows": nd, "d": 512, "gpu_ms": t * 1e3, "pair_gemm_tflops": 2 * 3 * 2.0 * nd * nd * 512 / t / 1e12}
        if not a.no_cpu and nd == 8192:
            tc = cpu_time(lambda: O.estimate_ratio_dbscan_features(feats[:nd].numpy(), 28.0, 3), 1, 0)
            c.update(cpu_ms=tc * 1e3)
        out["configs"][f"C3_dbscan512_clean_ratio_{nd}"] = c

    # ---- C4: auto-encoder reconstruction straining, batch 512 ---------------------------------------------------
    torch.manual_seed(3)
    ae = O.AutoEncoder().eval()
    imgs = sb.synth_images(0, 4096, O.SEED, dev)
    ds = torch.utils.data.TensorDataset(imgs, torch.zeros(4096, dtype=torch.long, device=dev))
    t = gpu_time(lambda: sb.detect_outliers_autoencoder(ae, ds, dev, 2.0, conv_mode="bf16"), max(3, a.iters // 3), 2)
    c = {"samples": 4096, "conv_mode": "bf16", "gpu_ms": t * 1e3, "gpu_samples_per_s": 4096 / t,
         "gflops_per_s": 46.6e-3 * 4096 / t}
    tb = gpu_time(lambda: sb.ae_errors(ae, imgs[:512], dev, conv_mode="bf16"), a.iters)
    c["gpu_us_per_batch512_scoring_only"] = tb * 1e6
    big = sb.synth_images(0, 32768, O.SEED, dev)
    tb = gpu_time(lambda: sb.ae_errors(ae, big, dev, chunk=8192, conv_mode="bf16"), 5, 2)
    c["gpu_samples_per_s_32768_resident"] = 32768 / tb
    t32 = gpu_time(lambda: sb.ae_errors(ae, big, dev, chunk=8192, conv_mode="fp32"), 3, 1)
    c["gpu_samples_per_s_fp32_parity_mode_tensor_cores"] = 32768 / t32
    t16 = gpu_time(lambda: sb.ae_errors(ae, big, dev, chunk=8192, conv_mode="fp16"), 5, 2)
    c["gpu_samples_per_s_fp16_mode_32768_resident"] = 32768 / t16
    e_16 = sb.ae_errors(ae, imgs, dev, conv_mode="fp16").double()
    e_ref = O.ae_errors(ae.cpu(), imgs.cpu()).double().to(dev)          # the oracle (torch fp32 on the CPU) as the checker
    ae.to("cpu")
    c["max_rel_diff_fp16_vs_oracle"] = float(((e_16 - e_ref).abs() / e_ref).max())
    e_tc = sb.ae_errors(ae, imgs, dev, conv_mode="fp32").double()
    e_bf = sb.ae_errors(ae, imgs, dev, conv_mode="bf16").double()
    c["max_rel_diff_fp32parity_vs_oracle"] = float(((e_tc - e_ref).abs() / e_ref).max())
    c["max_rel_diff_bf16_vs_oracle"] = float(((e_bf - e_ref).abs() / e_ref).max())
    del big
    if not a.no_cpu:
        ic = imgs[:512].cpu()
        tc = cpu_time(lambda: O.detect_outliers_autoencoder(ae, ic, 2.0), 2, 1)
        c.update(cpu_ms_per_512=tc * 1e3, cpu_samples_per_s=512 / tc)
    out["configs"]["C4_autoencoder_b512"] = c

    # ---- second headline metric: DCGAN 64x64 train iters/sec, batch 128 --------------------------------------
    out["train_iters_per_sec"] = {
        "batch": 128,
        "note": "generator forward/backward and Adam = torch autograd (cuDNN) in all arms; *_d_on_tcgen05 arms run the "
                "discriminator's forward + backward (D step and G step) on this repo's kernels (sb.accelerate_discriminator)",
        "plain_dcgan_no_strain": train_iters(dev, a.iters, "none"),
        "reference_eager_strain_block_on_gpu": train_iters(dev, a.iters, "torch"),
        "b200_strain_batch_concat_fake": train_iters(dev, a.iters, "b200"),
        "plain_dcgan_d_on_tcgen05": train_iters(dev, a.iters, "none_train"),
        "b200_strain_and_d_on_tcgen05": train_iters(dev, a.iters, "b200_train"),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
