"""BASELINE config 1 at dataset scale (N = 60 000 28x28 images): MLP D on the tcgen05 GEMM chain vs the fp32 CUDA-core
kernels, and the DCGAN-28 conv D.  Prints one JSON object with samples/s and the fraction of the tensor roofline.
    python tools/config1_bench.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402

N = 60000
MLP_FLOP = 2 * (784 * 1024 + 1024 * 512 + 512 * 256 + 256)             # 2.92 MFLOP / sample
D28_FLOP = 2 * (196 * 64 * 16 + 49 * 128 * 1024 + 6272)                # 13.3 MFLOP / sample


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
    x = torch.from_numpy(O.synth_images28(0, 4096)).to(dev).repeat(15, 1, 1, 1)[:N].contiguous()
    out = {"images": N, "tensor_peak_tflops": peak}
    torch.manual_seed(3)
    mlp = O.MLPDiscriminator().eval()
    for mode, chunk in (("fp16", 16384), ("fp16", 65536), ("fp32", 16384)):
        sc = sb.MLPScorer(mlp, dev, max_batch=chunk, mode=mode)
        t = timed(lambda: sc.score(x, ("loss",)), 10 if mode == "fp16" else 3)
        out[f"mlp_{mode}_chunk{chunk}"] = {"samples_per_s": N / t, "ms": t * 1e3, "tflops": MLP_FLOP * N / t / 1e12,
                                           "frac_of_tensor_peak": MLP_FLOP * N / t / 1e12 / peak}
    d28 = O.make_discriminator28().eval()
    for chunk in (8192, 32768):
        sc = sb.D28Scorer(d28, dev, max_batch=chunk)
        t = timed(lambda: sc.score(x, ("loss",)))
        out[f"d28_chunk{chunk}"] = {"samples_per_s": N / t, "ms": t * 1e3, "tflops": D28_FLOP * N / t / 1e12,
                                    "frac_of_tensor_peak": D28_FLOP * N / t / 1e12 / peak,
                                    "im2col_bytes_per_sample": 2 * 49 * 1024 * 2}
    ds = torch.utils.data.TensorDataset(x, torch.zeros(N))
    t0 = time.perf_counter()
    for _ in range(5):
        sb.refine_dataset_by_loss(ds, mlp, dev, 0.1)
    torch.cuda.synchronize()
    out["mlp_refine_dataset_by_loss_samples_per_s"] = 5 * N / (time.perf_counter() - t0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
