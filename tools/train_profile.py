"""Where a DCGAN 64x64 training iteration (B = 128, "#strainer gan.py:581-633") spends its GPU time: torch.profiler over
the plain autograd loop of tools/config_bench.py, kernels grouped by family.  One JSON object on stdout.

  python tools/train_profile.py [--iters 20] [--mode none|torch|b200|b200_train]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import config_bench  # noqa: E402


def family(name: str) -> str:
    n = name.lower()
    for key, fam in (("wgrad", "conv wgrad"), ("dgrad", "conv dgrad"), ("bn_bw", "batchnorm bwd"), ("bn_fw", "batchnorm fwd"),
                     ("batch_norm_backward", "batchnorm bwd"), ("batch_norm", "batchnorm fwd"),
                     ("cudnn", "cudnn other"), ("cutlass", "cutlass conv/gemm"), ("xmma", "xmma conv"), ("sm100", "sm100 conv/gemm"),
                     ("sm90", "sm90 conv/gemm"), ("sm80", "sm80 conv/gemm"), ("implicit", "implicit gemm conv"),
                     ("adam", "adam"), ("multi_tensor", "foreach / multi_tensor"), ("leaky", "leaky relu"),
                     ("threshold", "relu"), ("sigmoid", "sigmoid"), ("tanh", "tanh"), ("binary_cross", "bce"),
                     ("elementwise", "elementwise"), ("reduce", "reduce"), ("memcpy", "memcpy"), ("memset", "memset"),
                     ("sg::", "strainer_b200"), ("d64::", "strainer_b200"), ("cmp::", "strainer_b200"), ("dtr::", "strainer_b200")):
        if key in n:
            return fam
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--mode", default="none")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    from torch.profiler import ProfilerActivity, profile

    # the timed loop of config_bench, re-entered under the profiler: train_iters(dev, iters, mode) runs warm-up + iters
    rate = config_bench.train_iters(dev, a.iters, a.mode)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        config_bench.train_iters(dev, a.iters, a.mode)
        torch.cuda.synchronize()
    total_iters = a.iters + 5
    kern = {}
    for ev in prof.key_averages():
        dt = getattr(ev, "self_device_time_total", 0) or getattr(ev, "self_cuda_time_total", 0)
        if dt <= 0 or ev.device_type.name != "CUDA":
            continue
        kern[ev.key] = (dt, ev.count)
    fams = {}
    for k, (dt, cnt) in kern.items():
        f = family(k)
        t, c = fams.get(f, (0.0, 0))
        fams[f] = (t + dt, c + cnt)
    top = sorted(kern.items(), key=lambda kv: -kv[1][0])[:40]
    busy = sum(v[0] for v in kern.values())
    out = {"mode": a.mode, "iters_per_s": rate, "ms_per_iter": 1e3 / rate,
           "gpu_busy_ms_per_iter": busy / total_iters * 1e-3,
           "launches_per_iter": sum(v[1] for v in kern.values()) / total_iters,
           "families_us_per_iter": {f: [round(t / total_iters, 1), round(c / total_iters, 1)]
                                    for f, (t, c) in sorted(fams.items(), key=lambda kv: -kv[1][0])},
           "top_kernels_us_per_iter": [[k[:110], round(dt / total_iters, 1), round(cnt / total_iters, 1)] for k, (dt, cnt) in top]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
