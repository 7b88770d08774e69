"""Profiling target: the in-batch strain block (train-mode BatchNorm scoring + selection + concat) and the training forward
at one batch size.   python tools/strain_block_target.py [--batch 128] [--time]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402


def gpu_time(fn, iters=100, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B = a.batch
    netD = O.make_discriminator(O.SEED).to(dev).train()
    real = sb.synth_images(0, B, O.SEED, dev)
    fake = torch.randn(B, 3, 64, 64, device=dev)
    D = sb.accelerate_discriminator(netD, max_batch=B)

    def block():
        fr, ff, _, _ = sb.strain_batch(netD, real, 0.1)
        return sb.concat_fake(fake[:fr.shape[0]], ff)

    def fwd():
        with torch.no_grad():
            return D(real)
    for _ in range(3):
        block()
        fwd()
    torch.cuda.synchronize()
    if a.time:
        netD.eval()
        t_eval = gpu_time(block)
        netD.train()
        print(json.dumps({"batch": B, "strain_block_train_bn_us": gpu_time(block), "strain_block_eval_bn_us": t_eval,
                          "train_forward_only_us": gpu_time(fwd)}))
    else:
        block()
        fwd()
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
