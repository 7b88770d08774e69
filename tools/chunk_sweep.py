"""Step time of the C5 strain (bench.py's strain_step) against the scoring chunk size: small chunks keep the
activations in the 126 MB L2 (act1 of 512 samples = 67 MB) but pay launch gaps and wave quantisation.
  python tools/chunk_sweep.py [--mode bf16] [--shard 131072]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--shard", type=int, default=131072)
    ap.add_argument("--chunks", default="512,768,1024,2048,4096,8192,16384")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    netD = O.make_discriminator(O.SEED).eval()
    images = sb.synth_images(0, a.shard, O.SEED, dev)
    losses = torch.empty(a.shard, dtype=torch.float32, device=dev)
    out = {}
    for chunk in [int(c) for c in a.chunks.split(",")]:
        sc = sb.D64Scorer(netD, dev, a.mode, max_batch=chunk)

        def step():
            for i in range(0, a.shard, chunk):
                b = min(chunk, a.shard - i)
                sc.score_into(images[i:i + b], None, None, losses[i:i + b])
            thr = sb.percentile_device(losses, 90.0)
            return sb.compact_indices(losses, thr, 0, 0)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            step()
        e1.record()
        torch.cuda.synchronize()
        sc.check()
        ms = e0.elapsed_time(e1) / 8
        out[chunk] = {"ms_per_step": ms, "samples_per_s": a.shard / ms * 1e3}
        del sc
    print(json.dumps(out))


if __name__ == "__main__":
    main()
