"""End-to-end rate of refine_dataset_by_loss on a pinned fp32 host dataset against the share of rows the host threads
round to fp16 before the PCIe copy (scorer.host_pack) and the number of host threads.
python tools/host_pack_sweep.py [--n 65536]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--shares", default="0,0.5,0.7,0.8,0.9,1.0")
    ap.add_argument("--threads", default="0")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    netD = O.make_discriminator(O.SEED).to(dev).eval()
    host = torch.empty((a.n, 3, 64, 64), dtype=torch.float32).pin_memory()
    for i in range(0, a.n, 8192):
        host[i:i + 8192].copy_(sb.synth_images(i, min(8192, a.n - i), O.SEED, dev))
    ds = torch.utils.data.TensorDataset(host, torch.zeros(a.n, dtype=torch.long))
    sc = sb.scorer_for(netD, dev, "auto", sb.api._chunk_for(a.n))
    real_threads = sb.api._host_threads
    out = []
    for thr in [int(t) for t in a.threads.split(",")]:
        sb.api._host_threads = (lambda t=thr: t) if thr > 0 else real_threads
        for share in [float(x) for x in a.shares.split(",")]:
            sc.host_pack = share if share > 0 else False
            for _ in range(2):
                sb.refine_dataset_by_loss(ds, netD, dev, 0.1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(4):
                sb.refine_dataset_by_loss(ds, netD, dev, 0.1)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 4
            out.append({"threads": thr or real_threads(), "share": share, "samples_per_s": a.n / dt, "ms_per_8192": dt / a.n * 8192e3})
            print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
