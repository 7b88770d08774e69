"""Auto-encoder scoring throughput (BASELINE config 4) per conv mode + per-kernel launch times are taken with ncu
(tools/ncu_launch_table.py).  python tools/ae_bench.py [n_images]"""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    dev = torch.device("cuda", 0)
    torch.manual_seed(O.SEED)
    ae = O.AutoEncoder().eval()
    x = sb.synth_images(0, n, O.SEED, dev)
    out = {"images": n}
    ref = O.ae_errors(ae, x[:256].cpu()).numpy()
    for mode in ("bf16", "fp16", "fp32"):
        for chunk in (2048, 8192):
            e = sb.ae_errors(ae, x, dev, chunk=chunk, conv_mode=mode)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 5 if mode != "fp32" else 2
            for _ in range(reps):
                e = sb.ae_errors(ae, x, dev, chunk=chunk, conv_mode=mode)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            rel = float((abs(e[:256].cpu().numpy() - ref) / abs(ref).clip(1e-6)).max())
            out[f"{mode}_chunk{chunk}"] = {"samples_per_s": n / dt, "ms": dt * 1e3, "max_rel_err_vs_oracle_256": rel}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
