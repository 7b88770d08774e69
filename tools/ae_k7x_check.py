"""Development check of the shifted-window 7x7 kernels: reconstruction errors against the oracle (torch CPU fp32) with
random-scaled weights, per conv mode, and the scoring rate.  python tools/ae_k7x_check.py [n]"""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    dev = torch.device("cuda", 0)
    torch.manual_seed(O.SEED)
    ae = O.AutoEncoder().eval()
    with torch.no_grad():
        for p_ in ae.parameters():          # larger, sign-mixed weights than the default init: every tap matters
            p_.mul_(1.7)
    x = sb.synth_images(0, n, O.SEED, dev)
    ref = O.ae_errors(ae, x[:300].cpu()).numpy()
    out = {"k7x": os.environ.get("SG_AE_K7X", "1"), "images": n}
    for mode in ("bf16", "fp16"):
        e = sb.ae_errors(ae, x[:300], dev, conv_mode=mode).cpu().numpy()
        out[mode + "_max_rel"] = float((abs(e - ref) / abs(ref).clip(1e-6)).max())
        e = sb.ae_errors(ae, x, dev, chunk=8192, conv_mode=mode)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            e = sb.ae_errors(ae, x, dev, chunk=8192, conv_mode=mode)
        torch.cuda.synchronize()
        out[mode + "_samples_per_s"] = n / ((time.perf_counter() - t0) / 5)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
