"""Where does the step time go?  Compares, for the C5 workload of bench.py:
  (a) sum of per-layer CUDA-event times inside a loop,   (b) CUDA events around each whole-chunk score,
  (c) the full step loop, at several loop lengths (clock / power behaviour), with SM clocks sampled at 20 ms.
  python tools/gap_probe.py [--chunk 8192] [--mode bf16]"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402


class Clk:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "20"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for line in self.p.stdout:
            try:
                a, b = line.split(",")
                self.rows.append((time.perf_counter(), float(a), float(b)))
            except Exception:
                pass

    def window(self, t0, t1):
        r = [(c, p) for (t, c, p) in self.rows if t0 <= t <= t1]
        if not r:
            return None
        return {"sm_mhz_med": float(np.median([c for c, _ in r])), "sm_mhz_min": min(c for c, _ in r),
                "power_w_med": float(np.median([p for _, p in r])), "n": len(r)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=8192)
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--shard", type=int, default=131072)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    netD = O.make_discriminator(O.SEED).eval()
    images = sb.synth_images(0, a.shard, O.SEED, dev)
    sc = sb.D64Scorer(netD, dev, a.mode, max_batch=a.chunk)
    losses = torch.empty(a.shard, dtype=torch.float32, device=dev)
    nch = a.shard // a.chunk
    clk = Clk()
    out = {"chunk": a.chunk, "mode": a.mode}

    def chunk_loop(reps, per_layer):
        evs = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in range(reps):
            i = (r % nch) * a.chunk
            x = images[i:i + a.chunk]
            e = [torch.cuda.Event(enable_timing=True) for _ in range(7 if per_layer else 2)]
            e[0].record()
            if per_layer:
                for layer in range(1, 6):
                    sc.run_layer(x, layer, None, None, losses[i:i + a.chunk])
                    e[layer].record()
            else:
                sc.score_into(x, None, None, losses[i:i + a.chunk])
                e[1].record()
            evs.append(e)
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if per_layer:
            t = np.array([[e[l].elapsed_time(e[l + 1]) for l in range(5)] for e in evs])
        else:
            t = np.array([[e[0].elapsed_time(e[1])] for e in evs])
        total = evs[0][0].elapsed_time(evs[-1][5 if per_layer else 1])
        return {"reps": reps, "mean_ms": [float(v) for v in t.mean(axis=0)], "sum_ms": float(t.mean(axis=0).sum()),
                "first_quarter_sum": float(t[:max(1, reps // 4)].mean(axis=0).sum()),
                "last_quarter_sum": float(t[-max(1, reps // 4):].mean(axis=0).sum()),
                "wall_per_rep_ms": total / reps, "host_enqueue_ms_per_rep": t_host / reps * 1e3, "clk": clk.window(t0, t1)}

    for _ in range(2):
        chunk_loop(16, False)
    for reps in (16, 64, 256):
        out[f"per_layer_{reps}"] = chunk_loop(reps, True)
        time.sleep(0.5)
        out[f"whole_{reps}"] = chunk_loop(reps, False)
        time.sleep(0.5)
    sc.check()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
