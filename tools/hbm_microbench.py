"""HBM-bound kernels of the path at sizes far beyond L2 (BASELINE.json north_star: ">= 70 % of HBM peak on
the selection and compaction kernels").  Prints one JSON object; CUDA-event timing, inputs >> 126 MB L2.

  python tools/hbm_microbench.py [--log2n 28]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import strainer_b200 as sb  # noqa: E402

L = sb._lib


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    args = ap.parse_args()
    peak = 6551.4
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    src = "fallback 6650"
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
        src = "measured"
    else:
        peak = 6650.0
    dev = torch.device("cuda", 0)
    lib = L.init(0)
    st = L.P(torch.cuda.current_stream().cuda_stream)
    p = lambda t: L.P(t.data_ptr()) if t is not None else L.P(0)
    n = 1 << args.log2n
    out = {"n": n, "hbm_peak_gbs": peak, "peak_source": src, "kernels": {}}

    def rec(name, secs, nbytes, note=""):
        gbs = nbytes / secs / 1e9
        out["kernels"][name] = {"ms": secs * 1e3, "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_hbm_peak": gbs / peak, "note": note}

    g = torch.Generator(device=dev).manual_seed(1)
    v = torch.empty(n, dtype=torch.float32, device=dev).exponential_(generator=g)
    ws = torch.empty(L.SG_SELECT_WS_WORDS, dtype=torch.int32, device=dev)
    out2 = torch.empty(2, dtype=torch.float32, device=dev)
    L.check(lib.sg_select_begin(p(ws), n // 2, st))
    rec("select_hist_pass0", timeit(lambda: lib.sg_select_hist(p(v), n, p(ws), 0, st)), 4 * n, "one radix-histogram pass: 4 B/elem read")
    rec("radix_select_total", timeit(lambda: lib.sg_radix_select(p(v), n, (9 * n) // 10, p(ws), p(out2), st)), 16 * n,
        "x_(k) and x_(k+1): 4 streaming passes of 4 B/elem (8 key bits each)")
    sws_bytes = int(lib.sg_select_workspace_bytes(n))
    sws = torch.empty(sws_bytes, dtype=torch.uint8, device=dev)
    t = timeit(lambda: lib.sg_select_kth(p(v), n, (9 * n) // 10, p(sws), sws_bytes, p(out2), st))
    ref2 = out2.clone()
    L.check(lib.sg_radix_select(p(v), n, (9 * n) // 10, p(ws), p(out2), st))
    assert torch.equal(ref2, out2), (ref2, out2)
    rec("select_onepass_total", t, 4 * n,
        "x_(k), x_(k+1) by sampled pivots + ONE streaming read (SURVEY 8d: 4 B/elem) + radix passes over ~3 % candidates")
    del sws
    thr = torch.tensor([float(np.percentile(v[:1 << 20].cpu().numpy(), 90))], device=dev)
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    cnt = torch.empty(1, dtype=torch.int64, device=dev)
    cws = torch.empty(lib.sg_compact_workspace_bytes(n), dtype=torch.uint8, device=dev)
    t = timeit(lambda: lib.sg_compact_indices(p(v), n, p(thr), 0, 0, p(idx), p(cnt), L.P(0), p(cws), st))
    keep = cnt.item() / n
    rec("compact_indices", t, int(4 * n + 8 * keep * n), f"4 B read + 8 B written per kept index, keep={keep:.3f}")
    del idx
    # image-row compaction: rows of 49152 B, 90 % kept
    rows_n = 65536
    rows = torch.empty((rows_n, 3, 64, 64), dtype=torch.float32, device=dev).normal_(generator=g)
    mask = (torch.rand(rows_n, device=dev, generator=g) < 0.9).to(torch.uint8)
    kept = torch.empty_like(rows)
    dropped = torch.empty((rows_n // 4, 3, 64, 64), dtype=torch.float32, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev)
    rws = torch.empty(lib.sg_compact_workspace_bytes(rows_n), dtype=torch.uint8, device=dev)
    t = timeit(lambda: lib.sg_compact_rows(p(rows), rows_n, 49152, p(mask), p(kept), p(dropped), p(counts), p(rws), st))
    rec("compact_rows_49152B", t, 2 * 49152 * rows_n + rows_n, "every row read once and written once (kept or dropped) + 1 B mask")
    del rows, kept, dropped
    # feature z-score [N,512]
    fn = 1 << 21
    f = torch.empty((fn, 512), dtype=torch.float32, device=dev).normal_(generator=g)
    mean = torch.empty(512, dtype=torch.float32, device=dev)
    den = torch.empty(512, dtype=torch.float32, device=dev)
    zws = torch.empty(lib.sg_col_moments_workspace_bytes(fn, 512), dtype=torch.uint8, device=dev)
    zo = torch.empty(fn, dtype=torch.float32, device=dev)
    rec("col_moments_512", timeit(lambda: lib.sg_col_moments(p(f), fn, 512, 1, 0.0, p(mean), p(den), p(zws), st)), 2048 * fn, "stats pass: 2048 B/row")
    rec("row_max_absz_512", timeit(lambda: lib.sg_row_max_absz(p(f), fn, 512, p(mean), p(den), p(zo), st)), 2052 * fn, "z pass: 2048 B/row + 4 B out")
    del f
    mm = torch.empty(8, dtype=torch.float32, device=dev)
    rec("minmax", timeit(lambda: lib.sg_minmax(p(v), n, p(mm), st)), 4 * n)
    edges = torch.linspace(0, float(v[:1 << 20].max()) * 2, 101, device=dev)
    hc = torch.zeros(100, dtype=torch.int64, device=dev)
    rec("hist_uniform_100", timeit(lambda: lib.sg_hist_uniform(p(v), n, p(edges), 100, p(hc), st)), 4 * n)
    part = torch.empty(2 * (n // L.SG_MOMENT_CHUNK + 1), dtype=torch.float64, device=dev)
    rec("chunk_moments", timeit(lambda: lib.sg_chunk_moments(p(v), n, p(part), st)), 4 * n)
    # uint8 pixels -> fp32 NCHW (ToTensor + Normalize): 65536 images, 1 B read + 4 B written per element
    un = 65536
    cm = (L.c_float * 3)(0.5, 0.5, 0.5)
    f32 = torch.empty((un, 3, 64, 64), dtype=torch.float32, device=dev)
    for name, lay, shape in (("u8_normalize_nchw", L.SG_LAYOUT_NCHW, (un, 3, 64, 64)), ("u8_normalize_nhwc", L.SG_LAYOUT_NHWC, (un, 64, 64, 3))):
        u8 = torch.randint(0, 256, shape, dtype=torch.uint8, device=dev, generator=g)
        t = timeit(lambda: lib.sg_u8_normalize(p(u8), un, 3, 4096, lay, cm, cm, p(f32), st))
        rec(name, t, 5 * un * 12288, "1 B read + 4 B written per element (random pixels: worst case for the look-up table)")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
