"""Device radix sort (sg_sort_f32) and the 1-D DBSCAN clean ratio (sg_dbscan1d) at the path's size (2^20 losses) and far
beyond L2 (2^26).  Algorithmic bytes (SURVEY 8d "1-D DBSCAN K17": >= 4 passes x 8 B for keys alone): 4 B histogram read +
4 passes x (read + write of key and payload)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

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
    dev = torch.device("cuda", 0)
    lib = L.init(0)
    st = L.P(torch.cuda.current_stream().cuda_stream)
    p = lambda t: L.P(t.data_ptr()) if t is not None else L.P(0)
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    out = {"hbm_peak_gbs": peak}
    g = torch.Generator(device=dev).manual_seed(5)
    for log2n in (20, 24, 26):
        n = 1 << log2n
        v = torch.empty(n, dtype=torch.float32, device=dev).normal_(generator=g).exp_()
        ws = torch.empty(lib.sg_sort_workspace_bytes(n), dtype=torch.uint8, device=dev)
        so = torch.empty(n, dtype=torch.float32, device=dev)
        order = torch.empty(n, dtype=torch.int32, device=dev)
        t_pairs = timeit(lambda: L.check(lib.sg_sort_f32(p(v), n, p(so), p(order), p(ws), st)))
        t_keys = timeit(lambda: L.check(lib.sg_sort_f32(p(v), n, p(so), L.P(0), p(ws), st)))
        ref, ref_order = torch.sort(v, stable=True)
        assert torch.equal(so, ref) and torch.equal(order.long(), ref_order)
        t_torch = timeit(lambda: torch.sort(v, stable=True))
        t_torch_keys = timeit(lambda: torch.sort(v))
        dws = torch.empty(lib.sg_dbscan1d_workspace_bytes(n), dtype=torch.uint8, device=dev)
        counts = torch.empty(1, dtype=torch.int64, device=dev)
        t_db = timeit(lambda: L.check(lib.sg_dbscan1d(p(v), n, 1e-6, 3, p(counts), L.P(0), p(dws), st)))
        bytes_pairs = n * (4 + 12 + 3 * 16)
        bytes_keys = n * (4 + 4 * 8)
        out[f"n=2^{log2n}"] = {
            "sort_pairs_ms": t_pairs * 1e3, "sort_pairs_gbs": bytes_pairs / t_pairs / 1e9, "sort_pairs_frac": bytes_pairs / t_pairs / 1e9 / peak,
            "sort_keys_ms": t_keys * 1e3, "sort_keys_gbs": bytes_keys / t_keys / 1e9, "sort_keys_frac": bytes_keys / t_keys / 1e9 / peak,
            "torch_sort_stable_ms": t_torch * 1e3, "torch_sort_ms": t_torch_keys * 1e3, "dbscan1d_ms": t_db * 1e3,
            "melem_per_s_pairs": n / t_pairs / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
