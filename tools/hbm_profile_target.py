"""ncu target: ONE launch of each HBM-bound kernel of the path at sizes far beyond L2 (2^28 elements / 65536 image rows /
131072 feature rows) plus one 8192-sample D64 scoring chunk (conv kernels + head) in the default conv mode.
    ncu --set full --clock-control none -k regex:'filter_kernel|radix_phases|compact_indices|move_rows|partition_dest|col_partial|
        row_max_absz|hist_uniform|head_kernel|minmax|chunk_moments|conv|nchw' -o gpurun_out/r2_hbm python tools/hbm_profile_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402

dev = torch.device("cuda", 0)
n = 1 << 28
g = torch.Generator(device=dev).manual_seed(1)
v = torch.empty(n, dtype=torch.float32, device=dev).exponential_(generator=g)
# one-pass select (sample_pivot + filter + radix_phases) and the index compaction at keep = 0.9
thr = sb.percentile_device(v, 90.0)
idx, cnt, _ = sb.compact_indices(v, thr, 0, 0)
mm = sb.find_elbow_threshold(v[: 1 << 27])          # minmax + hist_uniform (2^27 elements)
t2 = sb.mean_plus_k_std(v, 2.0)                      # chunk_moments + finish
del idx
torch.cuda.synchronize()
# row partition of 65536 image rows (partition_dest + move_rows): 6.4 GB moved
rows = sb.synth_images(0, 32768, O.SEED, dev)
mask = (torch.arange(32768, device=dev) % 10 != 0)
kept, dropped, counts = sb.partition_rows(rows, mask)
del kept, dropped
# feature z-score: col_partial / col_finish / row_max_absz over [131072, 512] fp32 (268 MB)
f = torch.empty((131072, 512), dtype=torch.float32, device=dev).normal_(generator=g)
z = sb.zscore_max(f)
# u8 normalise of 32768 images
px = (torch.rand((32768, 3, 64, 64), device=dev) * 255).to(torch.uint8)
xf = sb.U8Images(px).to_f32(dev)
del px, xf, f
# one D64 scoring chunk in the default mode
netD = O.make_discriminator(O.SEED).eval()
loss = sb.get_scorer(netD, dev, "auto", 8192).score(rows[:8192], ("loss",))["loss"]
torch.cuda.synchronize()
print(float(thr), int(cnt), float(z.mean()), float(loss.mean()))
