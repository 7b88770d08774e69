"""One bf16-mode auto-encoder scoring pass over 8192 resident images (ncu target: per-kernel times of config 4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import strainer_b200 as sb  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402

torch.manual_seed(O.SEED)
ae = O.AutoEncoder().eval()
x = sb.synth_images(0, 8192, O.SEED, torch.device("cuda", 0))
for _ in range(3):
    e = sb.ae_errors(ae, x, "cuda", chunk=8192, conv_mode=sys.argv[1] if len(sys.argv) > 1 else "bf16")
torch.cuda.synchronize()
print(float(e.mean()))
