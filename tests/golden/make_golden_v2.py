"""Round-2 fixtures, generated like golden_v1 by EXECUTING THE REFERENCE'S OWN CODE (tests/golden/make_golden.py has
the loader): the in-batch strain block at the batch sizes of BASELINE.json configs 3 / 4 (B = 256, 512) with the
reference Discriminator in TRAIN mode, and the strained->fake concat block with its gradient.

    python tests/golden/make_golden_v2.py        # needs /root/reference (build container only)
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import TorchProxy, load_nodes  # noqa: E402
from oracle import strainer_oracle as O  # noqa: E402


def main():
    out = {}
    ns = load_nodes("#strainer gan.py", ["Discriminator"], dict(torch=TorchProxy()))
    od = O.make_discriminator(O.SEED)

    # ---- in-batch strain block, literal re-typing of "# 상위 10% 제거해서 fake image에 concate.py:243-251" with the
    #      reference Discriminator (train mode: batch-stat BN + running-stat update) ----------------
    for B in (256, 512):
        refD = ns["Discriminator"](1)
        refD.load_state_dict(od.state_dict())
        real_cpu = torch.from_numpy(O.synth_images(0, B))
        with torch.no_grad():
            real_scores = refD(real_cpu).view(-1)
            threshold = torch.quantile(real_scores, 0.1)
            mask = real_scores >= threshold
            filtered_real = real_cpu[mask]
            filtered_fake = real_cpu[~mask]
        out[f"g8_{B}_scores"] = real_scores.numpy()
        out[f"g8_{B}_threshold"] = threshold.numpy()
        out[f"g8_{B}_mask"] = mask.numpy()
        out[f"g8_{B}_nreal"] = np.int64(filtered_real.size(0))
        out[f"g8_{B}_nfake"] = np.int64(filtered_fake.size(0))
        for li, name in ((3, "bn1"), (6, "bn2"), (9, "bn3")):
            out[f"g8_{B}_{name}_mean"] = refD.main[li].running_mean.numpy().copy()
            out[f"g8_{B}_{name}_var"] = refD.main[li].running_var.numpy().copy()
        out[f"g8_{B}_nbt"] = np.int64(refD.main[3].num_batches_tracked.item())

    # ---- strained -> fake concat, ":265-273, 282-284": fake = cat([G(z), filtered_fake]); labels; the generator loss
    #      through netD(fake) back to the generator rows.  B = 64; "G(z)" is a seeded leaf standing in for netG(noise).
    B = 64
    refD = ns["Discriminator"](1)
    refD.load_state_dict(od.state_dict())
    refD.eval()                                    # deterministic D for the gradient fixture
    real_cpu = torch.from_numpy(O.synth_images(1000, B))
    with torch.no_grad():
        real_scores = refD(real_cpu).view(-1)
        threshold = torch.quantile(real_scores, 0.1)
        mask = real_scores >= threshold
        filtered_real = real_cpu[mask]
        filtered_fake = real_cpu[~mask]
    b_size_fake = filtered_fake.size(0)
    g = torch.Generator().manual_seed(1234)
    gz = torch.tanh(torch.randn(B - b_size_fake, 3, 64, 64, generator=g)).requires_grad_(True)
    fake = torch.cat([gz, filtered_fake], dim=0)
    label_fake = torch.full((fake.size(0),), 0.0, dtype=torch.float)
    label_g = torch.full((fake.size(0),), 1.0, dtype=torch.float)
    output = refD(fake).view(-1)
    errG = nn.BCELoss()(output, label_g)
    errG.backward()
    out["g9_scores"] = real_scores.numpy()
    out["g9_mask"] = mask.numpy()
    out["g9_nfake"] = np.int64(b_size_fake)
    out["g9_label_len"] = np.int64(label_fake.numel())
    out["g9_fake_rowsum"] = fake.detach().double().sum(dim=(1, 2, 3)).numpy()
    out["g9_output"] = output.detach().numpy()
    out["g9_errG"] = errG.detach().numpy()
    out["g9_grad_rowsum"] = gz.grad.double().sum(dim=(1, 2, 3)).numpy()
    out["g9_grad_sample"] = gz.grad[:, :, ::16, ::16].numpy().copy()

    path = os.path.join(HERE, "golden_v2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
