"""Generate tests/golden/golden_v1.npz by EXECUTING THE REFERENCE'S OWN CODE.

Runs only in the build container (needs /root/reference).  The reference scripts cannot be
imported (they fetch datasets at import time), so the wanted ``FunctionDef`` / ``ClassDef``
nodes are AST-extracted and exec'd into a namespace seeded with the globals they expect
(SURVEY.md §8c).  No reference source is copied into this repository: only the numeric
outputs on seeded synthetic inputs are stored.

    python tests/golden/make_golden.py
"""
import ast
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import strainer_oracle as O  # noqa: E402

REF = "/root/reference"


def load_nodes(fname, names, extra=None):
    from sklearn.cluster import DBSCAN
    from sklearn.mixture import GaussianMixture
    from sklearn.preprocessing import StandardScaler
    src = open(os.path.join(REF, fname), encoding="utf-8").read()
    tree = ast.parse(src)
    ns = dict(torch=torch, nn=nn, np=np, F=F, GaussianMixture=GaussianMixture, DBSCAN=DBSCAN,
              StandardScaler=StandardScaler, nc=3, ndf=64, ngpu=1, device=torch.device("cpu"),
              real_label=1.0, print=lambda *a, **k: None,
              visualize_z_scores=lambda *a, **k: None)
    ns.update(extra or {})
    want = set(names)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in want:
            if node.name == "visualize_z_scores":
                continue
            exec(compile(ast.Module([node], []), fname, "exec"), ns)
            want.discard(node.name)
    assert not want, f"missing {want} in {fname}"
    return ns


class NoWorkerLoader(torch.utils.data.DataLoader):
    """The reference asks for num_workers=2; results do not depend on it. Forked workers are
    avoided so the generator also runs in restricted sandboxes."""

    def __init__(self, *a, **k):
        k["num_workers"] = 0
        super().__init__(*a, **k)


class _DataNS:
    DataLoader = NoWorkerLoader
    Subset = torch.utils.data.Subset


class _UtilsNS:
    data = _DataNS


class TorchProxy:
    """``torch`` with ``torch.utils.data.DataLoader`` forced to num_workers=0."""
    utils = _UtilsNS

    def __getattr__(self, k):
        return getattr(torch, k)


def main():
    out = {}
    tp = TorchProxy()

    # ---- G1: refine_dataset_by_loss ("#strainer gan.py:364-392") -------------------------
    ns = load_nodes("#strainer gan.py", ["Discriminator", "refine_dataset_by_loss", "find_elbow_threshold",
                                         "detect_outliers"], dict(torch=tp))
    n1 = 160
    imgs = torch.from_numpy(O.synth_images(0, n1))
    od = O.make_discriminator(O.SEED)
    refD = ns["Discriminator"](1)
    refD.load_state_dict(od.state_dict())
    ds = torch.utils.data.TensorDataset(imgs, torch.zeros(n1, dtype=torch.long))
    for tag, ratio in (("a", 0.2), ("b", 1 - 0.8), ("c", 0.1)):
        sub, thr = ns["refine_dataset_by_loss"](ds, refD, torch.device("cpu"), ratio)
        out[f"g1{tag}_indices"] = np.asarray(sub.indices)
        out[f"g1{tag}_threshold"] = np.asarray(thr)
        out[f"g1{tag}_ratio"] = np.float64(ratio)
    with torch.no_grad():
        refD.eval()
        p = refD(imgs)
        out["g1_probs"] = p.reshape(-1).numpy()
        out["g1_logits"] = refD.main[:-1](imgs).reshape(-1).numpy()
        out["g1_losses"] = nn.BCELoss(reduction="none")(p, torch.ones_like(p)).reshape(-1).numpy()
    # degenerate fallback: all-equal losses through the reference itself
    const_imgs = torch.zeros(8, 3, 64, 64)
    ds0 = torch.utils.data.TensorDataset(const_imgs, torch.zeros(8, dtype=torch.long))
    sub, thr = ns["refine_dataset_by_loss"](ds0, refD, torch.device("cpu"), 0.2)
    out["g1z_indices"] = np.asarray(sub.indices)
    out["g1z_threshold"] = np.asarray(thr)

    # ---- G4/G5: elbow threshold + detect_outliers variants ------------------------------
    feats = torch.from_numpy(O.synth_features(4096))
    fds = torch.utils.data.TensorDataset(feats, torch.zeros(4096, dtype=torch.long))
    ident = nn.Identity()
    inl = ns["detect_outliers"](fds, ident, None)
    out["g5_elbow_inlier"] = np.asarray(inl)
    inl = ns["detect_outliers"](fds, ident, 5)
    out["g5_user5_inlier"] = np.asarray(inl)
    mz = O.zscore_max_torch(feats).numpy()
    thr, centers, hist = ns["find_elbow_threshold"](mz)
    out["g4_maxz"] = mz
    out["g4_threshold"] = np.asarray(thr)
    out["g4_centers"] = centers
    out["g4_hist"] = hist
    ns8 = load_nodes("#z_score.py", ["detect_outliers"], dict(torch=tp))
    out["g5_fixed5_inlier"] = ns8["detect_outliers"](fds, ident).numpy()
    out["g5_fixed45_inlier"] = ns8["detect_outliers"](fds, ident, 4.5).numpy()
    ns11 = load_nodes("# z_score + DBSCAN.py", ["detect_outliers"], dict(torch=tp))
    for tag, cr in (("a", 0.9), ("b", 0.8731)):
        out[f"g5_ratio{tag}_inlier"] = ns11["detect_outliers"](fds, ident, cr).numpy()
        out[f"g5_ratio{tag}"] = np.float64(cr)
    ns15 = load_nodes("# 1,2,8.py", ["compute_z_scores"], dict(torch=tp))
    out["g7_maxz_np"] = ns15["compute_z_scores"](fds, ident)

    # ---- G2: evaluate_dataset + GMM divide_dataset ("#clean ... .py:272-316") -----------
    fn = [f for f in os.listdir(REF) if f.startswith("#clean")][0]
    ns12 = load_nodes(fn, ["evaluate_dataset", "divide_dataset"], dict(torch=tp))
    refD2 = ns["Discriminator"](1)
    refD2.load_state_dict(od.state_dict())
    ev = ns12["evaluate_dataset"](refD2, ds, torch.device("cpu"))
    out["g2_eval_losses"] = ev
    lo = O.synth_losses(5000)
    out["g2_losses"] = lo
    np.random.seed(1234)
    dsl = torch.utils.data.TensorDataset(torch.zeros(5000, 1))
    clean, noisy = ns12["divide_dataset"](lo.copy(), dsl)
    out["g2_clean_idx"] = np.asarray(clean.indices)
    out["g2_noisy_idx"] = np.asarray(noisy.indices)

    # ---- G3: ensemble thresholds ("# 종합 loss.py:270-312") ------------------------------
    fn = [f for f in os.listdir(REF) if "종합" in f][0]
    ns13 = load_nodes(fn, ["get_gmm_threshold", "get_percentile_threshold", "get_iqr_threshold",
                           "get_ensemble_threshold", "divide_dataset"], dict(torch=tp))
    np.random.seed(1234)
    out["g3_gmm"] = np.asarray(ns13["get_gmm_threshold"](lo.copy()))
    out["g3_p75"] = np.asarray(ns13["get_percentile_threshold"](lo.copy()))
    out["g3_p75_dtype"] = np.asarray(str(np.asarray(ns13["get_percentile_threshold"](lo.copy())).dtype))
    out["g3_iqr"] = np.asarray(ns13["get_iqr_threshold"](lo.copy()))
    np.random.seed(1234)
    out["g3_ensemble"] = np.asarray(ns13["get_ensemble_threshold"](lo.copy()))
    np.random.seed(1234)
    clean, noisy = ns13["divide_dataset"](lo.copy(), dsl)
    out["g3_clean_idx"] = np.asarray(clean.indices)

    # ---- G6: auto-encoder straining ("#autoencoder.py:269-322") -------------------------
    ns14 = load_nodes("#autoencoder.py", ["AutoEncoder", "detect_outliers_autoencoder"], dict(torch=tp))
    torch.manual_seed(O.SEED)
    oae = O.AutoEncoder()
    rae = ns14["AutoEncoder"]()
    rae.load_state_dict(oae.state_dict())
    out["g6_inlier"] = ns14["detect_outliers_autoencoder"](rae, ds, torch.device("cpu")).numpy()
    out["g6_inlier_t05"] = ns14["detect_outliers_autoencoder"](rae, ds, torch.device("cpu"), 0.5).numpy()
    with torch.no_grad():
        rae.eval()
        o = rae(imgs)
        out["g6_errors"] = F.mse_loss(o, imgs, reduction="none").view(n1, -1).mean(dim=1).numpy()

    # ---- G8: in-batch strain block: literal re-typing of
    #      "# 상위 10% 제거해서 fake image에 concate.py:243-251, 265-269" with the reference D ----
    for B in (64, 128):
        refD3 = ns["Discriminator"](1)
        refD3.load_state_dict(od.state_dict())  # stays in train mode
        real_cpu = imgs[:B]
        with torch.no_grad():
            real_scores = refD3(real_cpu).view(-1)
            threshold = torch.quantile(real_scores, 0.1)
            mask = real_scores >= threshold
            filtered_real = real_cpu[mask]
            filtered_fake = real_cpu[~mask]
        out[f"g8_{B}_scores"] = real_scores.numpy()
        out[f"g8_{B}_threshold"] = threshold.numpy()
        out[f"g8_{B}_mask"] = mask.numpy()
        out[f"g8_{B}_nreal"] = np.int64(filtered_real.size(0))
        out[f"g8_{B}_nfake"] = np.int64(filtered_fake.size(0))
        bn = refD3.main[3]
        out[f"g8_{B}_bn1_mean"] = bn.running_mean.numpy().copy()
        out[f"g8_{B}_bn1_var"] = bn.running_var.numpy().copy()
        out[f"g8_{B}_bn3_mean"] = refD3.main[9].running_mean.numpy().copy()
        out[f"g8_{B}_bn3_var"] = refD3.main[9].running_var.numpy().copy()
        out[f"g8_{B}_nbt"] = np.int64(bn.num_batches_tracked.item())

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
