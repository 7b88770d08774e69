"""CPU ORACLE for the Strainer-GAN straining hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``strainer-gan_b200/`` may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` leg do, and only as the checker / the timed CPU
baseline -- never as a product code path.

The reference (hizibu7/Strainer-GAN) is a set of flat Python scripts that cannot
be imported (they download datasets at import time), has no tests and no golden
vectors.  Every function below is a *restatement* of one reference function or
inline block (cited ``file:line``), built on the same third-party arithmetic the
reference calls (torch CPU, numpy, scikit-learn -- versions un-pinned upstream;
the de-facto pin is this image: torch 2.11, numpy 2.3.5, scikit-learn 1.9).

PARITY PINNING: the restatements are pinned against fixtures produced by
executing the reference's own ``FunctionDef``/``ClassDef`` nodes (AST-loaded
from /root/reference by ``tests/golden/make_golden.py``) on seeded synthetic
inputs; the fixtures are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py`` on CPU.  The bit-level helpers
(``np_percentile_f32``, ``torch_quantile_f32``, ``np_histogram_f32``) restate the
library algorithms with explicit rounding steps so that the CUDA kernels can be
compared stage by stage; they are pinned against ``np.percentile``,
``torch.quantile`` and ``np.histogram`` themselves.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Counter-based synthetic data.  Mirrored bit-for-bit by the CUDA
# generator ``sg_synth_images`` (strainer-gan_b200/csrc/api.cu): integer hash only,
# and the int->float map is exact in fp32, so CPU and GPU agree exactly.
# --------------------------------------------------------------------------------------
SEED = 999  # the reference's manualSeed, "#strainer gan.py:38"
_GOLD = np.uint32(0x9E3779B9)


def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def synth_sample_key(seed: int, idx: np.ndarray) -> np.ndarray:
    idx = np.asarray(idx, dtype=np.uint64)
    lo = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (idx >> np.uint64(32)).astype(np.uint32)
    with np.errstate(over="ignore"):
        return _mix32(np.uint32(seed) ^ _mix32(lo) ^ (hi * np.uint32(0x9E3779B1)))


def synth_is_noisy(seed: int, idx: np.ndarray) -> np.ndarray:
    """~20 % of the samples are iid-uniform 'noise' images (the contaminant)."""
    key = synth_sample_key(seed, idx)
    return (_mix32(key ^ np.uint32(0xA5A5A5A5)) % np.uint32(5)) == 0


def synth_images(start: int, count: int, seed: int = SEED, nc: int = 3, hw: int = 64) -> np.ndarray:
    """fp32 [count, nc, hw, hw] in [-1, 1); sample i depends only on (seed, start+i)."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    key = synth_sample_key(seed, idx)[:, None, None, None]
    noisy = synth_is_noisy(seed, idx)[:, None, None, None]
    c = np.arange(nc, dtype=np.uint32)[None, :, None, None]
    y = np.arange(hw, dtype=np.uint32)[None, None, :, None]
    x = np.arange(hw, dtype=np.uint32)[None, None, None, :]
    with np.errstate(over="ignore"):
        e_fine = (c * np.uint32(hw) + y) * np.uint32(hw) + x
        e_c8 = (c * np.uint32(8) + (y >> np.uint32(3))) * np.uint32(8) + (x >> np.uint32(3))
        e_c2 = (c * np.uint32(32) + (y >> np.uint32(1))) * np.uint32(32) + (x >> np.uint32(1))
        vn = _mix32(key + np.uint32(0x30000000) + e_fine * _GOLD) >> np.uint32(8)
        a = _mix32(key + np.uint32(0x10000000) + e_c8 * _GOLD) >> np.uint32(8)
        b = _mix32(key + np.uint32(0x20000000) + e_c2 * _GOLD) >> np.uint32(8)
        vc = (np.uint32(3) * a.astype(np.uint64) + b.astype(np.uint64)) >> np.uint64(2)
    v = np.where(noisy, vn.astype(np.uint64), vc).astype(np.float32)
    return v * np.float32(2.0 ** -23) - np.float32(1.0)


def to_tensor_normalize(pixels_u8_nchw: np.ndarray, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> np.ndarray:
    """The reference's dataset transform tail, ``transforms.ToTensor()`` + ``transforms.Normalize(mean, std)``
    ("#strainer gan.py:89-90, 115-116"), restated on uint8 NCHW pixels with explicit fp32 steps: x / 255, then
    (x - mean[c]) / std[c] (IEEE division / subtraction, one rounding each).  Pinned against torchvision itself in
    tests/test_u8_dataset.py; mirrored bit for bit by ``sg_u8_normalize``."""
    x = np.asarray(pixels_u8_nchw).astype(np.float32) / np.float32(255)
    m = np.asarray(mean, np.float32).reshape(1, -1, 1, 1)
    s = np.asarray(std, np.float32).reshape(1, -1, 1, 1)
    return (x - m) / s


def synth_features(n: int, d: int = 512, seed: int = SEED, outlier_frac: float = 0.1) -> np.ndarray:
    """[n, d] ~ N(0,1) with ``outlier_frac`` rows shifted +4 sigma in 8 columns."""
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((n, d)).astype(np.float32)
    rows = rng.choice(n, int(n * outlier_frac), replace=False)
    for r in rows:
        cols = rng.choice(d, 8, replace=False)
        f[r, cols] += np.float32(4.0)
    return f


def synth_losses(n: int, seed: int = SEED) -> np.ndarray:
    """80 % lognormal(-1.2, 0.5) + 20 % lognormal(0.7, 0.4) fp32."""
    rng = np.random.default_rng(seed)
    noisy = rng.random(n) < 0.2
    v = np.where(noisy, rng.lognormal(0.7, 0.4, n), rng.lognormal(-1.2, 0.5, n))
    return v.astype(np.float32)


# --------------------------------------------------------------------------------------
# Models (a1, a14, a15, a16)
# --------------------------------------------------------------------------------------
def weights_init(m: nn.Module) -> None:
    """DCGAN init, "#strainer gan.py:166-172": Conv ~ N(0, .02); BN gamma ~ N(1, .02), beta = 0."""
    name = type(m).__name__
    if "Conv" in name:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif "BatchNorm" in name:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class Discriminator(nn.Module):
    """64x64 DCGAN discriminator, layer-for-layer the stack at "#strainer gan.py:230-256"
    (``main.{0,2,5,8,11}`` convs, ``main.{3,6,9}`` BatchNorm2d, LeakyReLU 0.2, Sigmoid)."""

    def __init__(self, ngpu: int = 1, nc: int = 3, ndf: int = 64):
        super().__init__()
        self.ngpu = ngpu
        layers = [nn.Conv2d(nc, ndf, 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True)]
        ch = ndf
        for _ in range(3):
            layers += [nn.Conv2d(ch, ch * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ch * 2),
                       nn.LeakyReLU(0.2, inplace=True)]
            ch *= 2
        layers += [nn.Conv2d(ch, 1, 4, 1, 0, bias=False), nn.Sigmoid()]
        self.main = nn.Sequential(*layers)

    def forward(self, input):
        return self.main(input)


def make_discriminator(seed: int = SEED, perturb_bn: bool = True) -> Discriminator:
    """weights_init'ed D with non-trivial BN running stats so BN folding is exercised (§8d C2)."""
    g = torch.Generator().manual_seed(seed)
    d = Discriminator()
    for m in d.modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data = torch.randn(m.weight.shape, generator=g) * 0.02
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data = 1.0 + torch.randn(m.weight.shape, generator=g) * 0.02
            m.bias.data.zero_()
            if perturb_bn:
                m.running_mean = torch.randn(m.running_mean.shape, generator=g) * 0.1
                m.running_var = torch.rand(m.running_var.shape, generator=g) + 0.5
    return d


class AutoEncoder(nn.Module):
    """Conv auto-encoder of "#autoencoder.py:269-291" (all layers biased)."""

    def __init__(self):
        super().__init__()
        self.encoder = nn.Sequential(
            nn.Conv2d(3, 16, 3, stride=2, padding=1), nn.ReLU(),
            nn.Conv2d(16, 32, 3, stride=2, padding=1), nn.ReLU(),
            nn.Conv2d(32, 64, 7))
        self.decoder = nn.Sequential(
            nn.ConvTranspose2d(64, 32, 7), nn.ReLU(),
            nn.ConvTranspose2d(32, 16, 3, stride=2, padding=1, output_padding=1), nn.ReLU(),
            nn.ConvTranspose2d(16, 3, 3, stride=2, padding=1, output_padding=1), nn.Tanh())

    def forward(self, x):
        return self.decoder(self.encoder(x))


class MLPDiscriminator(nn.Module):
    """28x28 MLP discriminator, "Untitled-2.py:79-94" (784-1024-512-256-1, LeakyReLU .2, Sigmoid);
    ``dropout=True`` gives the "# 1,2,8.py:110-128" variant (Dropout .3, identity in eval)."""

    def __init__(self, dropout: bool = False):
        super().__init__()
        dims = [784, 1024, 512, 256]
        layers = []
        for a, b in zip(dims[:-1], dims[1:]):
            layers += [nn.Linear(a, b), nn.LeakyReLU(0.2)]
            if dropout:
                layers.append(nn.Dropout(0.3))
        layers += [nn.Linear(dims[-1], 1), nn.Sigmoid()]
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class Discriminator28(nn.Module):
    """DCGAN-28 conv discriminator for BASELINE config 1.  The reference has NO 28x28 DCGAN (its 28x28
    scripts use the MLP above); SURVEY 8d C1 option (ii) defines this net for the config -- the 64x64 Discriminator's
    recipe ("#strainer gan.py:230-256": k4 s2 p1 convs without bias, BatchNorm from the second layer on, LeakyReLU .2,
    a final full-map conv + Sigmoid) cut down to 28x28 grayscale.  torch.nn on the CPU IS its oracle."""

    def __init__(self, ndf: int = 64):
        super().__init__()
        self.main = nn.Sequential(
            nn.Conv2d(1, ndf, 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 2), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(ndf * 2, 1, 7, 1, 0, bias=False), nn.Sigmoid())

    def forward(self, x):
        return self.main(x)


def make_discriminator28(seed: int = SEED) -> Discriminator28:
    """weights_init'ed DCGAN-28 D with non-trivial BN running statistics (as make_discriminator)."""
    g = torch.Generator().manual_seed(seed)
    d = Discriminator28()
    for m in d.modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data = torch.randn(m.weight.shape, generator=g) * 0.02
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data = 1.0 + torch.randn(m.weight.shape, generator=g) * 0.02
            m.bias.data.zero_()
            m.running_mean = torch.randn(m.running_mean.shape, generator=g) * 0.1
            m.running_var = torch.rand(m.running_var.shape, generator=g) + 0.5
    return d


def synth_images28(start: int, count: int, seed: int = SEED) -> np.ndarray:
    """Config-1 inputs: fp32 [count,1,28,28] in [-1,1], 80 % smooth "clean" images / 20 % iid noise --
    the centre 28x28 crop of channel 0 of the 64x64 counter-based stream (same sample keys, same noisy flags)."""
    return np.ascontiguousarray(synth_images(start, count, seed)[:, :1, 18:46, 18:46])


# --------------------------------------------------------------------------------------
# Scoring (a1+a2, a3, a4)
# --------------------------------------------------------------------------------------
def bce_vs_ones(p: torch.Tensor) -> torch.Tensor:
    """``nn.BCELoss(reduction='none')(p, ones)`` ("#strainer gan.py:369,375"):
    ``-max(log p, -100)``; p == 1 gives -0.0."""
    return F.binary_cross_entropy(p, torch.ones_like(p), reduction="none")


@torch.no_grad()
def score_losses(discriminator: nn.Module, images: torch.Tensor, batch_size: int = 64) -> np.ndarray:
    """Per-sample D loss vs label 1, eval-mode BN, batches of ``batch_size`` in index order:
    the scoring loop of ``refine_dataset_by_loss`` "#strainer gan.py:364-378" with the
    ``(B,1,1,1) -> .mean(dim=1)`` no-op removed; returns (N,) fp32."""
    discriminator.eval()
    out = []
    for i in range(0, images.shape[0], batch_size):
        p = discriminator(images[i:i + batch_size])
        out.append(bce_vs_ones(p).reshape(-1).numpy())
    return np.concatenate(out) if out else np.zeros((0,), np.float32)


def refine_select(losses: np.ndarray, loss_ratio: float = 0.2, n_dataset: int | None = None):
    """Selection half of ``refine_dataset_by_loss`` "#strainer gan.py:380-388".

    ``losses`` is given in the reference's (N,1,1) form or flat; the reference's array is
    (N,1,1) so the fallback ``np.argsort(losses)[:max(N//2,1)]`` sorts along the last axis
    (length 1) and yields all-zero indices of shape (k,1,1); that is
    reproduced when ``losses.ndim == 3``.  Returns (clean_indices, threshold)."""
    losses = np.asarray(losses)
    n = losses.shape[0] if n_dataset is None else n_dataset
    threshold = np.percentile(losses, (1 - loss_ratio) * 100)
    clean_indices = np.where(losses < threshold)[0]
    if len(clean_indices) == 0:
        clean_indices = np.argsort(losses)[:max(n // 2, 1)]
    return clean_indices, threshold


def refine_dataset_by_loss(images: torch.Tensor, discriminator: nn.Module, loss_ratio: float = 0.2):
    """``refine_dataset_by_loss`` "#strainer gan.py:364-392" on a resident image tensor.
    Returns (clean_indices int64, threshold np.float32, losses (N,1,1))."""
    losses = score_losses(discriminator, images, 64).reshape(-1, 1, 1)
    idx, thr = refine_select(losses, loss_ratio, images.shape[0])
    return idx, thr, losses


def evaluate_dataset(netD: nn.Module, images: torch.Tensor) -> np.ndarray:
    """``evaluate_dataset`` "#clean ... .py:272-287": same scoring at batch 128, returns (N,) fp32."""
    return score_losses(netD, images, 128)


# --------------------------------------------------------------------------------------
# Thresholds (a5, a6)
# --------------------------------------------------------------------------------------
def gmm_intersection(means: np.ndarray, stds: np.ndarray) -> float:
    """Closed-form clean/noisy Gaussian-pdf intersection, "#clean ... .py:298-307".  Evaluated in
    the dtype sklearn returns: float32 for float32 losses (scikit-learn >= 1.x keeps float32)."""
    means = np.asarray(means)
    stds = np.asarray(stds)
    ci = np.argmin(means)
    ni = 1 - ci
    a = 1 / (2 * stds[ci] ** 2) - 1 / (2 * stds[ni] ** 2)
    b = means[ni] / (stds[ni] ** 2) - means[ci] / (stds[ci] ** 2)
    c = means[ci] ** 2 / (2 * stds[ci] ** 2) - means[ni] ** 2 / (2 * stds[ni] ** 2) - np.log(stds[ni] / stds[ci])
    return (-b + np.sqrt(b ** 2 - 4 * a * c)) / (2 * a)


def gmm_fit(losses: np.ndarray):
    """2-component fit of "#clean ... .py:290-296" (sklearn; k-means init draws from the
    GLOBAL np.random state -- seed it before calling)."""
    from sklearn.mixture import GaussianMixture
    gmm = GaussianMixture(n_components=2, max_iter=10, tol=1e-2, reg_covar=5e-4)
    gmm.fit(np.asarray(losses).reshape(-1, 1))
    return gmm.means_.flatten(), np.sqrt(gmm.covariances_.flatten())


def get_gmm_threshold(losses):
    """"# 종합 loss.py:270-285"."""
    return gmm_intersection(*gmm_fit(losses))


def get_percentile_threshold(losses, percentile=75):
    """"# 종합 loss.py:287-288"."""
    return np.percentile(losses, percentile)


def get_iqr_threshold(losses):
    """"# 종합 loss.py:290-294"."""
    q1 = np.percentile(losses, 25)
    q3 = np.percentile(losses, 75)
    return q3 + 1.5 * (q3 - q1)


def get_ensemble_threshold(losses):
    """"# 종합 loss.py:296-301": median of the three thresholds."""
    return np.median([get_gmm_threshold(losses), get_percentile_threshold(losses), get_iqr_threshold(losses)])


def divide_by_threshold(losses: np.ndarray, threshold):
    """Mask + index split shared by both ``divide_dataset`` forms
    ("#clean ... .py:310-314", "# 종합 loss.py:306-310"): strict ``<``; ascending, complementary."""
    clean = np.asarray(losses).flatten() < threshold
    return np.where(clean)[0], np.where(~clean)[0]


# --------------------------------------------------------------------------------------
# Bit-level restatements of the library selection arithmetic
# --------------------------------------------------------------------------------------
def np_percentile_plan(n: int, q):
    """Index plan of ``np.percentile(a_f32, q)`` (numpy 2.3 ``_quantile``, method 'linear'):
    q/100 is evaluated in the *array's* dtype (``np.true_divide(q, a.dtype.type(100))``), the
    virtual index ``(n-1)*q`` therefore in fp32 for a python-float q.
    Returns (k_prev, k_next, gamma) with gamma in the virtual index's dtype."""
    qq = np.true_divide(q, np.float32(100))
    v = np.asanyarray((n - 1) * qq)
    prev = np.floor(v)
    nxt = prev + 1
    if v >= n - 1:
        prev = nxt = np.float64(n - 1)
    if v < 0:
        prev = nxt = np.float64(0)
    gamma = np.asanyarray(v - np.floor(v), dtype=v.dtype)
    if v >= n - 1 or v < 0:
        gamma = np.asanyarray(v - prev, dtype=v.dtype)
    return int(prev), int(nxt), gamma[()]


def np_lerp(a, b, t):
    """numpy ``_lerp``: ``a + (b-a)*t``, replaced by ``b - (b-a)*(1-t)`` when t >= 0.5; numpy
    scalar arithmetic (no FMA), result dtype by numpy promotion."""
    d = np.subtract(b, a)
    r = np.add(a, d * t)
    if t >= 0.5:
        r = np.subtract(b, d * (1 - t)).astype(r.dtype)
    return r


def np_percentile_f32(values: np.ndarray, q):
    """``np.percentile(values_f32, q)`` from two order statistics; NaN anywhere -> NaN."""
    values = np.asarray(values, dtype=np.float32).ravel()
    n = values.size
    k0, k1, g = np_percentile_plan(n, q)
    s = np.sort(values)  # NaNs last
    if np.isnan(s[-1]):
        return np.float32(np.nan)
    return np_lerp(s[k0], s[k1], g)


def _fma32(a, b, c) -> np.float32:
    """Correctly rounded fp32 fused multiply-add (exact rational arithmetic)."""
    a, b, c = float(np.float32(a)), float(np.float32(b)), float(np.float32(c))
    if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
        return np.float32(a * b + c)
    return np.float32(_round_fraction_to_f32(Fraction(a) * Fraction(b) + Fraction(c)))


def _round_fraction_to_f32(x: Fraction) -> float:
    if x == 0:
        return 0.0
    # float(Fraction) rounds correctly to double; round to fp32 from an exact representation
    # by scaling to a 24-bit significand with round-half-even.
    sign = -1 if x < 0 else 1
    x = abs(x)
    e = x.numerator.bit_length() - x.denominator.bit_length()
    if Fraction(2) ** e > x:
        e -= 1
    e = max(e, -126)
    scaled = x / (Fraction(2) ** (e - 23))
    fl = scaled.numerator // scaled.denominator
    rem = scaled - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (fl & 1)):
        fl += 1
    return sign * float(fl) * 2.0 ** (e - 23)


def torch_quantile_plan(n: int, q: float):
    """Index plan of ``torch.quantile(x_f32, q)`` (ATen ``quantile_compute``): the rank
    ``q * (n-1)`` is computed in fp32; weight = rank - floor(rank) in fp32."""
    rank = np.float32(q) * np.float32(n - 1)
    below = np.floor(rank)
    w = np.float32(rank - below)
    return int(below), int(np.ceil(rank)), w


def torch_lerp_f32(a, b, w) -> np.float32:
    """``torch.lerp`` for fp32: ``w < .5 ? fma(w, b-a, a) : fma(-(b-a), 1-w, b)`` (the ATen
    kernels contract to FMA; SURVEY §7 'bit-exact thresholds')."""
    a, b, w = np.float32(a), np.float32(b), np.float32(w)
    d = np.float32(b - a)
    if abs(w) < np.float32(0.5):
        return _fma32(w, d, a)
    return _fma32(-d, np.float32(np.float32(1) - w), b)


def torch_quantile_f32(values: np.ndarray, q: float) -> np.float32:
    values = np.asarray(values, dtype=np.float32).ravel()
    k0, k1, w = torch_quantile_plan(values.size, q)
    s = np.sort(values)
    if np.isnan(s[-1]):
        return np.float32(np.nan)
    return torch_lerp_f32(s[k0], s[k1], w)


def np_histogram_f32(values: np.ndarray, bins: int = 100):
    """Counts + edges of ``np.histogram(values_f32, bins)`` via numpy's uniform fast path
    (_histograms_impl.py): fp32 ``linspace`` edges, index ``((x-first)/(last-first))*bins`` in
    fp32, truncation, +-1 edge correction, last bin closed.  Returns (counts int64, edges f32)."""
    a = np.asarray(values, dtype=np.float32).ravel()
    first, last = a.min(), a.max()
    if not (np.isfinite(first) and np.isfinite(last)):
        raise ValueError(f"autodetected range of [{first}, {last}] is not finite")
    if first == last:
        first, last = first - 0.5, last + 0.5
    edges = np.linspace(first, last, bins + 1, endpoint=True, dtype=np.result_type(first, last, a))
    denom = last - first
    f = ((a - first) / denom) * bins
    idx = f.astype(np.intp)
    idx[idx == bins] -= 1
    idx[a < edges[idx]] -= 1
    inc = (a >= edges[idx + 1]) & (idx != bins - 1)
    idx[inc] += 1
    return np.bincount(idx, minlength=bins).astype(np.int64), edges


def elbow_from_hist(counts: np.ndarray, edges: np.ndarray):
    """Tail of ``find_elbow_threshold`` "#strainer gan.py:293-309" given the raw histogram:
    density normalisation as ``np.histogram(density=True)``, peak = argmax, target = first bin
    right of the peak whose density is closest to 0.01, threshold = mean of the two centres."""
    db = np.array(np.diff(edges), float)
    hist = counts / db / counts.sum()
    centers = (edges[:-1] + edges[1:]) / 2
    peak = np.argmax(hist)
    target = np.argmin(np.abs(hist[peak:] - 0.01))
    thr = (centers[peak] + centers[peak:][target]) / 2
    return thr, centers, hist


def find_elbow_threshold(z_scores, bins: int = 100):
    """``find_elbow_threshold`` "#strainer gan.py:291-309"."""
    hist, bin_edges = np.histogram(z_scores, bins=bins, density=True)
    bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2
    peak_index = np.argmax(hist)
    right_hist = hist[peak_index:]
    target_index = np.argmin(np.abs(right_hist - 0.01))
    threshold = (bin_centers[peak_index] + bin_centers[peak_index:][target_index]) / 2
    return threshold, bin_centers, hist


# --------------------------------------------------------------------------------------
# Feature z-score (a10, a11)
# --------------------------------------------------------------------------------------
def zscore_max_torch(features: torch.Tensor) -> torch.Tensor:
    """``detect_outliers`` core "#z_score.py:286-291": column mean, UNBIASED std, max |z| per row."""
    mean = features.mean(dim=0)
    std = features.std(dim=0)
    z = torch.abs((features - mean) / std)
    return z.max(dim=1)[0]


def zscore_max_numpy(features: np.ndarray) -> np.ndarray:
    """``compute_z_scores`` "# 1,2,8.py:164-168": np.std (ddof 0) + 1e-7."""
    mean = np.mean(features, axis=0)
    std = np.std(features, axis=0)
    return np.max(np.abs((features - mean) / (std + 1e-7)), axis=1)


def detect_outliers_fixed(features: torch.Tensor, threshold: float = 5.0) -> torch.Tensor:
    """"#z_score.py:276-294": strict ``<`` against a fixed threshold; torch bool."""
    return zscore_max_torch(features) < threshold


def detect_outliers_elbow(features: torch.Tensor, user_threshold=None):
    """"#strainer gan.py:331-360": user threshold or elbow threshold; numpy bool."""
    mz = zscore_max_torch(features).numpy()
    thr = find_elbow_threshold(mz)[0] if user_threshold is None else user_threshold
    return mz < thr, thr


def detect_outliers_ratio(features: torch.Tensor, clean_ratio: float) -> torch.Tensor:
    """"# z_score + DBSCAN.py:305-326": ``torch.quantile(max_z, clean_ratio)`` and ``<=``."""
    mz = zscore_max_torch(features)
    return mz <= torch.quantile(mz, clean_ratio)


# --------------------------------------------------------------------------------------
# DBSCAN clean ratio (a13)
# --------------------------------------------------------------------------------------
def estimate_ratio_dbscan_features(features: np.ndarray, eps=20, min_samples=3) -> float:
    """"# z_score + DBSCAN.py:291-299": StandardScaler -> DBSCAN -> fraction of non-noise."""
    from sklearn.cluster import DBSCAN
    from sklearn.preprocessing import StandardScaler
    labels = DBSCAN(eps=eps, min_samples=min_samples).fit_predict(StandardScaler().fit_transform(features))
    return np.sum(labels != -1) / len(labels)


def dbscan1d_noise_sklearn(values: np.ndarray, eps: float, min_samples: int) -> np.ndarray:
    """Noise mask of sklearn DBSCAN on an (N,1) array (the north_star's 1-D variant)."""
    from sklearn.cluster import DBSCAN
    return DBSCAN(eps=eps, min_samples=min_samples).fit_predict(np.asarray(values).reshape(-1, 1)) == -1


def dbscan1d_noise(values: np.ndarray, eps: float, min_samples: int) -> np.ndarray:
    """Sort-based restatement of DBSCAN noise membership in 1-D (what the CUDA kernel does):
    i is core iff #{j: |x_j - x_i| <= eps} >= min_samples (float64 distances as sklearn);
    i is noise iff it is not core and no core point lies within eps."""
    x = np.asarray(values, dtype=np.float32).ravel()
    order = np.argsort(x, kind="stable")
    s = x[order].astype(np.float64)
    n = s.size
    lo = np.searchsorted(s, s - eps, side="left")
    hi = np.searchsorted(s, s + eps, side="right")
    # searchsorted on s -/+ eps rounds the bound; fix up with the exact predicate |d| <= eps
    lo = np.clip(lo, 0, n - 1)
    while True:
        m = (lo > 0) & (s - s[np.maximum(lo - 1, 0)] <= eps)
        if not m.any():
            break
        lo[m] -= 1
    while True:
        m = s - s[lo] > eps
        if not m.any():
            break
        lo[m] += 1
    hi = np.clip(hi, 1, n)
    while True:
        m = (hi < n) & (s[np.minimum(hi, n - 1)] - s <= eps)
        if not m.any():
            break
        hi[m] += 1
    while True:
        m = s[hi - 1] - s > eps
        if not m.any():
            break
        hi[m] -= 1
    core = (hi - lo) >= min_samples
    idx = np.arange(n)
    prev_core = np.maximum.accumulate(np.where(core, idx, -1))
    next_core = np.minimum.accumulate(np.where(core, idx, n)[::-1])[::-1]
    near_prev = (prev_core >= 0) & (s - s[np.maximum(prev_core, 0)] <= eps)
    near_next = (next_core < n) & (s[np.minimum(next_core, n - 1)] - s <= eps)
    noise_sorted = ~(core | near_prev | near_next)
    noise = np.empty(n, dtype=bool)
    noise[order] = noise_sorted
    return noise


def dbscan1d_clean_ratio(values: np.ndarray, eps: float, min_samples: int = 3) -> float:
    noise = dbscan1d_noise(values, eps, min_samples)
    return np.sum(~noise) / len(noise)


# --------------------------------------------------------------------------------------
# Auto-encoder straining (a14)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def ae_errors(autoencoder: nn.Module, images: torch.Tensor, batch_size: int = 64) -> torch.Tensor:
    """Per-sample reconstruction MSE, "#autoencoder.py:311-318"."""
    autoencoder.eval()
    errs = []
    for i in range(0, images.shape[0], batch_size):
        img = images[i:i + batch_size]
        out = autoencoder(img)
        errs.append(F.mse_loss(out, img, reduction="none").view(img.size(0), -1).mean(dim=1))
    return torch.cat(errs, dim=0)


def detect_outliers_autoencoder(autoencoder: nn.Module, images: torch.Tensor, threshold: float = 2.0):
    """"#autoencoder.py:307-322": ``err < err.mean() + threshold * err.std()`` (unbiased std)."""
    e = ae_errors(autoencoder, images)
    thr = e.mean() + threshold * e.std()
    return e < thr, thr, e


# --------------------------------------------------------------------------------------
# In-batch strain + concat (a7, a8, a9): the reference has no function for these; the block
# below restates "# 상위 10% 제거해서 fake image에 concate.py:243-251, 265-269" literally.
# --------------------------------------------------------------------------------------
@torch.no_grad()
def strain_batch(netD: nn.Module, real: torch.Tensor, q: float = 0.1):
    """netD is used AS IS (train-mode BN under no_grad updates running stats, quirk 2)."""
    real_scores = netD(real).view(-1)
    threshold = torch.quantile(real_scores, q)
    mask = real_scores >= threshold
    return real[mask], real[~mask], mask, threshold, real_scores


def strain_scores(real: torch.Tensor, real_scores: torch.Tensor, q: float = 0.1):
    """Selection half of the block above, given the scores."""
    threshold = torch.quantile(real_scores, q)
    mask = real_scores >= threshold
    return real[mask], real[~mask], mask, threshold


def concat_fake(fake: torch.Tensor, strained: torch.Tensor) -> torch.Tensor:
    """``torch.cat([fake, filtered_fake], dim=0)`` ":268"."""
    return torch.cat([fake, strained], dim=0)


def sample_pool(pool: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """``potential_fake_data[indices]`` "# strainer gan + concate.py:623-624" for given indices
    (the reference draws ``torch.randperm(P)[:b]`` on the host)."""
    return pool[indices]


# --------------------------------------------------------------------------------------
# Training step (8f item 3): the D step and the G step's pass through D, "#strainer gan.py:586-615",
# restated on the CPU in fp32 (torch autograd IS the reference's arithmetic here).
# --------------------------------------------------------------------------------------
D64_PARAM_NAMES = ["w1", "w2", "w3", "w4", "w5", "g2", "b2", "g3", "b3", "g4", "b4"]


def d64_params(netD: nn.Module):
    """conv1..conv5 weights, then (gamma, beta) of the three BatchNorm2d, in D64_PARAM_NAMES order."""
    convs = [m for m in netD.modules() if isinstance(m, nn.Conv2d)]
    bns = [m for m in netD.modules() if isinstance(m, nn.BatchNorm2d)]
    return [c.weight for c in convs] + [t for bn in bns for t in (bn.weight, bn.bias)]


def train_step_through_d(netD: nn.Module, real: torch.Tensor, fake: torch.Tensor, real_label: float = 1.0,
                         fake_label: float = 0.0):
    """``netD.zero_grad(); netD(real) -> errD_real.backward()`` ":586-592", ``netD(fake.detach()) -> errD_fake.backward()``
    ":598-603", then (without an optimiser step in between, as the fixture does) ``netD(fake) -> errG.backward()``
    ":610-615".  netD is used as it is (train mode: batch statistics, running statistics updated three times).
    Returns the three output vectors, errD, errG, the accumulated D-step gradients and d errG / d fake.  The labels take the
    dtype of ``real`` (the reference's are float32): a float64 module / input gives the float64 truth the fp32 runs are
    measured against."""
    criterion = nn.BCELoss()
    fake = fake.detach().clone().requires_grad_(True)
    netD.zero_grad()
    label = torch.full((real.size(0),), real_label, dtype=real.dtype, device=real.device)
    out_real = netD(real).view(-1)
    errD_real = criterion(out_real, label)
    errD_real.backward()
    label = torch.full((fake.size(0),), fake_label, dtype=real.dtype, device=real.device)
    out_fake = netD(fake.detach()).view(-1)
    errD_fake = criterion(out_fake, label)
    errD_fake.backward()
    d_grads = [p.grad.detach().clone() for p in d64_params(netD)]
    label = torch.full((fake.size(0),), real_label, dtype=real.dtype, device=real.device)
    out_g = netD(fake).view(-1)
    errG = criterion(out_g, label)
    errG.backward()
    return dict(out_real=out_real.detach(), out_fake=out_fake.detach(), out_g=out_g.detach(),
                errD=(errD_real + errD_fake).detach(), errG=errG.detach(), d_grads=d_grads, dfake=fake.grad.detach())


def gmm_fit_deterministic(losses: np.ndarray, max_iter: int = 10, tol: float = 1e-2, reg_covar: float = 5e-4,
                          kmeans_iters: int = 30):
    """float64 restatement of the device EM (csrc/gmm.cu): scikit-learn's GaussianMixture equations
    (mixture/_gaussian_mixture.py: _estimate_gaussian_parameters, _estimate_log_prob_resp, the tol test on the
    mean log-likelihood) with Lloyd iterations from the 25 % / 75 % order statistics as the start -- the
    documented deviation from "#clean ... .py:290-292", whose k-means start draws from the global numpy RNG."""
    x = np.asarray(losses, np.float32).reshape(-1).astype(np.float64)
    n = x.size
    s = np.sort(np.asarray(losses, np.float32).reshape(-1))
    mu = np.array([s[(n - 1) // 4], s[(3 * (n - 1)) // 4]], np.float64)
    eps10 = 10 * np.finfo(np.float64).eps
    for left in range(kmeans_iters - 1, -1, -1):
        c1 = np.abs(x - mu[1]) < np.abs(x - mu[0])
        n0, n1 = float((~c1).sum()), float(c1.sum())
        m = np.array([x[~c1].sum() / n0 if n0 > 0 else mu[0], x[c1].sum() / n1 if n1 > 0 else mu[1]])
        moved = not np.array_equal(m, mu)
        if moved and left > 0:
            mu = m
            continue
        break
    r = np.stack([(~c1).astype(np.float64), c1.astype(np.float64)], 1)

    def m_step(r):
        nk = r.sum(0) + eps10
        mu = (r * x[:, None]).sum(0) / nk
        var = np.maximum((r * (x * x)[:, None]).sum(0) / nk - mu * mu, 0.0) + reg_covar
        return nk / n, mu, var
    w, mu, var = m_step(r)
    lb, n_iter, converged = -np.inf, 0, False
    for n_iter in range(1, max_iter + 1):
        lp = np.log(w) - 0.5 * (np.log(2 * np.pi) + np.log(var)) - (x[:, None] - mu) ** 2 * (0.5 / var)
        mx = lp.max(1)
        lse = mx + np.log(np.exp(lp - mx[:, None]).sum(1))
        r = np.exp(lp - lse[:, None])
        w, mu, var = m_step(r)
        new_lb = lse.mean()
        change, lb = new_lb - lb, new_lb
        if abs(change) < tol:
            converged = True
            break
    return {"weights": w, "means": mu, "stds": np.sqrt(var), "n_iter": n_iter, "converged": converged}
