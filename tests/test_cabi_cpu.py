"""CPU: the C-ABI library builds, loads, and exports every symbol include/strainer_b200.h declares;
host-side logic that needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sb():
    import strainer_b200
    return strainer_b200


def declared_symbols(release_only=True):
    src = open(os.path.join(ROOT, "include", "strainer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    if release_only:     # declarations of the experiment build (-DSG_AB_VARIANTS) are not part of the release library
        src = re.sub(r"#ifdef SG_AB_VARIANTS.*?#endif", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(sb):
    lib = ctypes.CDLL(sb._lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding covers exactly the header (the experiment-build entry points are optional bindings)
    assert sorted(set(sb._lib._SIGS) - sb._lib._OPTIONAL) == names
    assert sorted(sb._lib._SIGS) == declared_symbols(release_only=False)


def test_no_gpu_fails_loudly(sb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = sb._lib.load()
    assert lib.sg_version() == 100
    assert lib.sg_init(0) == -2  # SG_EARCH: no device, no fallback
    assert b"no CPU fallback" in lib.sg_last_error_string()
    with pytest.raises(RuntimeError):
        sb.synth_images(0, 4)
    with pytest.raises(RuntimeError):
        sb.refine_dataset_by_loss(torch.zeros(4, 3, 64, 64), torch.nn.Identity(), "cpu")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "strainer-gan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_percentile_plan_matches_numpy(sb):
    from oracle import strainer_oracle as O
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.integers(1, 5000))
        q = [float(rng.uniform(0, 100)), (1 - 0.8) * 100, 90.0, 75, 0, 100][int(rng.integers(0, 6))]
        v = np.sort(rng.standard_normal(n).astype(np.float32))
        k0, k1, g, dt = sb.api._np_percentile_plan(n, q) if hasattr(sb, "api") else sb._np_percentile_plan(n, q)
        assert dt == np.float32
        assert O.np_lerp(v[k0], v[k1], g) == np.percentile(v, q)


def test_f32_threshold_directed_rounding(sb):
    f = sb.api._f32_threshold if hasattr(sb, "api") else sb._f32_threshold
    rng = np.random.default_rng(4)
    v = rng.standard_normal(20000).astype(np.float32)
    for t in list(rng.standard_normal(50)) + [float(v[3]), float(np.float64(v[7]) + 1e-12)]:
        t = np.float64(t)
        assert np.array_equal(v < f(t, 0), v < t)
        assert np.array_equal(v <= f(t, 1), v <= t)
        assert np.array_equal(v >= f(t, 2), v >= t)
        assert np.array_equal(v > f(t, 3), v > t)


def test_entry_scripts_compile():
    """bench.py, __graft_entry__.py and the tools byte-compile (a `global` after use is a SyntaxError only at compile time)."""
    import glob
    import os
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] + glob.glob(os.path.join(root, "tools", "*.py")):
        py_compile.compile(path, doraise=True)


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` is pure CPU (oracle port): one tiny step must print a JSON line with the contract keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SG_BENCH_REF_SAMPLES="128")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "strained_samples_per_sec"
    for k in ("value", "unit", "cpu_baseline", "e2e", "config", "higher_is_better"):
        assert k in line


def test_trainable_d64_fails_loudly_without_gpu():
    """the training-step wrapper has no CPU or autograd fallback: CPU input (or a foreign module) raises"""
    import torch
    import strainer_b200 as sb
    from oracle import strainer_oracle as O
    d = O.make_discriminator(O.SEED).train()
    w = sb.accelerate_discriminator(d)
    assert sb.accelerate_discriminator(d) is w                      # one wrapper (packed weights, workspaces) per module
    assert [id(p) for p in w.parameters()] == [id(p) for p in d.parameters()]
    with pytest.raises(RuntimeError):
        w(torch.zeros(4, 3, 64, 64))
    with pytest.raises(NotImplementedError):
        sb.accelerate_discriminator(torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3)))
    w.eval()
    with torch.no_grad():                                           # eval mode: the wrapped module itself
        assert torch.equal(w(torch.zeros(2, 3, 64, 64)), d(torch.zeros(2, 3, 64, 64)))
