import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names(prefix="fc_"):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in z.files:
        v = z[k]
        out[k] = torch.from_numpy(v) if v.ndim else v.item()
    return out


def rel_max(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    return float(torch.linalg.vector_norm((a - b).reshape(-1)) / torch.linalg.vector_norm(b.reshape(-1)).clamp_min(1e-30))


def assert_close_normwise(a, b, tol, what=""):
    """max|a-b| <= tol*max|b| and ||a-b||_2 <= tol*||b||_2 (element-wise relative error is
    meaningless on near-cancelling entries; the fp32 reference differs from its own fp64 run by ~3e-7 normwise)."""
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    rm, r2 = rel_max(a, b), rel_l2(a, b)
    assert rm <= tol and r2 <= tol, "%s: rel max %.3e, rel l2 %.3e > %.1e" % (what, rm, r2, tol)
