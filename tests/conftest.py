import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_counts(name):
    """CSC count matrix stored by tests/golden/make_golden.py."""
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLDEN, name + "_counts.npz"))
    return sp.csc_matrix((z["data"].astype(np.float64), z["indices"], z["indptr"]),
                         shape=tuple(z["shape"]))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


RUN_CASES = {
    "run_c1s1_r3": "c1s1", "run_c1s2_r3": "c1s2", "run_c1s3_r3": "c1s3", "run_c1s1_r2": "c1s1",
    "run_c1s1_r5": "c1s1", "run_pbmc_r2": "pbmc", "run_pbmc_r3": "pbmc", "run_pbmc_r5": "pbmc",
    "run_c1s2_r3_conv": "c1s2", "run_tiny_r2_fixed": "tiny",
}
STEP_CASES = {"step_tiny_r2": "tiny", "step_c1s1_r3": "c1s1", "step_pbmc_r5_hyp": "pbmc"}


def run_kwargs(g):
    """Loop configuration a golden run was made with."""
    kw = dict(Itmax=int(g["cfg_Itmax"]), Tol=float(g["cfg_Tol"]), n0=int(g["cfg_n0"]),
              dn=int(g["cfg_dn"]))
    if "cfg_hyper_update_flags" in g:
        kw["hyper_update_flags"] = tuple(bool(v) for v in g["cfg_hyper_update_flags"])
    return kw


def hyper_dict(v):
    return dict(aw=float(v[0]), bw=float(v[1]), ah=float(v[2]), bh=float(v[3]))


@pytest.fixture(scope="session")
def have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
