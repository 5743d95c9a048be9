"""CPU checks of the drop-in boundary: the library builds, loads, and exports exactly the symbols
include/vbnmf.h declares.  No compute call is made here (no GPU in the build container)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "vbnmf.h")).read()
    return sorted(set(re.findall(r"VBNMF_API[^;(]*?\b((?:vbnmf|mlnmf)_\w+)\s*\(", txt)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("vbnmf_create", "vbnmf_set_state", "vbnmf_step", "vbnmf_run", "vbnmf_get_state",
                 "vbnmf_cluster_id", "mlnmf_run", "vbnmf_last_error", "vbnmf_destroy",
                 "vbnmf_attach_comm"):
        assert must in syms


def test_library_builds_and_exports_every_declared_symbol():
    from ccfindr_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    # the ctypes signature table covers the same set
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_cfg_struct_layout_matches_header():
    from ccfindr_b200._lib import VbnmfCfg
    # int, double, int[4], int, int, double with natural alignment
    assert VbnmfCfg.itmax.offset == 0 and VbnmfCfg.tol.offset == 8
    assert VbnmfCfg.hyper_update.offset == 16 and VbnmfCfg.n0.offset == 32
    assert VbnmfCfg.dn.offset == 36 and VbnmfCfg.fudge.offset == 40
    assert C.sizeof(VbnmfCfg) == 48


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import scipy.sparse as sp
    from ccfindr_b200 import _lib
    from ccfindr_b200.engine import Engine
    with pytest.raises(_lib.VbnmfError):
        Engine(sp.csc_matrix(np.eye(4)))
