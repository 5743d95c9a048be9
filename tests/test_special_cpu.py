"""The special functions of the CUDA kernels (ccfindr_b200/csrc/special.cuh) through their host
twin, against mpmath at 40 digits: psi and lgamma of the fused evaluation the posterior kernel uses
(one shift to x >= 10, two Stirling series), digamma and trigamma of the hyper-parameter update."""
import ctypes as C
import os
import subprocess

import mpmath
import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def twin():
    out = os.path.join(ROOT, "tests", "_build", "libspecial_twin.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    res = subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", out,
                          os.path.join(ROOT, "tests", "special_twin.cpp")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lib = C.CDLL(out)
    lib.twin_digamma.restype = C.c_double
    lib.twin_digamma.argtypes = [C.c_double]
    lib.twin_trigamma.restype = C.c_double
    lib.twin_trigamma.argtypes = [C.c_double]
    lib.twin_psi_lgamma.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return lib


def test_fused_psi_lgamma_against_mpmath(twin):
    mpmath.mp.dps = 40
    rng = np.random.default_rng(0)
    xs = np.concatenate([10.0 ** rng.uniform(-3, 6, size=400), [1e-3, 0.5, 1.0, 1.4616321449683623,
                         2.0, 9.999999, 10.0, 10.000001, 1e6], np.arange(1, 40) * 0.25])
    for x in xs:
        psi, lg = C.c_double(), C.c_double()
        twin.twin_psi_lgamma(float(x), C.byref(psi), C.byref(lg))
        d = float(mpmath.digamma(mpmath.mpf(float(x))))
        g = float(mpmath.loggamma(mpmath.mpf(float(x))))
        assert abs(psi.value - d) <= 4e-16 * abs(d) + 8e-16, (x, psi.value, d)
        assert abs(lg.value - g) <= 4e-16 * abs(g) + 4e-15, (x, lg.value, g)
        assert abs(twin.twin_digamma(float(x)) - d) <= 4e-16 * abs(d) + 8e-16


def test_trigamma_against_mpmath(twin):
    mpmath.mp.dps = 40
    for x in [0.01, 0.3, 1.0, 2.5, 9.9, 10.0, 57.0, 1e4]:
        t = float(mpmath.polygamma(1, mpmath.mpf(x)))
        assert abs(twin.twin_trigamma(x) - t) <= 1e-15 * abs(t)   # series truncation at x = 10: 7e-16
