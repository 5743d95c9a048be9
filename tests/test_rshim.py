"""The R .Call shim (r-shim/src/vbnmf_shim.c) compiled, linked against libvbnmf.so and executed.

R is not installed in this environment, so the shim is built against stand-in headers with R's own
prototypes and a minimal object runtime (tests/rstub/).  Checked without a GPU: it compiles with
-Wall -Werror against include/vbnmf.h, registers its routines the way the reference's
src/RcppExports.cpp:24-32 does (R_registerRoutines + R_useDynamicSymbols(FALSE)), and every
`.Call(C_xxx, ...)` in r-shim/R/vb_gpu.R names a registered routine with the right number of
arguments.  With a GPU: one factorization driven through the shim's entry points reproduces the
golden vectors made by the reference's own update code."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, hyper_dict, load_counts, load_golden, relerr, run_kwargs

STUB = os.path.join(ROOT, "tests", "rstub")
OUT = os.path.join(STUB, "_build", "libshimtest.so")
SHIM_C = os.path.join(ROOT, "r-shim", "src", "vbnmf_shim.c")
SHIM_R = os.path.join(ROOT, "r-shim", "R", "vb_gpu.R")


@pytest.fixture(scope="module")
def shim():
    from ccfindr_b200 import _lib, build as vb_build
    vb_build.build()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-shared", "-fPIC", "-O1", "-Wall", "-Wextra", "-Werror",
           "-Wno-cast-function-type",   # the (DL_FUNC) casts of every R registration table
           "-I", STUB,
           "-I", os.path.join(ROOT, "include"), SHIM_C, os.path.join(STUB, "rstub.c"),
           "-L", libdir, "-lvbnmf", "-Wl,-rpath," + libdir, "-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lib = C.CDLL(OUT)
    lib.rstub_init()
    for name in ("rstub_real", "rstub_int", "rstub_data", "Rf_ScalarReal", "Rf_ScalarInteger",
                 "VECTOR_ELT"):
        getattr(lib, name).restype = C.c_void_p
    lib.rstub_real.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int]
    lib.rstub_int.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int]
    lib.rstub_data.argtypes = [C.c_void_p]
    lib.Rf_ScalarReal.argtypes = [C.c_double]
    lib.Rf_ScalarInteger.argtypes = [C.c_int]
    lib.VECTOR_ELT.argtypes = [C.c_void_p, C.c_ssize_t]
    lib.XLENGTH.restype = C.c_ssize_t
    lib.XLENGTH.argtypes = [C.c_void_p]
    lib.R_init_ccfindRgpu(None)
    yield lib
    lib.rstub_free_all()


def registered(lib):
    out = {}
    name, nargs, fun = C.c_char_p(), C.c_int(), C.c_void_p()
    i = 0
    while lib.rstub_registered(i, C.byref(name), C.byref(nargs), C.byref(fun)):
        out[name.value.decode()] = (nargs.value, fun.value)
        i += 1
    return out


def call_sites(text):
    """(routine, number of arguments after it) of every .Call(...) in an R source."""
    sites = []
    for m in re.finditer(r"\.Call\(", text):
        depth, i, args, cur = 1, m.end(), [], ""
        while depth:
            ch = text[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            if ch == "," and depth == 1:
                args.append(cur); cur = ""
            else:
                cur += ch
            i += 1
        args.append(cur)
        sites.append((args[0].strip(), len(args) - 1))
    return sites


def test_shim_compiles_and_registers_like_the_reference(shim):
    reg = registered(shim)
    assert shim.rstub_dynamic_symbols() == 0          # R_useDynamicSymbols(dll, FALSE)
    assert len(reg) >= 16
    for name, (nargs, fun) in reg.items():
        sym = C.cast(getattr(shim, name), C.c_void_p).value
        assert sym == fun, name                        # the table points at the exported routine
        assert 0 <= nargs <= 8


def test_every_dot_call_matches_a_registered_routine(shim):
    reg = registered(shim)
    sites = call_sites(open(SHIM_R).read())
    assert len(sites) >= 10
    for name, nargs in sites:
        assert name in reg, name
        assert reg[name][0] == nargs, (name, nargs, reg[name][0])
    # nothing unexported by ccfindR is called unqualified (vb_init is internal: NAMESPACE has no export)
    src = open(SHIM_R).read()
    assert "ccfindR:::vb_init(" in src and not re.search(r"(?<![:\w])vb_init\(", src)


@pytest.mark.gpu
def test_factorization_through_the_shim_matches_the_reference_golden(shim):
    g = load_golden("run_pbmc_r3")
    X = load_counts("pbmc")
    kw = run_kwargs(g)
    flags = kw.pop("hyper_update_flags", (True,) * 4)
    n, m = X.shape
    r = g["w0"].shape[1]

    def real(a, nrow=0, ncol=0):
        a = np.asfortranarray(a, dtype=np.float64)
        return C.c_void_p(shim.rstub_real(a.ctypes.data_as(C.c_void_p), a.size, nrow, ncol))

    def ints(a, logical=0):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return C.c_void_p(shim.rstub_int(a.ctypes.data_as(C.c_void_p), a.size, logical))

    def vec(sexp, dtype, count):
        buf = (C.c_char * (np.dtype(dtype).itemsize * count)).from_address(shim.rstub_data(sexp))
        return np.frombuffer(buf, dtype=dtype).copy()

    for f in ("C_vbnmf_create", "C_vbnmf_run", "C_vbnmf_get_state", "C_vbnmf_cluster_id",
              "C_vbnmf_uniform_columns", "C_vbnmf_set_state", "C_vbnmf_destroy"):
        getattr(shim, f).restype = C.c_void_p
    h = C.c_void_p(shim.C_vbnmf_create(ints(X.indptr), ints(X.indices), real(X.data),
                                       ints([n, m]), ints([0])))
    nil = C.c_void_p.in_dll(shim, "R_NilValue")
    shim.C_vbnmf_set_state(h, real(g["w0"], n, r), real(g["h0"], r, m), nil, nil)
    hy = g["hyper0"]
    res = C.c_void_p(shim.C_vbnmf_run(h, real(hy), ints([kw["Itmax"]]), real([kw["Tol"]]),
                                      ints([int(f) for f in flags], logical=1), ints([kw["n0"]]),
                                      ints([kw["dn"]]), real([float(np.finfo(np.float64).eps)])))
    elt = lambda lst, i: C.c_void_p(shim.VECTOR_ELT(lst, i))
    niter = int(vec(elt(res, 2), np.int32, 1)[0])
    assert niter == int(g["niter"])
    assert relerr(vec(elt(res, 1), np.float64, 1)[0], g["lml"]) < 1e-9
    assert relerr(vec(elt(res, 0), np.float64, 4), g["hyper_trace"][niter - 1]) < 1e-9
    assert shim.XLENGTH(elt(res, 4)) == niter
    assert relerr(vec(elt(res, 4), np.float64, niter), g["lkh_trace"]) < 1e-9
    st = C.c_void_p(shim.C_vbnmf_get_state(h))
    for i, k in enumerate(("lw", "lh", "ew", "eh", "dw", "dh")):
        shape = (n, r) if k[1] == "w" else (r, m)
        got = vec(elt(st, i), np.float64, shape[0] * shape[1]).reshape(shape, order="F")
        assert relerr(got, g[k]) < 1e-9, k
    cid = vec(C.c_void_p(shim.C_vbnmf_cluster_id(h)), np.int32, m)
    assert np.array_equal(cid, g["cid"])
    unif = vec(C.c_void_p(shim.C_vbnmf_uniform_columns(h, real([kw["Tol"]]))), np.int32, r)
    assert not unif.any()
    shim.C_vbnmf_destroy(h)


@pytest.mark.gpu
def test_remaining_shim_entry_points_execute(shim):
    """C_vbnmf_init_random, C_vbnmf_step, C_vbnmf_set_precision, C_vbnmf_set_host_threads,
    C_mlnmf_run and C_vbnmf_nccl_unique_id through the stand-in runtime, against the ctypes path."""
    from ccfindr_b200 import synth
    from ccfindr_b200.engine import Engine
    X = load_counts("c1s1")
    n, m = X.shape
    r = 3
    hyper = np.array([1.0, 1.0, 1.0, 1.0])

    def real(a, nrow=0, ncol=0):
        a = np.asfortranarray(a, dtype=np.float64)
        return C.c_void_p(shim.rstub_real(a.ctypes.data_as(C.c_void_p), a.size, nrow, ncol))

    def ints(a, logical=0):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return C.c_void_p(shim.rstub_int(a.ctypes.data_as(C.c_void_p), a.size, logical))

    def vec(sexp, dtype, count):
        buf = (C.c_char * (np.dtype(dtype).itemsize * count)).from_address(shim.rstub_data(sexp))
        return np.frombuffer(buf, dtype=dtype).copy()

    for f in ("C_vbnmf_create", "C_vbnmf_step", "C_vbnmf_get_state", "C_mlnmf_run",
              "C_vbnmf_nccl_unique_id", "C_vbnmf_init_random", "C_vbnmf_set_precision",
              "C_vbnmf_set_host_threads", "C_vbnmf_destroy"):
        getattr(shim, f).restype = C.c_void_p
    shim.C_vbnmf_set_host_threads(ints([4]))
    h = C.c_void_p(shim.C_vbnmf_create(ints(X.indptr), ints(X.indices), real(X.data),
                                       ints([n, m]), ints([0])))
    shim.C_vbnmf_set_precision(h, ints([0]))
    shim.C_vbnmf_init_random(h, ints([r]), real(hyper), real([7.0]))
    lkh = vec(C.c_void_p(shim.C_vbnmf_step(h, real(hyper), real([float(np.finfo(float).eps)]))),
              np.float64, 1)[0]
    with Engine(X) as eng:                                      # same calls through ctypes
        eng.init_random(r, hyper, 7)
        ref = eng.step(hyper)
        w0, h0 = synth.uniform_init(n, m, r, 3)
        mref = eng.ml_run(w0, h0, Itmax=12, Tol=1e-7)
    assert lkh == ref
    out = C.c_void_p(shim.C_mlnmf_run(h, real(w0, n, r), real(h0, r, m), ints([12]), real([1e-7])))
    niter = int(vec(C.c_void_p(shim.VECTOR_ELT(out, 3)), np.int32, 1)[0])
    assert niter == mref["niter"]
    wgot = vec(C.c_void_p(shim.VECTOR_ELT(out, 0)), np.float64, n * r).reshape((n, r), order="F")
    assert np.array_equal(wgot, mref["w"])
    uid = C.c_void_p(shim.C_vbnmf_nccl_unique_id())
    assert shim.XLENGTH(uid) == 128 and vec(uid, np.uint8, 128).any()
    shim.C_vbnmf_destroy(h)
