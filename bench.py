#!/usr/bin/env python
"""Benchmark of the VB-NMF hot path (BASELINE.json metric: nnz*rank updates/s per VB iteration).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (rank 0)

Default workload = the north-star configuration C3 under STRONG scaling: 20,000 genes x 1,300,000
cells, ~8 % nonzero 10x-shaped Poisson counts (SURVEY.md 8d generator), rank 20, fp64, the cells
sharded over the N ranks by (expected) nonzero count, one NCCL all-reduce of the W-side statistics
per iteration.  It fits one B200 (2.07e9 nonzeros), so N = 1 runs the same matrix.  BASELINE
config 2 (20k x 100k cells per GPU, rank 10, weak scaling) is measured in the same run and nested
under "secondary".

A "step" is one VB iteration of the product loop (vbnmf_run's device-controlled loop): posterior
update of W and H, the two passes of the nonzero sweep, the all-reduce, the lower bound and
hyper_update + stop rules in the control kernel; the host reads the control block back once per 8
iterations.  Timed iterations run with hyper_update after every iteration (the loop's steady state
past hyper.update.n0 = 10, R/bayesian.R:342); the untimed warm-up holds the hypers fixed.

The line also carries: `e2e` (host CSC -> vbnmf_create -> set_state -> vbnmf_run(K) -> get_state,
cold and warm pass), `parity` (the same matrix through oracle/oracle_sparse.c on the host cores:
per-iteration bound, factors, cluster ids; at N > 1 the sharded bound against a single-GPU run of
the whole matrix), `cpu_baseline`, `roofline` and the fp32-storage mode.  One JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    del os.environ["NCCL_DEBUG"]  # keeps NCCL's banner off stdout (one JSON line is the contract)

WORKLOADS = {
    # name: genes, cells (per GPU = weak scaling, total = strong), true rank, density, seed, fit rank
    "c2": dict(n=20000, m_per_gpu=100000, r_true=10, density=0.08, seed=2, rank=10, ml_rank=15,
               label="C2: vb_factorize rank=10, 20k genes x 100k cells per GPU, ~8% nonzero"),
    "c3": dict(n=20000, m_total=1300000, r_true=20, density=0.08, seed=3, rank=20,
               label="C3: vb_factorize rank=20, 20k genes x 1.3M cells, ~8% nonzero, cells "
                     "sharded over the GPUs (strong scaling)"),
    "small": dict(n=2000, m_per_gpu=8000, r_true=5, density=0.08, seed=4, rank=6,
                  label="small: plumbing check, 2k genes x 8k cells per GPU"),
    "smallstrong": dict(n=2000, m_total=32000, r_true=5, density=0.08, seed=4, rank=6,
                        label="smallstrong: plumbing check, 2k genes x 32k cells sharded"),
}
HYPER = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)  # gamma.a = gamma.b = 1 (R/bayesian.R:231)
N0 = 10                                        # hyper.update.n0 (R/bayesian.R:233)
METRIC = "VB-NMF nnz*rank updates/s per iteration"
UNIT = "nnz*rank updates/s"
SM_COUNT, SMEM_BYTES_PER_CLK = 148, 128        # shared-memory pipe: 128 B/clk/SM


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(nnz, n, m, r, sp=8):
    """B_alg = nnz*(4+4) + 8*(m+1) + 2*r*m*s_p + 2*n*r*s_p bytes per VB iteration,
    s_p = 8 (fp64) or 4 (fp32-storage mode)."""
    return nnz * 8 + 8 * (m + 1) + 2 * r * m * sp + 2 * n * r * sp


def measured_traffic(name):
    """dram bytes of the two sweep launches from the last `ncu --set full` capture
    (profiles/traffic.json, written by profiles/ncu_traffic.py).  The file is stamped with the
    hash of the kernel sources it was captured for; a stale stamp returns None."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return None, "no capture"
    t = json.load(open(tp))
    try:
        from ccfindr_b200 import build as vb_build
        if t.get("source_hash") != vb_build.kernel_hash():
            return None, "stale: profiles/traffic.json was captured for other kernel sources"
    except Exception as e:  # pragma: no cover
        return None, "cannot verify: %s" % e
    return t.get(name), t.get("note")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread
    (nvidia_ml_py, one sample per millisecond); falls back to an `nvidia-smi -lms` subprocess when
    NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = self.f = self.thread = None
        self.sm, self.reasons, self.smax, self.stop_flag = [], set(), None, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # CUDA_VISIBLE_DEVICES may renumber: resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(gpu_index)
            self.h = None
            try:
                pci = "%08x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(pci.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                "sw_power_cap": 0x4}
        while True:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = int(get(self.h))
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
            except Exception:
                pass
            if self.stop_flag:
                return
            time.sleep(0.001)

    def start(self):
        if self.nvml is not None:
            import threading
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if not self.sm:
                return None
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.smax,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.p is None:
            return None
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(",") for ln in open(self.f.name) if ln.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def env_rank():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def restore_omp_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arms use all host cores."""
    from oracle import bindings as ob
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    ob.set_omp_threads(n)
    return n


def shard_plan(wl, m_total, world, dev):
    """Contiguous cell ranges in whole generator chunks with (nearly) equal EXPECTED numbers of
    nonzeros (nnz-balanced), the same on every rank."""
    from ccfindr_b200 import sharding, synth
    nch = -(-m_total // synth.TENX_CHUNK)
    if world == 1:
        return [0, m_total], "single shard"
    exp_nnz = synth.tenx_expected_nnz(wl["n"], m_total, wl["r_true"], wl["density"], wl["seed"], dev)
    pseudo = np.concatenate([[0.0], np.cumsum(exp_nnz)])
    cb = sharding.balanced_bounds(pseudo, world)
    for i in range(1, world):           # every rank gets at least one chunk
        cb[i] = max(cb[i], cb[i - 1] + 1)
    for i in range(world - 1, 0, -1):
        cb[i] = min(cb[i], cb[i + 1] - 1)
    assert cb[0] == 0 and cb[-1] == nch and all(cb[i] < cb[i + 1] for i in range(world))
    return [min(m_total, c * synth.TENX_CHUNK) for c in cb], \
        "nnz-balanced (expected nonzeros per %d-cell generator chunk)" % synth.TENX_CHUNK


def init_factors(n, m_total, rank, seed):
    from ccfindr_b200 import synth
    return synth.random_init(n, m_total, rank, HYPER, seed)  # vb_init 'random'


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


# ------------------------------------------------------------------------------------------------
def gpu_parity_run(eng, w0, h0_loc, iters, precision):
    """`iters` iterations of the product loop from (w0, h0) on the engine: bound per iteration,
    ew, eh and the cluster ids of the local cells."""
    eng.set_precision(precision)
    eng.set_state(w0, h0_loc)
    out = eng.run(HYPER, Itmax=iters, Tol=0.0)   # Tol = 0 never converges: exactly `iters`
    st = eng.get_state(("ew", "eh"))
    return dict(lkh=out["lkh_trace"], ew=st["ew"], eh=st["eh"], cid=eng.cluster_id())


def oracle_parity(n, m, colptr, rowidx, values, w0, h0, iters, gpu64, gpu32):
    """The same matrix and start through oracle/oracle_sparse.c (fp64, OpenMP) on the host."""
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    threads = restore_omp_threads()
    t0 = time.time()
    ref = ob.sparse_vb_run((n, m, colptr, rowidx, values), w0, h0, HYPER, Itmax=iters, Tol=0.0)
    sec = time.time() - t0
    cid = od.cluster_id(ref["eh"])
    blk = {"oracle": "oracle/oracle_sparse.c (fp64, OpenMP, %d threads) on the WHOLE matrix "
                     "(%d cells, %d nonzeros), %d iterations from the same w0, h0"
                     % (threads, m, len(rowidx), iters),
           "iterations": iters, "oracle_seconds": round(sec, 2),
           "lkh_oracle": [float(v) for v in ref["lkh_trace"]]}
    for key, g in (("fp64", gpu64), ("fp32_storage", gpu32)):
        if g is None:
            continue
        blk[key] = {"lkh": [float(v) for v in g["lkh"]],
                    "lkh_rel_err": [abs(a - b) / abs(b) for a, b in zip(g["lkh"], ref["lkh_trace"])],
                    "ew_max_rel_err": relerr(g["ew"], ref["ew"]),
                    "eh_max_rel_err": relerr(g["eh"], ref["eh"]),
                    "cid_mismatches": int(np.count_nonzero(g["cid"] != cid)), "cells": int(m)}
    blk["lkh_rel_err"] = max(blk["fp64"]["lkh_rel_err"])
    blk["cid_mismatches"] = blk["fp64"]["cid_mismatches"]
    blk["tolerance"] = {"fp64": 1e-9, "fp32_storage": 1e-4}
    cpu = {"value": len(rowidx) * w0.shape[1] / (sec / (iters + 1)), "unit": UNIT, "cores": threads,
           "kind": "port",
           "sample": "whole matrix (%d nonzeros): statistics pass + %d iterations of "
                     "oracle_sparse.c (CSC, OpenMP, fp64), %.1f s" % (len(rowidx), iters, sec),
           "seconds_per_iteration": sec / (iters + 1)}
    return blk, cpu


def cpu_baseline_port(n, r, colptr, rowidx, values, w0, h0, max_cols=24000, iters=2):
    """oracle/oracle_sparse.c (OpenMP, all host threads) on the first max_cols cells."""
    from oracle import bindings as ob
    restore_omp_threads()
    mc = min(max_cols, len(colptr) - 1)
    end = int(colptr[mc])
    arrays = (n, mc, np.ascontiguousarray(colptr[:mc + 1], dtype=np.int64),
              np.ascontiguousarray(rowidx[:end], dtype=np.int32),
              np.ascontiguousarray(values[:end], dtype=np.float64))
    hy = np.array([HYPER[k] for k in ("aw", "bw", "ah", "bh")])
    sec, lkh, threads = ob.sparse_time_iterations(arrays, w0, h0[:, :mc], hy,
                                                  np.finfo(np.float64).eps, iters)
    return {"value": end * r / sec, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d cells (%d nnz) of rank 0's shard, %d iterations of "
                      "oracle_sparse.c (CSC, OpenMP, fp64)" % (mc, end, iters),
            "seconds_per_iteration": sec}


def ml_leg(eng, n, m, nnz, r, peak, k1=5, k2=25, parity_iters=3):
    """BASELINE config 5: factorize()'s maximum-likelihood loop (R/factorize.R:189-212) on the
    engine's matrix.  mlnmf_run is called through the C ABI with k1 and k2 > k1 iterations
    (Tol = 0: the likelihood rule never fires); the per-iteration time is the difference quotient,
    so the upload of w0/h0 and the download of w/h drop out.  One iteration = cell-owner sweep +
    h update + gene-owner sweep + w update + the likelihood and stopping rule on the device."""
    import torch
    from ccfindr_b200 import synth
    w0, h0 = synth.uniform_init(n, m, r, seed=5)              # init(): w, h ~ U(0, 1) (:30-38)
    eng.set_precision(0)
    g = eng.ml_run(w0, h0, Itmax=parity_iters, Tol=0.0)      # warm-up + the state the oracle checks
    ts = {}
    for k in (k1, k2, k1, k2):
        torch.cuda.synchronize()
        t0 = time.time()
        res = eng.ml_run(w0, h0, Itmax=k, Tol=0.0)
        torch.cuda.synchronize()
        ts.setdefault(k, []).append(time.time() - t0)
        assert res["niter"] == k and np.isfinite(res["lik"])
    t_iter = (min(ts[k2]) - min(ts[k1])) / (k2 - k1)
    b_alg = 2 * nnz * 8 + 2 * 8 * (m + 1) + 3 * r * m * 8 + 3 * n * r * 8    # SURVEY.md 8(d), ML row
    return {"workload": "C5: factorize() ML path, rank %d, on the same matrix" % r,
            "value": nnz * r / t_iter, "unit": UNIT, "ms_per_iteration": t_iter * 1e3,
            "algorithmic_bytes_per_iteration": b_alg, "roofline_frac": b_alg / t_iter / 1e9 / peak,
            "timing": "wall-clock difference quotient of mlnmf_run(%d) and mlnmf_run(%d) through the "
                      "C ABI (host buffers in and out)" % (k1, k2),
            "_gpu": g}


def run_config(args, name, ctx):
    """One workload through every leg.  Returns the record (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from ccfindr_b200 import synth
    from ccfindr_b200.engine import Engine
    rank, local_rank, world, dev, comm = ctx
    wl = WORKLOADS[name]
    n, r = wl["n"], wl["rank"]
    strong = "m_total" in wl
    m_total = wl["m_total"] if strong else wl["m_per_gpu"] * world
    bounds, shard_how = shard_plan(wl, m_total, world, dev)
    c0, c1 = bounds[rank], bounds[rank + 1]
    t_gen = time.time()
    colptr, rowidx, values, scale = synth.tenx_like_device(n, m_total, wl["r_true"], wl["density"],
                                                           wl["seed"], dev, c0, c1)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    m_loc, nnz_loc = c1 - c0, int(rowidx.numel())
    w0, h0 = init_factors(n, m_total, r, seed=1000 * r + 1)
    h0_loc = np.asfortranarray(h0[:, c0:c1])
    P = args.parity_iters

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: inputs already in HBM when the timed region starts --------------
    eng = Engine.from_device_csc(n, m_loc, nnz_loc, colptr, rowidx, values, device=local_rank)
    if comm is not None:
        eng.attach_comm(comm)
    nnz_t = torch.tensor([float(nnz_loc)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz_total = int(nnz_t.item())

    def timed(precision, sample_clocks):
        eng.set_precision(precision)
        eng.set_state(w0, h0_loc)
        # untimed: hyper.update.n0 iterations with fixed hypers (at least the requested warm-up)
        eng.bench_iterations(HYPER, max(args.warmup, N0), hyper_on=False)
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        barrier()
        if sampler:
            sampler.start()
        t0 = time.time()
        res = eng.bench_iterations(HYPER, args.steps, hyper_on=True)   # EXACTLY K iterations
        torch.cuda.synchronize()
        wall = (time.time() - t0) * 1e3
        barrier()
        clocks = sampler.stop() if sampler else None
        tot, tc, tr, wall = allmax([res["ms_total"], res["ms_cols"], res["ms_rows"], wall])
        return dict(res=res, ms_step=tot / args.steps, ms_cols=tc / args.steps,
                    ms_rows=tr / args.steps, wall_ms=wall / args.steps, clocks=clocks)

    t64 = timed(0, True)
    layout = eng.layout_info()
    value = nnz_total * r / (t64["ms_step"] * 1e-3)
    gp64 = gpu_parity_run(eng, w0, h0_loc, P, 0) if P else None
    t32 = timed(1, False)
    layout32 = eng.layout_info()
    gp32 = gpu_parity_run(eng, w0, h0_loc, P, 1) if P else None
    eng.set_precision(0)
    peak, peak_src = peaks()
    b32 = algorithmic_bytes(nnz_loc, n, m_loc, r, sp=4)
    mixed = {"value": nnz_total * r / (t32["ms_step"] * 1e-3), "unit": UNIT,
             "ms_per_step": t32["ms_step"],
             "ms_per_launch": {"sweep_cols": t32["ms_cols"], "sweep_rows": t32["ms_rows"]},
             "roofline_frac": b32 / ((t32["ms_cols"] + t32["ms_rows"]) * 1e-3) / 1e9 / peak,
             "algorithmic_bytes_per_launch": b32, "lkh_last": t32["res"]["lkh"],
             "tile_rows": layout32["tile_rows"],
             "what": "panels lw/lh held in fp32, per-nonzero arithmetic fp32, every sum over "
                     "lanes/slabs/ranks and the posterior update in fp64 (tolerance 1e-4)"}
    # ---- BASELINE config 5 on the same matrix: the ML path of factorize(), rank 15 ----------------
    ml = None
    if wl.get("ml_rank") and world == 1:
        ml = ml_leg(eng, n, m_loc, nnz_loc, wl["ml_rank"], peak)
    eng.close()

    # ---- end-to-end arm: HOST buffers through the C ABI, copies inside the timed region ---------
    e2e = parity = cpu = None
    h_colptr = h_rowidx = h_values = None
    if not args.no_e2e or (P and world == 1):
        h_colptr = colptr.cpu().numpy()
        h_rowidx = rowidx.cpu().numpy()
        h_values = values.double().cpu().numpy()                  # dgCMatrix @x is double
    del colptr, rowidx, values
    torch.cuda.empty_cache()
    if not args.no_e2e:
        import scipy.sparse as sp
        csc = sp.csc_matrix((h_values, h_rowidx, h_colptr), shape=(n, m_loc))
        csc.has_sorted_indices = True                             # generator output is sorted

        def e2e_pass():
            barrier()
            t0 = time.time()
            eng2 = Engine(csc, device=local_rank)                 # H2D of X + device layouts
            if comm is not None:
                eng2.attach_comm(comm)
            eng2.set_state(w0, h0_loc)                            # H2D of the initial factors
            # Tol = 0 never satisfies |1 - lkh/lk0| < Tol: exactly K iterations of the reference's
            # own loop (hyper updates from iteration 11 on, R/bayesian.R:342)
            out = eng2.run(HYPER, Itmax=args.steps, Tol=0.0)
            st = eng2.get_state(("ew", "eh"))                     # D2H of the result
            torch.cuda.synchronize()
            sec = time.time() - t0
            assert out["niter"] == args.steps and np.isfinite(st["ew"]).all()
            d2h = st["ew"].nbytes + st["eh"].nbytes + 5 * 8 * args.steps
            eng2.close()
            return allmax([sec])[0], d2h, out

        cold_s, _, _ = e2e_pass()   # pays pinned-buffer creation and first-touch of the memory pool
        warm_s, d2h, out = e2e_pass()
        h2d = (h_values.nbytes + h_rowidx.nbytes + h_colptr.nbytes + w0.nbytes + h0_loc.nbytes)
        e2e = {"value": nnz_total * r * args.steps / warm_s, "unit": UNIT,
               "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "seconds": warm_s, "seconds_cold": cold_s, "iterations": args.steps,
               "value_cold": nnz_total * r * args.steps / cold_s,
               "host_threads": max(1, host_cores() // world),
               "lkh_first": [float(v) for v in out["lkh_trace"][:3]],
               "what": "vbnmf_create(host CSC: fp64 values, int32 indices) + set_state + "
                       "vbnmf_run(K, reference loop defaults) + get_state(ew, eh), max over ranks; "
                       "`seconds` = second (warm) of two identical passes, `seconds_cold` = the "
                       "first, which also creates the pinned staging buffers and maps the device "
                       "memory pool"}
        del csc

    # ---- parity ------------------------------------------------------------------------------
    if P and world == 1 and not args.no_parity:
        parity, cpu = oracle_parity(n, m_loc, h_colptr, h_rowidx, h_values, w0, h0_loc, P, gp64, gp32)
    elif P and world > 1 and not args.no_parity:
        # sharded bound against ONE GPU holding the whole matrix (rank 0), same start
        parity = {"iterations": P, "lkh_sharded": [float(v) for v in gp64["lkh"]]}
        if rank == 0 and strong and nnz_total < 3.5e9:
            fc, fr, fv, _ = synth.tenx_like_device(n, m_total, wl["r_true"], wl["density"],
                                                   wl["seed"], dev, 0, m_total)
            e1 = Engine.from_device_csc(n, m_total, int(fr.numel()), fc, fr, fv, device=local_rank)
            one = gpu_parity_run(e1, w0, np.asfortranarray(h0), P, 0)
            e1.close()
            del fc, fr, fv
            torch.cuda.empty_cache()
            parity.update({
                "lkh_single_gpu": [float(v) for v in one["lkh"]],
                "lkh_rel_err_vs_single_gpu": [abs(a - b) / abs(b)
                                              for a, b in zip(gp64["lkh"], one["lkh"])],
                "ew_max_rel_err_vs_single_gpu": relerr(gp64["ew"], one["ew"]),
                "eh_max_rel_err_vs_single_gpu": relerr(gp64["eh"], one["eh"][:, c0:c1]),
                "cid_mismatches_vs_single_gpu": int(np.count_nonzero(gp64["cid"] !=
                                                                     one["cid"][c0:c1])),
                "what": "the %d-rank sharded run against a single-GPU run of the whole matrix on "
                        "rank 0 (same w0, h0); the single-GPU path is checked against the CPU "
                        "oracle in the N = 1 line" % world})
            parity["lkh_rel_err"] = max(parity["lkh_rel_err_vs_single_gpu"])
            parity["cid_mismatches"] = parity["cid_mismatches_vs_single_gpu"]
        barrier()
    if ml is not None and P and not args.no_parity and h_colptr is not None:
        from ccfindr_b200 import synth as _synth
        from oracle import bindings as ob
        restore_omp_threads()
        w0m, h0m = _synth.uniform_init(n, m_loc, wl["ml_rank"], seed=5)
        t0 = time.time()
        ref = ob.sparse_ml_run((n, m_loc, h_colptr, h_rowidx, h_values), w0m, h0m, Itmax=P, Tol=0.0)
        g = ml.pop("_gpu")
        ml["parity"] = {"iterations": P, "oracle": "oracle_sparse.c osp_ml_run on the whole matrix",
                        "oracle_seconds": round(time.time() - t0, 2),
                        "lik_rel_err": [abs(a - b) / abs(b) for a, b in zip(g["lik_trace"], ref["lik_trace"])],
                        "w_max_rel_err": relerr(g["w"], ref["w"]), "h_max_rel_err": relerr(g["h"], ref["h"])}
    elif ml is not None:
        ml.pop("_gpu", None)
    if cpu is None and rank == 0 and not args.no_cpu and h_colptr is not None:
        cpu = cpu_baseline_port(n, r, h_colptr, h_rowidx, h_values, w0, h0_loc)
    barrier()
    if rank != 0:
        return None

    b_alg = algorithmic_bytes(nnz_loc, n, m_loc, r)
    t_sweep = (t64["ms_cols"] + t64["ms_rows"]) * 1e-3
    ach = b_alg / t_sweep / 1e9
    traffic, traffic_note = measured_traffic(name)
    clk = (t64["clocks"] or {}).get("sm_mhz") or 1965.0
    smem_peak = SM_COUNT * SMEM_BYTES_PER_CLK * clk * 1e6 / 1e9          # GB/s at the sampled clock
    smem_ach = 2.0 * nnz_loc * r * 8 / t_sweep / 1e9
    p16 = layout["format"] == "p16"
    kname = "sweep_p16_kernel" if p16 else "sweep_tiled_kernel"
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t64["ms_step"], "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "genes": n, "cells_total": m_total,
                   "nnz_total": nnz_total, "rank": r, "precision": "fp64",
                   "sharding": "cells over %d GPU(s), %s; 1 all-reduce/iter" % (world, shard_how),
                   "shard_cells_rank0": m_loc, "shard_nnz_rank0": nnz_loc,
                   "loop": "device-controlled product loop; %d untimed iterations with fixed "
                           "hyper-parameters (hyper.update.n0), then K timed iterations with "
                           "hyper_update after each; one control readback per 8 iterations"
                           % max(args.warmup, N0),
                   "l2": "inputs larger than L2 (two tiled copies of X, %.2f GB per GPU vs 126 MB)"
                         % (layout["bytes"] / 1e9),
                   "layout": layout, "generator_scale": scale, "gen_seconds": round(t_gen, 2)},
        "clocks": t64["clocks"],
        "e2e": e2e,
        "gpu_launches": int(t64["res"]["launches"]),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic, "traffic_note": traffic_note,
                     "kernel": "%s<COLS=1> + %s<COLS=0> (the two passes of the nonzero sweep, "
                               "incl. their combine kernels)" % (kname, kname),
                     "algorithmic_bytes_per_launch": b_alg,
                     "ms_per_launch": {"sweep_cols": t64["ms_cols"], "sweep_rows": t64["ms_rows"]},
                     "iteration_frac": b_alg / (t64["ms_step"] * 1e-3) / 1e9 / peak,
                     "secondary": {"bound": "smem", "achieved": smem_ach, "peak": smem_peak,
                                   "unit": "GB/s", "frac": smem_ach / smem_peak,
                                   "what": "rank-r fp64 rows gathered from shared memory: 2 passes x "
                                           "nnz x r x 8 B against 148 SMs x 128 B/clk at the "
                                           "sampled SM clock (the pipe that limits this "
                                           "formulation, DESIGN.md section 5)"},
                     "limiter": "shared-memory gather wavefronts (LSU data pipe), not HBM "
                                "(DESIGN.md section 5)",
                     "peak_source": peak_src},
        "cpu_baseline": cpu,
        "parity": parity,
        "fp32_storage_mode": mixed,
        "ml_path": ml,
        "wall_ms_per_step": t64["wall_ms"],
        "lkh_last": t64["res"]["lkh"],
        "hyper_last": t64["res"]["hyper"],
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ccfindr_b200.engine import Comm, Engine, set_host_threads

    rank, local_rank, world = env_rank()
    if args.gpus != world and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    if args.workload.startswith("small"):   # plumbing matrices: a 2,000-gene matrix has empty genes
        os.environ.setdefault("VBNMF_ALLOW_EMPTY", "1")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.tensor(list(Engine.nccl_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        comm = Comm(world, rank, bytes(uid.cpu().tolist()), device=local_rank)
    # the ranks of one box share its cores: staging threads = cores / ranks
    set_host_threads(max(1, min(16, host_cores() // world)))
    ctx = (rank, local_rank, world, dev, comm)
    line = run_config(args, args.workload, ctx)
    sec = None
    if args.secondary != "none" and args.secondary != args.workload:
        sec = run_config(args, args.secondary, ctx)
    if rank == 0:
        if sec is not None:
            line["secondary"] = sec
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        comm.close()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own vbnmf_update (src/vbnmf_update.cpp:16-102, compiled in place into
    oracle/_ref) on the host cores; each step = one call on a dense slab of the workload."""
    rank, _, world = env_rank()
    if rank != 0:
        return
    import torch
    from ccfindr_b200 import synth
    from oracle import bindings as ob
    cores = restore_omp_threads()
    wl = WORKLOADS[args.workload]
    n, r = wl["n"], wl["rank"]
    m_total = wl["m_total"] if "m_total" in wl else wl["m_per_gpu"] * max(world, args.gpus)
    ms = min(args.ref_cells, synth.TENX_CHUNK, m_total)
    # first generator chunk of the same matrix (sampled on the CPU here), first ms cells
    colptr, rowidx, values, _ = synth.tenx_like_device(n, m_total, wl["r_true"], wl["density"],
                                                       wl["seed"], torch.device("cpu"), 0,
                                                       synth.TENX_CHUNK)
    colptr, rowidx, values = colptr.numpy(), rowidx.numpy(), values.numpy().astype(np.float64)
    end = int(colptr[ms])
    import scipy.sparse as sp
    csc = sp.csc_matrix((values[:end], rowidx[:end], colptr[:ms + 1]), shape=(n, ms))
    w0, h0 = init_factors(n, m_total, r, seed=1000 * r + 1)
    h0 = np.asfortranarray(h0[:, :ms])
    have_ref = ob.ref_lib() is not None
    if have_ref:
        X = np.asfortranarray(csc.toarray())
        kind = "reference"
        wh = dict(lw=w0, lh=h0, ew=w0, eh=h0)
        step = lambda wh: ob.ref_vbnmf_update(X, wh, HYPER, np.finfo(np.float64).eps)
        what = ("src/vbnmf_update.cpp compiled in place (oracle/_ref; stand-in Eigen/Rcpp/GSL "
                "headers, GEMM loops OpenMP-parallel, the rest serial as in the reference)")
    else:
        kind = "port"
        wh = dict(lw=w0, lh=h0, ew=w0, eh=h0)
        step = lambda wh: ob.sparse_vb_step(csc, wh, HYPER, np.finfo(np.float64).eps)
        what = "oracle_sparse.c osp_vb_step (oracle/_ref not available)"
    for _ in range(args.warmup):
        wh = step(wh)
    t0 = time.time()
    for _ in range(args.steps):
        wh = step(wh)
    sec = (time.time() - t0) / args.steps
    value = end * r / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": max(world, args.gpus), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong" if "m_total" in wl else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "genes": n, "cells_total": m_total, "rank": r,
                   "precision": "fp64"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "dense slab: first %d cells (%d nnz) of the same matrix per "
                                   "step (the reference densifies X, R/bayesian.R:339; the whole "
                                   "matrix would need %.0f GB per n x m temporary); %s"
                                   % (ms, end, n * m_total * 8 / 1e9, what)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lkh_last": float(wh["lkh"]),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--secondary", default=None,
                    help="second workload nested under 'secondary' (default: c2 beside c3; 'none')")
    ap.add_argument("--parity-iters", type=int, default=3,
                    help="iterations compared with the CPU oracle (0: skip the parity block)")
    ap.add_argument("--ref-cells", type=int, default=400,
                    help="cells in the dense slab one reference step processes")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the CPU-oracle comparison")
    args = ap.parse_args()
    if args.secondary is None:
        args.secondary = "c2" if args.workload == "c3" else "none"
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
