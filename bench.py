#!/usr/bin/env python
"""Benchmark of the VB-NMF hot path (BASELINE.json metric: nnz*rank updates/s per VB iteration).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (rank 0)

A "step" is one VB iteration (posterior update of W and H, nonzero sweep, lower bound, the
per-iteration host readback of the loop) over the whole synthetic matrix.  Workload at N GPUs
(weak scaling): BASELINE config 2 per GPU -- 20,000 genes x 100,000 cells per GPU, ~8 % nonzero
10x-shaped Poisson counts (SURVEY.md 8d generator), rank 10, fp64; cells sharded over the ranks,
one NCCL all-reduce of the W-side statistics per iteration.  One JSON line on rank 0.
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
    # name: genes, cells per GPU, true rank, density, seed, fit rank
    "c2": dict(n=20000, m_per_gpu=100000, r_true=10, density=0.08, seed=2, rank=10,
               label="C2: vb_factorize rank=10, 20k genes x 100k cells per GPU, ~8% nonzero"),
    "c3": dict(n=20000, m_per_gpu=None, m_total=1300000, r_true=20, density=0.08, seed=3, rank=20,
               label="C3: vb_factorize rank=20, 20k genes x 1.3M cells sharded over the GPUs"),
    "small": dict(n=2000, m_per_gpu=8000, r_true=5, density=0.08, seed=4, rank=6,
                  label="small: plumbing check, 2k genes x 8k cells per GPU"),
}
HYPER = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)  # gamma.a = gamma.b = 1 (R/bayesian.R:231)
MIXED = {}  # filled by run_ours: the fp32-storage mode measured beside the fp64 headline
METRIC = "VB-NMF nnz*rank updates/s per iteration"
UNIT = "nnz*rank updates/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(nnz, n, m, r, sp=8):
    """B_alg = nnz*(4+4) + 8*(m+1) + 2*r*m*s_p + 2*n*r*s_p bytes per VB iteration,
    s_p = 8 (fp64) or 4 (fp32-storage mode)."""
    return nnz * 8 + 8 * (m + 1) + 2 * r * m * sp + 2 * n * r * sp


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread
    (nvidia_ml_py, one sample per millisecond -- the timed region of the default run is ~40 ms);
    falls back to an `nvidia-smi -lms` subprocess when NVML cannot be loaded."""
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


def shard_bounds(m_total, nranks, chunk):
    """Contiguous cell ranges in whole generator chunks, as even as possible."""
    nch = -(-m_total // chunk)
    base, extra = divmod(nch, nranks)
    b = [0]
    for r in range(nranks):
        b.append(min(m_total, b[-1] + (base + (1 if r < extra else 0)) * chunk))
    return b


def init_factors(n, m_total, rank, seed):
    from ccfindr_b200 import synth
    return synth.random_init(n, m_total, rank, HYPER, seed)  # vb_init 'random'


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from ccfindr_b200 import synth
    from ccfindr_b200.engine import Comm, Engine

    rank, local_rank, world = env_rank()
    if args.gpus != world and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
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

    wl = WORKLOADS[args.workload]
    n, r = wl["n"], wl["rank"]
    m_total = wl["m_per_gpu"] * world if wl.get("m_per_gpu") else wl["m_total"]
    bounds = shard_bounds(m_total, world, synth.TENX_CHUNK)
    c0, c1 = bounds[rank], bounds[rank + 1]
    t_gen = time.time()
    colptr, rowidx, values, scale = synth.tenx_like_device(n, m_total, wl["r_true"], wl["density"],
                                                           wl["seed"], dev, c0, c1)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    m_loc, nnz_loc = c1 - c0, int(rowidx.numel())
    w0, h0 = init_factors(n, m_total, r, seed=1000 * r + 1)
    h0_loc = np.asfortranarray(h0[:, c0:c1])

    # ---- device-resident arm: inputs already in HBM when the timed region starts --------------
    eng = Engine.from_device_csc(n, m_loc, nnz_loc, colptr, rowidx, values, device=local_rank)
    if comm is not None:
        eng.attach_comm(comm)
    eng.set_state(w0, h0_loc)
    eng.bench_iterations(HYPER, max(args.warmup, 1))          # warm-up (untimed)
    nnz_t = torch.tensor([float(nnz_loc)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz_total = int(nnz_t.item())
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    t0 = time.time()
    res = eng.bench_iterations(HYPER, args.steps)             # EXACTLY K iterations, CUDA events
    res["layout"] = eng.layout_info()
    torch.cuda.synchronize()
    wall_ms = (time.time() - t0) * 1e3
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    tms = torch.tensor([res["ms_total"], res["ms_cols"], res["ms_rows"], wall_ms],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total, ms_cols, ms_rows, wall_ms = [float(v) for v in tms.tolist()]
    ms_step = ms_total / args.steps
    value = nnz_total * r / (ms_step * 1e-3)
    lkh_dev = res["lkh"]

    # ---- same measurement in the fp32-storage / fp64-accumulate mode (reported beside fp64) -----
    eng.set_precision(1)
    eng.set_state(w0, h0_loc)
    eng.bench_iterations(HYPER, max(args.warmup, 1))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    r32 = eng.bench_iterations(HYPER, args.steps)
    torch.cuda.synchronize()
    t32 = torch.tensor([r32["ms_total"], r32["ms_cols"], r32["ms_rows"]], dtype=torch.float64,
                       device=dev)
    if world > 1:
        dist.all_reduce(t32, op=dist.ReduceOp.MAX)
    t32 = [float(v) / args.steps for v in t32.tolist()]
    b32 = algorithmic_bytes(nnz_loc, n, m_loc, r, sp=4)
    MIXED.update({"value": nnz_total * r / (t32[0] * 1e-3), "unit": UNIT, "ms_per_step": t32[0],
                  "ms_per_launch": {"sweep_cols": t32[1], "sweep_rows": t32[2]},
                  "roofline_frac": b32 / ((t32[1] + t32[2]) * 1e-3) / 1e9 / peaks()[0],
                  "algorithmic_bytes_per_launch": b32, "lkh_last": r32["lkh"],
                  "rel_diff_lkh_vs_fp64_same_iteration_count": abs(r32["lkh"] - lkh_dev) / abs(lkh_dev),
                  "what": "panels lw/lh held in fp32, per-nonzero arithmetic fp32, every sum over "
                          "lanes/slabs/ranks and the posterior update in fp64 (tolerance 1e-4)"})
    eng.set_precision(0)

    if args.no_e2e:
        if rank == 0:
            emit_line(args, wl, world, n, r, m_total, m_loc, nnz_loc, nnz_total, scale, t_gen,
                      value, ms_step, ms_cols, ms_rows, wall_ms, clocks, res, None, None, lkh_dev)
        if world > 1:
            dist.barrier()
            comm.close()
            dist.destroy_process_group()
        return
    # ---- end-to-end arm: HOST buffers through the C ABI, copies inside the timed region ---------
    h_colptr = colptr.cpu().numpy()
    h_rowidx = rowidx.cpu().numpy()
    h_values = values.cpu().numpy().astype(np.float64)        # dgCMatrix @x is double
    eng.close()
    del colptr, rowidx, values
    torch.cuda.empty_cache()
    import scipy.sparse as sp
    csc = sp.csc_matrix((h_values, h_rowidx, h_colptr), shape=(n, m_loc))
    def e2e_pass():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        eng2 = Engine(csc, device=local_rank)                  # H2D of X + device layouts
        if comm is not None:
            eng2.attach_comm(comm)
        eng2.set_state(w0, h0_loc)                             # H2D of the initial factors
        # Tol = 0 never satisfies |1 - lkh/lk0| < Tol: exactly K iterations with the reference's
        # own loop (hyper updates from iteration 11 on, R/bayesian.R:342)
        out = eng2.run(HYPER, Itmax=args.steps, Tol=0.0)
        st = eng2.get_state(("ew", "eh"))                      # D2H of the result
        torch.cuda.synchronize()
        return time.time() - t0, eng2, out, st

    # one untimed pass first (pinned staging buffers, memory pool, page cache of the host arrays),
    # except for the workloads whose upload alone takes seconds
    if nnz_loc < 5e8:
        _, eng_w, _, _ = e2e_pass()
        eng_w.close()
    e2e_s, eng2, out, st = e2e_pass()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    assert out["niter"] == args.steps and np.isfinite(st["ew"]).all()
    h2d = (h_values.nbytes + h_rowidx.nbytes + h_colptr.nbytes + w0.nbytes + h0_loc.nbytes)
    d2h = st["ew"].nbytes + st["eh"].nbytes + 5 * 8 * args.steps
    e2e_value = nnz_total * r * args.steps / e2e_s
    eng2.close()

    # ---- CPU baseline on rank 0 (bounded sample of the same workload) -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_port(n, r, h_colptr, h_rowidx, h_values, w0, h0_loc)

    if rank == 0:
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
               "d2h_bytes_per_step": d2h / args.steps, "seconds": e2e_s, "iterations": args.steps,
               "what": "vbnmf_create(host CSC) + set_state + vbnmf_run(K) + get_state(ew, eh); "
                       "second of two identical passes (the first is the warm-up)"}
        emit_line(args, wl, world, n, r, m_total, m_loc, nnz_loc, nnz_total, scale, t_gen, value,
                  ms_step, ms_cols, ms_rows, wall_ms, clocks, res, e2e, cpu, lkh_dev)
    if world > 1:
        dist.barrier()
        if comm is not None:
            comm.close()
        dist.destroy_process_group()


def emit_line(args, wl, world, n, r, m_total, m_loc, nnz_loc, nnz_total, scale, t_gen, value,
              ms_step, ms_cols, ms_rows, wall_ms, clocks, res, e2e, cpu, lkh_dev):
    peak, peak_src = peaks()
    b_alg_local = algorithmic_bytes(nnz_loc, n, m_loc, r)
    t_sweep = (ms_cols + ms_rows) / args.steps * 1e-3
    ach = b_alg_local / t_sweep / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if wl.get("m_per_gpu") else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "genes": n, "cells_total": m_total,
                   "nnz_total": nnz_total, "rank": r, "precision": "fp64",
                   "sharding": "cells over %d GPU(s), 1 all-reduce/iter" % world,
                   "l2": "inputs larger than L2 (two tiled copies of X, %.2f GB per GPU vs 126 MB)"
                         % (res["layout"]["bytes"] / 1e9),
                   "layout": res["layout"],
                   "generator_scale": scale, "gen_seconds": round(t_gen, 2)},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(res["launches"]),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic,
                     "kernel": "%s<COLS=1> + %s<COLS=0> (the two passes of the nonzero sweep, "
                               "incl. their combine kernels)"
                               % (("sweep_p16_kernel",) * 2 if res["layout"]["format"] == "p16"
                                  else ("sweep_tiled_kernel",) * 2),
                     "algorithmic_bytes_per_launch": b_alg_local,
                     "ms_per_launch": {"sweep_cols": ms_cols / args.steps,
                                       "sweep_rows": ms_rows / args.steps},
                     "iteration_frac": b_alg_local / (ms_step * 1e-3) / 1e9 / peak,
                     "limiter": "shared-memory gather wavefronts (LSU data pipe ~90% busy in the "
                                "gene-owner pass), not HBM (DESIGN.md section 5)",
                     "peak_source": peak_src},
        "cpu_baseline": cpu,
        "fp32_storage_mode": MIXED or None,
        "wall_ms_per_step": wall_ms / args.steps,
        "lkh_last": lkh_dev,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_port(n, r, colptr, rowidx, values, w0, h0, max_cols=24000, iters=2):
    """oracle/oracle_sparse.c (OpenMP, all host threads) on the first max_cols cells."""
    from oracle import bindings as ob
    mc = min(max_cols, len(colptr) - 1)
    end = int(colptr[mc])
    arrays = (n, mc, np.ascontiguousarray(colptr[:mc + 1], dtype=np.int64),
              np.ascontiguousarray(rowidx[:end], dtype=np.int32),
              np.ascontiguousarray(values[:end], dtype=np.float64))
    hy = np.array([HYPER[k] for k in ("aw", "bw", "ah", "bh")])
    sec, lkh, threads = ob.sparse_time_iterations(arrays, w0, h0[:, :mc], hy,
                                                  np.finfo(np.float64).eps, iters)
    return {"value": end * r / sec, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d cells (%d nnz) of the same matrix, %d iterations of "
                      "oracle_sparse.c (CSC, OpenMP, fp64)" % (mc, end, iters),
            "seconds_per_iteration": sec}


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
    wl = WORKLOADS[args.workload]
    n, r = wl["n"], wl["rank"]
    m_total = wl["m_per_gpu"] * max(world, args.gpus) if wl.get("m_per_gpu") else wl["m_total"]
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
        kind, cores = "reference", ob.sparse_lib().osp_num_threads()
        wh = dict(lw=w0, lh=h0, ew=w0, eh=h0)
        step = lambda wh: ob.ref_vbnmf_update(X, wh, HYPER, np.finfo(np.float64).eps)
        what = ("src/vbnmf_update.cpp compiled in place (oracle/_ref; stand-in Eigen/Rcpp/GSL "
                "headers, GEMM loops OpenMP-parallel, the rest serial as in the reference)")
    else:
        kind, cores = "port", ob.sparse_lib().osp_num_threads()
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
        "scaling": "weak" if wl.get("m_per_gpu") else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "genes": n, "cells_total": m_total, "rank": r,
                   "precision": "fp64"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "dense slab: first %d cells (%d nnz) of the same matrix per "
                                   "step; %s" % (ms, end, what)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lkh_last": float(wh["lkh"]),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-cells", type=int, default=400,
                    help="cells in the dense slab one reference step processes")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
