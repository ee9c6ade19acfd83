#!/usr/bin/env python
"""bench.py -- genotypes/sec per Arnoldi iteration (k=20) of the B200 path.

A "step" is one perform_op (y = X X' x, the body of one Arnoldi iteration,
svdwide.cpp:21-68) over the whole synthetic 500,000 x 100,000 2-bit genotype
matrix resident in HBM (BASELINE.json config 2/3), SNP-sharded over --gpus
ranks with one shard sum of y per step (the library's peer-memory kernel, else ncclAllReduce).

  value     N*P / t_step with x, y resident in HBM (CUDA events on the library's
            launch stream, max over ranks)
  e2e       same metric through the C-ABI host-pointer call fpb_perform_op:
            pinned host x in, host y out, copies inside the timed region
  roofline  the whole perform_op against the single-read roofline of SURVEY.md section 8d
            (ceil(N/4)*P + 16N + 16P algorithmic bytes per op) and MEASURED_PEAKS.json hbm_gbs;
            roofline.launch = the dominant contraction kernel alone (one launch per half,
            ceil(N/4)*P_local packed bytes / launch time, CUDA events around the launch)
  cpu_baseline  the CPU oracle (oracle/, a port of read_snp_block + perform_op)
            timed on >= 20 reference blocks of the same matrix, all host threads (extrapolated)
  solve     full k=20 solve + the reference's --check mse on the result, at every N
  config_1m_x_500k  BASELINE configs[4], measured when run on 8 GPUs

`--impl reference` times that CPU port alone (the upstream binary cannot be
built in this image: Eigen/Spectra/Boost are absent).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genotypes/sec per Arnoldi iter (k=20)"
UNIT = "genotypes/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=500000, help="individuals")
    ap.add_argument("--p", type=int, default=100000, help="SNPs")
    ap.add_argument("--k", type=int, default=20, help="ndim of the solve")
    ap.add_argument("--no-solve", action="store_true", help="skip the full k=20 solve")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-small", action="store_true", help="skip the 10k x 100k side measurement")
    ap.add_argument("--no-block", action="store_true", help="skip the k = 2 / k = 20 block ops")
    ap.add_argument("--no-cfg5", action="store_true", help="skip 1M x 500k at --gpus 8")
    ap.add_argument("--cfg5", action="store_true", help="force the 1M x 500k side measurement")
    ap.add_argument("--cfg5-shape", default="1000000x500000",
                    help="N x P of the side measurement (tests of the code path at small sizes)")
    ap.add_argument("--cpu-sample-snps", type=int, default=0)
    return ap.parse_args()


def workload_name(a):
    return "synthetic Balding-Nichols bed %d x %d, k=%d, binom2, HBM-resident 2-bit" % (
        a.n, a.p, a.k)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def wait_first(self, timeout=3.0):
        """Block until nvidia-smi has produced its first sample (its start-up is CPU work that must
        not overlap a timed region)."""
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.02)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads():
    return (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity")
            else (os.cpu_count() or 1))


def cpu_sample_snps(a, nblocks=20):
    """SNP columns of the CPU sample: `nblocks` blocks of the reference's own --memory 2048 block
    size for the FULL problem (flashpca.cpp:649-676; 465 SNPs at 500k x 100k -> 9300 columns),
    BASELINE.md section 3: ">= 20 blocks extrapolated to 216 and labelled extrapolated"."""
    from oracle import oracle as O
    import ctypes
    lib = ctypes.CDLL(O.build())
    lib.fo_block_size_from_memory.restype = ctypes.c_uint
    lib.fo_block_size_from_memory.argtypes = [ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_uint,
                                              ctypes.c_int, ctypes.c_uint]
    bs = int(lib.fo_block_size_from_memory(a.n, a.p, a.k, 0, 2048)) or 1
    bs = max(1, min(bs, a.p))
    snps = a.cpu_sample_snps or min(a.p, nblocks * bs)
    return snps, bs


def cpu_port_run(a, payload, n, snps, bs, steps, warmup, threads=None):
    """Time the CPU oracle's perform_op (the reference's block loop, block size `bs`) over `snps`
    SNP columns; returns (genotypes/s, [ms per step], threads)."""
    import numpy as np
    from oracle import oracle as O
    orc = O.COracle(payload, n, snps, threads=threads)
    x = np.random.default_rng(0).standard_normal(n)
    for _ in range(warmup):
        orc.perform_op(x, bs)
    ms = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.perform_op(x, bs)
        ms.append((time.perf_counter() - t0) * 1e3)
    dt = sum(ms) / len(ms) * 1e-3
    return n * snps / dt, ms, orc.threads


def run_reference(a):
    """Reference arm: the CPU port of the path (the upstream binary cannot be built here), all
    host threads, each step = the reference's block loop over a >= 20-block sample of the same
    synthetic matrix; the per-genotype rate is what `value` reports (extrapolated to the full
    matrix: the loop is the same 216 blocks, 20 of them timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from flashpca_b200.synth import SynthSpec
    from oracle import oracle as O
    snps, bs = cpu_sample_snps(a)
    spec = SynthSpec(a.n, a.p)
    payload = O.synth_packed_bed(spec, 0, snps)     # host generator, same hash as the device one
    # all host threads, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    ncpu = host_threads()
    val, ms, threads = cpu_port_run(a, payload, a.n, snps, bs, a.steps, max(a.warmup, 1),
                                    threads=ncpu)
    nblocks_full = (a.p + bs - 1) // bs
    sample = ("EXTRAPOLATED: each step runs the reference block loop over the first %d of %d SNP "
              "columns (%d of %d blocks of block_size %d, --memory 2048 formula) x %d individuals; "
              "bed bytes served from RAM; OpenMP port on %d threads (upstream itself is "
              "single-threaded)" % (snps, a.p, (snps + bs - 1) // bs, nblocks_full, bs, a.n, threads))
    ms_mean = sum(ms) / len(ms)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": max(a.warmup, 1), "ms_per_step": ms_mean,
        "median_ms_per_step": statistics.median(ms),
        "ms_per_full_op_extrapolated": ms_mean * a.p / snps,
        "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "n": a.n, "p": a.p, "k": a.k,
                   "note": "CPU port of Data::read_snp_block + SVDWideOnline::perform_op "
                           "(oracle/flashpca_oracle.c); upstream binary unbuildable here "
                           "(Eigen/Spectra/Boost absent)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _stats(ms):
    ms = [float(v) for v in ms]
    return {"mean": sum(ms) / len(ms), "median": statistics.median(ms), "min": min(ms),
            "max": max(ms), "n": len(ms)}


def measure_op(lib, _lib, op, n, steps, warmup, world, dist, torch):
    """Device-resident metric of one operator: `warmup` untimed ops, then exactly `steps` ops with a
    CUDA event after each (library stream), bracketed by barrier + synchronize; max over ranks."""
    import ctypes
    x = torch.randn(n, dtype=torch.float64, device="cuda",
                    generator=torch.Generator(device="cuda").manual_seed(1234))
    y = torch.empty_like(x)
    ms = ctypes.c_float()
    each = (ctypes.c_float * steps)()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.check(lib.fpb_sync(op.h), op.h)

    barrier()
    _lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), max(warmup, 3),
                                       ctypes.byref(ms), None), op.h)
    barrier()
    l0 = lib.fpb_launch_count(op.h)
    _lib.check(lib.fpb_time_perform_op_steps(op.h, x.data_ptr(), y.data_ptr(), steps, each), op.h)
    launches = lib.fpb_launch_count(op.h) - l0
    barrier()
    st = _stats(list(each))
    t = torch.tensor([st["mean"], st["median"], st["max"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st["mean"], st["median"], st["max"] = t[0].item(), t[1].item(), t[2].item()
    return st, int(launches), x, y, barrier


def solve_block(op, k, p, n, rank, world, barrier):
    """Full k-dimensional solve to convergence (tol 1e-6, maxiter 500, ncv = 2k+1 as
    randompca.cpp:174-178), twice (the first call allocates the Lanczos workspace), then the
    reference's --check criterion on the solver's own eigenpairs, computed on the device:
    mse = sum_j ||XX'u_j/p - u_j d_j||^2 / (N k)  (randompca.cpp:663-703; README.md:207 "< 1e-8")."""
    import numpy as np
    runs = []
    res = None
    # the eigenvectors land in a caller-owned buffer that exists before the timed region (a fresh
    # 80 MB array costs ~5 ms of page faults per call, which is not the library's time)
    ubuf = np.empty((n, k), order="F") if rank == 0 else None
    if ubuf is not None:
        ubuf.fill(0.0)
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        res = op.pca(k, 2 * k + 1, 500, 1e-6, want_vectors=(rank == 0), out_vectors=ubuf)
        barrier()
        runs.append((time.perf_counter() - t0, op.pca_phase_seconds()))
    op_ms = op.op_times_ms()
    err = op.pca_residual(k, float(p))
    sec, ph = runs[1]
    # extension: block Krylov on the tcgen05 block operator (8 columns per pass), same tolerance
    blk = None
    try:
        bruns = []
        for _ in range(2):
            barrier()
            t0 = time.perf_counter()
            bres = op.pca_block(k, 1e-6, want_vectors=(rank == 0), out_vectors=ubuf)
            barrier()
            bruns.append(time.perf_counter() - t0)
        berr = op.pca_residual(k, float(p))
        blk = {"seconds": bruns[1], "first_call_seconds": bruns[0], "passes_of_8_columns": int(bres["npasses"]),
               "nconv": int(bres["nconv"]), "check_mse": float(berr.sum() / (n * k)),
               "max_rel_eigenvalue_difference_to_spectra_schedule":
                   float(np.abs(bres["values"] / res["values"] - 1).max()),
               "note": "fpb_pca_block: block Lanczos + Rayleigh-Ritz, not upstream's algorithm; "
                       "Spectra's convergence criterion, tol 1e-6"}
    except Exception as e:  # noqa: BLE001 -- the extension must never take the contract line down
        blk = {"error": str(e)[:200]}
    return {"seconds": sec, "block_krylov": blk, "first_call_seconds": runs[0][0], "iterate_seconds": ph["iterate"],
            "eigenvector_assemble_seconds": ph["assemble"],
            "eigenvector_download_seconds": ph["download"],
            "eigenvectors_downloaded_on": ("rank 0 only" if world > 1 else "the single rank")
                                          + ", into a caller-owned buffer allocated before the timed region",
            "nops": int(res["nops"]), "nconv": int(res["nconv"]), "restarts": int(res["niter"]) - 1,
            "median_op_ms": float(np.median(op_ms)) if op_ms.size else None,
            "check_mse": float(err.sum() / (n * k)), "check_max_err": float(err.max()),
            "eigenvalues_over_p": [float(v / p) for v in res["values"]]}


def run_b200(a):
    import ctypes

    import numpy as np
    import torch
    import torch.distributed as dist

    from flashpca_b200 import _lib
    from flashpca_b200 import dist as fdist
    from flashpca_b200.synth import SynthSpec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    # clocks are sampled on EVERY rank from before the warm-up, so that starting nvidia-smi never
    # lands inside a timed region (round 1: forking it on rank 0 right before 20 timed steps
    # stalled that rank's launches and doubled the N=2 mean)
    sampler = ClockSampler(local)
    sampler.start()
    # NCCL announces its version on stdout when the first communicator comes up; stdout carries
    # exactly one JSON line, so fd 1 points at stderr until the communicators exist
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    lib = _lib.load()

    n, p, k = a.n, a.p, a.k
    spec = SynthSpec(n, p)
    j0, j1 = fdist.shard_range(p, world, rank)
    t_stage = time.perf_counter()
    op = spec.create_operator(device=local, j0=j0, j1=j1)
    t_stage = time.perf_counter() - t_stage
    if world > 1:
        fdist.attach_nccl(op, world, rank)
        y0 = torch.zeros(n, dtype=torch.float64, device="cuda")
        y1 = torch.empty_like(y0)
        _lib.check(lib.fpb_perform_op_dev(op.h, y0.data_ptr(), y1.data_ptr()), op.h)
        _lib.check(lib.fpb_sync(op.h), op.h)      # first collective of the library's communicator
        del y0, y1
    shard_sum = fdist.comm_kind(op)
    slice_upload = shard_sum == "peer" and os.environ.get("FPB_SLICE_UPLOAD", "1") != "0"
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    sampler.wait_first()

    # ---- device-resident metric: exactly --steps ops, a CUDA event after each
    st, launches, x_dev, y_dev, barrier = measure_op(lib, _lib, op, n, a.steps, a.warmup, world,
                                                     dist, torch)
    ms_step = st["mean"]
    ms = ctypes.c_float()
    kms = (ctypes.c_float * 4)()
    _lib.check(lib.fpb_time_perform_op(op.h, x_dev.data_ptr(), y_dev.data_ptr(), 1,
                                       ctypes.byref(ms), kms), op.h)
    t = torch.tensor([kms[0], kms[1], kms[2], kms[3]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    k_ms = [t[0].item(), t[1].item()]        # the two halves, small kernels included
    g_ms = [t[2].item(), t[3].item()]        # the contraction kernel of each half alone

    # ---- end to end through the host-pointer C ABI (H2D + op + shard sum + D2H per step)
    gen = torch.Generator(device="cpu").manual_seed(1234)
    x_host = torch.randn(n, dtype=torch.float64, generator=gen).pin_memory()
    y_host = torch.empty(n, dtype=torch.float64).pin_memory()
    for _ in range(max(a.warmup, 3)):
        _lib.check(lib.fpb_perform_op(op.h, x_host.data_ptr(), y_host.data_ptr()), op.h)
    barrier()
    e2e_each = []
    t0 = time.perf_counter()
    for _ in range(a.steps):
        t1 = time.perf_counter()
        _lib.check(lib.fpb_perform_op(op.h, x_host.data_ptr(), y_host.data_ptr()), op.h)
        e2e_each.append((time.perf_counter() - t1) * 1e3)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    t = torch.tensor([e2e_ms, statistics.median(e2e_each)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, e2e_median = t[0].item(), t[1].item()
    clocks = sampler.stop()                  # sampled over both timed regions
    y_check = float(torch.linalg.norm(y_host))

    # ---- full k=20 solve to convergence + the reference's --check criterion
    solve = None if a.no_solve else solve_block(op, k, p, n, rank, world, barrier)

    # ---- BASELINE configs[4]: 1,000,000 x 500,000, k=20, SNP-sharded over the 8 GPUs of the box
    cfg5 = None
    if (world >= 8 and (n, p) == (500000, 100000) and not a.no_cfg5) or a.cfg5:
        n5, p5 = (int(v) for v in a.cfg5_shape.split("x"))
        spec5 = SynthSpec(n5, p5)
        a5, b5 = fdist.shard_range(p5, world, rank)
        t5 = time.perf_counter()
        op5 = spec5.create_operator(device=local, j0=a5, j1=b5)
        t5 = time.perf_counter() - t5
        if world > 1:
            fdist.attach_nccl(op5, world, rank)
        st5, l5, _, _, barrier5 = measure_op(lib, _lib, op5, n5, 20, 3, world, dist, torch)
        solve5 = solve_block(op5, k, p5, n5, rank, world, barrier5)
        npb5 = (n5 + 3) // 4
        alg5 = npb5 * p5 + 16 * n5 + 16 * p5
        peak5, _ = peaks()
        cfg5 = {"workload": "synthetic Balding-Nichols bed %d x %d, k=%d, SNP-sharded x%d, "
                            "HBM-resident (%.1f GB of packed genotypes per GPU)"
                            % (n5, p5, k, world, npb5 * (b5 - a5) / 1e9),
                "value": n5 * p5 / (st5["mean"] * 1e-3), "unit": UNIT, "ms_per_step": st5["mean"],
                "step_ms": st5, "gpu_launches": l5, "stage_seconds": t5,
                "perform_op_frac_of_single_read_roofline":
                    alg5 / world / (st5["mean"] * 1e-3) / 1e9 / peak5,
                "solve": solve5}
        op5.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    npb = (n + 3) // 4
    alg_bytes = npb * p + 16 * n + 16 * p          # one perform_op, whole job
    peak, peak_src = peaks()
    kern_bytes = npb * (j1 - j0)                   # one contraction launch = one half, per GPU
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    fused = bool(lib.fpb_path_info(op.h) & _lib.PATH_FUSED)
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("n") == n and tj.get("p") == p and world == 1:
                traffic = tj.get("k_fused_op_dram_bytes_per_launch" if fused
                                 else "perform_op_dram_bytes")
        except (OSError, ValueError):
            pass
    dom = max(g_ms) if max(g_ms) > 0 else max(k_ms)
    op_gbs = alg_bytes / world / (ms_step * 1e-3) / 1e9
    # Top level = the whole perform_op against the SINGLE-read roofline of SURVEY 8d
    # (ceil(N/4) P + 16 N + 16 P algorithmic bytes per op, per GPU when sharded); the per-launch
    # figures of the dominant kernel sit under "launch".
    roofline = {
        "bound": "hbm", "achieved": op_gbs, "peak": peak, "unit": "GB/s", "frac": op_gbs / peak,
        "traffic": traffic, "peak_source": peak_src,
        "what": "one perform_op (y = X X'x) per GPU: algorithmic bytes of SURVEY 8d / mean step time",
        "algorithmic_bytes_per_op": alg_bytes / world, "ms": ms_step,
        "frac_at_median_step": alg_bytes / world / (st["median"] * 1e-3) / 1e9 / peak,
        "halves_ms": {"Xtx": k_ms[0], "Xt": k_ms[1]},
        "launch": {
            "kernel": ("k_fused_op (one launch per op, both halves)" if fused else
                       "k_imma_gemv_tma / k_imma_gemv_tma_t (one launch per half; slower one)"),
            "algorithmic_bytes_per_launch": kern_bytes,
            "achieved": kern_bytes / (dom * 1e-3) / 1e9,
            "frac": kern_bytes / (dom * 1e-3) / 1e9 / peak,
            "launch_ms": ({"fused_op": g_ms[0]} if fused else
                          {"Xtx_contraction": g_ms[0], "Xt_contraction": g_ms[1]})},
        "note": ("fused single-pass kernel" if fused else
                 "two-kernel path: each half streams the single packed copy once (2 HBM reads per "
                 "op), so the op-level fraction is capped near 0.55; see DESIGN.md section 4.7 for "
                 "why the single-read kernel is operand-ingest bound on B200"),
    }

    # BASELINE.json configs[1] (synthetic 10,000 x 100,000, k=20, one GPU) next to the headline
    small = None
    if world == 1 and (n, p) == (500000, 100000) and not a.no_small:
        sn, sp_ = 10000, 100000
        sop = SynthSpec(sn, sp_).create_operator(device=local)
        sst, sl, _, _, sbar = measure_op(lib, _lib, sop, sn, 200, 10, 1, dist, torch)
        s_alg = ((sn + 3) // 4) * sp_ + 16 * sn + 16 * sp_
        ssolve = solve_block(sop, k, sp_, sn, 0, 1, sbar)
        small = {"workload": "synthetic Balding-Nichols bed %d x %d, k=%d (BASELINE configs[1])"
                             % (sn, sp_, k),
                 "value": sn * sp_ / (sst["mean"] * 1e-3), "unit": UNIT,
                 "ms_per_step": sst["mean"], "step_ms": sst, "gpu_launches": sl,
                 "perform_op_frac_of_single_read_roofline":
                     s_alg / (sst["mean"] * 1e-3) / 1e9 / peak,
                 "solve": ssolve}
        sop.close()

    # the block variants (perform_op_mat, svdwide.cpp:71-118) at k = 20 (the loadings / --check /
    # --project shape of the headline solve): tcgen05 kernels, 8 columns per pass
    block = None
    if world == 1 and not a.no_block:
        lstream = torch.cuda.ExternalStream(lib.fpb_stream(op.h))
        block = {}
        for kk in (2, 20):
            xb = torch.randn(kk * n, dtype=torch.float64, device="cuda")
            yb = torch.empty_like(xb)
            _lib.check(lib.fpb_perform_op_multi_dev(op.h, xb.data_ptr(), kk, yb.data_ptr()), op.h)
            _lib.check(lib.fpb_sync(op.h), op.h)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if kk == 2 else 3
            with torch.cuda.stream(lstream):
                e0.record()
                for _ in range(reps):
                    _lib.check(lib.fpb_perform_op_multi_dev(op.h, xb.data_ptr(), kk, yb.data_ptr()),
                               op.h)
                e1.record()
            _lib.check(lib.fpb_sync(op.h), op.h)
            b_ms = e0.elapsed_time(e1) / reps
            block["k%d" % kk] = {
                "columns": kk, "ms_per_call": b_ms, "ms_per_column": b_ms / kk,
                "value_per_column": kk * n * p / (b_ms * 1e-3), "unit": UNIT,
                "kernels": ("k_imma_gemv_tma_2v / _t_2v (mma.sync, two columns per pass)" if kk == 2
                            else "k_umma_xt / k_umma_xv (tcgen05.mma kind::i8, 8 + 8 + 4 columns "
                                 "per pass)")}
            del xb, yb

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        from oracle import oracle as O
        snps, bs = cpu_sample_snps(a)
        payload = O.synth_packed_bed(spec, 0, snps)
        ncpu = host_threads()
        val, cms, threads = cpu_port_run(a, payload, n, snps, bs, steps=3, warmup=1, threads=ncpu)
        # BASELINE.md variant A: the reference is single-threaded in practice (2 blocks, 1 thread)
        s1 = min(snps, 2 * bs)
        val1, _, _ = cpu_port_run(a, payload[: s1 * npb], n, s1, bs, steps=1, warmup=1, threads=1)
        cpu = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
               "single_thread_value": val1,
               "ms_per_full_op_extrapolated": sum(cms) / len(cms) * p / snps,
               "sample": "EXTRAPOLATED: reference block loop over the first %d of %d SNP columns "
                         "(%d blocks of %d, --memory 2048 formula) x %d individuals, 3 steps after 1 "
                         "warm-up, %.0f ms per sample step; single-thread figure from 2 blocks"
                         % (snps, p, (snps + bs - 1) // bs, bs, n, sum(cms) / len(cms))}

    line = {
        "metric": METRIC, "value": n * p / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step,
        "median_ms_per_step": st["median"], "step_ms": st,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "n": n, "p": p, "k": k,
                   "path": "fused single-pass" if fused else "two-kernel",
                   "sharding": ("snp-columns x%d, shard sum of the N-vector per step: %s" % (world, {
                       "peer": "one kernel over NVLink peer memory, fused with the op's finalize "
                               "step (csrc/fpb_peer.cuh)",
                       "nccl": "ncclAllReduce(f64, N)"}[shard_sum]))
                   if world > 1 else "single GPU",
                   "l2": "inputs %.1f GB per GPU >> 126 MB L2; no flush between steps"
                         % (npb * (j1 - j0) / 1e9),
                   "timing": "CUDA event after every step on the library stream; value from the "
                             "mean of exactly --steps steps, max over ranks; median reported beside",
                   "stage_seconds": t_stage},
        "clocks": clocks,
        "e2e": {"value": n * p / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "median_ms_per_step": e2e_median,
                # per rank: with the peer-memory path a rank uploads its 1/world slice of x (the
                # rest arrives over NVLink) and downloads all of y
                "h2d_bytes_per_step": 8 * (-(-n // world)) if slice_upload else 8 * n,
                "d2h_bytes_per_step": 8 * n,
                "bytes_are": "per rank",
                "y_norm": y_check},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "solve": solve,
        "config_10k_x_100k": small,
        "config_1m_x_500k": cfg5,
        "block_variant": block,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
