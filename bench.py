#!/usr/bin/env python
"""bench.py -- genotypes/sec per Arnoldi iteration (k=20) of the B200 path.

A "step" is one perform_op (y = X X' x, the body of one Arnoldi iteration,
svdwide.cpp:21-68) over the whole synthetic 500,000 x 100,000 2-bit genotype
matrix resident in HBM (BASELINE.json config 2/3), SNP-sharded over --gpus
ranks with one NCCL all-reduce of y per step.

  value     N*P / t_step with x, y resident in HBM (CUDA events on the library's
            launch stream, max over ranks)
  e2e       same metric through the C-ABI host-pointer call fpb_perform_op:
            pinned host x in, host y out, copies inside the timed region
  roofline  dominant kernel (k_imma_gemv_tma, one launch per half): ceil(N/4)*P_local
            packed bytes / launch time (CUDA events around the launch) against
            MEASURED_PEAKS.json hbm_gbs; roofline.perform_op = the whole op against
            the single-read roofline ceil(N/4)*P + 16N + 16P of SURVEY.md section 8d
  cpu_baseline  the CPU oracle (oracle/, a port of read_snp_block + perform_op)
            timed on a bounded SNP sample of the same matrix, all host threads

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


def cpu_port_run(a, payload, n, snps, steps, warmup, threads=None):
    """Time the CPU oracle's perform_op over `snps` SNP columns; returns
    (genotypes/s, ms per step, threads, block_size)."""
    import numpy as np
    from oracle import oracle as O
    orc = O.COracle(payload, n, snps, threads=threads)
    # the reference's own --memory 2048 block size for the FULL problem
    # (flashpca.cpp:649-676), capped to the sample
    bs = orc.block_size_from_memory(a.k, False, 2048)
    full_bs = int(orc.lib.fo_block_size_from_memory(a.n, a.p, a.k, 0, 2048))
    bs = max(1, min(full_bs if full_bs else bs, snps))
    x = np.random.default_rng(0).standard_normal(n)
    for _ in range(warmup):
        orc.perform_op(x, bs)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.perform_op(x, bs)
    dt = (time.perf_counter() - t0) / steps
    return n * snps / dt, dt * 1e3, orc.threads, bs


def run_reference(a):
    """Reference arm: the CPU port of the path, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from flashpca_b200.synth import SynthSpec
    snps = a.cpu_sample_snps or max(64, min(a.p, int(2.5e8 // a.n)))
    spec = SynthSpec(a.n, a.p)
    payload = spec.packed_bed(0, snps)
    # all host threads, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    val, ms, threads, bs = cpu_port_run(a, payload, a.n, snps, a.steps, a.warmup, threads=ncpu)
    sample = ("first %d of %d SNP columns x %d individuals per step (block_size %d from the "
              "--memory 2048 formula), bed bytes served from RAM" % (snps, a.p, a.n, bs))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "n": a.n, "p": a.p, "k": a.k,
                   "note": "CPU port of Data::read_snp_block + SVDWideOnline::perform_op "
                           "(oracle/flashpca_oracle.c); upstream binary unbuildable here"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(a):
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    n, p, k = a.n, a.p, a.k
    spec = SynthSpec(n, p)
    j0, j1 = fdist.shard_range(p, world, rank)
    t_stage = time.perf_counter()
    op = spec.create_operator(device=local, j0=j0, j1=j1)
    t_stage = time.perf_counter() - t_stage
    if world > 1:
        fdist.attach_nccl(op, world, rank)

    gen = torch.Generator(device="cpu").manual_seed(1234)
    x_host = torch.randn(n, dtype=torch.float64, generator=gen).pin_memory()
    y_host = torch.empty(n, dtype=torch.float64).pin_memory()
    x_dev = x_host.cuda()
    y_dev = torch.empty_like(x_dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.check(lib.fpb_sync(op.h), op.h)

    import ctypes
    ms = ctypes.c_float()
    kms = (ctypes.c_float * 4)()

    def timed(reps, want_kernels=False):
        _lib.check(lib.fpb_time_perform_op(op.h, x_dev.data_ptr(), y_dev.data_ptr(), reps,
                                           ctypes.byref(ms), kms if want_kernels else None), op.h)
        return ms.value

    # ---- device-resident metric
    barrier()
    timed(max(a.warmup, 3))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.fpb_launch_count(op.h)
    ms_step = timed(a.steps)
    launches = lib.fpb_launch_count(op.h) - l0
    barrier()
    timed(1, want_kernels=True)
    t = torch.tensor([ms_step, kms[0], kms[1], kms[2], kms[3]], dtype=torch.float64,
                     device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t[0].item()
    k_ms = [t[1].item(), t[2].item()]        # the two halves, small kernels included
    g_ms = [t[3].item(), t[4].item()]        # the contraction kernel of each half alone

    # ---- end to end through the host-pointer C ABI (H2D + op + all-reduce + D2H)
    for _ in range(max(a.warmup, 3)):
        _lib.check(lib.fpb_perform_op(op.h, x_host.data_ptr(), y_host.data_ptr()), op.h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        _lib.check(lib.fpb_perform_op(op.h, x_host.data_ptr(), y_host.data_ptr()), op.h)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()
    y_check = float(torch.linalg.norm(y_host))

    # ---- full k=20 solve to convergence (time-to-solution, op count)
    solve = None
    if not a.no_solve:
        runs = []
        for _ in range(2):   # first call allocates the Lanczos workspace; the second is steady state
            barrier()
            t0 = time.perf_counter()
            res = op.pca(k, 2 * k + 1, 500, 1e-6)
            barrier()
            runs.append((time.perf_counter() - t0, op.pca_phase_seconds()))
        op_ms = op.op_times_ms()
        sec, ph = runs[1]
        solve = {"seconds": sec, "first_call_seconds": runs[0][0], "iterate_seconds": ph["iterate"],
                 "eigenvector_assemble_seconds": ph["assemble"],
                 "eigenvector_download_seconds": ph["download"], "nops": int(res["nops"]),
                 "nconv": int(res["nconv"]), "restarts": int(res["niter"]) - 1,
                 "median_op_ms": float(np.median(op_ms)) if op_ms.size else None,
                 "eigenvalue_1_over_p": float(res["values"][0] / p),
                 "eigenvalue_k_over_p": float(res["values"][k - 1] / p)}

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
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("n") == n and tj.get("p") == p and world == 1:
                traffic = tj.get("k_imma_gemv_tma_dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
    fused = bool(lib.fpb_path_info(op.h) & _lib.PATH_FUSED)
    dom = max(g_ms) if max(g_ms) > 0 else max(k_ms)
    op_gbs = alg_bytes / world / (ms_step * 1e-3) / 1e9
    if fused:
        # ONE launch of k_fused_op does both halves and reads the packed matrix once:
        # its algorithmic bytes are the whole local matrix
        traffic = None
        if os.path.exists(tpath):
            try:
                with open(tpath) as f:
                    tj = json.load(f)
                if tj.get("n") == n and tj.get("p") == p and world == 1:
                    traffic = tj.get("k_fused_op_dram_bytes_per_launch")
            except (OSError, ValueError):
                pass
        roofline = {
            "bound": "hbm",
            "kernel": "k_fused_op (one launch per perform_op: both halves of y = X X'x, the "
                      "packed matrix is read from HBM once and re-read from L2)",
            "achieved": kern_bytes / (g_ms[0] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": kern_bytes / (g_ms[0] * 1e-3) / 1e9 / peak, "traffic": traffic,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": kern_bytes,
            "launch_ms": {"fused_op": g_ms[0]},
            "perform_op": {"algorithmic_bytes": alg_bytes / world, "ms": ms_step,
                           "achieved": op_gbs, "frac": op_gbs / peak,
                           "note": "1 fused launch + 2 missing-genotype gathers + 5 small "
                                   "kernels per op, against the single-read roofline of "
                                   "SURVEY 8d"},
        }
    else:
      roofline = {
        # the dominant kernel, as the bench contract defines it: algorithmic bytes of
        # the units ONE launch processes (one half = ceil(N/4) * P_local packed bytes)
        "bound": "hbm",
        "kernel": "k_imma_gemv_tma / k_imma_gemv_tma_t (one launch per half of perform_op; "
                  "the slower of the two is reported)",
        "achieved": kern_bytes / (dom * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": kern_bytes / (dom * 1e-3) / 1e9 / peak, "traffic": traffic,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": kern_bytes,
        "launch_ms": {"Xtx_contraction": g_ms[0], "Xt_contraction": g_ms[1]},
        # the whole perform_op against the SINGLE-read roofline of SURVEY 8d: the
        # two-kernel path reads the packed matrix once per half, i.e. twice per op
        "perform_op": {"algorithmic_bytes": alg_bytes / world, "ms": ms_step,
                       "achieved": op_gbs, "frac": op_gbs / peak,
                       "halves_ms": {"Xtx": k_ms[0], "Xt": k_ms[1]},
                       "note": "2 contraction launches + 7 small kernels per op; each half "
                               "streams the single packed copy once, so the single-read "
                               "roofline fraction is capped near 0.55"},
      }

    # BASELINE.json configs[1] (synthetic 10,000 x 100,000, k=20, one GPU) next to the headline
    # configuration: same metric, device-resident vectors, CUDA events
    small = None
    if world == 1 and (n, p) == (500000, 100000) and not a.no_small:
        sn, sp_ = 10000, 100000
        sspec = SynthSpec(sn, sp_)
        sop = sspec.create_operator(device=local)
        sx = torch.randn(sn, dtype=torch.float64, device="cuda")
        sy = torch.empty_like(sx)
        _lib.check(lib.fpb_time_perform_op(sop.h, sx.data_ptr(), sy.data_ptr(), 10,
                                           ctypes.byref(ms), None), sop.h)
        _lib.check(lib.fpb_time_perform_op(sop.h, sx.data_ptr(), sy.data_ptr(), 200,
                                           ctypes.byref(ms), kms), sop.h)
        s_ms = ms.value
        s_alg = ((sn + 3) // 4) * sp_ + 16 * sn + 16 * sp_
        t0 = time.perf_counter()
        sres = sop.pca(k, 2 * k + 1, 500, 1e-6)
        s_solve = time.perf_counter() - t0
        small = {"workload": "synthetic Balding-Nichols bed %d x %d, k=%d (BASELINE configs[1])"
                             % (sn, sp_, k),
                 "value": sn * sp_ / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms,
                 "contraction_launch_ms": [kms[2], kms[3]],
                 "perform_op_frac_of_single_read_roofline": s_alg / (s_ms * 1e-3) / 1e9 / peak,
                 "solve_seconds_first_call": s_solve, "solve_nops": int(sres["nops"]),
                 "note": "250 MB matrix: two HBM passes + 9 launches per op; launch-bound"}
        sop.close()

    # the block variant (perform_op_mat, svdwide.cpp:71-118): two columns per pass over the matrix
    block = None
    if world == 1:
        xb = torch.randn(2 * n, dtype=torch.float64, device="cuda")
        yb = torch.empty_like(xb)
        lstream = torch.cuda.ExternalStream(lib.fpb_stream(op.h))
        _lib.check(lib.fpb_perform_op_multi_dev(op.h, xb.data_ptr(), 2, yb.data_ptr()), op.h)
        _lib.check(lib.fpb_sync(op.h), op.h)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(lstream):
            e0.record()
            for _ in range(5):
                _lib.check(lib.fpb_perform_op_multi_dev(op.h, xb.data_ptr(), 2, yb.data_ptr()), op.h)
            e1.record()
        _lib.check(lib.fpb_sync(op.h), op.h)
        b_ms = e0.elapsed_time(e1) / 5
        block = {"columns": 2, "ms_per_call": b_ms, "value_per_column": 2 * n * p / (b_ms * 1e-3),
                 "unit": UNIT, "note": "fpb_perform_op_multi_dev, k = 2: both columns from one pass "
                                       "over the packed matrix per half (k_imma_gemv_tma*_2v)"}
        del xb, yb

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        snps = a.cpu_sample_snps or max(64, min(p, int(1e9 // n)))
        sub = spec.create_operator(device=local, j0=0, j1=snps)
        payload = sub.bed_payload()
        sub.close()
        ncpu = (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity")
                else (os.cpu_count() or 1))
        val, cms, threads, bs = cpu_port_run(a, payload, n, snps, steps=3, warmup=1, threads=ncpu)
        # BASELINE.md variant A: the reference is single-threaded in practice
        s1 = max(64, snps // 8)
        val1, _, _, _ = cpu_port_run(a, payload[: s1 * ((n + 3) // 4)], n, s1, steps=1, warmup=1,
                                     threads=1)
        cpu = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
               "single_thread_value": val1,
               "sample": "first %d of %d SNP columns x %d individuals, 3 steps after 1 warm-up, "
                         "block_size %d (--memory 2048 formula), %.0f ms per sample step"
                         % (snps, p, n, bs, cms)}

    line = {
        "metric": METRIC, "value": n * p / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "n": n, "p": p, "k": k,
                   "path": "fused single-pass" if fused else "two-kernel",
                   "sharding": "snp-columns x%d, 1 ncclAllReduce(f64, N) per step" % world
                   if world > 1 else "single GPU",
                   "l2": "inputs %.1f GB per GPU >> 126 MB L2; no flush between steps"
                         % (npb * (j1 - j0) / 1e9),
                   "stage_seconds": t_stage},
        "clocks": clocks,
        "e2e": {"value": n * p / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "y_norm": y_check},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "solve": solve,
        "config_10k_x_100k": small,
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
