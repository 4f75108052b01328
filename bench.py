#!/usr/bin/env python
"""Headline benchmark: cells/s of ``DensityEstimator.fit_predict`` (the sparse-GP density hot path)
on synthetic cells, plus the HBM roofline of the fused N x M covariance build (K1).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

Under ``python -m torch.distributed.run --nproc-per-node N`` one rank drives one GPU; the cell
axis is sharded in contiguous row blocks (strong scaling: the total number of cells is fixed) and
the library all-reduces the Gram matrix and the (loss, gradient) vector over NCCL.

A "step" is one complete pass of the hot path over the synthetic cell matrix:
``Lp = chol(K_MM + jitter I)`` -> ``L = K_NM Lp^-T`` -> Ridge start (Gram + all-reduce) ->
L-BFGS-B to SciPy's default stop (every evaluation one fused pass over L + all-reduce) ->
``log_density = L z + mu``.  Nearest-neighbour distances and landmarks are inputs (computed once,
before the timed region, identically for every arm).  Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

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

METRIC = "cells/sec fit_predict"
UNIT = "cells/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=1_000_000, help="N: total cells (all GPUs together)")
    ap.add_argument("--landmarks", type=int, default=5000, help="M")
    ap.add_argument("--dims", type=int, default=50, help="D")
    ap.add_argument("--cov", default="Matern52", choices=["Matern52", "Matern32", "ExpQuad"])
    ap.add_argument("--cpu-sample", type=int, default=20_000, help="cells of the workload the CPU baseline times")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region (A/B only)")
    ap.add_argument("--predict-queries", type=int, default=1_000_000,
                    help="out-of-sample queries for the extra predict() measurement (BASELINE config 5 shape); 0 = skip")
    return ap.parse_args()


# ---- workload ------------------------------------------------------------------------------------
def make_cells(n, d, seed=0):
    """Synthetic cell matrix of BASELINE.json's shape: U[0, 1)^D, float64, C-contiguous."""
    return np.random.default_rng(seed).random((n, d))


def pick_landmarks(x, m, seed=1):
    """Seeded row sample (identical on every rank and for both arms)."""
    idx = np.sort(np.random.default_rng(seed).choice(x.shape[0], size=m, replace=False))
    return np.ascontiguousarray(x[idx])


def workload_config(args, world):
    return {
        "workload": f"DensityEstimator.fit_predict N={args.cells} cells D={args.dims} M={args.landmarks} landmarks "
                    f"{args.cov} sparse_cholesky L-BFGS-B",
        "n_cells": args.cells,
        "n_landmarks": args.landmarks,
        "dims": args.dims,
        "cov": args.cov,
        "gp_type": "sparse_cholesky",
        "optimizer": "L-BFGS-B (SciPy defaults, maxiter=500)",
        "inputs": "nn_distances (exact 1-NN) and landmarks (seeded row sample) precomputed outside the timed region",
        "parallelism": f"cells row-sharded over {world} GPU(s); NCCL all-reduce of Gram and (loss, grad)",
        "l2_policy": "inputs larger than L2 (L is N x M float64 = %.1f GB per job)" % (args.cells * args.landmarks * 8 / 1e9),
    }


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.device)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---- CPU arms (oracle port of the reference path) ----------------------------------------------------
def oracle_cov(name):
    from oracle import mellon_oracle as O

    return getattr(O, name)


def cpu_fit(x, lm, nn, cov_name):
    """The identical region on the host: Lp -> L -> z0 -> L-BFGS-B -> L z + mu, with the NumPy/SciPy/
    sklearn restatement of the reference (the reference itself needs JAX, absent from this image)."""
    from oracle import mellon_oracle as O

    timings = {}
    t0 = time.perf_counter()
    fit = O.fit_density(x, cov_func_curry=oracle_cov(cov_name), landmarks=lm, nn_distances=nn, timings=timings)
    return time.perf_counter() - t0, fit, timings


def cpu_sample(args, x, lm, nn):
    ns = min(args.cpu_sample, x.shape[0])
    return np.ascontiguousarray(x[:ns]), lm, np.ascontiguousarray(nn[:ns]), ns


def threads_used():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return int(max(n)) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port; the reference
    cannot be installed here: it needs jax/jaxlib/jaxopt/pynndescent and there is no network) on a
    bounded sample of the same workload.  Rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sklearn.neighbors import NearestNeighbors

    ns = min(args.cpu_sample, args.cells)
    x = make_cells(args.cells, args.dims)
    lm = pick_landmarks(x, args.landmarks)
    xs = np.ascontiguousarray(x[:ns])
    del x
    # the sample's own exact 1-NN distances (host, outside the timed region)
    nn = NearestNeighbors(n_neighbors=2).fit(xs).kneighbors(xs)[0][:, 1]
    times, last = [], None
    for i in range(args.warmup + args.steps):
        dt, fit, tm = cpu_fit(xs, lm, nn, args.cov)
        if i >= args.warmup:
            times.append(dt)
            last = tm
    total = float(np.sum(times))
    value = ns * len(times) / total
    cores = threads_used()
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {ns} cells of the workload (all {args.landmarks} landmarks), every stage "
                                   f"is O(N): cells/s is size-independent; host has {os.cpu_count()} logical CPUs",
                         "stages_s": {k: round(float(v), 3) for k, v in (last or {}).items()}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- CUDA arm -----------------------------------------------------------------------------------------
def run_b200(args):
    import mellon_b200 as mb
    from mellon_b200 import cov as C
    from mellon_b200 import distributed as dist

    logger = mb.setup_logging()
    logger.setLevel("WARNING")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    be = mb.get_backend()  # attaches the NCCL communicator when WORLD_SIZE > 1
    cov_curry = getattr(C, args.cov)

    x = make_cells(args.cells, args.dims)
    lm = pick_landmarks(x, args.landmarks)
    nn = be.nn_distances(x)  # exact 1-NN on the device, outside the timed region
    xd = be.upload(x, sharded=True)  # resident cell block of this rank

    def step(x_in):
        est = mb.DensityEstimator(cov_func_curry=cov_curry, landmarks=lm, nn_distances=nn, check_rank=False)
        dens = est.fit_predict(x_in)
        return est, dens

    # --- device-resident arm: `value` -------------------------------------------------------------
    for _ in range(args.warmup):
        est, dens = step(xd)
        del est
    be.prof_enable(True)
    be.prof_reset()
    sampler = ClockSampler(be.device)
    dist.barrier()
    be.sync()
    launches0 = be.launch_count()
    if rank == 0 and not args.no_clocks:
        sampler.start()
    be.timer_start(0)
    nfev = []
    last_est = None
    for _ in range(args.steps):
        last_est = None  # frees the previous step's 40 GB factor before the next one is built
        est, dens = step(xd)
        nfev.append(int(est.opt_state.num_fun_eval))
        nit = int(est.opt_state.iter_num)
        last_est = est
        del est
    ms = be.timer_stop(0)
    be.sync()
    dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = be.launch_count() - launches0
    prof = be.prof_read()
    be.prof_enable(False)
    ms = dist.host_max(ms)
    ms_per_step = ms / args.steps
    value = args.cells / (ms_per_step * 1e-3)
    checksum = float(np.sum(dens))

    # --- predict(): conditional-mean kernel K7 on out-of-sample queries (host in, host out) ---------
    predict = None
    if args.predict_queries > 0:
        q_total = args.predict_queries
        xq = np.random.default_rng(2).random((q_total, args.dims))
        predictor = last_est.predict          # builds the predictor: weights = Lp^-T z
        predictor(xq[:1024])                  # warm-up
        be.prof_enable(True)
        be.prof_reset()
        dist.barrier()
        be.sync()
        t0 = time.perf_counter()
        pq = predictor(xq)
        be.sync()
        dt_q = dist.host_max(time.perf_counter() - t0)
        n_mv, ms_mv, _ = be.prof_read()["matvec"]
        be.prof_enable(False)
        predict = {"queries": q_total, "value": q_total / dt_q, "unit": "queries/s (host in, host out, all GPUs)",
                   "k7_kernel_ms": ms_mv, "k7_launches": n_mv,
                   "k7_elements_per_s": (q_total / max(world, 1)) * args.landmarks / (ms_mv * 1e-3) if ms_mv else None,
                   "checksum": float(np.sum(pq))}
        del predictor, xq
    last_est = None

    # --- end-to-end arm: host buffers in, host result out -----------------------------------------
    e2e = None
    if not args.no_e2e:
        xh = be.pinned_empty(x.shape)
        xh[...] = x
        step(xh)  # one warm-up through the host path
        dist.barrier()
        be.sync()
        be.h2d_bytes = be.d2h_bytes = 0
        be.timer_start(1)
        for _ in range(args.steps):
            est, dens_h = step(xh)
            del est
        ms_e = dist.host_max(be.timer_stop(1))
        h2d = dist.host_sum(be.h2d_bytes) / args.steps
        d2h = dist.host_sum(be.d2h_bytes) / args.steps
        e2e = {"value": args.cells / (ms_e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e / args.steps}
        dist.barrier()

    if rank != 0:
        return

    # --- roofline of the dominant HBM-bound kernel of north_star: K1, the fused N x M build --------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    n_cov, ms_cov, bytes_cov = prof["cov"]
    achieved = bytes_cov / (ms_cov * 1e-3) / 1e9 if ms_cov > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json"))).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {
        "kernel": "cov_mma_kernel (K1: fused pairwise distance + covariance, K_NM build; FP64-issue bound, see DESIGN.md)",
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak if hbm_peak else None, "traffic": traffic, "peak_source": peak_src,
        "launches": n_cov, "avg_launch_ms": ms_cov / max(n_cov, 1),
        "algorithmic_bytes_per_launch": bytes_cov / max(n_cov, 1),
    }
    n_lg, ms_lg, bytes_lg = prof["lossgrad"]
    n_gm, ms_gm, flops_gm = prof["gemm"]
    kernels = {
        "k1_cov_build": {"launches": n_cov, "ms_per_step": ms_cov / args.steps, "share": ms_cov / ms},
        "k5_loss_grad": {"launches": n_lg, "ms_per_step": ms_lg / args.steps, "share": ms_lg / ms,
                         "achieved_GBps": bytes_lg / (ms_lg * 1e-3) / 1e9 if ms_lg else None,
                         "hbm_frac": bytes_lg / (ms_lg * 1e-3) / 1e9 / hbm_peak if ms_lg else None},
        "fp64_gemm (K3 trsm + K4 gram + K2 updates)": {
            "launches": n_gm, "ms_per_step": ms_gm / args.steps, "share": ms_gm / ms,
            "achieved_TFLOPs": flops_gm / (ms_gm * 1e-3) / 1e12 if ms_gm else None},
    }

    cpu_baseline = None
    parity = None
    if not args.no_cpu_baseline and world == 1:
        xs, lms, nns, ns = cpu_sample(args, x, lm, nn)
        dt, fit, tm = cpu_fit(xs, lms, nns, args.cov)
        # parity on the very sample the CPU arm just fitted: the same cells through the CUDA path
        dens_s = mb.DensityEstimator(cov_func_curry=cov_curry, landmarks=lms, nn_distances=nns,
                                     check_rank=False).fit_predict(xs)
        ref_s = np.asarray(fit.log_density_x)
        diff = dens_s - ref_s
        parity = {
            "sample": f"first {ns} cells, all {args.landmarks} landmarks: CUDA fit_predict vs the CPU oracle's",
            # the reference's own acceptance metric (tests/test_density_estimator.py:30-44): std(a - b) / std(b)
            "rel_std_err_log_density": float(np.std(diff) / np.std(ref_s)),
            "max_abs_err_over_max_abs": float(np.max(np.abs(diff)) / np.max(np.abs(ref_s))),
            # element-wise relative error: log densities cross zero on this workload, so this one is dominated by
            # the cells whose log density is ~0 (reported for completeness, not judged)
            "max_elementwise_rel_err": float(np.max(np.abs(diff) / np.abs(ref_s))),
            "min_abs_log_density": float(np.min(np.abs(ref_s))), "max_abs_log_density": float(np.max(np.abs(ref_s))),
            "tolerance": 1e-5,
        }
        parity["ok"] = bool(parity["rel_std_err_log_density"] < 1e-5 and parity["max_abs_err_over_max_abs"] < 1e-5)
        cpu_baseline = {
            "value": ns / dt, "unit": UNIT, "cores": threads_used(), "kind": "port",
            "sample": f"first {ns} cells of the workload (all {args.landmarks} landmarks, their nn_distances taken "
                      f"from the full set); every stage is O(N) so cells/s is size-independent; {dt:.1f} s on "
                      f"{os.cpu_count()} logical CPUs",
            "stages_s": {k: round(float(v), 3) for k, v in tm.items()},
        }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels,
        "predict": predict, "cpu_baseline": cpu_baseline, "parity": parity, "lbfgsb": {"nfev_per_step": nfev, "nit_last": nit},
        "log_density_checksum": checksum,
    }
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
