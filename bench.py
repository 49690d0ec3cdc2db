#!/usr/bin/env python
"""bench.py -- genes fitted/sec for a full runLineagePulse-equivalent (impulse vs constant, ZINB,
logistic drop-out: fits A-D + LRT) on simulateContinuousDataSet-style synthetic data.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the host cores (CPU oracle)

One "step" = one pass of the whole hot path over the workload (BASELINE.json configs[1]:
5 000 genes x 10 000 cells per GPU; at N > 1 every rank holds its own 5 000-gene shard of a
5 000*N-gene matrix => weak scaling, gene-sharded, with the per-cell drop-out sums and the EM
log-likelihood all-reduced over NCCL).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genes fitted/sec (runLineagePulse impulse ZINB)"
UNIT = "genes/s"

# canonical algorithmic FP64 flops per observation-likelihood evaluation
FLOP_NONZERO = 260.0      # NB / ZINB non-zero term (1 lgamma + 2 log)
FLOP_ZERO_ZINB = 150.0    # 2 log + 1 exp
FLOP_ZERO_NB = 55.0       # 1 log
FLOP_IMPULSE = 140.0      # 2 exp + 3 div
FLOP_DROPFIT = 62.0       # pi from theta (q = 1): 60 + 2q


def flops_per_ole(mu_impulse, zinb, zero_frac):
    f = (1 - zero_frac) * FLOP_NONZERO + zero_frac * (FLOP_ZERO_ZINB if zinb else FLOP_ZERO_NB)
    return f + (FLOP_IMPULSE if mu_impulse else 0.0)


def make_workload(n_genes, n_cells, seed):
    from lineagepulse_b200.simulate import simulateContinuousDataSet
    n_const = n_genes // 2
    n_lin = n_genes // 4
    n_imp = n_genes - n_const - n_lin
    sim = simulateContinuousDataSet(n_cells, n_const, n_lin, n_imp, scaMumax=100, scaSDMuAmplitude=3,
                                    vecGeneWiseDropoutRates=np.full(n_genes, 0.1), seed=seed, dtype=np.int32)
    X = sim["counts"]
    X = np.ascontiguousarray(X[X.sum(1) > 0])
    return X, sim["annot"]["continuous"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def native_pipeline(eng, design, A, min_rowmean):
    """fitLPModels (fitWrapperLP.R:71-289) + runDEAnalysis (runHypothesisTests.R:40-130) on one shard.
    Returns (result dict, per-fit work stats)."""
    from lineagepulse_b200.engine import lrt
    ctl = A.default_control()
    mA = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mB = A.Model(A.MU_IMPULSE, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mC = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_NONE, 0)
    mD = A.Model(A.MU_IMPULSE, A.DISP_CONSTANT, A.DROP_NONE, 0)
    work = {}

    def run(tag, m, p, fit_drop):
        eng.reset_stats()
        r = eng.fit_model(m, p, fit_drop, ctl, min_rowmean)
        work[tag] = eng.stats()
        return r

    pA = A.ParamsData(design, mA)
    run("A", mA, pA, True)
    # fits B, C, D are independent given A's drop-out model: one concurrent call (lpb200_fit_model_multi)
    pB = A.ParamsData(design, mB)
    pB.mat_drop = pA.mat_drop.copy(order="F")
    pC = A.ParamsData(design, mC)
    pD = A.ParamsData(design, mD)
    eng.reset_stats()
    res = eng.fit_model_multi([(mB, pB), (mC, pC), (mD, pD)], ctl, min_rowmean)
    st = eng.stats()
    for tag, r in zip("BCD", res):
        work[tag] = {"fn_evals": 0, "ole": r["ole"], "launches": 0, "kernel_ms": 0.0}
    work["BCD"] = st
    eng.reset_stats()
    llf, llr = eng.eval_loglik(mB, pB), eng.eval_loglik(mA, pA)
    llf_nb, llr_nb = eng.eval_loglik(mD, pD), eng.eval_loglik(mC, pC)
    work["DE"] = eng.stats()
    ddf = design.deg_freedom(mB) - design.deg_freedom(mA)
    p, padj = lrt(llf, llr, ddf)
    p_nb, padj_nb = lrt(llf_nb, llr_nb, ddf)
    return {"p": p, "padj": padj, "padj_nb": padj_nb, "llf": llf, "llr": llr}, work


def oracle_pipeline(oracle, X, t, A, ns):
    """The same pipeline through the CPU oracle (reference algorithm on the host cores)."""
    G, N = X.shape
    d = A.DesignData(G, N, pseudotime=t, impulse_basis=ns(t, 4))
    ctl = A.default_control()
    mA = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mB = A.Model(A.MU_IMPULSE, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mC = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_NONE, 0)
    mD = A.Model(A.MU_IMPULSE, A.DISP_CONSTANT, A.DROP_NONE, 0)
    pA = A.ParamsData(d, mA)
    oracle.fit_model(X, d, mA, pA, True, ctl)
    pB = A.ParamsData(d, mB)
    pB.mat_drop = pA.mat_drop.copy(order="F")
    oracle.fit_model(X, d, mB, pB, False, ctl)
    pC = A.ParamsData(d, mC)
    oracle.fit_model(X, d, mC, pC, False, ctl)
    pD = A.ParamsData(d, mD)
    oracle.fit_model(X, d, mD, pD, False, ctl)
    llf, llr = oracle.eval_loglik(X, d, mB, pB), oracle.eval_loglik(X, d, mA, pA)
    oracle.eval_loglik(X, d, mD, pD), oracle.eval_loglik(X, d, mC, pC)
    return oracle.lrt(llf, llr, 6)


def cpu_sample(X, n_sample):
    idx = np.linspace(0, X.shape[0] - 1, n_sample).astype(int)
    return np.ascontiguousarray(X[idx])


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path.  R is not installable
    here, so this is the CPU oracle (plain-C restatement), one thread per host core."""
    if rank != 0:
        return
    from lineagepulse_b200 import _cabi as A
    from lineagepulse_b200.splines import ns
    from tests.oracle_binding import Oracle
    oracle = Oracle(nthreads=os.cpu_count() or 1)
    X, t = make_workload(args.genes, args.cells, seed=2)
    Xs = cpu_sample(X, args.cpu_sample_genes)
    for _ in range(args.warmup):
        oracle_pipeline(oracle, Xs[: max(2, len(Xs) // 4)], t, A, ns)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_pipeline(oracle, Xs, t, A, ns)
    dt = (time.perf_counter() - t0) / args.steps
    value = len(Xs) / dt
    sample = (f"{len(Xs)} genes (evenly spaced rows) x {args.cells} cells of the {X.shape[0]}-gene workload per step; "
              "genes/s extrapolates linearly in genes")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C2: {args.genes} genes x {args.cells} cells, impulse vs constant, ZINB, "
                                   "logistic drop-out, fits A-D + LRT (simulateContinuousDataSet recipe, seed 2)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--genes", type=int, default=5000, help="genes per GPU (C2: 5000)")
    ap.add_argument("--cells", type=int, default=10000)
    ap.add_argument("--cpu-sample-genes", type=int, default=16)
    ap.add_argument("--warmup-genes", type=int, default=0, help="0: warm-up steps are full steps; >0: warm up on a gene subset")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from lineagepulse_b200 import _cabi as A
    from lineagepulse_b200 import api
    from lineagepulse_b200.distributed import make_allreduce
    from lineagepulse_b200.engine import Engine
    from lineagepulse_b200.splines import ns

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()

    X, t = make_workload(args.genes, args.cells, seed=2 + 1000 * rank)
    G, N = X.shape
    zero_frac = float((X == 0).mean())
    design = A.DesignData(G, N, pseudotime=t, impulse_basis=ns(t, 4))
    # rowMeans minimum over ALL shards (only logistic_ofMu needs it; kept for completeness)
    min_rowmean = float("nan")

    eng = Engine(local_rank)
    eng.set_stream(stream.cuda_stream)
    if world > 1:
        eng.set_collective(rank, world, make_allreduce())
    fp64_peak = eng.fp64_peak_tflops()
    eng.load_counts_i32(X)
    eng.set_design(design)

    # warm-up steps (untimed)
    if 0 < args.warmup_genes < G:
        Xw = np.ascontiguousarray(X[: args.warmup_genes])
        dw = A.DesignData(Xw.shape[0], N, pseudotime=t, impulse_basis=ns(t, 4))
        engw = Engine(local_rank)
        engw.set_stream(stream.cuda_stream)
        if world > 1:
            engw.set_collective(rank, world, make_allreduce())
        engw.load_counts_i32(Xw)
        engw.set_design(dw)
        for _ in range(args.warmup):
            native_pipeline(engw, dw, A, min_rowmean)
        del engw
    else:
        for _ in range(args.warmup):
            native_pipeline(eng, design, A, min_rowmean)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: counts resident in HBM ("value") -------------------------------------
    sampler = ClockSampler(local_rank)
    step_ms, works = [], []
    launches = 0
    barrier()
    if rank == 0:
        sampler.start()
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res, work = native_pipeline(eng, design, A, min_rowmean)
        e1.record(stream)
        barrier()
        step_ms.append(e0.elapsed_time(e1))
        works.append(work)
        launches += sum(int(w["launches"]) for w in work.values())
    clocks = sampler.stop() if rank == 0 else None
    t_local = float(np.sum(step_ms)) / 1e3
    tt = torch.tensor([t_local, float(G)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        gsum = tt.clone()
        dist.all_reduce(gsum, op=dist.ReduceOp.SUM)
        t_all, g_all = float(tmax[0]), float(gsum[1])
    else:
        t_all, g_all = t_local, float(G)
    value = g_all * args.steps / t_all
    ms_per_step = t_all / args.steps * 1e3

    # ---- timed region 2: end to end through the public API with HOST buffers ("e2e") ----------
    # runLineagePulse(counts, dfAnnotation, ...) from pinned host memory: context creation, H2D of the
    # counts, packing, all fits, D2H of every result.  Single-GPU call; at N > 1 each rank runs it on
    # its own shard (no cross-shard drop-out sums) and the slowest rank counts.
    pinned = torch.empty((G, N), dtype=torch.int32).pin_memory()
    pinned.numpy()[...] = X
    annot = {"cell": np.array([f"cell_{j + 1}" for j in range(N)]), "continuous": t}
    e2e_ms = []
    n_e2e = max(1, min(args.steps, 2))
    if args.warmup > 0:  # one untimed pass of the public API path (first-use allocations of a fresh context)
        api.runLineagePulse(pinned.numpy(), annot, strMuModel="impulse", strDropModel="logistic",
                            boolVerbose=False, device=local_rank)
    for _ in range(n_e2e):
        flush.fill_(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        obj = api.runLineagePulse(pinned.numpy(), annot, strMuModel="impulse", strDropModel="logistic",
                                  boolVerbose=False, device=local_rank)
        n_de = int((obj.dfResults["padj"] < 0.05).sum())
        e1.record(stream)
        barrier()
        e2e_ms.append(e0.elapsed_time(e1))
    te = torch.tensor([float(np.sum(e2e_ms)) / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = g_all * n_e2e / float(te[0])
    h2d = int(X.nbytes + 8 * N * 6)  # counts + pseudotime + 4-column impulse basis + size factors
    d2h = int(obj.dfResults.shape[0] * 14 * 8 + 8 * G * (7 + 1 + 1 + 1) * 2 + 8 * N)  # results table, parameter matrices, drop-out model

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fit_mudisp_kernel on the impulse fits B and D) --------
    w = works[-1]
    flop_B = w["B"]["ole"] * flops_per_ole(True, True, zero_frac)
    flop_C = w["C"]["ole"] * flops_per_ole(False, False, zero_frac)
    flop_D = w["D"]["ole"] * flops_per_ole(True, False, zero_frac)
    ker_s = w["BCD"]["kernel_ms"] / 1e3   # span of the three overlapping fit_mudisp launches
    achieved = (flop_B + flop_C + flop_D) / ker_s / 1e12
    share = ker_s * 1e3 / step_ms[-1]
    ct_bytes = 2 if int(X.max()) < 65535 else 4   # the library stores uint16 unless a count needs int32
    ct_name = "uint16_t" if ct_bytes == 2 else "int32_t"
    count_bytes = float(ct_bytes) * G * N   # gene-major count matrix, read once per fitMuDisp call
    perm_bytes = 4.0 * G * N               # per-gene cell permutation (non-zero cells first)
    roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                # dram__bytes_read + write per launch from profiles/ncu_fit_mudisp_impulse_zinb_r1.csv
                # (110.4 MB read + 44.4 MB written at 1 184 genes x 10 000 cells, u16 counts), scaled to this launch's matrix
                "traffic": 154.74e6 * (G * N) / (1184.0 * 10000.0) * ((ct_bytes + 4.0) / 6.0),
                "traffic_unit": "bytes per launch (ncu, scaled by matrix size)",
                "kernel": f"fit_mudisp_kernel<{ct_name}> (fits B, C, D: three overlapping launches)",
                "kernel_share_of_step": share,
                "ole_per_s": (w["B"]["ole"] + w["C"]["ole"] + w["D"]["ole"]) / ker_s,
                "peak_source": "FP64 DFMA chain probe measured live in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "hbm": {"algorithmic_bytes_per_launch": count_bytes + perm_bytes,
                        "achieved_GBps": 2 * (count_bytes + perm_bytes) / ker_s / 1e9,
                        "peak_GBps": _hbm_peak(), "note": "count bytes are <0.1% of HBM bandwidth: FP64-bound (SURVEY.md F7)"}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from tests.oracle_binding import Oracle
        oracle = Oracle(nthreads=os.cpu_count() or 1)
        Xs = cpu_sample(X, args.cpu_sample_genes)
        t0 = time.perf_counter()
        p_o, padj_o = oracle_pipeline(oracle, Xs, t, A, ns)
        dt = time.perf_counter() - t0
        cpu = {"value": len(Xs) / dt, "unit": UNIT, "cores": oracle.nthreads, "kind": "port",
               "sample": f"{len(Xs)} evenly spaced genes x {N} cells of the workload, full pipeline (fits A-D + LRT), "
                         f"{dt:.1f} s; CPU restatement of the reference algorithm (R not installable here)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C2: {G} genes x {N} cells per GPU, impulse vs constant, ZINB, logistic drop-out, "
                                   "fits A-D + LRT (simulateContinuousDataSet recipe 50/25/25 const/lin/impulse, "
                                   "mumax 100, sd 3, drop-out 0.1, seed 2)",
                       "genes_total": int(g_all), "cells": N, "zero_fraction": zero_frac,
                       "l2": "flushed between steps (256 MiB write)", "parallelism": f"gene-sharded x{world}",
                       "de_genes_q05": n_de},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te[0]) / n_e2e * 1e3, "api": "lineagepulse_b200.runLineagePulse(host counts)"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "work": {k: {kk: (float(vv) if kk == "kernel_ms" else int(vv)) for kk, vv in v.items()} for k, v in w.items()}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0  # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    main()
