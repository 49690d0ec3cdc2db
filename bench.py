#!/usr/bin/env python
"""bench.py -- genes fitted/sec for a full runLineagePulse-equivalent (H1 vs constant, ZINB, logistic drop-out:
fits A-D + LRT) on simulateContinuousDataSet-style synthetic data.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the host cores (CPU oracle)
    python bench.py --config c4|c5 ...        # group-wise DE with a batch confounder / splines mean model

One "step" = one pass of the whole hot path over the workload.  Default workload = BASELINE.json configs[1]
(C2: 5 000 genes x 10 000 cells, impulse vs constant) per GPU; at N > 1 every rank holds its own 5 000-gene
shard of a 5 000 N-gene matrix (weak scaling, gene-sharded, the per-cell drop-out sums and the EM
log-likelihood all-reduced by the library's own NCCL communicator).  In addition every run makes ONE pass over
the north_star matrix (C3: 20 000 genes x 100 000 cells, the same matrix for every N, gene-sharded over the N
ranks) and reports it under "strong" with a digest of the gathered per-gene results, which must not depend on N.
Prints ONE JSON line on rank 0.
"""
import argparse
import hashlib
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
FLOP_MU = {"impulse": 140.0, "constant": 0.0, "groups": 0.0, "splines": 2 * 6 + 40.0}   # + mean model
FLOP_DROPFIT = 62.0       # pi from theta (q = 1): 60 + 2q

CONFIGS = {
    # name: (genes per GPU, cells, H1 mean model, description)
    "c2": (5000, 10000, "impulse", "C2: impulse vs constant, ZINB, logistic drop-out"),
    "c4": (20000, 50000, "groups", "C4: group-wise DE (4 groups) with a batch confounder (3 batches, vecConfoundersMu), ZINB, logistic drop-out"),
    "c5": (10000, 50000, "splines", "C5: splines mean model (df 6) vs constant, ZINB, logistic drop-out"),
}


def flops_per_ole(mu_model, zinb, zero_frac, n_conf=0):
    f = (1 - zero_frac) * FLOP_NONZERO + zero_frac * (FLOP_ZERO_ZINB if zinb else FLOP_ZERO_NB)
    return f + FLOP_MU[mu_model] + n_conf


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


def cell_covariates(n_cells, seed):
    """pseudotime, 4 pseudotime-quartile groups, 3 random batches."""
    t = np.arange(n_cells, dtype=np.float64) / max(n_cells - 1, 1)
    groups = np.minimum((t * 4).astype(np.int64), 3)
    batch = np.random.default_rng(seed + 17).integers(0, 3, n_cells)
    return t, groups, batch


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


def models_for(A, h1):
    mu1 = A.MU_MODELS[h1]
    mA = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mB = A.Model(mu1, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mC = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_NONE, 0)
    mD = A.Model(mu1, A.DISP_CONSTANT, A.DROP_NONE, 0)
    return mA, mB, mC, mD


def native_pipeline(eng, design, A, min_rowmean, h1="impulse", with_lrt=True):
    """fitLPModels (fitWrapperLP.R:71-289) + runDEAnalysis (runHypothesisTests.R:40-130) on one shard.
    Returns (result dict, per-fit work stats)."""
    from lineagepulse_b200.engine import lrt
    ctl = A.default_control()
    mA, mB, mC, mD = models_for(A, h1)
    work = {}
    pA = A.ParamsData(design, mA)
    eng.reset_stats()
    eng.fit_model(mA, pA, True, ctl, min_rowmean)
    work["A"] = eng.stats()
    work["A"]["allreduce_calls"], work["A"]["allreduce_bytes"] = eng.comm_stats()
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
    out = {"llf": llf, "llr": llr, "llf_nb": llf_nb, "llr_nb": llr_nb, "pA": pA, "pB": pB, "pC": pC, "pD": pD,
           "ddf": design.deg_freedom(mB) - design.deg_freedom(mA)}
    if with_lrt:
        out["p"], out["padj"] = lrt(llf, llr, out["ddf"])
        out["p_nb"], out["padj_nb"] = lrt(llf_nb, llr_nb, out["ddf"])
    return out, work


def oracle_pipeline(oracle, X, t, A, ns, h1="impulse", design=None):
    """The same pipeline through the CPU oracle (reference algorithm on the host cores)."""
    G, N = X.shape
    d = design or A.DesignData(G, N, pseudotime=t, impulse_basis=ns(t, 4))
    ctl = A.default_control()
    mA, mB, mC, mD = models_for(A, h1)
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
    llf_nb, llr_nb = oracle.eval_loglik(X, d, mD, pD), oracle.eval_loglik(X, d, mC, pC)
    p, padj = oracle.lrt(llf, llr, d.deg_freedom(mB) - d.deg_freedom(mA))
    return {"llf": llf, "llr": llr, "llf_nb": llf_nb, "llr_nb": llr_nb, "p": p, "padj": padj, "pA": pA, "pC": pC, "pD": pD}


def cpu_sample_rows(n_rows, n_sample, offset=0):
    """Evenly spaced rows of the workload (disjoint sets for different offsets)."""
    n_sample = min(n_sample, n_rows)
    step = n_rows / n_sample
    return np.unique(((np.arange(n_sample) + 0.5 * (offset % 2)) * step).astype(int) % n_rows)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path.  R is not installable here, so this is
    the CPU oracle (plain-C restatement of the reference algorithm) with one worker thread per host core -- the
    analogue of BiocParallel's MulticoreParam(workers = scaNProc), main_LineagePulse.R:271-275.  The K steps are
    K disjoint gene batches of the workload handed to the worker pool together, as MulticoreParam would get a
    larger matrix: the pool is never drained between steps, so the number measures the cores, not the longest
    gene of a small batch."""
    if rank != 0:
        return
    from lineagepulse_b200 import _cabi as A
    from lineagepulse_b200.splines import ns
    from tests.oracle_binding import Oracle
    cores = os.cpu_count() or 1
    oracle = Oracle(nthreads=cores)
    genes, cells, h1, desc = CONFIGS[args.config]
    genes, cells = args.genes or genes, args.cells or cells
    if args.config != "c2":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm is implemented for the headline config (c2) only"}))
        return
    X, t = make_workload(genes, cells, seed=2)
    # warm-up: W small batches; they also measure the per-gene cost that sizes the timed batches
    t0 = time.perf_counter()
    n_warm = 0
    for w in range(max(args.warmup, 1)):
        idx = cpu_sample_rows(X.shape[0], cores, offset=w)
        oracle_pipeline(oracle, np.ascontiguousarray(X[idx]), t, A, ns)
        n_warm += len(idx)
        if time.perf_counter() - t0 > 40:
            break
    core_s_per_gene = (time.perf_counter() - t0) * cores / n_warm      # an over-estimate (batches of one gene per core)
    budget_s = args.cpu_budget_s
    total = int(budget_s * cores / max(core_s_per_gene, 1e-3))
    per_step = args.cpu_sample_genes or max(1, min(total // args.steps, 8 * cores))
    per_step = max(per_step, -(-4 * cores // args.steps))          # the pool holds >= 4 genes per core
    idx = cpu_sample_rows(X.shape[0], per_step * args.steps)
    Xs = np.ascontiguousarray(X[idx])
    t0 = time.perf_counter()
    oracle_pipeline(oracle, Xs, t, A, ns)
    dt_total = time.perf_counter() - t0
    dt = dt_total / args.steps
    value = len(idx) / dt_total
    sample = (f"{args.steps} steps x {len(idx) // args.steps} genes = {len(idx)} evenly spaced rows x {cells} cells of the "
              f"{X.shape[0]}-gene workload, handed to the {cores}-thread pool as one batch (full pipeline: fits A-D + LRT), "
              f"{dt_total:.1f} s; genes/s extrapolates linearly in genes")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C2: {genes} genes x {cells} cells, impulse vs constant, ZINB, "
                                   "logistic drop-out, fits A-D + LRT (simulateContinuousDataSet recipe, seed 2)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "genes_per_s_per_core": value / cores},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def parity_block(eng_res, X, t, idx, oracle, A, ns, cpu):
    """GPU fits against the oracle at the benchmark's own size (the oracle gets the GPU's fitted drop-out model,
    fitWrapperLP.R:152-238: fits B, C, D given A's lsDropModel)."""
    Xs = np.ascontiguousarray(X[idx])
    G, N = Xs.shape
    d = A.DesignData(G, N, pseudotime=t, impulse_basis=ns(t, 4))
    ctl = A.default_control()
    mA, mB, mC, mD = models_for(A, "impulse")
    drop = eng_res["pA"].mat_drop
    pA = A.ParamsData(d, mA)
    pA.mat_drop = drop.copy(order="F")
    rA = oracle.fit_model(Xs, d, mA, pA, False, ctl)          # constant ZINB given the GPU's drop-out model
    pB = A.ParamsData(d, mB)
    pB.mat_drop = drop.copy(order="F")
    rB = oracle.fit_model(Xs, d, mB, pB, False, ctl)          # impulse ZINB given the same
    rel = lambda a, b: np.abs(a - b) / np.abs(b)
    r_h0 = rel(eng_res["llr"][idx], rA["ll"])
    r_c = rel(eng_res["llr_nb"][idx], cpu["llr_nb"])
    r_b = rel(eng_res["llf"][idx], rB["ll"])
    r_d = rel(eng_res["llf_nb"][idx], cpu["llf_nb"])
    from lineagepulse_b200.engine import lrt
    p_g, padj_g = lrt(eng_res["llf"][idx], eng_res["llr"][idx], 6)
    p_o, padj_o = oracle.lrt(rB["ll"], rA["ll"], 6)
    de_g, de_o = padj_g < 0.05, padj_o < 0.05
    near = np.abs(padj_o - 0.05) < 1e-3
    union = ((de_g | de_o) & ~near).sum()
    from scipy.stats import spearmanr
    return {"n_genes": int(G), "n_cells": int(N), "H0_ll_max_rel": float(r_h0.max()), "NB_const_ll_max_rel": float(r_c.max()),
            "mu_const_max_rel": float(rel(eng_res["pC"].mat_mu[idx, 0], cpu["pC"].mat_mu[:, 0]).max()),
            "impulse_frac_1e-6": float((r_b < 1e-6).mean()), "impulse_frac_1e-4": float((r_b < 1e-4).mean()),
            "impulse_nb_frac_1e-6": float((r_d < 1e-6).mean()), "impulse_ll_median_rel": float(np.median(r_b)),
            "p_spearman": float(spearmanr(p_g, p_o).correlation),
            "de_set_jaccard": float(((de_g & de_o & ~near).sum() / union) if union else 1.0),
            "de_genes_gpu": int(de_g.sum()), "de_genes_oracle": int(de_o.sum()),
            "note": "GPU vs CPU oracle on the sampled genes at full N; the oracle's ZINB fits use the GPU's fitted drop-out "
                    "model (fit A over all genes); impulse fits are optimiser-path chaotic in the reference algorithm itself "
                    "(tests/test_mirror.py), constant fits are the north_star 1e-6 / 1e-4 bar"}


def strong_pass(args, A, torch, dist, rank, world, local_rank, fp64_peak):
    """ONE pass of the hot path over the north_star matrix (C3), the same matrix for every N, gene-sharded over the
    N ranks by shard_bounds; counts generated on the device shard by shard (the draws of a gene do not depend on
    the sharding).  Returns the record for the JSON line (rank 0) -- the digest of the gathered per-gene results
    must be the same for N = 1, 2, 4, 8."""
    from lineagepulse_b200.distributed import init_comm, shard_bounds
    from lineagepulse_b200.engine import Engine, lrt
    from lineagepulse_b200.simulate import simulate_rows_device, workload_recipe
    from lineagepulse_b200.splines import ns
    G, N = args.strong_genes, args.strong_cells
    rec = workload_recipe(G, seed=3)
    b = shard_bounds(G, world)
    lo, hi = b[rank], b[rank + 1]
    t = np.arange(N, dtype=np.float64) / max(N - 1, 1)
    t_gen0 = time.perf_counter()
    rows = simulate_rows_device(rec, N, lo, hi, f"cuda:{local_rank}")
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen0
    zero_frac = float((rows == 0).double().mean())
    eng = Engine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    if world > 1:
        init_comm(eng)
    eng.load_counts_device_i32(rows.data_ptr(), hi - lo, N)
    del rows
    torch.cuda.empty_cache()
    design = A.DesignData(hi - lo, N, pseudotime=t, impulse_basis=ns(t, 4))
    eng.set_design(design)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    res, work = native_pipeline(eng, design, A, float("nan"), "impulse", with_lrt=False)
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_s = float(tt[0]) / 1e3
    local = np.stack([res["llf"], res["llr"], res["llf_nb"], res["llr_nb"]], 1)
    stats = {"ole": [int(work[k]["ole"]) for k in "BCD"], "kernel_ms": float(work["BCD"]["kernel_ms"]),
             "fitA_ms": float(work["A"]["kernel_ms"]), "allreduce_calls": int(work["A"]["allreduce_calls"]),
             "allreduce_bytes": int(work["A"]["allreduce_bytes"])}
    if world > 1:
        parts, allst = [None] * world, [None] * world
        dist.all_gather_object(parts, local)
        dist.all_gather_object(allst, stats)
        ll_all = np.concatenate(parts, 0)
    else:
        ll_all, allst = local, [stats]
    eng.close()
    if rank != 0:
        return None
    p, padj = lrt(ll_all[:, 0], ll_all[:, 1], res["ddf"])
    h = hashlib.sha256()
    for a in (ll_all, p, padj):
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    ole = np.sum([s["ole"] for s in allst], 0)
    flop = (ole[0] * flops_per_ole("impulse", True, zero_frac) + ole[1] * flops_per_ole("constant", False, zero_frac)
            + ole[2] * flops_per_ole("impulse", False, zero_frac))
    ker_s = max(s["kernel_ms"] for s in allst) / 1e3
    achieved = flop / ker_s / 1e12
    return {"workload": f"C3 {G}x{N}: impulse vs constant, ZINB, logistic drop-out, fits A-D + LRT; ONE matrix (seed 3) "
                        f"gene-sharded over {world} GPU(s)", "n_gpus": world, "step_s": step_s, "genes_per_s": G / step_s,
            "roofline_frac": achieved / (fp64_peak * world), "achieved_tflops": achieved, "kernel_s_max_over_ranks": ker_s,
            "kernel_share_of_step": ker_s / step_s, "fitA_ms_max_over_ranks": max(s["fitA_ms"] for s in allst),
            "fitpi_allreduce": {"calls": allst[0]["allreduce_calls"], "payload_bytes_per_rank": allst[0]["allreduce_bytes"]},
            "de_genes_q05": int((padj < 0.05).sum()), "digest": h.hexdigest()[:32], "data_generation_s": t_gen,
            "zero_fraction": zero_frac,
            "what_bounds_it": "fit_mudisp kernels (FP64); the per-rank item-length tail of the dynamic queue grows as the "
                              "shard shrinks; fitPi all-reduce payload above"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--genes", type=int, default=0, help="genes per GPU (default: the config's)")
    ap.add_argument("--cells", type=int, default=0)
    ap.add_argument("--cpu-sample-genes", type=int, default=0, help="0: 4 x host cores (cpu_baseline) / sized by --cpu-budget-s (reference arm)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="reference arm: target duration of the timed CPU run")
    ap.add_argument("--warmup-genes", type=int, default=0, help="0: warm-up steps are full steps; >0: warm up on a gene subset")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the C3 strong-scaling pass")
    ap.add_argument("--strong-genes", type=int, default=20000)
    ap.add_argument("--strong-cells", type=int, default=100000)
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
    from lineagepulse_b200.distributed import init_comm
    from lineagepulse_b200.engine import Engine
    from lineagepulse_b200.splines import ns

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # a host-side group: ranks that wait while rank 0 drives all GPUs (the e2e leg) must not wait inside an NCCL
        # barrier -- its kernel would spin on the very GPUs rank 0 is using from another process
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")   # one node: never depend on the container hostname resolving
        cpu_group = dist.new_group(backend="gloo")
    stream = torch.cuda.current_stream()

    genes, cells, h1, desc = CONFIGS[args.config]
    genes, cells = args.genes or genes, args.cells or cells
    conf_mu = None
    if args.config == "c2":
        X, t = make_workload(genes, cells, seed=2 + 1000 * rank)
        G, N = X.shape
        zero_frac = float((X == 0).mean())
        design = A.DesignData(G, N, pseudotime=t, impulse_basis=ns(t, 4))
        annot = {"cell": np.array([f"cell_{j + 1}" for j in range(N)]), "continuous": t}
        api_kw = dict(strMuModel="impulse")
    else:
        from lineagepulse_b200.simulate import simulate_rows_device, workload_recipe
        t, groups, batch = cell_covariates(cells, seed=4)
        G, N = genes, cells
        if args.config == "c4":
            rec = workload_recipe(G, seed=4 + 1000 * rank, n_groups=4, n_batches=3)
            rows = simulate_rows_device(rec, N, 0, G, f"cuda:{local_rank}", groups, batch)
            design = A.DesignData(G, N, pseudotime=t, groups=groups, confounders_mu=[batch])
            annot = {"cell": np.array([f"cell_{j + 1}" for j in range(N)]), "continuous": t, "groups": groups, "batch": batch}
            api_kw = dict(strMuModel="groups", vecConfoundersMu=["batch"])
            conf_mu = 1
        else:
            rec = workload_recipe(G, seed=5 + 1000 * rank)
            rows = simulate_rows_device(rec, N, 0, G, f"cuda:{local_rank}")
            design = A.DesignData(G, N, pseudotime=t, spline_basis_mu=ns(t, 6))
            annot = {"cell": np.array([f"cell_{j + 1}" for j in range(N)]), "continuous": t}
            api_kw = dict(strMuModel="splines", scaDFSplinesMu=6)
        zero_frac = float((rows == 0).double().mean())
        X = rows.cpu().numpy()      # host copy: the e2e leg starts from host counts
        del rows
        torch.cuda.empty_cache()
    min_rowmean = float("nan")

    eng = Engine(local_rank)
    eng.set_stream(stream.cuda_stream)
    if world > 1:
        init_comm(eng)      # the library's own NCCL communicator (no Python in the fitPi loop)
    fp64_peak = eng.fp64_peak_tflops()
    eng.load_counts_i32(X)
    eng.set_design(design)

    # warm-up steps (untimed)
    if 0 < args.warmup_genes < G and args.config == "c2":
        Xw = np.ascontiguousarray(X[: args.warmup_genes])
        dw = A.DesignData(Xw.shape[0], N, pseudotime=t, impulse_basis=ns(t, 4))
        engw = Engine(local_rank)
        engw.set_stream(stream.cuda_stream)
        if world > 1:
            init_comm(engw)
        engw.load_counts_i32(Xw)
        engw.set_design(dw)
        for _ in range(args.warmup):
            native_pipeline(engw, dw, A, min_rowmean, h1)
        engw.close()
    else:
        for _ in range(args.warmup):
            native_pipeline(eng, design, A, min_rowmean, h1)

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
        res, work = native_pipeline(eng, design, A, min_rowmean, h1)
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
    # per-rank account of the last step: with independent matrices per rank the step is the slowest rank's, and the
    # lock-step collectives of fit A make every rank wait for the slowest one there too
    mine = {"step_ms": float(step_ms[-1]), "fit_BCD_ms": float(works[-1]["BCD"]["kernel_ms"]),
            "fit_A_ms": float(works[-1]["A"]["kernel_ms"])}
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine, group=cpu_group)
    else:
        per_rank = [mine]

    # ---- timed region 2: end to end through the public API with HOST buffers ("e2e") ----------
    # runLineagePulse(counts, dfAnnotation, ..., devices=[0 .. N-1]) from pinned host memory: ONE call of the
    # public API on the whole N-shard matrix (context creation, H2D of every shard's counts, packing, all fits with
    # the cross-shard sums over NCCL, D2H of every result).  Rank 0 makes the call through a multi-device context
    # (lpb200_create_multi) while the other ranks wait at the barrier with their GPUs idle; its matrix is the
    # concatenation of the ranks' shards.
    eng.close()
    del flush
    torch.cuda.empty_cache()
    if world > 1:
        torch.cuda.synchronize()
        parts = [None] * world if rank == 0 else None
        dist.gather_object(X, parts, dst=0, group=cpu_group)
        X_all = np.concatenate(parts, 0) if rank == 0 else None
        del parts
    else:
        X_all = X
    e2e_value = e2e_ms_step = None
    n_de = 0
    h2d = d2h = 0
    n_e2e = max(1, min(args.steps, 2))
    e2e_error = None
    if rank == 0:
        try:
            Gt = X_all.shape[0]
            pinned = torch.empty((Gt, N), dtype=torch.int32).pin_memory()
            pinned.numpy()[...] = X_all
            devices = list(range(world)) if world > 1 else None
            kw = dict(strDropModel="logistic", boolVerbose=False, device=local_rank, devices=devices, **api_kw)
            if args.warmup > 0:  # one untimed pass of the public API path (first-use allocations of fresh contexts)
                api.runLineagePulse(pinned.numpy(), annot, **kw).matCountsProc.engine().close()
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            e2e_ms = []
            obj = None
            for _ in range(n_e2e):
                if obj is not None:      # the previous result's GPU context is released outside the timed region
                    obj.matCountsProc.engine().close()
                    obj = None
                flush.fill_(1)
                for dv in range(world):
                    torch.cuda.synchronize(dv)
                # device-side timing, max over the devices of the call: events on every device's default stream
                # around the blocking call (its results are on the host when it returns)
                ev = []
                for dv in range(world):
                    with torch.cuda.device(dv):
                        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(torch.cuda.default_stream(dv))
                        ev.append((a, b_))
                obj = api.runLineagePulse(pinned.numpy(), annot, **kw)
                n_de = int((obj.dfResults["padj"] < 0.05).sum())
                for dv in range(world):
                    with torch.cuda.device(dv):
                        ev[dv][1].record(torch.cuda.default_stream(dv))
                for dv in range(world):
                    torch.cuda.synchronize(dv)
                e2e_ms.append(max(a.elapsed_time(b_) for a, b_ in ev))
            e2e_ms_step = float(np.mean(e2e_ms))
            e2e_value = Gt / (e2e_ms_step / 1e3)
            h2d = int(X_all.nbytes + 8 * N * 6 * world)  # counts + pseudotime + 4-column impulse basis + size factors, per device
            d2h = int(obj.dfResults.shape[0] * 14 * 8 + 8 * Gt * (7 + 1 + 1 + 1) * 2 + 8 * N)  # results table, parameter matrices, drop-out model
            del flush
        except Exception as ex:   # the other ranks wait at the host barrier below: report, do not hang them
            import traceback
            traceback.print_exc()
            e2e_error = f"{type(ex).__name__}: {ex}"
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)     # host-side wait (gloo): the other ranks' GPUs stay idle while rank 0 uses them
    barrier()

    # ---- the north_star matrix once: strong scaling ---------------------------------------------
    strong = None
    if not args.no_strong and args.config == "c2":
        torch.cuda.empty_cache()
        strong = strong_pass(args, A, torch, dist, rank, world, local_rank, fp64_peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fit_mudisp_kernel on fits B, C, D) --------------------
    w = works[-1]
    flop_B = w["B"]["ole"] * flops_per_ole(h1, True, zero_frac, conf_mu or 0)
    flop_C = w["C"]["ole"] * flops_per_ole("constant", False, zero_frac, conf_mu or 0)
    flop_D = w["D"]["ole"] * flops_per_ole(h1, False, zero_frac, conf_mu or 0)
    ker_s = w["BCD"]["kernel_ms"] / 1e3   # span of the three overlapping fit_mudisp launches
    achieved = (flop_B + flop_C + flop_D) / ker_s / 1e12
    share = ker_s * 1e3 / step_ms[-1]
    ct_bytes = 2 if int(X.max()) < 65535 else 4   # the library stores uint16 unless a count needs int32
    ct_name = "uint16_t" if ct_bytes == 2 else "int32_t"
    count_bytes = float(ct_bytes) * G * N   # gene-major count matrix, read once per fitMuDisp call
    perm_bytes = 4.0 * G * N               # per-gene cell permutation (non-zero cells first)
    roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                # dram__bytes_read + write of ONE launch (fit B) from the ncu --set full capture at exactly this shape and
                # count type (profiles/ncu_fit_mudisp_impulse_zinb_c2_r2_final.csv: 433.6 MB read + 22.1 MB written against
                # 400 MB of counts + permutation); any other shape: not captured
                "traffic": 455.7e6 if (G, N, ct_bytes, h1) == (5000, 10000, 4, "impulse") else None,
                "traffic_source": "ncu --set full of fit B at 5 000 x 10 000 int32 (profiles/ncu_fit_mudisp_impulse_zinb_c2_r2_final.csv); not re-measured per run",
                "kernel": f"fit_mudisp_kernel<{ct_name}> (fits B, C, D: three overlapping launches)",
                "kernel_share_of_step": share,
                "kernel_share_of_step_slowest_rank": max(r["fit_BCD_ms"] for r in per_rank) / max(r["step_ms"] for r in per_rank),
                "ole_per_s": (w["B"]["ole"] + w["C"]["ole"] + w["D"]["ole"]) / ker_s,
                "flops_per_ole": {"B": flops_per_ole(h1, True, zero_frac, conf_mu or 0), "C": flops_per_ole("constant", False, zero_frac, conf_mu or 0),
                                  "D": flops_per_ole(h1, False, zero_frac, conf_mu or 0)},
                "peak_source": "FP64 DFMA chain probe measured live in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "hbm": {"algorithmic_bytes_per_launch": count_bytes + perm_bytes,
                        "achieved_GBps": 2 * (count_bytes + perm_bytes) / ker_s / 1e9,
                        "peak_GBps": _hbm_peak(), "note": "count bytes are <0.1% of HBM bandwidth: FP64-bound (SURVEY.md F7)"}}

    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline and args.config == "c2":
        from tests.oracle_binding import Oracle
        cores = os.cpu_count() or 1
        oracle = Oracle(nthreads=cores)
        idx = cpu_sample_rows(G, args.cpu_sample_genes or 4 * cores)
        t0 = time.perf_counter()
        cpu_res = oracle_pipeline(oracle, np.ascontiguousarray(X[idx]), t, A, ns)
        dt = time.perf_counter() - t0
        cpu = {"value": len(idx) / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "genes_per_s_per_core": len(idx) / dt / cores,
               "sample": f"{len(idx)} evenly spaced genes (4 per core) x {N} cells of the workload, full pipeline (fits A-D + LRT), "
                         f"{dt:.1f} s; CPU restatement of the reference algorithm (R not installable here)"}
        parity = parity_block(res, X, t, idx, oracle, A, ns, cpu_res)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{desc}: {G} genes x {N} cells per GPU, fits A-D + LRT (simulateContinuousDataSet recipe "
                                   "50/25/25 const/lin/impulse, mumax 100, sd 3, drop-out 0.1)",
                       "genes_total": int(g_all), "cells": N, "zero_fraction": zero_frac,
                       "l2": "flushed between steps (256 MiB write)", "parallelism": f"gene-sharded x{world}",
                       "de_genes_q05": n_de},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_step, "error": e2e_error,
                    "api": "lineagepulse_b200.runLineagePulse(host counts" + (f", devices=[0..{world - 1}]): one sharded call" if world > 1 else ")")},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "strong": strong,
            "ranks": {k: [round(r[k], 2) for r in per_rank] for k in ("step_ms", "fit_BCD_ms", "fit_A_ms")},
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
