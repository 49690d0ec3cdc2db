"""Generates tests/golden/reference_exec.json by EXECUTING the reference's own R source text (run in the build container,
where /root/reference exists; the tests and the GPU box only read the JSON):

    python tests/golden/make_reference_exec_golden.py [/root/reference]

R is not installed here, so the text is run by tests/r_eval.py (an evaluator for the subset of R the hot path is written
in, on the syntax tree of tests/r_parse.py).  What runs is the reference's code, unmodified and in full: fitModel's EM
loop, initMuModel / initDispModel / initDropModel, fitMuDisp / fitMuDispGene / fitMuDispGeneImpulse with
initialiseImpulseParameters' three starts, fitPi / fitPi_SingleCell / fitPi_ManyCells, evalLogLikMatrix, the decompress*
functions, evalLogLikMuDispGeneFit with its clamps, evalLogLikZINB / NB.  Only what R itself implements in C is restated
(tests/r_builtins.py): IEEE arithmetic through the C library, long-double sum(), optim(method = "BFGS") = vmmin with
central differences (checked against ?optim's Rosenbrock answers), splines::ns, lm, and dnbinom -- 50-digit mpmath for
the objective-level vectors; for whole fits (millions of terms) the oracle's restated nmath dnbinom, which is pinned
against mpmath and scipy on its own (tests/test_oracle_pins.py).

The JSON holds, per case, the inputs (counts, pseudotime, groups, batches, size factors, model names) and what the
reference's code returned (parameter matrices, per-gene log-likelihoods and convergence codes, the EM trajectory).
tests/test_reference_exec.py runs the CPU oracle on the same inputs and compares."""
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.r_eval import Interp, RList, RVec, as_py  # noqa: E402


def S(s):
    return RVec(np.array([s] if isinstance(s, str) else list(s), dtype=object))


def D(x):
    return RVec(np.atleast_1d(np.asarray(x, dtype=np.float64)))


def B(x):
    return RVec(np.array([bool(x)]))


class Reference:
    """The reference's R functions, loaded from its source files into the evaluator."""

    def __init__(self, root, dnbinom=None):
        self.I = Interp()
        if dnbinom is not None:
            self.I.dnbinom = dnbinom
        files = sorted(glob.glob(os.path.join(root, "R", "*.R")))
        assert len(files) == 23, files
        for f in files:
            self.I.load(open(f).read())

    def call(self, name, **kw):
        g = self.I.globals
        return self.I.apply(self.I.lookup_function(name, g), list(kw.items()), g)


def lm_in_the_oracle_s_order(X, y):
    """Least squares y ~ 0 + X by twice-iterated Gram-Schmidt with long-double dot products: the operation order of the
    oracle's solver (oracle/lp_oracle.c thinqr_build / thinqr_solve).  Any two correct solvers differ in the last bit of
    the fitted values; the impulse start values inherit that bit and BFGS amplifies it, so the
    impulse cases are recorded twice: with numpy's QR (independent solver, compared statistically) and with this order
    (every gene must then agree with the oracle to the last bit: control flow, clamps, start selection)."""
    n, k = X.shape
    Q = np.array(X, dtype=np.float64, order="F")
    R = np.zeros((k, k))

    def ldot(a, b):
        return np.cumsum(a.astype(np.longdouble) * b.astype(np.longdouble))[-1]

    for c in range(k):
        for _ in range(2):
            for p in range(c):
                dot = float(ldot(Q[:, p], Q[:, c]))
                Q[:, c] = Q[:, c] - dot * Q[:, p]
                R[p, c] += dot
        nr = float(np.sqrt(ldot(Q[:, c], Q[:, c])))
        R[c, c] = nr
        Q[:, c] = Q[:, c] / nr
    z = np.array([float(ldot(Q[:, c], y)) for c in range(k)])
    fitted = np.zeros(n)
    for c in range(k):
        fitted = fitted + Q[:, c] * z[c]
    coef = np.zeros(k)
    for c in range(k - 1, -1, -1):
        s_ = z[c]
        for p in range(c + 1, k):
            s_ -= R[c, p] * coef[p]
        coef[c] = s_ / R[c, c]
    return coef, fitted


def make_data(seed, G=9, N=26, size_factors=False):
    """A tiny simulateContinuousDataSet-style matrix with group and batch annotation (plain numpy; any counts would do)."""
    rng = np.random.default_rng(seed)
    t = np.round(np.sort(rng.uniform(0, 1, N)), 3)
    t[0], t[-1] = 0.0, 1.0
    groups = np.array(["early", "mid", "late"])[np.minimum((t * 3).astype(int), 2)]
    batch = rng.permutation(np.array(["b1", "b2", "b3"])[np.arange(N) % 3])
    sf = np.round(rng.uniform(0.7, 1.4, N), 3) if size_factors else np.ones(N)
    X = np.zeros((G, N), dtype=np.int64)
    bf = {"b1": 1.0, "b2": 1.5, "b3": 0.7}
    for i in range(G):
        h0, h1, h2 = rng.uniform(5, 80, 3)
        kind = i % 3
        if kind == 0:
            mu = np.full(N, h0)
        elif kind == 1:
            mu = h0 + (h1 - h0) * t
        else:
            mu = (h0 + (h1 - h0) / (1 + np.exp(-12 * (t - 0.3)))) * (h2 + (h1 - h2) / (1 + np.exp(12 * (t - 0.7)))) / h1
        mu = mu * sf * np.array([bf[b] for b in batch])
        phi = max(3 + rng.normal(0, 0.5), 0.05)
        x = rng.negative_binomial(phi, phi / (phi + mu))
        x[rng.uniform(size=N) < 0.12] = 0
        X[i] = x
    X[0, :3] = [0, 1, 2]
    return dict(counts=X, t=t, groups=groups, batch=batch, sf=sf)


def data_json(data):
    """The inputs as stored in the JSON, with the natural-spline bases the run used (splines::ns restated in
    lineagepulse_b200/splines.py goes through LAPACK's QR, whose last bits depend on the BLAS kernels of the machine: the
    tests feed the oracle these stored bases, so the stored outputs do not depend on the BLAS of the test machine)."""
    from lineagepulse_b200.splines import ns
    out = {k: np.asarray(v).tolist() for k, v in data.items()}
    out["ns"] = {str(df): ns(np.asarray(data["t"], dtype=np.float64), df).tolist() for df in (3, 4)}
    return out


def r_inputs(data):
    X = data["counts"].astype(np.float64)
    G, N = X.shape
    cells = [f"cell_{j + 1}" for j in range(N)]
    genes = [f"gene_{i + 1}" for i in range(G)]
    mat = RVec(X.reshape(-1, order="F"), dim=(G, N), dimnames=(genes, cells))
    df = RList([S(cells), D(data["t"]), S(list(data["groups"])), S(list(data["batch"]))], ["cell", "continuous", "groups", "batch"])
    sf = RVec(np.asarray(data["sf"], dtype=np.float64), names=cells)
    return mat, df, sf


def fit_model(ref, data, strMuModel, strDispModel, strDropModel, strDropFitGroup="PerCell", lsDropModel=None,
              vecConfoundersMu=None, vecConfoundersDisp=None, scaDFSplinesMu=None, scaDFSplinesDisp=None,
              matPiConstPredictors=None, cycles=20):
    mat, df, sf = r_inputs(data)
    t0 = time.time()
    n0 = ref.I.counters["optim_fn"]
    res = ref.call("fitModel", matCounts=mat, dfAnnotation=df, vecConfoundersMu=None if vecConfoundersMu is None else S(vecConfoundersMu),
                   vecConfoundersDisp=None if vecConfoundersDisp is None else S(vecConfoundersDisp), vecNormConst=sf,
                   scaDFSplinesMu=None if scaDFSplinesMu is None else D(scaDFSplinesMu),
                   scaDFSplinesDisp=None if scaDFSplinesDisp is None else D(scaDFSplinesDisp), matWeights=None,
                   matPiConstPredictors=matPiConstPredictors, lsDropModel=lsDropModel, strMuModel=S(strMuModel),
                   strDispModel=S(strDispModel), strDropModel=S(strDropModel), strDropFitGroup=S(strDropFitGroup),
                   scaMaxEstimationCycles=D(cycles), boolVerbose=B(False), boolSuperVerbose=B(False))
    ll = ref.call("evalLogLikMatrix", matCounts=mat, lsMuModel=res.get("lsMuModel"), lsDispModel=res.get("lsDispModel"),
                  lsDropModel=res.get("lsDropModel"))
    mu, disp, drop = res.get("lsMuModel"), res.get("lsDispModel"), res.get("lsDropModel")
    out = dict(matMuModel=as_py(mu.get("matMuModel")), matDispModel=as_py(disp.get("matDispModel")),
               lsmatBatchModelMu=[as_py(m) for m in mu.get("lsmatBatchModel").v] if vecConfoundersMu else None,
               lsmatBatchModelDisp=[as_py(m) for m in disp.get("lsmatBatchModel").v] if vecConfoundersDisp else None,
               matDropoutLinModel=as_py(drop.get("matDropoutLinModel")) if drop.get("matDropoutLinModel") is not None else None,
               vecLL=as_py(ll), vecEMLogLikModel=[v for v in as_py(res.get("vecEMLogLikModel")) if v == v],
               boolConvergenceModel=bool(res.get("boolConvergenceModel").v[0]), strReport=res.get("strReport").v[0],
               scaDegFreedomMu=float(mu.get("lsMuModelGlobal").get("scaDegFreedom").v[0]),
               scaDegFreedomDisp=float(disp.get("lsDispModelGlobal").get("scaDegFreedom").v[0]),
               objective_evaluations=ref.I.counters["optim_fn"] - n0, seconds=round(time.time() - t0, 1))
    return res, out


def run_lineagepulse(ref, data, **kw):
    """The public call, main_LineagePulse.R:219-328: processSCData (checks, sparse conversion, all-zero filters),
    calcNormConst, fitLPModels (fits A-D), runDEAnalysis (LRT, BH, results table with its NA rows)."""
    mat, df, sf = r_inputs(data)
    df.attrs = {"class": "data.frame", "row.names": list(df.v[0].v)}
    args = dict(counts=mat, dfAnnotation=df, boolVerbose=B(False))
    for k, v in kw.items():
        args[k] = S(v) if isinstance(v, (str, list)) else (B(v) if isinstance(v, bool) else D(v))
    if not np.all(np.asarray(data["sf"]) == 1.0):
        args["vecNormConstExternal"] = sf
    t0 = time.time()
    obj = ref.call("runLineagePulse", **args)
    slot = lambda name: obj.v[obj.names.index(name)]     # noqa: E731
    res = slot("dfResults")
    conv = slot("lsFitConvergence")
    table = {}
    for name, col in zip(res.names, res.v):
        table[name] = [None if (isinstance(x, float) and x != x) else (x.item() if hasattr(x, "item") else x) for x in col.v]
    return dict(dfResults=table, row_names=res.attrs["row.names"],
                vecEMLogLikH0=[v for v in as_py(conv.get("vecEMLogLikH0")) if v == v],
                vecEMLogLikH1=[v for v in as_py(conv.get("vecEMLogLikH1")) if v == v],
                genes_kept=list(slot("matCountsProc").dimnames[0]), cells_kept=list(slot("matCountsProc").dimnames[1]),
                strReport=slot("strReport").v[0], seconds=round(time.time() - t0, 1))


def objective_vectors(ref, data, res, rng, n_random=6):
    """evalLogLikMuDispGeneFit (fitMeanDispersion.R:62-222) of every gene at random and at extreme parameter vectors,
    called exactly as fitMuDispGene calls it (:367-383), dnbinom by mpmath."""
    mat, df, sf = r_inputs(data)
    mu, disp, drop = res.get("lsMuModel"), res.get("lsDispModel"), res.get("lsDropModel")
    gm, gd, gp = mu.get("lsMuModelGlobal"), disp.get("lsDispModelGlobal"), drop.get("lsDropModelGlobal")
    G, N = data["counts"].shape
    n = int(gd.get("scaDegFreedom").v[0] + gm.get("scaDegFreedom").v[0])
    out = []
    for i in range(G):
        x = data["counts"][i]
        thetas = [rng.normal(0, 1.5, n) for _ in range(n_random)]
        thetas += [np.full(n, -30.0), np.full(n, 30.0), rng.choice([-25.0, 0.5, 25.0], n)]   # clamps
        pred = None if drop.get("matPiConstPredictors") is None else RVec(drop.get("matPiConstPredictors").mat()[i].copy())
        strDrop = gp.get("strDropModel").v[0]
        pi_param = None
        if strDrop == "logistic":      # fitMeanDispersion.R:887-892: the drop-out rates are fixed during the gene's fit
            pi_param = ref.call("decompressDropoutRateByGene", matDropModel=drop.get("matDropoutLinModel"), vecMu=None,
                                vecPiConstPredictors=pred, lsDropModelGlobal=gp)
        vals = []
        for th in thetas:
            v = ref.call("evalLogLikMuDispGeneFit", vecTheta=D(th), vecCounts=D(x), lsMuModelGlobal=gm, lsDispModelGlobal=gd,
                         vecTimepoints=gm.get("vecTimepoints"), vecindTimepointAssign=gm.get("vecindTimepointAssign"),
                         matDropoutLinModel=drop.get("matDropoutLinModel"), vecPiConstPredictors=pred, lsDropModelGlobal=gp,
                         vecPiParam=pi_param, vecidxNotZero=RVec(np.nonzero(x > 0)[0] + 1), vecidxZero=RVec(np.nonzero(x == 0)[0] + 1))
            vals.append(float(v.v[0]))
        out.append(dict(gene=i, thetas=[t.tolist() for t in thetas], values=vals))
    return out


def reexecute(case, cases, root="/root/reference"):
    """Runs ONE fit case of the JSON again, now, on this machine: the reference's fitModel from its source text on the
    stored inputs (a case that was fitted given another case's drop-out model -- `given` -- runs that case first).  For
    tests/test_reference_exec.py on a machine whose C library rounds exp / log differently from the one the JSON was made
    on: the stored bits are then not this machine's, the comparison to the last bit is made against this run instead."""
    from tests.oracle_binding import Oracle
    orc = Oracle()
    ref = Reference(root, dnbinom=lambda x, s, m: orc.dnbinom_mu_log(x, s, m))
    ref.I.lm_solver = lm_in_the_oracle_s_order

    def run(c):
        d = c["data"]
        data = dict(counts=np.asarray(d["counts"], dtype=np.int64), t=np.asarray(d["t"], dtype=np.float64),
                    groups=np.asarray(d["groups"]), batch=np.asarray(d["batch"]), sf=np.asarray(d["sf"], dtype=np.float64))
        how = {k: v for k, v in c["how"].items() if k != "matDropoutLinModel"}
        drop = None
        if c.get("given"):    # that case's lsDropModel list, holding the STORED matrix (what the oracle is handed, too)
            drop = run(next(x for x in cases if x["name"] == c["given"]))[0].get("lsDropModel")
            m = drop.get("matDropoutLinModel")
            m.v = np.asarray(c["how"]["matDropoutLinModel"], dtype=np.float64).reshape(m.dim).reshape(-1, order="F")
        pred = None
        if "pi_const_pred" in d:
            pm = np.asarray(d["pi_const_pred"], dtype=np.float64)
            pred = RVec(pm.reshape(-1, order="F"), dim=pm.shape,
                        dimnames=([f"gene_{i + 1}" for i in range(pm.shape[0])], [f"p{k + 1}" for k in range(pm.shape[1])]))
        return fit_model(ref, data, lsDropModel=drop, matPiConstPredictors=pred, **how)

    return run(case)[1]


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    from tests.oracle_binding import Oracle
    orc = Oracle()
    fast = Reference(root, dnbinom=lambda x, s, m: orc.dnbinom_mu_log(x, s, m))
    exact = Reference(root)                      # dnbinom by mpmath
    rng = np.random.default_rng(2024)
    cases = []

    def add(name, data, how, res_out, objective=None, given=None):
        res, out = res_out
        case = dict(name=name, how=how, data=data_json(data), reference=out)
        if given is not None:
            case["given"] = given         # the case whose fitted drop-out model this fit was handed (lsDropModel)
        if objective is not None:
            case["objective"] = objective
        cases.append(case)
        print(f"{name}: {out['seconds']} s, {out['objective_evaluations']} objective evaluations, EM {out['vecEMLogLikModel']}", flush=True)
        return res

    d1 = make_data(11)
    how = dict(strMuModel="constant", strDispModel="constant", strDropModel="logistic", strDropFitGroup="PerCell")
    r = fit_model(fast, d1, **how)
    fitA = add("H0 constant ZINB, drop-out EM (fit A)", d1, how, r, objective_vectors(exact, d1, r[0], rng))
    r_exact = fit_model(exact, d1, **how)
    add("H0 constant ZINB, drop-out EM (fit A), dnbinom by mpmath", d1, how, r_exact)
    dropA = fitA.get("lsDropModel")
    dropA_py = as_py(dropA.get("matDropoutLinModel"))

    numpy_lm = fast.I.lm_solver
    for solver, tag in ((numpy_lm, "lm by numpy QR"), (lm_in_the_oracle_s_order, "lm in the oracle's operation order")):
        fast.I.lm_solver = solver
        how = dict(strMuModel="impulse", strDispModel="constant", strDropModel="logistic")
        r = fit_model(fast, d1, lsDropModel=dropA, **how)
        add(f"H1 impulse ZINB given fit A's drop-out model (fit B), {tag}", d1, dict(how, matDropoutLinModel=dropA_py), r,
            objective_vectors(exact, d1, r[0], rng, n_random=3) if solver is numpy_lm else None,
            given="H0 constant ZINB, drop-out EM (fit A)")
        how = dict(strMuModel="impulse", strDispModel="constant", strDropModel="none")
        r = fit_model(fast, d1, **how)
        add(f"impulse NB (fit D), {tag}", d1, how, r)
    # from here on lm() (the spline models' initial values) runs in the oracle's operation order: these models are
    # well conditioned, the fits are compared to the last bit
    how = dict(strMuModel="constant", strDispModel="constant", strDropModel="none")
    r = fit_model(fast, d1, **how)
    add("constant NB (fit C)", d1, how, r, objective_vectors(exact, d1, r[0], rng, n_random=3))
    starts = []
    fast.I.lm_solver = numpy_lm
    mat1, df1, sf1 = r_inputs(d1)
    lsMu = fast.call("initMuModel", matCounts=mat1, dfAnnotation=df1, vecConfoundersMu=None, vecNormConst=sf1, scaDFSplinesMu=None,
                     matWeights=None, matMuModelInit=None, lsmatBatchModelInitMu=None, strMuModel=S("impulse"),
                     MAXIT_BFGS_MuDisp=D(1000), RELTOL_BFGS_MuDisp=D(1e-8))
    for i in range(d1["counts"].shape[0]):
        st = fast.call("initialiseImpulseParameters", vecCounts=D(d1["counts"][i]), lsMuModelGlobal=lsMu.get("lsMuModelGlobal"))
        starts.append(dict(peak=as_py(st.get("peak")), valley=as_py(st.get("valley"))))
    cases.append(dict(name="initialiseImpulseParameters (lm by numpy QR)", how=dict(strMuModel="impulse", strDispModel="constant",
                                                                                      strDropModel="none"),
                      data=data_json(d1), starts=starts))
    fast.I.lm_solver = lm_in_the_oracle_s_order

    d2 = make_data(12, size_factors=True)
    how = dict(strMuModel="constant", strDispModel="constant", strDropModel="logistic", strDropFitGroup="PerCell")
    r = fit_model(fast, d2, **how)
    fitA2 = add("H0 constant ZINB with size factors, drop-out EM", d2, how, r)
    drop2 = fitA2.get("lsDropModel")
    drop2_py = as_py(drop2.get("matDropoutLinModel"))
    how = dict(strMuModel="groups", strDispModel="constant", strDropModel="logistic", vecConfoundersMu=["batch"])
    r = fit_model(fast, d2, lsDropModel=drop2, **how)
    add("groups + batch confounder ZINB (C4 shape), size factors", d2, dict(how, matDropoutLinModel=drop2_py), r,
        objective_vectors(exact, d2, r[0], rng, n_random=3), given="H0 constant ZINB with size factors, drop-out EM")
    how = dict(strMuModel="splines", strDispModel="constant", strDropModel="logistic", scaDFSplinesMu=4)
    r = fit_model(fast, d2, lsDropModel=drop2, **how)
    add("splines (df 4) ZINB (C5 shape), size factors", d2, dict(how, matDropoutLinModel=drop2_py), r,
        objective_vectors(exact, d2, r[0], rng, n_random=3), given="H0 constant ZINB with size factors, drop-out EM")
    how = dict(strMuModel="constant", strDispModel="groups", strDropModel="none", vecConfoundersDisp=["batch"])
    r = fit_model(fast, d2, **how)
    add("dispersion groups + dispersion batch confounder, NB", d2, how, r, objective_vectors(exact, d2, r[0], rng, n_random=3))
    how = dict(strMuModel="groups", strDispModel="splines", strDropModel="none", scaDFSplinesDisp=3)
    r = fit_model(fast, d2, **how)
    add("groups mean, dispersion splines (df 3), NB", d2, how, r, objective_vectors(exact, d2, r[0], rng, n_random=3))

    d3 = make_data(13, G=8, N=22)
    how = dict(strMuModel="constant", strDispModel="constant", strDropModel="logistic_ofMu", strDropFitGroup="PerCell")
    r = fit_model(fast, d3, **how)
    add("H0 constant ZINB, logistic_ofMu drop-out EM", d3, how, r, objective_vectors(exact, d3, r[0], rng, n_random=3))
    how = dict(strMuModel="constant", strDispModel="constant", strDropModel="logistic", strDropFitGroup="ForAllCells")
    r = fit_model(fast, d3, **how)
    add("H0 constant ZINB, ForAllCells drop-out EM", d3, how, r)

    # gene-wise drop-out predictors (matPiConstPredictors: decompressDropoutRateByCell / ByGene with constant predictors,
    # fitPi's parameter layout with them, fitDropout.R:59-140)
    pred = np.round(np.random.default_rng(5).normal(0, 1, (d3["counts"].shape[0], 2)), 3)
    genes3 = [f"gene_{i + 1}" for i in range(pred.shape[0])]
    for drop, grp in (("logistic", "PerCell"), ("logistic_ofMu", "PerCell"), ("logistic_ofMu", "ForAllCells")):
        how = dict(strMuModel="constant", strDispModel="constant", strDropModel=drop, strDropFitGroup=grp)
        r = fit_model(fast, d3, matPiConstPredictors=RVec(pred.reshape(-1, order="F"), dim=pred.shape, dimnames=(genes3, ["p1", "p2"])), **how)
        add(f"H0 constant ZINB, {drop} drop-out EM ({grp}) with two gene-wise predictors", dict(d3, pi_const_pred=pred), how, r)

    # H1 fits GIVEN a logistic_ofMu drop-out model: the drop-out rate of every observation moves with the gene's mean
    # parameters inside the fit (fitMeanDispersion.R:887-892 and evalLogLikMuDispGeneFit's vecPiParam path)
    how = dict(strMuModel="constant", strDispModel="constant", strDropModel="logistic_ofMu")
    fitA3 = fit_model(fast, d3, **how)[0]
    drop3 = fitA3.get("lsDropModel")
    drop3_py = as_py(drop3.get("matDropoutLinModel"))
    for how in (dict(strMuModel="groups", strDispModel="constant", strDropModel="logistic_ofMu"),
                dict(strMuModel="splines", strDispModel="groups", strDropModel="logistic_ofMu", scaDFSplinesMu=3),
                dict(strMuModel="groups", strDispModel="splines", strDropModel="logistic_ofMu", scaDFSplinesDisp=3, vecConfoundersMu=["batch"])):
        r = fit_model(fast, d3, lsDropModel=drop3, **how)
        add(f"{how['strMuModel']} mean, {how['strDispModel']} dispersion ZINB given a logistic_ofMu drop-out model", d3,
            dict(how, matDropoutLinModel=drop3_py), r, given="H0 constant ZINB, logistic_ofMu drop-out EM")

    # two confounders each for the mean and the dispersion (lsmatBatchModel with two matrices; the annotation's group column
    # serves as the second confounder): order of the batch parameters in the parameter vector, first level fixed at 1
    how = dict(strMuModel="splines", strDispModel="constant", strDropModel="none", scaDFSplinesMu=3, vecConfoundersMu=["batch", "groups"])
    r = fit_model(fast, d2, **how)
    add("splines (df 3) with two mean confounders, NB, size factors", d2, how, r)
    how = dict(strMuModel="constant", strDispModel="constant", strDropModel="none", vecConfoundersMu=["batch", "groups"],
               vecConfoundersDisp=["groups", "batch"])
    r = fit_model(fast, d2, **how)
    add("constant with two mean and two dispersion confounders, NB, size factors", d2, how, r)

    # the public API end to end
    d4 = make_data(21, G=8, N=24, size_factors=True)
    d4["counts"][3, :] = 0          # an all-zero gene and an all-zero cell: filtered, NA row in dfResults
    d4["counts"][:, 5] = 0
    for name, kw in (("splines (df 4) vs constant", dict(strMuModel="splines", scaDFSplinesMu=4, strDropModel="logistic")),
                     ("impulse vs constant", dict(strMuModel="impulse", strDropModel="logistic")),
                     ("groups + batch confounder vs constant + batch", dict(strMuModel="groups", vecConfoundersMu=["batch"],
                                                                            strDropModel="logistic")),
                     ("groups vs constant, noise model on H1", dict(strMuModel="groups", strDropModel="logistic",
                                                                   boolEstimateNoiseBasedOnH0=False)),
                     ("splines with dispersion splines under H1", dict(strMuModel="splines", scaDFSplinesMu=4, strDispModelFull="splines",
                                                                       scaDFSplinesDisp=3, strDropModel="logistic"))):
        out = run_lineagepulse(fast, d4, **kw)
        cases.append(dict(name="runLineagePulse: " + name, how=kw, data=data_json(d4), api=out))
        print("runLineagePulse:", name, out["seconds"], "s; p =", out["dfResults"]["p"][:3], flush=True)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_exec.json")
    json.dump(dict(note="outputs of the reference's own R source executed by tests/r_eval.py; see make_reference_exec_golden.py",
                   cases=cases), open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
