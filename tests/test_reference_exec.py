"""The CPU oracle against outputs of the reference's OWN R source (CPU-only).

tests/golden/reference_exec.json was made by tests/golden/make_reference_exec_golden.py: the reference's R files, parsed by
tests/r_parse.py and executed by tests/r_eval.py -- fitModel's EM loop, the init*Model functions, fitMuDisp /
fitMuDispGene / fitMuDispGeneImpulse, fitPi, evalLogLikMatrix, the decompress* functions and evalLogLikMuDispGeneFit run
as written; only R's compiled builtins (arithmetic, sum, optim's vmmin, dnbinom, ns, lm) are restated.  This pins what a
restatement gets wrong: parameter unpacking order, clamps, start values, the three-start selection, the EM stopping rule,
which drop-out rates are held fixed during a gene's fit."""
import glob
import os
import warnings

import numpy as np
import pytest

from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.splines import ns
from tests.helpers import golden, model, rel

CASES = golden("reference_exec.json")["cases"]


def _design(case):
    d, how = case["data"], case["how"]
    X = np.ascontiguousarray(d["counts"], dtype=np.int32)
    G, N = X.shape
    t = np.asarray(d["t"], dtype=np.float64)
    sf = np.asarray(d["sf"], dtype=np.float64)
    groups = d["groups"] if "groups" in (how["strMuModel"], how["strDispModel"]) else None

    def basis(df):      # the basis the reference's run used (stored: ns() goes through LAPACK, machine-dependent last bits)
        return np.asarray(d["ns"][str(int(df))], dtype=np.float64) if "ns" in d else ns(t, int(df))

    design = A.DesignData(
        G, N, size_factors=None if np.all(sf == 1.0) else sf, pseudotime=t, groups=groups,
        spline_basis_mu=basis(how["scaDFSplinesMu"]) if how["strMuModel"] == "splines" else None,
        spline_basis_disp=basis(how["scaDFSplinesDisp"]) if how["strDispModel"] == "splines" else None,
        impulse_basis=basis(4), confounders_mu=[d[c] for c in how["vecConfoundersMu"]] if how.get("vecConfoundersMu") else None,
        confounders_disp=[d[c] for c in how["vecConfoundersDisp"]] if how.get("vecConfoundersDisp") else None,
        pi_const_pred=np.asarray(d["pi_const_pred"], dtype=np.float64) if "pi_const_pred" in d else None)
    m = model(how["strMuModel"], how["strDispModel"], how["strDropModel"], how.get("strDropFitGroup", "PerCell"))
    return X, design, m


def _fit(oracle, case):
    X, design, m = _design(case)
    how = case["how"]
    fit_drop = how["strDropModel"] != "none" and "matDropoutLinModel" not in how
    p = A.ParamsData(design, m)
    if "matDropoutLinModel" in how:
        p.mat_drop = np.asfortranarray(np.asarray(how["matDropoutLinModel"], dtype=np.float64))
    r = oracle.fit_model(X, design, m, p, fit_drop)
    return X, design, m, p, r


FITS = [c for c in CASES if "reference" in c]


def _reexecuted(case, root="/root/reference"):
    """The reference's fitModel executed again on this machine for one stored case (tests/golden/make_reference_exec_golden.py
    reexecute); None where the reference's source is absent."""
    if not os.path.isdir(os.path.join(root, "R")):
        return None
    from tests.golden.make_reference_exec_golden import reexecute
    return reexecute(case, CASES, root)


@pytest.mark.parametrize("case", FITS, ids=[c["name"] for c in FITS])
def test_oracle_reproduces_the_executed_reference(oracle, case):
    X, design, m, p, r = _fit(oracle, case)
    ref = case["reference"]
    impulse = case["how"]["strMuModel"] == "impulse"
    mpmath_run = "mpmath" in case["name"]
    em = np.asarray(ref["vecEMLogLikModel"])
    ll = oracle.eval_loglik(X, design, m, p)
    ll_ref = np.asarray(ref["vecLL"])
    if impulse and "numpy QR" in case["name"]:
        # The start values come from an lm() fit; two correct least-squares solvers differ in its last bit (4e-16 here,
        # test_impulse_start_values_of_the_executed_reference) and BFGS from three starts over a 9-parameter surface on
        # 26 cells amplifies that bit (tests/test_oracle_sensitivity.py).  With the solver's operation order made equal
        # (the other impulse cases) every gene agrees to the last bit, so this spread is conditioning, not a difference
        # of algorithm: about half the genes land on the same optimum, the rest on neighbouring ones.
        assert r["n_iter"] == len(em) and bool(r["converged"]) == ref["boolConvergenceModel"]
        assert np.mean(rel(ll, ll_ref) < 1e-6) >= 0.3 and np.max(rel(ll, ll_ref)) < 5e-3
        assert rel(r["em_ll"][0], em[0]) < 1e-3
        return
    if not mpmath_run:
        # same libm, same dnbinom, same long-double sums, same vmmin: the reference's R text and the oracle take every
        # optimiser decision alike -- log-likelihoods, parameters and the EM trajectory are equal to the last bit
        def differences(ref):
            em = np.asarray(ref["vecEMLogLikModel"])
            pairs = [("n_iter", r["n_iter"], len(em)), ("converged", bool(r["converged"]), ref["boolConvergenceModel"]),
                     ("vecLL", ll, np.asarray(ref["vecLL"])), ("vecEMLogLikModel", r["em_ll"][: len(em)], em),
                     ("matMuModel", p.mat_mu, np.asarray(ref["matMuModel"], dtype=np.float64).reshape(p.mat_mu.shape)),
                     ("matDispModel", p.mat_disp, np.asarray(ref["matDispModel"], dtype=np.float64).reshape(p.mat_disp.shape))]
            for k, bm in enumerate(ref["lsmatBatchModelMu"] or []):
                pairs.append((f"lsmatBatchModelMu[{k}]", p.batch_mu[k], np.asarray(bm)))
            for k, bm in enumerate(ref["lsmatBatchModelDisp"] or []):
                pairs.append((f"lsmatBatchModelDisp[{k}]", p.batch_disp[k], np.asarray(bm)))
            if ref["matDropoutLinModel"] is not None and "matDropoutLinModel" not in case["how"]:
                pairs.append(("matDropoutLinModel", p.mat_drop,
                              np.asarray(ref["matDropoutLinModel"], dtype=np.float64).reshape(p.mat_drop.shape)))
            return [name for name, got, want in pairs if not np.array_equal(got, want)]

        if differences(ref):
            # "Same libm" holds on ONE machine.  The stored bits are those of the machine the JSON was made on; the C
            # library picks its exp / log by CPU feature (glibc ifunc: an FMA and a non-FMA variant that round a few
            # arguments per million differently), and one such bit moves an ill-conditioned fit to a neighbouring optimum.
            # So on a machine that does not reproduce the stored bits the reference's text is executed again, here and now,
            # for this case, and the comparison to the last bit is made against that run.
            again = _reexecuted(case)
            if again is None:   # no /root/reference on this machine: what remains checkable is the north_star tolerance
                warnings.warn(f"{case['name']}: stored reference bits are another machine's and /root/reference is absent")
                assert np.mean(rel(ll, ll_ref) < 1e-6) >= (0.3 if impulse else 1.0) and np.max(rel(ll, ll_ref)) < 5e-3
                return
            assert not differences(again), (differences(again), np.max(rel(ll, np.asarray(again["vecLL"]))))
        return
    assert r["n_iter"] == len(em) and bool(r["converged"]) == ref["boolConvergenceModel"]
    tol_ll, tol_par = 1e-8, 1e-5        # dnbinom by mpmath instead of the restated nmath routine: ~1e-16 per term
    assert np.max(rel(r["em_ll"][: len(em)], em)) < tol_ll
    assert np.max(rel(ll, ll_ref)) < tol_ll
    mu_ref = np.asarray(ref["matMuModel"], dtype=np.float64).reshape(p.mat_mu.shape)
    disp_ref = np.asarray(ref["matDispModel"], dtype=np.float64).reshape(p.mat_disp.shape)
    ok = disp_ref.max(axis=1) < 100          # a dispersion that ran off to the Poisson limit is flat, not determined
    if case["how"]["strMuModel"] == "splines":
        assert np.max(np.abs(p.mat_mu - mu_ref)) < 1e-4
    else:
        assert np.max(rel(p.mat_mu, mu_ref)) < tol_par * 10
    if case["how"]["strDispModel"] != "splines":
        assert np.max(rel(p.mat_disp[ok], disp_ref[ok])) < tol_par * 100
    if ref["lsmatBatchModelMu"]:
        assert np.max(rel(p.batch_mu[0], np.asarray(ref["lsmatBatchModelMu"][0]))) < 1e-4
    if ref["lsmatBatchModelDisp"]:
        assert np.max(rel(p.batch_disp[0][ok], np.asarray(ref["lsmatBatchModelDisp"][0])[ok])) < 1e-3
    if ref["matDropoutLinModel"] is not None and "matDropoutLinModel" not in case["how"]:
        drop_ref = np.asarray(ref["matDropoutLinModel"], dtype=np.float64).reshape(p.mat_drop.shape)
        assert np.max(np.abs(p.mat_drop - drop_ref) / np.maximum(1.0, np.abs(drop_ref))) < 1e-4


@pytest.mark.parametrize("case", [c for c in CASES if "objective" in c], ids=[c["name"] for c in CASES if "objective" in c])
def test_objective_equals_the_executed_reference(oracle, case):
    """evalLogLikMuDispGeneFit of the reference (executed, dnbinom by 50-digit mpmath) at random, clamped and extreme
    parameter vectors against the oracle's objective: unpacking order, every clamp, the -700 floor, which drop-out rates
    enter."""
    X, design, m = _design(case)
    how, ref = case["how"], case["reference"]
    p = A.ParamsData(design, m)
    if m.drop_model != A.DROP_NONE:
        src = how.get("matDropoutLinModel", ref["matDropoutLinModel"])
        p.mat_drop = np.asfortranarray(np.asarray(src, dtype=np.float64).reshape(design.n_cells, -1))
    worst = 0.0
    for item in case["objective"]:
        got = oracle.objective_mudisp(X, design, m, p, item["gene"], np.asarray(item["thetas"]))
        want = np.asarray(item["values"])
        worst = max(worst, float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want)))))
    # R's dnbinom loses digits at the 1e10 dispersion clamp (n - x cancellation, ~1e-7 relative);
    # the extreme vectors sit exactly there, the random ones do not
    assert worst < 1e-6, worst
    typical = 0.0
    for item in case["objective"]:
        th = np.asarray(item["thetas"])
        keep = np.max(np.abs(th), axis=1) < 20
        if keep.any():
            got = oracle.objective_mudisp(X, design, m, p, item["gene"], th[keep])
            want = np.asarray(item["values"])[keep]
            typical = max(typical, float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want)))))
    assert typical < 1e-12, typical


def test_impulse_start_values_of_the_executed_reference(oracle):
    """initialiseImpulseParameters (initialiseImpulseParameters.R:29-111) executed from the reference's text, lm() by an
    independent least-squares solver, against the oracle's start values."""
    case = next(c for c in CASES if "starts" in c)
    X, design, m = _design(case)
    peak, valley = oracle.impulse_starts(X, design)
    for i, st in enumerate(case["starts"]):
        assert np.max(rel(peak[i], st["peak"])) < 1e-13 and np.max(rel(valley[i], st["valley"])) < 1e-13


_THIS_MACHINE = []


def _stored_bits_are_this_machines(oracle):
    """Re-executes three cheap cases from /root/reference and says whether the committed JSON is reproduced bit for bit,
    i.e. whether this machine's C library rounds like the one the JSON was made on (see the comment in
    test_oracle_reproduces_the_executed_reference).  The list of keys that differ; empty = reproduced."""
    if not _THIS_MACHINE:
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
        import make_reference_exec_golden as gen
        ref = gen.Reference("/root/reference", dnbinom=lambda x, s, mu: oracle.dnbinom_mu_log(x, s, mu))
        ref.I.lm_solver = gen.lm_in_the_oracle_s_order
        diff = []
        for name in ("H0 constant ZINB, drop-out EM (fit A)", "dispersion groups + dispersion batch confounder, NB",
                     "groups mean, dispersion splines (df 3), NB"):
            case = next(c for c in CASES if c["name"] == name)
            data = {k: np.asarray(v) for k, v in case["data"].items() if k != "ns"}
            _, out = gen.fit_model(ref, data, **case["how"])
            for key in ("matMuModel", "matDispModel", "vecLL", "vecEMLogLikModel", "matDropoutLinModel", "lsmatBatchModelDisp"):
                if out[key] != case["reference"][key]:
                    diff.append((name, key))
        _THIS_MACHINE.append(diff)
    return _THIS_MACHINE[0]


def _exact(oracle):
    """True where a comparison with the stored API outputs is to the last bit: always, except on a build machine that
    demonstrably rounds differently from the one the JSON was made on (then: north_star tolerances, with a warning)."""
    if not glob.glob("/root/reference/R/*.R") or not _stored_bits_are_this_machines(oracle):
        return True
    warnings.warn("tests/golden/reference_exec.json holds another machine's bits (C library exp / log variant): comparing "
                  "at tolerances; regenerate it here with tests/golden/make_reference_exec_golden.py")
    return False


def _same(got, want, exact, tol=1e-7):
    """Equality of two result columns (lists with None for NA, or arrays): to the last bit, or within tol relative."""
    if exact:
        return got == want if isinstance(got, list) else np.array_equal(got, want)
    if isinstance(got, list) and any(isinstance(v, (str, bool)) for v in got if v is not None):
        return got == want
    a = np.array([np.nan if v is None else v for v in np.asarray(got, dtype=object).ravel()], dtype=np.float64)
    b = np.array([np.nan if v is None else v for v in np.asarray(want, dtype=object).ravel()], dtype=np.float64)
    ok = ~np.isnan(b)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and bool(np.all(rel(a[ok], b[ok]) < tol))


@pytest.mark.skipif(not glob.glob("/root/reference/R/*.R"), reason="the reference's sources exist in the build container only")
def test_committed_golden_is_what_the_reference_text_produces(oracle):
    """Re-executes three cheap cases from /root/reference and compares with the committed JSON (bit for bit).  On a
    machine whose C library rounds differently this is where it shows: the JSON has to be regenerated there."""
    diff = _stored_bits_are_this_machines(oracle)
    if diff:
        pytest.skip(f"the committed JSON holds another machine's bits ({diff[:2]}...): regenerate it on this machine")


def test_kernel_mirror_against_the_executed_reference(oracle):
    """The GPU fit kernel's host mirror (tests/csrc/mirror_harness.cpp: bit for bit the kernel, asserted on the GPU) on
    the inputs of the executed reference.  What separates the two is arithmetic mode only -- device exp / log and 32
    partial sums against libm and a long-double running sum -- so the well-conditioned fits land on the reference's
    optimum (north_star: log-likelihood 1e-6, parameters 1e-4) and the impulse fits on it or a neighbouring one."""
    from tests.helpers import best_start, theta0_items
    from tests.mirror_harness import Mirror
    mir = Mirror()
    n = 0
    for case in FITS:
        how = case["how"]
        if how["strDropModel"] != "none" and "matDropoutLinModel" not in how:
            continue                       # the drop-out EM is not a fit-kernel item
        if "mpmath" in case["name"] or "oracle's operation order" in case["name"]:
            continue
        X, design, m = _design(case)
        p0 = oracle.init_params(X, design, m)
        drop = None
        if "matDropoutLinModel" in how:
            drop = np.asfortranarray(np.asarray(how["matDropoutLinModel"], dtype=np.float64))
            p0.mat_drop = drop
        th0, n_starts = theta0_items(oracle, X, design, m, p0)
        mr = mir.fit_items(design, m, X, th0, n_starts, drop=drop)
        _, ll = best_start(mr["ll_items"], n_starts)
        ll_ref = np.asarray(case["reference"]["vecLL"])
        r = rel(ll, ll_ref)
        if how["strMuModel"] == "impulse":
            assert np.mean(r < 1e-6) >= 0.3 and r.max() < 5e-3, (case["name"], r)
        else:
            assert r.max() < 1e-8, (case["name"], r)
        n += 1
    assert n == 12     # fits B-D style cases: constant / impulse / groups / splines, one and two confounders, given drop-out models


API_CASES = [c for c in CASES if "api" in c]


@pytest.mark.parametrize("case", API_CASES, ids=[c["name"] for c in API_CASES])
def test_oracle_pipeline_reproduces_the_executed_run_lineagepulse(oracle, case):
    """runLineagePulse (main_LineagePulse.R:219-328) executed from the reference's text -- processSCData's filters,
    calcNormConst, fitLPModels' fits A-D in either noise-model order, runDEAnalysis' table -- against the oracle pipeline
    the GPU tests and bench.py compare with: log-likelihood columns, mean_H0 and the EM trajectories to the last bit,
    p-values at 1e-10 (R's pchisq is restated twice, by mpmath there and in C here), the NA row of the all-zero gene."""
    d, how, api = case["data"], case["how"], case["api"]
    res = api["dfResults"]
    all_genes = [f"gene_{i + 1}" for i in range(len(d["counts"]))]
    all_cells = [f"cell_{j + 1}" for j in range(len(d["t"]))]
    gi = [all_genes.index(g) for g in api["genes_kept"]]
    ci = [all_cells.index(c) for c in api["cells_kept"]]
    counts = np.asarray(d["counts"])
    assert gi == [i for i in range(len(all_genes)) if counts[i].sum() > 0]          # processSCData.R:285-289
    assert ci == [j for j in range(len(all_cells)) if counts[:, j].sum() > 0]
    X = np.ascontiguousarray(counts[np.ix_(gi, ci)], dtype=np.int32)
    G, N = X.shape
    t = np.asarray(d["t"], dtype=np.float64)[ci]
    sf = np.asarray(d["sf"], dtype=np.float64)[ci]
    batch = [d["batch"][j] for j in ci]
    groups = [d["groups"][j] for j in ci]
    h1_mu = how["strMuModel"]
    h1_disp = how.get("strDispModelFull", "constant")
    design = A.DesignData(
        G, N, size_factors=sf, pseudotime=t, groups=groups if h1_mu == "groups" else None,
        spline_basis_mu=ns(t, int(how["scaDFSplinesMu"])) if h1_mu == "splines" else None,
        spline_basis_disp=ns(t, int(how["scaDFSplinesDisp"])) if h1_disp == "splines" else None,
        impulse_basis=ns(t, 4), confounders_mu=[batch] if how.get("vecConfoundersMu") else None)
    drop = how["strDropModel"]
    m_h1, m_h0 = model(h1_mu, h1_disp, drop), model("constant", "constant", drop)
    m_h1_nb, m_h0_nb = model(h1_mu, h1_disp), model("constant", "constant")
    noise_on_h0 = how.get("boolEstimateNoiseBasedOnH0", True)
    m_a, m_b = (m_h0, m_h1) if noise_on_h0 else (m_h1, m_h0)
    p_a = A.ParamsData(design, m_a)
    r_a = oracle.fit_model(X, design, m_a, p_a, True)
    p_b = A.ParamsData(design, m_b)
    p_b.mat_drop = p_a.mat_drop.copy(order="F")
    oracle.fit_model(X, design, m_b, p_b, False)
    p_h1, p_h0 = (p_b, p_a) if noise_on_h0 else (p_a, p_b)
    p_h0_nb, p_h1_nb = A.ParamsData(design, m_h0_nb), A.ParamsData(design, m_h1_nb)
    oracle.fit_model(X, design, m_h0_nb, p_h0_nb, False)
    oracle.fit_model(X, design, m_h1_nb, p_h1_nb, False)
    ll = {"loglik_full_zinb": oracle.eval_loglik(X, design, m_h1, p_h1), "loglik_red_zinb": oracle.eval_loglik(X, design, m_h0, p_h0),
          "loglik_full_nb": oracle.eval_loglik(X, design, m_h1_nb, p_h1_nb), "loglik_red_nb": oracle.eval_loglik(X, design, m_h0_nb, p_h0_nb)}
    rows = [api["row_names"].index(g) for g in api["genes_kept"]]
    col = lambda name: np.array([res[name][r] for r in rows], dtype=np.float64)     # noqa: E731
    exact = _exact(oracle)
    em_a = api["vecEMLogLikH0"] if noise_on_h0 else api["vecEMLogLikH1"]
    if not exact and h1_mu == "impulse":      # another machine's bits and an ill-conditioned fit: neighbouring optima
        for name in ("loglik_full_zinb", "loglik_full_nb"):
            assert np.mean(rel(ll[name], col(name)) < 1e-6) >= 0.3 and np.max(rel(ll[name], col(name))) < 5e-3, name
        assert _same(ll["loglik_red_zinb"], col("loglik_red_zinb"), False) and _same(ll["loglik_red_nb"], col("loglik_red_nb"), False)
        return
    for name, v in ll.items():      # lm() ran in the oracle's operation order: the impulse run is exact as well
        assert _same(v, col(name), exact), name
    assert _same(r_a["em_ll"][: r_a["n_iter"]], np.asarray(em_a), exact)
    assert _same(p_h0_nb.mat_mu[:, 0], col("mean_H0"), exact, 1e-5)                # runHypothesisTests.R:99-103
    ddf = design.deg_freedom(m_h1) - design.deg_freedom(m_h0)
    assert np.all(col("df_full_zinb") == design.deg_freedom(m_h1)) and np.all(col("df_red_zinb") == design.deg_freedom(m_h0))
    assert np.all(col("df_full_nb") == design.deg_freedom(m_h1)) and np.all(col("df_red_nb") == design.deg_freedom(m_h0))
    p, padj = oracle.lrt(ll["loglik_full_zinb"], ll["loglik_red_zinb"], ddf)
    _, padj_nb = oracle.lrt(ll["loglik_full_nb"], ll["loglik_red_nb"], ddf)
    tol_p = 1e-10 if exact else 1e-4      # p moves with the deviance: 2 x |LL| x 1e-7
    assert np.max(rel(p, col("p"))) < tol_p and np.max(rel(padj, col("padj"))) < tol_p
    assert np.array_equal(col("p_nb"), col("p"))                                   # the reference's quirk (:110): p_nb = p
    assert np.max(rel(padj_nb, col("padj_nb"))) < tol_p
    # rows of filtered genes: all NA, allZero TRUE; the table is in the order of the input genes
    assert api["row_names"] == all_genes
    for r, g in enumerate(all_genes):
        assert res["allZero"][r] == (g not in api["genes_kept"])
        if g not in api["genes_kept"]:
            assert res["p"][r] is None and res["loglik_full_zinb"][r] is None and res["gene"][r] is None


def test_product_host_layer_against_the_executed_reference():
    """The host side of the product that runs without a GPU -- processSCData's filters and report, the LRT arithmetic
    of lpb200_lrt -- on the inputs and outputs of the executed runLineagePulse."""
    import lineagepulse_b200 as lp
    from lineagepulse_b200.engine import lrt
    for case in API_CASES:
        d, how, api = case["data"], case["how"], case["api"]
        X = np.asarray(d["counts"], dtype=np.float64)
        cells = [f"cell_{j + 1}" for j in range(X.shape[1])]
        genes = [f"gene_{i + 1}" for i in range(X.shape[0])]
        sf = dict(zip(cells, d["sf"]))
        r = lp.processSCData(X, {"cell": cells, "continuous": d["t"], "groups": d["groups"], "batch": d["batch"]},
                             strMuModel=how["strMuModel"], vecConfoundersMu=how.get("vecConfoundersMu"),
                             strDispModelFull=how.get("strDispModelFull", "constant"), scaDFSplinesMu=how.get("scaDFSplinesMu", 6),
                             scaDFSplinesDisp=how.get("scaDFSplinesDisp", 3), vecNormConstExternal=sf, boolVerbose=False,
                             rownames=genes, colnames=cells)
        obj = r["objLP"]
        ref_lines = api["strReport"].split("\n")
        got_lines = obj.strReport.split("\n")
        n = len([ln for ln in got_lines if ln])
        assert n >= 4 and got_lines[1:n] == ref_lines[1:n], (got_lines[:n], ref_lines[:n])   # line 0 carries the version
        assert list(obj.vecAllGenes) == genes
        assert list(np.asarray(obj.dfAnnotationProc["cell"])) == api["cells_kept"]
        assert [float(r["vecNormConstExternalProc"][k]) for k in range(len(api["cells_kept"]))] == [sf[c] for c in api["cells_kept"]]
        res = api["dfResults"]
        rows = [api["row_names"].index(g) for g in api["genes_kept"]]
        col = lambda name: np.array([res[name][k] for k in rows], dtype=np.float64)     # noqa: E731
        ddf = int(res["df_full_zinb"][rows[0]] - res["df_red_zinb"][rows[0]])
        p, padj = lrt(col("loglik_full_zinb"), col("loglik_red_zinb"), ddf)
        assert np.max(rel(p, col("p"))) < 1e-10 and np.max(rel(padj, col("padj"))) < 1e-10
        _, padj_nb = lrt(col("loglik_full_nb"), col("loglik_red_nb"), ddf)
        assert np.max(rel(padj_nb, col("padj_nb"))) < 1e-10


@pytest.mark.skipif(not glob.glob("/root/reference/R/*.R"), reason="the reference's sources exist in the build container only")
def test_r_wrappers_executed_inside_the_reference_s_run_lineagepulse(oracle):
    """r_shim/R/lpb200_wrappers.R EXECUTED: the reference's R files are loaded into the evaluator, then the wrapper file
    on top of them (its fitLPModels / fitModel / fitMuDisp / fitPi / evalLogLikMatrix replace the reference's), and the
    reference's own runLineagePulse is called.  .Call() is answered by tests/r_shim_mock.py: lp_rshim.c's argument and
    result conventions over the CPU oracle.  The results table must equal the one the unmodified reference produced
    (tests/golden/reference_exec.json) -- every line of lpAttachCounts, lpContext, lpSetDesign, lpDesign, lpModel,
    lpParams, lpControl, lpSplitBatch, fitModel, fitModelsConcurrently and fitLPModels runs on the way."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_reference_exec_golden as gen
    from tests.r_shim_mock import ShimMock
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = gen.Reference("/root/reference", dnbinom=lambda x, s, mu: oracle.dnbinom_mu_log(x, s, mu))
    ref.I.lm_solver = gen.lm_in_the_oracle_s_order
    ref.I.load(open(os.path.join(root, "r_shim", "R", "lpb200_wrappers.R")).read(), only_definitions=False)
    mock = ShimMock(oracle)
    ref.I.dot_call = mock
    exact = _exact(oracle)
    for case in API_CASES:
        if not exact and "impulse" in case["name"]:
            continue        # another machine's bits: the impulse table is not comparable entry by entry
        mock.calls.clear()
        data = {k: np.asarray(v) for k, v in case["data"].items() if k != "ns"}
        out = gen.run_lineagepulse(ref, data, **case["how"])
        want = case["api"]
        assert out["row_names"] == want["row_names"] and out["genes_kept"] == want["genes_kept"]
        for name in ("loglik_full_zinb", "loglik_red_zinb", "loglik_full_nb", "loglik_red_nb", "mean_H0", "p", "padj", "p_nb",
                     "padj_nb", "df_full_zinb", "df_red_zinb", "df_full_nb", "df_red_nb", "allZero", "gene"):
            assert _same(out["dfResults"][name], want["dfResults"][name], exact, 1e-4 if name.startswith("p") or name == "mean_H0" else 1e-7), \
                (case["name"], name)
        noise_on_h0 = case["how"].get("boolEstimateNoiseBasedOnH0", True)
        key = "vecEMLogLikH0" if noise_on_h0 else "vecEMLogLikH1"
        assert _same(out[key], want[key], exact), case["name"]
        # one context and one upload per run, the whole EM and the three other fits in one call each, four likelihood
        # evaluations for the table; the design is re-sent only when it changes (H0 <-> H1 covariates)
        assert mock.calls.count("lp_create") == 1 and mock.calls.count("lp_load_counts_dense") == 1
        assert mock.calls.count("lp_fit_model") == 1 and mock.calls.count("lp_fit_models") == 1
        assert mock.calls.count("lp_eval_loglik") == 4 and mock.calls.count("lp_set_control") == 2
        assert 2 <= mock.calls.count("lp_set_design") <= 6, mock.calls


@pytest.mark.skipif(not glob.glob("/root/reference/R/*.R"), reason="the reference's sources exist in the build container only")
def test_r_wrappers_under_the_reference_s_own_em_loop(oracle):
    """The other integration INTEGRATION.md offers: fitModel and fitLPModels stay the reference's R code (its EM loop, its
    sequence of fits) and only the L2 functions -- fitMuDisp, fitPi, evalLogLikMatrix -- are the wrappers.  Executed as
    above; the table must again equal the unmodified reference's.  lpLRT is called directly."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_reference_exec_golden as gen
    from tests import r_parse
    from tests.r_eval import RVec
    from tests.r_shim_mock import ShimMock
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = gen.Reference("/root/reference", dnbinom=lambda x, s, mu: oracle.dnbinom_mu_log(x, s, mu))
    ref.I.lm_solver = gen.lm_in_the_oracle_s_order
    keep = {name: ref.I.globals.vars[name] for name in ("fitModel", "fitLPModels")}
    ref.I.load(open(os.path.join(root, "r_shim", "R", "lpb200_wrappers.R")).read(), only_definitions=False)
    ref.I.globals.vars.update(keep)
    # the one line INTEGRATION.md asks the maintainer to add to processSCData: attach the count matrix to its GPU context
    ref.I.load("""
processSCData_reference <- processSCData
processSCData <- function(...) {
    lsProcessed <- processSCData_reference(...)
    lsProcessed$objLP@matCountsProc <- lpAttachCounts(lsProcessed$objLP@matCountsProc)
    lsProcessed
}
""", only_definitions=False)
    mock = ShimMock(oracle)
    ref.I.dot_call = mock
    exact = _exact(oracle)
    for case in API_CASES:
        if "impulse" in case["name"]:
            continue
        mock.calls.clear()
        data = {k: np.asarray(v) for k, v in case["data"].items() if k != "ns"}
        out = gen.run_lineagepulse(ref, data, **case["how"])
        want = case["api"]
        for name in ("loglik_full_zinb", "loglik_red_zinb", "loglik_full_nb", "loglik_red_nb", "mean_H0", "p", "padj", "padj_nb", "gene"):
            assert _same(out["dfResults"][name], want["dfResults"][name], exact, 1e-4 if name.startswith("p") or name == "mean_H0" else 1e-7), \
                (case["name"], name)
        assert _same(out["vecEMLogLikH0"], want["vecEMLogLikH0"], exact) and _same(out["vecEMLogLikH1"], want["vecEMLogLikH1"], exact)
        n_cycles = len(want["vecEMLogLikH0" if case["how"].get("boolEstimateNoiseBasedOnH0", True) else "vecEMLogLikH1"])
        assert mock.calls.count("lp_fit_pi") == n_cycles and mock.calls.count("lp_fit_mudisp") == n_cycles + 3
        assert mock.calls.count("lp_fit_model") == 0 and mock.calls.count("lp_fit_models") == 0
        # the context travels with the matrix: one upload although every L2 call receives the matrix anew
        assert mock.calls.count("lp_create") == 1 and mock.calls.count("lp_load_counts_dense") == 1, mock.calls
    res = ref.call("lpLRT", vecLogLikFull=RVec(np.array([-100.0, -50.0, -70.0])), vecLogLikRed=RVec(np.array([-104.0, -50.5, -90.0])),
                   scaDeltaDegFreedom=RVec(np.array([3.0])))
    p, padj = oracle.lrt(np.array([-100.0, -50.0, -70.0]), np.array([-104.0, -50.5, -90.0]), 3)
    assert res.names == ["p", "padj"] and np.array_equal(res.v[0].v, p) and np.array_equal(res.v[1].v, padj)


@pytest.mark.skipif(not glob.glob("/root/reference/R/*.R"), reason="the reference's sources exist in the build container only")
def test_warm_start_hooks_of_the_executed_reference(oracle):
    """fitModel's warm-start hooks (fitModel.R:147-151; initialiseMuModel.R:93,131-133, initialiseDispModel.R) executed from
    the reference's text: with matMuModelInit / matDispModelInit / lsmatBatchModelInitMu supplied and no drop-out model to
    estimate, the reference's result is -- to the last bit -- fitMuDisp started from exactly those matrices (taken as stored:
    means and batch factors on the natural scale).  That is the semantics lpb200_fit_model_from implements (keep_mask;
    test_gpu_parity.py::test_warm_start_hooks asserts fit_model(keep) == fit_mudisp on the device)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_reference_exec_golden as gen
    from tests.r_eval import RList, RVec, as_py
    ref = gen.Reference("/root/reference", dnbinom=lambda x, s, mu: oracle.dnbinom_mu_log(x, s, mu))
    ref.I.lm_solver = gen.lm_in_the_oracle_s_order
    d = gen.make_data(12, size_factors=True)
    G, N = d["counts"].shape
    t = np.asarray(d["t"], dtype=np.float64)
    X = np.ascontiguousarray(d["counts"], dtype=np.int32)
    mat, df, sf = gen.r_inputs(d)
    genes = [f"gene_{i + 1}" for i in range(G)]
    rng = np.random.default_rng(3)

    def rmat(a):
        return RVec(a.reshape(-1, order="F"), dim=a.shape, dimnames=(genes, None))

    for mu_model, conf in (("constant", None), ("groups", ["batch"])):
        design = A.DesignData(G, N, size_factors=np.asarray(d["sf"], dtype=np.float64), pseudotime=t,
                              groups=d["groups"] if mu_model == "groups" else None, impulse_basis=ns(t, 4),
                              confounders_mu=[d["batch"]] if conf else None)
        m = model(mu_model, "constant", "none")
        p = oracle.init_params(X, design, m)
        p.mat_mu = np.asfortranarray(p.mat_mu * rng.uniform(0.5, 2.0, p.mat_mu.shape))
        p.mat_disp = np.asfortranarray(rng.uniform(0.5, 3.0, p.mat_disp.shape))
        kw = dict(matMuModelInit=rmat(p.mat_mu), matDispModelInit=rmat(p.mat_disp))
        if conf:
            p.batch_mu = [np.asfortranarray(np.column_stack([np.ones(G), rng.uniform(0.7, 1.5, (G, 2))]))]
            kw["lsmatBatchModelInitMu"] = RList([rmat(p.batch_mu[0])])
        res = ref.call("fitModel", matCounts=mat, dfAnnotation=df, vecConfoundersMu=None if conf is None else gen.S(conf),
                       vecConfoundersDisp=None, vecNormConst=sf, scaDFSplinesMu=None, scaDFSplinesDisp=None, matWeights=None,
                       matPiConstPredictors=None, lsDropModel=None, strMuModel=gen.S(mu_model), strDispModel=gen.S("constant"),
                       strDropModel=gen.S("none"), strDropFitGroup=gen.S("PerCell"), scaMaxEstimationCycles=gen.D(20),
                       boolVerbose=gen.B(False), boolSuperVerbose=gen.B(False), **kw)
        oracle.fit_mudisp(X, design, m, p)
        mu_ref = np.asarray(as_py(res.get("lsMuModel").get("matMuModel")), dtype=np.float64).reshape(p.mat_mu.shape)
        disp_ref = np.asarray(as_py(res.get("lsDispModel").get("matDispModel")), dtype=np.float64).reshape(p.mat_disp.shape)
        assert np.array_equal(p.mat_mu, mu_ref) and np.array_equal(p.mat_disp, disp_ref), mu_model
        if conf:
            assert np.array_equal(p.batch_mu[0], np.asarray(as_py(res.get("lsMuModel").get("lsmatBatchModel").v[0])))


@pytest.mark.skipif(not glob.glob("/root/reference/R/*.R"), reason="the reference's sources exist in the build container only")
def test_fit_extraction_of_the_executed_reference(oracle):
    """getFitsMean / getFitsDispersion / getFitsDropout / getPostDrop (calcPostDrop_Matrix) / getNormData with its two
    switches (getFits.R:52-363, calcPostDrop.R:69-125), executed from the reference's text on a model the reference's
    fitModel has just fitted -- groups mean and spline dispersion, a batch confounder on each, logistic_ofMu drop-out, size
    factors -- against the oracle's expansions, which are what the GPU's lpb200_expand_fits is compared with
    (test_gpu_parity.py::test_fit_extraction_parity, test_post_drop_parity): equal to the last bit, the quirk that the
    posterior ignores size factors (calcPostDrop.R:110) included."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_reference_exec_golden as gen
    ref = gen.Reference("/root/reference", dnbinom=lambda x, s, mu: oracle.dnbinom_mu_log(x, s, mu))
    ref.I.lm_solver = gen.lm_in_the_oracle_s_order
    d = gen.make_data(12, size_factors=True)
    G, N = d["counts"].shape
    t = np.asarray(d["t"], dtype=np.float64)
    X = np.ascontiguousarray(d["counts"], dtype=np.int32)
    mat = gen.r_inputs(d)[0]
    # the drop-out model of a constant-mean EM (seconds), then the rich model given it (the per-cell fits are the slow part of
    # the evaluator)
    drop_a = gen.fit_model(ref, d, strMuModel="constant", strDispModel="constant", strDropModel="logistic_ofMu")[0].get("lsDropModel")
    res, out = gen.fit_model(ref, d, strMuModel="groups", strDispModel="splines", strDropModel="logistic_ofMu", scaDFSplinesDisp=3,
                             vecConfoundersMu=["batch"], vecConfoundersDisp=["batch"], lsDropModel=drop_a)
    mu, disp, drop = res.get("lsMuModel"), res.get("lsDispModel"), res.get("lsDropModel")
    design = A.DesignData(G, N, size_factors=np.asarray(d["sf"], dtype=np.float64), pseudotime=t, groups=d["groups"],
                          spline_basis_disp=ns(t, 3), impulse_basis=ns(t, 4), confounders_mu=[d["batch"]], confounders_disp=[d["batch"]])
    m = model("groups", "splines", "logistic_ofMu")
    p = A.ParamsData(design, m)
    p.mat_mu = np.asfortranarray(np.asarray(out["matMuModel"], dtype=np.float64).reshape(p.mat_mu.shape))
    p.mat_disp = np.asfortranarray(np.asarray(out["matDispModel"], dtype=np.float64).reshape(p.mat_disp.shape))
    p.batch_mu = [np.asfortranarray(np.asarray(out["lsmatBatchModelMu"][0], dtype=np.float64))]
    p.batch_disp = [np.asfortranarray(np.asarray(out["lsmatBatchModelDisp"][0], dtype=np.float64))]
    p.mat_drop = np.asfortranarray(np.asarray(out["matDropoutLinModel"], dtype=np.float64).reshape(p.mat_drop.shape))
    ids, gi = gen.S(["gene_2", "gene_7", "gene_1"]), [1, 6, 0]
    got = {"getFitsMean": (ref.call("getFitsMean", lsMuModel=mu, vecGeneIDs=ids), A.EXPAND_MEAN),
           "getFitsDispersion": (ref.call("getFitsDispersion", lsDispModel=disp, vecGeneIDs=ids), A.EXPAND_DISP),
           "getFitsDropout": (ref.call("getFitsDropout", lsMuModel=mu, lsDropModel=drop, vecGeneIDs=ids), A.EXPAND_DROP),
           "getNormData": (ref.call("getNormData", matCounts=mat, lsMuModel=mu, vecGeneIDs=ids), A.EXPAND_NORMDATA),
           "getNormData(boolDepth = FALSE)": (ref.call("getNormData", matCounts=mat, lsMuModel=mu, vecGeneIDs=ids, boolDepth=gen.B(False)),
                                              A.EXPAND_NORMDATA | A.EXPAND_NO_DEPTH),
           "getNormData(boolBatch = FALSE)": (ref.call("getNormData", matCounts=mat, lsMuModel=mu, vecGeneIDs=ids, boolBatch=gen.B(False)),
                                              A.EXPAND_NORMDATA | A.EXPAND_NO_BATCH)}
    for name, (r, what) in got.items():
        assert r.dim == (3, N), name
        assert np.array_equal(oracle.expand_fits(X, design, m, p, gi, what), r.mat()), name
    r = ref.call("getPostDrop", matCounts=mat, lsMuModel=mu, lsDispModel=disp, lsDropModel=drop, vecGeneIDs=ids)
    assert np.array_equal(oracle.post_drop(X, design, m, p, gi), r.mat())


@pytest.mark.skipif(not glob.glob("/root/reference/R/*.R"), reason="the reference's sources exist in the build container only")
def test_test_dropout_of_the_executed_reference(oracle):
    """testDropout (runHypothesisTests.R:179-214) executed from the reference's text on the object its own runLineagePulse
    returned, against the product's testDropout on the same results table and drop-out model: every column equal to the last
    bit for the per-cell drop-out model, and for "ForAllCells" the reference's quirk -- it tests for "AllCells", so no
    drop-out parameter is counted, the difference in degrees of freedom is 0 and the p-value 0."""
    import sys
    import pandas as pd
    import lineagepulse_b200 as lp
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_reference_exec_golden as gen
    from tests.r_eval import as_py
    ref = gen.Reference("/root/reference", dnbinom=lambda x, s, mu: oracle.dnbinom_mu_log(x, s, mu))
    ref.I.lm_solver = gen.lm_in_the_oracle_s_order
    d = gen.make_data(21, G=8, N=24, size_factors=True)
    mat, df, sf = gen.r_inputs(d)
    df.attrs = {"class": "data.frame", "row.names": list(df.v[0].v)}
    for grp in ("PerCell", "ForAllCells"):
        obj = ref.call("runLineagePulse", counts=mat, dfAnnotation=df, boolVerbose=gen.B(False), strMuModel=gen.S("groups"),
                       strDropModel=gen.S("logistic"), strDropFitGroup=gen.S(grp), vecNormConstExternal=sf)
        want = ref.call("testDropout", objLP=obj)
        slot = lambda name: obj.v[obj.names.index(name)]     # noqa: E731
        res, drop = slot("dfResults"), slot("lsDropModel")
        table = pd.DataFrame({n: [np.nan if x is None else x for x in as_py(c)] for n, c in zip(res.names, res.v)
                              if n not in ("gene", "allZero")})
        mine = lp.LineagePulseObject(dfResults=table, lsDropModel={
            "matDropoutLinModel": np.asarray(as_py(drop.get("matDropoutLinModel")), dtype=np.float64),
            "lsDropModelGlobal": {"strDropFitGroup": drop.get("lsDropModelGlobal").get("strDropFitGroup").v[0]}})
        got = lp.testDropout(mine)
        assert list(got.columns) == want.names
        for name, col in zip(want.names, want.v):
            assert got[name].tolist() == as_py(col), (grp, name)
        assert (got["df_delta"][0] == 0 and got["p"][0] == 0.0) == (grp == "ForAllCells")
