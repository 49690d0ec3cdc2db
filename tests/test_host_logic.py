"""CPU-only tests of the host side: C-ABI surface, ns(), simulator, data preparation, LRT."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import lineagepulse_b200 as lp
from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.splines import _bspline_basis_all, ns
from tests.helpers import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lpb200.h")).read()
    declared = sorted(set(re.findall(r"\b(lpb200_[a-z0-9_]+)\s*\(", hdr)) - {"lpb200_allreduce_fn"})
    assert len(declared) >= 25
    lib = C.CDLL(A.default_library_path())
    for name in declared:
        assert hasattr(lib, name), f"liblpb200.so does not export {name}"
    bound = A.declare_product_prototypes(lib)
    assert sorted(bound) == declared  # the ctypes mirror covers exactly the header
    assert lib.lpb200_abi_version() == 1
    ctl = A.Control()
    lib.lpb200_default_control(C.byref(ctl))
    assert (ctl.maxit_mudisp, ctl.maxit_pi, ctl.max_cycles) == (1000, 10000, 20)  # fitModel.R:164-175
    assert ctl.reltol_mudisp == 1e-8 and ctl.ndeps == 1e-3 and ctl.prec_em == 1 - 1e-4


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lp.LPB200Error):
        lp.Engine(0)
    with pytest.raises(lp.LPB200Error):
        lp.Engine(devices=[0, 1])       # lpb200_create_multi: NULL without devices
    with pytest.raises(lp.LPB200Error):
        lp.evalLogLikMatrix(np.ones((3, 4), dtype=np.int32), {"lsMuModelGlobal": {"strMuModel": "constant", "vecNormConst": np.ones(4)}, "matMuModel": np.ones((3, 1))},
                            {"lsDispModelGlobal": {"strDispModel": "constant"}, "matDispModel": np.ones((3, 1))},
                            {"lsDropModelGlobal": {"strDropModel": "none"}})


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "lineagepulse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "lporacle" not in src and "liblporacle" not in src, f
                assert not re.search(r"^\s*(from|import)\s+\S*oracle", src, re.M), f
                assert not re.search(r"#include\s+\"[^\"]*oracle", src), f


def test_ns_against_scipy_and_r_pattern():
    from scipy.interpolate import BSpline
    x = np.linspace(0, 1, 200)
    B = ns(x, 6)
    assert B.shape == (200, 6) and np.linalg.matrix_rank(B) == 6
    # the boundary row R prints for this design
    assert np.allclose(B[-1], [0, 0, 0, -1 / 7, 3 / 7, 5 / 7], atol=1e-12)
    ik = np.quantile(x, np.linspace(0, 1, 6)[1:-1])
    t = np.sort(np.r_[np.repeat([0.0, 1.0], 4), ik])
    D = BSpline.design_matrix(x, t, 3).toarray()
    assert np.abs(D - _bspline_basis_all(t, x, 4)).max() < 1e-14
    # natural: second derivative vanishes at both boundary knots for every basis function
    d2 = BSpline(t, np.eye(len(t) - 4), 3).derivative(2)
    q = np.linalg.qr(_bspline_basis_all(t, np.array([0.0, 1.0]), 4, deriv=2).T, mode="complete")[0][:, 2:]
    assert np.abs(d2(np.array([0.0, 1 - 1e-13])) @ q).max() < 1e-6
    # constants and linear functions lie in the span
    for y in (np.ones_like(x), x):
        c = np.linalg.lstsq(B, y, rcond=None)[0]
        assert np.abs(B @ c - y).max() < 1e-12
    rng = np.random.default_rng(0)
    xr = np.sort(rng.uniform(size=57))
    assert ns(xr, 4).shape == (57, 4) and np.linalg.matrix_rank(ns(xr, 4)) == 4


def test_simulator_recipe():
    s = lp.simulateContinuousDataSet(300, 20, 20, 20, scaMumax=100, scaSDMuAmplitude=3,
                                     vecGeneWiseDropoutRates=np.full(60, 0.1), seed=7)
    X = s["counts"]
    assert X.shape == (60, 300) and X.min() >= 0
    assert np.allclose(s["annot"]["continuous"], np.arange(300) / 299)  # simulateDataSet.R:123
    zero_frac = (X == 0).mean()
    assert 0.1 <= zero_frac <= 0.6
    s2 = lp.simulateContinuousDataSet(300, 20, 20, 20, scaMumax=100, scaSDMuAmplitude=3,
                                      vecGeneWiseDropoutRates=np.full(60, 0.1), seed=7)
    assert np.array_equal(X, s2["counts"])
    with pytest.raises(ValueError):
        lp.simulateContinuousDataSet(10, 1, 1, 1, vecGeneWiseDropoutRates=np.full(5, 0.1))
    with pytest.raises(ValueError):
        lp.simulateContinuousDataSet(10, 1, 1, 1)


def test_first_appearance_index():
    lab, idx = A.first_appearance_index(np.array(["b", "a", "b", "c", "a"]))
    assert list(lab) == ["b", "a", "c"] and list(idx) == [0, 1, 0, 2, 1]  # match(v, unique(v)) - 1


def test_process_sc_data_filters_and_errors():
    X = np.array([[0, 0, 0, 0], [1, 0, 2, 0], [0, 0, 5, 1]], dtype=float)
    X[1, 3] = np.nan  # NA silently becomes 0 on sparsification (processSCData.R:262-264)
    annot = {"cell": np.array(["a", "b", "c", "d"]), "continuous": np.array([0.0, 0.3, np.nan, 1.0])}
    r = lp.processSCData(X, annot, strMuModel="impulse", boolVerbose=False, rownames=["g1", "g2", "g3"])
    obj = r["objLP"]
    # cell c (NA pseudotime) dropped, then all-zero gene g1 and all-zero cell b dropped
    assert list(obj.matCountsProc.rownames) == ["g2", "g3"]
    assert list(obj.matCountsProc.colnames) == ["a", "d"]
    assert list(obj.vecAllGenes) == ["g1", "g2", "g3"]
    assert np.array_equal(obj.matCountsProc.dense(), np.array([[1, 0], [0, 1]]))
    assert "1 out of 4 cells did not have a continuous covariate" in obj.strReport
    with pytest.raises(lp.LineagePulseError, match="non-integer"):
        lp.processSCData(np.array([[0.5, 1.0]]), {"cell": np.array(["a", "b"])}, strMuModel="constant", boolVerbose=False)
    with pytest.raises(lp.LineagePulseError, match="negative"):
        lp.processSCData(np.array([[-1.0, 1.0]]), {"cell": np.array(["a", "b"])}, strMuModel="constant", boolVerbose=False)
    with pytest.raises(lp.LineagePulseError, match="strMuModel not recognised"):
        lp.processSCData(X, annot, strMuModel="MM", boolVerbose=False)  # processSCData.R:231-234
    with pytest.raises(lp.LineagePulseError, match="Supply dfAnnotation"):
        lp.processSCData(X, None, strMuModel="constant", boolVerbose=False)
    annot_b = dict(annot, batch=np.array(["b1", None, "b2", "b1"], dtype=object))
    with pytest.raises(lp.LineagePulseError, match="Supply batch assignments for all cells"):   # processSCData.R:185-190
        lp.processSCData(X, annot_b, vecConfoundersMu=["batch"], strMuModel="constant", boolVerbose=False)
    with pytest.raises(lp.LineagePulseError, match="Not all confounders given in vecConfoundersMu"):
        lp.processSCData(X, annot, vecConfoundersMu=["batch"], strMuModel="constant", boolVerbose=False)


def test_process_sc_data_mtx_and_sparse_ingest(tmp_path):
    """processSCData.R:102-112: counts given as a MatrixMarket file name (readMM), as a sparse matrix or as a dense one
    lead to the same processed matrix, annotation and report; any other file name is the reference's error."""
    import scipy.io
    import scipy.sparse as sp
    rng = np.random.default_rng(4)
    X = rng.poisson(0.6, size=(7, 9)).astype(np.int64)
    X[2, :] = 0                        # an all-zero gene
    X[:, 4] = 0                        # an all-zero cell
    cells = np.array([f"c{j}" for j in range(9)])
    genes = [f"g{i}" for i in range(7)]
    t = np.linspace(0.0, 1.0, 9)
    t[7] = np.nan                      # a cell without pseudotime
    annot = {"cell": cells, "continuous": t}
    path = str(tmp_path / "counts.mtx")
    scipy.io.mmwrite(path, sp.csc_matrix(X))
    kw = dict(strMuModel="splines", scaDFSplinesMu=3, boolVerbose=False, rownames=genes, colnames=cells)
    objs = [lp.processSCData(src, annot, **kw)["objLP"] for src in (path, sp.csc_matrix(X), X.astype(float), X.astype(np.int32))]
    keep_g = [i for i in range(7) if i != 2 and X[i, [j for j in range(9) if j not in (4, 7)]].sum() > 0]
    keep_c = [j for j in range(9) if j not in (4, 7) and X[:, j].sum() > 0]
    want = X[np.ix_(keep_g, keep_c)]
    for obj in objs:
        assert np.array_equal(obj.matCountsProc.dense(), want)
        assert list(obj.matCountsProc.rownames) == [genes[i] for i in keep_g]
        assert list(obj.matCountsProc.colnames) == [cells[j] for j in keep_c]
        assert list(obj.vecAllGenes) == genes
        assert obj.strReport.split("\n")[1:] == objs[0].strReport.split("\n")[1:]
    with pytest.raises(lp.LineagePulseError, match="Input matrix file is not .mtx"):
        lp.processSCData(str(tmp_path / "counts.csv"), annot, **kw)


def test_process_sc_data_column_order_is_the_executed_reference_s():
    """processSCData.R:255-258 on counts whose columns are ordered differently from dfAnnotation$cell: the reference sorts the
    count columns by match(colnames(counts), dfAnnotation$cell) -- the inverse of the lining-up permutation
    -- and pairs by position afterwards, size factors by column name.  tests/golden/process_sc_data_order.json holds what the
    reference's own text returns (executed by tests/r_eval.py, make_process_order_golden.py): swaps line up, cycles do not,
    and the product returns the same object either way -- with a warning where the names no longer line up."""
    import json
    import warnings as W
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "process_sc_data_order.json")))["cases"]
    assert len(cases) == 6
    for case in cases:
        ref = case["reference"]
        X = np.asarray(case["counts"], dtype=np.int64)
        t = np.array([np.nan if v is None else v for v in case["t"]])
        annot = {"cell": np.array(case["cells"]), "continuous": t}
        sf = dict(zip(case["cells"], case["sf"]))
        lined_up = ref["colnames"] == ref["annotation_cell"]
        assert lined_up == (case["name"] in ("same order", "one swap", "two swaps", "reversed"))
        for counts in (X.astype(float), X.astype(np.int32)):
            with W.catch_warnings(record=True) as caught:
                W.simplefilter("always")
                r = lp.processSCData(counts, annot, strMuModel="splines", scaDFSplinesMu=3, vecNormConstExternal=sf,
                                     boolVerbose=False, rownames=case["genes"], colnames=case["colnames"])
            assert any("processSCData.R:255-258" in str(w.message) for w in caught) == (not lined_up), case["name"]
            obj = r["objLP"]
            assert list(obj.matCountsProc.colnames) == ref["colnames"], case["name"]
            assert list(obj.matCountsProc.rownames) == ref["rownames"]
            assert np.array_equal(obj.matCountsProc.dense(), np.asarray(ref["counts"])), case["name"]
            assert list(obj.dfAnnotationProc["cell"]) == ref["annotation_cell"]
            assert list(obj.dfAnnotationProc["continuous"]) == ref["annotation_continuous"]
            assert list(r["vecNormConstExternalProc"]) == ref["vecNormConstExternalProc"], case["name"]
            assert obj.strReport.split("\n")[1:4] == ref["report"][:3]


def test_object_accessors():
    obj = lp.LineagePulseObject(strReport="x", vecAllGenes=np.array(["g"]))
    assert "dfResults" in obj.names() and obj["strReport"] == "x"
    from lineagepulse_b200 import api
    assert api.strReport(obj) == "x" and api.vecAllGenes(obj)[0] == "g" and api.lsDropModel(obj) is None
    assert obj.boolFixedPopulations is False


def test_lrt_host_arithmetic_vs_mpmath_and_oracle(oracle):
    rng = np.random.default_rng(0)
    for df in (1, 3, 6, 7):
        q = np.r_[rng.exponential(8, 100), 0.0, -2.0, 900.0, 1600.0]
        llr = -rng.uniform(100, 1000, len(q))
        llf = llr + q / 2
        p, padj = lp.lrt(llf, llr, df)
        po, pao = oracle.lrt(llf, llr, df)
        assert np.allclose(p, po, rtol=1e-12, atol=1e-300)
        assert np.allclose(padj, pao, rtol=1e-12, atol=1e-300)
        assert p[-3] == 1.0 and p[-4] == 1.0  # deviance <= 0 => p = 1
    for r in golden("pchisq_upper.json"):
        p, _ = lp.lrt(np.array([r["q"] / 2]), np.array([0.0]), r["df"])
        assert abs(p[0] - r["value"]) <= 1e-12 * r["value"] + 1e-300
    p, padj = lp.lrt(np.array([1.0, np.nan, 3.0]), np.array([0.0, 0.0, 0.0]), 2)
    assert np.isnan(p[1]) and np.isnan(padj[1]) and not np.isnan(padj[0])


def _build_rshim_harness():
    """r_shim/lp_rshim.c + the functional stand-in R runtime + the driver, linked against liblpb200.so."""
    import subprocess
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    out = os.path.join(build, "librshimharness.so")
    srcs = [os.path.join(ROOT, "r_shim", "lp_rshim.c"), os.path.join(ROOT, "tests", "csrc", "fake_r_runtime.c"),
            os.path.join(ROOT, "tests", "csrc", "rshim_driver.c")]
    libdir = os.path.join(ROOT, "lineagepulse_b200")
    deps = srcs + [os.path.join(ROOT, "include", "lpb200.h"), os.path.join(ROOT, "r_shim", "fake_R", "Rinternals.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["gcc", "-std=gnu11", "-O1", "-fPIC", "-shared", "-Werror=implicit-function-declaration",
                               "-Werror=incompatible-pointer-types", "-I", os.path.join(ROOT, "r_shim", "fake_R")] + srcs +
                              ["-o", out, "-Wl,-Bsymbolic", "-L", libdir, "-l:liblpb200.so", f"-Wl,-rpath,{libdir}", "-lm"])   # -Bsymbolic: error() is also a glibc symbol
    return out


def test_r_shim_compiles_and_links():
    """R is absent from the image: the .Call shim is compiled against r_shim/fake_R (a stand-in declaring only the R
    API it uses), linked with the functional stand-in runtime and liblpb200.so; without a GPU its lp_create must
    raise an R error (no CPU fallback).  The GPU run is tests/test_gpu_parity.py::test_r_shim_runs_on_the_gpu."""
    lib = C.CDLL(_build_rshim_harness())
    src = open(os.path.join(ROOT, "r_shim", "lp_rshim.c")).read()
    for sym in ("lpb200_fit_mudisp", "lpb200_fit_pi", "lpb200_eval_loglik", "lpb200_fit_model", "lpb200_lrt",
                "lpb200_load_counts_csc", "lpb200_set_design", "lpb200_create_multi", "lpb200_fit_model_multi"):
        assert sym in src
    wr = open(os.path.join(ROOT, "r_shim", "R", "lpb200_wrappers.R")).read()
    for fn in ("lpAttachCounts <- function", "fitMuDisp <- function", "fitPi <- function", "evalLogLikMatrix <- function",
               "fitModel <- function", "fitLPModels <- function", "fitModelsConcurrently <- function"):
        assert fn in wr
    assert wr.count("(") == wr.count(")") and wr.count("{") == wr.count("}")   # R cannot parse it here; at least balanced
    err = C.create_string_buffer(512)
    assert lib.rs_set_control_checks(err) == 0, err.value
    assert b"bad optimiser constants" in err.value
    import torch
    if not torch.cuda.is_available():
        err = C.create_string_buffer(512)
        assert lib.rs_released_context_raises(err) == 1
        assert b"no usable CUDA device" in err.value


def test_test_dropout_table():
    """testDropout (runHypothesisTests.R:179-214): LRT of the ZINB against the NB fits over the whole data set;
    the drop-out model's parameters count as N x q for "PerCell", q for "AllCells" (the reference tests that
    string although fitPi knows "ForAllCells": with "ForAllCells" nothing is added), NA rows are skipped."""
    import pandas as pd
    from scipy.stats import chi2
    res = pd.DataFrame({"df_full_zinb": [8.0, 8.0, np.nan, 8.0], "df_full_nb": [8.0, 8.0, np.nan, 8.0],
                        "loglik_full_zinb": [-100.0, -250.5, np.nan, -80.25], "loglik_full_nb": [-130.0, -300.0, np.nan, -99.0]})
    for grp, extra in (("PerCell", 5 * 2), ("AllCells", 2), ("ForAllCells", 0)):
        obj = lp.LineagePulseObject(dfResults=res, lsDropModel={"matDropoutLinModel": np.zeros((5, 2)),
                                                              "lsDropModelGlobal": {"strDropFitGroup": grp}})
        out = lp.testDropout(obj)
        assert list(out.columns) == ["p", "df_delta", "deviance", "df_full", "df_red", "loglik_full", "loglik_red"]
        row = out.iloc[0]
        assert row["df_full"] == 24 + extra and row["df_red"] == 24 and row["df_delta"] == extra
        assert row["loglik_full"] == -430.75 and row["loglik_red"] == -529.0 and row["deviance"] == 2 * (529.0 - 430.75)
        if extra:
            assert abs(row["p"] - chi2.sf(row["deviance"], extra)) <= 1e-12 * chi2.sf(row["deviance"], extra) + 1e-300
        else:
            assert row["p"] == 0.0   # pchisq(q > 0, df = 0, lower.tail = FALSE)


def test_bench_contract_without_a_gpu():
    """bench.py's two arms on a box without a GPU: the reference arm (the CPU oracle pipeline on the host cores) prints
    ONE JSON line with the contract's keys -- the driver computes the GPU / CPU ratio from it -- and the native arm
    refuses to run instead of falling back to the CPU."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU for the native-arm half")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    common = ["--steps", "2", "--warmup", "1", "--genes", "64", "--cells", "300"]
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--cpu-budget-s", "3"] + common, cwd=root,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "genes/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "genes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    r = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-strong"] + common, cwd=root,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())


def test_r_wrappers_static_checks():
    """R cannot parse r_shim/R/lpb200_wrappers.R here (no R in the image); a lexer (tests/r_lint.py) checks what a
    never-run wrapper gets wrong: strings terminate and brackets nest; every .Call names a routine of the shim's
    registration table with that table's argument count, which is also the C function's; every list element the C
    side looks up by name is set somewhere on the R side; the wrappers keep the reference's function signatures
    (R/srcLineagePulse_fitMeanDispersion.R:853-859, fitDropout.R:418-422, evalLogLik.R:307-313, fitModel.R:137-158)."""
    from tests import r_lint
    wr = open(os.path.join(ROOT, "r_shim", "R", "lpb200_wrappers.R")).read()
    csrc = open(os.path.join(ROOT, "r_shim", "lp_rshim.c")).read()
    toks = r_lint.check_nesting(wr)
    table = {m.group(1): int(m.group(3)) for m in re.finditer(r'\{"(lp_\w+)",\s*\(DL_FUNC\)&(\w+),\s*(\d+)\}', csrc)}
    assert len(table) >= 11
    for name, arity in table.items():   # the C definitions take as many SEXPs as the table says
        m = re.search(r"^SEXP\s+" + name + r"\s*\(([^)]*)\)", csrc, re.M)
        assert m, name
        assert m.group(1).count("SEXP") == arity, (name, m.group(1), arity)
    called = set()
    for line, args in r_lint.calls(toks, ".Call"):
        assert args and args[0][0][0] == "str", f"line {line}: .Call without a routine name"
        routine = args[0][0][1]
        assert routine in table, f"line {line}: .Call of unregistered routine {routine}"
        assert len(args) - 1 == table[routine], f"line {line}: {routine} takes {table[routine]} arguments, {len(args) - 1} given"
        called.add(routine)
    # R's unary ! binds more loosely than arithmetic: `2L * !a + 4L * !b` is 2L * !(a + 4L * !b)
    for (k0, t0), (k1, t1) in zip([t[:2] for t in toks], [t[:2] for t in toks[1:]]):
        assert not (k0 == "op" and t0 in ("*", "+", "-", "/", "^") and (k1, t1) == ("op", "!")), "unparenthesised ! after " + t0
    assert {"lp_set_control", "lp_create", "lp_load_counts_csc", "lp_load_counts_dense", "lp_set_design", "lp_eval_loglik", "lp_fit_mudisp",
            "lp_fit_pi", "lp_fit_model", "lp_fit_models", "lp_lrt"} <= called
    # names the shim reads from R lists are names the R side writes (as list(... name = ...) or x$name <- / x[["name"]] <-)
    r_names = {t[1] for t in toks if t[0] in ("name", "str")}
    for name in set(re.findall(r'list_get\(\w+,\s*"(\w+)"\)', csrc)):
        assert name in r_names, f"the shim looks up list element {name!r} which the R wrappers never mention"
    # reference signatures (argument names in order; the wrappers may append optional arguments)
    defs = r_lint.function_defs(toks)
    ref = {"fitMuDisp": ["matCountsProc", "vecNormConst", "lsMuModel", "lsDispModel", "lsDropModel", "matWeights"],
           "fitPi": ["matCounts", "lsMuModel", "lsDispModel", "lsDropModel"],
           "evalLogLikMatrix": ["matCounts", "lsMuModel", "lsDispModel", "lsDropModel", "matWeights", "boolConstModelOnMMLL"]}
    for fn, formals in ref.items():
        assert defs[fn][: len(formals)] == formals, (fn, defs[fn])
    for fn in ("fitModel", "fitLPModels", "fitModelsConcurrently", "lpAttachCounts"):
        assert fn in defs
    # the lexer itself rejects the defects it is there for
    for bad in ('f <- function(x) { g(x[[1]) }', 'x <- "abc', 'f(a, b))'):
        with pytest.raises(SyntaxError):
            r_lint.check_nesting(bad)


R_BASE = {  # base / stats / methods names the wrappers may use (all exist in every R >= 3.4)
    ".Call", "TRUE", "FALSE", "NULL", "NA", "NA_real_", "any", "as.double", "as.integer", "as.logical", "attr", "attr<-",
    "c", "cbind", "colnames", "cumsum", "dim", "do.call", "emptyenv", "getOption", "identical", "integer", "invisible",
    "is", "is.null", "lapply", "list", "matrix", "max", "message", "ncol", "new.env", "nrow", "numeric", "paste0",
    "return", "rownames", "rownames<-", "seq_along", "seq_len", "stop", "sum", "unlist", "vapply", "which.max"}


def test_r_parser_accepts_the_reference_and_rejects_broken_sources():
    """tests/r_parse.py restates R's grammar; it is trusted only as far as it is checked: every R file of the reference
    parses (in the build container, where /root/reference exists) and yields as many function definitions as the text
    has `function(`; sources with the defects of a never-run file are rejected with the line."""
    import glob
    from tests import r_parse
    files = sorted(glob.glob("/root/reference/R/*.R"))
    n_fun = 0
    for f in files:
        src = open(f).read()
        prog = r_parse.parse(src)
        n_fun += sum(1 for n in r_parse.walk(prog) if n.kind == "function")
        code = "\n".join(ln for ln in src.splitlines() if not ln.lstrip().startswith("#"))
        assert sum(1 for n in r_parse.walk(prog) if n.kind == "function") == len(re.findall(r"\bfunction\s*\(", code)), f
    if files:
        assert len(files) == 23 and n_fun == 182
    wr = open(os.path.join(ROOT, "r_shim", "R", "lpb200_wrappers.R")).read()
    assert len(r_parse.parse(wr)) == 19          # 4 code tables + 15 functions
    # precedence and newline rules that matter for reading the wrappers
    e = r_parse.parse("-2^2; -1:3; !a == b; a <- b <- c; 1L * (!x) + 2L * (!y); x$y[[1]](z)@s")
    assert e[0].kind == "unop" and e[0].a[1].a[0] == "^"
    assert e[1].a[0] == ":" and e[1].a[1].kind == "unop"
    assert e[2].kind == "unop" and e[2].a[1].a[0] == "=="
    assert e[3].a[2].kind == "binop" and e[3].a[2].a[0] == "<-"
    assert e[4].a[0] == "+" and e[5].kind == "at"
    two = r_parse.parse("f <- function(x) {\n  y <- x\n  + 1\n}\n")   # a leading operator starts a new expression
    assert len(two[0].a[2].a[1].a[0]) == 2
    assert len(r_parse.parse("y <- x +\n  1\n")) == 1                   # a trailing operator continues the line
    broken = {
        "missing )": wr.replace("state <- new.env(parent = emptyenv())", "state <- new.env(parent = emptyenv()", 1),
        "missing ,": wr.replace("list(n_genes = nrow(matCounts), n_cells", "list(n_genes = nrow(matCounts) n_cells", 1),
        "dangling operator": wr.replace("ends <- cumsum(vecNBatches);", "ends <- cumsum(vecNBatches) + ;", 1),
        "else after a closed top-level if": "f <- function(x) {\n if (x) 1\n}\nelse 2\n",
        "chained comparison": "a < b < c\n", "unclosed {": "f <- function(x) {\n x\n", "two expressions": "a b\n",
        "formals without ,": "f <- function(x y) x\n", "[[ closed by ]": "x[[1]\n", "open string": 'x <- "abc\n',
        "stray )": "f(a, b))\n", "missing operand": "x <- \n", "in outside for": "a in b\n"}
    for what, src in broken.items():
        assert src != wr, what
        with pytest.raises(SyntaxError):
            r_parse.parse(src)


def test_r_wrappers_semantic_checks():
    """What R would only find at run time, found on the syntax tree of r_shim/R/lpb200_wrappers.R (tests/r_parse.py):
    (1) every variable a function reads is bound -- a formal, a local, a top-level object of the file, a function of the
    reference (tests/golden/reference_names.json, extracted from /root/reference/R) or a base-R name; (2) every call of
    a function whose formals are known (the file's own, the reference's) matches them: no unknown or repeated argument
    name, not too many positional ones, no formal without default left unsupplied; the do.call()s of fitLPModels
    included; (3) every x$name / x@name that is read is a name something sets -- the reference, the wrappers, or the
    shim's C code for the lists it returns; (4) S4 slots are slots of the reference's LineagePulseObject."""
    from tests import r_parse
    wr = open(os.path.join(ROOT, "r_shim", "R", "lpb200_wrappers.R")).read()
    csrc = open(os.path.join(ROOT, "r_shim", "lp_rshim.c")).read()
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_names.json")))
    prog = r_parse.parse(wr)
    defs = r_parse.top_level_definitions(prog)
    functions = {k: v for k, v in defs.items() if v.kind == "function"}
    assert len(functions) == 15
    # (1)
    known = set(defs) | set(ref["functions"]) | R_BASE
    for name, fn in functions.items():
        unbound = {v: line for v, line in r_parse.free_variables(fn).items() if v not in known}
        assert not unbound, f"{name}(): unbound variables {unbound}"
    used_ref = {v for fn in functions.values() for v in r_parse.free_variables(fn) if v in ref["functions"] and v not in defs}
    assert used_ref == {"initMuModel", "initDispModel", "initDropModel"}
    # (2)
    formals = {k: [(a, d) for a, d in v] for k, v in ref["functions"].items()}
    formals.update({k: [(a, d is not None) for a, d in fn.a[0]] for k, fn in functions.items()})
    n_checked = 0
    for callee, args, line in r_parse.calls_of(prog):
        if callee in formals:
            problems = r_parse.check_call(callee, args, formals[callee])
            assert not problems, f"line {line}: {problems}"
            n_checked += 1
    assert n_checked >= 40
    fit_lp = functions["fitLPModels"]
    lists = {}
    for n in r_parse.walk(fit_lp):         # do.call(f, c(common, list(...))): the two lists together are the arguments
        if n.kind == "binop" and n.a[0] == "<-" and n.a[1].kind == "name" and n.a[2].kind == "call" \
                and n.a[2].a[0].a[0] == "list":
            lists[n.a[1].a[0]] = n.a[2].a[1]
    n_docall = 0
    for callee, args, line in r_parse.calls_of(fit_lp):
        if callee == "do.call":
            target, packed = args[0][1].a[0], args[1][1]
            assert packed.kind == "call" and packed.a[0].a[0] == "c" and packed.a[1][0][1].a[0] == "common"
            merged = list(lists["common"]) + list(packed.a[1][1][1].a[1])
            problems = r_parse.check_call(target, merged, formals[target])
            assert not problems, f"line {line}: do.call {problems}"
            n_docall += 1
    assert n_docall == 2
    # the wrappers keep every formal of the reference functions they replace, in order, with the same defaults present
    for fn in ("fitMuDisp", "fitPi", "evalLogLikMatrix", "fitModel", "fitLPModels"):
        want = [(a, d) for a, d in ref["functions"][fn]]
        got = formals[fn]
        assert got[: len(want)] == want, (fn, got, want)
    # (3), (4)
    set_somewhere = set(ref["elements"]) | r_parse.element_writes(prog) | set(re.findall(r'"(\w+)"', csrc))
    for kind, name, line in r_parse.element_reads(prog):
        if kind == "at":
            assert name in ref["slots"] or name in ("p", "i", "x"), f"line {line}: @{name} is not a slot of LineagePulseObject / dgCMatrix"
        else:
            assert name in set_somewhere, f"line {line}: ${name} is read but nothing sets it"
    for n in r_parse.walk(prog):
        if n.kind == "binop" and n.a[0] == "<-" and n.a[1].kind == "at":
            assert n.a[1].a[1] in ref["slots"], f"line {n.line}: slot {n.a[1].a[1]} assigned"
    # (5) res <- .Call("routine", ...); res[[k]]: k within the list the C routine allocates
    out_len = {}
    for m in re.finditer(r"^SEXP\s+(lp_\w+)\s*\([^)]*\)\s*\{(.*?)^\}", csrc, re.M | re.S):
        alloc = re.search(r"SEXP out = PROTECT\(allocVector\(VECSXP, (\d+)\)\)", m.group(2))
        if alloc:
            out_len[m.group(1)] = int(alloc.group(1))
    assert out_len == {"lp_fit_mudisp": 3, "lp_fit_pi": 3, "lp_fit_model": 5, "lp_fit_models": 6, "lp_lrt": 2}
    n_indexed = 0
    for name, fn in functions.items():
        results = {}
        for n in r_parse.walk(fn):
            if n.kind == "binop" and n.a[0] == "<-" and n.a[1].kind == "name" and n.a[2].kind == "call" \
                    and n.a[2].a[0].kind == "name" and n.a[2].a[0].a[0] == ".Call" and n.a[2].a[1][0][1].a[0] in out_len:
                results[n.a[1].a[0]] = n.a[2].a[1][0][1].a[0]
        for n in r_parse.walk(fn):
            if n.kind == "index" and n.a[0] == "[[" and n.a[1].kind == "name" and n.a[1].a[0] in results \
                    and n.a[2][0][1].kind == "num":
                k = int(n.a[2][0][1].a[0].rstrip("L"))
                assert 1 <= k <= out_len[results[n.a[1].a[0]]], f"line {n.line}: {n.a[1].a[0]}[[{k}]] beyond {results[n.a[1].a[0]]}'s list"
                n_indexed += 1
    assert n_indexed >= 20
    # the analyses find what they are there for
    bad = r_parse.parse("f <- function(a) { b <- a + cc; g(b, zz = 1) }\ng <- function(x, y) x\n")
    d = r_parse.top_level_definitions(bad)
    assert set(r_parse.free_variables(d["f"])) == {"cc", "g"}
    (callee, args, _), = [c for c in r_parse.calls_of(d["f"]) if c[0] == "g"]
    assert r_parse.check_call("g", args, [("x", False), ("y", False)]) == ["g(): unknown argument 'zz'",
                                                                          "g(): argument 'y' has no default and is not supplied"]


def test_api_signatures_match_the_reference():
    """tests/golden/reference_formals.json holds the formal arguments (names, order, literal defaults) of the reference's
    R functions, extracted from /root/reference/R by tests/golden/make_reference_formals.py.  The Python mirror keeps every
    reference argument, in the reference's order, with the reference's literal defaults (it may add arguments after them:
    device lists, row names, a seed), and so do the R wrappers that replace reference functions."""
    import inspect
    import json
    from tests import r_lint
    from tests.helpers import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "reference_formals.json")))
    mirrored = ["runLineagePulse", "processSCData", "calcNormConst", "fitLPModels", "fitModel", "fitMuDisp", "fitPi",
                "evalLogLikMatrix", "runDEAnalysis", "testDropout", "calcPostDrop_Matrix", "getPostDrop", "getNormData",
                "getFitsMean", "getFitsDispersion", "getFitsDropout", "simulateContinuousDataSet"]
    n_defaults = 0
    for fn in mirrored:
        sig = inspect.signature(getattr(lp, fn)).parameters
        names = [a for a, _ in ref[fn]]
        assert [a for a in sig if a in names] == names, (fn, list(sig), names)      # all there, reference order
        for a, d in ref[fn]:
            if d is None:
                continue
            have = sig[a].default
            want = d.get("str", d.get("num", {"TRUE": True, "FALSE": False, "NULL": None}.get(d.get("lit"))))
            if fn == "fitModel" and a in ("strMuModel", "strDispModel"):
                continue
            assert have == want and (not isinstance(want, bool) or isinstance(have, bool)), (fn, a, have, want)
            n_defaults += 1
    assert n_defaults >= 50
    # the R wrappers that replace reference functions
    wr = open(os.path.join(ROOT, "r_shim", "R", "lpb200_wrappers.R")).read()
    defs = r_lint.function_defs(r_lint.check_nesting(wr))
    for fn in ("fitMuDisp", "fitPi", "evalLogLikMatrix", "fitModel", "fitLPModels"):
        names = [a for a, _ in ref[fn]]
        assert defs[fn][: len(names)] == names, (fn, defs[fn], names)
    # and the reference's init functions are called by the wrappers with exactly their formals
    toks = r_lint.check_nesting(wr)
    for fn in ("initMuModel", "initDispModel", "initDropModel"):
        for line, args in r_lint.calls(toks, fn):
            given = [a[0][1] for a in args if len(a) >= 3 and a[1][:2] == ("op", "=")]
            assert given == [a for a, _ in ref[fn]], (fn, line, given)


def test_error_messages_are_the_reference_s():
    """tests/golden/reference_messages.json: the string literals of every stop() / warning() of the reference (extracted by
    tests/golden/make_reference_formals.py).  The host mirror raises the reference's own texts for the conditions it checks:
    every literal fragment (>= 12 characters) of the messages listed here occurs verbatim in the product sources."""
    import json
    from tests.helpers import GOLDEN
    msgs = json.load(open(os.path.join(GOLDEN, "reference_messages.json")))
    src = "".join(open(os.path.join(ROOT, "lineagepulse_b200", f)).read() for f in ("api.py", "simulate.py", os.path.join("csrc", "lp_api.cu")))
    src = re.sub(r'"\s*\n\s*f?"', "", src)     # adjacent string literals of one message
    mirrored = {  # file -> lines of the stop() / warning() calls the mirror reproduces
        "srcLineagePulse_calcNormConst.R": [32],
        "srcLineagePulse_evalLogLik.R": [271, 319],
        "srcLineagePulse_fitDropout.R": [457, 501],
        "srcLineagePulse_fitMeanDispersion.R": [804],
        "srcLineagePulse_getFits.R": [60, 64, 140, 202, 268, 349],
        "srcLineagePulse_initialiseMuModel.R": [128],
        "srcLineagePulse_initialiseDispModel.R": [105],
        "srcLineagePulse_processSCData.R": [107, 117, 124, 134, 139, 144, 152, 167, 182, 188, 194, 202, 217, 232, 236, 240, 244],
        "srcLineagePulse_simulateDataSet.R": [95, 101, 107, 221, 260, 281],
    }
    by_key = {(m["file"], m["line"]): m for m in msgs}
    n = 0
    for f, lines in mirrored.items():
        for line in lines:
            m = by_key[(f, line)]
            for lit in m["literals"]:
                frag = lit.strip()
                if len(frag) >= 12:
                    # (messages shared by several arguments are formatted with the argument's name: {nm})
                    templ = re.sub(r"vecConfoundersMu|strDispModelFull|strDispModelRed|vecNormConstExternal|vecDispExternal|"
                                   r"vecGeneWiseDropoutRates", "\x00", frag)
                    pat = r"\{\w+\}".join(re.escape(part) for part in templ.split("\x00"))
                    assert frag in src or re.search(pat, src), (f, line, frag)
                    n += 1
    assert n >= 45
    # a few of them raised for real (no GPU needed)
    X = np.ones((3, 4), dtype=np.int32)
    for kw, text in ((dict(strMuModel="foo", strDispModel="constant"), "ERROR fitZINB(): strMuModel=foo not recognised."),
                     (dict(strMuModel="constant", strDispModel="bar"), "ERROR fitZINB(): strDispModel=bar not recognised."),
                     (dict(strMuModel="MM", strDispModel="constant"), "fitMMZINB() not yet supported")):
        with pytest.raises(lp.LineagePulseError, match=re.escape(text)):
            lp.fitModel(X, {"cell": np.arange(4)}, **kw)


def test_return_names_are_the_reference_s():
    """tests/golden/reference_returns.json: the element names of the lists the reference's boundary functions return and the
    columns of its result tables (extracted from /root/reference/R by tests/golden/make_reference_formals.py).  The dict /
    DataFrame literals of the Python mirror carry the same names in the same order (read off the source: the functions
    themselves need a GPU), and the R wrappers return lists with those names."""
    import ast
    import json
    from tests import r_lint
    from tests.helpers import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "reference_returns.json")))
    tree = ast.parse(open(os.path.join(ROOT, "lineagepulse_b200", "api.py")).read())
    funcs = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}

    def dict_keys(fn, pick):
        """String keys of the dict literals inside function ``fn``; the one ``pick`` selects."""
        ds = [[k.value for k in d.keys if isinstance(k, ast.Constant)] for d in ast.walk(funcs[fn]) if isinstance(d, ast.Dict)]
        return pick(ds)

    first_with = lambda key: (lambda ds: next(d for d in ds if key in d))   # noqa: E731
    assert dict_keys("fitMuDisp", first_with("vecConvergence")) == ref["fitMuDisp"]
    assert dict_keys("fitPi", first_with("vecConverged")) == ref["fitPi"]
    assert dict_keys("_fit_result", first_with("boolConvergenceModel"))[: len(ref["fitModel"])] == ref["fitModel"]
    assert dict_keys("processSCData", first_with("objLP")) == ref["processSCData"]
    cols = [c for c in ref["runDEAnalysis"] if c != "stringsAsFactors"]
    assert dict_keys("runDEAnalysis", first_with("padj_nb")) == cols
    assert 'df["allZero"]' in ast.get_source_segment(open(os.path.join(ROOT, "lineagepulse_b200", "api.py")).read(), funcs["runDEAnalysis"])
    sim = ast.parse(open(os.path.join(ROOT, "lineagepulse_b200", "simulate.py")).read())
    sim_keys = [[k.value for k in d.keys if isinstance(k, ast.Constant)] for d in ast.walk(sim) if isinstance(d, ast.Dict)]
    assert any(set(ref["simulateContinuousDataSet"]) <= set(d) for d in sim_keys)
    # the R wrappers
    wr = open(os.path.join(ROOT, "r_shim", "R", "lpb200_wrappers.R")).read()
    toks = r_lint.check_nesting(wr)
    lists = [[a[0][1] for a in args if len(a) >= 3 and a[1][:2] == ("op", "=")] for _, args in r_lint.calls(toks, "list")]
    for fn in ("fitMuDisp", "fitPi", "fitModel"):
        assert ref[fn] in lists, (fn, ref[fn])
