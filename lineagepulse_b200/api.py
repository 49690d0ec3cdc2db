"""Host-side mirror of the reference's operator interface for the fitting path.

The reference is an R package; R is not available in this build image, so the host side above
the C ABI is written in Python with the reference's own function names, argument names,
defaults and error behaviour.  Every function cites the R function it mirrors
(paths relative to /root/reference/R/, prefix srcLineagePulse_ dropped).  All fitting /
likelihood arithmetic happens in the CUDA library; this module only marshals.

Model containers are dicts with the reference's field names:
  lsMuModel   = {matMuModel, lsmatBatchModel, lsMuModelGlobal{...}}
  lsDispModel = {matDispModel, lsmatBatchModel, lsDispModelGlobal{...}}
  lsDropModel = {matDropoutLinModel, matPiConstPredictors, lsDropModelGlobal{...}}
Group / batch index vectors are 0-based here (R: 1-based).
"""
from __future__ import annotations

import sys
import time
import warnings

import numpy as np

from . import _cabi as A
from .engine import Engine, lrt
from .splines import ns

__version__ = "0.99.20-b200"

_MU = A.MU_MODELS
_DISP = A.DISP_MODELS
_DROP = A.DROP_MODELS
_GROUP = A.DROP_FIT_GROUPS


class LineagePulseError(Exception):
    """R's stop()."""


def message(*args):
    """R's message(): diagnostics go to stderr, never to stdout."""
    print(*args, file=sys.stderr)


# ---------------------------------------------------------------------------------------------
# count matrix handle: host integers + the GPU context that keeps them resident
# ---------------------------------------------------------------------------------------------
class CountMatrix:
    """matCountsProc (a dgCMatrix in the reference, classLineagePulseObject.R:105): genes x cells.

    ``data`` may be a dense array (NaN = NA) or a scipy.sparse matrix.  The device copy is created
    on first use and stays resident for every later fit on this object.
    """

    def __init__(self, data, rownames=None, colnames=None, device=0, devices=None):
        import scipy.sparse as sp
        self.sparse = sp.issparse(data)
        self.data = data.tocsc() if self.sparse else np.asarray(data)
        self.shape = self.data.shape
        G, N = self.shape
        self.rownames = np.asarray(rownames if rownames is not None else [f"Gene_{i + 1}" for i in range(G)])
        self.colnames = np.asarray(colnames if colnames is not None else [f"cell_{j + 1}" for j in range(N)])
        self.device = device
        self.devices = None if devices is None else [int(v) for v in devices]   # several GPUs: genes are dealt to them
        self._engine = None

    def engine(self) -> Engine:
        if self._engine is None:
            e = Engine(self.device, devices=self.devices)
            G, N = self.shape
            if self.sparse:
                m = self.data
                e.load_counts_csc(G, N, m.indptr, m.indices, m.data)
            elif np.issubdtype(self.data.dtype, np.integer):
                e.load_counts_i32(self.data)
            else:
                e.load_counts_dense(self.data)
            self._engine = e
        return self._engine

    def dense(self):
        return np.asarray(self.data.todense()) if self.sparse else self.data

    def row_means(self):
        return self.engine().row_means()

    def subset_genes(self, idx):
        d = self.data[idx, :] if not self.sparse else self.data.tocsr()[idx, :].tocsc()
        return CountMatrix(d, self.rownames[idx], self.colnames, self.device, self.devices)


def _as_counts(matCounts) -> CountMatrix:
    return matCounts if isinstance(matCounts, CountMatrix) else CountMatrix(matCounts)


# ---------------------------------------------------------------------------------------------
# LineagePulseObject (classLineagePulseObject.R:88-119) with list-like access (:622-712)
# ---------------------------------------------------------------------------------------------
_SLOTS = ["dfAnnotationProc", "dfResults", "lsDispModelH0", "lsDispModelH1", "lsDispModelConst",
          "lsDispModelH0_NB", "lsDispModelH1_NB", "lsDropModel", "lsMuModelH0", "lsMuModelH1",
          "lsMuModelConst", "lsMuModelH0_NB", "lsMuModelH1_NB", "lsFitConvergence", "matCountsProc",
          "matWeights", "scaDFSplinesDisp", "scaDFSplinesMu", "strReport", "vecAllGenes",
          "vecConfoundersDisp", "vecConfoundersMu", "scaOmega", "boolFixedPopulations",
          "vecNCentroidsPerPop", "vecH0Pop", "vecNormConst", "strVersion"]


class LineagePulseObject:
    def __init__(self, **kw):
        for s in _SLOTS:
            setattr(self, s, kw.get(s))
        if self.boolFixedPopulations is None:
            self.boolFixedPopulations = False
        if self.strVersion is None:
            self.strVersion = __version__

    def names(self):
        return list(_SLOTS)

    def __getitem__(self, name):
        if name not in _SLOTS:
            raise KeyError(name)
        return getattr(self, name)


def _accessor(name):
    def get(objLP):
        return getattr(objLP, name)
    get.__name__ = name
    get.__doc__ = f"Accessor of slot {name} (classLineagePulseObject.R:209-345)."
    return get


for _s in _SLOTS:
    globals()[_s] = _accessor(_s)


def writeReport(objLP, file=""):
    """classLineagePulseObject.R:748."""
    if file:
        with open(file, "w") as fh:
            fh.write(objLP.strReport + "\n")
    else:
        print(objLP.strReport)


# ---------------------------------------------------------------------------------------------
# model <-> C ABI marshalling
# ---------------------------------------------------------------------------------------------
def _model_struct(lsMuModel, lsDispModel, lsDropModel) -> A.Model:
    strMu = lsMuModel["lsMuModelGlobal"]["strMuModel"]
    strDisp = lsDispModel["lsDispModelGlobal"]["strDispModel"]
    g = (lsDropModel or {}).get("lsDropModelGlobal", {"strDropModel": "none"})
    strDrop = g.get("strDropModel", "none")
    if strMu not in _MU:
        raise LineagePulseError(f"ERROR decompressMeans(): lsMuModelGlobal$strMuModel={strMu} not recognised.")
    if strDisp not in _DISP:
        raise LineagePulseError(f"ERROR decompressDispersions(): lsDispModelGlobal$strDispModel={strDisp} not recognised.")
    if strDrop not in _DROP:
        raise LineagePulseError(f"lsDropModelGlobal$strDropModel={strDrop} not recognised.")
    strGroup = g.get("strDropFitGroup", "PerCell")
    if strDrop != "none" and strGroup not in _GROUP:
        raise LineagePulseError(f"ERROR fitZINBPi: lsDropModel$lsDropModelGlobal$strDropFitGroup={strGroup} not recognised.")
    return A.Model(_MU[strMu], _DISP[strDisp], _DROP[strDrop], _GROUP.get(strGroup, 0))


def _design(counts: CountMatrix, lsMuModel, lsDispModel, lsDropModel) -> A.DesignData:
    gm = lsMuModel["lsMuModelGlobal"]
    gd = lsDispModel["lsDispModelGlobal"]
    G, N = counts.shape
    groups = gm.get("vecidxGroups")
    if groups is None:
        groups = gd.get("vecidxGroups")
    t = gm.get("vecContinuousCovar")
    if t is None:
        t = gd.get("vecPseudotime")
    impulse_basis = None
    if gm["strMuModel"] == "impulse":
        impulse_basis = gm.get("_matImpulseBasis")
        if impulse_basis is None:
            impulse_basis = ns(t, 4, intercept=True)  # initialiseImpulseParameters.R:34-35
            gm["_matImpulseBasis"] = impulse_basis
    pred = (lsDropModel or {}).get("matPiConstPredictors")
    return A.DesignData(
        G, N, size_factors=gm.get("vecNormConst"), pseudotime=t, groups=groups,
        spline_basis_mu=gm.get("matSplineBasis"), spline_basis_disp=gd.get("matSplineBasis"),
        impulse_basis=impulse_basis,
        confounders_mu=gm.get("lsvecidxBatchAssign"), confounders_disp=gd.get("lsvecidxBatchAssign"),
        pi_const_pred=pred)


def _prepare(matCounts, lsMuModel, lsDispModel, lsDropModel):
    counts = _as_counts(matCounts)
    model = _model_struct(lsMuModel, lsDispModel, lsDropModel)
    if model.mu_model == A.MU_MM:
        raise LineagePulseError("LineagePulse ERROR: fitMMZINB() not yet supported. Contact developer.")
    design = _design(counts, lsMuModel, lsDispModel, lsDropModel)
    eng = counts.engine()
    eng.set_design(design)
    return counts, eng, design, model


def _params_from_models(design, model, lsMuModel, lsDispModel, lsDropModel) -> A.ParamsData:
    """Copies: the library updates parameter matrices in place, R functions have value semantics -- the caller's
    lsMuModel / lsDispModel / lsDropModel stay as they were."""
    def own(a, rows):
        return np.array(np.asarray(a, dtype=np.float64).reshape(rows, -1), dtype=np.float64, order="F", copy=True)
    p = A.ParamsData(design, model)
    p.mat_mu = own(lsMuModel["matMuModel"], design.n_genes)
    p.mat_disp = own(lsDispModel["matDispModel"], design.n_genes)
    if design.n_batches_mu:
        p.batch_mu = [own(b, design.n_genes) for b in lsMuModel["lsmatBatchModel"]]
    if design.n_batches_disp:
        p.batch_disp = [own(b, design.n_genes) for b in lsDispModel["lsmatBatchModel"]]
    if model.drop_model != A.DROP_NONE:
        p.mat_drop = own(lsDropModel["matDropoutLinModel"], design.n_cells)
    return p


def _control_from_models(lsMuModel=None, lsDropModel=None, control=None) -> A.Control:
    """The optimiser constants of a boundary call.  The reference reads them from the model lists --
    lsMuModelGlobal$MAXIT_BFGS_MuDisp / RELTOL_BFGS_MuDisp (fitMeanDispersion.R:912-913) and
    lsDropModelGlobal$MAXIT_BFGS_Pi / RELTOL_BFGS_Pi (fitDropout.R:291-292, :364-365) -- where fitModel put its
    constants (fitModel.R:164-175); a list without them (hand-built models) gets those constants.  An explicit
    ``control`` wins (testing)."""
    if control is not None:
        return control
    ctl = A.default_control()
    gm = (lsMuModel or {}).get("lsMuModelGlobal") or {}
    gp = (lsDropModel or {}).get("lsDropModelGlobal") or {}
    if gm.get("MAXIT_BFGS_MuDisp") is not None:
        ctl.maxit_mudisp = int(gm["MAXIT_BFGS_MuDisp"])
    if gm.get("RELTOL_BFGS_MuDisp") is not None:
        ctl.reltol_mudisp = float(gm["RELTOL_BFGS_MuDisp"])
    if gp.get("MAXIT_BFGS_Pi") is not None:
        ctl.maxit_pi = int(gp["MAXIT_BFGS_Pi"])
    if gp.get("RELTOL_BFGS_Pi") is not None:
        ctl.reltol_pi = float(gp["RELTOL_BFGS_Pi"])
    return ctl


# ---------------------------------------------------------------------------------------------
# L2 boundary functions
# ---------------------------------------------------------------------------------------------
def evalLogLikMatrix(matCounts, lsMuModel, lsDispModel, lsDropModel, matWeights=None,
                     boolConstModelOnMMLL=False):
    """evalLogLik.R:307-431: per-gene log-likelihood of the stored models."""
    if boolConstModelOnMMLL and (matWeights is None or lsMuModel["lsMuModelGlobal"]["strMuModel"] != "constant"):
        raise LineagePulseError("boolConstModelOnMMLL usage wrong in evalLogLikMatrix().")
    if boolConstModelOnMMLL or lsMuModel["lsMuModelGlobal"]["strMuModel"] == "MM":
        raise LineagePulseError("LineagePulse ERROR: evalLogLikGeneMM() not yet supported. Contact developer.")
    counts, eng, design, model = _prepare(matCounts, lsMuModel, lsDispModel, lsDropModel)
    p = _params_from_models(design, model, lsMuModel, lsDispModel, lsDropModel)
    return eng.eval_loglik(model, p)


def fitMuDisp(matCountsProc, vecNormConst=None, lsMuModel=None, lsDispModel=None, lsDropModel=None,
              matWeights=None, control=None):
    """fitMeanDispersion.R:853-1039.  ``vecNormConst`` is accepted and ignored exactly as in the
    reference (size factors travel inside lsMuModelGlobal, initialiseMuModel.R:73)."""
    counts, eng, design, model = _prepare(matCountsProc, lsMuModel, lsDispModel, lsDropModel)
    p = _params_from_models(design, model, lsMuModel, lsDispModel, lsDropModel)
    r = eng.fit_mudisp(model, p, _control_from_models(lsMuModel, lsDropModel, control))
    return {"matMuModel": p.mat_mu, "lsmatBatchModelMu": p.batch_mu or None, "matDispModel": p.mat_disp,
            "lsmatBatchModelDisp": p.batch_disp or None, "vecConvergence": r["conv"], "vecLL": r["ll"]}


def fitPi(matCounts, lsMuModel, lsDispModel, lsDropModel, control=None):
    """fitDropout.R:418-515."""
    counts, eng, design, model = _prepare(matCounts, lsMuModel, lsDispModel, lsDropModel)
    if model.drop_model == A.DROP_NONE:
        raise LineagePulseError("ERROR fitZINBPi: lsDropModel$lsDropModelGlobal$strDropModel=none not recognised.")
    p = _params_from_models(design, model, lsMuModel, lsDispModel, lsDropModel)
    r = eng.fit_pi(model, p, _control_from_models(lsMuModel, lsDropModel, control))
    return {"matDropoutLinModel": p.mat_drop, "vecConverged": r["conv"], "vecLL": r["ll"]}


def calcPostDrop_Matrix(matCounts, lsMuModel, lsDispModel, lsDropModel, vecIDs=None):
    """calcPostDrop.R:69-125 (mu enters WITHOUT size factors, :110)."""
    counts, eng, design, model = _prepare(matCounts, lsMuModel, lsDispModel, lsDropModel)
    p = _params_from_models(design, model, lsMuModel, lsDispModel, lsDropModel)
    ids = np.arange(counts.shape[0]) if vecIDs is None else _gene_indices(counts, vecIDs)
    return eng.post_drop(model, p, ids)


def _gene_indices(counts, vecIDs):
    ids = np.asarray(vecIDs)
    if np.issubdtype(ids.dtype, np.integer):
        return ids
    lookup = {g: i for i, g in enumerate(counts.rownames)}
    return np.array([lookup[g] for g in ids], dtype=np.int64)


# ---------------------------------------------------------------------------------------------
# model initialisation (metadata; the values come from the library)
# ---------------------------------------------------------------------------------------------
def _confounder_meta(dfAnnotation, vecConfounders):
    if not vecConfounders:
        return {"vecConfounders": None, "scaNConfounders": None, "vecNBatches": None,
                "lsvecidxBatchAssign": None, "lsvecBatchUnique": None}
    uniq, idx = [], []
    for cname in vecConfounders:
        lab, ix = A.first_appearance_index(np.asarray(dfAnnotation[cname]))
        uniq.append(lab)
        idx.append(ix.astype(np.int32))
    return {"vecConfounders": list(vecConfounders), "scaNConfounders": len(vecConfounders),
            "vecNBatches": [len(u) for u in uniq], "lsvecidxBatchAssign": idx, "lsvecBatchUnique": uniq}


def _global_models(counts, dfAnnotation, vecConfoundersMu, vecConfoundersDisp, vecNormConst, scaDFSplinesMu,
                   scaDFSplinesDisp, strMuModel, strDispModel, strDropModel, strDropFitGroup,
                   matPiConstPredictors, MAXIT_BFGS_MuDisp, RELTOL_BFGS_MuDisp, MAXIT_BFGS_Pi, RELTOL_BFGS_Pi):
    """The ...Global metadata of initMuModel / initDispModel / initDropModel
    (initialiseMuModel.R:67-91,134-203; initialiseDispModel.R:71-88,132-176; initialiseDropModel.R:61-70)."""
    G, N = counts.shape
    t = None if dfAnnotation is None or "continuous" not in dfAnnotation else np.asarray(dfAnnotation["continuous"], dtype=np.float64)
    gm = {"scaDegFreedom": None, "strMuModel": strMuModel, "vecNormConst": np.asarray(vecNormConst, dtype=np.float64),
          "scaNumCells": N, "vecContinuousCovar": t, "scaNSplines": None, "matSplineBasis": None,
          "vecidxGroups": None, "vecGroups": None, "scaNGroups": None, "scaNMix": None,
          "MAXIT_BFGS_MuDisp": MAXIT_BFGS_MuDisp, "RELTOL_BFGS_MuDisp": RELTOL_BFGS_MuDisp}
    gm.update(_confounder_meta(dfAnnotation, vecConfoundersMu))
    gd = {"scaDegFreedom": None, "strDispModel": strDispModel, "scaNumCells": N, "vecPseudotime": t,
          "scaNSplines": None, "matSplineBasis": None, "vecidxGroups": None, "vecGroups": None,
          "scaNGroups": None, "scaNMix": None}
    gd.update(_confounder_meta(dfAnnotation, vecConfoundersDisp))
    if strMuModel == "groups" or strDispModel == "groups":
        lab, idx = A.first_appearance_index(np.asarray(dfAnnotation["groups"]))
        for g, used in ((gm, strMuModel == "groups"), (gd, strDispModel == "groups")):
            if used:
                g["vecGroups"], g["vecidxGroups"], g["scaNGroups"] = lab, idx.astype(np.int32), len(lab)
    if strMuModel == "splines":
        gm["scaNSplines"] = scaDFSplinesMu
        gm["matSplineBasis"] = ns(t, scaDFSplinesMu, intercept=True)
    if strDispModel == "splines":
        gd["scaNSplines"] = scaDFSplinesDisp
        gd["matSplineBasis"] = ns(t, scaDFSplinesDisp, intercept=True)
    if strMuModel in ("splines", "impulse"):
        gm["vecTimepoints"] = np.unique(t)
        gm["vecindTimepointAssign"] = np.searchsorted(gm["vecTimepoints"], t)
    gdrop = {"strDropModel": strDropModel, "strDropFitGroup": strDropFitGroup, "scaNumGenes": G,
             "scaNumCells": N, "MAXIT_BFGS_Pi": MAXIT_BFGS_Pi, "RELTOL_BFGS_Pi": RELTOL_BFGS_Pi}
    lsMu = {"matMuModel": None, "lsmatBatchModel": None, "lsMuModelGlobal": gm}
    lsDisp = {"matDispModel": None, "lsmatBatchModel": None, "lsDispModelGlobal": gd}
    lsDrop = {"matDropoutLinModel": None,
              "matPiConstPredictors": None if matPiConstPredictors is None else np.asarray(matPiConstPredictors, dtype=np.float64),
              "lsDropModelGlobal": gdrop}
    return lsMu, lsDisp, lsDrop


def _fill_models(design, model, p: A.ParamsData, lsMu, lsDisp, lsDrop, fill_drop):
    lsMu["matMuModel"] = p.mat_mu
    lsDisp["matDispModel"] = p.mat_disp
    # no confounders: absent after the first fitMuDisp (fitModel.R:303,305)
    lsMu["lsmatBatchModel"] = p.batch_mu or None
    lsDisp["lsmatBatchModel"] = p.batch_disp or None
    lsMu["lsMuModelGlobal"]["scaDegFreedom"] = (design.n_mu_params(model) + sum(design.n_batches_mu) - len(design.n_batches_mu))
    lsDisp["lsDispModelGlobal"]["scaDegFreedom"] = (design.n_disp_params(model) + sum(design.n_batches_disp)
                                                   - len(design.n_batches_disp))
    if fill_drop:
        lsDrop["matDropoutLinModel"] = p.mat_drop


# ---------------------------------------------------------------------------------------------
# fitModel: the EM loop runs inside the library (fused fast path)
# ---------------------------------------------------------------------------------------------
def fitModel(matCounts, dfAnnotation, vecConfoundersMu=None, vecConfoundersDisp=None, vecNormConst=None,
             scaDFSplinesMu=None, scaDFSplinesDisp=None, matWeights=None, matPiConstPredictors=None,
             lsDropModel=None, matMuModelInit=None, lsmatBatchModelInitMu=None, matDispModelInit=None,
             lsmatBatchModelInitDisp=None, strMuModel=None, strDispModel=None, strDropModel="logistic_ofMu",
             strDropFitGroup="PerCell", scaMaxEstimationCycles=20, boolVerbose=True, boolSuperVerbose=True,
             min_rowmean_global=float("nan")):
    """fitModel.R:137-356.  Returns the reference's list: lsMuModel, lsDispModel, lsDropModel,
    boolConvergenceModel, vecEMLogLikModel, strReport."""
    counts = _as_counts(matCounts)
    G, N = counts.shape
    if vecNormConst is None:
        vecNormConst = np.ones(N)
    # constants of fitModel.R:164-175
    ctl = A.default_control()
    ctl.max_cycles = int(scaMaxEstimationCycles)
    boolFitDrop = lsDropModel is None and strDropModel != "none"
    if strMuModel == "MM":
        raise LineagePulseError("LineagePulse ERROR: fitMMZINB() not yet supported. Contact developer.")
    if strMuModel not in _MU:        # initialiseMuModel.R:127-130
        raise LineagePulseError(f"ERROR fitZINB(): strMuModel={strMuModel} not recognised.")
    if strDispModel not in _DISP:    # initialiseDispModel.R:104-107
        raise LineagePulseError(f"ERROR fitZINB(): strDispModel={strDispModel} not recognised.")
    lsMu, lsDisp, lsDropNew = _global_models(
        counts, dfAnnotation, vecConfoundersMu, vecConfoundersDisp, vecNormConst, scaDFSplinesMu, scaDFSplinesDisp,
        strMuModel, strDispModel, strDropModel, strDropFitGroup, matPiConstPredictors,
        ctl.maxit_mudisp, ctl.reltol_mudisp, ctl.maxit_pi, ctl.reltol_pi)
    if lsDropModel is None:
        lsDrop = lsDropNew
    else:
        lsDrop = lsDropModel  # supplied model is kept as is (initialiseDropModel.R:60)
    counts2, eng, design, model = _prepare(counts, lsMu, lsDisp, lsDrop)
    p = A.ParamsData(design, model)
    if not boolFitDrop and model.drop_model != A.DROP_NONE:
        p.mat_drop = np.asfortranarray(lsDrop["matDropoutLinModel"], dtype=np.float64)
    # warm-start hooks (fitModel.R:147-151; initialiseMuModel.R:93,131-133; initialiseDispModel.R:89,109-129)
    keep = 0
    if matMuModelInit is not None:
        p.mat_mu = np.array(np.asarray(matMuModelInit, dtype=np.float64).reshape(G, -1), order="F", copy=True)
        keep |= A.KEEP_MU
    if lsmatBatchModelInitMu is not None and design.n_batches_mu:
        p.batch_mu = [np.array(b, dtype=np.float64, order="F", copy=True) for b in lsmatBatchModelInitMu]
        keep |= A.KEEP_BATCH_MU
    if matDispModelInit is not None:
        p.mat_disp = np.array(np.asarray(matDispModelInit, dtype=np.float64).reshape(G, -1), order="F", copy=True)
        keep |= A.KEEP_DISP
    if lsmatBatchModelInitDisp is not None and design.n_batches_disp:
        p.batch_disp = [np.array(b, dtype=np.float64, order="F", copy=True) for b in lsmatBatchModelInitDisp]
        keep |= A.KEEP_BATCH_DISP
    for nm, a, want in (("matMuModelInit", p.mat_mu, design.n_mu_params(model)), ("matDispModelInit", p.mat_disp, design.n_disp_params(model))):
        if a.shape != (G, want):
            raise LineagePulseError(f"{nm} must be genes x model parameters ({G} x {want}), got {a.shape}")
    t0 = time.time()
    r = eng.fit_model(model, p, boolFitDrop, ctl, min_rowmean_global, keep_mask=keep)
    dt = time.time() - t0
    _fill_models(design, model, p, lsMu, lsDisp, lsDrop, fill_drop=boolFitDrop)
    return _fit_result(lsMu, lsDisp, lsDrop, r, ctl.max_cycles if boolFitDrop else 1, dt, boolVerbose)


def _union(*vals):
    for v in vals:
        if v is not None:
            return v
    return None


def fitModelsConcurrently(matCounts, dfAnnotation, lsSpecs, vecConfoundersMu=None, vecConfoundersDisp=None,
                          vecNormConst=None, scaDFSplinesMu=None, scaDFSplinesDisp=None, boolVerbose=True,
                          min_rowmean_global=float("nan")):
    """Several fitModel() calls that do not estimate the drop-out model (fits B, C, D of fitLPModels,
    fitWrapperLP.R:152-238) in one library call: identical results to sequential fitModel() calls, but
    the fit kernels overlap on the GPU (lpb200_fit_model_multi).  ``lsSpecs`` = list of dicts with
    strMuModel, strDispModel, strDropModel and (unless "none") lsDropModel."""
    counts = _as_counts(matCounts)
    G, N = counts.shape
    if vecNormConst is None:
        vecNormConst = np.ones(N)
    ctl = A.default_control()
    built = []
    for spec in lsSpecs:
        if spec["strMuModel"] == "MM":
            raise LineagePulseError("LineagePulse ERROR: fitMMZINB() not yet supported. Contact developer.")
        lsDropModel = spec.get("lsDropModel")
        strDrop = spec.get("strDropModel", "none")
        if strDrop != "none" and lsDropModel is None:
            raise LineagePulseError("fitModelsConcurrently(): a drop-out model must be supplied (use fitModel() to estimate it)")
        lsMu, lsDisp, lsDropNew = _global_models(
            counts, dfAnnotation, vecConfoundersMu, vecConfoundersDisp, vecNormConst, scaDFSplinesMu, scaDFSplinesDisp,
            spec["strMuModel"], spec["strDispModel"], strDrop, spec.get("strDropFitGroup", "PerCell"), None,
            ctl.maxit_mudisp, ctl.reltol_mudisp, ctl.maxit_pi, ctl.reltol_pi)
        built.append((lsMu, lsDisp, lsDropModel if lsDropModel is not None else lsDropNew))
    # one design holding every covariate any of the models needs
    gms = [b[0]["lsMuModelGlobal"] for b in built]
    gds = [b[1]["lsDispModelGlobal"] for b in built]
    t = _union(*[g.get("vecContinuousCovar") for g in gms])
    need_impulse = any(g["strMuModel"] == "impulse" for g in gms)
    design = A.DesignData(
        G, N, size_factors=vecNormConst, pseudotime=t,
        groups=_union(*[g.get("vecidxGroups") for g in gms + gds]),
        spline_basis_mu=_union(*[g.get("matSplineBasis") for g in gms]),
        spline_basis_disp=_union(*[g.get("matSplineBasis") for g in gds]),
        impulse_basis=ns(t, 4, intercept=True) if need_impulse else None,
        confounders_mu=_union(*[g.get("lsvecidxBatchAssign") for g in gms]),
        confounders_disp=_union(*[g.get("lsvecidxBatchAssign") for g in gds]),
        pi_const_pred=_union(*[(b[2] or {}).get("matPiConstPredictors") for b in built]))
    eng = counts.engine()
    eng.set_design(design)
    jobs = []
    for lsMu, lsDisp, lsDrop in built:
        model = _model_struct(lsMu, lsDisp, lsDrop)
        p = A.ParamsData(design, model)
        if model.drop_model != A.DROP_NONE:
            p.mat_drop = np.asfortranarray(lsDrop["matDropoutLinModel"], dtype=np.float64)
        jobs.append((model, p))
    t0 = time.time()
    res = eng.fit_model_multi(jobs, ctl, min_rowmean_global)
    dt = time.time() - t0
    out = []
    for (lsMu, lsDisp, lsDrop), (model, p), r in zip(built, jobs, res):
        _fill_models(design, model, p, lsMu, lsDisp, lsDrop, fill_drop=False)
        out.append(_fit_result(lsMu, lsDisp, lsDrop, r, 1, dt / len(jobs), boolVerbose))
    return out


def _fit_result(lsMu, lsDisp, lsDrop, r, ncyc, dt, boolVerbose):
    strReport = f"#  .   Initialisation: ll {r['ll_init']}\n"
    for k in range(r["n_iter"]):
        strReport += f"# {k + 1}.  Iteration with ll   {r['em_ll'][k]}\n"
    n_bad = int((r["conv"] != 0).sum())
    if n_bad:
        codes = ",".join(str(c) for c in np.unique(r["conv"][r["conv"] != 0]))
        strReport += f"(Mean-) Dispersion estimation did not converge in {n_bad} cases [codes: {codes}].\n"
    if boolVerbose:
        message(strReport.rstrip("\n"))
    return {"lsMuModel": lsMu, "lsDispModel": lsDisp, "lsDropModel": lsDrop,
            "boolConvergenceModel": r["converged"], "vecEMLogLikModel": r["em_ll"][:ncyc],
            "strReport": strReport, "vecLL": r["ll"], "vecConvergence": r["conv"], "scaTime": dt}


# ---------------------------------------------------------------------------------------------
# processSCData / calcNormConst / fitLPModels / runDEAnalysis / runLineagePulse
# ---------------------------------------------------------------------------------------------
def processSCData(counts, dfAnnotation, vecConfoundersDisp=None, vecConfoundersMu=None, matPiConstPredictors=None,
                  vecNormConstExternal=None, strDispModelFull="constant", strDispModelRed="constant",
                  strMuModel="splines", scaDFSplinesDisp=3, scaDFSplinesMu=6, scaMaxEstimationCycles=20,
                  boolVerbose=True, boolSuperVerbose=False, rownames=None, colnames=None, device=0, devices=None):
    """processSCData.R:85-365: input checks, cell / gene filtering, object construction."""
    import scipy.sparse as sp
    if isinstance(counts, str):
        if not counts.endswith(".mtx"):
            raise LineagePulseError("ERROR IN INPUT DATA: Input matrix file is not .mtx. "
                                    "Supply counts as .mtx file or as R matrix object.")
        import scipy.io
        try:
            counts = sp.csc_matrix(scipy.io.mmread(counts, spmatrix=True))
        except TypeError:      # scipy without the spmatrix switch
            counts = sp.csc_matrix(scipy.io.mmread(counts))
    if counts is None:
        raise LineagePulseError("ERROR: counts was not given as input.")
    if dfAnnotation is None:
        raise LineagePulseError("Supply dfAnnotation.")
    sparse = sp.issparse(counts)
    device_checks = False
    vals = counts.data if sparse else np.asarray(counts)
    if not np.issubdtype(vals.dtype, np.number):
        raise LineagePulseError("ERROR: counts contains non-numeric elements.")
    if not np.issubdtype(vals.dtype, np.integer):
        fin = vals[~np.isnan(vals)]
        if np.any(np.isinf(fin)):
            raise LineagePulseError("ERROR: counts contains infinite elements. Requires count data.")
        if np.any(fin % 1 != 0):
            raise LineagePulseError("ERROR: counts contains non-integer elements. Requires count data.")
        if np.any(fin < 0):
            raise LineagePulseError("ERROR: counts contains negative elements. Requires count data.")
    else:
        # dense int32 input without a cell filter: the negative check and the all-zero gene / cell filters are read off
        # the uploaded matrix (lpb200_count_flags) instead of three host passes over G x N
        device_checks = (not sparse and isinstance(counts, np.ndarray) and counts.dtype == np.int32 and counts.flags.c_contiguous
                         and counts.size > 0)
        if not device_checks and vals.size and vals.min() < 0:   # one reduction pass instead of a comparison pass plus any()
            raise LineagePulseError("ERROR: counts contains negative elements. Requires count data.")
    for name, v in (("boolVerbose", boolVerbose), ("boolSuperVerbose", boolSuperVerbose)):
        if not isinstance(v, (bool, np.bool_)):
            raise LineagePulseError(f"ERROR IN INPUT DATA CHECK: {name} must be logical (TRUE or FALSE).")
    if strMuModel not in ("groups", "impulse", "splines", "constant"):
        raise LineagePulseError(f"strMuModel not recognised: {strMuModel} Must be one of: groups, splines, impulse, constant.")
    for nm, v in (("strDispModelFull", strDispModelFull), ("strDispModelRed", strDispModelRed)):
        if v not in ("groups", "splines", "constant"):
            raise LineagePulseError(f"{nm} not recognised: {v} Must be one of: groups, splines, constant.")
    G, N = counts.shape
    dfAnnotation = {k: np.asarray(v) for k, v in dict(dfAnnotation).items()}
    if colnames is None:
        colnames = dfAnnotation.get("cell", np.array([f"cell_{j + 1}" for j in range(N)]))
    colnames = np.asarray(colnames)
    cells = np.asarray(dfAnnotation["cell"]) if "cell" in dfAnnotation else colnames
    if not np.all(np.isin(colnames, cells)):
        raise LineagePulseError("Not all cells given in counts (colnames) are given in dfAnnotation (rownames).")
    if strMuModel in ("splines", "impulse"):
        if "continuous" not in dfAnnotation:
            raise LineagePulseError("ERROR: dfAnnotation$continuous was not given as input.")
        if not np.issubdtype(dfAnnotation["continuous"].dtype, np.number):
            raise LineagePulseError("ERROR: dfAnnotation$continuous contains non-numeric elements.")
    for confs, nm in ((vecConfoundersMu, "vecConfoundersMu"), (vecConfoundersDisp, "vecConfoundersDisp")):
        if confs:
            if not all(c in dfAnnotation for c in confs):
                raise LineagePulseError(f"Not all confounders given in {nm} given in dfAnnotation (columns).")
            for c in confs:   # processSCData.R:185-190 (the reference tests vecConfoundersMu only)
                col = dfAnnotation[c]
                na = np.array([v is None or (isinstance(v, (float, np.floating)) and np.isnan(v)) for v in col.ravel()])
                if na.any():
                    raise LineagePulseError(f"Supply batch assignments for all cells and all confounders given in {nm}")
    if matPiConstPredictors is not None:
        matPiConstPredictors = np.asarray(matPiConstPredictors, dtype=np.float64)
        if matPiConstPredictors.ndim == 1:
            matPiConstPredictors = matPiConstPredictors.reshape(-1, 1)
        if matPiConstPredictors.shape[0] != G:
            raise LineagePulseError("ERROR: Some genes named in rows of counts do not occur in rows of matPiConstPredictors.")
    mperm = None
    if len(cells) == N and not np.array_equal(colnames, cells):
        # processSCData.R:255-258 sorts the COUNT columns by match(colnames(counts), dfAnnotation$cell): column k of the sorted
        # matrix is the old column at the annotation position of the k-th column name.  That is the INVERSE of the permutation
        # that lines the columns up with the annotation; the two agree only for permutations that are their
        # own inverse (e.g. one swap).  Reproduced as it is: everything downstream pairs count columns and annotation rows by
        # position, size factors by column name (:292).
        pos = {c: k for k, c in enumerate(cells)}
        mperm = np.array([pos[c] for c in colnames])
        counts = counts.tocsc()[:, mperm] if sparse else np.ascontiguousarray(np.asarray(counts)[:, mperm])
        colnames = colnames[mperm]
        if vecNormConstExternal is not None and not isinstance(vecNormConstExternal, dict):
            vecNormConstExternal = np.asarray(vecNormConstExternal, dtype=np.float64)[mperm]   # positional: travels with its column
        if not np.array_equal(colnames, cells):
            warnings.warn("processSCData: the columns of counts were sorted as the reference sorts them (match(colnames(counts), "
                          "dfAnnotation$cell), processSCData.R:255-258), which does not line them up with dfAnnotation for this "
                          "order of cells; annotation rows and count columns are paired by position from here on, as in the "
                          "reference. Supply counts and dfAnnotation in the same order of cells.")
    if isinstance(vecNormConstExternal, dict):
        # R's named vector (processSCData.R:215-221, :292): looked up by the cell names of the count matrix
        if not all(c in vecNormConstExternal for c in colnames):
            raise LineagePulseError("ERROR IN INPUT DATA CHECK: Not all cells in counts are given in vecNormConstExternal")
        vecNormConstExternal = np.array([vecNormConstExternal[c] for c in colnames], dtype=np.float64)
    if vecNormConstExternal is not None and len(vecNormConstExternal) != N:
        raise LineagePulseError("ERROR IN INPUT DATA CHECK: Not all cells in counts are given in vecNormConstExternal")
    if rownames is None:
        rownames = np.array([f"Gene_{i + 1}" for i in range(G)])
    rownames = np.asarray(rownames)
    # an annotation with MORE rows than counts has columns (the reference fails on it, subscript out of bounds): the rows of the
    # cells of counts, in the order of the count columns
    if mperm is None and not np.array_equal(colnames, cells):
        pos = {c: k for k, c in enumerate(cells)}
        order = np.array([pos[c] for c in colnames])
        dfAnnotation = {k: v[order] for k, v in dfAnnotation.items()}
    # dense -> sparse keeps entries > 0, so NA silently become 0 (processSCData.R:260-269)
    if not sparse:
        dense = np.asarray(counts)
        if not np.issubdtype(dense.dtype, np.integer):
            dense = np.where(np.isnan(dense), 0, dense)
        dense = dense.astype(np.int32, copy=False)
    strReport = f"LineagePulse for count data: v{__version__}\n--- Data preprocessing\n"
    keep_cells = np.ones(N, dtype=bool)
    if strMuModel in ("splines", "impulse"):
        ptok = ~np.isnan(dfAnnotation["continuous"].astype(np.float64))
        strReport += f"# {int((~ptok).sum())} out of {N} cells did not have a continuous covariate and were excluded.\n"
        keep_cells &= ptok
    cm_dev = None
    if device_checks and keep_cells.all():
        from .engine import LPB200Error
        try:
            cm_dev = CountMatrix(dense, rownames, colnames, device, devices)
            rs, cs, n_neg = cm_dev.engine().count_flags()
        except LPB200Error:      # no device here: the input is still validated (on the host); any fit will fail loudly
            cm_dev, device_checks = None, False
            if vals.min() < 0:
                raise LineagePulseError("ERROR: counts contains negative elements. Requires count data.")
    if cm_dev is not None:
        m = dense
        if n_neg > 0:
            cm_dev.engine().close()
            raise LineagePulseError("ERROR: counts contains negative elements. Requires count data.")
    elif device_checks and vals.min() < 0:
        raise LineagePulseError("ERROR: counts contains negative elements. Requires count data.")
    elif sparse:
        m = counts.tocsc()[:, keep_cells]
        rs = np.asarray(m.sum(axis=1)).ravel()
        cs = np.asarray(m.sum(axis=0)).ravel()
    else:
        m = dense if keep_cells.all() else dense[:, keep_cells]
        # counts are non-negative here, so rowSums(m) > 0 (processSCData.R:285-289) is max > 0: half the time of
        # the int64 sums on a 5 000 x 10 000 matrix, and no overflow
        rs, cs = (m.max(axis=1), m.max(axis=0)) if m.size else (np.zeros(m.shape[0]), np.zeros(m.shape[1]))
    vecboolGenes, vecboolCells = rs > 0, cs > 0
    strReport += (f"# {int((~vecboolGenes).sum())} out of {len(vecboolGenes)} genes did not contain non-zero observations"
                  " and are excluded from analysis.\n")
    strReport += (f"# {int((~vecboolCells).sum())} out of {len(vecboolCells)} cells did not contain non-zero observations"
                  " and are excluded from analysis.\n")
    cell_idx = np.nonzero(keep_cells)[0][vecboolCells]
    gene_idx = np.nonzero(vecboolGenes)[0]
    if sparse:
        m = m.tocsr()[gene_idx, :].tocsc()[:, vecboolCells]
    elif not (vecboolGenes.all() and vecboolCells.all()):
        m = np.ascontiguousarray(m[gene_idx][:, vecboolCells])  # otherwise: no copy, H2D straight from the caller's buffer
        if cm_dev is not None:      # the uploaded matrix had all-zero rows / columns: upload the filtered one instead
            cm_dev.engine().close()
            cm_dev = None
    dfProc = {k: v[cell_idx] for k, v in dfAnnotation.items()}
    if "cell" not in dfProc:
        dfProc["cell"] = colnames[cell_idx]
    if boolVerbose:
        message(strReport.rstrip("\n"))
    if cm_dev is not None:
        cm_dev.rownames, cm_dev.colnames = rownames[gene_idx], colnames[cell_idx]
    objLP = LineagePulseObject(
        dfAnnotationProc=dfProc,
        matCountsProc=cm_dev if cm_dev is not None else CountMatrix(m, rownames[gene_idx], colnames[cell_idx], device, devices),
        scaDFSplinesDisp=scaDFSplinesDisp, scaDFSplinesMu=scaDFSplinesMu, strReport=strReport,
        vecAllGenes=rownames, vecConfoundersMu=vecConfoundersMu, vecConfoundersDisp=vecConfoundersDisp)
    return {"objLP": objLP,
            "vecNormConstExternalProc": None if vecNormConstExternal is None else np.asarray(vecNormConstExternal, dtype=np.float64)[cell_idx],
            "matPiConstPredictorsProc": None if matPiConstPredictors is None else matPiConstPredictors[gene_idx]}


def calcNormConst(objLP, vecNormConstExternal):
    """calcNormConst.R:20-38."""
    if vecNormConstExternal is not None:
        v = np.array(vecNormConstExternal, dtype=np.float64)
    else:
        message("# All size factors are set to one.")
        v = np.ones(objLP.matCountsProc.shape[1])
    if np.any(v == 0):
        warnings.warn("WARNING IN LINEAGEPULSE: Found size factors==0. Setting these to 1.")
        v[v == 0] = 1
    return v


def fitLPModels(objLP, matPiConstPredictors=None, strMuModel="constant", strDispModelFull="constant",
                strDispModelRed="constant", strDropModel="logistic_ofMu", strDropFitGroup="PerCell",
                boolEstimateNoiseBasedOnH0=True, scaMaxEstimationCycles=20, boolVerbose=False,
                boolSuperVerbose=False, min_rowmean_global=float("nan")):
    """fitWrapperLP.R:71-289: fits A (ZINB + drop-out EM), B (ZINB given the drop-out model),
    C and D (NB).  ``scaMaxEstimationCycles`` is accepted and NOT forwarded, as in the reference
    (fitWrapperLP.R:112-127)."""
    if boolEstimateNoiseBasedOnH0:
        strMuModelA, strMuModelB = "constant", strMuModel
        strDispModelA, strDispModelB = strDispModelRed, strDispModelFull
        nameA, nameB = f"H0: mu={strMuModelA} disp={strDispModelA}", f"H1: mu={strMuModelB} disp={strDispModelB}"
    else:
        strMuModelA, strMuModelB = strMuModel, "constant"
        strDispModelA, strDispModelB = strDispModelFull, strDispModelRed
        nameA, nameB = f"H1: mu={strMuModelA} disp={strDispModelA}", f"H0: mu={strMuModelB} disp={strDispModelB}"
    common = dict(matCounts=objLP.matCountsProc, dfAnnotation=objLP.dfAnnotationProc,
                  vecConfoundersMu=objLP.vecConfoundersMu, vecConfoundersDisp=objLP.vecConfoundersDisp,
                  vecNormConst=objLP.vecNormConst, scaDFSplinesDisp=objLP.scaDFSplinesDisp,
                  scaDFSplinesMu=objLP.scaDFSplinesMu, boolVerbose=boolVerbose, boolSuperVerbose=boolSuperVerbose,
                  min_rowmean_global=min_rowmean_global)

    def log(msg):
        objLP.strReport = (objLP.strReport or "") + msg + "\n"
        if boolVerbose:
            message(msg)

    log(f"### a) Fit ZINB model A ({nameA}) with noise model.")
    fitA = fitModel(lsDropModel=None, strMuModel=strMuModelA, strDispModel=strDispModelA, strDropModel=strDropModel,
                    strDropFitGroup=strDropFitGroup, matPiConstPredictors=matPiConstPredictors, **common)
    objLP.strReport += fitA["strReport"]
    log(f"Finished fitting zero-inflated negative binomial model A with noise model in {round(fitA['scaTime'] / 60, 2)} min.")
    lsDropModelA = fitA["lsDropModel"]
    # b) - d): the three remaining fits are independent of each other (fitWrapperLP.R:152-238); the
    # reference runs them one after the other, here their kernels overlap on the GPU
    log(f"### b) Fit ZINB model B ({nameB}).")
    log(f"### c) Fit NB model A ({nameA}).")
    log(f"### d) Fit NB model B ({nameB}).")
    fitB, fitA_NB, fitB_NB = fitModelsConcurrently(
        objLP.matCountsProc, objLP.dfAnnotationProc,
        [dict(strMuModel=strMuModelB, strDispModel=strDispModelB, strDropModel=strDropModel,
              strDropFitGroup=strDropFitGroup, lsDropModel=lsDropModelA),
         dict(strMuModel=strMuModelA, strDispModel=strDispModelA, strDropModel="none"),
         dict(strMuModel=strMuModelB, strDispModel=strDispModelB, strDropModel="none")],
        vecConfoundersMu=objLP.vecConfoundersMu, vecConfoundersDisp=objLP.vecConfoundersDisp,
        vecNormConst=objLP.vecNormConst, scaDFSplinesMu=objLP.scaDFSplinesMu,
        scaDFSplinesDisp=objLP.scaDFSplinesDisp, boolVerbose=boolVerbose, min_rowmean_global=min_rowmean_global)
    for fit, what in ((fitB, "zero-inflated negative binomial model B"), (fitA_NB, "NB model B"), (fitB_NB, "NB model B")):
        objLP.strReport += fit["strReport"]
        log(f"Finished fitting {what} in {round(fit['scaTime'] / 60, 2)} min.")
    H1, H0, H1n, H0n = (fitB, fitA, fitB_NB, fitA_NB) if boolEstimateNoiseBasedOnH0 else (fitA, fitB, fitA_NB, fitB_NB)
    objLP.lsFitConvergence = {
        "boolConvergenceH1": H1["boolConvergenceModel"], "boolConvergenceH0": H0["boolConvergenceModel"],
        "vecEMLogLikH1": H1["vecEMLogLikModel"], "vecEMLogLikH0": H0["vecEMLogLikModel"],
        "boolConvergenceH1_NB": H1n["boolConvergenceModel"], "boolConvergenceH0_NB": H0n["boolConvergenceModel"]}
    objLP.lsMuModelH1, objLP.lsDispModelH1 = H1["lsMuModel"], H1["lsDispModel"]
    objLP.lsMuModelH0, objLP.lsDispModelH0 = H0["lsMuModel"], H0["lsDispModel"]
    objLP.lsMuModelH1_NB, objLP.lsDispModelH1_NB = H1n["lsMuModel"], H1n["lsDispModel"]
    objLP.lsMuModelH0_NB, objLP.lsDispModelH0_NB = H0n["lsMuModel"], H0n["lsDispModel"]
    objLP.lsDropModel = lsDropModelA
    return objLP


_NB_DROP = {"lsDropModelGlobal": {"strDropModel": "none"}}


def runDEAnalysis(objLP):
    """runHypothesisTests.R:40-130, including its quirks: ``p_nb`` holds the ZINB p-values (:110),
    the NB degrees of freedom are taken from the ZINB models (:89-92), ``mean_H0`` is the NB-H0 mean."""
    import pandas as pd
    cm = objLP.matCountsProc
    llf = evalLogLikMatrix(cm, objLP.lsMuModelH1, objLP.lsDispModelH1, objLP.lsDropModel, objLP.matWeights)
    llr = evalLogLikMatrix(cm, objLP.lsMuModelH0, objLP.lsDispModelH0, objLP.lsDropModel, objLP.matWeights)
    llf_nb = evalLogLikMatrix(cm, objLP.lsMuModelH1_NB, objLP.lsDispModelH1_NB, _NB_DROP, objLP.matWeights)
    llr_nb = evalLogLikMatrix(cm, objLP.lsMuModelH0_NB, objLP.lsDispModelH0_NB, _NB_DROP, objLP.matWeights)
    dfH1 = objLP.lsMuModelH1["lsMuModelGlobal"]["scaDegFreedom"] + objLP.lsDispModelH1["lsDispModelGlobal"]["scaDegFreedom"]
    dfH0 = objLP.lsMuModelH0["lsMuModelGlobal"]["scaDegFreedom"] + objLP.lsDispModelH0["lsDispModelGlobal"]["scaDegFreedom"]
    p, padj = lrt(llf, llr, dfH1 - dfH0)
    p_nb, padj_nb = lrt(llf_nb, llr_nb, dfH1 - dfH0)
    if objLP.lsMuModelH0["lsMuModelGlobal"]["strMuModel"] == "constant":
        vecMu = np.asarray(objLP.lsMuModelH0_NB["matMuModel"]).ravel()
    else:
        vecMu = np.full(len(p), np.nan)
    genes = cm.rownames
    df = pd.DataFrame({"gene": genes, "p": p, "padj": padj, "mean_H0": vecMu, "p_nb": p, "padj_nb": padj_nb,
                       "df_full_zinb": dfH1, "df_red_zinb": dfH0, "df_full_nb": dfH1, "df_red_nb": dfH0,
                       "loglik_full_zinb": llf, "loglik_red_zinb": llr, "loglik_full_nb": llf_nb,
                       "loglik_red_nb": llr_nb}, index=genes)
    df = df.reindex(objLP.vecAllGenes)  # filtered genes become NA rows (:123-126)
    df["allZero"] = ~np.isin(objLP.vecAllGenes, genes)
    objLP.dfResults = df
    return objLP


def testDropout(objLP):
    """runHypothesisTests.R:179-214 (tests strDropFitGroup == "AllCells", as the reference does)."""
    import pandas as pd
    res = objLP.dfResults
    df_zinb = np.nansum(res["df_full_zinb"])
    df_nb = np.nansum(res["df_full_nb"])
    drop = objLP.lsDropModel
    grp = drop["lsDropModelGlobal"]["strDropFitGroup"]
    if grp == "PerCell":
        df_zinb += drop["matDropoutLinModel"].shape[0] * drop["matDropoutLinModel"].shape[1]
    elif grp == "AllCells":
        df_zinb += drop["matDropoutLinModel"].shape[1]
    llz, lln = np.nansum(res["loglik_full_zinb"]), np.nansum(res["loglik_full_nb"])
    dev = 2 * (llz - lln)
    p, _ = lrt(np.array([llz]), np.array([lln]), int(df_zinb - df_nb))
    return pd.DataFrame({"p": p, "df_delta": df_zinb - df_nb, "deviance": dev, "df_full": df_zinb, "df_red": df_nb,
                         "loglik_full": llz, "loglik_red": lln})


def runLineagePulse(counts, dfAnnotation=None, vecConfoundersMu=None, vecConfoundersDisp=None, strMuModel="splines",
                    strDispModelFull="constant", strDispModelRed="constant", strDropModel="logistic",
                    strDropFitGroup="PerCell", scaDFSplinesMu=6, scaDFSplinesDisp=3, matPiConstPredictors=None,
                    vecNormConstExternal=None, boolEstimateNoiseBasedOnH0=True, scaMaxEstimationCycles=20,
                    scaNProc=1, boolVerbose=True, boolSuperVerbose=False, rownames=None, colnames=None, device=0,
                    devices=None):
    """main_LineagePulse.R:219-323.  ``scaNProc`` stays in the signature; the BiocParallel fork pool it sized
    (main_LineagePulse.R:271-275) is replaced by the GPUs: ``device`` names one, ``devices=[0, 1, ...]`` several
    (one gene-sharded context, lpb200_create_multi; results are bit-identical to the one-GPU run)."""
    ls = processSCData(counts=counts, dfAnnotation=dfAnnotation, vecConfoundersMu=vecConfoundersMu,
                       vecConfoundersDisp=vecConfoundersDisp, matPiConstPredictors=matPiConstPredictors,
                       vecNormConstExternal=vecNormConstExternal, strMuModel=strMuModel,
                       strDispModelFull=strDispModelFull, strDispModelRed=strDispModelRed,
                       scaDFSplinesMu=scaDFSplinesMu, scaDFSplinesDisp=scaDFSplinesDisp,
                       scaMaxEstimationCycles=scaMaxEstimationCycles, boolVerbose=boolVerbose,
                       boolSuperVerbose=boolSuperVerbose, rownames=rownames, colnames=colnames, device=device,
                       devices=devices)
    objLP = ls["objLP"]

    def log(msg):
        objLP.strReport += msg + "\n"
        if boolVerbose:
            message(msg)

    log("--- Compute normalisation constants:")
    objLP.vecNormConst = calcNormConst(objLP, ls["vecNormConstExternalProc"])
    log("--- Fit ZINB model for both H1 and H0.")
    t0 = time.time()
    objLP = fitLPModels(objLP, matPiConstPredictors=ls["matPiConstPredictorsProc"],
                        boolEstimateNoiseBasedOnH0=boolEstimateNoiseBasedOnH0, strMuModel=strMuModel,
                        strDispModelFull=strDispModelFull, strDispModelRed=strDispModelRed,
                        strDropModel=strDropModel, strDropFitGroup=strDropFitGroup,
                        scaMaxEstimationCycles=scaMaxEstimationCycles, boolVerbose=boolVerbose,
                        boolSuperVerbose=boolSuperVerbose)
    log(f"Time elapsed during ZINB fitting: {round((time.time() - t0) / 60, 2)} min")
    log("--- Run differential expression analysis.")
    objLP = runDEAnalysis(objLP)
    objLP.strReport += "Finished runLineagePulse()."
    if boolVerbose:
        message("Finished runLineagePulse().")
    return objLP


# ---------------------------------------------------------------------------------------------
# fit extraction (getFits.R): post-hoc expansions for user-selected genes
# ---------------------------------------------------------------------------------------------
# the reference's own messages (getFits.R:63-66, :138-141, :201-204, :267-270): three of the four name getFitsDropout()
_ID_ERRORS = {"getNormData": "ERROR getNormData(): Not all vecGeneIDs were rownames of lsMuModel$matMuModel",
              "getFitsMean": "ERROR getFitsDropout(): Not all vecGeneIDs were rownames of lsMuModel$matMuModel",
              "getFitsDispersion": "ERROR getFitsDropout(): Not all vecGeneIDs were rownames of lsDispModel$matDispModel",
              "getFitsDropout": "ERROR getFitsDropout(): Not all vecGeneIDs were rownames of lsMuModel$matMuModel"}


def _check_ids(counts, vecGeneIDs, what):
    ids = np.asarray(vecGeneIDs)
    if not np.issubdtype(ids.dtype, np.integer) and not np.all(np.isin(ids, counts.rownames)):
        raise LineagePulseError(_ID_ERRORS[what])
    return _gene_indices(counts, vecGeneIDs)


def _no_drop(lsDropModel=None):
    return lsDropModel if lsDropModel is not None else _NB_DROP


def _dummy_disp(counts):
    return {"matDispModel": np.ones((counts.shape[0], 1)), "lsmatBatchModel": None,
            "lsDispModelGlobal": {"strDispModel": "constant"}}


def getNormData(matCounts, lsMuModel, vecGeneIDs, boolDepth=True, boolBatch=True):
    """getFits.R:52-98: counts divided by size factors and / or batch factors."""
    if not boolDepth and not boolBatch:
        warnings.warn("No normalisation chosen")
    counts, eng, design, model = _prepare(matCounts, lsMuModel, _dummy_disp(_as_counts(matCounts)), _NB_DROP)
    ids = _check_ids(counts, vecGeneIDs, "getNormData")
    p = _params_from_models(design, model, lsMuModel, _dummy_disp(counts), _NB_DROP)
    what = A.EXPAND_NORMDATA | (0 if boolDepth else A.EXPAND_NO_DEPTH) | (0 if boolBatch else A.EXPAND_NO_BATCH)
    return eng.expand_fits(model, p, ids, what)


def getFitsMean(lsMuModel, vecGeneIDs=None, matCounts=None):
    """getFits.R:134-154 (decompressMeansByGene; no size factors).  ``matCounts`` names the GPU-resident
    matrix whose context evaluates the expansion (the R function needs none because it runs on the host)."""
    counts = _as_counts(matCounts)
    counts, eng, design, model = _prepare(counts, lsMuModel, _dummy_disp(counts), _NB_DROP)
    ids = _check_ids(counts, vecGeneIDs, "getFitsMean")
    return eng.expand_fits(model, _params_from_models(design, model, lsMuModel, _dummy_disp(counts), _NB_DROP), ids,
                           A.EXPAND_MEAN)


def getFitsDispersion(lsDispModel, vecGeneIDs=None, matCounts=None):
    """getFits.R:196-216."""
    counts = _as_counts(matCounts)
    lsMu = {"matMuModel": np.ones((counts.shape[0], 1)), "lsmatBatchModel": None,
            "lsMuModelGlobal": {"strMuModel": "constant", "vecNormConst": None,
                                "vecContinuousCovar": lsDispModel["lsDispModelGlobal"].get("vecPseudotime")}}
    counts, eng, design, model = _prepare(counts, lsMu, lsDispModel, _NB_DROP)
    ids = _check_ids(counts, vecGeneIDs, "getFitsDispersion")
    return eng.expand_fits(model, _params_from_models(design, model, lsMu, lsDispModel, _NB_DROP), ids, A.EXPAND_DISP)


def getFitsDropout(lsMuModel, lsDropModel, vecGeneIDs=None, matCounts=None):
    """getFits.R:261-288."""
    counts = _as_counts(matCounts)
    counts, eng, design, model = _prepare(counts, lsMuModel, _dummy_disp(counts), lsDropModel)
    ids = _check_ids(counts, vecGeneIDs, "getFitsDropout")
    return eng.expand_fits(model, _params_from_models(design, model, lsMuModel, _dummy_disp(counts), lsDropModel), ids,
                           A.EXPAND_DROP)


def getPostDrop(matCounts, lsMuModel, lsDispModel, lsDropModel, vecGeneIDs):
    """getFits.R:340-363."""
    counts = _as_counts(matCounts)
    if not np.issubdtype(np.asarray(vecGeneIDs).dtype, np.integer) and not np.all(np.isin(vecGeneIDs, counts.rownames)):
        raise LineagePulseError("ERROR getFitsDropout(): Not all vecGeneIDs were rownames of matCounts.")
    return calcPostDrop_Matrix(counts, lsMuModel, lsDispModel, lsDropModel, vecIDs=vecGeneIDs)
