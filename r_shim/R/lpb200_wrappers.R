### lpb200_wrappers.R -- drop-in bodies for the L2 boundary functions of LineagePulse
### (SURVEY.md 8b).  Names, signatures and return values are the reference's
### (R/srcLineagePulse_fitMeanDispersion.R:853, R/srcLineagePulse_fitDropout.R:418,
### R/srcLineagePulse_evalLogLik.R:307, R/srcLineagePulse_fitModel.R:137); the bodies call the
### CUDA library through the .Call shim src/lp_rshim.c.  BiocParallel is not used around these
### calls (fork + CUDA context is unsafe); scaNProc stays in runLineagePulse's signature.
### UNTESTED against a real R session: R is not available in the build image (see INTEGRATION.md).

.lpStrMu   <- c(constant=0L, impulse=1L, splines=2L, groups=3L, MM=4L)
.lpStrDisp <- c(constant=0L, splines=2L, groups=3L)
.lpStrDrop <- c(none=0L, logistic=1L, logistic_ofMu=2L)
.lpStrGrp  <- c(PerCell=0L, ForAllCells=1L)

## one GPU context per count matrix, cached on the dgCMatrix via an attribute
lpContext <- function(matCounts, device = 0L) {
    ctx <- attr(matCounts, "lpb200_ctx")
    if (is.null(ctx)) {
        ctx <- .Call("lp_create", as.integer(device))
        if (is(matCounts, "dgCMatrix")) {
            .Call("lp_load_counts_csc", ctx, dim(matCounts), matCounts@p, matCounts@i, as.double(matCounts@x))
        } else {
            .Call("lp_load_counts_dense", ctx, matrix(as.double(matCounts), nrow(matCounts)))
        }
    }
    ctx
}

lpDesign <- function(matCounts, lsMuModel, lsDispModel, lsDropModel) {
    gm <- lsMuModel$lsMuModelGlobal; gd <- lsDispModel$lsDispModelGlobal
    t <- gm$vecContinuousCovar
    idxGroups <- if (!is.null(gm$vecidxGroups)) gm$vecidxGroups else gd$vecidxGroups
    list(n_genes = nrow(matCounts), n_cells = ncol(matCounts),
         vecNormConst = as.double(gm$vecNormConst),
         vecContinuousCovar = if (is.null(t)) NULL else as.double(t),
         vecidxGroups = if (is.null(idxGroups)) NULL else as.integer(idxGroups),
         scaNGroups = as.integer(max(0L, gm$scaNGroups, gd$scaNGroups)),
         matSplineBasisMu = gm$matSplineBasis, matSplineBasisDisp = gd$matSplineBasis,
         matImpulseBasis = if (gm$strMuModel == "impulse")
             matrix(as.double(splines::ns(t, df = 4, intercept = TRUE)), ncol = 4) else NULL,
         vecNBatchesMu = if (is.null(gm$vecConfounders)) NULL else as.integer(gm$vecNBatches),
         vecidxBatchAssignMu = if (is.null(gm$vecConfounders)) NULL else as.integer(unlist(gm$lsvecidxBatchAssign)),
         vecNBatchesDisp = if (is.null(gd$vecConfounders)) NULL else as.integer(gd$vecNBatches),
         vecidxBatchAssignDisp = if (is.null(gd$vecConfounders)) NULL else as.integer(unlist(gd$lsvecidxBatchAssign)),
         matPiConstPredictors = lsDropModel$matPiConstPredictors)
}

lpModel <- function(lsMuModel, lsDispModel, lsDropModel) {
    strDrop <- lsDropModel$lsDropModelGlobal$strDropModel
    strGrp <- lsDropModel$lsDropModelGlobal$strDropFitGroup
    if (is.null(strGrp)) strGrp <- "PerCell"
    c(.lpStrMu[[lsMuModel$lsMuModelGlobal$strMuModel]], .lpStrDisp[[lsDispModel$lsDispModelGlobal$strDispModel]],
      .lpStrDrop[[strDrop]], .lpStrGrp[[strGrp]])
}

lpParams <- function(lsMuModel, lsDispModel, lsDropModel) {
    bind <- function(ls, used) if (used) do.call(cbind, ls) + 0 else NULL   # "+ 0" forces a private copy
    list(matMuModel = lsMuModel$matMuModel + 0,
         matBatchMu = bind(lsMuModel$lsmatBatchModel, !is.null(lsMuModel$lsMuModelGlobal$vecConfounders)),
         matDispModel = lsDispModel$matDispModel + 0,
         matBatchDisp = bind(lsDispModel$lsmatBatchModel, !is.null(lsDispModel$lsDispModelGlobal$vecConfounders)),
         matDropoutLinModel = if (is.null(lsDropModel$matDropoutLinModel)) NULL else lsDropModel$matDropoutLinModel + 0)
}

lpSplitBatch <- function(mat, vecNBatches) {
    if (is.null(mat)) return(NULL)
    ends <- cumsum(vecNBatches); starts <- ends - vecNBatches + 1
    lapply(seq_along(vecNBatches), function(b) mat[, starts[b]:ends[b], drop = FALSE])
}

evalLogLikMatrix <- function(matCounts, lsMuModel, lsDispModel, lsDropModel,
                             matWeights = NULL, boolConstModelOnMMLL = FALSE) {
    if (boolConstModelOnMMLL | lsMuModel$lsMuModelGlobal$strMuModel == "MM")
        stop("LineagePulse ERROR: evalLogLikGeneMM() not yet supported. Contact developer.")
    ctx <- lpContext(matCounts)
    .Call("lp_set_design", ctx, lpDesign(matCounts, lsMuModel, lsDispModel, lsDropModel))
    .Call("lp_eval_loglik", ctx, lpModel(lsMuModel, lsDispModel, lsDropModel),
          lpParams(lsMuModel, lsDispModel, lsDropModel), nrow(matCounts))
}

fitMuDisp <- function(matCountsProc, vecNormConst, lsMuModel, lsDispModel, lsDropModel, matWeights) {
    if (lsMuModel$lsMuModelGlobal$strMuModel == "MM")
        stop("LineagePulse ERROR: fitMMZINB() not yet supported. Contact developer.")
    ctx <- lpContext(matCountsProc)
    .Call("lp_set_design", ctx, lpDesign(matCountsProc, lsMuModel, lsDispModel, lsDropModel))
    res <- .Call("lp_fit_mudisp", ctx, lpModel(lsMuModel, lsDispModel, lsDropModel),
                 lpParams(lsMuModel, lsDispModel, lsDropModel), nrow(matCountsProc))
    p <- res[[1]]
    rownames(p$matMuModel) <- rownames(matCountsProc); rownames(p$matDispModel) <- rownames(matCountsProc)
    list(matMuModel = p$matMuModel,
         lsmatBatchModelMu = lpSplitBatch(p$matBatchMu, lsMuModel$lsMuModelGlobal$vecNBatches),
         matDispModel = p$matDispModel,
         lsmatBatchModelDisp = lpSplitBatch(p$matBatchDisp, lsDispModel$lsDispModelGlobal$vecNBatches),
         vecConvergence = res[[2]], vecLL = res[[3]])
}

fitPi <- function(matCounts, lsMuModel, lsDispModel, lsDropModel) {
    ctx <- lpContext(matCounts)
    .Call("lp_set_design", ctx, lpDesign(matCounts, lsMuModel, lsDispModel, lsDropModel))
    nOut <- if (lsDropModel$lsDropModelGlobal$strDropFitGroup == "PerCell") ncol(matCounts) else 1L
    res <- .Call("lp_fit_pi", ctx, lpModel(lsMuModel, lsDispModel, lsDropModel),
                 lpParams(lsMuModel, lsDispModel, lsDropModel), nOut)
    matDropoutLinModel <- res[[1]]$matDropoutLinModel
    rownames(matDropoutLinModel) <- colnames(matCounts)
    list(matDropoutLinModel = matDropoutLinModel, vecConverged = res[[2]], vecLL = res[[3]])
}

## LRT part of runDEAnalysis (R/srcLineagePulse_runHypothesisTests.R:81-86); the data.frame assembly stays in R
lpLRT <- function(vecLogLikFull, vecLogLikRed, scaDeltaDegFreedom) {
    res <- .Call("lp_lrt", as.double(vecLogLikFull), as.double(vecLogLikRed), as.integer(scaDeltaDegFreedom))
    list(p = res[[1]], padj = res[[2]])
}
