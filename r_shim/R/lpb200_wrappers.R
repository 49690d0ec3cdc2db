### lpb200_wrappers.R -- drop-in bodies for the L2 boundary functions of LineagePulse
### (SURVEY.md 8b).  Names, signatures and return values are the reference's
### (R/srcLineagePulse_fitMeanDispersion.R:853, R/srcLineagePulse_fitDropout.R:418,
### R/srcLineagePulse_evalLogLik.R:307, R/srcLineagePulse_fitModel.R:137); the bodies call the
### CUDA library through the .Call shim src/lp_rshim.c.  BiocParallel is not used around these
### calls (fork + CUDA context is unsafe); scaNProc stays in runLineagePulse's signature.
### UNTESTED against a real R session: R is not available in the build image (see INTEGRATION.md); the C side
### of every .Call below is run on the GPU through a stand-in R runtime (tests/csrc/fake_r_runtime.c).

.lpStrMu   <- c(constant=0L, impulse=1L, splines=2L, groups=3L, MM=4L)
.lpStrDisp <- c(constant=0L, splines=2L, groups=3L)
.lpStrDrop <- c(none=0L, logistic=1L, logistic_ofMu=2L)
.lpStrGrp  <- c(PerCell=0L, ForAllCells=1L)

## One GPU context per count matrix.  R copies on modify, so a context cannot be hung on the caller's matrix
## from inside fitMuDisp(); instead lpAttachCounts() returns the matrix WITH an attribute holding an environment
## (a reference object: the context pointer and the key of the design last sent to the device live in it and
## are shared by every copy of the matrix).  processSCData / fitModel attach once (objLP@matCountsProc <-
## lpAttachCounts(objLP@matCountsProc, devices)); every later fitMuDisp / fitPi / evalLogLikMatrix call on that
## object reuses the resident counts.  A matrix that was never attached still works: it gets a context for the
## duration of the call (one upload per call) and a message says so.
## devices: integer vector of CUDA device ids -- several ids make one gene-sharded multi-GPU context, the
## replacement of scaNProc (main_LineagePulse.R:271-275).
lpAttachCounts <- function(matCounts, devices = getOption("LineagePulse.devices", 0L)) {
    if (!is.null(attr(matCounts, "lpb200"))) return(matCounts)
    state <- new.env(parent = emptyenv())
    state$ctx <- .Call("lp_create", as.integer(devices))
    if (is(matCounts, "dgCMatrix")) {
        .Call("lp_load_counts_csc", state$ctx, dim(matCounts), matCounts@p, matCounts@i, as.double(matCounts@x))
    } else {
        .Call("lp_load_counts_dense", state$ctx, matrix(as.double(matCounts), nrow(matCounts)))
    }
    state$design_key <- NULL
    attr(matCounts, "lpb200") <- state
    matCounts
}

lpContext <- function(matCounts) {
    state <- attr(matCounts, "lpb200")
    if (is.null(state)) {
        message("LineagePulse: count matrix not attached to a GPU context (lpAttachCounts); uploading it for this call only.")
        state <- attr(lpAttachCounts(matCounts), "lpb200")
    }
    state
}

## send the design only when it differs from the one the device already holds
lpSetDesign <- function(state, lsDesign) {
    key <- list(lsDesign$n_genes, lsDesign$n_cells, lsDesign$vecNormConst, lsDesign$vecContinuousCovar,
                lsDesign$vecidxGroups, dim(lsDesign$matSplineBasisMu), dim(lsDesign$matSplineBasisDisp),
                !is.null(lsDesign$matImpulseBasis), lsDesign$vecidxBatchAssignMu, lsDesign$vecidxBatchAssignDisp,
                lsDesign$matPiConstPredictors)
    if (!identical(key, state$design_key)) {
        .Call("lp_set_design", state$ctx, lsDesign)
        state$design_key <- key
    }
    invisible(state$ctx)
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

## optimiser constants of the next fit: the reference reads them from the model lists (fitMeanDispersion.R:912-913,
## fitDropout.R:291-292) and from fitModel's scaMaxEstimationCycles (fitModel.R:156); absent entries -> NA -> the
## constants of fitModel.R:164-175
lpControl <- function(lsMuModel, lsDropModel, scaMaxEstimationCycles = NA) {
    pick <- function(x) if (is.null(x)) NA_real_ else as.double(x)
    gm <- lsMuModel$lsMuModelGlobal; gp <- lsDropModel$lsDropModelGlobal
    invisible(.Call("lp_set_control", c(pick(gm$MAXIT_BFGS_MuDisp), pick(gm$RELTOL_BFGS_MuDisp), pick(gp$MAXIT_BFGS_Pi),
                                        pick(gp$RELTOL_BFGS_Pi), as.double(scaMaxEstimationCycles))))
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
    ctx <- lpSetDesign(lpContext(matCounts), lpDesign(matCounts, lsMuModel, lsDispModel, lsDropModel))
    .Call("lp_eval_loglik", ctx, lpModel(lsMuModel, lsDispModel, lsDropModel),
          lpParams(lsMuModel, lsDispModel, lsDropModel), nrow(matCounts))
}

fitMuDisp <- function(matCountsProc, vecNormConst, lsMuModel, lsDispModel, lsDropModel, matWeights) {
    if (lsMuModel$lsMuModelGlobal$strMuModel == "MM")
        stop("LineagePulse ERROR: fitMMZINB() not yet supported. Contact developer.")
    ctx <- lpSetDesign(lpContext(matCountsProc), lpDesign(matCountsProc, lsMuModel, lsDispModel, lsDropModel))
    lpControl(lsMuModel, lsDropModel)
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
    ctx <- lpSetDesign(lpContext(matCounts), lpDesign(matCounts, lsMuModel, lsDispModel, lsDropModel))
    nOut <- if (lsDropModel$lsDropModelGlobal$strDropFitGroup == "PerCell") ncol(matCounts) else 1L
    lpControl(lsMuModel, lsDropModel)
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

## fitModel (R/srcLineagePulse_fitModel.R:137-356): the whole EM loop is one .Call.  Metadata (the ...Global lists,
## degrees of freedom, spline bases, batch assignments) still comes from the reference's own initMuModel /
## initDispModel / initDropModel; the initial VALUES they compute are recomputed identically on the device.
fitModel <- function(matCounts, dfAnnotation, vecConfoundersMu = NULL, vecConfoundersDisp = NULL, vecNormConst,
                     scaDFSplinesMu = NULL, scaDFSplinesDisp = NULL, matWeights = NULL, matPiConstPredictors = NULL,
                     lsDropModel = NULL, matMuModelInit = NULL, lsmatBatchModelInitMu = NULL,
                     matDispModelInit = NULL, lsmatBatchModelInitDisp = NULL, strMuModel, strDispModel,
                     strDropModel = "logistic_ofMu", strDropFitGroup = "PerCell", scaMaxEstimationCycles = 20,
                     boolVerbose = TRUE, boolSuperVerbose = TRUE) {
    if (strMuModel == "MM") stop("LineagePulse ERROR: fitMMZINB() not yet supported. Contact developer.")
    matCounts <- lpAttachCounts(matCounts)
    boolFitDrop <- is.null(lsDropModel) & strDropModel != "none"
    lsMuModel <- initMuModel(matCounts = matCounts, dfAnnotation = dfAnnotation, vecConfoundersMu = vecConfoundersMu,
                             vecNormConst = vecNormConst, scaDFSplinesMu = scaDFSplinesMu, matWeights = matWeights,
                             matMuModelInit = matMuModelInit, lsmatBatchModelInitMu = lsmatBatchModelInitMu,
                             strMuModel = strMuModel, MAXIT_BFGS_MuDisp = 1000, RELTOL_BFGS_MuDisp = 10^(-8))
    lsDispModel <- initDispModel(matCounts = matCounts, dfAnnotation = dfAnnotation,
                                 vecConfoundersDisp = vecConfoundersDisp, scaDFSplinesDisp = scaDFSplinesDisp,
                                 matWeights = matWeights, matDispModelInit = matDispModelInit,
                                 lsmatBatchModelInitDisp = lsmatBatchModelInitDisp, strDispModel = strDispModel,
                                 strMuModel = strMuModel, MAXIT_BFGS_MuDisp = 1000, RELTOL_BFGS_MuDisp = 10^(-8))
    if (is.null(lsDropModel))
        lsDropModel <- initDropModel(matCounts = matCounts, matPiConstPredictors = matPiConstPredictors,
                                     lsDropModel = lsDropModel, strDropModel = strDropModel,
                                     strDropFitGroup = strDropFitGroup, MAXIT_BFGS_Pi = 10000, RELTOL_BFGS_Pi = 10^(-8))
    ctx <- lpSetDesign(lpContext(matCounts), lpDesign(matCounts, lsMuModel, lsDispModel, lsDropModel))
    ## warm-start hooks (fitModel.R:147-151): which of the matrices are the caller's initial values
    ## (unary ! binds more loosely than * and + in R: the parentheses are needed)
    scaKeepMask <- 1L * (!is.null(matMuModelInit)) + 2L * (!is.null(lsmatBatchModelInitMu)) +
        4L * (!is.null(matDispModelInit)) + 8L * (!is.null(lsmatBatchModelInitDisp))
    lpControl(lsMuModel, lsDropModel, scaMaxEstimationCycles)
    res <- .Call("lp_fit_model", ctx, lpModel(lsMuModel, lsDispModel, lsDropModel),
                 lpParams(lsMuModel, lsDispModel, lsDropModel), boolFitDrop, nrow(matCounts), as.integer(scaKeepMask))
    p <- res[[1]]; info <- res[[5]]
    rownames(p$matMuModel) <- rownames(matCounts); rownames(p$matDispModel) <- rownames(matCounts)
    lsMuModel$matMuModel <- p$matMuModel
    lsMuModel$lsmatBatchModel <- lpSplitBatch(p$matBatchMu, lsMuModel$lsMuModelGlobal$vecNBatches)
    lsDispModel$matDispModel <- p$matDispModel
    lsDispModel$lsmatBatchModel <- lpSplitBatch(p$matBatchDisp, lsDispModel$lsDispModelGlobal$vecNBatches)
    if (boolFitDrop) {
        lsDropModel$matDropoutLinModel <- p$matDropoutLinModel
        rownames(lsDropModel$matDropoutLinModel) <- colnames(matCounts)
    }
    nIter <- as.integer(info[1])
    vecEMLogLikModel <- res[[2]][seq_len(if (boolFitDrop) scaMaxEstimationCycles else 1L)]
    strReport <- paste0("#  .   Initialisation: ll ", info[3], "\n",
                        paste0("# ", seq_len(nIter), ".  Iteration with ll   ", res[[2]][seq_len(nIter)], collapse = "\n"), "\n")
    if (any(res[[4]] != 0))
        strReport <- paste0(strReport, "(Mean-) Dispersion estimation did not converge in ", sum(res[[4]] != 0), " cases.\n")
    if (boolVerbose) message(strReport)
    list(lsMuModel = lsMuModel, lsDispModel = lsDispModel, lsDropModel = lsDropModel,
         boolConvergenceModel = as.logical(info[2]), vecEMLogLikModel = vecEMLogLikModel, strReport = strReport)
}

## fits B, C, D of fitLPModels (R/srcLineagePulse_fitWrapperLP.R:152-238): independent given fit A's drop-out model,
## so one .Call runs their fit kernels concurrently.  lsSpecs: list of list(strMuModel, strDispModel, strDropModel,
## lsDropModel); returns one fitModel()-style list per spec.
fitModelsConcurrently <- function(matCounts, dfAnnotation, lsSpecs, vecConfoundersMu = NULL, vecConfoundersDisp = NULL,
                                  vecNormConst, scaDFSplinesMu = NULL, scaDFSplinesDisp = NULL) {
    matCounts <- lpAttachCounts(matCounts)
    lsInit <- lapply(lsSpecs, function(s) {
        lsMuModel <- initMuModel(matCounts = matCounts, dfAnnotation = dfAnnotation, vecConfoundersMu = vecConfoundersMu,
                                 vecNormConst = vecNormConst, scaDFSplinesMu = scaDFSplinesMu, matWeights = NULL,
                                 matMuModelInit = NULL, lsmatBatchModelInitMu = NULL, strMuModel = s$strMuModel,
                                 MAXIT_BFGS_MuDisp = 1000, RELTOL_BFGS_MuDisp = 10^(-8))
        lsDispModel <- initDispModel(matCounts = matCounts, dfAnnotation = dfAnnotation,
                                     vecConfoundersDisp = vecConfoundersDisp, scaDFSplinesDisp = scaDFSplinesDisp,
                                     matWeights = NULL, matDispModelInit = NULL, lsmatBatchModelInitDisp = NULL,
                                     strDispModel = s$strDispModel, strMuModel = s$strMuModel,
                                     MAXIT_BFGS_MuDisp = 1000, RELTOL_BFGS_MuDisp = 10^(-8))
        lsDropModel <- if (s$strDropModel == "none") list(lsDropModelGlobal = list(strDropModel = "none")) else s$lsDropModel
        list(lsMuModel = lsMuModel, lsDispModel = lsDispModel, lsDropModel = lsDropModel)
    })
    ## one design holding every covariate any of the models needs: the H1 model's (impulse / splines / groups)
    idxFull <- which.max(vapply(lsInit, function(m) m$lsMuModel$lsMuModelGlobal$scaDegFreedom, numeric(1)))
    ctx <- lpSetDesign(lpContext(matCounts), lpDesign(matCounts, lsInit[[idxFull]]$lsMuModel, lsInit[[idxFull]]$lsDispModel,
                                                      lsSpecs[[1]]$lsDropModel))
    matModels <- vapply(lsInit, function(m) lpModel(m$lsMuModel, m$lsDispModel, m$lsDropModel), integer(4))
    lpControl(lsInit[[idxFull]]$lsMuModel, lsSpecs[[1]]$lsDropModel)
    res <- .Call("lp_fit_models", ctx, matModels,
                 lapply(lsInit, function(m) lpParams(m$lsMuModel, m$lsDispModel, m$lsDropModel)), nrow(matCounts))
    G <- nrow(matCounts)
    lapply(seq_along(lsInit), function(j) {
        m <- lsInit[[j]]; p <- res[[1]][[j]]
        rownames(p$matMuModel) <- rownames(matCounts); rownames(p$matDispModel) <- rownames(matCounts)
        m$lsMuModel$matMuModel <- p$matMuModel
        m$lsMuModel$lsmatBatchModel <- lpSplitBatch(p$matBatchMu, m$lsMuModel$lsMuModelGlobal$vecNBatches)
        m$lsDispModel$matDispModel <- p$matDispModel
        m$lsDispModel$lsmatBatchModel <- lpSplitBatch(p$matBatchDisp, m$lsDispModel$lsDispModelGlobal$vecNBatches)
        list(lsMuModel = m$lsMuModel, lsDispModel = m$lsDispModel, lsDropModel = m$lsDropModel,
             boolConvergenceModel = as.logical(res[[4]][j]), vecEMLogLikModel = res[[3]][j],
             strReport = paste0("#  .   Initialisation: ll ", res[[2]][j], "\n# 1.  Iteration with ll   ", res[[3]][j], "\n"),
             vecLL = res[[5]][(j - 1) * G + seq_len(G)], vecConvergence = res[[6]][(j - 1) * G + seq_len(G)])
    })
}

## fitLPModels (R/srcLineagePulse_fitWrapperLP.R:71-289): fit A with the drop-out EM, then B, C, D in one call.
fitLPModels <- function(objLP, matPiConstPredictors, strMuModel = "constant", strDispModelFull = "constant",
                        strDispModelRed = "constant", strDropModel = "logistic_ofMu", strDropFitGroup = "PerCell",
                        boolEstimateNoiseBasedOnH0 = TRUE, scaMaxEstimationCycles = 20, boolVerbose = FALSE,
                        boolSuperVerbose = FALSE, devices = getOption("LineagePulse.devices", 0L)) {
    objLP@matCountsProc <- lpAttachCounts(objLP@matCountsProc, devices)
    if (boolEstimateNoiseBasedOnH0) {
        strMuModelA <- "constant"; strMuModelB <- strMuModel; strDispModelA <- strDispModelRed; strDispModelB <- strDispModelFull
    } else {
        strMuModelA <- strMuModel; strMuModelB <- "constant"; strDispModelA <- strDispModelFull; strDispModelB <- strDispModelRed
    }
    common <- list(matCounts = objLP@matCountsProc, dfAnnotation = objLP@dfAnnotationProc,
                   vecConfoundersMu = objLP@vecConfoundersMu, vecConfoundersDisp = objLP@vecConfoundersDisp,
                   vecNormConst = objLP@vecNormConst, scaDFSplinesMu = objLP@scaDFSplinesMu,
                   scaDFSplinesDisp = objLP@scaDFSplinesDisp)
    fitA <- do.call(fitModel, c(common, list(matPiConstPredictors = matPiConstPredictors, lsDropModel = NULL,
                                             strMuModel = strMuModelA, strDispModel = strDispModelA,
                                             strDropModel = strDropModel, strDropFitGroup = strDropFitGroup,
                                             boolVerbose = boolVerbose, boolSuperVerbose = boolSuperVerbose)))
    lsFits <- do.call(fitModelsConcurrently, c(common, list(lsSpecs = list(
        list(strMuModel = strMuModelB, strDispModel = strDispModelB, strDropModel = strDropModel, lsDropModel = fitA$lsDropModel),
        list(strMuModel = strMuModelA, strDispModel = strDispModelA, strDropModel = "none"),
        list(strMuModel = strMuModelB, strDispModel = strDispModelB, strDropModel = "none")))))
    fitB <- lsFits[[1]]; fitA_NB <- lsFits[[2]]; fitB_NB <- lsFits[[3]]
    if (boolEstimateNoiseBasedOnH0) { H1 <- fitB; H0 <- fitA; H1n <- fitB_NB; H0n <- fitA_NB
    } else { H1 <- fitA; H0 <- fitB; H1n <- fitA_NB; H0n <- fitB_NB }
    objLP@lsMuModelH1 <- H1$lsMuModel; objLP@lsDispModelH1 <- H1$lsDispModel
    objLP@lsMuModelH0 <- H0$lsMuModel; objLP@lsDispModelH0 <- H0$lsDispModel
    objLP@lsMuModelH1_NB <- H1n$lsMuModel; objLP@lsDispModelH1_NB <- H1n$lsDispModel
    objLP@lsMuModelH0_NB <- H0n$lsMuModel; objLP@lsDispModelH0_NB <- H0n$lsDispModel
    objLP@lsDropModel <- fitA$lsDropModel
    objLP@lsFitConvergence <- list(boolConvergenceH1 = H1$boolConvergenceModel, boolConvergenceH0 = H0$boolConvergenceModel,
                                   vecEMLogLikH1 = H1$vecEMLogLikModel, vecEMLogLikH0 = H0$vecEMLogLikModel,
                                   boolConvergenceH1_NB = H1n$boolConvergenceModel, boolConvergenceH0_NB = H0n$boolConvergenceModel)
    objLP@strReport <- paste0(objLP@strReport, fitA$strReport, fitB$strReport, fitA_NB$strReport, fitB_NB$strReport)
    objLP
}
