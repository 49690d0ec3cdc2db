"""lineagepulse_b200 -- B200-native implementation of LineagePulse's per-gene (ZI)NB model-fitting
path (fitMuDisp / fitPi / evalLogLikMatrix / fitModel + LRT) behind the reference's own function
names.  All arithmetic runs in the CUDA library ``liblpb200.so`` (sm_100a); there is no CPU
fallback.  See DESIGN.md and INTEGRATION.md.
"""
from ._cabi import (DesignData, Model, ParamsData, Control, default_control, LibraryMissing,  # noqa: F401
                    MU_MODELS, DISP_MODELS, DROP_MODELS, DROP_FIT_GROUPS)
from .engine import Engine, LPB200Error, lrt  # noqa: F401
from .splines import ns  # noqa: F401
from .simulate import simulateContinuousDataSet  # noqa: F401
from .api import (  # noqa: F401
    CountMatrix, LineagePulseObject, LineagePulseError, runLineagePulse, processSCData, calcNormConst,
    fitLPModels, fitModel, fitModelsConcurrently, fitMuDisp, fitPi, evalLogLikMatrix, runDEAnalysis, testDropout,
    calcPostDrop_Matrix, getPostDrop, getNormData, getFitsMean, getFitsDispersion, getFitsDropout,
    writeReport)
