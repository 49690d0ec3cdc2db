// lp_fit_table.cuh -- the table of fit_mudisp_kernel instantiations, one translation unit per count type.
#pragma once
#include "lp_fit_kernel.cuh"

typedef void (*FitKernelFn)(const FitArgs);
#define LP_N_FIT_VARIANTS 11
// variant: 0 generic; 1 / 2 constant mean NB / ZINB; 3 / 4 impulse NB / ZINB; 5 / 6 groups NB / ZINB; 7 / 8 splines NB / ZINB;
// 9 / 10 constant mean with batch factors NB / ZINB
FitKernelFn lp_fit_kernel_u16(int variant, bool cluster);
FitKernelFn lp_fit_kernel_i32(int variant, bool cluster);

#ifdef LP_FIT_TABLE_COUNT_TYPE
template <typename CT, bool CL>
static FitKernelFn fit_kernel_ct(int variant) {
    switch (variant) {
    case 1: return fit_mudisp_kernel<CT, LP_MU_FAST_CONST, false, CL>;
    case 2: return fit_mudisp_kernel<CT, LP_MU_FAST_CONST, true, CL>;
    case 3: return fit_mudisp_kernel<CT, LP_MU_FAST_IMPULSE, false, CL>;
    case 4: return fit_mudisp_kernel<CT, LP_MU_FAST_IMPULSE, true, CL>;
    default: break;
    }
    if (!CL) {   // the covariate-model kernels are built for one CTA per item only
        switch (variant) {
        case 5: return fit_mudisp_kernel<CT, LP_MU_FAST_GROUPS, false, false>;
        case 6: return fit_mudisp_kernel<CT, LP_MU_FAST_GROUPS, true, false>;
        case 7: return fit_mudisp_kernel<CT, LP_MU_FAST_SPLINES, false, false>;
        case 8: return fit_mudisp_kernel<CT, LP_MU_FAST_SPLINES, true, false>;
        case 9: return fit_mudisp_kernel<CT, LP_MU_FAST_CONSTCOV, false, false>;
        case 10: return fit_mudisp_kernel<CT, LP_MU_FAST_CONSTCOV, true, false>;
        default: break;
        }
    }
    return fit_mudisp_kernel<CT, -1, false, CL>;
}
#endif
