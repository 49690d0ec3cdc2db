// fit_mudisp_kernel for int32_t counts (see lp_fit_table.cuh)
#define LP_FIT_TABLE_COUNT_TYPE 1
#include "lp_fit_table.cuh"
FitKernelFn lp_fit_kernel_i32(int variant, bool cluster) {
    return cluster ? fit_kernel_ct<int32_t, true>(variant) : fit_kernel_ct<int32_t, false>(variant);
}
