// libm_harness.cpp -- TEST INFRASTRUCTURE: compiles the device routines of lineagepulse_b200/csrc/lp_libm.cuh
// (exp, log, check-free division / reciprocal) for the host through the header's LP_LIBM_HOST_EMULATION
// switch, so their constants, operation order and exponent handling can be checked against the C library
// on a box without a GPU.  The hardware reciprocal seed is emulated (upper word of the true reciprocal);
// the bit-for-bit comparison with the CUDA math library itself is a GPU test.  Not part of the product.
#define LP_LIBM_HOST_EMULATION 1
#include "../../lineagepulse_b200/csrc/lp_libm.cuh"

extern "C" void lm_exp(const double *x, int n, double *single, double *paired) {
    for (int k = 0; k < n; k++) single[k] = lp_exp_device(x[k]);
    for (int k = 0; k + 1 < n; k += 2) lp_exp_pair_device(x[k], x[k + 1], paired[k], paired[k + 1]);
    if (n & 1) paired[n - 1] = lp_exp_device(x[n - 1]);
}
extern "C" void lm_log(const double *x, int n, double *full, double *positive_only) {
    for (int k = 0; k < n; k++) {
        full[k] = lp_log_device(x[k]);
        positive_only[k] = lp_log_pos_device(x[k]);   // meaningful for positive normal finite x only
    }
}
extern "C" void lm_div(const double *a, const double *b, int n, double *quot, double *rcp) {
    for (int k = 0; k < n; k++) {
        quot[k] = lp_div_fast_device(a[k], b[k]);
        rcp[k] = lp_rcp_fast_device(b[k]);
    }
}
