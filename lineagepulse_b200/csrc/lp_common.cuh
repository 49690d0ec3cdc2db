// lp_common.cuh -- shared declarations of the CUDA library (device-side views of the design,
// per-point natural-scale parameters).  Headers marked LP_HD also compile as plain C++ so that
// tests/csrc/host_harness.cpp can exercise the scalar logic on a CPU-only box.
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/lpb200.h"

#ifdef __CUDACC__
#define LP_HD __host__ __device__ __forceinline__
#define LP_D __device__ __forceinline__
#else
#define LP_HD inline
#define LP_D inline
#endif

// separately rounded multiply / add: the vmmin restatement must not be FMA-contracted
// (the "no change" test `10.0 + X[i] == 10.0 + b[i]` of R's vmmin is rounding sensitive).
#if defined(__CUDA_ARCH__)
#define LP_MUL(a, b) __dmul_rn((a), (b))
#define LP_ADD(a, b) __dadd_rn((a), (b))
#else
static inline double lp_mul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double lp_add_rn(double a, double b) { volatile double r = a + b; return r; }
#define LP_MUL(a, b) lp_mul_rn((a), (b))
#define LP_ADD(a, b) lp_add_rn((a), (b))
#endif

#define LP_LL_FLOOR (-700.0)      // scaLogLikPrecLim, evalLogLik.R:37,123
#define LP_CLAMP_LO 1e-10         // 10^(-10), fitMeanDispersion.R:82
#define LP_CLAMP_HI 1e10          // 10^(10),  fitMeanDispersion.R:83
#define LP_NA_U16 0xFFFFu

// capacity of one point's natural-scale parameter block (checked by the API)
#define LP_MAX_PDISP 12           // dispersion-model parameters (groups / spline coefficients)
#define LP_MAX_PMU 16             // mean-model parameters
#define LP_MAX_BDISP 12           // dispersion batch factors incl. the leading 1 of every confounder
#define LP_MAX_BMU 20             // mean batch factors incl. leading ones

// device-side view of lpb200_design (all pointers are device pointers inside kernels)
struct DevDesign {
    int G, N;
    const double *sf;        // [N] or nullptr (all 1)
    const double *t;         // [N]
    const int32_t *group;    // [N]
    int n_groups;
    const double *bs_mu;     // N x k_mu column-major
    int k_mu;
    const double *bs_disp;   // N x k_disp
    int k_disp;
    int n_conf_mu;
    const int32_t *bidx_mu;  // [n_conf_mu * N]
    int nb_mu[LPB200_MAX_CONF];
    int n_conf_disp;
    const int32_t *bidx_disp;
    int nb_disp[LPB200_MAX_CONF];
    const double *pi_pred;   // G x P column-major
    int P;
};

struct DevModel {
    int mu_model, disp_model, drop_model;
    int p_mu, p_disp, q_drop;
    int n_theta;             // length of the BFGS parameter vector of fitMuDispGene
    int nbf_mu, nbf_disp;    // sum nB_c (batch-factor columns incl. leading ones)
};

// one gene's natural-scale parameters (after the reference clamps).  61 doubles: the odd
// 8-byte stride (122 words) keeps lane-per-point shared-memory reads bank-conflict free.
struct PointNat {
    double disp[LP_MAX_PDISP];
    double mu[LP_MAX_PMU];
    double bdisp[LP_MAX_BDISP];
    double bmu[LP_MAX_BMU];
    double pad;
};

LP_HD int lp_n_mu_params(int mu_model, int k_mu, int n_groups) {
    switch (mu_model) {
    case LPB200_MU_CONSTANT: return 1;
    case LPB200_MU_IMPULSE: return 7;
    case LPB200_MU_SPLINES: return k_mu;
    case LPB200_MU_GROUPS: return n_groups;
    default: return -1;
    }
}
LP_HD int lp_n_disp_params(int disp_model, int k_disp, int n_groups) {
    switch (disp_model) {
    case LPB200_DISP_CONSTANT: return 1;
    case LPB200_DISP_SPLINES: return k_disp;
    case LPB200_DISP_GROUPS: return n_groups;
    default: return -1;
    }
}

// compact "nat vector" of one gene: [disp p_disp | bdisp nbf_disp | mu p_mu | bmu nbf_mu]
LP_HD int lp_n_nat(const DevModel &M) { return M.p_disp + M.nbf_disp + M.p_mu + M.nbf_mu; }
LP_HD void lp_nat_load(const DevModel &M, const double *v, PointNat &g) {
    int u = 0;
    for (int k = 0; k < M.p_disp; k++) g.disp[k] = v[u++];
    for (int k = 0; k < M.nbf_disp; k++) g.bdisp[k] = v[u++];
    for (int k = 0; k < M.p_mu; k++) g.mu[k] = v[u++];
    for (int k = 0; k < M.nbf_mu; k++) g.bmu[k] = v[u++];
}
LP_HD void lp_nat_store(const DevModel &M, const PointNat &g, double *v) {
    int u = 0;
    for (int k = 0; k < M.p_disp; k++) v[u++] = g.disp[k];
    for (int k = 0; k < M.nbf_disp; k++) v[u++] = g.bdisp[k];
    for (int k = 0; k < M.p_mu; k++) v[u++] = g.mu[k];
    for (int k = 0; k < M.nbf_mu; k++) v[u++] = g.bmu[k];
}
