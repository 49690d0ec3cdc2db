// lp_eval.cuh -- the pieces of the fit kernel's objective evaluation that do not depend on the thread layout:
// per-observation arithmetic of the specialised evaluator, the (x, phi)-table rule, and the construction of
// the evaluation points of one optimiser request.  LP_HD: lp_kernels.cuh runs them on the device, and
// tests/csrc/mirror_harness.cpp (test infrastructure) runs the same source on the host under LP_LIBM_MIRROR
// while restating the kernel's lane / slot / warp layout and reduction order -- the GPU tests then require
// the kernel and that mirror to agree bit for bit on every (gene, start) fit (theta, log-likelihood,
// convergence code, evaluation count).
#pragma once
#include "lp_bfgs.cuh"
#include "lp_model.cuh"

#define LP_KC 32          // evaluation points per chunk (one warp row)
#ifndef LP_TAB
#define LP_TAB 1024       // entries of one (x, phi) table
#endif
#define LP_NTAB 3         // tables per pass: base phi, phi + eps, phi - eps
#ifdef LP_VW_OVERRIDE     // timing experiments only (profiles/README.md); changes the summation order
#define LP_VW LP_VW_OVERRIDE
#else
#define LP_VW 32          // virtual warps of the evaluation layout (see lp_kernels.cuh): fixes the summation order
#endif

#define LP_MU_FAST_CONST 0
#define LP_MU_FAST_IMPULSE 1
#define LP_MU_FAST_GROUPS 2     // group means + mean batch factors (C4)
#define LP_MU_FAST_SPLINES 3    // spline mean model + mean batch factors (C5)
#define LP_MU_FAST_CONSTCOV 4   // constant mean + mean batch factors (the H0 fits of C4)

// entries of each (x, phi) table of a pass over N cells with nk points, or 0: tabulate when that is cheaper
// than N x nk evaluations of lp_nb_xphi_part
LP_HD int lp_table_len(int use_table, int ntabs, int xmax, int nk, int N) {
    if (!use_table || ntabs <= 0) return 0;
    const int want = (xmax + 1 < LP_TAB) ? xmax + 1 : LP_TAB;
    return ((long long)ntabs * want * 2 <= (long long)nk * N) ? want : 0;
}

// point kk of an optimiser request at b (LP_REQ_F: b itself; LP_REQ_GRAD: b +- ndeps e_c, fmingr's order):
// natural-scale parameters in out; returns the (x, phi) table the point reads (0 base, 1 / 2: phi + / - eps)
LP_HD int lp_request_point(const DevDesign &D, const DevModel &M, const double *b, int n, int req, int kk, double ndeps,
                           bool scalar_phi, PointNat &out) {
    double th[LPB200_MAX_PARAMS];
    for (int i = 0; i < n; i++) th[i] = b[i];
    int tsel = 0;
    if (req == LP_REQ_GRAD) {
        const int c = bfgs_grad_coord(kk);
        th[c] = bfgs_grad_point(b[c], kk, ndeps);
        if (scalar_phi && c == 0) tsel = 1 + (kk & 1);  // the dispersion coordinate
    }
    lp_unpack_theta(D, M, th, out);
    return tsel;
}
LP_HD double lp_base_phi(const double *b) { return lp_clamp(lp_exp(b[0]), LP_CLAMP_LO, LP_CLAMP_HI); }
LP_HD int lp_request_ntabs(int req, int base) { return (req == LP_REQ_F) ? 1 : (base == 0 ? 3 : 1); }

#ifdef __CUDACC__
static __host__ __device__ __noinline__ double lp_nb_xphi_part_cold(double x, double phi) { return lp_nb_xphi_part(x, phi); }
#else
inline double lp_nb_xphi_part_cold(double x, double phi) { return lp_nb_xphi_part(x, phi); }
#endif

// the lane's point of the specialised evaluator, in registers
struct FastPoint {
    double phi, logphi, p0, b2, h0, h1, h2, t1, t2, inv_h1;
};
template <int MU>
LP_HD void lp_fast_point(const PointNat &P, FastPoint &F) {
    F.phi = P.disp[0];
    F.logphi = lp_log(F.phi);
    F.p0 = P.mu[0];
    F.b2 = F.h0 = F.h1 = F.h2 = F.t1 = F.t2 = F.inv_h1 = 0.0;
    if (MU == LP_MU_FAST_IMPULSE) {
        F.b2 = P.mu[1]; F.h0 = P.mu[2]; F.h1 = P.mu[3]; F.h2 = P.mu[4]; F.t1 = P.mu[5]; F.t2 = P.mu[6];
        F.inv_h1 = 1.0 / F.h1;
    }
}

// log-likelihood of observation x (>= 0) of cell j under the lane's point; the per-cell vectors are read only
// where the term needs them (tt: impulse; pi_ptr: ZINB zero; l1m_ptr: ZINB non-zero)
// ... given the cell's mean mu (before the size factor)
template <bool ZINB>
LP_HD double lp_fast_cell_mu(const FastPoint &F, double mu, int x, int j, const double *sf, const double *pi_ptr,
                             const double *l1m_ptr, const double *tabp, int tab_len) {
    const double mus = sf ? mu * sf[j] : mu;
    if (x == 0) return ZINB ? lp_ll_zero_zinb(mus, F.phi, pi_ptr[j]) : lp_ll_zero_nb(mus, F.phi, F.logphi);
    const double xd = (double)x;
    const double xphi = (x < tab_len) ? tabp[x] : lp_nb_xphi_part_cold(xd, F.phi);  // cold: keeps the table code out of the loop body
    return lp_ll_nonzero(xd, xphi, mus, F.phi, ZINB ? l1m_ptr[j] : 0.0);
}
template <int MU, bool ZINB>
LP_HD double lp_fast_cell(const FastPoint &F, int x, int j, const double *tt, const double *sf, const double *pi_ptr,
                          const double *l1m_ptr, const double *tabp, int tab_len) {
    double mu;
    if (MU == LP_MU_FAST_IMPULSE) {
        const double t = tt[j];
        double s1, s2;
        lp_sigmoid_pair(-F.p0 * (t - F.t1), F.b2 * (t - F.t2), s1, s2);
        mu = F.inv_h1 * (F.h0 + (F.h1 - F.h0) * s1) * (F.h2 + (F.h1 - F.h2) * s2);
    } else {
        mu = F.p0;
    }
    return lp_fast_cell_mu<ZINB>(F, mu, x, j, sf, pi_ptr, l1m_ptr, tabp, tab_len);
}
