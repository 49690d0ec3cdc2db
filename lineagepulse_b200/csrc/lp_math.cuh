// lp_math.cuh -- scalar fp64 math of the likelihood: impulse model, logistic drop-out, and the
// per-observation NB / ZINB log-likelihood terms.  No fast-math; every function is LP_HD so the
// CPU-only test harness can run the same code.
#pragma once
#include "lp_common.cuh"
#include "lp_libm.cuh"

// evalImpulseModel (evalImpulseModel.R:23-37); p = (beta1, beta2, h0, h1, h2, t1, t2)
// The two sigmoids 1 / (1 + e): with both exponentials below 1e300 (not infinite, not NaN) the reciprocals
// take the check-free path of lp_libm.cuh.
LP_HD void lp_sigmoid_pair(double a1, double a2, double &s1, double &s2) {
    double e1, e2;
    lp_exp_pair(a1, a2, e1, e2);
    if (e1 < 1e300 && e2 < 1e300) {
        s1 = lp_rcp<true>(1.0 + e1);
        s2 = lp_rcp<true>(1.0 + e2);
    } else {
        s1 = lp_rcp<false>(1.0 + e1);
        s2 = lp_rcp<false>(1.0 + e2);
    }
}

LP_HD double lp_impulse(const double *p, double t) {
    double d1 = t - p[5];
    double d2 = t - p[6];
    double s1, s2;
    lp_sigmoid_pair(-p[0] * d1, p[1] * d2, s1, s2);
    return (1.0 / p[3]) * (p[2] + (p[3] - p[2]) * s1) * (p[4] + (p[3] - p[4]) * s2);
}

// evalDropoutModel (evalDropoutModel.R:24-32) given the linear predictor x'theta
LP_HD double lp_dropout_rate(double lin) {
    const double off = 0.001;
    return off + (1.0 - 2.0 * off) * 1.0 / (1.0 + lp_exp(-lin));
}

// ---- dnbinom(x, mu=, size=phi, log=TRUE) for x >= 1, restating R's nmath dnbinom_mu ->
// dbinom_raw (C. Loader's saddle-point form: stirlerr + bd0, no cancellation between large
// terms).  The reference's non-zero terms come from exactly this algorithm (evalLogLik.R:44-48,
// :136-140), and BFGS paths are sensitive to the last digits of the objective, so the device
// code follows the same operation sequence instead of an lgamma-difference formula.
#define LP_LN_SQRT_2PI 0.918938533204672741780329736406
#define LP_LN_2PI 1.837877066409345483560659472811
#define LP_SFERR_VALUES                                                                          \
    {0.0, 0.1534264097200273452913839, 0.08106146679532725821967026, 0.0548141210519176538961387, \
     0.04134069595540929409382208, 0.03316287351993628748511051, 0.02767792568499833914878929,    \
     0.02374616365629749597133028, 0.02079067210376509311152277, 0.01848845053267318523077936,    \
     0.01664469118982119216319487, 0.01513497322191737887351384, 0.01387612882307074799874573,    \
     0.01281046524292022692425066, 0.01189670994589177009505572, 0.01110455975820691732663076,    \
     0.01041126526197209649747857, 0.009799416126158803298390373, 0.009255462182712732917728637,  \
     0.008768700134139385462955047, 0.008330563433362871256469319, 0.007934114564314020547249562, \
     0.007573675487951840794972024, 0.007244554301320383179546197, 0.006942840107209529865664153, \
     0.006665247032707682442356181, 0.006408994188004207068439631, 0.006171712263039457647534605, \
     0.005951370112758847735624416, 0.005746216513010115682026102, 0.00555473355196280137103869}
static const double lp_sferr_host[31] = LP_SFERR_VALUES;
// 1/3, 1/5, ..., 1/19 (correctly rounded quotients) for the bd0 series; in __constant__ memory on the
// device for the same reason as the exp / log coefficients (lp_libm.cuh)
#define LP_ODD_RECIP_VALUES {1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0, 1.0 / 19.0}
static const double lp_odd_recip_host[9] = LP_ODD_RECIP_VALUES;
#ifdef __CUDACC__
static __constant__ double lp_sferr_dev[31] = LP_SFERR_VALUES;
static __constant__ double lp_odd_recip_dev[9] = LP_ODD_RECIP_VALUES;
#endif

// log1p and lgamma of the table-building code.  Not the CUDA library's: own routines over lp_log and IEEE
// operations only, so that every operation of the fit path is reproducible on the host (the mirror of
// tests/csrc/mirror_harness.cpp, which the GPU tests compare with the kernels bit for bit).
// log1p(a) by the classical correction log(u) a / (u - 1), u = fl(1 + a): a few ulp for every a > -1.
LP_HD double lp_log1p(double a) {
    const double u = 1.0 + a;
    if (u == 1.0) return a;
    return lp_log(u) * (a / (u - 1.0));
}
// lgamma(z) for 1 < z <= 16: the recurrence lifts the argument to w >= 17, where six Stirling terms leave less
// than 1e-17; absolute error <= 2e-14 against 40-digit arithmetic over the whole interval (as good as the lgamma
// difference R's stirlerr forms with it: the result is a difference of numbers of size ~30).
LP_HD double lp_lgamma_1_16(double z) {
    double prod = 1.0, w = z;
    while (w < 17.0) {
        prod *= w;
        w += 1.0;
    }
    const double iw = 1.0 / w, iw2 = iw * iw;
    const double ser = iw * (1.0 / 12.0 - iw2 * (1.0 / 360.0 - iw2 * (1.0 / 1260.0 - iw2 * (1.0 / 1680.0 - iw2 * (1.0 / 1188.0 - iw2 * (691.0 / 360360.0))))));
    return (w - 0.5) * lp_log(w) - w + LP_LN_SQRT_2PI + ser - lp_log(prod);
}

// stirlerr(n) = log(n!) - log(sqrt(2 pi n) (n/e)^n)
LP_HD double lp_stirlerr(double n) {
    if (n <= 15.0) {
        double nn = n + n;
        if (nn == (double)(int)nn) {
#ifdef __CUDA_ARCH__
            return lp_sferr_dev[(int)nn];
#else
            return lp_sferr_host[(int)nn];
#endif
        }
        return lp_lgamma_1_16(n + 1.0) - (n + 0.5) * lp_log(n) + n - LP_LN_SQRT_2PI;
    }
    const double S0 = 0.083333333333333333333, S1 = 0.00277777777777777777778, S2 = 0.00079365079365079365079365,
                 S3 = 0.000595238095238095238095238, S4 = 0.0008417508417508417508417508;
    double nn = n * n;
    if (n > 500.0) return (S0 - S1 / nn) / n;
    if (n > 80.0) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35.0) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

// bd0(x, np) = x log(x/np) + np - x.  For |x - np| < 0.1 (x + np) R sums the series
//   s + sum_j 2 x v^(2j+1) / (2j+1),  v = (x - np)/(x + np),
// until the sum stops changing.  With |v| < 0.1 the terms fall below half an ulp of s after 8 terms,
// so a fixed 9-term loop returns the value R's early-exit loop converges to; the divisions by
// (2j+1) become multiplications by the rounded reciprocals (<= 1 ulp of a term that is itself a
// small correction).  The early-exit / division form cost 29 % of the fit kernel's instructions
// (profiles/README.md) and diverged between lanes.
template <bool FAST>
LP_HD double lp_bd0_t(double x, double np) {
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = lp_div<FAST>(x - np, x + np);
        double s = (x - np) * v;
#if LP_LIBM_DEVICE && !defined(LP_NO_INTCMP)   /* the device and its host mirror */
        if ((__double2hiint(s) & 0x7ff00000) == 0) return s;   // |s| < 2^-1022: the same test on the exponent field (no fp64 compare)
#else
        if (fabs(s) < 2.2250738585072014e-308) return s;
#endif
        double ej = 2.0 * x * v;
        v = v * v;
        // explicit fma: the translation unit is compiled with -fmad=false (see the Makefile), these nine
        // are the only fused operations of the likelihood code
#ifdef __CUDA_ARCH__
#define LP_ODD_RECIP(k) lp_odd_recip_dev[k]
#else
#define LP_ODD_RECIP(k) lp_odd_recip_host[k]
#endif
        ej *= v; s = fma(ej, LP_ODD_RECIP(0), s);
        ej *= v; s = fma(ej, LP_ODD_RECIP(1), s);
        ej *= v; s = fma(ej, LP_ODD_RECIP(2), s);
        ej *= v; s = fma(ej, LP_ODD_RECIP(3), s);
        ej *= v; s = fma(ej, LP_ODD_RECIP(4), s);
        ej *= v; s = fma(ej, LP_ODD_RECIP(5), s);
        ej *= v; s = fma(ej, LP_ODD_RECIP(6), s);
        ej *= v; s = fma(ej, LP_ODD_RECIP(7), s);
        ej *= v; s = fma(ej, LP_ODD_RECIP(8), s);
#undef LP_ODD_RECIP
        return s;
    }
    return x * lp_logp<FAST>(lp_div<FAST>(x, np)) + np - x;
}
LP_HD double lp_bd0(double x, double np) { return lp_bd0_t<false>(x, np); }

// mu and phi both in (1e-100, 1e100): p, q >= 5e-201, n p and n q >= 5e-201, and every quotient of the NB / ZINB
// terms (p, q, x / np <= 2^31 / 5e-201, (phi + mu) / n >= 1e-200, (x - np) / (x + np) with 1 <= x < 2^31) lies
// well within [1e-290, 1e290]: the check-free division and logarithm apply
LP_HD bool lp_div_safe(double mu, double phi) {
#if LP_LIBM_DEVICE && !defined(LP_NO_INTCMP)   /* the device and its host mirror */
    // the interval test on the upper words, [2^-332, 2^331) inside (1e-100, 1e100): one integer subtract + compare per operand
    // instead of two fp64 compares (the fp64 pipe is the binding unit); negative, NaN and infinite operands fail it as they
    // fail the comparisons, and outside the interval the checked path returns the same bits
    const unsigned lo = 0x2b300000u, span = 0x54a00000u - 0x2b300000u;
    return ((unsigned)__double2hiint(mu) - lo) < span && ((unsigned)__double2hiint(phi) - lo) < span;
#else
    return mu > 1e-100 && mu < 1e100 && phi > 1e-100 && phi < 1e100;
#endif
}

// Part of dnbinom_mu that depends on (x, phi) only; tabulated per evaluation point when phi is
// one scalar per gene.  n = x + phi and xm = n - phi are formed as R forms them (so the loss of
// digits of n - phi at huge phi is reproduced).
// The two pieces that do not depend on x are arguments (logphi = lp_log(phi), sphi = lp_stirlerr(phi)): a table
// over x forms them once.
LP_HD double lp_nb_xphi_part_pre(double x, double phi, double logphi, double sphi) {
    double n = x + phi;
    double xm = n - phi;
    double lf = LP_LN_2PI + logphi + lp_log1p(-phi / n);
    return lp_log(phi / (phi + x)) + (lp_stirlerr(n) - sphi - lp_stirlerr(xm)) - 0.5 * lf;
}
LP_HD double lp_nb_xphi_part(double x, double phi) { return lp_nb_xphi_part_pre(x, phi, lp_log(phi), lp_stirlerr(phi)); }

// Part that depends on mu: -bd0(phi, n p) - bd0(n - phi, n q)
template <bool FAST>
LP_HD double lp_nb_mu_part_t(double x, double mu, double phi) {
    double n = x + phi;
    double p = lp_div<FAST>(phi, phi + mu), q = lp_div<FAST>(mu, phi + mu);
    if (!FAST && (p == 0.0 || q == 0.0)) return -INFINITY;  // dbinom_raw: R_D__0 (FAST: p, q >= 1e-300)
    return -lp_bd0_t<FAST>(phi, n * p) - lp_bd0_t<FAST>(n - phi, n * q);
}
LP_HD double lp_nb_mu_part(double x, double mu, double phi) {
    return lp_div_safe(mu, phi) ? lp_nb_mu_part_t<true>(x, mu, phi) : lp_nb_mu_part_t<false>(x, mu, phi);
}

// dnbinom_mu's other branch: for x < 1e-10 * size R does not go through dbinom_raw but evaluates
//   x log(mu / (1 + mu/size)) - mu - lgamma(x + 1) + log1p(x (x - 1) / (2 size))
// (with log(size / (1 + size/mu)) when size < mu).  Constant and group dispersions are clamped at 1e10 and
// x >= 1, so only products with dispersion batch factors, spline dispersions or stored parameters handed to
// evalLogLikMatrix reach it: a cold function, tested by the callers that can see such a phi.
LP_HD bool lp_nb_small_x(double x, double phi) { return x < 1e-10 * phi; }
#ifdef __CUDACC__
static __host__ __device__ __noinline__
#else
static inline
#endif
double lp_dnbinom_small_x(double x, double mu, double phi) {
    const double p = (phi < mu) ? lp_log(phi / (1.0 + phi / mu)) : lp_log(mu / (1.0 + mu / phi));
    return x * p - mu - lgamma(x + 1.0) + lp_log1p(x * (x - 1.0) / (2.0 * phi));
}

// Non-zero observation: dnbinom(x, mu, size=phi, log) [+ log(1-pi)], floored
// (evalLogLik.R:44-49 NB, :135-142 ZINB).  xphi = lp_nb_xphi_part(x, phi) (possibly from a
// table), l1mpi = log(1-pi) or 0 for NB.
LP_HD double lp_ll_nonzero(double x, double xphi, double mu, double phi, double l1mpi) {
    double v = l1mpi + (xphi + lp_nb_mu_part(x, mu, phi));
    return (v < LP_LL_FLOOR) ? LP_LL_FLOOR : v;
}

// Zero observation of the NB model: phi*(log(phi) - log(phi+mu)), floored (evalLogLik.R:40-42)
LP_HD double lp_ll_zero_nb(double mu, double phi, double logphi) {
    double v = phi * (logphi - lp_log(phi + mu));
    return (v < LP_LL_FLOOR) ? LP_LL_FLOOR : v;
}

// Zero observation of the ZINB model: log((1-pi)*(phi/(phi+mu))^phi + pi), floored
// (evalLogLik.R:125-132).  The ratio is formed exactly as the reference forms it, so the rounding
// of phi/(phi+mu) at huge phi is reproduced; the power is exp(phi*log(ratio)).
template <bool FAST>
LP_HD double lp_ll_zero_zinb_t(double mu, double phi, double pi) {
    double r = lp_div<FAST>(phi, phi + mu);
    double nb0 = lp_exp(phi * lp_logp<FAST>(r));
    double v = lp_log((1.0 - pi) * nb0 + pi);
    return (v < LP_LL_FLOOR) ? LP_LL_FLOOR : v;
}
LP_HD double lp_ll_zero_zinb(double mu, double phi, double pi) {
    return lp_div_safe(mu, phi) ? lp_ll_zero_zinb_t<true>(mu, phi, pi) : lp_ll_zero_zinb_t<false>(mu, phi, pi);
}

// posterior of drop-out at a zero (calcPostDrop.R:27-45)
LP_HD double lp_post_drop_zero(double mu, double phi, double pi) {
    double nb0 = lp_exp(phi * lp_log(phi / (phi + mu)));
    return pi / (pi + (1.0 - pi) * nb0);
}

LP_HD double lp_clamp(double v, double lo, double hi) {
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}
