// lp_bfgs.cuh -- stats::optim(method="BFGS", control=list(fnscale=-1)) as a RESUMABLE state
// machine.  The reference calls optim() without a gradient (fitMeanDispersion.R:367-383,
// fitDropout.R:279-294), i.e. R's vmmin (J.C. Nash, Compact Numerical Methods, Alg. 21) with
// central differences of step ndeps.  Results are optimiser-path dependent, so the
// control flow below follows vmmin decision by decision; it is cut at every objective evaluation
// so that (a) a CTA can evaluate the requested points cooperatively between two advance() calls
// and (b) thousands of per-cell problems can run in lock-step with one batched evaluation kernel
// (and one all-reduce) per round.
#pragma once
#include "lp_common.cuh"

#define LP_STEPREDN 0.2
#define LP_ACCTOL 0.0001
#define LP_RELTEST 10.0

enum { LP_REQ_NONE = 0, LP_REQ_F = 1, LP_REQ_GRAD = 2 };
enum { LP_PH_F0 = 0, LP_PH_G0 = 1, LP_PH_LS = 2, LP_PH_G = 3, LP_PH_DONE = 4 };

template <int NP>
struct BfgsState {
    int n, maxit;
    double reltol, ndeps;
    double b[NP];   // current trial point (what the evaluator reads)
    double X[NP], c[NP], g[NP], t[NP];
    double B[NP * (NP + 1) / 2];  // lower triangle of the inverse-Hessian approximation
    double f, Fmin, gradproj, step;
    int count, funcount, gradcount, ilast, iter, accpoint;
    int phase, req;
    int fail;            // optim's convergence: 0, 1 (maxit), 1001 = R would have raised an error
    unsigned int evals;  // objective evaluations incl. the gradient's
};

template <int NP>
LP_HD void bfgs_start(BfgsState<NP> &s, int n, const double *theta0, int maxit, double reltol, double ndeps) {
    s.n = n;
    s.maxit = maxit;
    s.reltol = reltol;
    s.ndeps = ndeps;
    for (int i = 0; i < n; i++) s.b[i] = theta0[i];
    s.f = 0.0;
    s.Fmin = 0.0;
    s.count = 0;
    s.funcount = s.gradcount = 0;
    s.ilast = 0;
    s.iter = 0;
    s.accpoint = 0;
    s.fail = 0;
    s.evals = 0;
    s.gradproj = 0.0;
    s.step = 1.0;
    s.phase = LP_PH_F0;
    s.req = LP_REQ_F;
}

// coordinate / offset of gradient evaluation point k (k = 0 .. 2n-1): fmingr of optim.c
LP_HD int bfgs_grad_coord(int k) { return k >> 1; }
LP_HD double bfgs_grad_point(double bi, int k, double eps) { return (k & 1) ? LP_ADD(bi, -eps) : LP_ADD(bi, eps); }

// ll[] are objective values to MAXIMISE (fnscale = -1): ll[0] for LP_REQ_F, ll[2i], ll[2i+1] =
// value at b + eps e_i, b - eps e_i for LP_REQ_GRAD.  Returns with s.req set to the next request
// (or LP_REQ_NONE when s.phase == LP_PH_DONE).
template <int NP>
LP_HD void bfgs_advance(BfgsState<NP> &s, const double *ll) {
    const int n = s.n;
    enum { L_TOP, L_LS_TRY, L_LS_CONT, L_LS_END, L_BOTTOM };
    int lbl = L_TOP;

    if (s.phase == LP_PH_F0) {
        double f = ll[0] / (-1.0);
        s.evals += 1;
        if (!isfinite(f)) {  // "initial value in 'vmmin' is not finite"
            s.fail = 1001;
            s.phase = LP_PH_DONE;
            s.req = LP_REQ_NONE;
            return;
        }
        s.f = f;
        s.Fmin = f;
        s.funcount = s.gradcount = 1;
        s.phase = LP_PH_G0;
        s.req = LP_REQ_GRAD;
        return;
    } else if (s.phase == LP_PH_G0 || s.phase == LP_PH_G) {
        bool bad = false;
        double gn[NP];
        for (int i = 0; i < n; i++) {
            double v1 = ll[2 * i] / (-1.0), v2 = ll[2 * i + 1] / (-1.0);
            gn[i] = (v1 - v2) / (2.0 * s.ndeps);
            if (!isfinite(gn[i])) bad = true;  // "non-finite finite-difference value [i]"
        }
        s.evals += 2u * (unsigned)n;
        if (bad) {
            s.fail = 1001;
            s.phase = LP_PH_DONE;
            s.req = LP_REQ_NONE;
            return;
        }
        if (s.phase == LP_PH_G0) {
            for (int i = 0; i < n; i++) s.g[i] = gn[i];
            s.iter = 1;
            s.ilast = s.gradcount;
            lbl = L_TOP;
        } else {
            for (int i = 0; i < n; i++) s.g[i] = gn[i];
            s.gradcount++;
            s.iter++;
            double D1 = 0.0;
            for (int i = 0; i < n; i++) {
                s.t[i] = LP_MUL(s.step, s.t[i]);
                s.c[i] = LP_ADD(s.g[i], -s.c[i]);
                D1 = LP_ADD(D1, LP_MUL(s.t[i], s.c[i]));
            }
            if (D1 > 0) {
                double D2 = 0.0;
                for (int i = 0; i < n; i++) {
                    double sum = 0.0;
                    for (int j = 0; j <= i; j++) sum = LP_ADD(sum, LP_MUL(s.B[i * (i + 1) / 2 + j], s.c[j]));
                    for (int j = i + 1; j < n; j++) sum = LP_ADD(sum, LP_MUL(s.B[j * (j + 1) / 2 + i], s.c[j]));
                    s.X[i] = sum;
                    D2 = LP_ADD(D2, LP_MUL(sum, s.c[i]));
                }
                D2 = 1.0 + D2 / D1;
                for (int i = 0; i < n; i++)
                    for (int j = 0; j <= i; j++) {
                        double num = LP_ADD(LP_ADD(LP_MUL(LP_MUL(D2, s.t[i]), s.t[j]), -LP_MUL(s.X[i], s.t[j])),
                                            -LP_MUL(s.t[i], s.X[j]));
                        s.B[i * (i + 1) / 2 + j] = LP_ADD(s.B[i * (i + 1) / 2 + j], num / D1);
                    }
            } else {
                s.ilast = s.gradcount;
            }
            lbl = L_BOTTOM;
        }
    } else if (s.phase == LP_PH_LS) {
        s.f = ll[0] / (-1.0);
        s.evals += 1;
        s.funcount++;
        s.accpoint = isfinite(s.f) && (s.f <= LP_ADD(s.Fmin, LP_MUL(LP_MUL(s.gradproj, s.step), LP_ACCTOL)));
        if (!s.accpoint) s.step *= LP_STEPREDN;
        lbl = L_LS_CONT;
    } else {
        s.req = LP_REQ_NONE;
        return;
    }

    for (;;) {
        if (lbl == L_TOP) {
            if (s.ilast == s.gradcount) {
                for (int i = 0; i < n; i++) {
                    for (int j = 0; j < i; j++) s.B[i * (i + 1) / 2 + j] = 0.0;
                    s.B[i * (i + 1) / 2 + i] = 1.0;
                }
            }
            for (int i = 0; i < n; i++) {
                s.X[i] = s.b[i];
                s.c[i] = s.g[i];
            }
            double gp = 0.0;
            for (int i = 0; i < n; i++) {
                double sum = 0.0;
                for (int j = 0; j <= i; j++) sum = LP_ADD(sum, -LP_MUL(s.B[i * (i + 1) / 2 + j], s.c[j]));
                for (int j = i + 1; j < n; j++) sum = LP_ADD(sum, -LP_MUL(s.B[j * (j + 1) / 2 + i], s.c[j]));
                s.t[i] = sum;
                gp = LP_ADD(gp, LP_MUL(sum, s.c[i]));
            }
            s.gradproj = gp;
            if (gp < 0.0) {  // downhill
                s.step = 1.0;
                s.accpoint = 0;
                lbl = L_LS_TRY;
            } else {  // uphill: reset unless it has just been reset
                s.count = 0;
                if (s.ilast == s.gradcount) s.count = n;
                else s.ilast = s.gradcount;
                lbl = L_BOTTOM;
            }
        } else if (lbl == L_LS_TRY) {
            int cnt = 0;
            for (int i = 0; i < n; i++) {
                s.b[i] = LP_ADD(s.X[i], LP_MUL(s.step, s.t[i]));
                if (LP_ADD(LP_RELTEST, s.X[i]) == LP_ADD(LP_RELTEST, s.b[i])) cnt++;  // no change
            }
            s.count = cnt;
            if (cnt < n) {
                s.phase = LP_PH_LS;
                s.req = LP_REQ_F;
                return;
            }
            lbl = L_LS_CONT;
        } else if (lbl == L_LS_CONT) {
            lbl = (s.count == n || s.accpoint) ? L_LS_END : L_LS_TRY;
        } else if (lbl == L_LS_END) {
            // abstol = -Inf: (f > abstol) is false only for f = -Inf / NaN
            bool enough = (s.f > -INFINITY) && fabs(s.f - s.Fmin) > s.reltol * (fabs(s.Fmin) + s.reltol);
            if (!enough) {
                s.count = n;
                s.Fmin = s.f;
            }
            if (s.count < n) {  // making progress: new gradient at b
                s.Fmin = s.f;
                s.phase = LP_PH_G;
                s.req = LP_REQ_GRAD;
                return;
            }
            if (s.ilast < s.gradcount) {  // no progress: restart once from steepest descent
                s.count = 0;
                s.ilast = s.gradcount;
            }
            lbl = L_BOTTOM;
        } else {  // L_BOTTOM
            bool stop = false;
            if (s.iter >= s.maxit) stop = true;
            if (!stop) {
                if (s.gradcount - s.ilast > 2 * n) s.ilast = s.gradcount;  // periodic restart
                if (s.count != n || s.ilast != s.gradcount) {
                    lbl = L_TOP;
                    continue;
                }
            }
            s.fail = (s.iter < s.maxit) ? 0 : 1;
            if (!isfinite(s.Fmin)) s.fail = 1001;
            s.phase = LP_PH_DONE;
            s.req = LP_REQ_NONE;
            return;
        }
    }
}
