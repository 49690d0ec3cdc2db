// lp_model.cuh -- parameter transforms and per-observation model expansion:
//   theta <-> natural scale with the reference clamps (fitMeanDispersion.R:78-200, :349-364, :391-498),
//   decompressMeansByGene / decompressDispByGene / decompressDropoutRateByGene for ONE cell
//   (decompressParameters.R:26-88,155-211,295-320).
#pragma once
#include "lp_math.cuh"

// theta = [disp params | disp batch logs | mu params | mu batch logs]  ->  natural scale
// (lp_exp: on the device bit-identical to the CUDA library's exp, on the host the C library's)
LP_HD void lp_unpack_theta(const DevDesign &D, const DevModel &M, const double *th, PointNat &g) {
    int u = 0;
    if (M.disp_model == LPB200_DISP_SPLINES) {  // raw coefficients, clamp +-10 log 10
        const double lim = 10.0 * lp_log(10.0);
        for (int k = 0; k < M.p_disp; k++) g.disp[k] = lp_clamp(th[u + k], -lim, lim);
    } else {
        for (int k = 0; k < M.p_disp; k++) g.disp[k] = lp_clamp(lp_exp(th[u + k]), LP_CLAMP_LO, LP_CLAMP_HI);
    }
    u += M.p_disp;
    int o = 0;
    for (int c = 0; c < D.n_conf_disp; c++) {
        g.bdisp[o] = 1.0;
        for (int b = 1; b < D.nb_disp[c]; b++) g.bdisp[o + b] = lp_clamp(lp_exp(th[u + b - 1]), LP_CLAMP_LO, LP_CLAMP_HI);
        u += D.nb_disp[c] - 1;
        o += D.nb_disp[c];
    }
    if (M.mu_model == LPB200_MU_IMPULSE) {
        for (int k = 0; k < 7; k++) g.mu[k] = th[u + k];
        for (int k = 2; k < 5; k++) g.mu[k] = lp_exp(g.mu[k]);
        for (int k = 0; k < 2; k++)
            if (g.mu[k] < LP_CLAMP_LO) g.mu[k] = LP_CLAMP_LO;
        for (int k = 2; k < 5; k++) g.mu[k] = lp_clamp(g.mu[k], LP_CLAMP_LO, LP_CLAMP_HI);
    } else if (M.mu_model == LPB200_MU_SPLINES) {
        for (int k = 0; k < M.p_mu; k++) g.mu[k] = lp_clamp(th[u + k], -LP_CLAMP_HI, LP_CLAMP_HI);
    } else {
        for (int k = 0; k < M.p_mu; k++) g.mu[k] = lp_clamp(lp_exp(th[u + k]), LP_CLAMP_LO, LP_CLAMP_HI);
    }
    u += M.p_mu;
    o = 0;
    for (int c = 0; c < D.n_conf_mu; c++) {
        g.bmu[o] = 1.0;
        for (int b = 1; b < D.nb_mu[c]; b++) g.bmu[o + b] = lp_clamp(lp_exp(th[u + b - 1]), LP_CLAMP_LO, LP_CLAMP_HI);
        u += D.nb_mu[c] - 1;
        o += D.nb_mu[c];
    }
}

// natural-scale guesses -> theta (fitMeanDispersion.R:349-364); impulse_coords: the 7 mu
// coordinates already in optimiser space (slots 3:5 logged, :621-622) or nullptr
LP_HD void lp_pack_theta(const DevDesign &D, const DevModel &M, const PointNat &g, const double *impulse_coords,
                         double *th) {
    int u = 0;
    for (int k = 0; k < M.p_disp; k++) th[u++] = (M.disp_model == LPB200_DISP_SPLINES) ? g.disp[k] : log(g.disp[k]);
    int o = 0;
    for (int c = 0; c < D.n_conf_disp; c++) {
        for (int b = 1; b < D.nb_disp[c]; b++) th[u++] = log(g.bdisp[o + b]);
        o += D.nb_disp[c];
    }
    if (M.mu_model == LPB200_MU_IMPULSE) {
        for (int k = 0; k < 7; k++) th[u++] = impulse_coords[k];
    } else {
        for (int k = 0; k < M.p_mu; k++) th[u++] = (M.mu_model == LPB200_MU_SPLINES) ? g.mu[k] : log(g.mu[k]);
    }
    o = 0;
    for (int c = 0; c < D.n_conf_mu; c++) {
        for (int b = 1; b < D.nb_mu[c]; b++) th[u++] = log(g.bmu[o + b]);
        o += D.nb_mu[c];
    }
}

// mean of cell j WITHOUT size factor
LP_HD double lp_mu_at(const DevDesign &D, const DevModel &M, const PointNat &g, int j) {
    double mu;
    if (M.mu_model == LPB200_MU_CONSTANT) {
        mu = g.mu[0];
    } else if (M.mu_model == LPB200_MU_IMPULSE) {
        mu = lp_impulse(g.mu, D.t[j]);
    } else if (M.mu_model == LPB200_MU_SPLINES) {
        double lin = 0.0;
        for (int k = 0; k < D.k_mu; k++) lin += D.bs_mu[j + (size_t)k * D.N] * g.mu[k];
        mu = lp_exp(lin);
    } else {
        mu = g.mu[D.group[j]];
    }
    int o = 0;
    for (int c = 0; c < D.n_conf_mu; c++) {
        mu = mu * g.bmu[o + D.bidx_mu[(size_t)c * D.N + j]];
        o += D.nb_mu[c];
    }
    return mu;
}

LP_HD double lp_disp_at(const DevDesign &D, const DevModel &M, const PointNat &g, int j) {
    double phi;
    if (M.disp_model == LPB200_DISP_CONSTANT) {
        phi = g.disp[0];
    } else if (M.disp_model == LPB200_DISP_SPLINES) {
        double lin = 0.0;
        for (int k = 0; k < D.k_disp; k++) lin += D.bs_disp[j + (size_t)k * D.N] * g.disp[k];
        phi = lp_exp(lin);
    } else {
        phi = g.disp[D.group[j]];
    }
    int o = 0;
    for (int c = 0; c < D.n_conf_disp; c++) {
        phi = phi * g.bdisp[o + D.bidx_disp[(size_t)c * D.N + j]];
        o += D.nb_disp[c];
    }
    return phi;
}

// linear predictor of the drop-out model of gene i at cell j; drop = N x q column-major,
// mu = mean without size factor (only read for logistic_ofMu)
LP_HD double lp_drop_lin(const DevDesign &D, const DevModel &M, const double *drop, int i, int j, double mu) {
    int k = 0;
    double lin = 1.0 * drop[j];
    k = 1;
    if (M.drop_model == LPB200_DROP_LOGISTIC_OFMU) {
        lin += lp_log(mu) * drop[j + (size_t)k * D.N];
        k++;
    }
    for (int p = 0; p < D.P; p++, k++) lin += D.pi_pred[i + (size_t)p * D.G] * drop[j + (size_t)k * D.N];
    return lin;
}

// does the dispersion reduce to one scalar per point? (constant model, no dispersion confounders)
LP_HD bool lp_disp_is_scalar(const DevDesign &D, const DevModel &M) {
    return M.disp_model == LPB200_DISP_CONSTANT && D.n_conf_disp == 0;
}

// full per-observation log-likelihood for the generic path.
//   x < 0: NA (contributes nothing);  s = size factor
//   pi_fixed / l1m_fixed: drop-out rate fixed during the fit (logistic), ignored otherwise
//   xphi_tab/tab_len: optional table of lp_nb_xphi_part(x, phi) for scalar phi (nullptr = none)
LP_HD double lp_cell_ll(const DevDesign &D, const DevModel &M, const PointNat &g, const double *drop, int gene,
                        int j, int x, double s, double pi_fixed, double l1m_fixed,
                        double phi_scalar, double logphi_scalar, const double *xphi_tab, int tab_len) {
    if (x < 0) return 0.0;
    double mu = lp_mu_at(D, M, g, j);
    double phi, logphi;
    if (lp_disp_is_scalar(D, M)) {
        phi = phi_scalar;
        logphi = logphi_scalar;
    } else {
        phi = lp_disp_at(D, M, g, j);
        logphi = lp_log(phi);
    }
    double pi = pi_fixed, l1m = l1m_fixed;
    if (M.drop_model == LPB200_DROP_LOGISTIC_OFMU) {
        pi = lp_dropout_rate(lp_drop_lin(D, M, drop, gene, j, mu));
        l1m = lp_log(1.0 - pi);
    }
    double mus = mu * s;
    if (x == 0) {
        if (M.drop_model == LPB200_DROP_NONE) return lp_ll_zero_nb(mus, phi, logphi);
        return lp_ll_zero_zinb(mus, phi, pi);
    }
    double xd = (double)x;
    if (lp_nb_small_x(xd, phi)) {   // dnbinom_mu's x < 1e-10 * size branch (phi beyond the 1e10 clamp of a scalar dispersion)
        double v = ((M.drop_model == LPB200_DROP_NONE) ? 0.0 : l1m) + lp_dnbinom_small_x(xd, mus, phi);
        return (v < LP_LL_FLOOR) ? LP_LL_FLOOR : v;
    }
    double xphi = (xphi_tab && x < tab_len) ? xphi_tab[x] : lp_nb_xphi_part(xd, phi);
    return lp_ll_nonzero(xd, xphi, mus, phi, (M.drop_model == LPB200_DROP_NONE) ? 0.0 : l1m);
}
