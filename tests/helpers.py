"""Shared test helpers: synthetic inputs of SURVEY.md 8(d) and comparison utilities."""
import json
import os

import numpy as np

from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.simulate import simulateContinuousDataSet
from lineagepulse_b200.splines import ns

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def make_dataset(n_cells, n_const, n_lin, n_imp, seed, mumax=100, sd=3, drop=0.1, size_factors=None):
    G = n_const + n_lin + n_imp
    sim = simulateContinuousDataSet(n_cells, n_const, n_lin, n_imp, scaMumax=mumax, scaSDMuAmplitude=sd,
                                    vecNormConstExternal=size_factors,
                                    vecGeneWiseDropoutRates=np.full(G, drop), seed=seed)
    X = sim["counts"]
    keep = X.sum(1) > 0
    X = np.ascontiguousarray(X[keep], dtype=np.int32)
    t = sim["annot"]["continuous"]
    return {"X": X, "t": t, "G": X.shape[0], "N": X.shape[1], "sf": size_factors, "sim": sim, "keep": keep}


def design_for(ds, mu="impulse", groups=None, conf_mu=None, conf_disp=None, k_mu=6, k_disp=3, pred=None,
               disp="constant"):
    t = ds["t"]
    return A.DesignData(
        ds["G"], ds["N"], size_factors=ds.get("sf"), pseudotime=t, groups=groups,
        spline_basis_mu=ns(t, k_mu) if mu == "splines" else None,
        spline_basis_disp=ns(t, k_disp) if disp == "splines" else None,
        impulse_basis=ns(t, 4), confounders_mu=conf_mu, confounders_disp=conf_disp, pi_const_pred=pred)


def model(mu, disp="constant", drop="none", group="PerCell"):
    return A.Model(A.MU_MODELS[mu], A.DISP_MODELS[disp], A.DROP_MODELS[drop], A.DROP_FIT_GROUPS[group])


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def random_params(oracle, X, d, m, seed=5):
    """Plausible (not fitted) parameters for fixed-parameter likelihood comparisons."""
    rng = np.random.default_rng(seed)
    p = oracle.init_params(X, d, m)
    G, N = d.n_genes, d.n_cells
    if m.mu_model == A.MU_IMPULSE:
        p.mat_mu[:, 0:2] = rng.uniform(0.5, 20, (G, 2))
        p.mat_mu[:, 5:7] = np.sort(rng.uniform(0, 1, (G, 2)), 1)
        p.mat_mu[:, 2:5] *= rng.uniform(0.3, 3, (G, 3))
    elif m.mu_model == A.MU_SPLINES:
        p.mat_mu += rng.normal(0, 0.2, p.mat_mu.shape)
    else:
        p.mat_mu *= rng.uniform(0.5, 2, p.mat_mu.shape)
    if m.disp_model == A.DISP_SPLINES:
        p.mat_disp = np.asfortranarray(rng.normal(0.5, 0.3, p.mat_disp.shape))
    else:
        p.mat_disp = np.asfortranarray(rng.uniform(0.3, 10, p.mat_disp.shape))
    p.batch_mu = [np.asfortranarray(np.c_[np.ones(G), rng.uniform(0.5, 2, (G, b.shape[1] - 1))]) for b in p.batch_mu]
    p.batch_disp = [np.asfortranarray(np.c_[np.ones(G), rng.uniform(0.5, 2, (G, b.shape[1] - 1))]) for b in p.batch_disp]
    if p.mat_drop is not None:
        p.mat_drop[:, 0] = rng.normal(-2, 1, N)
        if m.drop_model == A.DROP_LOGISTIC_OFMU:
            p.mat_drop[:, 1] = -rng.uniform(0.2, 2, N)
            for k in range(2, p.mat_drop.shape[1]):
                p.mat_drop[:, k] = rng.normal(0, 0.3, N)
        else:
            for k in range(1, p.mat_drop.shape[1]):
                p.mat_drop[:, k] = rng.normal(0, 0.3, N)
    return p


def theta0_items(oracle, X, d, m, p0):
    """Start vectors of the fit kernel's work items as lpb200_fit_mudisp builds them (fitMeanDispersion.R:349-364,
    :621-663): theta = [log disp | log disp batch factors | mu coordinates | log mu batch factors], impulse fits
    with 3 starts per gene in the order Peak, Valley, Prior.  Returns (theta0 [G * n_starts, n], n_starts)."""
    G = d.n_genes
    disp_cols = p0.mat_disp if m.disp_model == A.DISP_SPLINES else np.log(p0.mat_disp)
    bdisp = [np.log(b[:, 1:]) for b in p0.batch_disp]
    bmu = [np.log(b[:, 1:]) for b in p0.batch_mu]
    if m.mu_model == A.MU_IMPULSE:
        peak, valley = oracle.impulse_starts(X, d)
        prior = p0.mat_mu.copy()
        prior[:, 2:5] = np.log(prior[:, 2:5])
        mus = [peak, valley, prior]
    elif m.mu_model == A.MU_SPLINES:
        mus = [p0.mat_mu]
    else:
        mus = [np.log(p0.mat_mu)]
    rows = []
    for i in range(G):
        for mu in mus:
            rows.append(np.concatenate([disp_cols[i]] + [b[i] for b in bdisp] + [mu[i]] + [b[i] for b in bmu]))
    return np.ascontiguousarray(rows, dtype=np.float64), len(mus)


def best_start(ll_items, n_starts):
    """Per gene: index and value of the first maximum over its starts (fitMeanDispersion.R:664-671)."""
    ll = np.asarray(ll_items).reshape(-1, n_starts)
    best = np.argmax(ll, axis=1)   # first maximum
    return best, ll[np.arange(len(ll)), best]


def edge_datasets():
    """Degenerate and ragged shapes for the fit path: one gene x five cells, a gene with a single non-zero
    observation, gene / cell counts that are not multiples of the 16-gene summation blocks, the 32-lane warps or
    the 64-element row pitch, counts beyond the (x, phi) table (1 024 entries) and beyond uint16 (int32 storage),
    NA observations (-1) and size factors.  Yields (tag, dataset)."""
    rng = np.random.default_rng(7)
    for G, N, single, big, na, sf in ((1, 5, False, False, False, False), (1, 7, True, False, False, False),
                                      (3, 33, False, False, True, True), (17, 97, False, True, False, False),
                                      (2, 64, False, True, True, True), (5, 1025, False, False, False, False)):
        t = np.linspace(0, 1, N)
        mu = rng.uniform(0.5, 60, (G, 1)) * (1 + 0.5 * np.sin(3 * t)[None, :])
        X = rng.negative_binomial(3, 3 / (3 + mu)).astype(np.int32)
        X[rng.uniform(size=X.shape) < 0.1] = 0
        if single:
            X[0, :] = 0
            X[0, N // 2] = 7
        if big:
            X[-1, 0] = 70000
            X[-1, 1] = 1500
        if na:
            X[0, 1] = -1
            X[G - 1, N - 1] = -1
        yield f"{G}x{N}", {"X": X, "t": t, "G": G, "N": N, "sf": rng.uniform(0.6, 1.5, N) if sf else None}
