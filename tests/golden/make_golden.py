"""Generates the golden fixtures in this directory with 50-digit mpmath arithmetic.

The reference ships no golden vectors and R cannot be run in the build image, so
these pins are ours: they state what the *mathematical* definitions used by the reference
(dnbinom, the ZINB zero term, pchisq, the impulse function, the logistic drop-out function)
evaluate to, independent of both the oracle and the CUDA code.

    python tests/golden/make_golden.py        # rewrites the *.json files next to this script
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))


def nb_logpmf(x, size, mu):
    x, size, mu = mp.mpf(x), mp.mpf(size), mp.mpf(mu)
    return (mp.loggamma(x + size) - mp.loggamma(size) - mp.loggamma(x + 1)
            + size * mp.log(size / (size + mu)) + x * mp.log(mu / (size + mu)))


def zinb_obs(x, mu, size, pi):
    """One term of evalLogLikZINB / evalLogLikNB (evalLogLik.R:30-53,107-149); pi < 0 => NB."""
    xm, mum, sm, pim = mp.mpf(x), mp.mpf(mu), mp.mpf(size), mp.mpf(pi)
    if pi >= 0:
        v = mp.log((1 - pim) * (sm / (sm + mum)) ** sm + pim) if x == 0 else mp.log(1 - pim) + nb_logpmf(x, size, mu)
    else:
        v = sm * (mp.log(sm) - mp.log(sm + mum)) if x == 0 else nb_logpmf(x, size, mu)
    return max(v, mp.mpf(-700))


def main():
    rng = np.random.default_rng(20171009)
    # 1. dnbinom(x, mu=, size=, log=TRUE): typical box + clamp extremes
    pts = []
    for _ in range(300):
        pts.append((float(np.floor(10 ** rng.uniform(0, 3.5))), 10 ** rng.uniform(-2, 3), 10 ** rng.uniform(-3, 4)))
    for _ in range(150):
        pts.append((float(np.floor(10 ** rng.uniform(0, 5))), 10 ** rng.uniform(-10, 10), 10 ** rng.uniform(-10, 7)))
    for size in (1e-10, 1e-5, 1.0, 1e5, 1e10):
        for mu in (1e-10, 1e-3, 1.0, 50.0, 1e5, 1e10):
            for x in (1.0, 7.0, 250.0):
                pts.append((x, size, mu))
    dn = [{"x": x, "size": s, "mu": m, "value": float(nb_logpmf(x, s, m))} for x, s, m in pts]
    json.dump(dn, open(os.path.join(HERE, "dnbinom_mu_log.json"), "w"), indent=0)
    # 2. per-observation ZINB / NB terms incl. the -700 floor
    obs = []
    for _ in range(300):
        mu, size = 10 ** rng.uniform(-2, 3.5), 10 ** rng.uniform(-1, 2)
        x = float(rng.poisson(min(mu, 1e4))) if rng.uniform() < 0.75 else 0.0
        pi = float(rng.choice([-1.0, 0.001, 0.1, 0.5, 0.98902, 0.999]))
        obs.append({"x": x, "mu": mu, "size": size, "pi": pi, "value": float(zinb_obs(x, mu, size, pi))})
    for x, mu, size, pi in ((0.0, 1e9, 50.0, -1.0), (3000.0, 1e-8, 2.0, 0.1), (0.0, 1e10, 1e10, 0.001), (5.0, 1e10, 1e-10, -1.0)):
        obs.append({"x": x, "mu": mu, "size": size, "pi": pi, "value": float(zinb_obs(x, mu, size, pi))})
    json.dump(obs, open(os.path.join(HERE, "zinb_obs.json"), "w"), indent=0)
    # 3. pchisq(q, df, lower.tail=FALSE)
    pc = []
    for df in (1, 2, 3, 5, 6, 7, 12):
        for q in (1e-3, 0.5, 3.0, 12.3, 40.0, 200.0, 1400.0):
            pc.append({"q": q, "df": df, "value": float(mp.gammainc(mp.mpf(df) / 2, mp.mpf(q) / 2, mp.inf, regularized=True))})
    json.dump(pc, open(os.path.join(HERE, "pchisq_upper.json"), "w"), indent=0)
    # 4. impulse function (evalImpulseModel.R:23-37) and drop-out function (evalDropoutModel.R:24-32)
    imp = []
    for _ in range(60):
        p = [rng.uniform(0.1, 30), rng.uniform(0.1, 30), 10 ** rng.uniform(-2, 3), 10 ** rng.uniform(-2, 3),
             10 ** rng.uniform(-2, 3), rng.uniform(-0.5, 1.5), rng.uniform(-0.5, 1.5)]
        t = rng.uniform(0, 1)
        b1, b2, h0, h1, h2, t1, t2 = map(mp.mpf, p)
        v = (1 / h1) * (h0 + (h1 - h0) / (1 + mp.exp(-b1 * (mp.mpf(t) - t1)))) * (h2 + (h1 - h2) / (1 + mp.exp(b2 * (mp.mpf(t) - t2))))
        imp.append({"par": p, "t": t, "value": float(v)})
    json.dump(imp, open(os.path.join(HERE, "impulse.json"), "w"), indent=0)
    dr = []
    for lin in (-800.0, -155.0, -5.0, 0.0, 0.3, 4.59511985013459, 40.0, 800.0):
        dr.append({"lin": lin, "value": float(mp.mpf("0.001") + mp.mpf("0.998") / (1 + mp.exp(-mp.mpf(lin))))})
    json.dump(dr, open(os.path.join(HERE, "dropout.json"), "w"), indent=0)
    # 5. a tiny gene: exact per-gene ZINB log-likelihood of an impulse model with size factors
    N = 40
    t = np.linspace(0, 1, N)
    sf = rng.uniform(0.5, 2.0, N)
    par = [3.0, 7.0, 20.0, 80.0, 5.0, 0.3, 0.7]
    phi = 2.5
    theta = rng.normal(-2, 1, N)
    xs = rng.poisson(30, N) * (rng.uniform(size=N) > 0.2)
    total = mp.mpf(0)
    for j in range(N):
        b1, b2, h0, h1, h2, t1, t2 = map(mp.mpf, par)
        mu = (1 / h1) * (h0 + (h1 - h0) / (1 + mp.exp(-b1 * (mp.mpf(t[j]) - t1)))) * (h2 + (h1 - h2) / (1 + mp.exp(b2 * (mp.mpf(t[j]) - t2))))
        pi = mp.mpf("0.001") + mp.mpf("0.998") / (1 + mp.exp(-mp.mpf(theta[j])))
        total += zinb_obs(int(xs[j]), mu * mp.mpf(sf[j]), phi, float(pi))
    json.dump({"t": t.tolist(), "sf": sf.tolist(), "par": par, "phi": phi, "theta": theta.tolist(),
               "x": [int(v) for v in xs], "loglik": float(total)}, open(os.path.join(HERE, "tiny_gene.json"), "w"))
    # 6. dnbinom beyond the 1e10 dispersion clamp (products with dispersion batch factors, spline dispersions):
    #    R's dnbinom_mu takes its x < 1e-10 * size branch here
    huge = []
    for size in (1.5e10, 1e11, 1e12, 1e15, 1e20):
        for mu in (1e-3, 0.7, 35.0, 4e3, 1e6):
            for x in (1.0, 2.0, 9.0, 140.0, 60000.0):
                huge.append({"x": x, "size": size, "mu": mu, "value": float(nb_logpmf(x, size, mu))})
    json.dump(huge, open(os.path.join(HERE, "dnbinom_huge_size.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
