"""Synthetic data with the recipe of ``simulateContinuousDataSet`` (simulateDataSet.R:70-304).

R's RNG streams are not reproducible outside R, so the draws come from
``numpy.random.Generator(PCG64(seed))``; parity between implementations is judged on identical
*inputs*, not identical RNG.  Genes are generated in chunks so that large
configurations (20k x 100k) never hold a dense float64 matrix.
"""
from __future__ import annotations

import numpy as np

SCA_MIN_MU = 1e-5  # simulateDataSet.R:84


def _impulse(t, beta, t1, t2, h0, h1, h2):
    """The simulator's single-beta impulse (simulateDataSet.R:85-88)."""
    return (1.0 / h1) * (h0 + (h1 - h0) / (1.0 + np.exp(-beta * (t - t1)))) * \
        (h2 + (h1 - h2) / (1.0 + np.exp(beta * (t - t2))))


def simulate_hidden_means(rng, n_cells, n_const, n_lin, n_imp, mumax=100.0, sd_amp=1.0):
    """Per-gene trajectory parameters (simulateDataSet.R:157-201); evaluated lazily per chunk."""
    pt = np.arange(n_cells, dtype=np.float64) / max(n_cells - 1, 1)  # seq(0,1,by=1/(N-1)) :123
    mu_const = np.maximum(rng.uniform(size=n_const) * mumax, SCA_MIN_MU)
    lin_start = np.maximum(rng.uniform(size=n_lin) * mumax, SCA_MIN_MU)
    lin_end = np.maximum(lin_start * np.abs(1 + rng.normal(1.0, sd_amp, size=n_lin)), SCA_MIN_MU)
    beta = rng.uniform(size=n_imp) * 2 + 0.5
    ta, tb = rng.uniform(size=n_imp), rng.uniform(size=n_imp)
    t1, t2 = np.minimum(ta, tb), np.maximum(ta, tb)
    h0 = rng.uniform(size=n_imp) * mumax
    h1 = h0 * np.abs(1 + rng.normal(0.0, sd_amp, size=n_imp))
    h2 = h0 * np.abs(1 + rng.normal(0.0, sd_amp, size=n_imp))
    h0, h1, h2 = (np.maximum(h, SCA_MIN_MU) for h in (h0, h1, h2))
    return dict(pt=pt, mu_const=mu_const, lin_start=lin_start, lin_end=lin_end,
                beta=beta, t1=t1, t2=t2, h0=h0, h1=h1, h2=h2)


def hidden_mean_rows(hid, lo, hi):
    """Rows [lo, hi) of matMuHidden (const genes, then linear, then impulse)."""
    pt = hid["pt"]
    nc, nl = len(hid["mu_const"]), len(hid["lin_start"])
    rows = []
    for g in range(lo, hi):
        if g < nc:
            rows.append(np.full_like(pt, hid["mu_const"][g]))
        elif g < nc + nl:
            k = g - nc
            rows.append(hid["lin_start"][k] + pt / pt.max() * (hid["lin_end"][k] - hid["lin_start"][k]))
        else:
            k = g - nc - nl
            rows.append(_impulse(pt, hid["beta"][k], hid["t1"][k], hid["t2"][k],
                                 hid["h0"][k], hid["h1"][k], hid["h2"][k]))
    return np.stack(rows, 0) if rows else np.zeros((0, len(pt)))


def simulateContinuousDataSet(scaNCells, scaNConst, scaNLin, scaNImp, scaMumax=100.0,
                              scaSDMuAmplitude=1.0, vecNormConstExternal=None, vecDispExternal=None,
                              vecGeneWiseDropoutRates=None, matDropoutModelExternal=None,
                              seed=0, dtype=np.int32, chunk=512):
    """Returns ``{"annot": {"cell", "continuous"}, "counts": G x N integer array}``.

    Same argument names and checks as the reference; ``seed`` is the only addition.
    """
    G = scaNConst + scaNLin + scaNImp
    for name, v in (("vecDispExternal", vecDispExternal), ("vecGeneWiseDropoutRates", vecGeneWiseDropoutRates)):
        if v is not None and len(v) != G:
            raise ValueError(f"scaNConst + scaNLin + scaNImp has to be length({name})")
    if matDropoutModelExternal is not None and vecGeneWiseDropoutRates is not None:
        raise ValueError("Supply either matDropoutModelExternal or vecGeneWiseDropoutRates.")
    if matDropoutModelExternal is None and vecGeneWiseDropoutRates is None:
        raise ValueError("Supply either matDropoutModelExternal or vecGeneWiseDropoutRates.")
    if matDropoutModelExternal is not None:
        # the reference's shape checks (simulateDataSet.R:111-119) make this path unusable
        raise NotImplementedError("matDropoutModelExternal: only gene-wise drop-out rates are usable "
                                  "in the reference (simulateDataSet.R:111-119)")
    rng = np.random.Generator(np.random.PCG64(seed))
    hid = simulate_hidden_means(rng, scaNCells, scaNConst, scaNLin, scaNImp, scaMumax, scaSDMuAmplitude)
    if vecNormConstExternal is None:
        sf = np.ones(scaNCells)
    else:
        sf = np.asarray(vecNormConstExternal, dtype=np.float64)
        if len(sf) != scaNCells:
            raise ValueError("vecNormConstExternal has to be of the length of the number of cells scaNCells.")
    if vecDispExternal is None:
        disp = np.maximum(3 + rng.normal(0.0, 0.5, size=G), 0.05)  # :235-236
    else:
        disp = np.asarray(vecDispExternal, dtype=np.float64)
    rates = np.asarray(vecGeneWiseDropoutRates, dtype=np.float64)
    counts = np.empty((G, scaNCells), dtype=dtype)
    for lo in range(0, G, chunk):
        hi = min(G, lo + chunk)
        mu = hidden_mean_rows(hid, lo, hi) * sf[None, :]
        phi = disp[lo:hi, None]
        # rnbinom(mu=, size=): gamma-Poisson mixture
        lam = rng.gamma(shape=phi, scale=mu / phi)
        x = rng.poisson(lam)
        drop = rng.uniform(size=x.shape) < rates[lo:hi, None]  # rbinom(1, 1, rate) :285-297
        x[drop] = 0
        counts[lo:hi] = x.astype(dtype)
    cells = np.array([f"cell_{j + 1}" for j in range(scaNCells)])
    genes = np.array([f"gene_{i + 1}" for i in range(G)])
    return {"annot": {"cell": cells, "continuous": hid["pt"]}, "counts": counts, "genes": genes,
            "hidden": {"disp": disp, **hid}}


# ---------------------------------------------------------------------------------------------
# Benchmark workloads generated ON THE DEVICE, any gene range at a time (bench.py: the strong-scaling
# 20 000 x 100 000 matrix is 2e9 draws -- minutes of numpy and 8 GB of host memory per rank otherwise).
# Same recipe as simulateContinuousDataSet (+ group / batch effects for the C4 shape); the
# draws of genes [16 c, 16 c + 16) come from a generator seeded with (seed, c), so any sharding that starts on
# multiples of 16 genes sees the same matrix.
# ---------------------------------------------------------------------------------------------
def workload_recipe(n_genes, seed, mumax=100.0, sd_amp=3.0, dropout=0.1, n_groups=0, n_batches=0):
    """Per-gene hidden parameters (host, O(G)): 50 % constant / 25 % linear / 25 % impulse trajectories in a fixed
    random gene order (contiguous gene shards then carry the same mix), dispersions 3 + N(0, 0.5); with
    ``n_groups``: half the genes get group means m exp(N(0, 0.5)); with ``n_batches``: batch factors exp(N(0, 0.3))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_const, n_lin = n_genes // 2, n_genes // 4
    n_imp = n_genes - n_const - n_lin
    hid = simulate_hidden_means(rng, 2, n_const, n_lin, n_imp, mumax, sd_amp)
    kind = np.r_[np.zeros(n_const, np.int8), np.ones(n_lin, np.int8), np.full(n_imp, 2, np.int8)]
    par = np.zeros((n_genes, 6))
    par[:n_const, 0] = hid["mu_const"]
    par[n_const:n_const + n_lin, 0] = hid["lin_start"]
    par[n_const:n_const + n_lin, 1] = hid["lin_end"]
    for k, nm in enumerate(("beta", "t1", "t2", "h0", "h1", "h2")):
        par[n_const + n_lin:, k] = hid[nm]
    disp = np.maximum(3 + rng.normal(0.0, 0.5, size=n_genes), 0.05)
    order = rng.permutation(n_genes)
    rec = {"n_genes": n_genes, "seed": int(seed), "kind": kind[order], "par": par[order], "disp": disp[order],
           "dropout": float(dropout), "group_fac": None, "batch_fac": None}
    if n_groups:
        gf = np.ones((n_genes, n_groups))
        affected = rng.uniform(size=n_genes) < 0.5
        gf[affected] = np.exp(rng.normal(0.0, 0.5, size=(int(affected.sum()), n_groups)))
        rec["group_fac"] = gf
    if n_batches:
        bf = np.exp(rng.normal(0.0, 0.3, size=(n_genes, n_batches)))
        bf[:, 0] = 1.0
        rec["batch_fac"] = bf
    return rec


def simulate_rows_device(rec, n_cells, lo, hi, device, group_idx=None, batch_idx=None, dtype=None):
    """Rows [lo, hi) of the workload's count matrix as an int32 torch tensor on ``device`` (lo a multiple of 16)."""
    import torch
    assert lo % 16 == 0
    dev = torch.device(device)
    pt = torch.arange(n_cells, dtype=torch.float64, device=dev) / max(n_cells - 1, 1)
    out = torch.empty((hi - lo, n_cells), dtype=dtype or torch.int32, device=dev)
    gidx = None if group_idx is None else torch.as_tensor(np.asarray(group_idx), dtype=torch.long, device=dev)
    bidx = None if batch_idx is None else torch.as_tensor(np.asarray(batch_idx), dtype=torch.long, device=dev)
    gen = torch.Generator(device=dev)
    for c0 in range(lo, hi, 16):
        c1 = min(hi, c0 + 16)
        kind = torch.as_tensor(rec["kind"][c0:c1].astype(np.int64), device=dev)[:, None]
        par = torch.as_tensor(rec["par"][c0:c1], dtype=torch.float64, device=dev)
        a, b, c, d, e, f = (par[:, k:k + 1] for k in range(6))
        const = a.expand(-1, n_cells)
        lin = a + pt[None, :] * (b - a)
        # impulse (simulateDataSet.R:85-88): par = beta, t1, t2, h0, h1, h2; guarded for the non-impulse rows
        h1 = torch.where(kind == 2, e, torch.ones_like(e))
        imp = (1.0 / h1) * (d + (h1 - d) / (1.0 + torch.exp(-a * (pt[None, :] - b)))) * (f + (h1 - f) / (1.0 + torch.exp(a * (pt[None, :] - c))))
        mu = torch.where(kind == 0, const, torch.where(kind == 1, lin, imp))
        if rec["group_fac"] is not None and gidx is not None:
            mu = mu * torch.as_tensor(rec["group_fac"][c0:c1], dtype=torch.float64, device=dev)[:, gidx]
        if rec["batch_fac"] is not None and bidx is not None:
            mu = mu * torch.as_tensor(rec["batch_fac"][c0:c1], dtype=torch.float64, device=dev)[:, bidx]
        mu = torch.clamp(mu, min=SCA_MIN_MU)
        phi = torch.as_tensor(rec["disp"][c0:c1], dtype=torch.float64, device=dev)[:, None]
        gen.manual_seed(rec["seed"] * 1000003 + c0 // 16)
        lam = torch._standard_gamma(phi.expand(-1, n_cells).contiguous(), generator=gen) * (mu / phi)   # rnbinom(mu, size)
        x = torch.poisson(lam, generator=gen)
        keep = torch.rand(x.shape, generator=gen, device=dev, dtype=torch.float64) >= rec["dropout"]
        out[c0 - lo:c1 - lo] = (x * keep).to(out.dtype)
    return out
