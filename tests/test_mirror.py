"""The host mirror of fit_mudisp_kernel (tests/csrc/mirror_harness.cpp), CPU side.

The GPU tests (tests/test_gpu_parity.py::test_fit_kernel_equals_host_mirror_bitwise) establish kernel == mirror,
bit for bit.  Here, without a GPU:
  * the dumped MUFU.RCP64H table is sane and the mirror's exp / log (the device routines) are within 1 ulp of libm;
  * the committed GPU outputs of C1 (tests/golden/c1_fit_items.npz, written on a B200 by
    scripts/make_items_golden.py) are reproduced by the mirror bit for bit;
  * what separates the mirror -- hence the kernel -- from the oracle is arithmetic mode only (device exp / log
    instead of glibc's, 64 partial sums instead of R's long-double running sum): well-conditioned constant fits
    agree at 1e-8 (median 1e-15), and on the chaotic 3-start impulse fits the mirror-vs-oracle spread is the spread between
    the oracle's own two legitimate summation modes (long double vs double), with no bias in the mean.
"""
import os

import numpy as np
import pytest

from lineagepulse_b200 import _cabi as A
from tests.helpers import GOLDEN, best_start, design_for, edge_datasets, model, rel, theta0_items
from tests.mirror_harness import Mirror, rcp64h_table


@pytest.fixture(scope="module")
def mir():
    return Mirror()


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_rcp64h_table_and_device_libm_on_host(mir):
    tab = rcp64h_table().astype(np.int64)
    m = np.arange(1 << 20, dtype=np.int64)
    x = ((np.int64(0x3FF) << 52) | (m << 32)).view(np.float64)
    true_hi = (1.0 / x).view(np.int64) >> 32
    d = tab - true_hi
    assert tab[0] == 0x3FF00000 and np.all(np.diff(tab) <= 0)
    assert set(np.unique(d)) <= {0, 1} and 0.02 < (d == 1).mean() < 0.2   # a 20-bit seed, rounded up now and then
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 200000)), rng.uniform(0.5, 2, 200000), 1 + rng.normal(0, 1e-7, 50000)])
    _, l = mir.exp_log(x)
    assert np.all(np.abs(l - np.log(x)) <= np.spacing(np.abs(np.log(x))))
    xe = rng.uniform(-700, 700, 200000)
    e, _ = mir.exp_log(xe)
    assert np.all(np.abs(e - np.exp(xe)) <= np.spacing(np.exp(xe)))


def test_mirror_reproduces_the_committed_gpu_fits_bitwise(mir, oracle, c1):
    path = os.path.join(GOLDEN, "c1_fit_items.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c1_fit_items.npz not generated yet (scripts/make_items_golden.py on a B200)")
    z = np.load(path)
    d = design_for(c1)
    n_cases = 0
    for mu in ("impulse", "constant"):
        for drop in ("logistic", "none"):
            tag = f"{mu}_{drop}"
            m = model(mu, drop=drop)
            mat_drop = None
            if drop != "none":
                mat_drop = np.full((c1["N"], 1), -2.2)
            mr = mir.fit_items(d, m, c1["X"], z[tag + "_theta0"], int(z[tag + "_n_starts"]), drop=mat_drop)
            assert np.array_equal(mr["evals_items"], z[tag + "_evals"]), tag
            assert np.array_equal(mr["conv_items"], z[tag + "_conv"]), tag
            assert np.array_equal(_bits(mr["ll_items"]), _bits(z[tag + "_ll"])), tag
            assert np.array_equal(_bits(mr["theta"]), _bits(z[tag + "_theta"])), tag
            n_cases += 1
    assert n_cases == 4


def test_mirror_vs_oracle_is_arithmetic_mode_only(mir, oracle, c1):
    X = c1["X"]
    d = design_for(c1)
    # 1. constant fits (n = 2, well-conditioned): the mirror lands on the oracle's optimum
    for drop in ("logistic", "none"):
        m = model("constant", drop=drop)
        p0 = oracle.init_params(X, d, m)
        if drop != "none":
            p0.mat_drop[:, 0] = -2.2
        th0, ns = theta0_items(oracle, X, d, m, p0)
        mr = mir.fit_items(d, m, X, th0, ns, drop=p0.mat_drop)
        po = p0.copy()
        ro = oracle.fit_mudisp(X, d, m, po)
        assert np.array_equal(mr["conv_items"], ro["conv"])
        assert rel(mr["ll_items"], ro["ll"]).max() < 1e-8 and np.median(rel(mr["ll_items"], ro["ll"])) < 1e-14
        rp = rel(np.exp(mr["theta"]), np.c_[po.mat_disp[:, 0], po.mat_mu[:, 0]])
        assert np.quantile(rp, 0.99) < 1e-4 and rp.max() < 1e-3   # 1e-4 where the optimum is well-conditioned
    # 2. impulse fits (3 starts, n = 8): chaotic in the last digits of the objective.  Spread of the mirror against
    #    the oracle vs the spread between the oracle's own two summation modes (measured: ZINB 38 % / 76 % of genes
    #    within 1e-6 / 1e-4 for the mirror, 41 % / 78.5 % for the oracle against itself; NB 60 % / 89.5 % vs 61 % / 87 %).
    for drop in ("logistic", "none"):
        m = model("impulse", drop=drop)
        p0 = oracle.init_params(X, d, m)
        if drop != "none":
            p0.mat_drop[:, 0] = -2.2
        th0, ns = theta0_items(oracle, X, d, m, p0)
        mr = mir.fit_items(d, m, X, th0, ns, drop=p0.mat_drop)
        _, ll_m = best_start(mr["ll_items"], ns)
        lls = []
        try:
            for mode in (0, 1):
                oracle.lib.lporacle_set_sum_mode(mode)
                lls.append(oracle.fit_mudisp(X, d, m, p0.copy())["ll"])
        finally:
            oracle.lib.lporacle_set_sum_mode(0)
        ll_ld, ll_d = lls
        r_self = rel(ll_d, ll_ld)            # the reference algorithm against itself (double vs long double sums)
        r_mir = rel(ll_m, ll_ld)             # the kernel's arithmetic against the reference's
        f_self = [(r_self < t).mean() for t in (1e-6, 1e-4, 1e-3)]
        f_mir = [(r_mir < t).mean() for t in (1e-6, 1e-4, 1e-3)]
        # same agreement profile (each fraction is a binomial estimate over 200 genes: 3 sigma ~ 0.1)
        assert abs(f_mir[0] - f_self[0]) < 0.10 and abs(f_mir[1] - f_self[1]) < 0.08 and abs(f_mir[2] - f_self[2]) < 0.06, (drop, f_mir, f_self)
        assert np.median(r_mir) < 3 * np.median(r_self) + 1e-12
        # no bias: neither arithmetic finds systematically better optima (paired differences; the self-comparison gives the scale)
        dm, ds = ll_m - ll_ld, ll_d - ll_ld
        assert abs(dm.mean()) <= 3 * dm.std(ddof=1) / np.sqrt(len(dm)), (drop, dm.mean(), dm.std())
        assert abs(np.median(dm)) < 1e-6 * np.median(np.abs(ll_ld))
        assert dm.std() < 2 * ds.std() + 1e-9, (drop, dm.std(), ds.std())
        # what is not chaotic: every start ends with optim's code 0 / 1, and the total log-likelihood agrees
        assert np.all(np.isin(mr["conv_items"], (0, 1)))
        assert abs(ll_m.sum() - ll_ld.sum()) <= 1e-4 * abs(ll_ld.sum())


def test_mirror_on_degenerate_and_ragged_shapes(mir, oracle):
    """One gene x five cells, a single non-zero observation, shapes off every block / warp / pitch multiple, counts
    beyond the (x, phi) table and beyond uint16, NA observations with size factors (tests.helpers.edge_datasets):
    constant fits land on the oracle's optimum, every impulse start ends with optim's code 0 / 1 and a finite
    log-likelihood not below the constant model's by more than the optimiser's tolerance.  The GPU test
    test_edge_shapes_kernel_equals_mirror_bitwise runs the same inputs through the kernel."""
    n_cases = 0
    for tag, ds in edge_datasets():
        d = design_for(ds)
        for drop in ("none", "logistic"):
            ll = {}
            for mu in ("constant", "impulse"):
                m = model(mu, drop=drop)
                p0 = oracle.init_params(ds["X"], d, m)
                if drop != "none":
                    p0.mat_drop[:, 0] = -2.0
                th0, ns = theta0_items(oracle, ds["X"], d, m, p0)
                mr = mir.fit_items(d, m, ds["X"], th0, ns, drop=p0.mat_drop)
                assert np.all(np.isin(mr["conv_items"], (0, 1))) and np.isfinite(mr["ll_items"]).all(), (tag, mu, drop)
                _, ll[mu] = best_start(mr["ll_items"], ns)
                if mu == "constant":
                    po = p0.copy()
                    ro = oracle.fit_mudisp(ds["X"], d, m, po)
                    assert np.array_equal(mr["conv_items"], ro["conv"]), (tag, drop)
                    assert rel(ll[mu], ro["ll"]).max() < 1e-10, (tag, drop)
                n_cases += 1
            # the impulse model contains the constant one (h0 = h1 = h2): its best start cannot be worse
            assert np.all(ll["impulse"] >= ll["constant"] - 1e-6 * np.abs(ll["constant"])), (tag, drop)
    assert n_cases == 24
