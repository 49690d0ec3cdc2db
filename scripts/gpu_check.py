"""Development check on a GPU box: CUDA path vs the CPU oracle on C1-like data + timings."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.engine import Engine
from lineagepulse_b200.splines import ns
from lineagepulse_b200.simulate import simulateContinuousDataSet
from tests.oracle_binding import Oracle

def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)

def main():
    o = Oracle()
    e = Engine(0)
    print("fp64 peak TFLOP/s", e.fp64_peak_tflops(), flush=True)
    G0, N = 200, 200
    sim = simulateContinuousDataSet(N, 70, 65, 65, scaMumax=100, scaSDMuAmplitude=3,
                                    vecGeneWiseDropoutRates=np.full(G0, 0.1), seed=1)
    X = sim["counts"]; X = np.ascontiguousarray(X[X.sum(1) > 0], dtype=np.int32)
    t = sim["annot"]["continuous"]
    G = X.shape[0]
    d = A.DesignData(G, N, pseudotime=t, impulse_basis=ns(t, 4))
    e.load_counts_dense(X.astype(np.float64)); e.set_design(d)
    ctl = A.default_control()
    print("rowmeans", np.abs(e.row_means() - o.row_means(X, d)).max())
    mA = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mB = A.Model(A.MU_IMPULSE, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mC = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_NONE, 0)
    mD = A.Model(A.MU_IMPULSE, A.DISP_CONSTANT, A.DROP_NONE, 0)
    # eval parity at init params and at random params
    for m in (mA, mB, mC, mD):
        p0 = o.init_params(X, d, m)
        pg = e.init_params(m)
        print("init diff", np.abs(p0.mat_mu - pg.mat_mu).max(), np.abs(p0.mat_disp - pg.mat_disp).max())
        if p0.mat_drop is not None:
            p0.mat_drop[:, 0] = np.random.default_rng(3).normal(-2, 1, N); 
        rng = np.random.default_rng(5)
        if m.mu_model == A.MU_IMPULSE:
            p0.mat_mu[:, 0:2] = rng.uniform(0.5, 20, (G, 2)); p0.mat_mu[:, 5:7] = np.sort(rng.uniform(0, 1, (G, 2)), 1)
            p0.mat_mu[:, 2:5] *= rng.uniform(0.3, 3, (G, 3))
        p0.mat_disp[:, 0] = rng.uniform(0.3, 10, G)
        ll_o = o.eval_loglik(X, d, m, p0); ll_g = e.eval_loglik(m, p0)
        print("eval model", m.mu_model, m.drop_model, "max rel", rel(ll_g, ll_o).max())
    pk_o, vl_o = o.impulse_starts(X, d); pk_g, vl_g = e.impulse_starts()
    print("starts max rel", rel(pk_g, pk_o).max(), rel(vl_g, vl_o).max())
    # fits
    pA_o = A.ParamsData(d, mA); rA_o = o.fit_model(X, d, mA, pA_o, True, ctl)
    pA_g = A.ParamsData(d, mA); t0 = time.time(); rA_g = e.fit_model(mA, pA_g, True, ctl); tA = time.time() - t0
    print("A: n_iter", rA_o["n_iter"], rA_g["n_iter"], "em", rA_o["em_ll"][:3], rA_g["em_ll"][:3], "time", tA)
    print("A: ll rel", rel(rA_g["ll"], rA_o["ll"]).max(), "drop diff", np.abs(pA_g.mat_drop - pA_o.mat_drop).max(),
          "mu rel", rel(pA_g.mat_mu, pA_o.mat_mu).max(), "disp rel", rel(pA_g.mat_disp, pA_o.mat_disp).max())
    pB_o = A.ParamsData(d, mB); pB_o.mat_drop = pA_o.mat_drop.copy(); rB_o = o.fit_model(X, d, mB, pB_o, False, ctl)
    pB_g = A.ParamsData(d, mB); pB_g.mat_drop = pA_o.mat_drop.copy()
    e.reset_stats(); t0 = time.time(); rB_g = e.fit_model(mB, pB_g, False, ctl); tB = time.time() - t0
    r = rel(rB_g["ll"], rB_o["ll"])
    print("B: time", tB, e.stats(), "oracle evals", rB_o["fn_evals"])
    print("B: ll rel max", r.max(), "median", np.median(r), "frac<1e-6", (r < 1e-6).mean(), "frac<1e-4", (r < 1e-4).mean(),
          "sum ll", rB_g["ll"].sum(), rB_o["ll"].sum())
    pC_o = A.ParamsData(d, mC); rC_o = o.fit_model(X, d, mC, pC_o, False, ctl)
    pC_g = A.ParamsData(d, mC); rC_g = e.fit_model(mC, pC_g, False, ctl)
    print("C: ll rel", rel(rC_g["ll"], rC_o["ll"]).max(), "mu rel", rel(pC_g.mat_mu, pC_o.mat_mu).max())
    pD_o = A.ParamsData(d, mD); rD_o = o.fit_model(X, d, mD, pD_o, False, ctl)
    pD_g = A.ParamsData(d, mD); rD_g = e.fit_model(mD, pD_g, False, ctl)
    r = rel(rD_g["ll"], rD_o["ll"])
    print("D: ll rel max", r.max(), "median", np.median(r), "frac<1e-6", (r < 1e-6).mean())
    # table vs no table determinism check
    z_o = o.post_drop(X, d, mB, pB_o, [0, 5, 100]); z_g = e.post_drop(mB, pB_o, [0, 5, 100])
    print("postdrop", np.nanmax(np.abs(z_o - z_g)))
    # mid-size timing: 512 genes x 10000 cells impulse ZINB
    G2, N2 = 592, 10000
    sim = simulateContinuousDataSet(N2, G2 // 2, G2 // 4, G2 - G2 // 2 - G2 // 4, scaMumax=100, scaSDMuAmplitude=3,
                                    vecGeneWiseDropoutRates=np.full(G2, 0.1), seed=2)
    X2 = np.ascontiguousarray(sim["counts"], dtype=np.int32); t2 = sim["annot"]["continuous"]
    d2 = A.DesignData(G2, N2, pseudotime=t2, impulse_basis=ns(t2, 4))
    e2 = Engine(0); e2.load_counts_i32(X2); e2.set_design(d2)
    pA2 = A.ParamsData(d2, mA); t0 = time.time(); rA2 = e2.fit_model(mA, pA2, True, ctl); print("mid A time", time.time() - t0, rA2["n_iter"], e2.stats(), flush=True)
    pB2 = A.ParamsData(d2, mB); pB2.mat_drop = pA2.mat_drop.copy(); e2.reset_stats()
    t0 = time.time(); rB2 = e2.fit_model(mB, pB2, False, ctl); tB2 = time.time() - t0
    st = e2.stats(); print("mid B time", tB2, st, "OLE/s", st["ole"] / (st["kernel_ms"] * 1e-3), flush=True)
    # oracle on 16 genes of the same for parity/timing
    sub = np.arange(0, G2, G2 // 16)[:16]
    d2s = A.DesignData(len(sub), N2, pseudotime=t2, impulse_basis=ns(t2, 4))
    pBs = A.ParamsData(d2s, mB); pBs.mat_drop = pA2.mat_drop.copy()
    t0 = time.time(); rBs = o.fit_model(X2[sub], d2s, mB, pBs, False, ctl); tO = time.time() - t0
    r = rel(rB2["ll"][sub], rBs["ll"])
    print("mid oracle 16 genes time", tO, "evals", rBs["fn_evals"], "ll rel", r)

if __name__ == "__main__":
    main()
