"""The .Call routines of r_shim/lp_rshim.c restated over the CPU oracle (test infrastructure).

R is not installed here; tests/r_eval.py can nevertheless EXECUTE r_shim/R/lpb200_wrappers.R if something answers its
.Call()s.  This module does, with the argument and result conventions of lp_rshim.c (named design list, 1-based index
vectors, integer(4) model codes, parameter list with cbind-ed batch matrices, list(params, conv, ll) results ...), and
the CPU oracle in place of the CUDA library -- so that a test can run the reference's runLineagePulse with the wrappers'
fitLPModels / fitModel / fitMuDisp / fitPi / evalLogLikMatrix substituted for the reference's and compare the results
table with the one of the unmodified reference.  The C shim itself is exercised on the GPU under a stand-in R runtime
(tests/test_gpu_parity.py::test_r_shim_runs_on_the_gpu); this checks the R text above it."""
import numpy as np

from lineagepulse_b200 import _cabi as A
from tests.r_eval import RError, RList, RVec


class ShimMock:
    def __init__(self, oracle):
        self.oracle = oracle
        self.calls = []
        self.control = A.default_control()
        self.contexts = []

    # -- conversions (lp_rshim.c: fill_design, model_of, params_of) ---------------------------------------------------
    @staticmethod
    def _get(lst, name):
        return lst.get(name) if isinstance(lst, RList) else None

    def _design(self, ctx, lst):
        g = lambda k: self._get(lst, k)     # noqa: E731
        G, N = int(g("n_genes").v[0]), int(g("n_cells").v[0])
        mat = lambda v: None if v is None else np.asarray(v.mat(), dtype=np.float64)     # noqa: E731
        conf = {}
        for tag in ("Mu", "Disp"):
            nb = g("vecNBatches" + tag)
            if nb is not None:
                idx = g("vecidxBatchAssign" + tag).v.astype(np.int64)
                assert len(idx) == len(nb) * N
                conf[tag] = [idx[c * N:(c + 1) * N] for c in range(len(nb))]
                for c, col in enumerate(conf[tag]):
                    assert col.min() >= 1 and col.max() == int(nb.v[c])
        groups = g("vecidxGroups")
        if groups is not None:
            assert groups.v.min() >= 1 and groups.v.max() <= int(g("scaNGroups").v[0])
        ctx["design"] = A.DesignData(
            G, N, size_factors=None if g("vecNormConst") is None else g("vecNormConst").v.astype(np.float64),
            pseudotime=None if g("vecContinuousCovar") is None else g("vecContinuousCovar").v.astype(np.float64),
            groups=None if groups is None else groups.v.astype(np.int64), spline_basis_mu=mat(g("matSplineBasisMu")),
            spline_basis_disp=mat(g("matSplineBasisDisp")), impulse_basis=mat(g("matImpulseBasis")),
            confounders_mu=conf.get("Mu"), confounders_disp=conf.get("Disp"), pi_const_pred=mat(g("matPiConstPredictors")))

    @staticmethod
    def _model(v):
        m = [int(x) for x in v]
        assert len(m) == 4
        return A.Model(*m)

    def _params(self, ctx, model, lst):
        d = ctx["design"]
        p = A.ParamsData(d, model)
        take = lambda k: None if self._get(lst, k) is None else np.asfortranarray(self._get(lst, k).mat(), dtype=np.float64)   # noqa: E731
        p.mat_mu, p.mat_disp = take("matMuModel"), take("matDispModel")
        assert p.mat_mu.shape[0] == d.n_genes and p.mat_disp.shape[0] == d.n_genes
        for field, key, nbs in (("batch_mu", "matBatchMu", d.n_batches_mu), ("batch_disp", "matBatchDisp", d.n_batches_disp)):
            m = take(key)
            if nbs:
                assert m is not None and m.shape == (d.n_genes, sum(nbs)), (key, None if m is None else m.shape, nbs)
                ends = np.cumsum(nbs)
                setattr(p, field, [np.asfortranarray(m[:, e - n:e]) for n, e in zip(nbs, ends)])
        if model.drop_model != A.DROP_NONE:
            p.mat_drop = take("matDropoutLinModel")
        return p

    @staticmethod
    def _params_out(lst, p, d):
        out = RList(lst.v, lst.names)

        def put(key, m):
            if m is not None and out.get(key) is not None:
                old = out.get(key)
                out.v[out.names.index(key)] = RVec(np.asarray(m).reshape(-1, order="F"), dim=np.asarray(m).shape, dimnames=old.dimnames)
        put("matMuModel", p.mat_mu)
        put("matDispModel", p.mat_disp)
        if d.n_batches_mu:
            put("matBatchMu", np.concatenate(p.batch_mu, axis=1))
        if d.n_batches_disp:
            put("matBatchDisp", np.concatenate(p.batch_disp, axis=1))
        put("matDropoutLinModel", p.mat_drop)
        return out

    # -- the routines -------------------------------------------------------------------------------------------------
    def __call__(self, routine, args):
        self.calls.append(routine)
        fn = getattr(self, routine, None)
        if fn is None:
            raise RError(f"C symbol name \"{routine}\" not in load table")
        want = {"lp_set_control": 1, "lp_create": 1, "lp_load_counts_csc": 5, "lp_load_counts_dense": 2, "lp_set_design": 2,
                "lp_eval_loglik": 4, "lp_fit_mudisp": 4, "lp_fit_pi": 4, "lp_fit_model": 6, "lp_fit_models": 4, "lp_lrt": 3}[routine]
        if len(args) != want:
            raise RError(f"Incorrect number of arguments ({len(args)}), expecting {want} for '{routine}'")
        return fn(*args)

    def lp_set_control(self, v):
        c = A.default_control()
        x = v.v.astype(np.float64)
        for k, (name, cast) in enumerate((("maxit_mudisp", int), ("reltol_mudisp", float), ("maxit_pi", int), ("reltol_pi", float),
                                          ("max_cycles", int))):
            if k < len(x) and x[k] == x[k]:
                setattr(c, name, cast(x[k]))
        self.control = c
        return None

    def lp_create(self, devices):
        assert devices.v.dtype.kind == "i" and len(devices) >= 1
        ctx = {"devices": devices.v.tolist()}
        self.contexts.append(ctx)
        return RList([ctx], ["externalptr"])

    @staticmethod
    def _ctx(ptr):
        return ptr.v[0]

    def lp_load_counts_dense(self, ptr, mat):
        assert mat.dim is not None and mat.v.dtype.kind == "f"
        self._ctx(ptr)["counts"] = np.ascontiguousarray(np.nan_to_num(mat.mat(), nan=-1.0), dtype=np.int32)
        return None

    def lp_load_counts_csc(self, ptr, dim, p, i, x):
        G, N = (int(v) for v in dim.v)
        m = np.zeros((G, N), dtype=np.int32)
        for j in range(N):
            for k in range(int(p.v[j]), int(p.v[j + 1])):
                m[int(i.v[k]), j] = int(x.v[k])
        self._ctx(ptr)["counts"] = m
        return None

    def lp_set_design(self, ptr, lst):
        self._design(self._ctx(ptr), lst)
        return None

    def lp_eval_loglik(self, ptr, model, params, n_genes):
        ctx, m = self._ctx(ptr), self._model(model.v)
        p = self._params(ctx, m, params)
        ll = self.oracle.eval_loglik(ctx["counts"], ctx["design"], m, p)
        assert int(n_genes.v[0]) == len(ll)
        return RVec(ll)

    def lp_fit_mudisp(self, ptr, model, params, n_genes):
        ctx, m = self._ctx(ptr), self._model(model.v)
        p = self._params(ctx, m, params)
        r = self.oracle.fit_mudisp(ctx["counts"], ctx["design"], m, p, self.control)
        return RList([self._params_out(params, p, ctx["design"]), RVec(r["conv"].astype(np.int64)), RVec(r["ll"])])

    def lp_fit_pi(self, ptr, model, params, n_out):
        ctx, m = self._ctx(ptr), self._model(model.v)
        p = self._params(ctx, m, params)
        r = self.oracle.fit_pi(ctx["counts"], ctx["design"], m, p, self.control)
        assert int(n_out.v[0]) == len(r["ll"])
        return RList([self._params_out(params, p, ctx["design"]), RVec(r["conv"].astype(np.int64)), RVec(r["ll"])])

    def lp_fit_model(self, ptr, model, params, fit_drop, n_genes, keep_mask):
        ctx, m = self._ctx(ptr), self._model(model.v)
        if int(keep_mask.v[0]) != 0:
            raise RError("the oracle-backed stand-in has no warm starts")
        p = self._params(ctx, m, params)
        r = self.oracle.fit_model(ctx["counts"], ctx["design"], m, p, bool(fit_drop.v[0]), self.control)
        em = np.full(self.control.max_cycles, np.nan)
        em[: r["n_iter"]] = r["em_ll"][: r["n_iter"]]
        ll = self.oracle.eval_loglik(ctx["counts"], ctx["design"], m, p)
        return RList([self._params_out(params, p, ctx["design"]), RVec(em), RVec(ll), RVec(np.asarray(r["conv"], dtype=np.int64)),
                      RVec(np.array([float(r["n_iter"]), float(bool(r["converged"])), float(r["ll_init"])]))])

    def lp_fit_models(self, ptr, models, params, n_genes):
        ctx = self._ctx(ptr)
        G = int(n_genes.v[0])
        codes = models.mat().astype(np.int64)
        assert codes.shape == (4, len(params))
        outs, ll_init, em, convd, lls, convs = [], [], [], [], [], []
        for j, plist in enumerate(params.v):
            m = self._model(codes[:, j])
            p = self._params(ctx, m, plist)
            r = self.oracle.fit_model(ctx["counts"], ctx["design"], m, p, False, self.control)
            outs.append(self._params_out(plist, p, ctx["design"]))
            ll_init.append(float(r["ll_init"]))
            em.append(float(r["em_ll"][0]))
            convd.append(int(bool(r["converged"])))
            lls.append(self.oracle.eval_loglik(ctx["counts"], ctx["design"], m, p))
            convs.append(np.asarray(r["conv"], dtype=np.int64))
        assert all(len(v) == G for v in lls)
        return RList([RList(outs), RVec(np.array(ll_init)), RVec(np.array(em)), RVec(np.array(convd, dtype=np.int64)),
                      RVec(np.concatenate(lls)), RVec(np.concatenate(convs))])

    def lp_lrt(self, ll_full, ll_red, ddf):
        p, padj = self.oracle.lrt(ll_full.v.astype(np.float64), ll_red.v.astype(np.float64), int(ddf.v[0]))
        return RList([RVec(p), RVec(padj)])
