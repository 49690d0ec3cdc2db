"""ctypes binding of the CPU oracle (oracle/liblporacle.so) -- test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from lineagepulse_b200 import _cabi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liblporacle.so")


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "lp_oracle.c")
    if force or not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "liblporacle.so"],
                              stdout=subprocess.DEVNULL)
    return ORACLE_LIB


class Oracle:
    def __init__(self, nthreads=None):
        self.lib = C.CDLL(build_oracle())
        self.nthreads = nthreads or min(os.cpu_count() or 1, 32)
        P = C.POINTER
        L = self.lib
        ip, dp = A.c_int32_p, A.c_double_p
        u64p = P(C.c_uint64)
        L.lporacle_row_means.argtypes = [ip, P(A.Design), dp]
        L.lporacle_init_params.argtypes = [ip, P(A.Design), P(A.Model), P(A.Params), C.c_double]
        L.lporacle_eval_loglik.argtypes = [ip, P(A.Design), P(A.Model), P(A.Params), dp, C.c_int]
        L.lporacle_impulse_starts.argtypes = [ip, P(A.Design), dp, dp, C.c_int]
        L.lporacle_fit_mudisp.argtypes = [ip, P(A.Design), P(A.Model), P(A.Params), P(A.Control), dp, ip, u64p, C.c_int]
        L.lporacle_fit_pi.argtypes = [ip, P(A.Design), P(A.Model), P(A.Params), P(A.Control), dp, ip, u64p, C.c_int]
        L.lporacle_fit_model.argtypes = [ip, P(A.Design), P(A.Model), P(A.Params), C.c_int, P(A.Control), C.c_double,
                                         dp, ip, ip, dp, dp, ip, u64p, C.c_int]
        L.lporacle_post_drop.argtypes = [ip, P(A.Design), P(A.Model), P(A.Params), ip, C.c_int, dp]
        L.lporacle_lrt.argtypes = [dp, dp, C.c_int, C.c_int, dp, dp]
        L.lporacle_expand_fits.argtypes = [ip, P(A.Design), P(A.Model), P(A.Params), ip, C.c_int, C.c_int, dp]
        L.lporacle_dnbinom_mu_log.argtypes = [C.c_double] * 3
        L.lporacle_dnbinom_mu_log.restype = C.c_double
        L.lporacle_pchisq_upper.argtypes = [C.c_double] * 2
        L.lporacle_pchisq_upper.restype = C.c_double
        L.lporacle_impulse_value.argtypes = [dp, C.c_double]
        L.lporacle_impulse_value.restype = C.c_double
        L.lporacle_dropout_rate.argtypes = [dp, dp, C.c_int]
        L.lporacle_dropout_rate.restype = C.c_double
        L.lporacle_ll_zinb_obs.argtypes = [C.c_double] * 4
        L.lporacle_ll_zinb_obs.restype = C.c_double
        L.lporacle_vmmin_rosenbrock.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp, P(C.c_int), P(C.c_int), P(C.c_int)]

    @staticmethod
    def _counts(counts):
        c = np.ascontiguousarray(counts, dtype=np.int32)
        return c, c.ctypes.data_as(A.c_int32_p)

    @staticmethod
    def _check(rc, allow=()):
        if rc != 0 and rc not in allow:
            raise RuntimeError(f"oracle returned {rc}")
        return rc

    # scalar helpers
    def dnbinom_mu_log(self, x, size, mu):
        return self.lib.lporacle_dnbinom_mu_log(float(x), float(size), float(mu))

    def pchisq_upper(self, q, df):
        return self.lib.lporacle_pchisq_upper(float(q), float(df))

    def impulse_value(self, par7, t):
        p = np.ascontiguousarray(par7, dtype=np.float64)
        return self.lib.lporacle_impulse_value(p.ctypes.data_as(A.c_double_p), float(t))

    def dropout_rate(self, theta, pred):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        pr = np.ascontiguousarray(pred, dtype=np.float64)
        return self.lib.lporacle_dropout_rate(th.ctypes.data_as(A.c_double_p), pr.ctypes.data_as(A.c_double_p), len(th))

    def ll_obs(self, x, mu, disp, pi=-1.0):
        return self.lib.lporacle_ll_zinb_obs(float(x), float(mu), float(disp), float(pi))

    def rosenbrock(self, use_numgrad, maxit=100, reltol=np.sqrt(np.finfo(float).eps)):
        par = np.zeros(2)
        val = C.c_double()
        fc, gc, cv = C.c_int(), C.c_int(), C.c_int()
        rc = self.lib.lporacle_vmmin_rosenbrock(int(use_numgrad), maxit, reltol, par.ctypes.data_as(A.c_double_p),
                                                C.byref(val), C.byref(fc), C.byref(gc), C.byref(cv))
        return dict(rc=rc, par=par, value=val.value, fncount=fc.value, grcount=gc.value, convergence=cv.value)

    # matrix-level
    def row_means(self, counts, design):
        c, cp = self._counts(counts)
        out = np.zeros(design.n_genes)
        self._check(self.lib.lporacle_row_means(cp, C.byref(design.struct), out.ctypes.data_as(A.c_double_p)))
        return out

    def init_params(self, counts, design, model, min_rowmean_global=float("nan"), with_drop=True):
        c, cp = self._counts(counts)
        p = A.ParamsData(design, model, with_drop=with_drop)
        ps = p.as_struct()
        self._check(self.lib.lporacle_init_params(cp, C.byref(design.struct), C.byref(model), C.byref(ps),
                                                  min_rowmean_global))
        p.sync_back()
        return p

    def eval_loglik(self, counts, design, model, params):
        c, cp = self._counts(counts)
        ll = np.zeros(design.n_genes)
        ps = params.as_struct()
        self._check(self.lib.lporacle_eval_loglik(cp, C.byref(design.struct), C.byref(model), C.byref(ps),
                                                  ll.ctypes.data_as(A.c_double_p), self.nthreads))
        return ll

    def impulse_starts(self, counts, design):
        c, cp = self._counts(counts)
        G = design.n_genes
        peak = np.zeros((G, 7), order="F")
        valley = np.zeros((G, 7), order="F")
        self._check(self.lib.lporacle_impulse_starts(cp, C.byref(design.struct), peak.ctypes.data_as(A.c_double_p),
                                                     valley.ctypes.data_as(A.c_double_p), self.nthreads))
        return peak, valley

    def fit_mudisp(self, counts, design, model, params, control=None, allow_nonfinite=False):
        c, cp = self._counts(counts)
        control = control or A.default_control()
        G = design.n_genes
        ll = np.zeros(G)
        conv = np.zeros(G, dtype=np.int32)
        ev = C.c_uint64(0)
        ps = params.as_struct()
        rc = self.lib.lporacle_fit_mudisp(cp, C.byref(design.struct), C.byref(model), C.byref(ps), C.byref(control),
                                          ll.ctypes.data_as(A.c_double_p), conv.ctypes.data_as(A.c_int32_p),
                                          C.byref(ev), self.nthreads)
        self._check(rc, allow=(A.E_NONFINITE,) if allow_nonfinite else ())
        params.sync_back()
        return dict(ll=ll, conv=conv, fn_evals=ev.value, rc=rc)

    def objective_mudisp(self, counts, design, model, params, gene, thetas):
        """evalLogLikMuDispGeneFit of one gene at the rows of thetas (optimiser coordinates)."""
        c, cp = self._counts(counts)
        th = np.ascontiguousarray(thetas, dtype=np.float64)
        out = np.zeros(th.shape[0])
        ps = params.as_struct()
        self._check(self.lib.lporacle_objective_mudisp(cp, C.byref(design.struct), C.byref(model), C.byref(ps), int(gene),
                                                       th.ctypes.data_as(A.c_double_p), th.shape[0],
                                                       out.ctypes.data_as(A.c_double_p)))
        return out

    def fit_pi(self, counts, design, model, params, control=None):
        c, cp = self._counts(counts)
        control = control or A.default_control()
        n = design.n_cells if model.drop_fit_group == A.DROPFIT_PERCELL else 1
        ll = np.zeros(n)
        conv = np.zeros(n, dtype=np.int32)
        ev = C.c_uint64(0)
        ps = params.as_struct()
        self._check(self.lib.lporacle_fit_pi(cp, C.byref(design.struct), C.byref(model), C.byref(ps), C.byref(control),
                                             ll.ctypes.data_as(A.c_double_p), conv.ctypes.data_as(A.c_int32_p),
                                             C.byref(ev), self.nthreads))
        params.sync_back()
        return dict(ll=ll, conv=conv, fn_evals=ev.value)

    def fit_model(self, counts, design, model, params, fit_drop, control=None, min_rowmean_global=float("nan")):
        c, cp = self._counts(counts)
        control = control or A.default_control()
        G = design.n_genes
        em = np.full(control.max_cycles, np.nan)
        n_iter, converged = C.c_int32(), C.c_int32()
        ll_init = C.c_double()
        ll = np.zeros(G)
        conv = np.zeros(G, dtype=np.int32)
        ev = C.c_uint64(0)
        ps = params.as_struct()
        self._check(self.lib.lporacle_fit_model(cp, C.byref(design.struct), C.byref(model), C.byref(ps), int(fit_drop),
                                                C.byref(control), min_rowmean_global,
                                                em.ctypes.data_as(A.c_double_p), C.byref(n_iter), C.byref(converged),
                                                C.byref(ll_init), ll.ctypes.data_as(A.c_double_p),
                                                conv.ctypes.data_as(A.c_int32_p), C.byref(ev), self.nthreads))
        params.sync_back()
        return dict(em_ll=em, n_iter=n_iter.value, converged=bool(converged.value), ll_init=ll_init.value,
                    ll=ll, conv=conv, fn_evals=ev.value)

    def post_drop(self, counts, design, model, params, gene_ids):
        c, cp = self._counts(counts)
        ids = np.ascontiguousarray(gene_ids, dtype=np.int32)
        z = np.zeros((len(ids), design.n_cells), order="F")
        ps = params.as_struct()
        self._check(self.lib.lporacle_post_drop(cp, C.byref(design.struct), C.byref(model), C.byref(ps),
                                                ids.ctypes.data_as(A.c_int32_p), len(ids),
                                                z.ctypes.data_as(A.c_double_p)))
        return z

    def expand_fits(self, counts, design, model, params, gene_ids, what):
        c, cp = self._counts(counts)
        ids = np.ascontiguousarray(gene_ids, dtype=np.int32)
        z = np.zeros((len(ids), design.n_cells), order="F")
        ps = params.as_struct()
        self._check(self.lib.lporacle_expand_fits(cp, C.byref(design.struct), C.byref(model), C.byref(ps),
                                                  ids.ctypes.data_as(A.c_int32_p), len(ids), int(what),
                                                  z.ctypes.data_as(A.c_double_p)))
        return z

    def lrt(self, ll_full, ll_red, ddf):
        f = np.ascontiguousarray(ll_full, dtype=np.float64)
        r = np.ascontiguousarray(ll_red, dtype=np.float64)
        p = np.zeros(len(f))
        padj = np.zeros(len(f))
        self._check(self.lib.lporacle_lrt(f.ctypes.data_as(A.c_double_p), r.ctypes.data_as(A.c_double_p), int(ddf),
                                          len(f), p.ctypes.data_as(A.c_double_p), padj.ctypes.data_as(A.c_double_p)))
        return p, padj
