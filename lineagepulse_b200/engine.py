"""Thin object wrapper over the C ABI (``include/lpb200.h``): one ``Engine`` = one
``lpb200_ctx`` = one GPU holding one gene shard of the count matrix.

There is no CPU fallback.  Constructing an ``Engine`` without the built CUDA library or without
a usable CUDA device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi as A


class LPB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lpb200 error {code}: {msg}")
        self.code = code


class Engine:
    """``Engine(device)``: one GPU.  ``Engine(devices=[0, 1, ...])``: one context over several GPUs of this
    process (lpb200_create_multi): same methods, whole-matrix arguments and results, genes dealt to the devices,
    NCCL inside the library -- the replacement of the reference's ``scaNProc`` worker pool."""

    def __init__(self, device: int = 0, library_path: str | None = None, devices=None):
        self.lib = A.load_library(library_path)
        if self.lib.lpb200_abi_version() != 1:
            raise RuntimeError("liblpb200.so ABI version mismatch")
        if devices is not None and len(devices) > 1:
            devs = np.ascontiguousarray(devices, dtype=np.int32)
            self.devices = [int(v) for v in devs]
            self.ctx = self.lib.lpb200_create_multi(devs.ctypes.data_as(A.c_int32_p), len(devs))
            if not self.ctx:
                raise LPB200Error(A.E_CUDA, f"cannot open CUDA devices {self.devices} as one context (no GPU / driver, "
                                             "or NCCL not found: set LPB200_NCCL_LIBRARY) -- this library has no CPU fallback")
        else:
            if devices is not None and len(devices) == 1:
                device = devices[0]
            self.devices = [int(device)]
            self.ctx = self.lib.lpb200_create(int(device))
            if not self.ctx:
                raise LPB200Error(A.E_CUDA, f"cannot open CUDA device {device} (no GPU / driver?) -- "
                                             "this library has no CPU fallback")
        self.design: A.DesignData | None = None
        self._cb = None
        self.n_genes = self.n_cells = 0

    # -- lifecycle -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.lpb200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            msg = self.lib.lpb200_last_error(self.ctx)
            raise LPB200Error(rc, msg.decode() if msg else "")
        return rc

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.lpb200_set_stream(self.ctx, C.c_void_p(cuda_stream)))

    def set_collective(self, rank: int, world: int, allreduce):
        """``allreduce(dev_ptr: int, count: int, dtype: int, stream: int) -> None`` sums ``count``
        8-byte words (dtype _cabi.F64 or _cabi.I64) in place across the gene shards (e.g.
        torch.distributed.all_reduce over NCCL)."""
        def _cb(user, ptr, count, dtype, stream):
            try:
                allreduce(int(ptr), int(count), int(dtype), int(stream or 0))
                return 0
            except Exception:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        self._cb = A.ALLREDUCE_FN(_cb) if world > 1 else A.ALLREDUCE_FN()
        self._check(self.lib.lpb200_set_collective(self.ctx, rank, world, self._cb, None))

    def init_comm(self, rank: int, world: int, unique_id: bytes | None):
        """One process per GPU: join the library's own NCCL communicator (lpb200_comm_init).  ``unique_id`` are the
        bytes of ``Engine.comm_unique_id()`` created on rank 0 and distributed by the host (see
        ``lineagepulse_b200.distributed.init_comm``)."""
        self._check(self.lib.lpb200_comm_init(self.ctx, int(rank), int(world), unique_id))

    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(A.UNIQUE_ID_BYTES)
        rc = self.lib.lpb200_comm_unique_id(buf)
        if rc != 0:
            raise LPB200Error(rc, "NCCL library not found (set LPB200_NCCL_LIBRARY)")
        return buf.raw

    def comm_stats(self):
        """(all-reduce calls, payload bytes) since the last reset_stats()."""
        calls, nbytes = C.c_uint64(), C.c_uint64()
        self._check(self.lib.lpb200_comm_stats(self.ctx, C.byref(calls), C.byref(nbytes)))
        return int(calls.value), int(nbytes.value)

    def shard_bounds(self):
        n = self.lib.lpb200_n_shards(self.ctx)
        b = np.zeros(n + 1, dtype=np.int32)
        self._check(self.lib.lpb200_shard_bounds(self.ctx, b.ctypes.data_as(A.c_int32_p)))
        return [int(v) for v in b]

    # -- data ------------------------------------------------------------------------------
    def load_counts_dense(self, x):
        """G x N matrix as R stores it (column-major doubles, NaN = NA)."""
        x = np.asfortranarray(x, dtype=np.float64)
        G, N = x.shape
        self._check(self.lib.lpb200_load_counts_dense(self.ctx, G, N, x.ctypes.data_as(A.c_double_p)))
        self.n_genes, self.n_cells = G, N

    def load_counts_csc(self, G, N, indptr, indices, data):
        p = np.ascontiguousarray(indptr, dtype=np.int32)
        i = np.ascontiguousarray(indices, dtype=np.int32)
        x = np.ascontiguousarray(data, dtype=np.float64)
        self._check(self.lib.lpb200_load_counts_csc(self.ctx, G, N, p.ctypes.data_as(A.c_int32_p),
                                                    i.ctypes.data_as(A.c_int32_p), x.ctypes.data_as(A.c_double_p)))
        self.n_genes, self.n_cells = G, N

    def load_counts_i32(self, x):
        """Gene-major int32 G x N host array (NA = -1)."""
        x = np.ascontiguousarray(x, dtype=np.int32)
        G, N = x.shape
        self._check(self.lib.lpb200_load_counts_i32(self.ctx, G, N, C.c_void_p(x.ctypes.data), 0))
        self.n_genes, self.n_cells = G, N

    def load_counts_device_i32(self, dev_ptr: int, G: int, N: int):
        """Gene-major int32 G x N already resident in device memory (e.g. a torch tensor)."""
        self._check(self.lib.lpb200_load_counts_i32(self.ctx, G, N, C.c_void_p(dev_ptr), 1))
        self.n_genes, self.n_cells = G, N

    def set_design(self, design: A.DesignData):
        self._check(self.lib.lpb200_set_design(self.ctx, C.byref(design.struct)))
        self.design = design

    def count_flags(self):
        """(largest count per gene, largest count per cell, number of negative entries) of the uploaded matrix."""
        gm, cm = np.zeros(self.n_genes, dtype=np.int32), np.zeros(self.n_cells, dtype=np.int32)
        neg = C.c_int64()
        self._check(self.lib.lpb200_count_flags(self.ctx, gm.ctypes.data_as(A.c_int32_p), cm.ctypes.data_as(A.c_int32_p), C.byref(neg)))
        return gm, cm, int(neg.value)

    def row_means(self):
        out = np.zeros(self.n_genes)
        self._check(self.lib.lpb200_row_means(self.ctx, out.ctypes.data_as(A.c_double_p)))
        return out

    # -- compute ---------------------------------------------------------------------------
    def init_params(self, model: A.Model, min_rowmean_global=float("nan"), with_drop=True) -> A.ParamsData:
        p = A.ParamsData(self.design, model, with_drop=with_drop)
        ps = p.as_struct()
        self._check(self.lib.lpb200_init_params(self.ctx, C.byref(model), C.byref(ps), min_rowmean_global))
        p.sync_back()
        return p

    def eval_loglik(self, model: A.Model, params: A.ParamsData):
        ll = np.zeros(self.n_genes)
        ps = params.as_struct()
        self._check(self.lib.lpb200_eval_loglik(self.ctx, C.byref(model), C.byref(ps), ll.ctypes.data_as(A.c_double_p)))
        return ll

    def impulse_starts(self):
        G = self.n_genes
        peak = np.zeros((G, 7), order="F")
        valley = np.zeros((G, 7), order="F")
        self._check(self.lib.lpb200_impulse_starts(self.ctx, peak.ctypes.data_as(A.c_double_p),
                                                   valley.ctypes.data_as(A.c_double_p)))
        return peak, valley

    def fit_mudisp(self, model, params, control=None, allow_nonfinite=False):
        control = control or A.default_control()
        G = self.n_genes
        ll = np.zeros(G)
        conv = np.zeros(G, dtype=np.int32)
        ps = params.as_struct()
        rc = self.lib.lpb200_fit_mudisp(self.ctx, C.byref(model), C.byref(ps), C.byref(control),
                                        ll.ctypes.data_as(A.c_double_p), conv.ctypes.data_as(A.c_int32_p))
        self._check(rc, allow=(A.E_NONFINITE,) if allow_nonfinite else ())
        params.sync_back()
        return dict(ll=ll, conv=conv, rc=rc)

    def fit_mudisp_items(self, model, params, control=None):
        """fit_mudisp plus every (gene, start) work item of the fit kernel (lpb200_fit_mudisp_items)."""
        control = control or A.default_control()
        G = self.n_genes
        ll = np.zeros(G)
        conv = np.zeros(G, dtype=np.int32)
        ns, nt = C.c_int32(), C.c_int32()
        cap = G * 3
        theta0 = np.zeros((cap, A.MAX_PARAMS))
        theta = np.zeros((cap, A.MAX_PARAMS))
        ll_i = np.zeros(cap)
        conv_i = np.zeros(cap, dtype=np.int32)
        evals_i = np.zeros(cap, dtype=np.uint32)
        ps = params.as_struct()
        rc = self.lib.lpb200_fit_mudisp_items(
            self.ctx, C.byref(model), C.byref(ps), C.byref(control), ll.ctypes.data_as(A.c_double_p),
            conv.ctypes.data_as(A.c_int32_p), C.byref(ns), C.byref(nt), theta0.ctypes.data_as(A.c_double_p),
            theta.ctypes.data_as(A.c_double_p), ll_i.ctypes.data_as(A.c_double_p), conv_i.ctypes.data_as(A.c_int32_p),
            evals_i.ctypes.data_as(C.POINTER(C.c_uint32)))
        self._check(rc, allow=(A.E_NONFINITE,))
        params.sync_back()
        n_items, n = G * ns.value, nt.value
        flat0, flat1 = theta0.ravel()[: n_items * n].reshape(n_items, n), theta.ravel()[: n_items * n].reshape(n_items, n)
        return dict(ll=ll, conv=conv, rc=rc, n_starts=ns.value, theta0=flat0.copy(), theta=flat1.copy(),
                    ll_items=ll_i[:n_items].copy(), conv_items=conv_i[:n_items].copy(), evals_items=evals_i[:n_items].copy())

    def fit_pi(self, model, params, control=None):
        control = control or A.default_control()
        n = self.n_cells if model.drop_fit_group == A.DROPFIT_PERCELL else 1
        ll = np.zeros(n)
        conv = np.zeros(n, dtype=np.int32)
        ps = params.as_struct()
        self._check(self.lib.lpb200_fit_pi(self.ctx, C.byref(model), C.byref(ps), C.byref(control),
                                           ll.ctypes.data_as(A.c_double_p), conv.ctypes.data_as(A.c_int32_p)))
        params.sync_back()
        return dict(ll=ll, conv=conv)

    def fit_model(self, model, params, fit_drop, control=None, min_rowmean_global=float("nan"), keep_mask=0):
        """fitModel; ``keep_mask`` (_cabi.KEEP_* bits): matrices of ``params`` that are warm-start values."""
        control = control or A.default_control()
        G = self.n_genes
        em = np.full(control.max_cycles, np.nan)
        n_iter, converged = C.c_int32(), C.c_int32()
        ll_init = C.c_double()
        ll = np.zeros(G)
        conv = np.zeros(G, dtype=np.int32)
        ps = params.as_struct()
        self._check(self.lib.lpb200_fit_model_from(self.ctx, C.byref(model), C.byref(ps), int(bool(fit_drop)),
                                                   C.byref(control), min_rowmean_global, int(keep_mask),
                                                   em.ctypes.data_as(A.c_double_p), C.byref(n_iter), C.byref(converged),
                                              C.byref(ll_init), ll.ctypes.data_as(A.c_double_p),
                                              conv.ctypes.data_as(A.c_int32_p)))
        params.sync_back()
        return dict(em_ll=em, n_iter=n_iter.value, converged=bool(converged.value), ll_init=ll_init.value,
                    ll=ll, conv=conv)

    def fit_model_multi(self, jobs, control=None, min_rowmean_global=float("nan")):
        """Independent fitModel() calls without drop-out estimation, run concurrently
        (lpb200_fit_model_multi).  ``jobs`` = [(Model, ParamsData), ...]; returns one dict per job."""
        control = control or A.default_control()
        n, G = len(jobs), self.n_genes
        models = (A.Model * n)(*[m for m, _ in jobs])
        structs = (A.Params * n)(*[p.as_struct() for _, p in jobs])
        ll_init, em = np.zeros(n), np.zeros(n)
        converged = np.zeros(n, dtype=np.int32)
        ll = np.zeros((n, G))
        conv = np.zeros((n, G), dtype=np.int32)
        ole = (C.c_uint64 * n)()
        self._check(self.lib.lpb200_fit_model_multi(
            self.ctx, n, models, structs, C.byref(control), min_rowmean_global, ll_init.ctypes.data_as(A.c_double_p),
            em.ctypes.data_as(A.c_double_p), converged.ctypes.data_as(A.c_int32_p), ll.ctypes.data_as(A.c_double_p),
            conv.ctypes.data_as(A.c_int32_p), ole))
        out = []
        for j, (_, p) in enumerate(jobs):
            p.sync_back()
            out.append(dict(em_ll=np.array([em[j]]), n_iter=1, converged=bool(converged[j]), ll_init=ll_init[j],
                            ll=ll[j].copy(), conv=conv[j].copy(), ole=int(ole[j])))
        return out

    def post_drop(self, model, params, gene_ids):
        ids = np.ascontiguousarray(gene_ids, dtype=np.int32)
        z = np.zeros((len(ids), self.n_cells), order="F")
        ps = params.as_struct()
        self._check(self.lib.lpb200_post_drop(self.ctx, C.byref(model), C.byref(ps), ids.ctypes.data_as(A.c_int32_p),
                                              len(ids), z.ctypes.data_as(A.c_double_p)))
        return z

    def expand_fits(self, model, params, gene_ids, what):
        """getFits.R expansions (what = _cabi.EXPAND_*): n_ids x N matrix."""
        ids = np.ascontiguousarray(gene_ids, dtype=np.int32)
        z = np.zeros((len(ids), self.n_cells), order="F")
        ps = params.as_struct()
        self._check(self.lib.lpb200_expand_fits(self.ctx, C.byref(model), C.byref(ps), ids.ctypes.data_as(A.c_int32_p),
                                                len(ids), int(what), z.ctypes.data_as(A.c_double_p)))
        return z

    def stats(self):
        s = A.Stats()
        self._check(self.lib.lpb200_get_stats(self.ctx, C.byref(s)))
        return dict(fn_evals=s.fn_evals, ole=s.ole, launches=s.launches, kernel_ms=s.kernel_ms)

    def reset_stats(self):
        self._check(self.lib.lpb200_reset_stats(self.ctx))

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._check(self.lib.lpb200_fp64_peak(self.ctx, C.byref(v)))
        return v.value

    def libm_probe(self, x, use_library: bool):
        """exp(x), log(x) on the device: the inner loop's own implementation or the CUDA library's."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        e, l = np.zeros(len(x)), np.zeros(len(x))
        self._check(self.lib.lpb200_libm_probe(self.ctx, x.ctypes.data_as(A.c_double_p), len(x), int(use_library),
                                               e.ctypes.data_as(A.c_double_p), l.ctypes.data_as(A.c_double_p)))
        return e, l

    def div_probe(self, a, b, use_library: bool):
        """a / b and 1 / b on the device: the inner loop's check-free division or the compiler's."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        q, r = np.zeros(len(a)), np.zeros(len(a))
        self._check(self.lib.lpb200_div_probe(self.ctx, a.ctypes.data_as(A.c_double_p), b.ctypes.data_as(A.c_double_p),
                                              len(a), int(use_library), q.ctypes.data_as(A.c_double_p),
                                              r.ctypes.data_as(A.c_double_p)))
        return q, r

    def rcp64h_probe(self, hi, lo=0):
        """MUFU.RCP64H of the doubles (hi[k], lo): (result hi words, result lo words)."""
        hi = np.ascontiguousarray(hi, dtype=np.int32)
        oh, ol = np.zeros(len(hi), dtype=np.int32), np.zeros(len(hi), dtype=np.int32)
        self._check(self.lib.lpb200_rcp64h_probe(self.ctx, hi.ctypes.data_as(A.c_int32_p), len(hi), int(lo),
                                                 oh.ctypes.data_as(A.c_int32_p), ol.ctypes.data_as(A.c_int32_p)))
        return oh, ol


def lrt(ll_full, ll_red, ddf, library_path=None):
    """LRT p-values and BH-adjusted p-values (runHypothesisTests.R:81-86).  Pure host arithmetic
    inside the library; needs no GPU."""
    lib = A.load_library(library_path)
    f = np.ascontiguousarray(ll_full, dtype=np.float64)
    r = np.ascontiguousarray(ll_red, dtype=np.float64)
    p = np.zeros(len(f))
    padj = np.zeros(len(f))
    rc = lib.lpb200_lrt(f.ctypes.data_as(A.c_double_p), r.ctypes.data_as(A.c_double_p), int(ddf), len(f),
                        p.ctypes.data_as(A.c_double_p), padj.ctypes.data_as(A.c_double_p))
    if rc != 0:
        raise LPB200Error(rc, "lpb200_lrt failed")
    return p, padj
