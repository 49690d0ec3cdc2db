"""ctypes mirror of ``include/lpb200.h`` (the C ABI of the CUDA library).

Only plain C types cross this boundary.  The structures here are also reused by
``tests/oracle_binding.py`` to drive the CPU oracle with identical inputs; the
product never loads the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MAX_CONF = 4
MAX_PARAMS = 32
MAX_PI_PRED = 6

MU_CONSTANT, MU_IMPULSE, MU_SPLINES, MU_GROUPS, MU_MM = 0, 1, 2, 3, 4
DISP_CONSTANT, DISP_SPLINES, DISP_GROUPS = 0, 2, 3
DROP_NONE, DROP_LOGISTIC, DROP_LOGISTIC_OFMU = 0, 1, 2
DROPFIT_PERCELL, DROPFIT_FORALLCELLS = 0, 1
NA_COUNT = -1
EXPAND_MEAN, EXPAND_DISP, EXPAND_DROP, EXPAND_POSTDROP, EXPAND_NORMDATA = 0, 1, 2, 3, 4
EXPAND_NO_DEPTH, EXPAND_NO_BATCH = 16, 32

KEEP_MU, KEEP_BATCH_MU, KEEP_DISP, KEEP_BATCH_DISP = 1, 2, 4, 8

E_ARG, E_CUDA, E_STATE, E_NOTSUP, E_NONFINITE, E_COMM = -1, -2, -3, -4, -5, -6

MU_MODELS = {"constant": MU_CONSTANT, "impulse": MU_IMPULSE, "splines": MU_SPLINES,
             "groups": MU_GROUPS, "MM": MU_MM}
DISP_MODELS = {"constant": DISP_CONSTANT, "splines": DISP_SPLINES, "groups": DISP_GROUPS}
DROP_MODELS = {"none": DROP_NONE, "logistic": DROP_LOGISTIC, "logistic_ofMu": DROP_LOGISTIC_OFMU}
DROP_FIT_GROUPS = {"PerCell": DROPFIT_PERCELL, "ForAllCells": DROPFIT_FORALLCELLS}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Design(C.Structure):
    _fields_ = [
        ("n_genes", C.c_int32), ("n_cells", C.c_int32),
        ("size_factors", c_double_p), ("pseudotime", c_double_p),
        ("group_idx", c_int32_p), ("n_groups", C.c_int32),
        ("spline_basis_mu", c_double_p), ("k_spline_mu", C.c_int32),
        ("spline_basis_disp", c_double_p), ("k_spline_disp", C.c_int32),
        ("impulse_basis", c_double_p),
        ("n_conf_mu", C.c_int32), ("batch_idx_mu", c_int32_p), ("n_batches_mu", C.c_int32 * MAX_CONF),
        ("n_conf_disp", C.c_int32), ("batch_idx_disp", c_int32_p), ("n_batches_disp", C.c_int32 * MAX_CONF),
        ("pi_const_pred", c_double_p), ("n_pi_pred", C.c_int32),
    ]


class Model(C.Structure):
    _fields_ = [("mu_model", C.c_int32), ("disp_model", C.c_int32),
                ("drop_model", C.c_int32), ("drop_fit_group", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("mat_mu", c_double_p), ("batch_mu", c_double_p), ("mat_disp", c_double_p),
                ("batch_disp", c_double_p), ("mat_drop", c_double_p)]


class Control(C.Structure):
    _fields_ = [("maxit_mudisp", C.c_int32), ("reltol_mudisp", C.c_double),
                ("maxit_pi", C.c_int32), ("reltol_pi", C.c_double), ("ndeps", C.c_double),
                ("max_cycles", C.c_int32), ("prec_em", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("fn_evals", C.c_uint64), ("ole", C.c_uint64), ("launches", C.c_uint64),
                ("kernel_ms", C.c_double)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)
F64, I64 = 0, 1


def default_control() -> Control:
    """Constants of fitModel.R:164-175 (+ optim's ndeps)."""
    return Control(1000, 1e-8, 10000, 1e-8, 1e-3, 20, 1.0 - 1e-4)


def _dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _iptr(a):
    return a.ctypes.data_as(c_int32_p) if a is not None else None


def _f64(a, order="F"):
    return None if a is None else np.require(np.asarray(a, dtype=np.float64), requirements=["A", "O"] + [order])


class DesignData:
    """Owns the numpy buffers an ``lpb200_design`` points into.

    Mirrors the ``ls*ModelGlobal`` metadata (initialiseMuModel.R:67-91,134-203).
    Group / batch labels are mapped to first-appearance indices,
    ``match(v, unique(v))`` (initialiseMuModel.R:141-142,191-194).
    """

    def __init__(self, n_genes, n_cells, size_factors=None, pseudotime=None, groups=None,
                 spline_basis_mu=None, spline_basis_disp=None, impulse_basis=None,
                 confounders_mu=None, confounders_disp=None, pi_const_pred=None):
        self.n_genes, self.n_cells = int(n_genes), int(n_cells)
        N = self.n_cells
        self.size_factors = _f64(size_factors)
        self.pseudotime = _f64(pseudotime)
        self.group_labels = None
        self.group_idx = None
        self.n_groups = 0
        if groups is not None:
            self.group_labels, idx = first_appearance_index(groups)
            self.group_idx = np.ascontiguousarray(idx, dtype=np.int32)
            self.n_groups = len(self.group_labels)
        self.spline_basis_mu = _f64(spline_basis_mu)
        self.spline_basis_disp = _f64(spline_basis_disp)
        self.impulse_basis = _f64(impulse_basis)
        self.batch_labels_mu, self.batch_idx_mu, self.n_batches_mu = self._conf(confounders_mu)
        self.batch_labels_disp, self.batch_idx_disp, self.n_batches_disp = self._conf(confounders_disp)
        self.pi_const_pred = _f64(pi_const_pred)
        if self.pi_const_pred is not None and self.pi_const_pred.ndim == 1:
            self.pi_const_pred = _f64(self.pi_const_pred.reshape(-1, 1))
        for name, a, shape0 in (("size_factors", self.size_factors, N), ("pseudotime", self.pseudotime, N),
                                ("spline_basis_mu", self.spline_basis_mu, N),
                                ("spline_basis_disp", self.spline_basis_disp, N),
                                ("impulse_basis", self.impulse_basis, N),
                                ("pi_const_pred", self.pi_const_pred, self.n_genes)):
            if a is not None and a.shape[0] != shape0:
                raise ValueError(f"{name}: leading dimension {a.shape[0]} != {shape0}")
        if self.impulse_basis is not None and self.impulse_basis.shape[1] != 4:
            raise ValueError("impulse_basis must be N x 4 (ns(t, df=4, intercept=TRUE))")
        self.struct = self._build()

    def _conf(self, confs):
        if not confs:
            return [], None, []
        if len(confs) > MAX_CONF:
            raise ValueError(f"at most {MAX_CONF} confounders are supported")
        labels, idxs, nb = [], [], []
        for v in confs:
            lab, idx = first_appearance_index(v)
            labels.append(lab)
            idxs.append(idx.astype(np.int32))
            nb.append(len(lab))
        return labels, np.ascontiguousarray(np.stack(idxs, 0), dtype=np.int32), nb

    def _build(self) -> Design:
        d = Design()
        d.n_genes, d.n_cells = self.n_genes, self.n_cells
        d.size_factors = _dptr(self.size_factors)
        d.pseudotime = _dptr(self.pseudotime)
        d.group_idx = _iptr(self.group_idx)
        d.n_groups = self.n_groups
        d.spline_basis_mu = _dptr(self.spline_basis_mu)
        d.k_spline_mu = 0 if self.spline_basis_mu is None else self.spline_basis_mu.shape[1]
        d.spline_basis_disp = _dptr(self.spline_basis_disp)
        d.k_spline_disp = 0 if self.spline_basis_disp is None else self.spline_basis_disp.shape[1]
        d.impulse_basis = _dptr(self.impulse_basis)
        d.n_conf_mu = len(self.n_batches_mu)
        d.batch_idx_mu = _iptr(self.batch_idx_mu)
        for k, v in enumerate(self.n_batches_mu):
            d.n_batches_mu[k] = v
        d.n_conf_disp = len(self.n_batches_disp)
        d.batch_idx_disp = _iptr(self.batch_idx_disp)
        for k, v in enumerate(self.n_batches_disp):
            d.n_batches_disp[k] = v
        d.pi_const_pred = _dptr(self.pi_const_pred)
        d.n_pi_pred = 0 if self.pi_const_pred is None else self.pi_const_pred.shape[1]
        return d

    # sizes (lpb200_n_*_params)
    def n_mu_params(self, model: Model) -> int:
        return {MU_CONSTANT: 1, MU_IMPULSE: 7, MU_SPLINES: self.struct.k_spline_mu,
                MU_GROUPS: self.n_groups}.get(model.mu_model, -1)

    def n_disp_params(self, model: Model) -> int:
        return {DISP_CONSTANT: 1, DISP_SPLINES: self.struct.k_spline_disp,
                DISP_GROUPS: self.n_groups}.get(model.disp_model, -1)

    def n_drop_params(self, model: Model) -> int:
        if model.drop_model == DROP_NONE:
            return 0
        return 1 + (1 if model.drop_model == DROP_LOGISTIC_OFMU else 0) + self.struct.n_pi_pred

    def deg_freedom(self, model: Model) -> int:
        """scaDegFreedom of mu + dispersion model (initialiseMuModel.R:199-201,
        initialiseDispModel.R:173-176)."""
        return (self.n_mu_params(model) + sum(self.n_batches_mu) - len(self.n_batches_mu)
                + self.n_disp_params(model) + sum(self.n_batches_disp) - len(self.n_batches_disp))


def first_appearance_index(v):
    """``unique(v)`` and ``match(v, unique(v)) - 1``."""
    v = np.asarray(v)
    labels, first, inv = np.unique(v, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return labels[order], rank[inv.reshape(-1)]


class ParamsData:
    """numpy storage behind an ``lpb200_params`` (column-major, as R)."""

    def __init__(self, design: DesignData, model: Model, with_drop=True):
        G, N = design.n_genes, design.n_cells
        self.mat_mu = np.zeros((G, design.n_mu_params(model)), dtype=np.float64, order="F")
        self.mat_disp = np.zeros((G, design.n_disp_params(model)), dtype=np.float64, order="F")
        self.batch_mu = [np.ones((G, nb), dtype=np.float64, order="F") for nb in design.n_batches_mu]
        self.batch_disp = [np.ones((G, nb), dtype=np.float64, order="F") for nb in design.n_batches_disp]
        q = design.n_drop_params(model)
        self.mat_drop = np.zeros((N, q), dtype=np.float64, order="F") if (q > 0 and with_drop) else None
        self._flat_bmu = None
        self._flat_bdisp = None

    def as_struct(self) -> Params:
        """Pack into the C struct (batch matrices concatenated); call ``sync_back`` after a
        mutating call."""
        p = Params()
        self.mat_mu = np.asfortranarray(self.mat_mu, dtype=np.float64)
        self.mat_disp = np.asfortranarray(self.mat_disp, dtype=np.float64)
        p.mat_mu = _dptr(self.mat_mu)
        p.mat_disp = _dptr(self.mat_disp)
        self._flat_bmu = (np.concatenate([b.ravel(order="F") for b in self.batch_mu])
                          if self.batch_mu else None)
        self._flat_bdisp = (np.concatenate([b.ravel(order="F") for b in self.batch_disp])
                            if self.batch_disp else None)
        p.batch_mu = _dptr(self._flat_bmu)
        p.batch_disp = _dptr(self._flat_bdisp)
        if self.mat_drop is not None:
            self.mat_drop = np.asfortranarray(self.mat_drop, dtype=np.float64)
        p.mat_drop = _dptr(self.mat_drop)
        return p

    def sync_back(self):
        for flat, mats in ((self._flat_bmu, self.batch_mu), (self._flat_bdisp, self.batch_disp)):
            if flat is None:
                continue
            off = 0
            for k, b in enumerate(mats):
                n = b.size
                mats[k] = flat[off:off + n].reshape(b.shape, order="F").copy(order="F")
                off += n

    def copy(self):
        import copy as _copy
        o = _copy.copy(self)
        o.mat_mu = self.mat_mu.copy(order="F")
        o.mat_disp = self.mat_disp.copy(order="F")
        o.batch_mu = [b.copy(order="F") for b in self.batch_mu]
        o.batch_disp = [b.copy(order="F") for b in self.batch_disp]
        o.mat_drop = None if self.mat_drop is None else self.mat_drop.copy(order="F")
        return o


_LIB_NAME = "liblpb200.so"


def default_library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def declare_product_prototypes(lib):
    """argtypes / restype for every symbol declared in include/lpb200.h."""
    P = C.POINTER
    vp = C.c_void_p
    sig = {
        "lpb200_abi_version": (C.c_int, []),
        "lpb200_default_control": (None, [P(Control)]),
        "lpb200_n_mu_params": (C.c_int, [P(Model), P(Design)]),
        "lpb200_n_disp_params": (C.c_int, [P(Model), P(Design)]),
        "lpb200_n_drop_params": (C.c_int, [P(Model), P(Design)]),
        "lpb200_n_batch_cols": (C.c_int, [c_int32_p, C.c_int]),
        "lpb200_deg_freedom": (C.c_int, [P(Model), P(Design)]),
        "lpb200_create": (vp, [C.c_int]),
        "lpb200_create_multi": (vp, [c_int32_p, C.c_int]),
        "lpb200_n_shards": (C.c_int, [vp]),
        "lpb200_shard_bounds": (C.c_int, [vp, c_int32_p]),
        "lpb200_comm_unique_id": (C.c_int, [C.c_char_p]),
        "lpb200_comm_init": (C.c_int, [vp, C.c_int, C.c_int, C.c_char_p]),
        "lpb200_comm_stats": (C.c_int, [vp, P(C.c_uint64), P(C.c_uint64)]),
        "lpb200_destroy": (None, [vp]),
        "lpb200_release_cached_memory": (C.c_int, [C.c_int]),
        "lpb200_last_error": (C.c_char_p, [vp]),
        "lpb200_set_stream": (C.c_int, [vp, vp]),
        "lpb200_set_collective": (C.c_int, [vp, C.c_int, C.c_int, ALLREDUCE_FN, vp]),
        "lpb200_load_counts_dense": (C.c_int, [vp, C.c_int, C.c_int, c_double_p]),
        "lpb200_load_counts_csc": (C.c_int, [vp, C.c_int, C.c_int, c_int32_p, c_int32_p, c_double_p]),
        "lpb200_load_counts_i32": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int]),
        "lpb200_set_design": (C.c_int, [vp, P(Design)]),
        "lpb200_count_flags": (C.c_int, [vp, c_int32_p, c_int32_p, P(C.c_int64)]),
        "lpb200_init_params": (C.c_int, [vp, P(Model), P(Params), C.c_double]),
        "lpb200_eval_loglik": (C.c_int, [vp, P(Model), P(Params), c_double_p]),
        "lpb200_impulse_starts": (C.c_int, [vp, c_double_p, c_double_p]),
        "lpb200_fit_mudisp": (C.c_int, [vp, P(Model), P(Params), P(Control), c_double_p, c_int32_p]),
        "lpb200_fit_mudisp_items": (C.c_int, [vp, P(Model), P(Params), P(Control), c_double_p, c_int32_p, c_int32_p, c_int32_p,
                                              c_double_p, c_double_p, c_double_p, c_int32_p, P(C.c_uint32)]),
        "lpb200_fit_pi": (C.c_int, [vp, P(Model), P(Params), P(Control), c_double_p, c_int32_p]),
        "lpb200_fit_model": (C.c_int, [vp, P(Model), P(Params), C.c_int, P(Control), C.c_double,
                                       c_double_p, c_int32_p, c_int32_p, c_double_p, c_double_p, c_int32_p]),
        "lpb200_fit_model_from": (C.c_int, [vp, P(Model), P(Params), C.c_int, P(Control), C.c_double, C.c_int,
                                            c_double_p, c_int32_p, c_int32_p, c_double_p, c_double_p, c_int32_p]),
        "lpb200_fit_model_multi": (C.c_int, [vp, C.c_int, P(Model), P(Params), P(Control), C.c_double, c_double_p,
                                             c_double_p, c_int32_p, c_double_p, c_int32_p, P(C.c_uint64)]),
        "lpb200_post_drop": (C.c_int, [vp, P(Model), P(Params), c_int32_p, C.c_int, c_double_p]),
        "lpb200_expand_fits": (C.c_int, [vp, P(Model), P(Params), c_int32_p, C.c_int, C.c_int, c_double_p]),
        "lpb200_lrt": (C.c_int, [c_double_p, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]),
        "lpb200_row_means": (C.c_int, [vp, c_double_p]),
        "lpb200_get_stats": (C.c_int, [vp, P(Stats)]),
        "lpb200_reset_stats": (C.c_int, [vp]),
        "lpb200_fp64_peak": (C.c_int, [vp, c_double_p]),
        "lpb200_libm_probe": (C.c_int, [vp, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]),
        "lpb200_div_probe": (C.c_int, [vp, c_double_p, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]),
        "lpb200_rcp64h_probe": (C.c_int, [vp, c_int32_p, C.c_int, C.c_int32, c_int32_p, c_int32_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError => the library does not export the ABI
        fn.restype = res
        fn.argtypes = args
    return sorted(sig)


class LibraryMissing(RuntimeError):
    pass


UNIQUE_ID_BYTES = 128


def _point_at_bundled_nccl():
    """The library binds NCCL with dlopen (lp_comm.cuh).  Inside a Python environment that ships NCCL as a wheel
    (nvidia-nccl-cu12, which torch uses) point it at that copy, so that this process never holds two different
    NCCL builds under one SONAME."""
    if os.environ.get("LPB200_NCCL_LIBRARY"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["LPB200_NCCL_LIBRARY"] = cand
                return
    except Exception:
        pass


def load_library(path: str | None = None):
    """dlopen the CUDA library.  There is no CPU fallback: a missing library is an error."""
    _point_at_bundled_nccl()
    path = path or os.environ.get("LPB200_LIBRARY") or default_library_path()
    if not os.path.exists(path):
        raise LibraryMissing(
            f"{path} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C lineagepulse_b200/csrc)")
    lib = C.CDLL(path)
    declare_product_prototypes(lib)
    return lib
