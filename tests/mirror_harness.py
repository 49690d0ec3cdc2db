"""Builds and binds tests/csrc/mirror_harness.cpp: the host mirror of fit_mudisp_kernel (test infrastructure).

The mirror runs the kernel's own arithmetic source (lp_eval.cuh, lp_math.cuh, lp_libm.cuh's DEVICE routines in
host-emulation mode) in the kernel's own summation order.  The one hardware ingredient, the reciprocal seed
MUFU.RCP64H inside the device logarithm, comes from tests/golden/rcp64h.npz (dumped on a B200 by
scripts/dump_rcp64h.py, packed by scripts/make_rcp64h_golden.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

from lineagepulse_b200 import _cabi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
LIB = os.path.join(BUILD, "libmirrorharness.so")
SRC = os.path.join(ROOT, "tests", "csrc", "mirror_harness.cpp")
HDRS = [os.path.join(ROOT, "lineagepulse_b200", "csrc", h) for h in
        ("lp_common.cuh", "lp_math.cuh", "lp_libm.cuh", "lp_model.cuh", "lp_bfgs.cuh", "lp_eval.cuh")]
TABLE = os.path.join(ROOT, "tests", "golden", "rcp64h.npz")


def rcp64h_table():
    """Upper result words of MUFU.RCP64H for the 2^20 upper-word mantissas at exponent 0x3FF."""
    z = np.load(TABLE)
    # stored as the first word plus int8 steps between neighbours
    return np.ascontiguousarray(int(z["first"]) + np.concatenate([[0], np.cumsum(z["steps"].astype(np.int64))]), dtype=np.int32)


def build():
    os.makedirs(BUILD, exist_ok=True)
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-pthread", SRC, "-o", LIB])
    return LIB


class Mirror:
    def __init__(self):
        self.lib = C.CDLL(build())
        self.table = rcp64h_table()   # kept alive: the harness holds the pointer
        assert self.table.shape == (1 << 20,)
        self.lib.mh_set_rcp64h_table(self.table.ctypes.data_as(A.c_int32_p))
        self.lib.mh_fit_items.argtypes = [C.POINTER(A.Design), C.POINTER(A.Model), A.c_int32_p, A.c_double_p, C.c_int,
                                          A.c_double_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                          A.c_double_p, A.c_double_p, A.c_int32_p, C.POINTER(C.c_uint32)]
        self.lib.mh_exp_log.argtypes = [A.c_double_p, C.c_int, A.c_double_p, A.c_double_p]
        self.lib.mh_ll_obs.argtypes = [C.c_double] * 4
        self.lib.mh_ll_obs.restype = C.c_double

    def exp_log(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        e, l = np.zeros(len(x)), np.zeros(len(x))
        self.lib.mh_exp_log(x.ctypes.data_as(A.c_double_p), len(x), e.ctypes.data_as(A.c_double_p), l.ctypes.data_as(A.c_double_p))
        return e, l

    def fit_items(self, design, model, X, theta0, n_starts, drop=None, control=None, use_fast=True, use_table=True,
                  nthreads=None):
        """Every (gene, start) item of fitMuDisp as the kernel computes it: theta, ll, conv, evals."""
        control = control or A.default_control()
        X = np.ascontiguousarray(X, dtype=np.int32)
        theta0 = np.ascontiguousarray(theta0, dtype=np.float64)
        n_items, n = theta0.shape
        assert n_items == X.shape[0] * n_starts
        theta = np.zeros_like(theta0)
        ll = np.zeros(n_items)
        conv = np.zeros(n_items, dtype=np.int32)
        evals = np.zeros(n_items, dtype=np.uint32)
        dropf = None if drop is None else np.asfortranarray(drop, dtype=np.float64)
        got = self.lib.mh_fit_items(
            C.byref(design.struct), C.byref(model), X.ctypes.data_as(A.c_int32_p), theta0.ctypes.data_as(A.c_double_p),
            int(n_starts), None if dropf is None else dropf.ctypes.data_as(A.c_double_p), control.maxit_mudisp,
            control.reltol_mudisp, control.ndeps, int(use_fast), int(use_table), int(nthreads or os.cpu_count() or 1),
            theta.ctypes.data_as(A.c_double_p), ll.ctypes.data_as(A.c_double_p), conv.ctypes.data_as(A.c_int32_p),
            evals.ctypes.data_as(C.POINTER(C.c_uint32)))
        assert got == n, (got, n)
        return dict(theta=theta, ll_items=ll, conv_items=conv, evals_items=evals)
