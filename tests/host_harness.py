"""Builds and binds tests/csrc/host_harness.cpp: the LP_HD headers of the CUDA library compiled
with g++ so their scalar logic can be checked on a CPU-only box (test infrastructure)."""
import ctypes as C
import os
import subprocess

from lineagepulse_b200 import _cabi as A
from tests.oracle_binding import ORACLE_DIR, ROOT, build_oracle

BUILD = os.path.join(ROOT, "tests", "_build")
LIB = os.path.join(BUILD, "libhostharness.so")
SRC = os.path.join(ROOT, "tests", "csrc", "host_harness.cpp")
HDRS = [os.path.join(ROOT, "lineagepulse_b200", "csrc", h) for h in
        ("lp_common.cuh", "lp_math.cuh", "lp_libm.cuh", "lp_model.cuh", "lp_bfgs.cuh", "lp_exact.cuh")]


def build():
    build_oracle()
    os.makedirs(BUILD, exist_ok=True)
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-x", "c++", SRC,
                               "-o", LIB, "-L", ORACLE_DIR, "-l:liblporacle.so", f"-Wl,-rpath,{ORACLE_DIR}"])
    return LIB


def load():
    hh = C.CDLL(build())
    hh.hh_ll_obs.restype = C.c_double
    hh.hh_ll_obs.argtypes = [C.c_double] * 4
    hh.hh_fit_gene.argtypes = [C.POINTER(A.Design), C.POINTER(A.Model), A.c_int32_p, C.c_int, A.c_double_p,
                               A.c_double_p, A.c_double_p, C.c_int, C.c_double, C.c_double, A.c_double_p,
                               C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint)]
    hh.hh_rosenbrock.argtypes = [C.c_int, C.c_double, A.c_double_p, C.POINTER(C.c_double), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return hh
