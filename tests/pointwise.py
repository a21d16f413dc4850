"""Shared wrappers around the point-wise test hooks of the C ABI (both prefixes)."""
import ctypes as C

import numpy as np

_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)


def _p(a):
    return a.ctypes.data_as(_DP)


def ldfss(lib, cfg, face, uL, uR):
    n = len(face)
    face, uL, uR = (np.ascontiguousarray(x, dtype=np.float64) for x in (face, uL, uR))
    flux, parr = np.zeros((n, 4)), np.zeros(n)
    lib.check(lib.test_ldfss(C.byref(cfg), n, _p(face), _p(uL), _p(uR), _p(flux), _p(parr)), "test_ldfss")
    return flux, parr


def boundary_state(lib, cfg, btype, norm, uInt):
    n = len(btype)
    btype = np.ascontiguousarray(btype, dtype=np.int32)
    norm, uInt = (np.ascontiguousarray(x, dtype=np.float64) for x in (norm, uInt))
    ghost = np.zeros((n, 4))
    lib.check(lib.test_boundary_state(C.byref(cfg), n, btype.ctypes.data_as(_IP), _p(norm), _p(uInt), _p(ghost)),
              "test_boundary_state")
    return ghost


def update_state(lib, cfg, unk):
    unk = np.array(unk, dtype=np.float64, order="C")
    der = np.zeros((len(unk), 8))
    lib.check(lib.test_update_state(C.byref(cfg), len(unk), _p(unk), _p(der)), "test_update_state")
    return unk, der


def solve(lib, cfg, dt_ycc, unk, res):
    dt_ycc, unk, res = (np.ascontiguousarray(x, dtype=np.float64) for x in (dt_ycc, unk, res))
    dv = np.zeros((len(unk), 4))
    lib.check(lib.test_solve(C.byref(cfg), len(unk), _p(dt_ycc), _p(unk), _p(res), _p(dv)), "test_solve")
    return dv
