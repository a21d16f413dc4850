"""ctypes bindings of the C ABI declared in include/par2d_gpu.h and include/par2d_host.h.

The same struct layouts serve three shared libraries:
  * libpar2d_gpu.so   (product: CUDA path, symbols ``par2d_*``)
  * libpar2d_host.so  (product: case I/O, no CUDA, symbols ``par2d_case_*`` ...)
  * oracle/libpar2d_oracle.so (TEST ONLY: CPU restatement, symbols ``orc_*``)
Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may
pass ``prefix="orc_"``; the product path fails loudly when the CUDA library is
missing instead of falling back.
"""
from __future__ import annotations

import ctypes as C
import os

NVAR = 4
BOTTOM, RIGHT, TOP, LEFT = 0, 1, 2, 3
MODE_FAST, MODE_FAITHFUL = 0, 1

ERR_NAMES = {
    0: "OK", 1: "ERR_ARG", 2: "ERR_CUDA", 3: "ERR_NAN", 4: "ERR_DENSITY", 5: "ERR_STATE",
    6: "ERR_SINGULAR", 7: "ERR_THERMO_RANGE", 8: "ERR_COMM",
}

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
GPU_LIB = os.path.join(PKG_DIR, "libpar2d_gpu.so")
HOST_LIB = os.path.join(PKG_DIR, "libpar2d_host.so")
ORACLE_LIB = os.path.join(REPO_DIR, "oracle", "libpar2d_oracle.so")


class Config(C.Structure):
    """par2d_config (include/par2d_gpu.h)."""
    _fields_ = [
        ("iaxi", C.c_int32), ("mode", C.c_int32),
        ("igam", C.c_double), ("cfl", C.c_double), ("tol", C.c_double),
        ("uFS", C.c_double * NVAR), ("p_back", C.c_double), ("Rs", C.c_double),
        ("thermo_hi", C.c_double * 7), ("thermo_lo", C.c_double * 7),
        ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("accur", C.c_int32),
    ]

    def copy(self) -> "Config":
        out = Config()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(Config))
        return out


class BlockDesc(C.Structure):
    """par2d_block_desc (include/par2d_gpu.h)."""
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32),
        ("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double)), ("ibound", C.POINTER(C.c_int32)),
        ("bid", C.c_int32 * 4), ("bbound", C.c_int32 * 4),
        ("owner", C.c_int32), ("reserved", C.c_int32),
    ]


class StepInfo(C.Structure):
    """par2d_step_info (include/par2d_gpu.h)."""
    _fields_ = [
        ("iter", C.c_int64), ("dt", C.c_double), ("ressumsum", C.c_double), ("res0sum", C.c_double),
        ("relres", C.c_double), ("status", C.c_int32), ("reserved", C.c_int32),
    ]


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)
_CFGP = C.POINTER(Config)

# name -> (restype, argtypes); names are given without prefix.
SOLVER_API = {
    "default_config": (None, [_CFGP]),
    "last_error": (C.c_char_p, []),
    "create": (C.c_int, [_CFGP, C.c_int32, C.POINTER(BlockDesc), C.POINTER(C.c_void_p)]),
    "destroy": (C.c_int, [C.c_void_p]),
    "set_state": (C.c_int, [C.c_void_p, C.c_int32, _DP]),
    "get_state": (C.c_int, [C.c_void_p, C.c_int32, _DP]),
    "get_residual": (C.c_int, [C.c_void_p, C.c_int32, _DP]),
    "set_slopes": (C.c_int, [C.c_void_p, C.c_int32, _DP, _DP]),
    "get_slopes": (C.c_int, [C.c_void_p, C.c_int32, _DP, _DP]),
    "get_residual_dg": (C.c_int, [C.c_void_p, C.c_int32, _DP, _DP, _DP]),
    "step": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(StepInfo)]),
    "get_block_history": (C.c_int, [C.c_void_p, C.c_int32, _DP]),
    "iteration": (C.c_int64, [C.c_void_p]),
    "num_owned": (C.c_int32, [C.c_void_p]),
    "owned_block": (C.c_int32, [C.c_void_p, C.c_int32]),
    "owned_cells": (C.c_int64, [C.c_void_p]),
    "exchange_layout": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "dt_table": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "step_begin": (C.c_int, [C.c_void_p]),
    "step_end": (C.c_int, [C.c_void_p]),
    "step_collect": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(StepInfo)]),
    "test_ldfss": (C.c_int, [_CFGP, C.c_int64, _DP, _DP, _DP, _DP, _DP]),
    "test_boundary_state": (C.c_int, [_CFGP, C.c_int64, _IP, _DP, _DP, _DP]),
    "test_update_state": (C.c_int, [_CFGP, C.c_int64, _DP, _DP]),
    "test_solve": (C.c_int, [_CFGP, C.c_int64, _DP, _DP, _DP, _DP]),
}

# entry points only the CUDA library has
GPU_ONLY_API = {
    "peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "peer_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "peer_disconnect": (C.c_int, [C.c_void_p]),
    "peer_check": (C.c_int, [C.c_void_p]),
    "peer_snapshot": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), _DP, C.POINTER(C.c_int64)]),
    "peer_connect_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32]),
    "step_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "step_wait": (C.c_int, [C.c_void_p, C.POINTER(StepInfo)]),
    "set_states": (C.c_int, [C.c_void_p, C.c_int32, _IP, C.POINTER(_DP)]),
    "get_states": (C.c_int, [C.c_void_p, C.c_int32, _IP, C.POINTER(_DP)]),
    "set_stream": (C.c_int, [C.c_void_p, C.c_uint64]),
    "synchronize": (C.c_int, [C.c_void_p]),
    "probe_fp64": (C.c_int, [C.c_int32, _DP]),
    "device_count": (C.c_int, [_IP]),
    "persist_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "kernel_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "kernel_time": (C.c_int, [C.c_void_p, _DP, C.POINTER(C.c_int64)]),
    "launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
}

# entry points only the oracle has (test hooks)
ORACLE_ONLY_API = {
    "get_metrics": (C.c_int, [C.c_void_p, C.c_int32, _DP, _DP, _DP]),
    "reset_history": (None, [C.c_void_p]),
}

HOST_API = {
    "par2d_case_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "par2d_case_synthetic_ramp": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "par2d_case_synthetic_ramp_ranked": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                   C.POINTER(C.c_void_p)]),
    "par2d_case_split": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "par2d_case_write": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "par2d_gridpart": (C.c_int, [C.c_char_p, _IP, C.c_int32, C.POINTER(C.c_void_p)]),
    "par2d_case_write_gridpart": (C.c_int, [C.c_void_p, C.c_char_p]),
    "par2d_case_num_blocks": (C.c_int32, [C.c_void_p]),
    "par2d_case_num_cells": (C.c_int64, [C.c_void_p]),
    "par2d_case_block": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(BlockDesc)]),
    "par2d_case_partition": (C.c_int, [C.c_void_p, C.c_int32]),
    "par2d_case_free": (None, [C.c_void_p]),
    "par2d_thermo_load": (C.c_int, [C.c_char_p, _CFGP]),
    "par2d_write_solution": (C.c_int, [C.c_char_p, C.c_char_p, _CFGP, C.c_int32, C.c_int32, _DP, _DP, _DP]),
    "par2d_write_solution_dgp1": (C.c_int, [C.c_char_p, C.c_char_p, _CFGP, C.c_int32, C.c_int32, _DP, _DP, _DP, _DP, _DP, C.c_int32, C.c_double]),
    "par2d_read_solution": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, _DP, _DP, _DP]),
    "par2d_host_last_error": (C.c_char_p, []),
}


class Par2dError(RuntimeError):
    def __init__(self, code: int, where: str, msg: str = ""):
        self.code = code
        super().__init__(f"{where}: {ERR_NAMES.get(code, code)} {msg}".strip())


class _Lib:
    """A loaded shared library with prefixed, typed entry points as attributes."""

    def __init__(self, path: str, prefix: str, tables):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing - build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path, self.prefix = path, prefix
        self.cdll = C.CDLL(path, mode=C.RTLD_GLOBAL if prefix == "" else 0)
        for table in tables:
            for name, (res, args) in table.items():
                fn = getattr(self.cdll, prefix + name)
                fn.restype, fn.argtypes = res, args
                setattr(self, name, fn)

    def check(self, rc: int, where: str):
        if rc != 0:
            msg = self.last_error()
            raise Par2dError(rc, where, msg.decode() if msg else "")


_cache = {}


def load_gpu() -> _Lib:
    """The product CUDA library.  No fallback: raises if it has not been built."""
    if "gpu" not in _cache:
        # Kernels of contexts that share a process wait for each other (par2d_peer_connect_local): no kernel may be
        # loaded lazily while another one spins.  par2d_create preloads what it launches; eager module loading (only
        # effective if CUDA is not initialised yet) covers the rest.
        os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
        # PAR2D_GPU_LIB: A/B builds of the same CUDA library (tools/build_variants.sh), measurements only
        _cache["gpu"] = _Lib(os.environ.get("PAR2D_GPU_LIB", GPU_LIB), "par2d_", [SOLVER_API, GPU_ONLY_API])
    return _cache["gpu"]


def load_oracle() -> _Lib:
    """TEST INFRASTRUCTURE ONLY - the CPU restatement in oracle/."""
    if "oracle" not in _cache:
        _cache["oracle"] = _Lib(ORACLE_LIB, "orc_", [SOLVER_API, ORACLE_ONLY_API])
    return _cache["oracle"]


def load_host() -> _Lib:
    if "host" not in _cache:
        lib = _Lib(HOST_LIB, "", [HOST_API])
        lib.last_error = lib.par2d_host_last_error
        _cache["host"] = lib
    return _cache["host"]
