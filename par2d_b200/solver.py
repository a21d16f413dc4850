"""Python mirror of the reference driver loop (FV/main.cpp:263-532) over the C ABI.

``Solver`` owns one ``par2d_ctx``.  By default it binds the CUDA library
(libpar2d_gpu.so) and raises if that is missing - there is no CPU fallback.
Tests pass ``lib=_capi.load_oracle()`` to run the same calls against the CPU
restatement in oracle/ as the checker.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import _capi
from ._capi import Config, StepInfo
from .host import Case

_DP = C.POINTER(C.c_double)


def default_config(lib=None) -> Config:
    lib = lib or _capi.load_gpu()
    cfg = Config()
    lib.default_config(C.byref(cfg))
    return cfg


class Solver:
    def __init__(self, case: Case, cfg: Optional[Config] = None, lib=None, owners=None):
        self.lib = lib or _capi.load_gpu()
        self.case = case
        self.cfg = (cfg or default_config(self.lib)).copy()
        self._descs = case.descs()
        if owners is not None:
            for b, r in enumerate(owners):
                self._descs[b].owner = int(r)
        self.shapes = [(d.nx, d.ny) for d in self._descs]
        self.nblocks = len(self._descs)
        h = C.c_void_p()
        self.lib.check(self.lib.create(C.byref(self.cfg), self.nblocks, self._descs, C.byref(h)), "create")
        self._h = h
        self.owned = [self.lib.owned_block(self._h, k) for k in range(self.lib.num_owned(self._h))]

    # -- state ------------------------------------------------------------
    def ncells_block(self, b: int) -> int:
        nx, ny = self.shapes[b]
        return (nx - 1) * (ny - 1)

    def get_state(self, b: int) -> np.ndarray:
        """(ny-1, nx-1, 4) array = unk[(j*(nx-1)+i)*4+k]."""
        nx, ny = self.shapes[b]
        out = np.empty((ny - 1, nx - 1, 4), dtype=np.float64)
        self.lib.check(self.lib.get_state(self._h, b, out.ctypes.data_as(_DP)), "get_state")
        return out

    def set_state(self, b: int, unk: np.ndarray):
        nx, ny = self.shapes[b]
        unk = np.ascontiguousarray(unk, dtype=np.float64).reshape(ny - 1, nx - 1, 4)
        self.lib.check(self.lib.set_state(self._h, b, unk.ctypes.data_as(_DP)), "set_state")

    def get_states(self, blocks=None):
        """All (or the given) owned blocks in one batched transfer (par2d_get_states): {block: (ny-1, nx-1, 4)}."""
        blocks = list(self.owned if blocks is None else blocks)
        out = {b: np.empty((self.shapes[b][1] - 1, self.shapes[b][0] - 1, 4), dtype=np.float64) for b in blocks}
        ids = (C.c_int32 * len(blocks))(*blocks)
        ptrs = (_DP * len(blocks))(*[out[b].ctypes.data_as(_DP) for b in blocks])
        self.lib.check(self.lib.get_states(self._h, len(blocks), ids, ptrs), "get_states")
        return out

    def set_states(self, states: dict):
        blocks = list(states)
        arrs = [np.ascontiguousarray(states[b], dtype=np.float64) for b in blocks]
        ids = (C.c_int32 * len(blocks))(*blocks)
        ptrs = (_DP * len(blocks))(*[a.ctypes.data_as(_DP) for a in arrs])
        self.lib.check(self.lib.set_states(self._h, len(blocks), ids, ptrs), "set_states")

    def get_residual(self, b: int) -> np.ndarray:
        nx, ny = self.shapes[b]
        out = np.empty((ny - 1, nx - 1, 4), dtype=np.float64)
        self.lib.check(self.lib.get_residual(self._h, b, out.ctypes.data_as(_DP)), "get_residual")
        return out

    # -- second-order branch (cfg.accur == 1): xi / eta slopes of the DG(P1) representation -----------
    def get_slopes(self, b: int):
        nx, ny = self.shapes[b]
        ux = np.empty((ny - 1, nx - 1, 4), dtype=np.float64)
        uy = np.empty((ny - 1, nx - 1, 4), dtype=np.float64)
        self.lib.check(self.lib.get_slopes(self._h, b, ux.ctypes.data_as(_DP), uy.ctypes.data_as(_DP)), "get_slopes")
        return ux, uy

    def set_slopes(self, b: int, ux: np.ndarray, uy: np.ndarray):
        nx, ny = self.shapes[b]
        ux = np.ascontiguousarray(ux, dtype=np.float64).reshape(ny - 1, nx - 1, 4)
        uy = np.ascontiguousarray(uy, dtype=np.float64).reshape(ny - 1, nx - 1, 4)
        self.lib.check(self.lib.set_slopes(self._h, b, ux.ctypes.data_as(_DP), uy.ctypes.data_as(_DP)), "set_slopes")

    def get_residual_dg(self, b: int):
        """dudt, duxdt, duydt of one calc_dudt evaluation with ACCUR == 1."""
        nx, ny = self.shapes[b]
        out = [np.empty((ny - 1, nx - 1, 4), dtype=np.float64) for _ in range(3)]
        self.lib.check(self.lib.get_residual_dg(self._h, b, *[o.ctypes.data_as(_DP) for o in out]), "get_residual_dg")
        return out

    def nodal_dgp1(self, b: int) -> np.ndarray:
        """The node-averaged reconstruction print_state_DGP1 writes (FileIO.cpp:139-201): (ny, nx, 4)."""
        from .host import nodal_dgp1
        ux, uy = self.get_slopes(b)
        return nodal_dgp1(self.get_state(b), ux, uy)

    # -- stepping ---------------------------------------------------------
    def step(self, nsteps: int, stop_on_tol: bool = False, check: bool = True) -> List[StepInfo]:
        info = (StepInfo * nsteps)()
        rc = self.lib.step(self._h, nsteps, int(stop_on_tol), info)
        if check:
            self.lib.check(rc, "step")
        return [i for i in info if i.iter >= 0]

    def block_history(self, nsteps: int) -> np.ndarray:
        """(nsteps, nowned, 2): per owned block sum_k||res_k|| and sum_k res0_k."""
        out = np.empty((nsteps, len(self.owned), 2), dtype=np.float64)
        self.lib.check(self.lib.get_block_history(self._h, nsteps, out.ctypes.data_as(_DP)), "get_block_history")
        return out

    @property
    def iteration(self) -> int:
        return self.lib.iteration(self._h)

    @property
    def owned_cells(self) -> int:
        return self.lib.owned_cells(self._h)

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
