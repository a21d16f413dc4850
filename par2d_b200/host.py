"""Host-side mirror of the reference's case handling (no CUDA involved).

``Case`` wraps include/par2d_host.h: it reads the reference's case directories
(FV/FileIO.cpp:13-51 block files + FV/main.cpp:117-142 grid.conn), generates the
synthetic ramp meshes of the benchmark, splits blocks with GridPart's rules and
assigns blocks to ranks.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import BlockDesc, Config


class Case:
    def __init__(self, handle: C.c_void_p):
        self._lib = _capi.load_host()
        self._h = handle

    # -- constructors -----------------------------------------------------
    @classmethod
    def load(cls, directory: str) -> "Case":
        lib = _capi.load_host()
        h = C.c_void_p()
        lib.check(lib.par2d_case_load(os.fsencode(directory), C.byref(h)), "case_load")
        return cls(h)

    @classmethod
    def synthetic_ramp(cls, ncx: int, ncy: int, nbx: int, nby: int, nranks: int = 1, rank: int = -1) -> "Case":
        """rank >= 0: node coordinates only for the blocks that rank owns under partition(nranks)."""
        lib = _capi.load_host()
        h = C.c_void_p()
        lib.check(lib.par2d_case_synthetic_ramp_ranked(ncx, ncy, nbx, nby, nranks, rank, C.byref(h)),
                  "case_synthetic_ramp")
        return cls(h)

    @classmethod
    def gridpart(cls, directory: str, isteps=None) -> "Case":
        """GridPart's decomposition of <directory>/gridin.conn + gridin.dat (GridPart/main.c).  isteps: the manual
        table of i node breaks of the committed GridPart source ([0, 99, 178, 257, 356]); None = the automatic
        ceil rule in both directions."""
        lib = _capi.load_host()
        h = C.c_void_p()
        if isteps is None:
            arr, n = None, 0
        else:
            arr, n = (C.c_int32 * len(isteps))(*isteps), len(isteps)
        lib.check(lib.par2d_gridpart(os.fsencode(directory), arr, n, C.byref(h)), "gridpart")
        return cls(h)

    def write_gridpart(self, directory: str):
        """grid.conn + subgrids/ in the byte format of GridPart's own output."""
        self._lib.check(self._lib.par2d_case_write_gridpart(self._h, os.fsencode(directory)), "case_write_gridpart")

    def split(self, nbx: int, nby: int) -> "Case":
        h = C.c_void_p()
        self._lib.check(self._lib.par2d_case_split(self._h, nbx, nby, C.byref(h)), "case_split")
        return Case(h)

    # -- queries ----------------------------------------------------------
    @property
    def nblocks(self) -> int:
        return self._lib.par2d_case_num_blocks(self._h)

    @property
    def ncells(self) -> int:
        return self._lib.par2d_case_num_cells(self._h)

    def desc(self, b: int) -> BlockDesc:
        d = BlockDesc()
        self._lib.check(self._lib.par2d_case_block(self._h, b, C.byref(d)), "case_block")
        return d

    def descs(self):
        arr = (BlockDesc * self.nblocks)()
        for b in range(self.nblocks):
            self._lib.check(self._lib.par2d_case_block(self._h, b, C.byref(arr[b])), "case_block")
        return arr

    def shape(self, b: int):
        d = self.desc(b)
        return d.nx, d.ny

    def nodes(self, b: int):
        d = self.desc(b)
        n = d.nx * d.ny
        x = np.ctypeslib.as_array(d.x, shape=(n,)).reshape(d.ny, d.nx).copy()
        y = np.ctypeslib.as_array(d.y, shape=(n,)).reshape(d.ny, d.nx).copy()
        return x, y

    def ibound(self, b: int):
        d = self.desc(b)
        n = 2 * (d.nx - 1) + 2 * (d.ny - 1)
        return np.ctypeslib.as_array(d.ibound, shape=(n,)).copy()

    def partition(self, nranks: int):
        self._lib.check(self._lib.par2d_case_partition(self._h, nranks), "case_partition")
        return [self.desc(b).owner for b in range(self.nblocks)]

    def write(self, directory: str, one_based: bool = False):
        self._lib.check(self._lib.par2d_case_write(self._h, os.fsencode(directory), int(one_based)), "case_write")

    def close(self):
        if self._h:
            self._lib.par2d_case_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_thermo(path: str, cfg: Config) -> Config:
    lib = _capi.load_host()
    lib.check(lib.par2d_thermo_load(os.fsencode(path), C.byref(cfg)), "thermo_load")
    return cfg


def write_solution(path: str, title: str, cfg: Config, x: np.ndarray, y: np.ndarray, unk: np.ndarray):
    """solution.b<N>.tec in print_state's byte format (FV/FileIO.cpp:84-115)."""
    lib = _capi.load_host()
    ny, nx = x.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    unk = np.ascontiguousarray(unk, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    lib.check(lib.par2d_write_solution(os.fsencode(path), title.encode(), C.byref(cfg), nx, ny,
                                       x.ctypes.data_as(dp), y.ctypes.data_as(dp), unk.ctypes.data_as(dp)),
              "write_solution")


def read_solution(path: str, nx: int, ny: int):
    """Restart reader for solution.b<N>.tec (reference print_state format, FV/FileIO.cpp:84-115):
    returns (unk[ny-1, nx-1, 4], xc[ny-1, nx-1], yc[ny-1, nx-1])."""
    lib = _capi.load_host()
    unk = np.empty((ny - 1, nx - 1, 4), dtype=np.float64)
    xc = np.empty((ny - 1, nx - 1), dtype=np.float64)
    yc = np.empty((ny - 1, nx - 1), dtype=np.float64)
    dp = C.POINTER(C.c_double)
    lib.check(lib.par2d_read_solution(os.fsencode(path), nx, ny, unk.ctypes.data_as(dp), xc.ctypes.data_as(dp),
                                      yc.ctypes.data_as(dp)), "read_solution")
    return unk, xc, yc


def nodal_dgp1(unk: np.ndarray, ux: np.ndarray, uy: np.ndarray) -> np.ndarray:
    """Node values of the DG(P1) representation as print_state_DGP1 averages them (FV/FileIO.cpp:139-201):
    every node takes the mean of the standard reconstruction unk + xsi*ux + eta*uy of its 1, 2 or 4
    adjacent cells at their corner facing the node.  unk/ux/uy: (ncy, ncx, 4); returns (ncy+1, ncx+1, 4)."""
    ncy, ncx, _ = unk.shape
    out = np.zeros((ncy + 1, ncx + 1, 4))
    for j in range(ncy + 1):
        for i in range(ncx + 1):
            # the reference's summation order: pp, pm, mp, mm (interior), see the branches of FileIO.cpp:147-199
            terms = []
            for (di, dj, xsi, eta) in ((0, 0, -1.0, -1.0), (0, -1, -1.0, 1.0), (-1, 0, 1.0, -1.0), (-1, -1, 1.0, 1.0)):
                ci, cj = i + di, j + dj
                if 0 <= ci < ncx and 0 <= cj < ncy:
                    terms.append(unk[cj, ci] + (xsi * ux[cj, ci]) + (eta * uy[cj, ci]))
            if len(terms) == 1:
                out[j, i] = terms[0]
            elif len(terms) == 2:
                out[j, i] = 0.5 * (terms[0] + terms[1])
            else:
                out[j, i] = 0.25 * (terms[0] + terms[1] + terms[2] + terms[3])
    return out
