"""One-process-per-GPU driver: GridPart-decomposed blocks sharded over the ranks of one box.

Replaces the reference's MPI plumbing (one rank per block, FV/main.cpp:68-74,125-128):
  * halo exchange  : the four MPI_Sendrecv sites (SpatialDiscretization.cpp:259,320,380,441)
                     become one grouped NCCL send/recv per neighbouring rank on the device
                     buffers the C ABI exposes (par2d_exchange_layout);
  * dt sync        : the gather/fmax/bcast chain (main.cpp:270-292) becomes one all-reduce(MAX)
                     over the per-block dt table (par2d_dt_table);
  * residual gather: main.cpp:379-392 becomes a gather of per-block sums after the batch, added
                     in global block order so relres is independent of the number of GPUs.
torch.distributed is plumbing only; every byte of compute is in libpar2d_gpu.so.  The same class
drives the oracle with the gloo backend in the CPU tests (world_size 2).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._capi import Config, StepInfo
from .host import Case
from .solver import Solver, default_config


class _DevPtr:
    """Minimal __cuda_array_interface__ holder so torch can alias library-owned device memory."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {
            "shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None,
        }


def _alias(ptr: int, n: int, cuda: bool) -> Optional[torch.Tensor]:
    if n == 0 or not ptr:
        return None
    if cuda:
        return torch.as_tensor(_DevPtr(ptr, n), device="cuda")
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))
    return torch.from_numpy(arr)


class DistSolver:
    """All ranks construct this collectively with the same case; each advances the blocks it owns."""

    def __init__(self, case: Case, cfg: Optional[Config] = None, lib=None, group=None):
        self.lib = lib or _capi.load_gpu()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.cuda = self.lib.prefix == "par2d_"
        cfg = (cfg or default_config(self.lib)).copy()
        cfg.rank, cfg.nranks = self.rank, self.world
        if self.cuda:
            cfg.device = torch.cuda.current_device()
        self.owners = case.partition(self.world)
        self.solver = Solver(case, cfg, lib=self.lib, owners=self.owners)
        h = self.solver.handle
        if self.cuda:
            self.lib.check(self.lib.set_stream(h, torch.cuda.current_stream().cuda_stream), "set_stream")
        # exchange buffers per neighbouring rank
        self.send, self.recv = {}, {}
        for p in range(self.world):
            if p == self.rank:
                continue
            ns, nr = C.c_int64(), C.c_int64()
            ps, pr = C.c_void_p(), C.c_void_p()
            self.lib.check(self.lib.exchange_layout(h, p, C.byref(ns), C.byref(nr), C.byref(ps), C.byref(pr)), "exchange_layout")
            if ns.value != 0 or nr.value != 0:
                self.send[p] = _alias(ps.value, ns.value, self.cuda)
                self.recv[p] = _alias(pr.value, nr.value, self.cuda)
        pt, n = C.c_void_p(), C.c_int32()
        self.lib.check(self.lib.dt_table(h, C.byref(pt), C.byref(n)), "dt_table")
        self.dt_table = _alias(pt.value, n.value, self.cuda)
        self.nblocks = case.nblocks
        self.total_cells = case.ncells

    # -- one explicit step: pack -> transport -> ghost/residual/update ----------------------
    def _transport(self):
        dist.all_reduce(self.dt_table, op=dist.ReduceOp.MAX, group=self.group)
        ops = []
        for p in sorted(self.send):
            if self.send[p] is not None:
                ops.append(dist.P2POp(dist.isend, self.send[p], p, self.group))
            if self.recv[p] is not None:
                ops.append(dist.P2POp(dist.irecv, self.recv[p], p, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def queue_steps(self, nsteps: int):
        h = self.solver.handle
        for _ in range(nsteps):
            self.lib.check(self.lib.step_begin(h), "step_begin")
            self._transport()
            self.lib.check(self.lib.step_end(h), "step_end")

    def collect(self, nsteps: int, check: bool = True) -> List[StepInfo]:
        info = (StepInfo * nsteps)()
        rc = self.lib.step_collect(self.solver.handle, nsteps, info)
        if check:
            self.lib.check(rc, "step_collect")
        return list(info)

    def step(self, nsteps: int, check: bool = True):
        """Advance nsteps; returns (dt[nsteps], relres[nsteps]) with relres summed over ALL blocks
        in global block order on every rank."""
        self.queue_steps(nsteps)
        info = self.collect(nsteps, check=check)
        hist = self.solver.block_history(nsteps)            # (nsteps, nowned, 2)
        table = torch.zeros(nsteps, self.nblocks, 2, dtype=torch.float64)
        for k, g in enumerate(self.solver.owned):
            table[:, g, :] = torch.from_numpy(hist[:, k, :])
        if self.cuda:
            table = table.cuda()
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=self.group)   # disjoint supports: exact
        table = table.cpu().numpy()
        num = np.zeros(nsteps)
        den = np.zeros(nsteps)
        for g in range(self.nblocks):                                     # rank order of main.cpp:381-392
            num += table[:, g, 0]
            den += table[:, g, 1]
        return np.array([i.dt for i in info]), num / den

    def get_state(self, b: int) -> np.ndarray:
        return self.solver.get_state(b)

    @property
    def owned(self):
        return self.solver.owned

    def close(self):
        self.solver.close()
