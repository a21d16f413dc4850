"""One-process-per-GPU driver: GridPart-decomposed blocks sharded over the ranks of one box.

Replaces the reference's MPI plumbing (one rank per block, FV/main.cpp:68-74,125-128).  Two transports:
  "peer" (default on GPUs): every rank exports its receive arena (cudaIpc handle, all-gathered here
                     once), then the library's own kernels store edge lines and dt_b into the neighbours'
                     arenas over NVLink and a whole batch of steps runs without the host (par2d_peer_*);
  "nccl"/"gloo"    : external transport per step -
  * halo exchange  : the four MPI_Sendrecv sites (SpatialDiscretization.cpp:259,320,380,441)
                     become one grouped NCCL send/recv per neighbouring rank on the device
                     buffers the C ABI exposes (par2d_exchange_layout);
  * dt sync        : the gather/fmax/bcast chain (main.cpp:270-292) becomes one all-reduce(MAX)
                     over the per-block dt table (par2d_dt_table);
and for both
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


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


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

    def __init__(self, case: Case, cfg: Optional[Config] = None, lib=None, group=None, transport: str = "auto"):
        """transport: "peer" (NVLink peer-mapped arenas; raises if the GPUs have no peer access),
        "external" (torch.distributed send/recv per step) or "auto" (peer on GPUs, falling back to
        external with a printed note; always external for the CPU oracle used by the gloo tests)."""
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
        self.stream = None
        if self.cuda:
            # The external transport interleaves library kernels with NCCL operations, which torch orders
            # against the CURRENT torch stream: the library must therefore run on a real torch stream (the
            # legacy default stream's handle is 0 = "private stream" for par2d_set_stream and would not be
            # ordered with the NCCL work at all).
            self.stream = torch.cuda.Stream()
            self.lib.check(self.lib.set_stream(h, self.stream.cuda_stream), "set_stream")
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
        self.transport = "external"
        if self.cuda and transport in ("auto", "peer"):
            self._connect_peers(strict=(transport == "peer"))

    def _connect_peers(self, strict: bool):
        h = self.solver.handle
        mine = (C.c_ubyte * 64)()
        rc = self.lib.peer_export(h, mine)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device="cuda")
        handles = torch.zeros(self.world, 64, dtype=torch.uint8, device="cuda")
        handles[self.rank] = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device="cuda")
        dist.all_reduce(handles, op=dist.ReduceOp.SUM, group=self.group)      # disjoint rows: an all-gather
        if rc == 0:
            buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(bytes(handles.cpu().numpy().tobytes()))
            rc = self.lib.peer_connect(h, buf)
            ok[0] = 1 if rc == 0 else 0
        err = self.lib.last_error()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)            # all or nobody
        if int(ok.item()) == 1:
            # every rank is connected: the strip decomposition (summation order of the norms) must agree
            ok[0] = 1 if self.lib.peer_check(h) == 0 else 0
            err = self.lib.last_error()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 1:
                self.transport = "peer"
                return
        self.lib.peer_disconnect(h)
        msg = f"peer-mapped exchange unavailable ({err.decode() if err else 'a peer failed'})"
        if strict:
            raise RuntimeError(msg)
        if self.rank == 0:
            print(f"par2d_b200.dist: {msg}; using the external NCCL transport", flush=True)

    # -- one explicit step: pack -> transport -> ghost/residual/update ----------------------
    def _on_stream(self):
        return torch.cuda.stream(self.stream) if self.stream is not None else _NullCtx()

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

    def advance(self, nsteps: int, check: bool = True, stop_on_tol: bool = False) -> List[StepInfo]:
        """Advance nsteps on every rank (collective); returns this rank's log records (local sums).
        stop_on_tol (peer transport): every rank stops after the iteration whose GLOBAL relres fell below
        cfg.tol (main.cpp:506-531), decided on the device from the sums of all ranks."""
        if self.transport == "peer":
            info = (StepInfo * nsteps)()
            rc = self.lib.step(self.solver.handle, nsteps, int(stop_on_tol), info)
            if check:
                self.lib.check(rc, "step")
            return [i for i in info if i.iter >= 0]
        if stop_on_tol:
            raise ValueError("stop_on_tol across ranks needs the peer transport")
        self.queue_steps(nsteps)
        return self.collect(nsteps, check=check)

    def queue_steps(self, nsteps: int):
        h = self.solver.handle
        with self._on_stream():
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

    def global_relres(self, nsteps: int) -> np.ndarray:
        """relres of the last nsteps steps with the per-block sums of ALL ranks added in global block order
        (the rank-order gather of main.cpp:379-392), identical on every rank and for any number of ranks."""
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
        return num / den

    def step(self, nsteps: int, check: bool = True, stop_on_tol: bool = False):
        """Advance nsteps; returns (dt[n], relres[n]) with relres summed over ALL blocks in global block
        order on every rank (n < nsteps if the run stopped on the tolerance)."""
        info = self.advance(nsteps, check=check, stop_on_tol=stop_on_tol)
        return np.array([i.dt for i in info]), self.global_relres(len(info))

    def get_state(self, b: int) -> np.ndarray:
        return self.solver.get_state(b)

    @property
    def owned(self):
        return self.solver.owned

    def close(self):
        if self.transport == "peer":
            # unmap the peers' arenas everywhere before anybody frees its own
            self.lib.peer_disconnect(self.solver.handle)
            dist.barrier(group=self.group)
            self.transport = "external"
        self.solver.close()
