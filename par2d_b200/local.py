"""Single-process multi-GPU driver: N contexts, one per GPU, connected through peer-mapped arenas.

The Python twin of ``par2d_fv --gpus N`` (par2d_b200/csrc/host/main.cpp): replaces the reference's
rank setup (FV/main.cpp:68-74,125-142) without a second process - par2d_peer_connect_local enables
peer access between the devices, every batch is advanced by one host thread per context, the residual is gathered in global block order (main.cpp:379-392) and
the tolerance break (main.cpp:506-531) is decided on the device from the sums of all contexts.
Context r runs on ``devices[r % len(devices)]``: with fewer devices than contexts several contexts
share a GPU (their kernels run side by side) - that is how the exchange protocol is tested on a
one-GPU box.  No torch, no torch.distributed.
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import Config, StepInfo
from .host import Case
from .solver import Solver, default_config


class LocalMultiSolver:
    def __init__(self, case: Case, cfg: Optional[Config] = None, nranks: int = 2, devices: Optional[Sequence[int]] = None, lib=None):
        self.lib = lib or _capi.load_gpu()
        cfg = (cfg or default_config(self.lib)).copy()
        if devices is None:
            n = C.c_int32()
            self.lib.check(self.lib.device_count(C.byref(n)), "device_count")
            devices = list(range(n.value))
        self.nranks = nranks
        self.owners = case.partition(nranks)
        self.nblocks = case.nblocks
        self.solvers: List[Solver] = []
        for r in range(nranks):
            c = cfg.copy()
            c.rank, c.nranks, c.device = r, nranks, devices[r % len(devices)]
            self.solvers.append(Solver(case, c, lib=self.lib, owners=self.owners))
        arr = (C.c_void_p * nranks)(*[s.handle for s in self.solvers])
        self.lib.check(self.lib.peer_connect_local(arr, nranks), "peer_connect_local")
        # contexts that SHARE a GPU (tests on a one-GPU box) must not run far ahead of each other: a kernel spinning on a
        # peer's flag can sit in front of the very kernel that would raise it in the GPU's work queues.  Short batches
        # keep the queues shallow; with one GPU per context there is no such limit.
        self.max_batch = 16 if len(set(devices[r % len(devices)] for r in range(nranks))) < nranks else None
        self.connected = True
        self._pool = ThreadPoolExecutor(max_workers=nranks)

    def _of(self, b: int) -> Solver:
        return self.solvers[self.owners[b]]

    def step(self, nsteps: int, stop_on_tol: bool = False):
        """Advance nsteps on every context; returns (dt[n], relres[n]) with relres over ALL blocks in
        global block order; n < nsteps if the run stopped on the tolerance."""
        if self.max_batch and nsteps > self.max_batch:
            dts, rels = [], []
            left = nsteps
            while left > 0:
                n = min(left, self.max_batch)
                d, r = self.step(n, stop_on_tol)
                dts += list(d)
                rels += list(r)
                left -= n
                if len(d) < n or (stop_on_tol and len(r) and r[-1] < self.solvers[0].cfg.tol):
                    break
            return np.array(dts), np.array(rels)

        def run(s):
            info = (StepInfo * max(nsteps, 1))()
            rc = self.lib.step(s.handle, nsteps, int(stop_on_tol), info)
            msg = self.lib.last_error() if rc else b""          # the error text is per thread
            return rc, msg, [i for i in info[:nsteps] if i.iter >= 0]

        results = list(self._pool.map(run, self.solvers))
        for rc, msg, _ in results:
            if rc:
                raise _capi.Par2dError(rc, "step", msg.decode() if msg else "")
        infos = [r[2] for r in results]
        done = len(infos[0])
        assert all(len(i) == done for i in infos), "the contexts stopped after different iterations"
        table = np.zeros((done, self.nblocks, 2))
        for s in self.solvers:
            if done:
                hist = s.block_history(done)
                for k, g in enumerate(s.owned):
                    table[:, g, :] = hist[:, k, :]
        num, den = np.zeros(done), np.zeros(done)
        for g in range(self.nblocks):          # rank order of main.cpp:381-392
            num += table[:, g, 0]
            den += table[:, g, 1]
        self.infos = infos
        return np.array([i.dt for i in infos[0]]), (num / den if done else num)

    def get_state(self, b: int) -> np.ndarray:
        return self._of(b).get_state(b)

    def set_state(self, b: int, unk: np.ndarray):
        self._of(b).set_state(b, unk)

    def get_residual(self, b: int) -> np.ndarray:
        return self._of(b).get_residual(b)

    @property
    def iteration(self) -> int:
        return self.solvers[0].iteration

    def close(self):
        if self.connected:
            for s in self.solvers:
                self.lib.peer_disconnect(s.handle)
            self.connected = False
        for s in self.solvers:
            s.close()
        self._pool.shutdown()
