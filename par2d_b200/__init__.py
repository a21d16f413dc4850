"""par2d_b200 - B200-native explicit finite-volume residual+update path of PAR2D.

Layout (only what the hot path needs):
  csrc/        CUDA kernels + the C ABI (include/par2d_gpu.h) and the C++ host
               (case I/O, driver) behind include/par2d_host.h
  _capi.py     ctypes bindings of both headers
  host.py      case directories, synthetic meshes, partitioning
  solver.py    the reference driver loop re-expressed over the C ABI
  dist.py      one-process-per-GPU driver (torch.distributed for the plumbing)
"""
from ._capi import Config, BlockDesc, StepInfo, Par2dError, MODE_FAST, MODE_FAITHFUL  # noqa: F401
from .host import Case, load_thermo, write_solution  # noqa: F401
from .solver import Solver, default_config  # noqa: F401
