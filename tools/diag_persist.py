"""Per-step cycle breakdown of the persistent launch on a shipped case (PAR2D_PROFILE_PERSIST=1)."""
import ctypes as C, os, sys, time
os.environ["PAR2D_PROFILE_PERSIST"] = "1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from par2d_b200 import Case, Solver, _capi, default_config
gpu = _capi.load_gpu()
for name in sys.argv[1:] or ["Case_ffstep8"]:
    case = Case.load(os.path.join(REPO, "tests", "golden", "cases", name))
    cfg = default_config(gpu); cfg.device = 0
    s = Solver(case, cfg, lib=gpu)
    s.step(200)
    t0 = time.perf_counter(); s.step(4000); dt = time.perf_counter() - t0
    out = (C.c_int64 * (12 * 4000))()
    gpu.check(gpu.persist_profile(s.handle, out, 4000), "persist_profile")
    a = np.array(list(out)).reshape(-1, 12)[100:]
    print(f"{name}: {1e6 * dt / 4000:.2f} us/step wall; cycles/step median: march {np.median(a[:,0]):.0f}  barrier wait {np.median(a[:,1]):.0f}  bookkeeping {np.median(a[:,2]):.0f}  | in strip (cumulative): wait rows {np.median(a[:,3]):.0f}  +rows {np.median(a[:,4]):.0f}  +norms {np.median(a[:,5]):.0f}  +epilogue {np.median(a[:,6]):.0f} | inside the epilogue (cumulative): entry+syncwarp {np.median(a[:,7]):.0f}  left/right {np.median(a[:,8]):.0f}  bottom/top {np.median(a[:,9]):.0f}  dt+ticket {np.median(a[:,10]):.0f}")
    s.close()
