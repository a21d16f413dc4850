"""Per-step cycle breakdown of the persistent launch on a shipped case (PAR2D_PROFILE_PERSIST=1): the first strip and the
first bookkeeping CTA over 2000 steps, and every CTA of the launch for one step (march time, wait at the grid barrier)."""
import ctypes as C, os, sys, time
os.environ["PAR2D_PROFILE_PERSIST"] = "1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from par2d_b200 import Case, Solver, _capi, default_config
gpu = _capi.load_gpu()
NSTEP, HALF = 1000, 2048
for name in sys.argv[1:] or ["Case_ffstep8"]:
    case = Case.load(os.path.join(REPO, "tests", "golden", "cases", name))
    cfg = default_config(gpu); cfg.device = 0
    s = Solver(case, cfg, lib=gpu)
    s.step(200)
    t0 = time.perf_counter(); s.step(NSTEP); dt = time.perf_counter() - t0
    out = (C.c_int64 * (12 * 4096))()
    gpu.check(gpu.persist_profile(s.handle, out, 4096), "persist_profile")
    raw = np.array(list(out))
    a = raw[:12 * NSTEP].reshape(-1, 12)[100:]
    m = lambda k: np.median(a[:, k])
    print(f"{name}: {1e6 * dt / NSTEP:.2f} us/step wall; cycles/step median: march {m(0):.0f}  barrier wait {m(1):.0f}  bookkeeping {m(2):.0f}  | in strip (cumulative): "
          f"wait rows {m(3):.0f}  +rows {m(4):.0f}  +norms {m(5):.0f}  +epilogue {m(6):.0f} | inside the epilogue (cumulative): entry+syncwarp {m(7):.0f}  "
          f"ghost records {m(8):.0f}  dt+ticket {m(10):.0f} | before the rows: strip entry -> rows requested {m(9):.0f}  -> dt known {m(2):.0f} | barrier: arrival done after {m(11):.0f}")
    per = raw[12 * HALF:].reshape(-1, 2)
    n = int(np.count_nonzero(per[:, 0]))
    if n:
        work, wait = per[:n, 0], per[:n, 1]
        q = lambda v: " ".join(f"{np.percentile(v, p):.0f}" for p in (0, 10, 50, 90, 99, 100))
        order = np.argsort(-work)[:8]
        print(f"   all {n} CTAs at step 500: work cycles (min p10 p50 p90 p99 max) {q(work)} | barrier wait {q(wait)} | work+wait {q(work + wait)}")
        print("   longest CTAs (cta: work/wait): " + "  ".join(f"{int(i)}: {int(work[i])}/{int(wait[i])}" for i in order))
    fin = raw[12 * HALF + 8192:12 * HALF + 8192 + 16 * 1024].reshape(-1, 2, 8)[100:900]
    for blk in (0, 1):
        f = fin[:, blk, :]
        for last in (0, 1):
            sel = f[f[:, 2] == last]
            if len(sel):
                print(f"   bookkeeping CTA {blk}, last={last} ({len(sel)} steps): cumulative cycles  sums+tree+sqrt {np.median(sel[:,0]):.0f}  +fence+ticket {np.median(sel[:,1]):.0f}"
                      + (f"  +fence+counters+block sums {np.median(sel[:,3]):.0f}  +history record {np.median(sel[:,4]):.0f}" if last else ""))
    s.close()
