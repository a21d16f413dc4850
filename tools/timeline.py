"""In-situ kernel timeline of the benchmark step loop (CUPTI through torch.profiler): per kernel name the mean duration
and the mean idle gap before it.  Run from the repo root (or from a copy of an older tree) on the GPU box."""
import os, sys, collections
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
from par2d_b200 import Case, Solver, _capi, default_config
gpu = _capi.load_gpu()
blk = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
case = Case.synthetic_ramp(blk * 4, blk * 2, 4, 2)
cfg = default_config(gpu); cfg.device = 0
s = Solver(case, cfg, lib=gpu)
s.step(300)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    s.step(40)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
dur, gap, n = collections.defaultdict(float), collections.defaultdict(float), collections.Counter()
prev_end = None
for e in ev:
    name = e.name.split("(")[0][-40:]
    st, en = e.time_range.start, e.time_range.end
    dur[name] += en - st; n[name] += 1
    if prev_end is not None: gap[name] += max(st - prev_end, 0)
    prev_end = en if prev_end is None else max(prev_end, en)
tot = (ev[-1].time_range.end - ev[0].time_range.start)
print(f"{len(ev)} GPU activities over {tot:.0f} us")
for k in sorted(n, key=lambda k: -dur[k]):
    print(f"  {k:42s} n={n[k]:4d}  mean duration {dur[k]/n[k]:9.2f} us   mean gap before {gap[k]/n[k]:7.2f} us")
