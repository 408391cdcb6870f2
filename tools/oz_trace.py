"""Kernel timeline (CUPTI via torch.profiler) of one logpdf at n=16384: name, start, duration, stream."""
import sys, json, torch
import stheno_b200.torch as S
from stheno_b200 import B as Bns
from torch.profiler import profile, ProfilerActivity
Bns.precision = sys.argv[1] if len(sys.argv) > 1 else "auto"
n, d = 16384, 8
g = torch.Generator().manual_seed(1)
x = torch.rand(n, d, generator=g, dtype=torch.float64).cuda()
y = torch.randn(n, generator=g, dtype=torch.float64).cuda()
f = S.GP(S.EQ().stretch(1.5))
for _ in range(3):
    lp = f(x, 0.1).logpdf(y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    lp = f(x, 0.1).logpdf(y)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        ev.append((e.name[:60], e.time_range.start, e.time_range.end - e.time_range.start, getattr(e, "stream", -1)))
ev.sort(key=lambda t: t[1])
t0 = ev[0][1]
out = [{"name": n_, "t": (s - t0), "dur": d_, "stream": st} for n_, s, d_, st in ev]
json.dump(out, open("gpurun_out/oz_trace.json", "w"))
print(len(out), "kernels; span", (ev[-1][1] + ev[-1][2] - t0) / 1e3, "ms")
