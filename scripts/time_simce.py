"""Time the tcgen05 sim-CE kernels alone (CUDA events), optionally with profiling knobs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jodalrob_twotower_b200 as tt
from jodalrob_twotower_b200.native import ptr, TT_BF16
lib = tt.native.lib(); DEV = "cuda:0"
B, F = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 64
g = torch.Generator().manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(B, F, generator=g), dim=1).to(torch.bfloat16).to(DEV)
y = torch.nn.functional.normalize(torch.randn(B, F, generator=g), dim=1).to(torch.bfloat16).to(DEV)
f32 = dict(dtype=torch.float32, device=DEV)
rse, cse, rss, css, dg = [torch.empty(B, **f32) for _ in range(5)]
arg = torch.empty(B, dtype=torch.int32, device=DEV)
ws = torch.empty(lib.query("tt_simce_workspace_bytes", B, B, F), dtype=torch.uint8, device=DEV)
inv = torch.full((B,), 1.0 / B, **f32); dX = torch.empty(B, F, **f32)
s = torch.cuda.current_stream().cuda_stream
def fwd(): lib.call("tt_simce_forward", ptr(x), ptr(y), B, B, F, TT_BF16, 1.0, 0, ptr(rse), ptr(cse), ptr(rss), ptr(css), ptr(dg), ptr(arg), ptr(ws), ws.numel(), s)
def bwd(): lib.call("tt_simce_backward", ptr(x), ptr(y), B, B, F, TT_BF16, 1.0, 0, ptr(inv), ptr(inv), 0.5 / B, 0.0, B, ptr(dX), ptr(ws), ws.numel(), s)
if os.environ.get("TT_ONLY") == "fwd":  # for ncu: just the forward kernels
    for _ in range(3): fwd()
    torch.cuda.synchronize(); sys.exit(0)
if os.environ.get("TT_ONLY") == "bwd":  # for ncu: just the backward kernels
    fwd(); torch.cuda.synchronize()
    for _ in range(3): bwd()
    torch.cuda.synchronize(); sys.exit(0)
def timeit(fn, n=20):
    """device-bound timing: n calls captured in one CUDA graph, replayed 5 times"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    global s
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(n): fn()
    s = torch.cuda.current_stream().cuda_stream
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [g.replay() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1000
for knobs in (0,):
    lib.call("tt_debug_simce_knobs", knobs)
    print(f"knobs={knobs}: fwd call (2 passes + reduces) {timeit(fwd):7.1f} us   bwd call {timeit(bwd):7.1f} us", flush=True)
lib.call("tt_debug_simce_knobs", 0)

# ---- per-kernel device durations (CUPTI through torch.profiler)
if os.environ.get("TT_KINETO", "1") == "1":
    from torch.profiler import profile, ProfilerActivity
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3): fwd(); bwd()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10): fwd(); bwd()
        torch.cuda.synchronize()
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
        print(f"  {ev.key[:70]:70s} n={ev.count:3d} avg={ev.device_time_total / ev.count:8.1f} us")

# ---- timeline of CTA 0: where does a tile's time go
def timeline(fn, name):
    ts = torch.zeros(4 * 2048, dtype=torch.int64, device=DEV)
    lib.call("tt_debug_simce_timeline", ptr(ts))
    fn(); torch.cuda.synchronize()
    lib.call("tt_debug_simce_timeline", 0)
    t = ts.cpu().view(4, 512, 4)
    mma, epi, ld, mma2 = t[0], t[1], t[2], t[3]
    n = int((mma[:, 3] > 0).sum())
    t0 = int(mma[0, 0])
    print(name, "tiles stamped:", n)
    print("tile | MMA1: start yfull-wait sempty-wait issue+commit | EPI: top  ld+sync+pwait  half1  ldwait+release+half2 | LD: yempty-wait | MMA2: start pfull-wait issue+commit")
    for i in range(min(n, 30)):
        m, e, l, m2 = mma[i], epi[i], ld[i], mma2[i]
        print(f"{i:3d} | {int(m[0])-t0:7d} {int(m[1]-m[0]):6d} {int(m[2]-m[1]):6d} {int(m[3]-m[2]):6d} | {int(e[0])-t0:7d} {int(e[1]-e[0]):6d} {int(e[2]-e[1]):6d} {int(e[3]-e[2]):6d} | {int(l[0])-t0:7d} {int(l[1]-l[0]):6d} | {int(m2[0])-t0:7d} {int(m2[1]-m2[0]):6d} {int(m2[2]-m2[1]):6d}")
s = torch.cuda.current_stream().cuda_stream
def items(fn, name):
    ts = torch.zeros(4 * 2048, dtype=torch.int64, device=DEV)
    lib.call("tt_debug_simce_timeline", ptr(ts))
    fn(); torch.cuda.synchronize()
    lib.call("tt_debug_simce_timeline", 0)
    t = ts.cpu().view(4, 512, 4)
    t0 = int(t[0][0, 0])
    it = t[1][256:]
    w = t[2][256:256 + 148]
    w = w[w[:, 0] > 0]
    if len(w):
        k0 = int(w[:, 0].min())
        print(f"{name} CTAs {len(w)}: first entry 0, last entry {int(w[:,0].max())-k0} ns, set-up done (mean) {float((w[:,1]-w[:,0]).float().mean()):.0f} ns after entry, "
              f"first exit {int(w[:,2].min())-k0} ns, last exit {int(w[:,2].max())-k0} ns")
    print(name, "items of CTA 0: start | prologue | tiles | drain  (cycles)")
    for i in range(8):
        if int(it[i, 0]) == 0: break
        print(f"  {i}: {int(it[i,0])-t0:7d} | {int(it[i,1]-it[i,0]):6d} | {int(it[i,2]-it[i,1]):6d} | {int(it[i,3]-it[i,2]):6d}")
timeline(fwd, "fwd")
timeline(bwd, "bwd")
items(bwd, "bwd")
items(fwd, "fwd")
