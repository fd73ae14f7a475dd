"""Time the three-pass logits step (tt_simce3_*) kernel by kernel (CUPTI), print CTA 0's tile timeline, or run single passes for
ncu (TT_ONLY=p1|p2|p3). Usage: python scripts/time_simce3.py [B_local] [B_global]"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jodalrob_twotower_b200 as tt
from jodalrob_twotower_b200.native import ptr
lib = tt.native.lib(); DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
Bg = int(sys.argv[2]) if len(sys.argv) > 2 else B
F, T = 64, 1.0
g = torch.Generator().manual_seed(0)
c1 = math.log2(math.e) / T
n = torch.nn.functional.normalize(torch.randn(Bg, F, generator=g), dim=1)
c = torch.nn.functional.normalize(torch.randn(Bg, F, generator=g) + 0.5 * n, dim=1)
bf = lambda t: t.to(torch.bfloat16).to(DEV).contiguous()
N, C = bf(n), bf(c)
Ns, Cs = bf(n[:B] * c1), bf(c[:B] * c1)
N_loc, C_loc = N[:B].contiguous(), C[:B].contiguous()
f32 = dict(dtype=torch.float32, device=DEV)
nb = lib.query("tt_simce3_workspace_bytes", B, Bg, F)
ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
diag, rse, inv_r, cse, inv_c = [torch.empty(B, **f32) for _ in range(5)]
dN, dC = torch.empty(B, F, **f32), torch.empty(B, F, **f32)
xs0 = torch.empty(Bg, F, dtype=torch.bfloat16, device=DEV); xs1 = torch.empty_like(xs0)
xs0.copy_(N * 1e-4); xs1.copy_(C * 1e-4)
gs = 0.5 / Bg / T
single = B == Bg
s = torch.cuda.current_stream().cuda_stream
def p1(): lib.call("tt_simce3_rows", ptr(Ns), ptr(C), B, Bg, F, 1 / T, 0, ptr(diag), ptr(ws), nb, s)
def r1(): lib.call("tt_simce3_rows_finish", ptr(ws), nb, B, Bg, F, 1 / T, ptr(rse), ptr(inv_r), ptr(N_loc) if single else 0, gs, ptr(xs0) if single else 0, s)
def p2(): lib.call("tt_simce3_grad", ptr(Cs), ptr(N), ptr(xs0), B, Bg, F, 1 / T, 0, 0, gs, 0.0, ptr(ws), nb, s)
def r2(): lib.call("tt_simce3_grad_finish", ptr(ws), nb, B, Bg, F, 1 / T, 0, ptr(N), 1, gs, 0.0, Bg, ptr(dC), ptr(cse), ptr(inv_c), ptr(C_loc) if single else 0, ptr(xs1) if single else 0, s)
def p3(): lib.call("tt_simce3_grad", ptr(Ns), ptr(C), ptr(xs1), B, Bg, F, 1 / T, 0, ptr(inv_r), gs, 0.0, ptr(ws), nb, s)
def r3(): lib.call("tt_simce3_grad_finish", ptr(ws), nb, B, Bg, F, 1 / T, 0, ptr(C), 0, gs, 0.0, Bg, ptr(dN), 0, 0, 0, 0, s)
def step(): p1(); r1(); p2(); r2(); p3(); r3()
lib.call("tt_debug_simce_knobs", int(os.environ.get("TT_KNOBS", "0")))
only = os.environ.get("TT_ONLY")
if only:
    step(); torch.cuda.synchronize()  # ncu: -k regex:simce_tc_kernel -s 3 -c 1 captures the first launch below
    for _ in range(3): {"p1": p1, "p2": p2, "p3": p3}[only]()
    torch.cuda.synchronize(); sys.exit(0)
from torch.profiler import profile, ProfilerActivity
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): step()
    torch.cuda.synchronize()
print(f"B_local={B} B_global={Bg} F={F}")
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f"  {ev.key[:80]:80s} n={ev.count:3d} avg={ev.device_time_total / ev.count:8.1f} us")

def timeline(fn, name):
    ts = torch.zeros(4 * 2048, dtype=torch.int64, device=DEV)
    lib.call("tt_debug_simce_timeline", ptr(ts))
    fn(); torch.cuda.synchronize()
    lib.call("tt_debug_simce_timeline", 0)
    t = ts.cpu().view(4, 512, 4)
    mma, epi, ld, mma2 = t[0], t[1], t[2], t[3]
    nst = int((mma[:, 3] > 0).sum())
    t0 = int(mma[0, 0])
    print(name, "tiles stamped:", nst)
    print("tile | MMA1: start yfull-wait sempty-wait issue+commit | EPI: top  ld+sync+pwait  half1  ldwait+release+half2 | LD: yempty-wait | MMA2: start pfull-wait issue+commit")
    for i in range(min(nst, 30)):
        m, e, l, m2 = mma[i], epi[i], ld[i], mma2[i]
        print(f"{i:3d} | {int(m[0])-t0:7d} {int(m[1]-m[0]):6d} {int(m[2]-m[1]):6d} {int(m[3]-m[2]):6d} | {int(e[0])-t0:7d} {int(e[1]-e[0]):6d} {int(e[2]-e[1]):6d} {int(e[3]-e[2]):6d} | {int(l[0])-t0:7d} {int(l[1]-l[0]):6d} | {int(m2[0])-t0:7d} {int(m2[1]-m2[0]):6d} {int(m2[2]-m2[1]):6d}")
    w = t[2][256:256 + 148]
    w = w[w[:, 0] > 0]
    if len(w):
        k0 = int(w[:, 0].min())
        print(f"{name} CTAs {len(w)}: last entry {int(w[:,0].max())-k0} ns, set-up done (mean) {float((w[:,1]-w[:,0]).float().mean()):.0f} ns after entry, "
              f"first exit {int(w[:,2].min())-k0} ns, last exit {int(w[:,2].max())-k0} ns")
    it = t[1][256:]
    print(name, "items of CTA 0: start | prologue | tiles | drain  (cycles)")
    for i in range(8):
        if int(it[i, 0]) == 0: break
        print(f"  {i}: {int(it[i,0])-t0:7d} | {int(it[i,1]-it[i,0]):6d} | {int(it[i,2]-it[i,1]):6d} | {int(it[i,3]-it[i,2]):6d}")
timeline(p1, "P1")
timeline(p2, "P2")
timeline(p3, "P3")
