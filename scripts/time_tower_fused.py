"""Stand-alone device time of the fused tower kernels and the GEMMs around them (CUDA graph of N launches, CUDA events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jodalrob_twotower_b200 as tt
lib = tt.native.lib(); DEV = "cuda:0"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
H0, H1, F = 128, 64, 64
P = lambda t: 0 if t is None else t.data_ptr()
g = torch.Generator(device=DEV).manual_seed(0)
f32 = dict(dtype=torch.float32, device=DEV); bf = dict(dtype=torch.bfloat16, device=DEV)
nt = lib.query("tt_linear_stats_tiles", M)
hyper = torch.zeros(64, dtype=torch.uint8, device=DEV)
def timeit(fn, n=20):
    for _ in range(3): fn(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(n): fn(s)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [gr.replay() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1000
for name, LD in (("notice", 1152), ("company", 320)):
    c = torch.randn(M, LD, generator=g, device=DEV).to(torch.bfloat16)
    W1 = (torch.randn(H1, LD, generator=g, device=DEV) / LD ** 0.5).to(torch.bfloat16)
    W2 = (torch.randn(F, H1, generator=g, device=DEV) / 8).to(torch.bfloat16)
    b1, b2 = torch.zeros(H1, **f32), torch.zeros(F, **f32)
    gamma, beta = torch.ones(H1, **f32), torch.zeros(H1, **f32)
    h = torch.empty(M, H1, **bf); y = torch.empty(M, H1, **bf); e = torch.empty(M, F, **bf); es = torch.empty(M, F, **bf)
    part = torch.zeros(nt * 2 * H1, **f32); bn_part = torch.zeros(nt * 2 * H1, **f32)
    mean, rstd, nrm = torch.empty(H1, **f32), torch.empty(H1, **f32), torch.empty(M, **f32)
    rm, rv = torch.zeros(H1, **f32), torch.ones(H1, **f32); nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    de = torch.randn(M, F, generator=g, device=DEV) * 1e-3
    dz16 = torch.empty(M, F, **bf); dyp = torch.empty(M, H1, **bf); dh = torch.empty(M, H1, **bf); dc = torch.empty(M, LD, **bf)
    dz_cp, dh_cp, dc_cp = torch.zeros(nt * F, **f32), torch.zeros(nt * H1, **f32), torch.zeros(nt * H0, **f32)
    dg, db, dbias2 = torch.empty(H1, **f32), torch.empty(H1, **f32), torch.empty(F, **f32)
    for p in (0.0, 0.1):
        t1 = timeit(lambda s: lib.call("tt_linear_forward_bf16_stats", P(c), LD, P(W1), P(b1), P(h), H1, 1, M, H1, LD, 1, P(part), s))
        t1b = timeit(lambda s: lib.call("tt_linear_forward_bf16", P(c), LD, P(W1), P(b1), P(h), H1, 1, M, H1, LD, 1, s))
        t2 = timeit(lambda s: lib.call("tt_tower_tail_forward", P(h), M, H1, F, P(part), nt, 0, M, P(gamma), P(beta), 1e-5, 0.1, P(mean), P(rstd), P(rm), P(rv),
                                       P(nbt), p, 0, P(hyper), 3, P(y), P(W2), P(b2), P(e), P(es), 1.44, P(nrm), s))
        t3 = timeit(lambda s: lib.call("tt_tower_head_backward", P(de), P(e), P(nrm), M, F, H1, P(W2), P(dz16), P(dz_cp), P(h), P(mean), P(rstd), p, 0, P(hyper), 3,
                                       P(dyp), P(bn_part), s))
        t4 = timeit(lambda s: lib.call("tt_tower_mid_backward", P(dyp), P(h), M, H1, LD, P(W1), P(bn_part), nt, 0, M, P(mean), P(rstd), P(gamma), P(dg), P(db), 1.0,
                                       P(dz_cp), F, P(dbias2), P(dh), P(dh_cp), P(dc), H0, P(dc_cp), s))
        print(f"{name} M={M} LD={LD} p={p}: GEMM1+stats {t1:6.1f} us (plain {t1b:6.1f})  tail fwd {t2:6.1f}  head bwd {t3:6.1f}  mid bwd {t4:6.1f}", flush=True)
