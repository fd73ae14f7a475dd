"""Fused tower kernels (csrc/tower_fused.cu + the statistics epilogue of the tcgen05 GEMM) against plain torch math on the same
bf16 operands: Linear -> ReLU -> BatchNorm1d -> Dropout -> Linear -> F.normalize of base_tower.py:87-97,142-145 and its backward.
Every stage's stored tensor is compared (so a failure names the stage); the end-to-end trajectory tests are in
test_gpu_bench_path.py. Tolerances: bf16 tensors within one bf16 ulp of the fp32 math (2^-8 relative to max), fp32 reductions 1e-5."""
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16_ULP = 2.0 ** -8


def _lib():
    import jodalrob_twotower_b200 as tt
    return tt.native.lib()


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def _fold(part, n_tiles, width):
    return part.view(n_tiles, -1, width).double().sum(0)


@pytest.mark.parametrize("M,N,K", [(8192, 64, 1152), (1000, 64, 320), (300, 128, 256), (8192, 128, 128)])
def test_linear_forward_stats_epilogue(M, N, K):
    lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(M + N)
    A = torch.randn(M, K, generator=g, device=DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device=DEV) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device=DEV)
    C0 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    C1 = torch.empty_like(C0)
    nt = lib.query("tt_linear_stats_tiles", M)
    part = torch.full((nt * 2 * N,), float("nan"), device=DEV)
    lib.call("tt_linear_forward_bf16", _ptr(A), K, _ptr(W), _ptr(b), _ptr(C0), N, 1, M, N, K, 1, _s())
    lib.call("tt_linear_forward_bf16_stats", _ptr(A), K, _ptr(W), _ptr(b), _ptr(C1), N, 1, M, N, K, 1, _ptr(part), _s())
    assert torch.equal(C0, C1)
    ref = torch.relu(A.float() @ W.float().t() + b)
    assert rel_err(C1.float(), ref) < BF16_ULP
    sums = _fold(part, nt, N)
    assert rel_err(sums[0], C1.double().sum(0)) < 1e-5
    assert rel_err(sums[1], C1.double().pow(2).sum(0)) < 1e-5


def _tail_inputs(M, H1, F, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    h = torch.relu(torch.randn(M, H1, generator=g, device=DEV) + 0.3).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(H1, generator=g, device=DEV)
    beta = 0.1 * torch.randn(H1, generator=g, device=DEV)
    W2 = (torch.randn(F, H1, generator=g, device=DEV) / H1 ** 0.5).to(torch.bfloat16)
    b2 = 0.1 * torch.randn(F, generator=g, device=DEV)
    return g, h, gamma, beta, W2, b2


@pytest.mark.parametrize("sync_form", [False, True])
@pytest.mark.parametrize("M,H1,F,p", [(8192, 64, 64, 0.0), (1000, 64, 64, 0.25), (640, 128, 128, 0.25), (130, 64, 128, 0.0), (8192, 128, 64, 0.1)])
def test_tower_tail_forward(M, H1, F, p, sync_form):
    lib = _lib()
    g, h, gamma, beta, W2, b2 = _tail_inputs(M, H1, F, M + H1 + F)
    nt = lib.query("tt_linear_stats_tiles", M)
    # statistics partials exactly as the GEMM epilogue leaves them
    hp = torch.zeros(nt * 128, H1, device=DEV)
    hp[:M] = h.float()
    part = torch.stack([hp.view(nt, 128, H1).sum(1), hp.view(nt, 128, H1).pow(2).sum(1)], 1).contiguous().view(-1)
    sums = torch.cat([h.double().sum(0), h.double().pow(2).sum(0)]) if sync_form else None
    mask = (torch.rand(M, H1, generator=g, device=DEV) >= p).to(torch.uint8) if p > 0 else None
    f32 = dict(dtype=torch.float32, device=DEV)
    mean, rstd, nrm = torch.empty(H1, **f32), torch.empty(H1, **f32), torch.empty(M, **f32)
    rm, rv = torch.zeros(H1, **f32), torch.ones(H1, **f32)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    y = torch.empty(M, H1, dtype=torch.bfloat16, device=DEV)
    e, es = torch.empty(M, F, dtype=torch.bfloat16, device=DEV), torch.empty(M, F, dtype=torch.bfloat16, device=DEV)
    lib.call("tt_tower_tail_forward", _ptr(h), M, H1, F, _ptr(part), nt, _ptr(sums), M, _ptr(gamma), _ptr(beta), 1e-5, 0.1, _ptr(mean),
             _ptr(rstd), _ptr(rm), _ptr(rv), _ptr(nbt), p, _ptr(mask), 0, 0, _ptr(y), _ptr(W2), _ptr(b2), _ptr(e), _ptr(es), 1.4427,
             _ptr(nrm), _s())
    hd = h.double()
    mu, var = hd.mean(0), hd.var(0, unbiased=False)
    assert rel_err(mean, mu) < 1e-5 and rel_err(rstd, 1 / (var + 1e-5).sqrt()) < 1e-5
    assert rel_err(rm, 0.1 * mu) < 1e-5 and rel_err(rv, 0.9 + 0.1 * hd.var(0, unbiased=True)) < 1e-5 and int(nbt) == 1
    y_ref = ((hd - mu) / (var + 1e-5).sqrt() * gamma.double() + beta.double())
    if p > 0:
        y_ref = y_ref * mask.double() / (1 - p)
    assert rel_err(y.double(), y_ref) < BF16_ULP
    z = y.double() @ W2.double().t() + b2.double()     # from the bf16 y the kernel itself consumed
    n_ref = z.norm(dim=1).clamp_min(1e-12)
    assert rel_err(nrm, n_ref) < 1e-5
    assert rel_err(e.double(), z / n_ref[:, None]) < BF16_ULP
    assert rel_err(es.double(), z / n_ref[:, None] * 1.4427) < BF16_ULP


@pytest.mark.parametrize("M,H1,F,p", [(8192, 64, 64, 0.0), (1000, 64, 64, 0.25), (640, 128, 128, 0.25), (130, 128, 64, 0.0)])
def test_tower_head_backward(M, H1, F, p):
    lib = _lib()
    g, h, gamma, beta, W2, b2 = _tail_inputs(M, H1, F, 2 * M + H1)
    f32 = dict(dtype=torch.float32, device=DEV)
    z = torch.randn(M, F, generator=g, device=DEV)
    nrm = z.norm(dim=1)
    e = (z / nrm[:, None]).to(torch.bfloat16)
    de = torch.randn(M, F, generator=g, device=DEV) * 1e-3
    mean = h.float().mean(0)
    rstd = 1 / (h.float().var(0, unbiased=False) + 1e-5).sqrt()
    mask = (torch.rand(M, H1, generator=g, device=DEV) >= p).to(torch.uint8) if p > 0 else None
    nt = lib.query("tt_linear_stats_tiles", M)
    dz16 = torch.empty(M, F, dtype=torch.bfloat16, device=DEV)
    dyp = torch.empty(M, H1, dtype=torch.bfloat16, device=DEV)
    dz_cp = torch.full((nt * F,), float("nan"), **f32)
    bn_part = torch.full((nt * 2 * H1,), float("nan"), **f32)
    lib.call("tt_tower_head_backward", _ptr(de), _ptr(e), _ptr(nrm), M, F, H1, _ptr(W2), _ptr(dz16), _ptr(dz_cp), _ptr(h), _ptr(mean),
             _ptr(rstd), p, _ptr(mask), 0, 0, _ptr(dyp), _ptr(bn_part), _s())
    ed, ded = e.double(), de.double()
    dz_ref = (ded - ed * (ed * ded).sum(1, keepdim=True)) / nrm.double()[:, None]
    assert rel_err(dz16.double(), dz_ref) < BF16_ULP
    assert rel_err(dz_cp.view(nt, F).double().sum(0), dz16.double().sum(0)) < 1e-5
    dy = dz16.double() @ W2.double()
    if p > 0:
        dy = dy * mask.double() / (1 - p)
    assert rel_err(dyp.double(), dy) < BF16_ULP
    xhat = (h.double() - mean.double()) * rstd.double()
    s = _fold(bn_part, nt, H1)
    assert rel_err(s[0], dyp.double().sum(0)) < 2e-5
    assert rel_err(s[1], (dyp.double() * xhat).sum(0)) < 2e-5


@pytest.mark.parametrize("sync_form", [False, True])
@pytest.mark.parametrize("M,H1,LD,H0,F", [(8192, 64, 1152, 128, 64), (8192, 64, 320, 128, 64), (1000, 128, 448, 128, 128), (130, 64, 136, 128, 64)])
def test_tower_mid_backward(M, H1, LD, H0, F, sync_form):
    lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(M + LD)
    f32 = dict(dtype=torch.float32, device=DEV)
    h = torch.relu(torch.randn(M, H1, generator=g, device=DEV)).to(torch.bfloat16)
    dyp = (torch.randn(M, H1, generator=g, device=DEV) * 1e-3).to(torch.bfloat16)
    W1 = (torch.randn(H1, LD, generator=g, device=DEV) / LD ** 0.5).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(H1, generator=g, device=DEV)
    mean = h.float().mean(0)
    rstd = 1 / (h.float().var(0, unbiased=False) + 1e-5).sqrt()
    xhat = (h.double() - mean.double()) * rstd.double()
    nt = lib.query("tt_linear_stats_tiles", M)
    pad = lambda x: torch.cat([x, torch.zeros(nt * 128 - M, x.shape[1], dtype=x.dtype, device=DEV)])
    S0, S1 = dyp.double().sum(0), (dyp.double() * xhat).sum(0)
    bn_part = torch.stack([pad(dyp.float()).view(nt, 128, H1).sum(1), pad((dyp.double() * xhat).float()).view(nt, 128, H1).sum(1)], 1).contiguous().view(-1)
    sums = torch.cat([S0, S1]) if sync_form else None
    dz_cp = torch.randn(nt * F, generator=g, device=DEV)
    dgamma, dbeta, dbias2 = torch.empty(H1, **f32), torch.empty(H1, **f32), torch.empty(F, **f32)
    dh = torch.empty(M, H1, dtype=torch.bfloat16, device=DEV)
    dc = torch.full((M, LD), float("nan"), dtype=torch.bfloat16, device=DEV)
    dh_cp, dc_cp = torch.full((nt * H1,), float("nan"), **f32), torch.full((nt * H0,), float("nan"), **f32)
    lib.call("tt_tower_mid_backward", _ptr(dyp), _ptr(h), M, H1, LD, _ptr(W1), _ptr(bn_part), nt, _ptr(sums), M, _ptr(mean), _ptr(rstd),
             _ptr(gamma), _ptr(dgamma), _ptr(dbeta), 0.5, _ptr(dz_cp), F, _ptr(dbias2), _ptr(dh), _ptr(dh_cp), _ptr(dc), H0, _ptr(dc_cp), _s())
    assert rel_err(dbeta, 0.5 * S0) < 2e-5 and rel_err(dgamma, 0.5 * S1) < 2e-5
    assert rel_err(dbias2, dz_cp.view(nt, F).double().sum(0)) < 1e-5
    dh_ref = (rstd.double() * gamma.double()) * (dyp.double() - S0 / M - xhat * S1 / M) * (h.double() > 0)
    assert rel_err(dh.double(), dh_ref) < 2 * BF16_ULP
    assert rel_err(dh_cp.view(nt, H1).double().sum(0), dh.double().sum(0)) < 1e-5
    dc_ref = dh.double() @ W1.double()
    assert not torch.isnan(dc.float()).any()
    assert rel_err(dc.double(), dc_ref) < BF16_ULP
    assert rel_err(dc_cp.view(nt, H0).double().sum(0), dc[:, :H0].double().sum(0)) < 1e-5


def test_fused_and_per_stage_tower_paths_agree_on_a_train_step(monkeypatch):
    """Same model, same batch, dropout 0.1 (same Philox stream): the fused kernels and the per-stage kernels are two
    implementations of the same step — loss equal to bf16 noise, updated weights within the distance Adam's first steps allow."""
    import jodalrob_twotower_b200 as tt
    from oracle import bench_parity
    from tests.helpers import readme_schema
    from tests.test_gpu_bench_path import _batch, _task
    schema = readme_schema()
    batches = bench_parity.global_batches(schema, 2048, 2, seed=11)
    outs = []
    for fused in ("1", "0"):
        monkeypatch.setenv("TT_FUSED_TOWER", fused)
        torch.manual_seed(4)
        task = _task(tt, schema, "bf16", dropout=0.1)
        assert task.engine.fused_tower == (fused == "1")
        tt.FusedAdam(task, lr=1e-3, weight_decay=1e-5, warmup_steps=0)
        task.train()
        losses = [task.fused_step(_batch(tt, schema, b)).clone() for b in batches]
        outs.append((torch.stack(losses), task.engine.flat_w.clone(), [m.clone() for m in task.engine.bn_rm[0]]))
    (la, wa, ra), (lb, wb, rb) = outs
    assert rel_err(la[:, 0], lb[:, 0]) < 2e-3
    assert rel_err(ra[0], rb[0]) < 5e-3  # running mean after two steps (the second step's weights already differ by bf16 noise)
    assert float((wa - wb).abs().max()) <= 2 * 2e-3 + 1e-6  # two steps of lr 1e-3: sign flips of near-zero gradients at most
    assert float((wa - wb).pow(2).sum().sqrt() / wb.pow(2).sum().sqrt()) < 2e-3
