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


@pytest.mark.parametrize("E,vocabs,B,layout,bf16", [
    (32, [12, 25, 48, 49, 137], 1000, 0, False), (32, [12, 64, 100, 160], 8192, 0, True), (32, [5, 50], 300, 1, False),
    (64, [3, 24, 25, 80], 2049, 0, True), (128, [7, 12, 13, 40], 513, 1, False)])
def test_small_table_one_hot_reduction(E, vocabs, B, layout, bf16):
    """tt_ebc_small_backward (the dense one-hot gradient of the replicated small tables, two_tower_train_task.py's
    EmbeddingBagCollection backward for tables with a handful of rows) against torch index_add in fp64, at every chunking the
    single-wave grid picks (B from 300 to 8192). Ids include out-of-range values (clamped like the lookup clamps them).
    Deterministic: two runs are bit-identical."""
    import jodalrob_twotower_b200 as tt
    from jodalrob_twotower_b200.native import ptr, TT_BF16, TT_F32
    lib = tt.native.lib()
    g = torch.Generator().manual_seed(B + E)
    K = len(vocabs) + 2  # two keys of the tower are not small tables: their columns must be ignored
    key_of = [0, 2, 3, 4, 5, 6, 7][:len(vocabs)]
    H0 = 16
    ld = H0 + K * E
    ids = torch.stack([torch.randint(-2, max(vocabs) + 3, (B,), generator=g) for _ in range(K)], 1)  # [B, K]
    for t, k in enumerate(key_of):
        ids[:, k] = torch.randint(-1, vocabs[t] + 2, (B,), generator=g)
        ids[: B // 3, k] = t % vocabs[t]  # a hot row
    values = (ids.reshape(-1) if layout == 0 else ids.t().reshape(-1)).contiguous().to(DEV)
    grad = torch.randn(B, ld, generator=g)
    gdev = (grad.to(torch.bfloat16) if bf16 else grad).to(DEV).contiguous()
    gref = gdev.double().cpu()
    desc = np.zeros(len(vocabs), dtype=np.dtype([("key", "<i4"), ("vocab", "<i4"), ("col", "<i4"), ("pad", "<i4"), ("g_off", "<i8")]))
    off = 0
    for t, k in enumerate(key_of):
        desc[t] = (k, vocabs[t], H0 + k * E, 0, off)
        off += vocabs[t] * E
    small_floats = off
    tabs = torch.from_numpy(desc.view(np.uint8).copy()).to(DEV)
    nb = lib.query("tt_ebc_small_workspace_bytes", B, small_floats)
    s = torch.cuda.current_stream().cuda_stream
    outs = []
    for _ in range(2):
        ws = torch.full((nb,), 0x7F, dtype=torch.uint8, device=DEV)
        gout = torch.full((small_floats,), float("nan"), dtype=torch.float32, device=DEV)
        lib.call("tt_ebc_small_backward", ptr(values), B, K, layout, ptr(gdev), ld, TT_BF16 if bf16 else TT_F32, E, ptr(tabs), len(vocabs),
                 max(vocabs), small_floats, ptr(gout), ptr(ws), nb, s)
        torch.cuda.synchronize()
        outs.append(gout.cpu())
    assert torch.equal(outs[0], outs[1])
    off = 0
    for t, k in enumerate(key_of):
        V = vocabs[t]
        want = torch.zeros(V, E, dtype=torch.float64).index_add_(0, ids[:, k].clamp(0, V - 1), gref[:, H0 + k * E: H0 + (k + 1) * E])
        got = outs[0][off: off + V * E].view(V, E).double()
        assert rel_err(got, want) < 2e-6, (t, V)
        off += V * E
