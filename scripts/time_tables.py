"""Run the pooled-table kernels alone on a config-5-shaped shard (for ncu captures; sizes kept modest so that the profiler's
save/restore of written memory stays cheap): jagged lookup, sort-dedup, sparse segment-sum + Adam.
usage: [TT_MODE=sparse|dense_exact] python scripts/time_tables.py [rows_per_table=400000] [E=64] [L=20] [B=8192] [K=38]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jodalrob_twotower_b200.tables import PooledTables
a = [int(x) for x in sys.argv[1:]]
rows_per, E, L, B, K = (a + [400000, 64, 20, 8192, 38][len(a):])[:5]
dev = torch.device("cuda", 0)
tab = PooledTables([rows_per] * K, E, dev, optimizer_mode=os.environ.get("TT_MODE", "sparse"), seed=1)
g = torch.Generator(device=dev).manual_seed(7)
n = B * K * L
ids = (torch.rand(n, generator=g, device=dev) * rows_per).long().clamp_(max=rows_per - 1)
offs = torch.arange(B * K + 1, dtype=torch.int64, device=dev) * L
grad = torch.randn(B, K * E, generator=g, device=dev)
out = torch.empty(B, K * E, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(3):
    e = [ev() for _ in range(4)]
    e[0].record(); tab.forward(ids, B, offs, out=out); e[1].record(); tab.sort(); e[2].record()
    tab.hyper_advance(); tab.backward_adam(grad); e[3].record()
    torch.cuda.synchronize()
    print(f"iter {it}: lookup {e[0].elapsed_time(e[1]):.3f} ms  sort {e[1].elapsed_time(e[2]):.3f} ms  update {e[2].elapsed_time(e[3]):.3f} ms", flush=True)
uniq = int(torch.unique(tab.rows[:n]).numel())
print(f"lookup bytes {tab.forward_bytes(B, n)/1e9:.3f} GB, update bytes {tab.update_bytes(B, n, uniq)/1e9:.3f} GB, unique rows {uniq}")
