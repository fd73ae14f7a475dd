"""Stand-alone device time of the README-default embedding lookup (tt_ebc_forward_mixed: one id per bag, sample-major ids, bf16 rows
written into the tower's concat buffer) — a CUDA graph of N launches, CUDA events; `cold` writes a 512 MB buffer between launches.
Usage: python scripts/time_lookup.py [batch]   (TT_ONLY=notice|company restricts the run, for ncu -k regex:ebc_fwd_l1)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import jodalrob_twotower_b200 as tt
from jodalrob_twotower_b200.native import TT_BF16

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
schema = bench.load_schema(); M = bench.MODEL
nk, ck = [k for k, _ in schema["notice"]], [k for k, _ in schema["company"]]
torch.manual_seed(0)
task = tt.create_two_tower_train_task(
    nk, ck, categorical_embedding_dim=M["emb_dim"], notice_dense_input_dim=M["notice_dense"], company_dense_input_dim=M["company_dense"],
    tower_hidden_dims=M["hidden"], final_embedding_dim=M["final_dim"], dropout_rate=M["dropout"], temperature=M["temperature"],
    device=dev, vocab_sizes={"notice": dict(schema["notice"]), "company": dict(schema["company"])}, dtype="bf16")
eng = task.engine
lib = tt.native.lib()
gen = torch.Generator().manual_seed(1234)
nid, cid, _, _ = bench.synth_batch(schema, B, gen)
E, H0 = M["emb_dim"], M["hidden"][0]
P = lambda t: t.data_ptr()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
only = os.environ.get("TT_ONLY", "")
for t, (name, ids) in enumerate((("notice", nid), ("company", cid))):
    if only and only != name:
        continue
    K = eng.K[t]; k0 = 0 if t == 0 else eng.K[0]
    v = ids.reshape(-1).to(dev)
    c = torch.empty(B, H0 + K * E, dtype=torch.bfloat16, device=dev)
    rows = torch.empty(B * max(eng.KL[t], 1), dtype=torch.int32, device=dev)
    def launch(s):
        lib.call("tt_ebc_forward_mixed", P(eng.flat_w), eng.row_base.data_ptr() + 8 * k0, eng.vocab.data_ptr() + 4 * k0, P(v), B, K, E, 0,
                 P(c), TT_BF16, H0 + K * E, H0, eng.large_rank.data_ptr() + 4 * k0, eng.KL[t], eng.P_tot // E, P(rows), s)
    for cold in (False, True):
        n = 20
        for _ in range(3): launch(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for e0, e1 in evs:
            if cold: flush.fill_(1)
            e0.record(); launch(torch.cuda.current_stream().cuda_stream); e1.record()
        torch.cuda.synchronize()
        ts = sorted(e0.elapsed_time(e1) * 1000 for e0, e1 in evs)
        traffic = B * K * (8 + E * 4 + E * 2)
        print(f"{name} B={B} K={K} (large {eng.KL[t]}) {'cold' if cold else 'hot '}: median {ts[n // 2]:6.1f} us  min {ts[0]:6.1f}  "
              f"-> {traffic / ts[n // 2] / 1e3:7.1f} GB/s algorithmic ({traffic / 1e6:.1f} MB)", flush=True)
