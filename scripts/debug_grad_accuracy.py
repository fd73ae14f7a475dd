"""Gradient accuracy of the fp32 CUDA path vs the fp64 explicit-math oracle, next to the fp32 torch port's own error."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jodalrob_twotower_b200 as tt
from oracle import ref_port, two_tower_oracle as orc
from tests.helpers import readme_schema
from tests.test_gpu_parity import make_task, make_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
schema = readme_schema()
cfg = dict(emb_dim=32, notice_dense=256, company_dense=128, hidden=[128, 64], final_dim=64, temperature=1.0)
torch.manual_seed(11)
task = make_task(cfg, schema); port = ref_port.build_port_task(schema, seed=11)
g = torch.Generator().manual_seed(99)
mk = lambda vocab: torch.stack([torch.randint(-1, v + 2, (B,), generator=g) for _, v in vocab], 1)
nid, cid = mk(schema["notice"]), mk(schema["company"]); nd, cd = torch.randn(B, 256, generator=g), torch.randn(B, 128, generator=g)
task.train(); port.train()
o = task(make_batch(schema, nid, cid, nd, cd), return_metrics=True); o["loss"].backward()
po = port(ref_port.make_port_batch(schema, nid, cid, nd, cd), return_metrics=True); po["loss"].backward()
truth = orc.OracleStepper(port.state_dict(), schema, 1.0, 0.0, 1e-3, 1e-5, 0, dtype=torch.float64)
loss, met, cache, grads, embs, _ = truth.forward_backward(nid.reshape(-1).numpy(), cid.reshape(-1).numpy(), nd, cd)
from tests.test_oracle_vs_golden import st_export_grads
tg = st_export_grads(truth, grads)
print("loss ours/port/truth", o["loss"].item(), po["loss"].item(), float(loss))
for (n1, p1), (n2, p2) in zip(task.named_parameters(), port.named_parameters()):
    if "embeddings" in n1 and p1.numel() > 5000: continue
    t = tg[n1]; sc = t.abs().max()
    eo = (p1.grad.double().cpu() - t).abs().max() / sc; ep = (p2.grad.double() - t).abs().max() / sc
    print(f"{n1[16:]:55s} |g|max {sc:.2e}  ours {eo:.2e}  port {ep:.2e}")
