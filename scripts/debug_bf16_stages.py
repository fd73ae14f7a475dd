"""Stage-by-stage comparison of the bf16 (tcgen05) engine against the fp32 engine on identical inputs."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jodalrob_twotower_b200 as tt
from tests.helpers import readme_schema, rel_err
schema = readme_schema(); nk=[k for k,_ in schema["notice"]]; ck=[k for k,_ in schema["company"]]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
mk = lambda dt: tt.create_two_tower_train_task(nk, ck, categorical_embedding_dim=32, notice_dense_input_dim=256, company_dense_input_dim=128,
        tower_hidden_dims=[128,64], final_embedding_dim=64, dropout_rate=0.0, device="cuda:0",
        vocab_sizes={"notice": dict(schema["notice"]), "company": dict(schema["company"])}, dtype=dt)
torch.manual_seed(5); a = mk("fp32"); b = mk("bf16"); b.load_state_dict(a.state_dict())
g = torch.Generator().manual_seed(17)
nid = torch.stack([torch.randint(0, v, (B,), generator=g) for _, v in schema["notice"]], 1).cuda()
cid = torch.stack([torch.randint(0, v, (B,), generator=g) for _, v in schema["company"]], 1).cuda()
nd, cd = torch.randn(B,256,generator=g).cuda(), torch.randn(B,128,generator=g).cuda()
outs = {}
for name, t in (("fp32", a), ("bf16", b)):
    e = t.engine
    o, ws = e.forward(nd, nid.reshape(-1), cd, cid.reshape(-1), True)
    e.backward(ws)
    outs[name] = (o.clone(), ws, e)
wa, wb = outs["fp32"][1], outs["bf16"][1]
print("out5", outs["fp32"][0].tolist(), outs["bf16"][0].tolist())
for t in range(2):
    for nm in ("c","z","e","de","dz","dc"):
        x, y = getattr(wa,nm)[t].float(), getattr(wb,nm)[t].float()
        print(t, nm, f"{rel_err(y, x):.3e}", f"rms-rel {((y-x).pow(2).mean().sqrt()/x.pow(2).mean().sqrt()).item():.3e}")
    for nm in ("h","y","dy","dh"):
        x, y = getattr(wa,nm)[t][0].float(), getattr(wb,nm)[t][0].float()
        print(t, nm, f"{rel_err(y, x):.3e}", f"rms-rel {((y-x).pow(2).mean().sqrt()/x.pow(2).mean().sqrt()).item():.3e}")
ga, gb = outs["fp32"][2].dense_g, outs["bf16"][2].dense_g
print("dense_g", f"{rel_err(gb, ga):.3e}")
