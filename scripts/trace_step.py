"""Kernel timeline of ONE graph-replayed train step (CUPTI through torch.profiler; no nsys in this image).

Prints every kernel of the replay with its start offset, duration and stream, plus the idle gaps of the whole device, so the
critical path across the tower / sort / Adam streams can be read off. Usage: python scripts/trace_step.py [batch] [dtype]"""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import jodalrob_twotower_b200 as tt
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = torch.device("cuda", 0)
schema = bench.load_schema()
M = bench.MODEL
nk, ck = [k for k, _ in schema["notice"]], [k for k, _ in schema["company"]]
torch.manual_seed(0)
task = tt.create_two_tower_train_task(
    nk, ck, categorical_embedding_dim=M["emb_dim"], notice_dense_input_dim=M["notice_dense"], company_dense_input_dim=M["company_dense"],
    tower_hidden_dims=M["hidden"], final_embedding_dim=M["final_dim"], dropout_rate=M["dropout"], temperature=M["temperature"],
    device=dev, vocab_sizes={"notice": dict(schema["notice"]), "company": dict(schema["company"])}, dtype=dtype)
tt.FusedAdam(task, lr=M["lr"], weight_decay=M["weight_decay"], warmup_steps=0)
task.train()
gen = torch.Generator().manual_seed(1234)
nid, cid, nd, cd = bench.synth_batch(schema, B, gen)
batch = {"notice": {"dense": nd.to(dev), "kjt": tt.KeyedJaggedTensor(nk, nid.reshape(-1).to(dev))},
         "company": {"dense": cd.to(dev), "kjt": tt.KeyedJaggedTensor(ck, cid.reshape(-1).to(dev))}}
gstep = tt.GraphedFusedStep(task, batch)
for _ in range(5):
    gstep(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        gstep(batch)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "tt_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
# keep the last of the three replays: it starts with the staging copies in front of the last hyper_advance_kernel
first = [i for i, e in enumerate(ev) if "hyper_advance" in e["name"]]
start = first[-1]
while start > 0 and ev[start - 1].get("cat") == "gpu_memcpy":
    start -= 1
ev = ev[start:]
t0 = ev[0]["ts"]
streams = {s: i for i, s in enumerate(sorted({e["args"].get("stream", 0) for e in ev}))}
end = max(e["ts"] + e["dur"] for e in ev)
print(f"step span {end - t0:.1f} us, {len(ev)} device activities, streams {len(streams)}")
busy_until = t0
for e in ev:
    gap = e["ts"] - busy_until
    mark = f"  <-- device idle {gap:.1f} us" if gap > 1.0 else ""
    print(f"{e['ts'] - t0:8.1f} +{e['dur']:7.1f}  s{streams[e['args'].get('stream', 0)]}  {e['name'][:90]}{mark}")
    busy_until = max(busy_until, e["ts"] + e["dur"])
