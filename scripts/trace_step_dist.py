"""Kernel timeline of ONE graph-replayed sharded train step on rank 0 (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/trace_step_dist.py"""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from torch.profiler import profile, ProfilerActivity

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import jodalrob_twotower_b200 as tt
from jodalrob_twotower_b200 import dist as ttdist
B = 8192
schema = bench.load_schema()
M = bench.MODEL
nk, ck = [k for k, _ in schema["notice"]], [k for k, _ in schema["company"]]
torch.manual_seed(0)
task = ttdist.create_distributed_train_task(nk, ck, schema, M, dev, dtype="bf16")
tt.FusedAdam(task, lr=M["lr"], weight_decay=M["weight_decay"], warmup_steps=0)
task.train()
gen = torch.Generator().manual_seed(1234 + rank)
nid, cid, nd, cd = bench.synth_batch(schema, B, gen)
batch = {"notice": {"dense": nd.to(dev), "kjt": tt.KeyedJaggedTensor(nk, nid.reshape(-1).to(dev))},
         "company": {"dense": cd.to(dev), "kjt": tt.KeyedJaggedTensor(ck, cid.reshape(-1).to(dev))}}
gstep = tt.GraphedFusedStep(task, batch)
packed = gstep.pack(batch)
for _ in range(5):
    gstep(packed=packed)
torch.cuda.synchronize()
dist.barrier(device_ids=[local])
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        gstep(packed=packed)
    torch.cuda.synchronize()
if rank == 0:
    path = os.path.join(tempfile.gettempdir(), "tt_trace_dist.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    copies = [i for i, e in enumerate(ev) if e.get("cat") == "gpu_memcpy"]
    start = copies[-1] if copies else 0  # the last replay starts at its staging copy
    ev = ev[start:]
    t0 = ev[0]["ts"]
    streams = {s: i for i, s in enumerate(sorted({e["args"].get("stream", 0) for e in ev}))}
    end = max(e["ts"] + e["dur"] for e in ev)
    print(f"step span {end - t0:.1f} us, {len(ev)} device activities, streams {len(streams)}")
    busy = t0
    for e in ev:
        gap = e["ts"] - busy
        mark = f"  <-- device idle {gap:.1f} us" if gap > 1.0 else ""
        print(f"{e['ts'] - t0:8.1f} +{e['dur']:7.1f}  s{streams[e['args'].get('stream', 0)]}  {e['name'][:80]}{mark}")
        busy = max(busy, e["ts"] + e["dur"])
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
