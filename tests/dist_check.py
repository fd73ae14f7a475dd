"""Real multi-GPU check (run under torchrun on N GPUs of one node):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py
Each rank takes its slice of the golden global batch; NCCL + peer-mapped (symmetric memory) buffers carry the exchange.
The result must equal the single-device reference trajectory (tests/golden/tiny_*.npz)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import batch_inputs, load_golden, rel_err, sd_from  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import jodalrob_twotower_b200 as tt
    from jodalrob_twotower_b200 import dist as ttd
    ok = True
    for case in ("base", "oob_temp", "deep_ls"):
        d, cfg = load_golden(f"tiny_{case}.npz")
        if cfg["bsz"] % world:
            continue
        schema = cfg["schema"]
        model = dict(emb_dim=cfg["emb_dim"], notice_dense=cfg["notice_dense"], company_dense=cfg["company_dense"], hidden=cfg["hidden"],
                     final_dim=cfg["final_dim"], dropout=0.0, temperature=cfg["temperature"], label_smoothing=cfg["label_smoothing"])
        nk, ck = [k for k, _ in schema["notice"]], [k for k, _ in schema["company"]]
        task = ttd.create_distributed_train_task(nk, ck, schema, model, dev, dtype="fp32")
        ttd.load_reference_state_dict(task, sd_from(d, "init."))
        tt.FusedAdam(task, lr=cfg["lr"], weight_decay=cfg["weight_decay"], warmup_steps=cfg["warmup_steps"])
        task.train()
        Bl = cfg["bsz"] // world
        mk = lambda keys, ids: tt.KeyedJaggedTensor(keys, ids.reshape(-1).contiguous().to(dev))
        worst = 0.0
        for i in range(3):
            nid, cid, nd, cd = batch_inputs(d, i)
            sl = slice(rank * Bl, (rank + 1) * Bl)
            batch = {"notice": {"dense": nd[sl].to(dev), "kjt": mk(nk, nid[sl])}, "company": {"dense": cd[sl].to(dev), "kjt": mk(ck, cid[sl])}}
            out = task.fused_step(batch)
            worst = max(worst, abs(out[0].item() - float(d[f"s{i}.loss"])) / abs(float(d[f"s{i}.loss"])))
        sd = ttd.gather_reference_state_dict(task)
        fin = sd_from(d, "final.")
        werr = max(rel_err(v.cpu(), fin[k]) for k, v in sd.items() if v.is_floating_point())
        good = worst < 1e-5 and werr < 2e-5
        ok = ok and good
        if rank == 0:
            print(f"[dist_check] world={world} case={case}: loss rel err {worst:.2e}, weights rel err {werr:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
    # the BENCHMARKED multi-GPU path: bf16 (tcgen05), NVLink peer-memory exchanges, whole step as one CUDA graph per rank —
    # against the reference port stepping the concatenated global batch (oracle/bench_parity.py)
    from oracle import bench_parity
    from tests.helpers import readme_schema
    for dtype, bl in (("bf16", 1024), ("fp32", 256)):
        res = bench_parity.run(tt, readme_schema(), dev, dtype=dtype, B_local=bl, steps=3, world=world, rank=rank, graph=True,
                               ttdist=ttd)
        if rank == 0:
            ok = ok and res["ok"]
            print(f"[dist_check] world={world} README-default {dtype} graph+peer path, B_l={bl}: loss rel {res['loss_rel']:.2e}, "
                  f"weights rms rel {res['weights_rel']:.2e}, tables max rel {res['tables_rel_max']:.2e}, bad {res['bad_tensors']} "
                  f"(tol {res['tol']:g}) -> {'OK' if res['ok'] else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    ok = bool(flag.item())
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)  # captured graphs hold symmetric-memory handles: skip the destructors (as bench.py does)


if __name__ == "__main__":
    main()
