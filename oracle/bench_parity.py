"""TEST INFRASTRUCTURE ONLY — parity check of the BENCHMARKED configuration (checker, never the thing measured).

Runs a short golden-style trajectory (fixed seed, dropout 0, the reference's LambdaLR warm-up) through exactly the path
``bench.py`` times — ``GraphedFusedStep`` over the fused step, ``dtype="bf16"`` (tcgen05 towers + sim-CE), and at
world > 1 the row-wise sharded engine with its NVLink peer-memory exchanges — and compares it with ``oracle/ref_port.py``
(pinned bit-exactly to the unmodified reference) stepping the CONCATENATED GLOBAL batch on the CPU
(reference: src/towers/two_tower_train_task.py:40-97, scripts/train.py:319-330).

Used by ``tests/test_gpu_bench_path.py`` (-m gpu), ``tests/dist_check.py`` (torchrun) and ``bench.py``'s ``"parity"`` key.
Tolerances are BASELINE.json's: 2e-2 relative in bf16, 1e-5 in fp32, written below.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ref_port

MODEL = dict(emb_dim=32, notice_dense=256, company_dense=128, hidden=[128, 64], final_dim=64, temperature=1.0)


def build_port(schema, seed: int = 0):
    return ref_port.build_port_task(schema, MODEL["emb_dim"], MODEL["notice_dense"], MODEL["company_dense"], MODEL["hidden"],
                                    MODEL["final_dim"], 0.0, MODEL["temperature"], seed=seed)


def global_batches(schema, B_global: int, steps: int, seed: int = 4321):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        nid = torch.stack([torch.randint(0, v, (B_global,), generator=g) for _, v in schema["notice"]], 1)
        cid = torch.stack([torch.randint(0, v, (B_global,), generator=g) for _, v in schema["company"]], 1)
        out.append((nid, cid, torch.randn(B_global, MODEL["notice_dense"], generator=g),
                    torch.randn(B_global, MODEL["company_dense"], generator=g)))
    return out


def adam_travel(lr: float, warmup_steps: int, steps: int) -> float:
    """Upper bound of how far Adam can move one element over the trajectory: sum_t lr_t * |m_hat / sqrt(v_hat)|, and the ratio is
    <= 1/sqrt(1-beta2)... in practice <= ~1 once both moments hold the same history; 1.0 per step is the bound used."""
    return sum(lr * (min(s / warmup_steps, 1.0) if warmup_steps > 0 else 1.0) for s in range(steps))


def compare_state(ours: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor], tol: float, travel: float) -> dict:
    """ours / ref: reference-named state_dicts. Returns the numbers and `ok`.
    * every embedding table, per tensor: max |a-b| <= tol * max|b|                                  (north_star: updated weights)
    * every dense tensor, per tensor (BN beta included): max |a-b| <= tol * max|b| + 2 * travel     (Adam moves an element by at
      most `travel`; where reduced precision flips the sign of a near-zero gradient the two trajectories part by 2 * travel,
      which for a zero-initialised BatchNorm beta IS its whole magnitude — so the bound is stated against the step size)
    * all weights together, rms: ||a-b|| / ||b|| <= tol
    """
    num = den = 0.0
    worst_table = worst_dense = 0.0
    bad = []
    for k, b in ref.items():
        if not b.is_floating_point() or k.endswith("num_batches_tracked"):
            continue
        a = ours[k].detach().double().cpu()
        b = b.detach().double().cpu()
        diff = float((a - b).abs().max())
        scale = float(b.abs().max()) + 1e-30
        num += float((a - b).pow(2).sum())
        den += float(b.pow(2).sum())
        if ".embeddings." in k:
            worst_table = max(worst_table, diff / scale)
            # fp32: Adam's update is ill-conditioned where |g| ~ eps (oracle/compare.py) — single elements may sit up to a step apart
            if diff > tol * scale + (2 * travel if tol < 1e-3 else 0.0):
                bad.append(k)
        else:
            worst_dense = max(worst_dense, max(diff - 2 * travel, 0.0) / scale)
            if diff > tol * scale + 2 * travel:
                bad.append(k)
    rms = (num / max(den, 1e-30)) ** 0.5
    return {"weights_rel": rms, "tables_rel_max": worst_table, "dense_rel_beyond_step": worst_dense, "bad_tensors": bad[:4],
            "ok": (not bad) and rms <= tol}


def run(tt, schema, device, dtype: str = "bf16", B_local: int = 1024, steps: int = 3, world: int = 1, rank: int = 0,
        graph: bool = True, lr: float = 1e-3, weight_decay: float = 1e-5, warmup_steps: int = 2, seed: int = 0,
        ttdist=None, comm=None, return_task: bool = False, port=None, optimizer_mode: str = "dense_exact") -> Optional[dict]:
    """Every rank calls it (collective at world > 1); the dict comes back on rank 0, None elsewhere. `port`: a freshly built
    port shared by in-process rank threads (the global torch RNG is not thread-safe); rank 0 steps it at the end."""
    tol = 2e-2 if dtype == "bf16" else 1e-5
    nk, ck = [k for k, _ in schema["notice"]], [k for k, _ in schema["company"]]
    if port is None:
        port = build_port(schema, seed)
    init = {k: v.clone() for k, v in port.state_dict().items()}
    if world > 1:
        model = dict(MODEL, dropout=0.0)
        task = ttdist.create_distributed_train_task(nk, ck, schema, model, device, comm=comm, dtype=dtype, optimizer_mode=optimizer_mode)
        ttdist.load_reference_state_dict(task, init)
    else:
        task = tt.create_two_tower_train_task(
            nk, ck, categorical_embedding_dim=MODEL["emb_dim"], notice_dense_input_dim=MODEL["notice_dense"],
            company_dense_input_dim=MODEL["company_dense"], tower_hidden_dims=MODEL["hidden"], final_embedding_dim=MODEL["final_dim"],
            dropout_rate=0.0, temperature=MODEL["temperature"], device=device,
            vocab_sizes={"notice": dict(schema["notice"]), "company": dict(schema["company"])}, dtype=dtype,
            optimizer_mode=optimizer_mode)
        task.load_state_dict(init)
    tt.FusedAdam(task, lr=lr, weight_decay=weight_decay, warmup_steps=warmup_steps)
    task.train()
    Bg = B_local * world
    batches = global_batches(schema, Bg, steps)
    sl = slice(rank * B_local, (rank + 1) * B_local)

    def mine(b):
        nid, cid, nd, cd = b
        mk = lambda keys, ids: tt.KeyedJaggedTensor(keys, ids[sl].reshape(-1).contiguous().to(device))
        return {"notice": {"dense": nd[sl].to(device), "kjt": mk(nk, nid)}, "company": {"dense": cd[sl].to(device), "kjt": mk(ck, cid)}}

    stepper = tt.GraphedFusedStep(task, mine(batches[0])) if graph else None
    losses = []
    for b in batches:
        out = stepper(mine(b)) if graph else task.fused_step(mine(b))
        losses.append(float(out[0].item()))
    sd = ttdist.gather_reference_state_dict(task) if world > 1 else {k: v.detach().clone() for k, v in task.state_dict().items()}
    if rank != 0:
        return (None, task) if return_task else None
    tr = ref_port.PortTrainer(port, lr, weight_decay, warmup_steps)
    ref_losses = [float(tr.step(ref_port.make_port_batch(schema, *b))["loss"].detach()) for b in batches]
    loss_rel = max(abs(a - r) / abs(r) for a, r in zip(losses, ref_losses))
    res = compare_state(sd, port.state_dict(), tol, adam_travel(lr, warmup_steps, steps))
    res.update({"loss_rel": loss_rel, "losses": losses, "ref_losses": ref_losses, "tol": tol, "steps": steps,
                "batch_per_gpu": B_local, "global_batch": Bg, "dtype": dtype,
                "path": ("GraphedFusedStep" if graph else "fused_step (eager)") + (f", row-wise sharded over {world} ranks" if world > 1 else ""),
                "oracle": "oracle/ref_port.py (bit-exact to the reference on tests/golden) on the concatenated global batch, CPU fp32"})
    res["ok"] = bool(res["ok"] and loss_rel <= tol)
    return (res, task) if return_task else res
