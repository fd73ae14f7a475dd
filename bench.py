#!/usr/bin/env python
"""Benchmark of the two-tower train step (BASELINE.json metric: train samples/sec, device-timed).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on the host cores

A "step" is one whole train step (hyper advance, both towers forward, sim-CE loss, backward, sort-dedup +
segment-sum + exact-dense Adam on all embedding rows, Adam on the dense weights) over one synthetic batch of
the configured schema (tests/golden/schema_readme_default.json: 32 notice + 6 company keys, vocab = count+10),
README-default model (emb 32, towers [128,64] -> 64), B = 8192 per GPU (BASELINE.json configs[1]).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train_samples_per_sec"
UNIT = "samples/s"
MODEL = dict(emb_dim=32, notice_dense=256, company_dense=128, hidden=[128, 64], final_dim=64, dropout=0.1, temperature=1.0,
             lr=1e-3, weight_decay=1e-5)
N_ROTATE = 16  # distinct resident input batches cycled through the timed region


def load_schema():
    with open(os.path.join(ROOT, "tests", "golden", "schema_readme_default.json")) as f:
        s = json.load(f)
    return {"notice": s["notice"], "company": s["company"]}


def synth_batch(schema, B, gen):
    nid = torch.stack([torch.randint(0, v, (B,), generator=gen) for _, v in schema["notice"]], 1)
    cid = torch.stack([torch.randint(0, v, (B,), generator=gen) for _, v in schema["company"]], 1)
    return nid, cid, torch.randn(B, MODEL["notice_dense"], generator=gen), torch.randn(B, MODEL["company_dense"], generator=gen)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def cupti_kernel_table(prof, n_steps, drop_first=1):
    """{kernel name: (device us per step, launches per step)} over the LAST n_steps of the n_steps + drop_first identical steps run
    inside `prof`. The first step(s) are dropped per kernel (its earliest launches): under torchrun the ranks leave the profiler's
    start-up at different times, and the first peer-memory kernel of the first profiled step then spins for the slowest rank —
    28 ms "per launch" in an N=2 table once."""
    per = {}
    for e in prof.events():
        if e.device_time_total > 0:
            per.setdefault(e.key, []).append((e.time_range.start, e.device_time_total))
    out = {}
    for key, lst in per.items():
        lst.sort()
        kept = lst[len(lst) * drop_first // (n_steps + drop_first):]
        out[key] = (sum(d for _, d in kept) / n_steps, len(kept) / n_steps)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
REF_COPY = os.path.join(ROOT, "oracle", "_ref")


class RefStepper:
    """One reference train step on the host cores: zero_grad -> forward(return_metrics=True) -> backward -> Adam.step ->
    scheduler.step (scripts/train.py:319-330). kind "reference": the UNMODIFIED reference modules from oracle/_ref (verbatim copy
    made by oracle/make_ref.py; incl. its per-scalar KJT decode loop cat_embed.py:103-121, its per-forward parameter loops
    base_tower.py:115-127 and the allclose sync two_tower_model.py:81-87), through oracle/ref_harness.py's two shims. kind "port":
    oracle/ref_port.py (bit-exact restatement) when the copy is absent. Never reads /root/reference."""

    def __init__(self, schema, decode="loop"):
        self.schema = schema
        self.kind = "port"
        have_ref = os.path.isfile(os.path.join(REF_COPY, "src", "towers", "two_tower_train_task.py"))
        if have_ref and decode == "loop":
            os.environ["TT_REFERENCE_ROOT"] = REF_COPY
            from oracle import ref_harness as rh
            self.rh = rh
            self.task, self.nk, self.ck = rh.build_reference_task(
                MODEL["emb_dim"], MODEL["notice_dense"], MODEL["company_dense"], MODEL["hidden"], MODEL["final_dim"],
                MODEL["dropout"], MODEL["temperature"], seed=0)
            got = {"notice": [[k, self.task.two_tower_model.notice_tower.categorical_embedder.vocab_sizes[k]] for k in self.nk],
                   "company": [[k, self.task.two_tower_model.company_tower.categorical_embedder.vocab_sizes[k]] for k in self.ck]}
            assert got == {"notice": [list(x) for x in schema["notice"]], "company": [list(x) for x in schema["company"]]}, \
                "oracle/_ref schema differs from tests/golden/schema_readme_default.json"
            self.task.train()
            self.opt = torch.optim.Adam(self.task.parameters(), lr=MODEL["lr"], weight_decay=MODEL["weight_decay"])
            self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lambda s_: 1.0)
            self.kind = "reference"
        else:
            from oracle import ref_port
            self.rp = ref_port
            task = ref_port.build_port_task(schema, MODEL["emb_dim"], MODEL["notice_dense"], MODEL["company_dense"], MODEL["hidden"],
                                            MODEL["final_dim"], MODEL["dropout"], MODEL["temperature"], seed=0, decode=decode)
            self.tr = ref_port.PortTrainer(task, MODEL["lr"], MODEL["weight_decay"], 0)

    def make_batch(self, nid, cid, nd, cd):
        if self.kind == "reference":
            return self.rh.make_reference_batch(self.nk, self.ck, nid, cid, nd, cd)
        return self.rp.make_port_batch(self.schema, nid, cid, nd, cd)

    def step(self, batch) -> float:
        if self.kind == "reference":
            self.opt.zero_grad()
            with self.rh.silenced():
                out = self.task(batch, return_metrics=True)
            out["loss"].backward()
            self.opt.step()
            self.sched.step()
            return float(out["loss"].detach())
        return float(self.tr.step(batch)["loss"].detach())


def run_reference(args):
    """The reference's own CPU implementation of the path at the SAME configuration (README-default, B = --batch), all host
    threads. Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    schema = load_schema()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.batch
    ref = RefStepper(schema)
    gen = torch.Generator().manual_seed(1234)
    batches = [ref.make_batch(*synth_batch(schema, B, gen)) for _ in range(2)]
    # bounded run: one probe step decides how many warm-ups fit (a B=8192 step is ~2 s on 16 cores); B is only reduced if even
    # K timed steps + 1 warm-up would exceed ~6 minutes
    t0 = time.perf_counter()
    ref.step(batches[0])
    probe = time.perf_counter() - t0
    same = True
    while B > 256 and probe * (args.steps + 1) > 360.0:
        B //= 2
        same = False
        batches = [ref.make_batch(*synth_batch(schema, B, gen)) for _ in range(2)]
        t0 = time.perf_counter()
        ref.step(batches[0])
        probe = time.perf_counter() - t0
    warm = max(0, min(args.warmup - 1, int(120.0 / max(probe, 1e-3))))
    for i in range(warm):
        ref.step(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        ref.step(batches[i % 2])
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    what = ("unmodified reference modules (oracle/_ref), as-is python KJT decode loop" if ref.kind == "reference"
            else "oracle/ref_port.py (bit-exact port), as-is python KJT decode loop")
    sample = f"{args.steps} timed steps of B={B} after {warm + 1} warm-up steps, {what}, {cores} torch threads"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm + 1, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"README-default two-tower train step (emb 32, towers [128,64]->64, 38 tables / 64,141 rows), "
                                   f"B={B}, in-batch negatives, dropout 0.1, dense Adam lr 1e-3 wd 1e-5, fp32, CPU",
                       "batch_per_gpu": B, "same_config_as_gpu_arm": same and B == args.batch},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(schema, B, budget_s=25.0):
    """The reference timed on the host cores, bounded sample of the same workload (reported baseline, not the target)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {}
    kind = "port"
    for decode in ("loop", "vectorized"):
        ref = RefStepper(schema, decode=decode)
        if decode == "loop":
            kind = ref.kind
        gen = torch.Generator().manual_seed(1234)
        batch = ref.make_batch(*synth_batch(schema, B, gen))
        t0 = time.perf_counter()
        ref.step(batch)  # warm-up
        warm = time.perf_counter() - t0
        n = max(1, min(10, int((budget_s / 2) / max(warm, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(n):
            ref.step(batch)
        dt = time.perf_counter() - t0
        out[decode] = (B * n / dt, n)
    what = "unmodified reference modules (oracle/_ref)" if kind == "reference" else "oracle/ref_port.py (bit-exact port)"
    return {"value": out["loop"][0], "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{out['loop'][1]} timed steps of B={B} after 1 warm-up, {what} as-is (python KJT decode loop), {cores} torch threads",
            "vectorized_decode_value": out["vectorized"][0],
            "vectorized_decode_note": "fairness footnote: the port with the KJT decode vectorised (not the reference's code)"}



# ------------------------------------------------------------------------------------------------ large-table section
def tables_section(dev, pk, quick=False):
    """BASELINE.json's second metric ("lookup HBM GB/s vs peak") on the shapes where the embedding kernels are HBM-bound:
    one GPU's row-wise shard of config 5 (500 M rows x E=64 over 8 GPUs = 62.5 M rows, 16 GB fp32 + 32 GB Adam state; multi-hot
    bags, pooling factor 20) and of config 3 (38 tables x 10 M rows x E=32 over 8 GPUs, one id per bag). Uniform ids (no L2 reuse:
    the worst case; tables >> 126 MB L2), B = 8192 samples per step, 38 keys. Algorithmic bytes: SURVEY.md §8(d)."""
    from jodalrob_twotower_b200.tables import PooledTables
    out = {}
    K, B = 38, 8192
    cases = [("config5_shard_E64_L20", 62_500_000 // K, 64, 20), ("config3_shard_E32_L1", 1_250_000, 32, 1)]
    for name, rows_per, E, L in cases:
        res = {}
        for mode in ("sparse", "dense_exact"):
            if mode == "dense_exact" and quick:
                continue
            torch.cuda.empty_cache()
            tab = PooledTables([rows_per] * K, E, dev, optimizer_mode=mode, seed=1, init_std=0.0 if mode == "dense_exact" else 1.0)
            g = torch.Generator(device=dev).manual_seed(7)
            n = B * K * L
            key = (torch.arange(n, device=dev) // L) % K
            sets = []
            for _ in range(3):
                u = torch.rand(n, generator=g, device=dev)
                sets.append(((u * rows_per).long().clamp_(max=rows_per - 1), torch.randn(B, K * E, generator=g, device=dev)))
            offs = torch.arange(B * K + 1, dtype=torch.int64, device=dev) * L if L > 1 else None
            outbuf = torch.empty(B, K * E, dtype=torch.float32, device=dev)
            st = torch.cuda.current_stream(dev)

            def timed(fn, reps):
                evs = []
                for i in range(reps + 2):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(st); fn(i); b.record(st)
                    evs.append((a, b))
                torch.cuda.synchronize(dev)
                t = sorted(a.elapsed_time(b) for a, b in evs[2:])
                return t[len(t) // 2] * 1e-3

            reps = 3 if mode == "dense_exact" else 10
            if mode == "sparse":
                t_f = timed(lambda i: tab.forward(sets[i % 3][0], B, offs, out=outbuf), reps)
                fb = tab.forward_bytes(B, n, 4, jagged=L > 1)
                res["lookup"] = {"ms": t_f * 1e3, "bytes": fb, "GBps": fb / t_f / 1e9, "frac_hbm": fb / t_f / 1e9 / pk["hbm_gbs"]}
                t_s = timed(lambda i: tab.sort(), reps)
                res["sort_dedup"] = {"ms": t_s * 1e3, "pairs": n, "Mpairs_per_s": n / t_s / 1e6}
            tab.forward(sets[0][0], B, offs, out=outbuf)
            tab.sort()
            uniq = int(torch.unique(tab.rows[:n]).numel())

            def upd(i):
                tab.hyper_advance()
                tab.backward_adam(sets[0][1])
            t_u = timed(upd, reps)
            ub = tab.update_bytes(B, n, uniq, 4)
            res["update_" + mode] = {"ms": t_u * 1e3, "bytes": ub, "unique_rows": uniq, "GBps": ub / t_u / 1e9,
                                     "frac_hbm": ub / t_u / 1e9 / pk["hbm_gbs"]}
            del tab, sets
        res["shape"] = {"tables": K, "rows_per_table": rows_per, "emb_dim": E, "ids_per_bag": L, "batch": B, "ids": "uniform"}
        out[name] = res
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ BASELINE configs[2]: 10 M-row tables
def zipf_ids(n, V, alpha, gen, dev):
    """Zipf(alpha)-distributed ids over [0, V) (continuous inverse-CDF approximation), scrambled by a fixed bijection of [0, V) so
    that the hot rows spread over the row-wise shards the way hashed ids do in practice (rank r owns the contiguous block r)."""
    u = torch.rand(n, generator=gen, device=dev, dtype=torch.float64)
    a = 1.0 - alpha
    x = ((float(V) ** a - 1.0) * u + 1.0) ** (1.0 / a)
    ids = (x.floor().long() - 1).clamp_(0, V - 1)
    return (ids * 2_654_435_761 + 12_345) % V if V % 2_654_435_761 else ids


def run_cfg3(args):
    """BASELINE.json configs[2]: the README-default towers on 38 tables of --rows-per-table rows (10 M), row-wise sharded over the
    ranks, Zipf(1.05) ids, B = --batch per GPU, all-to-all id routing, fused lookup + row push and gradient pull over NVLink peer
    memory. One JSON line: samples/s, the kernel table of the graph replays (CUPTI) and the HBM fractions of the table kernels
    against their algorithmic bytes, NVLink bytes per rank against the formulas."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    dist.init_process_group("nccl", device_id=dev)
    import jodalrob_twotower_b200 as tt
    from jodalrob_twotower_b200 import dist as ttdist
    base = load_schema()
    cfg5 = args.config == "cfg5"
    L = 20 if cfg5 else 1
    model = dict(MODEL, emb_dim=64 if cfg5 else MODEL["emb_dim"])
    V = int(args.rows_per_table) if args.rows_per_table else (500_000_000 // 38 if cfg5 else 10_000_000)
    schema = {"notice": [[k, V] for k, _ in base["notice"]], "company": [[k, V] for k, _ in base["company"]]}
    nk, ck = [k for k, _ in schema["notice"]], [k for k, _ in schema["company"]]
    B, E, K = args.batch, model["emb_dim"], len(nk) + len(ck)
    torch.manual_seed(0)
    task = ttdist.create_distributed_train_task(nk, ck, schema, model, dev, dtype=args.dtype, optimizer_mode=args.optimizer,
                                                seed_full_init=False, bag_len=L)
    tt.FusedAdam(task, lr=MODEL["lr"], weight_decay=MODEL["weight_decay"], warmup_steps=0)
    task.train()
    eng = task.engine
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    n_rot = 4
    mk = lambda keys, v: tt.KeyedJaggedTensor(keys, v)

    def batch():
        nid = torch.stack([zipf_ids(B * L, V, args.zipf, gen, dev).view(B, L) for _ in nk], 1).reshape(-1)  # [B, K, L] sample-major
        cid = torch.stack([zipf_ids(B * L, V, args.zipf, gen, dev).view(B, L) for _ in ck], 1).reshape(-1)
        return {"notice": {"dense": torch.randn(B, MODEL["notice_dense"], generator=gen, device=dev), "kjt": mk(nk, nid)},
                "company": {"dense": torch.randn(B, MODEL["company_dense"], generator=gen, device=dev), "kjt": mk(ck, cid)}}

    batches = [batch() for _ in range(n_rot)]
    gstep = tt.GraphedFusedStep(task, batches[0])
    packed = [gstep.pack(b) for b in batches]

    def barrier():
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        gstep(packed=packed[i % n_rot])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream(dev)
    barrier()
    e0.record(st)
    for i in range(args.steps):
        out = gstep(packed=packed[i % n_rot])
    e1.record(st)
    barrier()
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    clocks = sampler.stop() if rank == 0 else None
    loss_last = float(out[0].item())
    # kernel table of the replays (CUPTI)
    gk = {}
    from torch.profiler import ProfilerActivity, profile
    n_cupti = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(n_cupti + 1):
            gstep(packed=packed[i % n_rot])
        torch.cuda.synchronize(dev)
    for key, (us, cnt) in cupti_kernel_table(prof, n_cupti).items():
        name = key.replace("void ", "").replace("tt::", "").split("(")[0]
        d = gk.setdefault(name, {"us": 0.0, "n": 0.0})
        d["us"] += us
        d["n"] += cnt
    barrier()
    # lazy exact-replay Adam: what the deferred dense update costs when it is finally paid — a flush after `lag` more steps replays
    # (rows x E x lag) element-steps in registers; reported as picoseconds per element-step next to the 24 B / 6.5 TB/s = 3.7 ps the
    # streaming dense update needs at the HBM peak
    lazy_replay = None
    if eng.lazy:
        eng.flush_lazy()
        lag = 32
        for i in range(lag):
            gstep(packed=packed[i % n_rot])
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(st)
        eng.flush_lazy()
        f1.record(st)
        torch.cuda.synchronize(dev)
        fl_ms = f0.elapsed_time(f1)
        lazy_replay = {"flush_ms_after_32_steps": fl_ms, "rows": eng.total_rows, "element_steps_upper_bound": eng.total_rows * E * lag,
                       "ps_per_element_step": fl_ms * 1e-3 / (eng.total_rows * E * lag) * 1e12,
                       "note": "every row-step is replayed exactly once (at the row's next touch, or at a flush): in steady state the "
                               "deferred arithmetic per step is rows x E element-steps at this rate, instead of rows x E x 24 B of HBM traffic"}
        barrier()
    ws = eng.workspace(B)
    n_owned = int(ws.n_owned.item())
    uniq = int(torch.unique(ws.sorted_rows[:n_owned]).numel()) if n_owned else 0
    stats = torch.tensor([n_owned, uniq], dtype=torch.float64, device=dev)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    if rank != 0:
        _finish(world, dev)
        return
    pk = peaks()
    n_l = B * K * L
    n_owned_avg, uniq_avg = float(stats[0]) / world, float(stats[1]) / world
    w_o = 2 if args.dtype == "bf16" else 4
    rows_local = eng.total_rows
    alg = {
        "push_pooled_kernel": n_owned_avg * (E * 4 + 8) + min(n_owned_avg, world * B * K) * E * 4,  # rows read + (row, slot) + partials pushed
        "sum_partials_kernel": B * K * E * (4 * world + w_o),                          # W partials read, pooled bag written
        "push_rows_kernel": n_owned_avg * (E * 4 + 8 + E * w_o),                       # row read + (row, slot) + row pushed to the requester
        "ebc_gather_sorted_kernel": n_owned_avg * (E * w_o + 4 + E * 4),               # gradient pulled over NVLink + slot + compacted copy
        "ebc_rows_dense_stream_kernel": rows_local * E * 24 + n_owned_avg * E * 4,     # every row: w, m, v read + written
        "ebc_rows_sparse_kernel": uniq_avg * E * 24 + n_owned_avg * (E * 4 + 4),       # touched rows only
    }
    kernels = []
    for k_, v in sorted(gk.items(), key=lambda kv: -kv[1]["us"]):
        row = {"kernel": k_, "us_per_step": round(v["us"], 1), "launches_per_step": round(v["n"], 2)}
        for key, by in alg.items():
            if k_.startswith(key):
                row["algorithmic_MB"] = round(by / 1e6, 2)
                row["GBps"] = round(by / (v["us"] * 1e-6) / 1e9, 1)
                row["frac_hbm"] = round(row["GBps"] / pk["hbm_gbs"], 3)
        kernels.append(row)
    f = (world - 1) / world
    nvlink = {"id_all_to_all_sent": n_l * 8 * f,
              "rows_received": (n_l * E * w_o if L == 1 else min(n_l, world * B * K) * E * 4) * f,  # L > 1: one fp32 partial per (owner, bag)
              "gradients_pulled_by_owner": n_owned_avg * E * w_o * f,
              "embedding_all_gathers_received": 2 * B * world * MODEL["final_dim"] * 2 * f, "sum_exchanges_received": 2 * B * world * 4 * f,
              "dense_gradient_all_reduce": 2 * eng.P_tot * 4 * f}
    top_tab = next((r for r in kernels if "frac_hbm" in r), None)
    roofline = None
    if top_tab is not None:
        roofline = {"kernel": top_tab["kernel"], "bound": "hbm", "achieved": top_tab["GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": top_tab["frac_hbm"], "traffic": None, "peak_source": pk["source"],
                    "algorithmic": f"{top_tab['algorithmic_MB']} MB per launch (SURVEY.md 8(d))"}
    line = {"metric": METRIC, "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[{4 if cfg5 else 2}]: README-default towers on {K} tables x {V:,} rows x E={E} ({K * V * E * 4 / 1e9:.1f} GB of "
                                   f"fp32 weights, x3 with the Adam moments), row-wise sharded over {world} GPUs, " +
                                   (f"multi-hot bags of {L} ids (sum pooling, owners pre-pool), " if L > 1 else "") +
                                   f"Zipf({args.zipf}) ids (scrambled), B={B} per GPU, global negatives",
                       "ids_per_bag": L,
                       "batch_per_gpu": B, "global_batch": B * world, "rows_per_table": V, "rows_per_gpu": rows_local,
                       "optimizer": args.optimizer, "routing": "all-to-all over NVLink peer memory (sender buckets by owner)",
                       "l2_policy": f"tables of {rows_local * E * 4 / 1e9:.1f} GB per GPU >> 126 MB L2; inputs rotate over {n_rot} batches",
                       "cuda_graph": True},
            "clocks": clocks, "gpu_launches": gstep.launches_per_step * args.steps, "launches_per_step": gstep.launches_per_step,
            "roofline": roofline, "kernels": kernels[:20],
            "ids_per_rank": {"requested": n_l, "owned_avg": n_owned_avg, "unique_rows_avg": uniq_avg},
            "nvlink_bytes_per_rank_per_step": {k_: round(v) for k_, v in nvlink.items()},
            "nvlink_total_MB_per_rank_per_step": round(sum(nvlink.values()) / 1e6, 2), "lazy_replay": lazy_replay, "loss_last": loss_last}
    print(json.dumps(line), flush=True)
    _finish(world, dev)

# ------------------------------------------------------------------------------------------------ our arm
def algorithmic(tag: str, B: int, schema, world: int = 1) -> dict:
    """Algorithmic bytes / flops of one C-ABI call. B = samples per rank; with global
    negatives a rank's logits block is B x (world * B)."""
    Bg = B * world
    E, F, H0, H1 = MODEL["emb_dim"], MODEL["final_dim"], MODEL["hidden"][0], MODEL["hidden"][1]
    Kn, Kc = len(schema["notice"]), len(schema["company"])
    rows = sum(v for _, v in schema["notice"]) + sum(v for _, v in schema["company"])
    K = Kn if "notice" in tag else Kc
    D = MODEL["notice_dense"] if "notice" in tag else MODEL["company_dense"]
    if tag.startswith("tt_ebc_forward"):
        return {"bytes": B * K * (E * 4 + 8) + B * K * E * 4}
    if tag.startswith("tt_ebc_backward_adam"):
        return {"bytes": B * (Kn + Kc) * (E * 4 + 4) + rows * E * 24}
    if tag.startswith("tt_ebc_sort_rows"):
        return {"bytes": B * (Kn + Kc) * 4 * 2 * 3}
    if tag.startswith("tt_dense_adam"):
        return {"bytes": 152320 * 28}
    # sim-CE flops: a forward call needs S = N C^T once (2 B^2 F; its second, transposed pass is an implementation choice and is
    # not counted); a backward call forms one tower's gradient from logits that were never stored: S again (2 B^2 F) and
    # dX = dS Y (2 B^2 F); the widened [Y|Ys] operand is not counted. `exps` is the number of exp evaluations a call executes
    # (one per logit per pass), the quantity the MUFU pipe bounds.
    # three-pass step: P1 = S once (2 B Bg F); P2 / P3 = S again + dS Y (4 B Bg F each); one exp per logit per pass
    if tag.startswith("tt_simce3_rows") and "finish" not in tag or tag.startswith("simce_tc_kernel<0"):
        return {"flops": 2 * B * Bg * F, "exps": B * Bg}
    if tag.startswith("tt_simce3_grad") and "finish" not in tag or tag.startswith("simce_tc_kernel<1") or tag.startswith("simce_tc_kernel<2"):
        return {"flops": 4 * B * Bg * F, "exps": B * Bg}
    if tag.startswith("tt_simce_forward"):
        return {"flops": 2 * B * Bg * F, "exps": 2 * B * Bg}
    if tag.startswith("tt_simce_backward"):
        return {"flops": 4 * B * Bg * F, "exps": B * Bg}
    return {}


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} must be launched under torchrun with {args.gpus} ranks")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import jodalrob_twotower_b200 as tt

    schema = load_schema()
    nk, ck = [k for k, _ in schema["notice"]], [k for k, _ in schema["company"]]
    B = args.batch
    torch.manual_seed(0)
    if world > 1:
        from jodalrob_twotower_b200 import dist as ttdist
        task = ttdist.create_distributed_train_task(nk, ck, schema, MODEL, dev, dtype=args.dtype, optimizer_mode=args.optimizer)
    else:
        task = tt.create_two_tower_train_task(
            nk, ck, categorical_embedding_dim=MODEL["emb_dim"], notice_dense_input_dim=MODEL["notice_dense"],
            company_dense_input_dim=MODEL["company_dense"], tower_hidden_dims=MODEL["hidden"], final_embedding_dim=MODEL["final_dim"],
            dropout_rate=MODEL["dropout"], temperature=MODEL["temperature"], device=dev,
            vocab_sizes={"notice": dict(schema["notice"]), "company": dict(schema["company"])}, dtype=args.dtype,
            optimizer_mode=args.optimizer)
    # ---- parity of THIS path (graph replay, this dtype, this world size and its peer-memory exchanges) against the reference
    # port on the concatenated global batch, before anything is timed (checker only; oracle/bench_parity.py)
    parity = None
    if not args.no_parity:
        from oracle import bench_parity
        parity = bench_parity.run(tt, schema, dev, dtype=args.dtype, B_local=args.parity_batch, steps=3, world=world, rank=rank,
                                  graph=not args.no_graph, ttdist=ttdist if world > 1 else None)
        if parity is not None:
            parity.pop("losses", None); parity.pop("ref_losses", None)
    tt.FusedAdam(task, lr=MODEL["lr"], weight_decay=MODEL["weight_decay"], warmup_steps=0)
    task.train()
    eng = task.engine
    gen = torch.Generator().manual_seed(1234 + rank)
    host = [synth_batch(schema, B, gen) for _ in range(N_ROTATE)]
    pinned = [tuple(t.pin_memory() for t in (nid.reshape(-1), cid.reshape(-1), nd, cd)) for nid, cid, nd, cd in host]
    resident = [tuple(t.to(dev) for t in p) for p in pinned]
    in_bytes = sum(t.numel() * t.element_size() for t in pinned[0])
    if args.dtype == "bf16" and not args.fp32_inputs and not args.no_graph:
        in_bytes = sum(t.numel() * (2 if t.is_floating_point() else t.element_size()) for t in pinned[0])  # dense staged as bf16

    mk = lambda keys, v: tt.KeyedJaggedTensor(keys, v)

    def as_batch(tensors):
        nv, cv, nd, cd = tensors
        return {"notice": {"dense": nd, "kjt": mk(nk, nv)}, "company": {"dense": cd, "kjt": mk(ck, cv)}}

    def step_resident(i):  # eager (no graph): used by the per-kernel profile pass and --no-graph
        return task.fused_step(as_batch(resident[i % N_ROTATE]))

    # ---- the whole step as one CUDA graph through the public API (launch-bound at this size) ----
    use_graph = not args.no_graph
    # compact_inputs: a bf16 engine takes its dense features in bf16 (bit-identical to its own device-side cast of fp32 inputs):
    # half the PCIe bytes in the end-to-end loop and no cast kernels in the step
    compact = args.dtype == "bf16" and not args.fp32_inputs
    gstep = tt.GraphedFusedStep(task, as_batch(resident[0]), compact_inputs=compact) if use_graph else None
    if gstep is None:
        for i in range(2):
            step_resident(i)
        l0 = eng.launches
        step_resident(0)
        launches_per_step = eng.launches - l0
    else:
        launches_per_step = gstep.launches_per_step

    # batches packed in the layout of the graph's static input buffer: staging is ONE copy (device->device here, and
    # host->device in the end-to-end loop)
    packed_res = [gstep.pack(as_batch(r)) for r in resident] if gstep is not None else None
    packed_pin = [gstep.pack(as_batch(h), pin=True) for h in pinned] if gstep is not None else None

    def run_step(i):
        if gstep is not None:
            return gstep(packed=packed_res[i % N_ROTATE])  # device-to-device stage of a resident batch + graph replay
        return step_resident(i)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        run_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream(dev)
    barrier()
    e0.record(st)
    for i in range(args.steps):
        out_static = run_step(i)
    e1.record(st)
    barrier()
    ms = e0.elapsed_time(e1)
    loss_last = float(out_static[0].item())
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the public API: pinned host inputs -> H2D -> step -> D2H of the result, every step.
    # H2D of batch i+1 (copy stream, double-buffered staging) overlaps the step of batch i. ----
    copy_stream = torch.cuda.Stream(dev)
    # Two captures of the same step, each with its own static input buffer: the host-to-device copy of batch i+1 goes straight into
    # the buffer of the graph that will run it while the other graph runs step i — no device-to-device staging copy, and nothing
    # between the host's return from step i and the replay of step i+1 but one event wait (--e2e-single-graph: one capture, staged).
    # Single GPU only: measured 27.6 -> 28.6 M samples/s there; under torchrun the strict loop is paced by the ranks waiting for
    # each other in the first exchange of every step, and one sample at N=2 showed no gain (26.2 vs 28.0 M).
    gpair = None
    if gstep is not None and world == 1 and not args.e2e_single_graph:
        gpair = [gstep, tt.GraphedFusedStep(task, as_batch(resident[0]), compact_inputs=compact)]
    if gpair is not None:
        stage = [(g_.static_blob,) for g_ in gpair]
        src = [(b,) for b in packed_pin]
    elif gstep is not None:
        stage = [(torch.empty_like(packed_res[0]),) for _ in range(2)]
        src = [(b,) for b in packed_pin]
    else:
        stage = [tuple(torch.empty_like(t) for t in resident[0]) for _ in range(2)]
        src = pinned
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    out_host = torch.empty(5, dtype=torch.float32).pin_memory()
    out_hosts = [torch.empty(5, dtype=torch.float32).pin_memory() for _ in range(2)]
    read_done = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n, lag=0):
        """lag=0: the host reads step i's result before it launches step i+1 (the strict `e2e`). lag=1: the read of step i's result
        is enqueued right behind the step and the host waits for it only after launching step i+1 (asynchronous logging; reported
        as the supplementary `e2e_pipelined`)."""
        main = torch.cuda.current_stream(dev)
        with torch.cuda.stream(copy_stream):
            for t, s_ in zip(src[0], stage[0]):
                s_.copy_(t, non_blocking=True)
            ready[0].record(copy_stream)

        def prefetch(i):  # H2D of batch i + 1 into the other staging buffer, on the copy stream
            nxt = (i + 1) % 2
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(freed[nxt])
                    for t, s_ in zip(src[(i + 1) % N_ROTATE], stage[nxt]):
                        s_.copy_(t, non_blocking=True)
                    ready[nxt].record(copy_stream)

        for i in range(n):
            cur, nxt = i % 2, (i + 1) % 2
            # step i is launched FIRST (the GPU is idle while the host gets here), the copy of batch i + 1 is issued underneath it;
            # the staging buffer is free again as soon as the device-to-device copy into the graph's input buffer has run
            main.wait_event(ready[cur])
            if gpair is not None:
                res = gpair[cur]()  # its input buffer was filled by the copy stream
                freed[cur].record(main)  # ... and is free again when this replay is over (the weight gradients read it last)
            elif gstep is not None:
                gstep.stage_packed(stage[cur][0])
                freed[cur].record(main)
                res = gstep()
            else:
                res = task.fused_step(as_batch(stage[cur]))
                freed[cur].record(main)
            if lag == 0:
                out_host.copy_(res, non_blocking=True)
                prefetch(i)
                main.synchronize()  # the D2H read of this step's loss/metrics
            else:
                out_hosts[cur].copy_(res, non_blocking=True)  # ordered before the next replay overwrites the result buffer
                read_done[cur].record(main)
                prefetch(i)
                if i >= 1:
                    read_done[nxt].synchronize()  # step i-1's result is on the host now
                    last = float(out_hosts[nxt][0])
        if lag:
            read_done[(n - 1) % 2].synchronize()
            return float(out_hosts[(n - 1) % 2][0])
        return float(out_host[0])

    e2e_loop(3)
    barrier()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record(torch.cuda.current_stream(dev))
    e2e_loop(args.steps)
    ee1.record(torch.cuda.current_stream(dev))
    barrier()
    e2e_ms = torch.tensor([ee0.elapsed_time(ee1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = world * B * args.steps / (float(e2e_ms.item()) * 1e-3)
    e2e_loop(3, lag=1)
    barrier()
    ee0.record(torch.cuda.current_stream(dev))
    e2e_loop(args.steps, lag=1)
    ee1.record(torch.cuda.current_stream(dev))
    barrier()
    e2e_ms = torch.tensor([ee0.elapsed_time(ee1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_pipe = world * B * args.steps / (float(e2e_ms.item()) * 1e-3)

    # ---- end to end with HBM-resident feature stores: the reference's Dataset yields row indices and
    # collates on the CPU; here only the two index vectors cross PCIe, the batch is gathered on the GPU straight into the
    # captured step's input buffers (tt_batch_gather), then step + D2H of the result, every step. ----
    e2e_store = None
    if gstep is not None and not args.no_device_store:
        from jodalrob_twotower_b200.loader import DeviceFeatureStore, GpuBatchAssembler
        n_notice, n_company = 1_000_000, 500_000
        gdev = torch.Generator(device=dev).manual_seed(99 + rank)
        mk_store = lambda n, d, sch: DeviceFeatureStore(
            torch.randn(n, d, generator=gdev, device=dev),
            torch.stack([torch.randint(0, v, (n,), generator=gdev, device=dev) for _, v in sch], 1), [k for k, _ in sch], dev)
        asm = GpuBatchAssembler(mk_store(n_notice, MODEL["notice_dense"], schema["notice"]),
                                mk_store(n_company, MODEL["company_dense"], schema["company"]))
        # the on-GPU collate writes fp32 dense features: a second capture of the same step with fp32 static inputs
        gstep_c, gstep = gstep, (tt.GraphedFusedStep(task, as_batch(resident[0])) if compact else gstep)
        idx_pin = [torch.stack([torch.randint(0, n_notice, (B,), generator=gen), torch.randint(0, n_company, (B,), generator=gen)]).pin_memory()
                   for _ in range(N_ROTATE)]
        idx_stage = [torch.empty(2, B, dtype=torch.int64, device=dev) for _ in range(2)]

        def store_loop(n):
            main = torch.cuda.current_stream(dev)
            with torch.cuda.stream(copy_stream):
                idx_stage[0].copy_(idx_pin[0], non_blocking=True)
                ready[0].record(copy_stream)
            for i in range(n):
                cur, nxt = i % 2, (i + 1) % 2
                main.wait_event(ready[cur])
                asm.assemble_into(gstep, idx_stage[cur][0], idx_stage[cur][1])
                freed[cur].record(main)
                res = gstep()
                out_host.copy_(res, non_blocking=True)
                if i + 1 < n:  # the next step's indices cross PCIe underneath this step (issued after the launch, like e2e_loop)
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(freed[nxt])
                        idx_stage[nxt].copy_(idx_pin[(i + 1) % N_ROTATE], non_blocking=True)
                        ready[nxt].record(copy_stream)
                main.synchronize()
            return float(out_host[0])

        store_loop(3)
        barrier()
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es0.record(torch.cuda.current_stream(dev))
        store_loop(args.steps)
        es1.record(torch.cuda.current_stream(dev))
        barrier()
        asm.check()
        es_ms = torch.tensor([es0.elapsed_time(es1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(es_ms, op=dist.ReduceOp.MAX)
        e2e_store = {"value": world * B * args.steps / (float(es_ms.item()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * 8 * B,
                     "d2h_bytes_per_step": 20, "gather_launches_per_step": 2,
                     "store": f"{n_notice:,} notice rows + {n_company:,} company rows resident in HBM "
                              f"({(asm.notice.nbytes() + asm.company.nbytes()) / 1e9:.2f} GB, fp32 dense + int64 ids), uniform row indices"}
        del asm
        gstep = gstep_c
        torch.cuda.empty_cache()

    # ---- per-kernel device times (eager launches bracketed by CUDA events on the launching stream) ----
    kern = {}
    if rank == 0 or world > 1:
        eng.profile = []
        n_prof = 5
        for i in range(n_prof):
            step_resident(i)
        torch.cuda.synchronize(dev)
        for tag, nker, a, b in eng.profile:
            d = kern.setdefault(tag, {"ms": 0.0, "calls": 0, "kernels": nker})
            d["ms"] += a.elapsed_time(b)
            d["calls"] += 1
        eng.profile = None
        for d in kern.values():
            d["ms"] = d["ms"] / n_prof
            d["calls"] = d["calls"] // n_prof
    # ---- per-kernel device times of the GRAPH REPLAYS themselves (CUPTI through torch.profiler, after the timed region): the
    # kernels as they run inside the step, concurrency included; every rank replays (the step has cross-rank exchanges)
    gk = {}
    n_cupti = 5
    try:
        from torch.profiler import ProfilerActivity, profile
        for i in range(2):
            run_step(i)
        barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(n_cupti + 1):
                run_step(i)
            torch.cuda.synchronize(dev)
        for key, (us, cnt) in cupti_kernel_table(prof, n_cupti).items():
            name = key.replace("void ", "").replace("tt::", "")
            name = name.split("(")[0]
            d = gk.setdefault(name, {"us_per_step": 0.0, "launches_per_step": 0.0})
            d["us_per_step"] += us
            d["launches_per_step"] += cnt
    except Exception as e:  # the line must still come out
        gk = {"error": {"us_per_step": 0.0, "launches_per_step": 0.0, "what": repr(e)}}
    barrier()
    if rank != 0:
        _finish(world, dev)
        return
    pk = peaks()
    graph_kernels = [dict(kernel=k, us_per_step=round(v["us_per_step"], 2), launches_per_step=round(v["launches_per_step"], 2),
                          us_per_launch=round(v["us_per_step"] / max(v["launches_per_step"], 1e-9), 2))
                     for k, v in sorted(gk.items(), key=lambda kv: -kv[1]["us_per_step"])]
    table = []
    for tag, d in sorted(kern.items(), key=lambda kv: -kv[1]["ms"]):
        a = algorithmic(tag, B, schema, world)
        row = {"call": tag, "ms": round(d["ms"], 5), "kernels": d["kernels"] * max(d["calls"], 1)}
        per_call = d["ms"] / max(d["calls"], 1) * 1e-3
        if "bytes" in a:
            row["GBps"] = round(a["bytes"] / per_call / 1e9, 1)
            row["frac_hbm"] = round(row["GBps"] / pk["hbm_gbs"], 4)
        if "flops" in a:
            row["TFLOPs"] = round(a["flops"] / per_call / 1e12, 2)
        table.append(row)
    # ---- roofline of the dominant kernel: the logits pass with the largest share of the step (graph-replay kernel time, CUPTI);
    # the eager CUDA-event bracket of the same C-ABI call is listed next to it
    roofline = None
    sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
    mufu_peak = 16 * 148 * sm_hz * 1e6  # ex2.approx per second: 16 / clk / SM (scripts/microbench/mufu2.cu)
    logits = [r for r in graph_kernels if r["kernel"].startswith("simce_tc_kernel")]
    if logits:
        top = logits[0]
        a = algorithmic(top["kernel"], B, schema, world)
        sec = top["us_per_launch"] * 1e-6
        ach = a["flops"] / sec / 1e12
        which = {"simce_tc_kernel<0": "p1", "simce_tc_kernel<2": "p2", "simce_tc_kernel<1": "p3"}.get(top["kernel"][:17], "p2")
        call = {"p1": "tt_simce3_rows", "p2": "tt_simce3_grad", "p3": "tt_simce3_grad"}[which]
        ev_ms = [r["ms"] for r in table if r["call"].startswith(call) and "finish" not in r["call"]]
        traffic = None
        prof_json = os.path.join(ROOT, "profiles", f"r02_ncu_simce_{which}.json")
        if os.path.isfile(prof_json) and B == 8192 and world == 1:
            with open(prof_json) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")  # dram read + write of one launch, ncu --set full capture
        roofline = {"kernel": top["kernel"] + {"p1": " (P1: row sums of exp)", "p2": " (P2: column sums of exp + company gradient)",
                                              "p3": " (P3: notice gradient)"}[which],
                    "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                    "traffic": traffic, "peak_source": pk["source"] + ", burst bf16",
                    "algorithmic": f"{a['flops'] / 1e9:.2f} GFLOP per launch (B_l x B_g x F logits" +
                                   (", computed once" if which == "p1" else " recomputed + dS.Y") + f"), {a['exps'] / 1e6:.1f} M exp",
                    "us_per_launch": top["us_per_launch"], "timing": "CUPTI kernel duration inside the graph replays",
                    "cuda_event_ms_per_call_eager": ev_ms[:2] or None,
                    "share_of_step": top["us_per_step"] / (ms / args.steps * 1e3),
                    "note": "K = F = 64 and one exp per logit: the pass is bound by the MUFU pipe (16 ex2/clk/SM) and the epilogue's issue "
                            "slots, not by the tensor pipe or HBM — see `mufu` (fraction of the MUFU peak at the sampled SM clock)",
                    "mufu": {"achieved_exps_per_s": a["exps"] / sec, "peak_exps_per_s": mufu_peak, "frac": a["exps"] / sec / mufu_peak}}
    # ---- where the step goes, and what global negatives allow when scaling out (weak scaling in B_l: every rank's logits block is
    # B_l x W B_l, so the three logits passes grow with W while everything else stays constant)
    logits_us = sum(r["us_per_step"] for r in logits)
    peer_us = sum(r["us_per_step"] for r in graph_kernels if r["kernel"].startswith("peer_") or "ncclDevKernel" in r["kernel"])
    scaling_model = {"logits_passes_us": round(logits_us, 1), "peer_exchange_kernels_us": round(peer_us, 1),
                     "step_us": round(ms / args.steps * 1e3, 1)}
    n1 = os.path.join(ROOT, "profiles", "r02_bench_n1.json")
    if os.path.isfile(n1):
        try:
            with open(n1) as f:
                d1 = json.load(f)
            s1, l1 = d1["ms_per_step"] * 1e3, d1["scaling_model"]["logits_passes_us"]
            ideal = s1 + (world - 1) * l1
            scaling_model.update({"n1_step_us": round(s1, 1), "n1_logits_passes_us": round(l1, 1),
                                  "ideal_step_us_given_global_negatives": round(ideal, 1),
                                  "ideal_weak_scaling_efficiency": round(s1 / ideal, 3),
                                  "achieved_over_ideal": round(ideal / (ms / args.steps * 1e3), 3),
                                  "model": "N=1 step + (W - 1) x N=1 logits passes (profiles/r02_bench_n1.json); exchanges cost nothing in the ideal"})
        except Exception:
            pass
    tables = None
    if world == 1 and not args.no_tables:
        del gstep, packed_res, resident
        torch.cuda.empty_cache()
        tables = tables_section(dev, pk)
    cpu = cpu_baseline(schema, B) if (world == 1 and not args.no_cpu_baseline) else None
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"precision": ("bf16 activations / gradients / tower outputs: every Linear (forward, dgrad, wgrad) and the B x B "
                                     "sim-CE on tcgen05 with fp32 accumulation in TMEM; BatchNorm statistics, reductions, "
                                     "master weights and Adam in fp32 (fp64 partials)") if args.dtype == "bf16"
                                    else "fp32 everywhere (SIMT)",
                       "workload": f"README-default two-tower train step (emb 32, towers [128,64]->64, 38 tables / 64,141 rows), "
                                   f"B={B} per GPU, in-batch negatives, dropout 0.1, exact dense Adam lr 1e-3 wd 1e-5",
                       "batch_per_gpu": B, "global_batch": B * world, "optimizer": args.optimizer,
                       "parallelism": f"dp{world}" + (" row-wise sharded tables" if world > 1 else ""),
                       "l2_policy": f"inputs rotate over {N_ROTATE} resident batches ({in_bytes * N_ROTATE / 1e6:.0f} MB > 126 MB L2)",
                       "cuda_graph": bool(use_graph),
                       "inputs": ("dense features staged as bf16 (GraphedFusedStep(compact_inputs=True): same bits as the engine's own "
                                  "cast of fp32 inputs), ids int64") if compact else "dense features fp32, ids int64"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 20},
            "e2e_pipelined": {"value": e2e_pipe, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 20,
                              "note": "supplementary: as e2e, but the host waits for step i's result only after launching step i+1"},
            "e2e_device_store": e2e_store,
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "parity": parity, "roofline": roofline, "scaling_model": scaling_model, "cpu_baseline": cpu,
            "hbm": None if not tables else {k: {kk: round(vv["frac_hbm"], 3) for kk, vv in t.items() if isinstance(vv, dict) and "frac_hbm" in vv}
                                            for k, t in tables.items()},
            "tables": tables, "kernels": graph_kernels[:24], "calls": table, "loss_last": loss_last}
    print(json.dumps(line), flush=True)
    _finish(world, dev)


def _finish(world, dev):
    """Multi-rank teardown: captured graphs hold NCCL work and symmetric-memory handles; skip the destructors."""
    if world > 1:
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8192, help="samples per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--fp32-inputs", action="store_true", help="stage the dense features as fp32 (the engine then casts them on the device)")
    ap.add_argument("--optimizer", default="dense_exact", choices=["dense_exact", "lazy_exact", "sparse"],
                    help="table update: dense_exact streams every row every step (torch.optim.Adam as the reference runs it); lazy_exact "
                         "gives the same bits by replaying skipped steps when a row is next touched; sparse is NOT the reference semantics")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity pre-check against the reference port")
    ap.add_argument("--parity-batch", type=int, default=1024, help="samples per GPU in the parity pre-check (the port steps the global batch on the CPU)")
    ap.add_argument("--tables-only", nargs="?", const="full", default=None, choices=["full", "quick"])
    ap.add_argument("--e2e-single-graph", action="store_true", help="end-to-end loop: one captured graph fed through a staging buffer instead of two captures fed directly")
    ap.add_argument("--no-device-store", action="store_true", help="skip the end-to-end loop with HBM-resident feature stores")
    ap.add_argument("--no-tables", action="store_true", help="skip the large-table (HBM-bound lookup/update) section")
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg5"],
                    help="cfg2 (default): BASELINE configs[1], the headline line. cfg3: BASELINE configs[2], 38 x 10 M-row tables sharded over the "
                         "ranks. cfg5: BASELINE configs[4], 500 M rows x E=64 over the ranks, bags of 20 ids (both under torchrun; "
                         "--optimizer lazy_exact / dense_exact / sparse)")
    ap.add_argument("--rows-per-table", type=int, default=0, help="cfg3 / cfg5: override the table size (default 10 M / 500 M / 38)")
    ap.add_argument("--zipf", type=float, default=1.05)
    args = ap.parse_args()
    if args.tables_only:  # development aid: only the large-table section (not the contract line)
        torch.cuda.set_device(0)
        print(json.dumps({"tables": tables_section(torch.device("cuda", 0), peaks(), quick=args.tables_only == "quick")}), flush=True)
    elif args.impl == "reference":
        run_reference(args)
    elif args.config in ("cfg3", "cfg5"):
        run_cfg3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
