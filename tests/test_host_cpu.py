"""CPU-side tests: the C-ABI library loads and exports every declared symbol, host logic (schema, KJT, shard plan) and a
world-size-2 gloo check of the multi-rank routing logic. No compute calls (no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as orc
from tests.helpers import GOLDEN, readme_schema

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    import jodalrob_twotower_b200 as tt
    protos = tt.native.parse_header()
    assert len(protos) >= 35
    assert os.path.isfile(tt.native.LIB_PATH), "run __graft_entry__.build() first"
    cdll = ctypes.CDLL(tt.native.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), f"{name} declared in include/twotower_b200.h but not exported"
    lib = tt.native.lib()
    assert lib.cdll.tt_version() == 1
    import re
    decl = set(re.findall(r"\b(tt_\w+)\s*\(", open(tt.native.HEADER).read()))
    assert decl == set(protos), decl ^ set(protos)  # the parser saw every prototype


def test_no_cpu_fallback():
    import jodalrob_twotower_b200 as tt
    with pytest.raises(tt.native.TwoTowerNativeError):
        tt.create_two_tower_train_task(["a"], ["b"], vocab_sizes={"notice": {"a": 5}, "company": {"b": 4}}, device="cpu")
    src = open(os.path.join(ROOT, "jodalrob-twotower_b200", "engine.py")).read() + \
        open(os.path.join(ROOT, "jodalrob-twotower_b200", "modules.py")).read()
    assert "oracle" not in src.replace("oracle = ", "")  # the product never imports the checker


def test_schema_loader_matches_golden_schema():
    import jodalrob_twotower_b200 as tt
    meta = os.path.join(GOLDEN, "tiny_metadata.csv")
    keys = tt.categorical_keys_from_metadata(meta, "notice")
    assert keys == ["flag_a", "org_code", "method", "nocount", "region"]
    vs = tt.vocab_sizes_from_metadata(meta, "notice", keys)
    assert vs == {"flag_a": 12, "org_code": 300, "method": 25, "nocount": 1000, "region": 51}
    assert tt.vocab_sizes_from_metadata("/nonexistent.csv", "notice", ["x"]) == {"x": 1000}  # cat_embed.py:82-85
    assert tt.vocab_sizes_from_metadata(meta, "notice", ["missing_col"]) == {"missing_col": 1000}
    s = readme_schema()
    assert sum(v for _, v in s["notice"]) + sum(v for _, v in s["company"]) == 64141


def test_kjt_surface():
    import jodalrob_twotower_b200 as tt
    v = torch.arange(12)
    k = tt.KeyedJaggedTensor.from_lengths_sync(["a", "b", "c"], v, torch.ones(12, dtype=torch.int64))
    assert k.keys() == ["a", "b", "c"] and torch.equal(k.values(), v) and k.lengths().sum() == 12
    assert torch.equal(k.offsets(), torch.arange(13)) and k.to("cpu").layout == "sample_major" and k.device().type == "cpu"
    j = tt.KeyedJaggedTensor.from_offsets_sync(["a"], v, torch.tensor([0, 5, 12]))
    assert torch.equal(j.lengths(), torch.tensor([5, 7]))


def test_shard_plan_properties():
    from jodalrob_twotower_b200.dist import ShardPlan
    s = readme_schema()
    vocab = [v for _, v in s["notice"]] + [v for _, v in s["company"]]
    for world in (1, 2, 4, 8):
        p = ShardPlan(vocab, world)
        assert p.local_rows == sum(-(-v // world) for v in vocab)
        rng = np.random.default_rng(world)
        for k, v in enumerate(vocab):
            ids = rng.integers(0, v, 50)
            own, loc = orc.rw_shard(ids[None, :], [v] * 50, world)
            for i, idx in enumerate(ids):
                assert p.owner(k, int(idx)) == own[0, i] and p.local_row(k, int(idx)) == p.local_base[k] + loc[0, i]
            cover = sum(p.shard_range(k, r)[1] - p.shard_range(k, r)[0] for r in range(world))
            assert cover == v  # shards tile the table exactly


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from jodalrob_twotower_b200.dist import ShardPlan
    vocab0, vocab1 = [12, 300, 25], [51, 7]
    K0, K1, Bl = 3, 2, 40
    plan = ShardPlan(vocab0 + vocab1, world)
    rng = np.random.default_rng(100 + rank)
    ids0 = np.stack([rng.integers(-2, v + 3, Bl) for v in vocab0], 1)
    ids1 = np.stack([rng.integers(-2, v + 3, Bl) for v in vocab1], 1)
    local = np.concatenate([orc.decode_clamp(ids0.reshape(-1), vocab0).reshape(-1), orc.decode_clamp(ids1.reshape(-1), vocab1).reshape(-1)])
    mine = torch.from_numpy(local.astype(np.int32))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)          # the id all-gather of the sharded step
    ids_all = torch.cat(gathered).numpy()
    n_l, n0 = Bl * (K0 + K1), Bl * K0
    g = np.arange(world * n_l)
    sl = g % n_l
    k = np.where(sl < n0, sl % K0, K0 + (sl - n0) % K1)
    owner = ids_all // np.asarray(plan.block)[k]
    own_count = torch.tensor([int((owner == rank).sum())])
    dist.all_reduce(own_count)               # every id is owned exactly once across ranks
    # loss normalisation composition: global mean == sum of local sums / B_g
    x = torch.from_numpy(rng.standard_normal(Bl))
    tot = x.sum().clone()
    dist.all_reduce(tot)
    allx = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(allx, x)
    ok = int(own_count.item()) == world * n_l and abs(tot.item() / (world * Bl) - torch.cat(allx).mean().item()) < 1e-12
    ok = ok and int(owner.min()) >= 0 and int(owner.max()) < world
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gloo_world2_routing_covers_every_id_once():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_sass_shows_blackwell_native_instructions():
    """What proves a Blackwell-native kernel without a GPU (B200_PROFILING.md): the built library's SASS carries tcgen05 MMAs
    (UTCHMMA), tensor-memory loads/stores (LDTM/STTM) and TMA loads (UTMALDG), and no legacy mma.sync path (HMMA)."""
    import shutil
    import subprocess
    from jodalrob_twotower_b200 import native
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump) and shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", native.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "MUFU.EX2"):
        assert mnemonic in sass, f"{mnemonic} missing from the SASS of {native.LIB_PATH}"
    kernels = {k: sass.count(k) for k in ("simce_tc_kernel", "gemm_tc_kernel")}
    assert all(v > 0 for v in kernels.values()), kernels
    assert " HMMA" not in sass and "HGMMA" not in sass  # no mma.sync / wgmma fallbacks


def test_reference_copy_recipe_and_bench_reference_arm_use_the_reference_itself():
    """oracle/make_ref.py leaves a verbatim copy of the hot-path reference files in oracle/_ref (git-ignored build output);
    bench.py's reference arm / cpu_baseline then time the UNMODIFIED reference (kind "reference"), and the port used everywhere
    else as the checker gives the same loss bit for bit on the same inputs."""
    import filecmp
    import bench
    from oracle import make_ref
    if os.path.isdir("/root/reference"):
        assert make_ref.make(verbose=False)
        for rel in make_ref.FILES:
            assert filecmp.cmp(os.path.join("/root/reference", rel), os.path.join(make_ref.DST, rel), shallow=False), rel
    if not os.path.isfile(os.path.join(make_ref.DST, make_ref.FILES[0])):
        pytest.skip("oracle/_ref not prepared (no /root/reference on this machine)")
    schema = bench.load_schema()
    gen = torch.Generator().manual_seed(3)
    raw = bench.synth_batch(schema, 64, gen)
    losses = {}
    for decode in ("loop", "vectorized"):
        ref = bench.RefStepper(schema, decode=decode)
        losses[ref.kind] = [ref.step(ref.make_batch(*raw)) for _ in range(2)]
    assert set(losses) == {"reference", "port"}
    # dropout 0.1 draws from the global torch RNG in both: seed equal => masks equal only if the RNG consumption matches,
    # which build_*_task(seed=0) + identical module order guarantees for the first step
    assert losses["reference"][0] == losses["port"][0]


def test_bench_cupti_table_drops_the_first_profiled_step():
    """bench.cupti_kernel_table: per kernel, the launches of the first profiled step are dropped (under torchrun the first
    peer-memory kernel of that step waits for the rank that left the profiler's start-up last) and the rest is averaged per step."""
    import bench

    class _Range:
        def __init__(self, start):
            self.start = start

    class _Ev:
        def __init__(self, key, start, dur):
            self.key, self.time_range, self.device_time_total = key, _Range(start), dur

    class _Prof:
        def events(self):
            skewed = [_Ev("peer_allreduce", 0.0, 28000.0)] + [_Ev("peer_allreduce", 10.0 * i, 9.0) for i in range(1, 8)]  # 2 per step
            gemm = [_Ev("gemm", 3.0 + 10 * i, 5.0 + (i % 2)) for i in range(8)]                                          # 2 per step
            host_only = [_Ev("cudaLaunchKernel", 1.0, 0.0)]
            return skewed + gemm + host_only

    t = bench.cupti_kernel_table(_Prof(), n_steps=3, drop_first=1)
    assert set(t) == {"peer_allreduce", "gemm"}
    assert t["peer_allreduce"] == (18.0, 2.0)          # 6 launches of 9 us over 3 steps: the 28 ms outlier is gone
    assert t["gemm"][1] == 2.0 and abs(t["gemm"][0] - 11.0) < 1e-9
