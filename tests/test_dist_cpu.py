"""CPU, world_size 2, gloo: host-side logic of the sharded engine (N > 1 path) -- gradient-block ranges,
the summed all-reduce over adjacent nets, the global Q_risk gate count, env sharding, identical replicas
from the shared torch seed and distinct per-rank sampler streams."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "recovery-rl_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recovery_rl import native, dist_utils
    from recovery_rl.model import build_reference_modules
    cfg = native.agent_config(max_batch=64)
    n = native.agent_arena_floats(cfg)
    ranges = dist_utils.grad_ranges(cfg)
    arena = torch.zeros(n)
    g_off, g_cnt = native.agent_grad_range(cfg, -1)
    arena[g_off:g_off + g_cnt] = float(rank + 1)
    params_before = arena[:g_off].clone()
    dist_utils.all_reduce_grads(arena, ranges, ["critic", "policy"])
    lo, hi = dist_utils.span(ranges, ["critic", "policy"])
    ok_sum = bool((arena[lo:hi] == 3.0).all())                       # 1 + 2
    qo, qc = ranges["qrisk"]
    untouched = bool((arena[qo:qo + qc] == float(rank + 1)).all()) and torch.equal(arena[:g_off], params_before)
    counters = torch.zeros(native.NUM_COUNTERS, dtype=torch.int64)
    counters[native.C_NUM_VIOLS] = 5 + rank
    counters[native.C_OFFLINE_VIOLS] = 100 * (rank + 1)
    dist_utils.sync_gate_counts(counters)
    ext = int(counters[native.C_EXT_VIOLS])
    torch.manual_seed(4)
    w = torch.cat([p.detach().reshape(-1) for m in build_reference_modules().values() for p in m.parameters()])
    ws = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    same_init = all(torch.equal(ws[0], x) for x in ws)
    mt = native.mt19937_seed(9 + rank)
    q.put((rank, ok_sum, untouched, ext, same_init, mt[:4].tolist(), dist_utils.shard(65537, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_host_logic_gloo_world2():
    if not os.path.exists(os.path.join(ROOT, "recovery-rl_b200", "librrl.so")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("rrl_build", os.path.join(ROOT, "recovery-rl_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, sum0, unt0, ext0, init0, mt0, sh0), (r1, sum1, unt1, ext1, init1, mt1, sh1) = out
    assert sum0 and sum1 and unt0 and unt1
    assert ext0 == 6 + 200 and ext1 == 5 + 100          # the OTHER rank's num_viols + offline_viols
    assert init0 and init1
    assert mt0 != mt1
    assert sh0 == (0, 32769) and sh1 == (32769, 65537)


def test_grad_block_is_contiguous_and_ordered():
    sys.path.insert(0, os.path.join(ROOT, "recovery-rl_b200"))
    from recovery_rl import native, dist_utils
    cfg = native.agent_config(max_batch=256)
    r = dist_utils.grad_ranges(cfg)
    off, cnt = native.agent_grad_range(cfg, -1)
    pos = off
    for name in dist_utils.GRAD_NETS:
        assert r[name][0] == pos
        pos += r[name][1]
    assert pos == off + cnt == off + 404008
    sizes = {k: v[1] for k, v in r.items()}
    assert sizes == {"critic": 134664, "policy": 67592, "qrisk": 134672, "recovery": 67080}


def test_checkpoint_shards_are_per_rank():
    """sharded runs write one checkpoint file per rank and refuse a shard written by another rank / world size
    (env copies, replay shards and the seed+rank sampler stream are per-rank state)."""
    import pytest
    from recovery_rl import checkpoint
    assert checkpoint.shard_path("/x/checkpoint.pt", 0, 1) == "/x/checkpoint.pt"
    assert checkpoint.shard_path("/x/checkpoint.pt", 3, 8) == "/x/checkpoint.pt.rank3of8"
    assert len({checkpoint.shard_path("c.pt", r, 4) for r in range(4)}) == 4
    checkpoint.check_shard({"rank": 1, "world": 2}, 1, 2)
    checkpoint.check_shard({}, 0, 1)                       # files written before shards existed: one GPU
    for e, r, w in (({"rank": 0, "world": 2}, 1, 2), ({"rank": 0, "world": 1}, 0, 2), ({}, 1, 2)):
        with pytest.raises(ValueError):
            checkpoint.check_shard(e, r, w)
