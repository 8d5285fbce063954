"""2-rank check of the peer-memory gradient sum (run under torch.distributed.run with 2 GPUs; launched by
tests/test_engine_gpu.py::test_peer_grads_match_nccl).  The same sharded run is done twice -- NCCL all-reduce, then the
fused peer-sum optimizer step -- and must leave bit-identical parameters on every rank (a two-term float sum is
commutative, so both orders agree exactly at world 2)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "recovery-rl_b200"))


def run(peer, rank, world, dev, tc, steps):
    import bench
    args = argparse.Namespace(env_name="navigation1", envs=2048, batch=256, seed=5, tc=tc, demos=2000, pretrain=20,
                              peer_grads=int(peer))
    eng = bench.build_engine(args, rank, world, dist.group.WORLD, False, dev)
    assert (eng.peer_arena is not None) == bool(peer), "peer mode not active: %r" % (eng.peer_error,)
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    eng.capture()
    for _ in range(steps):
        eng.replay()
    torch.cuda.synchronize()
    cn = eng.read_counters()
    assert cn["error"] == 0, cn
    lo, cnt = eng.agent.grad_off, eng.agent.grad_count
    params = eng.arena[:lo].clone()       # parameters, targets and images precede the gradient block
    return params, cn, eng.graph is not None


def resume_check(rank, world, dev):
    """sharded checkpoint: every rank writes / reads its own shard (checkpoint.pt.rank<r>of<w>); a resumed run continues
    bit-identically on every rank, and a shard of another rank is refused."""
    import tempfile
    import bench
    from recovery_rl import checkpoint
    args = argparse.Namespace(env_name="navigation1", envs=1024, batch=64, seed=7, tc=2, demos=1000, pretrain=10, peer_grads=1,
                              replay=1 << 16)
    box = [tempfile.mkdtemp(prefix="rrl_ck_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    path = os.path.join(box[0], "checkpoint.pt")
    a = bench.build_engine(args, rank, world, dist.group.WORLD, False, dev)
    for _ in range(4):
        a.step()
    torch.cuda.synchronize()
    saved = a.save(path)
    assert saved == checkpoint.shard_path(path, rank, world) and os.path.exists(saved), saved
    for _ in range(3):
        a.step()
    torch.cuda.synchronize()
    b = bench.build_engine(args, rank, world, dist.group.WORLD, False, dev)
    b.load(path)
    for _ in range(3):
        b.step()
    torch.cuda.synchronize()
    g = a.agent.grad_off
    same = torch.equal(a.arena[:g].view(torch.int32), b.arena[:g].view(torch.int32)) and torch.equal(a.state, b.state) and \
        torch.equal(a.mt_state, b.mt_state) and a.read_counters()["total_numsteps"] == b.read_counters()["total_numsteps"]
    refused = False
    try:
        st = torch.load(checkpoint.shard_path(path, (rank + 1) % world, world), map_location="cpu", weights_only=False)
        checkpoint.load_engine_state(b, st)
    except ValueError:
        refused = True
    if rank == 0:
        print("sharded resume bit-identical: %s  foreign shard refused: %s" % (same, refused), flush=True)
    assert same and refused


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    for tc in (0, 1):
        a, ca, _ = run(False, rank, world, dev, tc, 6)
        b, cb, graphed = run(True, rank, world, dev, tc, 6)
        assert graphed, "peer-mode step was not captured into a graph"
        assert ca["sac_updates"] == cb["sac_updates"] > 0 and ca["qrisk_updates"] == cb["qrisk_updates"] > 0, (ca, cb)
        # world 2: a two-term sum is order-independent -> bit-equal; more ranks: NCCL's reduction order is its own
        same = torch.equal(a, b) if world == 2 else torch.allclose(a, b, rtol=1e-3, atol=1e-4)
        other = [torch.empty_like(b) for _ in range(world)]
        dist.all_gather(other, b)
        across = all(torch.equal(o, b) for o in other)
        if rank == 0:
            print("tc=%d nccl==peer: %s  ranks identical: %s  sac_updates=%d qrisk_updates=%d max|diff|=%.3e" % (
                tc, same, across, cb["sac_updates"], cb["qrisk_updates"], float((a - b).abs().max())), flush=True)
        assert same and across
    resume_check(rank, world, dev)
    dist.barrier()
    if rank == 0:
        print("P2P_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
