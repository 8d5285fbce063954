"""Collective plumbing of the sharded engine (one process per GPU, torch.distributed; NCCL on the GPUs, gloo in
the CPU tests).  The path shards by env copies: every rank owns its envs, its replay shards and its sampler
stream; the only exchange is the gradient sum before each optimizer step (+ one scalar for the Q_risk gate).
"""
import torch

from . import native

GRAD_NETS = ("critic", "policy", "qrisk", "recovery")


def grad_ranges(cfg):
    """{net: (offset, count)} of the four trainable nets inside the flat arena (contiguous, in this order)."""
    return {name: native.agent_grad_range(cfg, native.NET_NAMES.index(name)) for name in GRAD_NETS}


def span(ranges, names):
    lo = min(ranges[n][0] for n in names)
    hi = max(ranges[n][0] + ranges[n][1] for n in names)
    return lo, hi


def all_reduce_grads(arena, ranges, names, group=None):
    """SUM the gradient block of `names` (adjacent nets -> one collective); Adam applies 1/world (grad_scale)."""
    import torch.distributed as dist
    lo, hi = span(ranges, names)
    dist.all_reduce(arena[lo:hi], op=dist.ReduceOp.SUM, group=group)


def all_reduce_sum(view, group=None):
    """in-place SUM of a small contiguous view (the scalar-multiplier gradients of the comparison branches)."""
    import torch.distributed as dist
    dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)


def sync_gate_counts(counters, group=None):
    """The Q_risk online gate (experiment.py:407-410) must open on every rank in the same step, so each rank adds
    the violations seen elsewhere: EXT_VIOLS = sum over other ranks of (num_viols + offline_viols)."""
    import torch.distributed as dist
    local = (counters[native.C_NUM_VIOLS] + counters[native.C_OFFLINE_VIOLS]).reshape(1).clone()
    total = local.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    counters[native.C_EXT_VIOLS:native.C_EXT_VIOLS + 1] = total - local


class PeerArena(object):
    """Symmetric-memory plumbing for the peer-summed optimizer step (include/rrl.h, rrl_peers_t).  torch's symmetric
    memory does the allocation and the handle exchange; the data path (barrier flags, gradient loads) is ours.
    Usage: pa = PeerArena(device, rank, world, group); arena = pa.allocate(n) [collective]; pa.peers -> native.Peers."""
    PAD_BASE = 256          # uint32 words; the first KB of the signal pad is left to torch's own primitives

    def __init__(self, device, rank, world, group=None):
        self.device, self.rank, self.world, self.group = torch.device(device), int(rank), int(world), group
        self.peers = None
        self.handle = None

    def allocate(self, n_floats):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = self.group if self.group is not None else dist.group.WORLD
        t = symm.empty(int(n_floats), dtype=torch.float32, device=self.device)
        self.handle = symm.rendezvous(t, group)
        t.zero_()
        pad = self.handle.get_signal_pad(self.rank, dtype=torch.int32)
        pad[self.PAD_BASE:self.PAD_BASE + 64].zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)
        h = self.handle
        self.epoch = torch.zeros(1, dtype=torch.int64, device=self.device)
        # `peers`: barriers are separate launches (rrl_peer_barrier);  `peers_fused`: the optimizer-step kernel runs the
        # barrier itself (rrl_peers_t::epoch) -- same pads, same generation counter
        self.peers = native.make_peers(self.rank, list(h.buffer_ptrs), list(h.signal_pad_ptrs))
        # multicast (NVLS) mapping of the arena, when the fabric has one: with RRL_MULTIMEM=1 the optimizer-step kernels read the
        # gradient SUM with multimem.ld_reduce instead of world peer loads.  Off by default: measured no faster (2 GPUs 0.3570 vs
        # 0.3555 ms, 4 GPUs 0.3787 vs 0.3782 ms per step, profiles/r2/sync_modes.txt) -- every requester's reduction still pulls
        # every source GPU's block through that GPU's links -- and the peer loads keep the rank-order (reproducible) sum
        import os
        mc = 0
        try:
            if os.environ.get("RRL_MULTIMEM", "0") == "1":
                mc = int(getattr(h, "multicast_ptr", 0) or 0)
        except Exception:
            mc = 0
        self.multicast = mc != 0
        self.peers_fused = native.make_peers(self.rank, list(h.buffer_ptrs), list(h.signal_pad_ptrs),
                                             epoch_ptr=self.epoch.data_ptr(), mc_ptr=mc)
        assert int(h.buffer_ptrs[self.rank]) == t.data_ptr()
        self.tensor = t
        return t


def shard(n_total, rank, world):
    """[lo, hi) of the env copies rank owns when a global env count is split (remainder to the first ranks)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
