"""GPU parity: rrl_replay_push / rrl_replay_flag_count / rrl_replay_sample against the index streams
recorded from the reference's ReplayMemory / ConstraintReplayMemory (CPython random.sample)."""
import os

import numpy as np
import pytest
import torch

from oracle import replay as oreplay

pytestmark = pytest.mark.gpu
CHUNK = 512


class GpuBuffers(object):
    def __init__(self, native, dev, cap, seed):
        self.n, self.dev, self.cap = native, dev, cap
        self.cap_pad = (cap + 15) // 16 * 16
        self.task = torch.zeros(cap, 8, device=dev)
        self.cons = torch.zeros(cap, 8, device=dev)
        self.flags = torch.zeros(self.cap_pad, dtype=torch.uint8, device=dev)
        self.n_chunks = (self.cap_pad + CHUNK - 1) // CHUNK
        self.chunk_counts = torch.zeros(2, self.n_chunks, dtype=torch.int32, device=dev)
        self.cnt = torch.zeros(native.NUM_COUNTERS, dtype=torch.int64, device=dev)
        self.mt = native.mt19937_seed(seed).to(dev)

    def push(self, rec_task, rec_cons):
        n = len(rec_task)
        self.n.replay_push(self.task, self.cap, torch.from_numpy(rec_task).to(self.dev), n, self.cnt)
        self.n.replay_push(self.cons, self.cap, torch.from_numpy(rec_cons).to(self.dev), n, self.cnt,
                           cons_flags=self.flags)

    def sample(self, B, is_cons, pf, gate_mode=0, gate_pf=-1.0):
        out = [torch.zeros(B, 2, device=self.dev), torch.zeros(B, 2, device=self.dev), torch.zeros(B, device=self.dev),
               torch.zeros(B, 2, device=self.dev), torch.zeros(B, device=self.dev)]
        idx = torch.full((B,), -1, dtype=torch.int64, device=self.dev)
        cfg = self.n.sample_config(self.cap_pad if is_cons and pf is not None else self.cap, B, is_cons, pf,
                                   gate_mode, CHUNK, gate_pf)
        if is_cons and pf is not None:
            self.n.replay_flag_count(self.flags, self.cap_pad, CHUNK, self.chunk_counts)
        rows_counter = self.n.C_QRISK_ROWS if is_cons else self.n.C_SAC_ROWS
        self.n.replay_sample(cfg, self.cons if is_cons else self.task, self.mt, self.cnt, rows_counter, *out,
                             out_idx=idx, cons_flags=self.flags if is_cons else None,
                             chunk_counts=self.chunk_counts if is_cons else None)
        torch.cuda.synchronize()
        rows = int(self.cnt[rows_counter].item())
        return rows, idx.cpu().numpy()[:rows], [o.cpu().numpy()[:rows] for o in out]


@pytest.mark.parametrize("name", ["poolset", "strat", "b1024", "strat1k"])
def test_sample_streams_bit_exact_vs_reference(native, cuda, golden_dir, name):
    z = np.load(os.path.join(golden_dir, "replay_idx.npz"))
    cap = int(z[name + "_cap"]); B = int(z[name + "_B"]); pf = float(z[name + "_pf"])
    pf = None if pf < 0 else pf
    flags = z[name + "_flags"]
    g = GpuBuffers(native, cuda, cap, int(z[name + "_seed"]))
    c = 0
    ti, ci = [], []
    for burst in z[name + "_bursts"]:
        ids = np.arange(c, c + burst, dtype=np.float32)
        rec = np.zeros((burst, 8), np.float32)
        rec[:, 0] = ids; rec[:, 4] = -1.0; rec[:, 5] = ids + 1; rec[:, 6] = 1.0; rec[:, 7] = 1.0
        rec_c = rec.copy(); rec_c[:, 4] = flags[c:c + burst]
        g.push(rec, rec_c)
        c += burst
        for _ in range(3):
            rows, idx, out = g.sample(B, False, None)
            assert rows == min(B, min(c, cap))
            assert np.array_equal(out[0][:, 0].astype(np.int64) % cap, idx)     # gather consistent with idx
            assert np.array_equal(out[3][:, 0], out[0][:, 0] + 1)
            ti.append(out[0][:, 0].astype(np.int64))
            rows, idx, out = g.sample(B, True, pf)
            ci.append(out[0][:, 0].astype(np.int64))
            assert np.array_equal(out[2], flags[out[0][:, 0].astype(np.int64)])
    assert np.array_equal(np.concatenate(ti), z[name + "_task_ids"])
    assert np.array_equal(np.concatenate(ci), z[name + "_cons_ids"])
    assert int(g.cnt[native.C_ERROR].item()) == 0


def test_gates_and_stream_not_advanced_when_closed(native, cuda):
    g = GpuBuffers(native, cuda, 2048, 9)
    rec = np.zeros((256, 8), np.float32); rec[:, 0] = np.arange(256)
    g.push(rec, rec)
    mt0 = g.mt.clone()
    rows, _, _ = g.sample(256, False, None, gate_mode=1)       # len == B: strict > fails (experiment.py:397)
    assert rows == 0 and torch.equal(mt0, g.mt)
    g.push(rec[:1], rec[:1])
    rows, idx, _ = g.sample(256, False, None, gate_mode=1)
    assert rows == 256 and not torch.equal(mt0, g.mt)
    st = oreplay.MT19937(9)
    assert np.array_equal(idx, st.sample_indices(257, 256))
    # Q_risk gate: (num_viols + offline_viols)/B > pos_fraction (experiment.py:407-410), nested in the SAC gate
    g.cnt[native.C_TASK_LEN] = 257
    mt1 = g.mt.clone()
    rows, _, _ = g.sample(256, True, None, gate_mode=2, gate_pf=0.3)
    assert rows == 0 and torch.equal(mt1, g.mt)
    g.cnt[native.C_OFFLINE_VIOLS] = 77                          # 77/256 > 0.3
    rows, idx, _ = g.sample(256, True, None, gate_mode=2, gate_pf=0.3)
    assert rows == 256
    assert np.array_equal(idx, st.sample_indices(257, 256))


def test_large_ring_set_path_matches_cpython(native, cuda):
    import random
    cap = 1 << 20
    g = GpuBuffers(native, cuda, cap, 2 ** 40 + 5)
    n = 700001
    rec = np.zeros((n, 8), np.float32); rec[:, 0] = np.arange(n) % 4096; rec[:, 1] = np.arange(n) // 4096
    g.push(rec, rec)
    random.seed(2 ** 40 + 5)
    for B in (256, 1024, 4096, 17):
        rows, idx, out = g.sample(B, False, None)
        ref = random.sample(range(n), B)
        assert np.array_equal(idx, ref)
        assert np.array_equal(out[0][:, 0] + 4096.0 * out[0][:, 1], np.array(ref, np.float64))


@pytest.mark.parametrize("cap,chunk,n_fill,p_pos", [(1 << 21, 512, 1 << 21, 0.0004), (3 * (1 << 20) + 48, 1024, 2500000, 0.3),
                                                     (1 << 23, 1024, 1 << 23, 0.0001), (70000, 2048, 65000, 0.05)])
def test_stratified_sample_large_rings_and_chunks(native, cuda, cap, chunk, n_fill, p_pos):
    """stratified (pos_fraction) sample over large rings with the larger flag chunks the engine picks for them:
    the slots must be the ones replay_memory.py:58-68 picks (np.argwhere order + random.sample index streams)."""
    rs = np.random.RandomState(cap % 1000 + chunk)
    flags_h = (rs.rand(n_fill) < p_pos).astype(np.float32)
    cap_pad = (cap + 15) // 16 * 16
    ring = torch.zeros(cap, 8, device=cuda)
    flags = torch.zeros(cap_pad, dtype=torch.uint8, device=cuda)
    cnt = torch.zeros(native.NUM_COUNTERS, dtype=torch.int64, device=cuda)
    rec = np.zeros((n_fill, 8), np.float32)
    rec[:, 0] = np.arange(n_fill) % 4096; rec[:, 1] = np.arange(n_fill) // 4096; rec[:, 4] = flags_h
    native.replay_push(ring, cap, torch.from_numpy(rec).to(cuda), n_fill, cnt, cons_flags=flags)
    n_chunks = (cap_pad + chunk - 1) // chunk
    counts = torch.zeros(2, n_chunks, dtype=torch.int32, device=cuda)
    native.replay_flag_count(flags, cap_pad, chunk, counts)
    f = flags.cpu().numpy()
    pad = np.zeros(n_chunks * chunk, np.uint8); pad[:cap_pad] = f
    assert np.array_equal(counts[0].cpu().numpy(), (pad.reshape(n_chunks, chunk) == 1).sum(1))
    assert np.array_equal(counts[1].cpu().numpy(), (pad.reshape(n_chunks, chunk) == 2).sum(1))
    seed = 77
    mt = native.mt19937_seed(seed).to(cuda)
    st = oreplay.MT19937(seed)
    pos_slots = np.argwhere(flags_h).ravel()
    neg_slots = np.argwhere(1 - flags_h).ravel()
    B, pf = 256, 0.3
    cfg = native.sample_config(cap_pad, B, True, pf, 0, chunk, -1.0)
    for _ in range(3):
        out = [torch.zeros(B, 2, device=cuda), torch.zeros(B, 2, device=cuda), torch.zeros(B, device=cuda),
               torch.zeros(B, 2, device=cuda), torch.zeros(B, device=cuda)]
        idx = torch.full((B,), -1, dtype=torch.int64, device=cuda)
        native.replay_sample(cfg, ring, mt, cnt, native.C_QRISK_ROWS, *out, out_idx=idx, cons_flags=flags, chunk_counts=counts)
        torch.cuda.synchronize()
        assert int(cnt[native.C_QRISK_ROWS].item()) == B and int(cnt[native.C_ERROR].item()) == 0
        pi = st.sample_indices(len(pos_slots), int(B * pf))
        ni = st.sample_indices(len(neg_slots), B - int(B * pf))
        want = np.concatenate([pos_slots[pi], neg_slots[ni]])
        assert np.array_equal(idx.cpu().numpy(), want)
        assert np.array_equal(out[0].cpu().numpy()[:, 0] + 4096.0 * out[0].cpu().numpy()[:, 1], want.astype(np.float64))
        assert np.array_equal(out[2].cpu().numpy(), flags_h[want])
