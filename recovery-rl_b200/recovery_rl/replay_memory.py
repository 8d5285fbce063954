"""Replay buffers resident in HBM (reference recovery_rl/replay_memory.py:11-75).

Same classes and methods as the reference (`push`, `sample`, `__len__`), but the buffer is a device ring of
32-byte fp32 records and `sample` is the CUDA kernel that reproduces CPython's `random.sample` index stream
bit for bit (csrc/replay.cu).  As in the reference, BOTH memories draw from ONE generator and each
constructor reseeds it (`random.seed(seed)`, replay_memory.py:16,41): here that generator is a device-side
MT19937 state shared through this module.
"""
import numpy as np
import torch

from . import native

FLAG_CHUNK = 512
_shared = {"mt": None}          # the module-level `random` state of the reference, on the device


def _seed_shared(seed, device):
    st = native.mt19937_seed(seed).to(device)
    if _shared["mt"] is None or _shared["mt"].device != st.device:
        _shared["mt"] = st
    else:
        _shared["mt"].copy_(st)
    return _shared["mt"]


class ReplayMemory(object):
    is_constraint = False

    def __init__(self, capacity, seed, device=None):
        native.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        _seed_shared(seed, self.device)                         # random.seed(seed)
        self.capacity = int(capacity)
        self.cap_pad = (self.capacity + 15) // 16 * 16
        self.ring = torch.zeros(self.cap_pad, 8, device=self.device)
        self.counters = torch.zeros(native.NUM_COUNTERS, dtype=torch.int64, device=self.device)
        self.flags = torch.zeros(self.cap_pad, dtype=torch.uint8, device=self.device) if self.is_constraint else None
        self.position = 0
        self._len = 0
        self._stage = torch.zeros(1, 8).pin_memory()
        self._stage_dev = torch.zeros(1, 8, device=self.device)
        self._out = None

    # ---- reference API -----------------------------------------------------------------------------
    def push(self, state, action, reward, next_state, done):
        s = self._stage[0]
        s[0], s[1] = float(state[0]), float(state[1])
        s[2], s[3] = float(action[0]), float(action[1])
        s[4] = float(reward)
        s[5], s[6] = float(next_state[0]), float(next_state[1])
        s[7] = float(done)
        self._stage_dev.copy_(self._stage, non_blocking=True)
        native.replay_push(self.ring, self.capacity, self._stage_dev, 1, self.counters, cons_flags=self.flags)
        self._len = min(self._len + 1, self.capacity)
        self.position = (self.position + 1) % self.capacity

    def push_many(self, rec):
        """rec: float32 [n, 8] rows (s0, s1, a0, a1, r|c, s2_0, s2_1, mask), pushed in order."""
        rec = np.ascontiguousarray(rec, np.float32)
        n = len(rec)
        if n:
            native.replay_push(self.ring, self.capacity, torch.from_numpy(rec).to(self.device), n, self.counters,
                               cons_flags=self.flags)
            self._len = min(self._len + n, self.capacity)
            self.position = (self.position + n) % self.capacity

    def _outputs(self, batch_size):
        if self._out is None or self._out[0].shape[0] < batch_size:
            d = self.device
            self._out = (torch.zeros(batch_size, 2, device=d), torch.zeros(batch_size, 2, device=d),
                         torch.zeros(batch_size, device=d), torch.zeros(batch_size, 2, device=d),
                         torch.zeros(batch_size, device=d), torch.zeros(batch_size, dtype=torch.int64, device=d))
        return self._out

    def _flag_chunk(self):
        """flag-chunk size of the stratified sampler: 512 B, grown in 512-B steps so that a ring never needs more than
        the 8,192 chunks replay.cu handles (--replay_size above 4M transitions)."""
        return FLAG_CHUNK * max(1, -(-self.cap_pad // (8192 * FLAG_CHUNK)))

    def _sample_device(self, batch_size, outs, pos_fraction=None, rows_counter=native.C_SAC_ROWS, idx=None):
        cfg = native.sample_config(self.cap_pad if (self.is_constraint and pos_fraction is not None) else self.capacity,
                                   int(batch_size), self.is_constraint, pos_fraction, gate_mode=0, chunk=self._flag_chunk())
        chunk_counts = None
        if self.is_constraint and pos_fraction is not None:
            chunk = self._flag_chunk()
            n_chunks = (self.cap_pad + chunk - 1) // chunk
            if getattr(self, "_chunk_counts", None) is None:
                self._chunk_counts = torch.zeros(2, n_chunks, dtype=torch.int32, device=self.device)
            chunk_counts = self._chunk_counts
            native.replay_flag_count(self.flags, self.cap_pad, chunk, chunk_counts)
        native.replay_sample(cfg, self.ring, _shared["mt"], self.counters, rows_counter, outs[0], outs[1], outs[2],
                             outs[3], outs[4], out_idx=idx, cons_flags=self.flags, chunk_counts=chunk_counts)

    def sample(self, batch_size, pos_fraction=None):
        if batch_size > self._len:
            raise ValueError("Sample larger than population or is negative")
        o = self._outputs(batch_size)
        self._sample_device(batch_size, o, pos_fraction, idx=o[5])
        self._raise_on_sampler_error()
        self.last_idx = o[5][:batch_size].cpu().numpy()
        return tuple(x[:batch_size].cpu().numpy() for x in o[:5])

    def sample_into(self, arena, which, batch_size, pos_fraction=None):
        """device-to-device: the batch lands in the agent's update scratch, the row count in its counters."""
        names = {"sac": ("sac_s", "sac_a", "sac_r", "sac_s2", "sac_m"), "qr": ("qr_s", "qr_a", "qr_c", "qr_s2", "qr_m")}[which]
        if batch_size > self._len:                           # random.sample raises (replay_memory.py:28,57-66)
            raise ValueError("Sample larger than population or is negative")
        outs = [arena.scratch(n) for n in names]
        rc = native.C_SAC_ROWS if which == "sac" else native.C_QRISK_ROWS
        self._sample_device(batch_size, outs, pos_fraction, rows_counter=rc)
        if self.is_constraint and pos_fraction is not None:
            # too few positives / negatives for the stratified draw: the kernel flags it (rows = 0 would silently skip
            # the update); the reference's random.sample raises here
            self._raise_on_sampler_error()
        arena.counters[rc:rc + 1].copy_(self.counters[rc:rc + 1])

    def _raise_on_sampler_error(self):
        if int(self.counters[native.C_ERROR].item()):
            self.counters[native.C_ERROR] = 0                # sticky on the device; the exception carries it from here
            raise ValueError("Sample larger than population or is negative")

    def __len__(self):
        return self._len


class ConstraintReplayMemory(ReplayMemory):
    """replay_memory.py:36-75: adds the positive-example flags and stratified sampling."""
    is_constraint = True

    @property
    def pos_idx(self):
        return (self.flags[:self.capacity] & 1).to(torch.float64).cpu().numpy()
