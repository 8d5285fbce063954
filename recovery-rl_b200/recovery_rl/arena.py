"""Device-resident agent state: one flat fp32 arena (parameters, gradients, Adam moments, weight images,
update scratch) plus the int64 counter block, laid out by librrl.so (csrc/agent_layout.cuh).

Host code only allocates, initialises (xavier draws from the torch global RNG in the reference's
construction order, sac.py:82-131 / qrisk.py:36-75) and views this memory; all arithmetic on it is done
by the CUDA kernels behind include/rrl.h.
"""
import numpy as np
import torch

from . import native

NETS = native.NET_NAMES


class AgentArena(object):
    def __init__(self, device, max_batch=256, allocator=None, **cfg_kwargs):
        native.require_cuda()
        self.device = torch.device(device)
        max_batch = (int(max_batch) + 31) // 32 * 32
        self.cfg = native.agent_config(max_batch=max_batch, **cfg_kwargs)
        self.max_batch = max_batch
        n = native.agent_arena_floats(self.cfg)
        if n <= 0:
            raise native.RRLError("bad agent config: %s" % native.lib().rrl_last_error().decode())
        # allocator(n) -> zeroed float32 device tensor; the sharded engine passes a symmetric-memory allocator so that
        # peers can read the gradient block over NVLink (dist_utils.PeerArena)
        self.arena = torch.zeros(n, dtype=torch.float32, device=self.device) if allocator is None else allocator(n)
        assert self.arena.dtype == torch.float32 and self.arena.numel() >= n and self.arena.is_contiguous()
        self.counters = torch.zeros(native.NUM_COUNTERS, dtype=torch.int64, device=self.device)
        self._views = {}
        self.grad_off, self.grad_count = native.agent_grad_range(self.cfg, -1)
        native.agent_init_scalars(self.cfg, self.arena)       # alpha / nu / lambda and their logs (sac.py:48,56-72,99-101)

    # ---- views ---------------------------------------------------------------------------------
    def tensor(self, net, i):
        """view of parameter i of `net` (torch parameters() order of the reference module)."""
        net = NETS.index(net) if isinstance(net, str) else net
        off, rows, cols = native.agent_tensor_info(self.cfg, net, i)
        n = rows * (cols if cols else 1)
        v = self.arena[off:off + n]
        return v.view(rows, cols) if cols else v

    def grad(self, net, i):
        net = NETS.index(net) if isinstance(net, str) else net
        off, rows, cols = native.agent_tensor_info(self.cfg, net, i)
        n = rows * (cols if cols else 1)
        v = self.arena[self.grad_off + off:self.grad_off + off + n]
        return v.view(rows, cols) if cols else v

    def adam_state(self, net, i):
        """views (exp_avg, exp_avg_sq) of parameter i of a trainable net (torch.optim.Adam state)."""
        net = NETS.index(net) if isinstance(net, str) else net
        off, rows, cols = native.agent_tensor_info(self.cfg, net, i)
        n = rows * (cols if cols else 1)
        m_off, _ = native.agent_scratch_info(self.cfg, "adam_m")
        v_off, _ = native.agent_scratch_info(self.cfg, "adam_v")
        return self.arena[m_off + off:m_off + off + n], self.arena[v_off + off:v_off + off + n]

    def load_optimizer(self, net, optimizer, t_counter):
        """copy a torch.optim.Adam state (exp_avg, exp_avg_sq, step) of `net` into the arena."""
        step = 0
        for i, p in enumerate(optimizer.param_groups[0]["params"]):
            st = optimizer.state.get(p, {})
            m, v = self.adam_state(net, i)
            if "exp_avg" in st:
                m.copy_(st["exp_avg"].detach().reshape(-1).to(self.device))
                v.copy_(st["exp_avg_sq"].detach().reshape(-1).to(self.device))
                step = int(st["step"])
            else:
                m.zero_()
                v.zero_()
        self.counters[t_counter] = step

    def num_tensors(self, net):
        return native.agent_num_tensors(NETS.index(net) if isinstance(net, str) else net)

    def scratch(self, name, width=None):
        if name not in self._views:
            off, cnt = native.agent_scratch_info(self.cfg, name)
            self._views[name] = self.arena[off:off + cnt]
        v = self._views[name]
        return v.view(-1, width) if width else v

    # ---- scalar multipliers of the comparison branches (sac.py:56-72, 95-101) --------------------------
    def scalars(self):
        """(float32 view, float64 view) of the scalar block: indices native.S_* / native.D_*."""
        v = self.scratch("scalars")
        return v, v[native.S_F64_BASE:].view(torch.float64)

    def set_nu_arg(self, nu):
        """the `nu` argument of SAC.update_parameters (experiment.py:406: nu_schedule(i_episode))."""
        self.scratch("scalars")[native.S_NU_ARG] = float(nu)

    def flat_grads(self):
        """the contiguous gradient block [critic | policy | qrisk | recovery] (NCCL all-reduce payload)."""
        return self.arena[self.grad_off:self.grad_off + self.grad_count]

    # ---- parameter IO ----------------------------------------------------------------------------
    def load_params(self, getter):
        """getter(net_name, i) -> array-like or None (keep).  Rebuilds the weight images afterwards."""
        for net in NETS:
            for i in range(self.num_tensors(net)):
                v = getter(net, i)
                if v is not None:
                    t = self.tensor(net, i)
                    t.copy_(torch.as_tensor(np.asarray(v), dtype=torch.float32).reshape(t.shape))
        self.refresh()

    def load_modules(self, modules):
        """modules: {net_name: torch.nn.Module with the reference's parameter order}."""
        def getter(net, i):
            if net not in modules:
                return None
            ps = list(modules[net].parameters())
            if i >= len(ps):             # DeterministicPolicy: no log_std head (the arena keeps the Gaussian layout, head unused)
                return None
            return ps[i].detach().cpu().numpy()
        self.load_params(getter)

    def params(self, net):
        return [self.tensor(net, i).detach().cpu().numpy().copy() for i in range(self.num_tensors(net))]

    def refresh(self):
        native.agent_refresh(self.cfg, self.arena)

    def hard_update(self, dst, src):
        native.hard_update(self.cfg, self.arena, NETS.index(dst), NETS.index(src))

    def soft_update(self, dst, src, tau):
        native.soft_update(self.cfg, self.arena, NETS.index(dst), NETS.index(src), tau)

    # ---- batches (tests / N = 1 drop-in path; the vector engine samples on the device) ------------
    def set_batch(self, which, s, a, r, s2, m):
        """which: 'sac' | 'qr'.  Copies a host batch into the update scratch and sets the row counter."""
        names = {"sac": ("sac_s", "sac_a", "sac_r", "sac_s2", "sac_m"), "qr": ("qr_s", "qr_a", "qr_c", "qr_s2", "qr_m")}[which]
        rows = len(s)
        if rows > self.max_batch:
            raise native.RRLError("batch of %d rows exceeds max_batch %d" % (rows, self.max_batch))
        for name, x, w in zip(names, (s, a, r, s2, m), (2, 2, 1, 2, 1)):
            t = torch.as_tensor(np.asarray(x), dtype=torch.float32).reshape(-1).to(self.device)
            self.scratch(name)[:rows * w].copy_(t)
        self.counters[native.C_SAC_ROWS if which == "sac" else native.C_QRISK_ROWS] = rows
        return rows
