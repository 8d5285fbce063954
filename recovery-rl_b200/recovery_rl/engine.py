"""VecEngine -- the vectorised Recovery RL training step on one GPU (one process per GPU).

One `step()` is what recovery_rl/experiment.py:396-452 does for one env step, for N env copies at once:

    [len(memory) > batch_size]   SAC.update_parameters            (sac.py:170-277)
    [Q_risk online gate]         QRiskWrapper.update_parameters   (qrisk.py:86-182)
    get_action                   policy -> Q_risk threshold -> recovery policy   (experiment.py:546-577)
    env.step + both memory.push  (env/*.py, replay_memory.py:21-25,47-52) + episode statistics

Everything lives in HBM; every call below only ENQUEUES kernels of librrl.so on the current stream, all
data-dependent control flow (gates, effective batch rows, ring positions) is read from the device counter
block, so the whole step is captured once into a CUDA graph and replayed.  With world_size > 1 each rank
owns N/world env copies and its own replay shards; gradients are summed inside the optimizer-step kernel over
NVLink peer memory (peer_grads: symmetric arena + flag barrier) or, as a fallback, with one NCCL all-reduce per
optimizer step over the flat gradient block (1/world applied inside Adam either way).
"""
import os

import numpy as np
import torch

from . import dist_utils, native
from .arena import AgentArena

HORIZON = {"navigation1": 100, "navigation2": 100, "maze": 100}
ACTION_SCALE = {"navigation1": 1.0, "navigation2": 1.0, "maze": float(np.float32(0.1))}
FLAG_CHUNK = 512


def _carve(specs, device=None, pinned=False):
    """specs: [(name, shape, dtype)] -> (flat uint8 buffer, {name: typed view}); every view starts on a 256-byte boundary.
    One flat buffer = ONE copy per direction for the whole set (the host face moves ~10 arrays per step)."""
    offs, total = [], 0
    for _, shape, dtype in specs:
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        offs.append((total, nbytes))
        total += (nbytes + 255) // 256 * 256
    flat = torch.zeros(total, dtype=torch.uint8).pin_memory() if pinned else torch.zeros(total, dtype=torch.uint8, device=device)
    views = {name: flat[o:o + nb].view(dtype).view(*shape) for (name, shape, dtype), (o, nb) in zip(specs, offs)}
    return flat, views


class HostInputs(object):
    """One step's random draws in ONE pinned host buffer laid out like the engine's device staging (VecEngine.new_host_inputs):
    fill the typed views in `.views` in place; submit() / step_host() then upload the whole set with a single copy."""

    def __init__(self, specs):
        self.flat, self.views = _carve(specs, pinned=True)

    def __getitem__(self, k):
        return self.views[k]

    def __contains__(self, k):
        return k in self.views

    def keys(self):
        return self.views.keys()


class VecEngine(object):
    def __init__(self, env_name, num_envs, batch_size=256, replay_size=1000000, safe_replay_size=1000000,
                 gamma=0.99, alpha=0.2, tau=0.005, lr=3e-4, gamma_safe=0.5, tau_safe=0.0002, eps_safe=0.1,
                 target_update_interval=1, use_recovery=True, mf_recovery=True, pos_fraction=-1.0,
                 disable_online_updates=False, constraint_reward_penalty=0.0, start_steps=100, seed=0,
                 device="cuda:0", rank=0, world_size=1, process_group=None, host_inputs=False, log_outputs=False,
                 use_tensor_cores=0, maze_substeps=500, dgd=False, update_nu=False, rcpo=False, auto_alpha=False,
                 nu=0.01, lambda_rcpo=0.01, disable_action_relabeling=False, mb_recovery=False, mpc_popsize=None,
                 mpc_num_elites=None, peer_grads=False, constraint_sampling=False, q_sampling_recovery=False,
                 add_both_transitions=False, deterministic=False, safe_samples=100, q_samples=1000):
        native.require_cuda()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.env_name = env_name
        self.kind = native.ENV_KIND[env_name]
        self.n = int(num_envs)
        self.B = int(batch_size)
        self.rank, self.world = int(rank), int(world_size)
        self.pg = process_group
        self.seed = int(seed)
        self.use_recovery = bool(use_recovery)
        self.mf_recovery = bool(mf_recovery)
        # the safety critic is trained (and the constraint buffer filled) for Recovery RL and for the LR / RSPO /
        # SQRL / RCPO comparisons (experiment.py:407-415, 443-445)
        self.uses_qrisk = self.use_recovery or bool(dgd) or bool(rcpo)
        self.online_qrisk = self.uses_qrisk and not disable_online_updates
        self.scalar_algos = bool(update_nu) or bool(rcpo) or bool(auto_alpha)
        self.gate_pos_fraction = float(pos_fraction)                          # experiment.py:410
        self.pos_fraction = pos_fraction if pos_fraction >= 0 else None        # qrisk.py:77
        self.start_steps = int(start_steps)
        # comparison branches of the acting path (csrc/select.cu): SQRL's action filter (sac.py:139-161), Q-sampling recovery
        # (qrisk.py:214-225), the second push of --add_both_transitions (experiment.py:446-448), --policy Deterministic
        self.constraint_sampling = bool(constraint_sampling)
        self.q_sampling = bool(q_sampling_recovery) and self.use_recovery and not self.mf_recovery
        self.add_both = bool(add_both_transitions) and self.use_recovery
        self.deterministic = bool(deterministic)
        self.safe_samples, self.q_samples = int(safe_samples), int(q_samples)    # sac.py:140, qrisk.py:216 hard-code 100 / 1000
        if self.constraint_sampling and (self.use_recovery or self.deterministic):
            raise ValueError("the SQRL action filter (use_constraint_sampling) replaces the Gaussian task policy's own draw; the "
                             "reference's scripts never combine it with use_recovery or --policy Deterministic")
        if self.q_sampling and mb_recovery:
            raise ValueError("Q_sampling_recovery and the model-based planner are alternatives (experiment.py:568-573)")
        self.relabel = not disable_action_relabeling            # experiment.py:438-441
        self.host_inputs = bool(host_inputs)
        self.act_staging = True                                 # set False to act with the single fused launch
        self.fused = int(use_tensor_cores) >= 2                 # use_tensor_cores 2: tcgen05 + fused update stages
        self.log_outputs = bool(log_outputs) or self.host_inputs
        sc = ACTION_SCALE[env_name]
        # peer_grads: the arena lives in symmetric memory and the optimizer-step kernel sums the ranks' gradient blocks
        # itself over NVLink (one flag barrier per optimizer step) instead of an NCCL all-reduce
        self.peer_arena = None
        if peer_grads and self.world > 1:
            self.peer_arena = dist_utils.PeerArena(self.device, self.rank, self.world, process_group)
        self.peer_error = None
        self.agent = AgentArena(self.device, max_batch=self.B, allocator=self._peer_alloc if self.peer_arena else None, gamma=gamma, alpha=alpha, tau=tau, lr=lr,
                                gamma_safe=gamma_safe, tau_safe=tau_safe, eps_safe=eps_safe,
                                target_update_interval=target_update_interval, mf_recovery=mf_recovery,
                                action_scale=(sc, sc), grad_scale=1.0 / self.world,
                                use_tensor_cores=use_tensor_cores, dgd=dgd, update_nu=update_nu, rcpo=rcpo,
                                auto_alpha=auto_alpha, nu=nu, lambda_rcpo=lambda_rcpo, deterministic=self.deterministic)
        self.cfg = self.agent.cfg
        self.arena = self.agent.arena
        self.counters = self.agent.counters
        self.env_cfg = native.env_config(self.kind, self.n, horizon=HORIZON[env_name],
                                         reward_penalty=constraint_reward_penalty, seed=self.seed, stream_id=self.rank,
                                         maze_substeps=maze_substeps)
        dev = self.device
        n = self.n
        self.task_cap = max(int(replay_size), 2 * n if self.add_both else n)
        self.cons_cap = (max(int(safe_replay_size), n) + 15) // 16 * 16
        self.state = torch.zeros(2, n, dtype=torch.float64, device=dev)
        self.ep_steps = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ep_return = torch.zeros(n, dtype=torch.float64, device=dev)
        self.action_task = torch.zeros(n, 2, device=dev)
        self.qrisk = torch.zeros(n, device=dev)
        # what the host face downloads every step lives in ONE flat device buffer (one D2H copy): per-env next_state / reward /
        # flags / executed action
        self._out_specs = [("next_state", (2, n), torch.float64), ("reward", (n,), torch.float64), ("done", (n,), torch.uint8),
                           ("constraint", (n,), torch.uint8), ("success", (n,), torch.uint8), ("recovery", (n,), torch.uint8),
                           ("action", (n, 2), torch.float32)]
        self._out_flat, ov = _carve(self._out_specs, device=dev)
        self.action_real = ov["action"]
        self.recovery = ov["recovery"]
        self.task_ring = torch.zeros(self.task_cap, 8, device=dev)
        self.cons_ring = torch.zeros(self.cons_cap, 8, device=dev)
        self.cons_flags = torch.zeros(self.cons_cap, dtype=torch.uint8, device=dev)
        # flag chunks of the stratified sampler: at most 8192 per ring (replay.cu), 512-byte granularity
        self.flag_chunk = FLAG_CHUNK * max(1, -(-self.cons_cap // (8192 * FLAG_CHUNK)))
        self.n_chunks = (self.cons_cap + self.flag_chunk - 1) // self.flag_chunk
        self.chunk_counts = torch.zeros(2, self.n_chunks, dtype=torch.int32, device=dev)
        # ONE CPython-compatible MT19937 stream shared by both buffers (replay_memory.py:16,41); every rank
        # gets its own stream (seed + rank) so that shards draw different batches
        self.mt_state = native.mt19937_seed(self.seed + self.rank).to(dev)
        self.losses = torch.zeros(16, device=dev)
        self.sac_sample_cfg = native.sample_config(self.task_cap, self.B, False, None, gate_mode=1)
        self.qr_sample_cfg = native.sample_config(self.cons_cap, self.B, True, self.pos_fraction, gate_mode=2,
                                                  chunk=self.flag_chunk, gate_pos_fraction=self.gate_pos_fraction)
        # per-step outputs (logging schema of experiment.py:421 / run_stats.pkl), only when asked for
        if self.log_outputs:
            self.out_next, self.out_reward, self.out_done = ov["next_state"], ov["reward"], ov["done"]
            self.out_cons, self.out_succ = ov["constraint"], ov["success"]
        else:
            self.out_next = self.out_reward = self.out_done = self.out_cons = self.out_succ = None
        # host-supplied randomness (parity / end-to-end mode): device staging + pinned host buffers
        if self.host_inputs:
            B = self.B
            self._in_specs = [("reset_draws", (2, n), torch.float64)] + \
                ([("env_noise", (2, n), torch.float64)] if self.kind != native.ENV_MAZE else []) + \
                [("eps_task", (n, 2), torch.float32), ("eps_rec", (n, 2), torch.float32), ("rand_u", (n, 2), torch.float32),
                 ("sac_eps_next", (B, 2), torch.float32), ("sac_eps_cur", (B, 2), torch.float32),
                 ("qr_eps_next", (B, 2), torch.float32), ("qr_eps_rec", (B, 2), torch.float32)] + \
                ([("sqrl_eps", (n, self.safe_samples, 2), torch.float32), ("sqrl_u", (n,), torch.float32)]
                 if self.constraint_sampling else []) + \
                ([("qs_u", (n, self.q_samples, 2), torch.float32)] if self.q_sampling else [])
            self._in_flat, self.in_dev = _carve(self._in_specs, device=dev)         # the step's kernels read these views
            if self.kind == native.ENV_MAZE:
                self.in_dev["env_noise"] = None
            self._in_host_set = HostInputs(self._in_specs)
            self.in_host = self._in_host_set.views
            self._out_host_flat, oh = _carve(self._out_specs, pinned=True)
            self.out_host = dict(losses=torch.zeros(16).pin_memory(),
                                 counters=torch.zeros(native.NUM_COUNTERS, dtype=torch.int64).pin_memory(), **oh)
        else:
            self.in_dev = {}
        # model-based recovery (BASELINE config 5): the PETS / CEM planner proposes the recovery action for every env
        # copy (recovery_rl/MPC.py, csrc/mpc.cu); only the copies whose Q_risk exceeds eps_safe use it
        self.mpc = None
        if mb_recovery:
            if not self.use_recovery or self.mf_recovery:
                raise ValueError("mb_recovery needs use_recovery=True and mf_recovery=False")
            from config import create_config
            from .dotmap_lite import DotMap
            from .MPC import MPC
            cc = create_config(env_name, "MPC", DotMap(), [], "").ctrl_cfg
            if mpc_popsize is not None:
                cc.opt_cfg.cfg = dict(cc.opt_cfg.cfg, popsize=int(mpc_popsize),
                                      num_elites=int(mpc_num_elites or max(1, mpc_popsize // 10)))
            self.mpc = MPC(cc, n_envs=n, seed=self.seed, stream_id=self.rank)

            class _VF(object):
                arena = self.agent
            self.mpc.update_value_func(_VF())
            self.mpc.state = self.state                      # the planner reads the env state in place
            self.mpc.counters = self.counters                # Philox step counter
            self.action64 = torch.zeros(n, 2, dtype=torch.float64, device=dev)
        self._sel_ws = None
        if self.constraint_sampling or self.q_sampling:
            # candidate workspace: at most ~8M candidate rows (288 MB) at a time, larger env counts loop over chunks
            k_s = self.safe_samples if self.constraint_sampling else self.q_samples
            chunk = min(n, max(1, (1 << 23) // k_s))
            self._sel_ws = torch.empty(native.select_workspace_floats(chunk, k_s), device=dev)
            self._sel_chunks = -(-n // chunk)
        self.graph = None
        self._side = torch.cuda.Stream(device=dev)
        self._ev_fork = torch.cuda.Event()
        self._ev_join = torch.cuda.Event()
        self._side2 = torch.cuda.Stream(device=dev)           # staged acting (policy / Q_risk stages next to the updates)
        self._ev_act = torch.cuda.Event()
        self._ev_stage = [torch.cuda.Event(), torch.cuda.Event()]
        # SMs a side-stream acting stage may occupy: 148 - the running update kernel's <= 16 CTAs (+ the next update kernel's,
        # resident early under programmatic dependent launch)
        self.stage_ctas = int(os.environ.get("RRL_STAGE_CTAS", "128"))
        self.launches_per_step = 0
        self._grad_views = None

    # ------------------------------------------------------------------------------------------
    def _in(self, name):
        return self.in_dev.get(name) if self.host_inputs else None

    def h2d_bytes_per_step(self):
        return sum(v.numel() * v.element_size() for v in self.in_host.values()) if self.host_inputs else 0

    def d2h_bytes_per_step(self):
        return sum(v.numel() * v.element_size() for v in self.out_host.values()) if self.host_inputs else 0

    def init_agent(self, modules=None):
        """xavier init in the reference's construction order (draws from the torch global RNG), identical on
        every rank when the torch seed is."""
        if modules is None:
            from .model import build_reference_modules
            sc = ACTION_SCALE[self.env_name]
            modules = build_reference_modules(hidden=256, action_scale=(sc, sc), deterministic=self.deterministic)
        self.agent.load_modules(modules)

    def train_mb(self, transitions=None, n_recent=50000, epochs=None):
        """MPC.train for the engine.  transitions: list of (s, a, c, s', mask) demos (experiment.py:298-305), or None:
        the most recent `n_recent` transitions of the constraint ring (stands in for experiment.py:464-478, which
        appends every finished episode; with thousands of env copies the data set is capped to a recent window)."""
        if self.mpc is None:
            return
        if transitions is not None:
            s = np.array([t[0] for t in transitions]); a = np.array([t[1] for t in transitions])
            s2 = np.array([t[3] for t in transitions])
            self.mpc.train(s, a, random=True, next_obs=s2, epochs=50 if epochs is None else epochs)
            return
        c = self.counters.cpu()
        ln, pos = int(c[native.C_CONS_LEN]), int(c[native.C_CONS_POS])
        k = min(int(n_recent), ln)
        if k == 0:
            return
        idx = (pos - 1 - torch.arange(k, device=self.device)) % self.cons_cap
        rec = self.cons_ring[idx].cpu().numpy().astype(np.float64)
        self.mpc.train_in = np.zeros((0, 4)); self.mpc.train_targs = np.zeros((0, 2))     # recent window, not cumulative
        self.mpc.train(rec[:, 0:2], rec[:, 2:4], random=True, next_obs=rec[:, 5:7], epochs=epochs)

    def set_nu(self, nu):
        """the `nu` argument of SAC.update_parameters (experiment.py:406); a device scalar the captured graph reads."""
        self.agent.set_nu_arg(nu)

    def reset(self, draws=None):
        """env.reset() for every env copy (experiment.py:383)."""
        native.env_reset(self.env_cfg, self.state, self.ep_steps, self.ep_return, self.counters, draws=draws)

    def push_offline(self, transitions):
        """pretrain_critic_recovery: recovery_memory.push(*transition) for the demos (experiment.py:277-282)."""
        rec = np.zeros((len(transitions), 8), np.float32)
        for i, t in enumerate(transitions):
            rec[i, 0:2] = t[0]; rec[i, 2:4] = t[1]; rec[i, 4] = float(t[2]); rec[i, 5:7] = t[3]; rec[i, 7] = float(t[4])
        self.push_offline_records(rec)

    def push_offline_records(self, rec):
        rec = np.ascontiguousarray(rec, np.float32)
        n = len(rec)
        if n == 0:
            return
        native.replay_push(self.cons_ring, self.cons_cap, torch.from_numpy(rec).to(self.device), n, self.counters,
                           cons_flags=self.cons_flags)
        self.counters[native.C_OFFLINE_VIOLS] += int((rec[:, 4] != 0).sum())

    # ---- one Q_risk (+ recovery policy) update -----------------------------------------------------
    def _all_reduce(self, net_names):
        if self.world == 1:
            return
        if self.peer_arena is not None:      # every rank's gradients are complete -> apply kernels read them in place
            if not self.fused_barrier:       # (fused: the optimizer-step kernel runs the flag barrier itself)
                native.peer_barrier(self.peer_arena.peers, self.peer_arena.epoch, self.counters)
            return
        if self._grad_views is None:
            self._grad_views = dist_utils.grad_ranges(self.cfg)
        dist_utils.all_reduce_grads(self.arena, self._grad_views, net_names, self.pg)

    def _peer_alloc(self, n):
        """symmetric allocation, agreed on by all ranks: if it fails anywhere every rank keeps the NCCL path."""
        import torch.distributed as dist
        t = None
        try:
            t = self.peer_arena.allocate(n)
        except Exception as ex:            # no symmetric memory on this system / driver
            self.peer_error = repr(ex)
        ok = torch.tensor([0 if t is None else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.pg)
        if int(ok.item()) == 0:
            self.peer_arena = None
            return torch.zeros(n, dtype=torch.float32, device=self.device)
        return t

    @property
    def _peers(self):
        if self.peer_arena is None:
            return None
        return self.peer_arena.peers_fused if self.fused_barrier else self.peer_arena.peers

    @property
    def staged_act(self):
        """acting split into stages that overlap the updates (see _enqueue_step): tcgen05 path, model-free recovery, online
        safety-critic updates (every stage then follows an optimizer step of its own)"""
        return (int(self.cfg.use_tensor_cores) >= 1 and self.use_recovery and self.mf_recovery and self.online_qrisk
                and self.mpc is None and self.act_staging and not self.constraint_sampling)

    @property
    def early_rec_forward(self):
        """the recovery policy's forward pass of the recovery update runs on the sampling side stream (tcgen05 path)"""
        return int(self.cfg.use_tensor_cores) >= 1 and self.mf_recovery and self.online_qrisk and self.act_staging

    @property
    def fused_barrier(self):
        """peer mode on the tcgen05 path: the tiled optimizer-step kernel publishes / waits for the gradient flags itself"""
        return self.peer_arena is not None and int(self.cfg.use_tensor_cores) >= 1

    def _sync_gate_counts(self):
        if self.peer_arena is not None:
            native.peer_sync_gate_counts(self.peer_arena.peers, self.peer_arena.epoch, self.counters, gate_batch=self.B,
                                         gate_pos_fraction=self.gate_pos_fraction)
        elif self.world > 1:
            dist_utils.sync_gate_counts(self.counters, self.pg)

    def _qr_sample(self, sample_cfg=None):
        sc = sample_cfg or self.qr_sample_cfg
        k = 0
        if sc.pos_fraction >= 0:
            native.replay_flag_count(self.cons_flags, self.cons_cap, self.flag_chunk, self.chunk_counts); k += 1
        a = self.agent
        native.replay_sample(sc, self.cons_ring, self.mt_state, self.counters, native.C_QRISK_ROWS, a.scratch("qr_s"),
                             a.scratch("qr_a"), a.scratch("qr_c"), a.scratch("qr_s2"), a.scratch("qr_m"),
                             cons_flags=self.cons_flags, chunk_counts=self.chunk_counts)
        return k + 1

    def _qr_compute(self, between=None, rec_forward_done=False):
        """safety-critic step, then (MF recovery) the recovery-policy step on the post-step critic.  `between`: called
        after the safety-critic optimizer step has been enqueued (staged acting forks its Q_risk stage there).
        rec_forward_done: the recovery policy's forward pass was already enqueued (side stream, next to the SAC update)."""
        cfg, ar, cn = self.cfg, self.arena, self.counters
        native.qrisk_backward(cfg, ar, cn, self.losses[8:], self._in("qr_eps_next"), seed=self.seed, stream_id=self.rank)
        self._all_reduce(["qrisk"])
        native.qrisk_apply(cfg, ar, cn, peers=self._peers)
        if between is not None:
            between()
        if rec_forward_done:
            native.recovery_backward_rest(cfg, ar, cn, self.losses[8:])
        else:
            native.recovery_backward(cfg, ar, cn, self.losses[8:], self._in("qr_eps_rec"), seed=self.seed, stream_id=self.rank)
        self._all_reduce(["recovery"])
        native.recovery_apply(cfg, ar, cn, peers=self._peers)
        nb = 1 if (self.peer_arena is not None and not self.fused_barrier) else 0           # peer barrier kernels
        # kernels launched (gpu_launches bookkeeping).  fused: the loss / sample-backward stages run as kernel tails and the
        # layer-1 backward in the epilogue of the backward GEMM -> fwd, fwd, bwd [, bwd] per update
        qr, rec = (3, 4) if self.fused else (5, 8)
        return qr + 1 + (rec if self.mf_recovery else 0) + 1 + 2 * nb

    def qrisk_update(self, sample_cfg=None):
        return self._qr_sample(sample_cfg) + self._qr_compute()

    def _sac_sample(self):
        a = self.agent
        native.replay_sample(self.sac_sample_cfg, self.task_ring, self.mt_state, self.counters, native.C_SAC_ROWS,
                             a.scratch("sac_s"), a.scratch("sac_a"), a.scratch("sac_r"), a.scratch("sac_s2"),
                             a.scratch("sac_m"))
        return 1

    def _sac_compute(self):
        cfg, ar, cn = self.cfg, self.arena, self.counters
        native.sac_backward(cfg, ar, cn, self.losses, self._in("sac_eps_next"), self._in("sac_eps_cur"), seed=self.seed,
                            stream_id=self.rank)
        self._all_reduce(["critic", "policy"])
        if self.world > 1 and self.scalar_algos:      # gradients of log_alpha (f32) and log_nu / log_lambda (f64)
            f32, f64 = self.agent.scalars()
            dist_utils.all_reduce_sum(f32[native.S_G_LOG_ALPHA:native.S_G_LOG_ALPHA + 1], self.pg)
            dist_utils.all_reduce_sum(f64[native.D_G_LOG_NU:native.D_G_LOG_LAMBDA + 1], self.pg)
        native.sac_apply(cfg, ar, cn, peers=self._peers)
        return (4 if self.fused else 8) + 1 + (1 if self.scalar_algos else 0) + \
            (1 if (self.peer_arena is not None and not self.fused_barrier) else 0)

    def sac_update(self):
        return self._sac_sample() + self._sac_compute()

    def pretrain_qrisk(self, steps, n_demos=None):
        """experiment.py:289-296: critic_safe_pretraining_steps x QRiskWrapper.update_parameters with
        batch_size = min(batch_size, len(constraint_demo_data)); no gate."""
        b = self.B if n_demos is None else min(self.B, int(n_demos))
        if self.peer_arena is not None:      # ranks reach the first flag barrier together (host-side skew stays out of it)
            import torch.distributed as dist
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.pg)
        sc = native.sample_config(self.cons_cap, b, True, self.pos_fraction, gate_mode=0, chunk=self.flag_chunk)
        for _ in range(int(steps)):
            self.qrisk_update(sc)

    # ---- the vector step -------------------------------------------------------------------------
    def _enqueue_step(self):
        k = 0
        main = torch.cuda.current_stream()
        k += self._sac_sample()          # draws first from the shared sampler stream (replay_memory.py:28)
        if self.online_qrisk:
            # the Q_risk batch (flag scan + stratified sample + gather) does not depend on the SAC update: it runs
            # on a side stream, forked after the SAC sample (stream order of the ONE shared generator) and joined
            # before the Q_risk kernels.  Captured as a fork/join in the CUDA graph.
            self._sync_gate_counts()
            k += 1 if self.peer_arena is not None else 0
            self._ev_fork.record(main)
            self._side.wait_event(self._ev_fork)
            with torch.cuda.stream(self._side):
                k += self._qr_sample()
                if self.early_rec_forward:
                    # the recovery policy's forward pass on the Q_risk batch needs neither the SAC nor the safety-critic step
                    native.recovery_forward(self.cfg, self.arena, self.counters, self._in("qr_eps_rec"), seed=self.seed,
                                            stream_id=self.rank)
                self._ev_join.record(self._side)
        k += self._sac_compute()                                                   # experiment.py:397-406
        # Staged acting (tcgen05 path, MF recovery, online safety-critic updates): each stage of the composite action only
        # needs SOME of the networks, so it is enqueued on a second side stream as soon as those have been stepped and runs
        # on a bounded number of SMs NEXT TO the remaining (latency-bound, <= 16-CTA) update kernels:
        #   task policy  after the SAC step        (while the safety critic is updated)
        #   Q_risk       after the safety-critic step (while the recovery policy is updated)
        #   recovery policy + select after the recovery step, on the main stream
        # Same kernels, same arithmetic, same results as the fused launch; only the schedule differs.
        staged = self.staged_act

        def act_stage(stages, max_ctas=0):
            native.agent_act(self.cfg, self.arena, self.n, self.state, self.counters, self.action_task, self.action_real,
                             self.recovery, self.qrisk, self._in("eps_task"), self._in("eps_rec"), self._in("rand_u"),
                             use_recovery=self.use_recovery, start_steps=self.start_steps, seed=self.seed,
                             stream_id=self.rank, stages=stages, max_ctas=max_ctas)                  # experiment.py:419

        def fork_stage(stages, slot):
            ev = self._ev_stage[slot]
            ev.record(main)
            self._side2.wait_event(ev)
            with torch.cuda.stream(self._side2):
                act_stage(stages, self.stage_ctas)
                self._ev_act.record(self._side2)

        if staged:
            fork_stage(native.ACT_STAGE_POLICY, 0)
            k += 1
        if self.online_qrisk:
            main.wait_event(self._ev_join)
            k += self._qr_compute(between=(lambda: fork_stage(native.ACT_STAGE_QRISK, 1)) if staged else None,
                                  rec_forward_done=self.early_rec_forward)   # experiment.py:407-415
            if staged:
                k += 1
        # (the fp16 hi/lo tcgen05 operand images are refreshed by the optimizer-step kernels themselves)
        if staged:
            main.wait_event(self._ev_act)                                          # the Q_risk stage (after the policy stage)
            act_stage(native.ACT_STAGE_RECOVERY)
        elif self.constraint_sampling:                                             # sac.py:139-161 (SQRL)
            native.sqrl_select_action(self.cfg, self.arena, self.n, self.safe_samples, self.state, self.counters, self._sel_ws,
                                      self.action_task, self.action_real, self.recovery, self.qrisk,
                                      eps_cand=self._in("sqrl_eps"), cat_u=self._in("sqrl_u"), rand_u=self._in("rand_u"),
                                      start_steps=self.start_steps, seed=self.seed, stream_id=self.rank)
            k += 4 * self._sel_chunks - 1
        else:
            act_stage(native.ACT_STAGE_ALL)
        if self.q_sampling:                                                        # qrisk.py:214-225
            native.qsample_recovery_action(self.cfg, self.arena, self.n, self.q_samples, self.state, self.counters, self._sel_ws,
                                           self.action_real, recovery=self.recovery, cand_u=self._in("qs_u"), seed=self.seed,
                                           stream_id=self.rank)
            k += 3 * self._sel_chunks
        a64 = None
        if self.mpc is not None:                                                   # experiment.py:568-573
            plan = self.mpc.plan(mask=self.recovery)
            rec = self.recovery.bool().unsqueeze(1)
            self.action64.copy_(torch.where(rec, plan, self.action_task.double()))   # MPC actions stay float64
            self.action_real.copy_(self.action64)
            a64 = self.action64
            k += 2 + 3 * self.mpc.optimizer.max_iters
        native.env_step(self.env_cfg, self.action_task if self.relabel else self.action_real, self.action_real,
                        self.state, self.ep_steps, self.ep_return,
                        self.counters, recovery=self.recovery, noise=self._in("env_noise"),
                        reset_draws=self._in("reset_draws"), task_ring=self.task_ring, task_capacity=self.task_cap,
                        cons_ring=self.cons_ring if self.uses_qrisk else None,
                        cons_flags=self.cons_flags if self.uses_qrisk else None,
                        cons_capacity=self.cons_cap if self.uses_qrisk else 0, out_next_state=self.out_next,
                        out_reward=self.out_reward, out_done=self.out_done, out_constraint=self.out_cons,
                        out_success=self.out_succ, action_f64=a64)                 # experiment.py:420-461
        native.counters_advance(self.counters, self.n, self.task_cap, self.cons_cap, True, self.uses_qrisk)
        if self.add_both:                                                          # experiment.py:446-448
            native.replay_push_both(self.task_ring, self.task_cap, self.n, self.recovery, self.action_real, self.counters)
            k += 1
        self.launches_per_step = k + 3
        return self.launches_per_step

    def step(self):
        """one vector step, eagerly enqueued."""
        return self._enqueue_step()

    def capture(self):
        """warm up (lazy cudaFuncSetAttribute calls must happen outside capture) and record the step."""
        saved = self.snapshot()
        self._enqueue_step()
        torch.cuda.synchronize()
        self.restore(saved)
        self.graph = self._try_capture()
        if self.graph is None and native.set_pdl(False):
            # programmatic dependent launches could not be captured on this driver: stream-ordered launches instead
            self.pdl_disabled = self.capture_error
            self.restore(saved)
            self.graph = self._try_capture()
            if self.graph is None:
                native.set_pdl(True)
        self.restore(saved)
        return self.graph

    def _try_capture(self):
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue_step()
            return g
        except Exception as ex:     # e.g. a collective that cannot be captured: keep the eager path
            self.capture_error = repr(ex)
            torch.cuda.synchronize()
            return None

    def replay(self):
        if self.graph is None:
            self._enqueue_step()
        else:
            self.graph.replay()

    # ---- host-buffer step (end-to-end face) --------------------------------------------------------
    def step_host(self, inputs):
        """inputs: dict of host arrays for this step's random draws (see self.in_host).  Uploads them from
        pinned memory, runs the step, downloads losses / counters / per-env outputs into pinned buffers."""
        assert self.host_inputs
        src = self._stage_host(inputs, self._in_host_set)
        self._in_flat.copy_(src.flat, non_blocking=True)                 # ONE H2D copy for all of the step's inputs
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue_step()
        o = self.out_host
        self._out_host_flat.copy_(self._out_flat, non_blocking=True)     # ONE D2H copy for the per-env results
        o["losses"].copy_(self.losses, non_blocking=True)
        o["counters"].copy_(self.counters, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return o

    def new_host_inputs(self):
        """a pinned, flat input set for this engine (HostInputs): fill `.views[name]` in place, hand it to step_host / submit."""
        assert self.host_inputs
        return HostInputs(self._in_specs)

    def _stage_host(self, inputs, slot):
        """inputs -> a HostInputs set ready for one flat upload: a HostInputs of this engine is used as is (zero-copy);
        a dict of arrays / tensors is packed into the pinned set `slot` (missing keys keep the slot's previous contents)."""
        if isinstance(inputs, HostInputs):
            assert inputs.flat.numel() == slot.flat.numel(), "HostInputs of another engine"
            return inputs
        if inputs is not None:
            for k, h in slot.views.items():
                if k in inputs:
                    x = inputs[k]
                    h.copy_(x if torch.is_tensor(x) else torch.as_tensor(x).reshape(h.shape))
        return slot

    # ---- pipelined host-buffer steps: copies of step k overlap the compute of its neighbours -----------------
    def enable_pipeline(self):
        """Two-slot staging on both sides of the step so that the H2D copy of step k+1 and the D2H copy of step k-1 run
        on their own streams (copy engines) while step k computes.  submit(inputs) -> ticket; collect(ticket) -> the
        pinned outputs of that step.  A ticket must be collected before the second-next submit (slot reuse).  Every set is
        ONE flat buffer: per step one H2D + one D2D on the way in, one D2D + one D2H (+ losses, counters) on the way out."""
        assert self.host_inputs
        dev = self.device
        self._pl_in_dev = [torch.empty_like(self._in_flat) for _ in range(2)]
        self._pl_in_host = [HostInputs(self._in_specs) for _ in range(2)]
        self._pl_out_dev = [torch.empty_like(self._out_flat) for _ in range(2)]
        self._pl_small_dev = [(torch.empty_like(self.losses), torch.empty_like(self.counters)) for _ in range(2)]
        self._pl_out_host = []
        for _ in range(2):
            flat, views = _carve(self._out_specs, pinned=True)
            views = dict(views, losses=torch.zeros(16).pin_memory(),
                         counters=torch.zeros(native.NUM_COUNTERS, dtype=torch.int64).pin_memory())
            self._pl_out_host.append((flat, views))
        self._s_h2d = torch.cuda.Stream(device=dev)
        self._s_d2h = torch.cuda.Stream(device=dev)
        mk = lambda: [torch.cuda.Event(), torch.cuda.Event()]
        self._ev_in_ready, self._ev_in_free, self._ev_out_ready, self._ev_out_done = mk(), mk(), mk(), mk()
        self._pl_k = 0

    def submit(self, inputs):
        k = self._pl_k
        b = k & 1
        main = torch.cuda.current_stream()
        if not isinstance(inputs, HostInputs) and k >= 2:
            self._ev_in_ready[b].synchronize()               # the H2D that last read this pinned slot has finished
        src = self._stage_host(inputs, self._pl_in_host[b])
        with torch.cuda.stream(self._s_h2d):
            if k >= 2:
                self._s_h2d.wait_event(self._ev_in_free[b])  # step k-2 has consumed this device slot
            self._pl_in_dev[b].copy_(src.flat, non_blocking=True)
            self._ev_in_ready[b].record(self._s_h2d)
        main.wait_event(self._ev_in_ready[b])
        self._in_flat.copy_(self._pl_in_dev[b], non_blocking=True)
        self._ev_in_free[b].record(main)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue_step()
        if k >= 2:
            main.wait_event(self._ev_out_done[b])            # step k-2's D2H has drained this device slot
        self._pl_out_dev[b].copy_(self._out_flat, non_blocking=True)
        self._pl_small_dev[b][0].copy_(self.losses, non_blocking=True)
        self._pl_small_dev[b][1].copy_(self.counters, non_blocking=True)
        self._ev_out_ready[b].record(main)
        with torch.cuda.stream(self._s_d2h):
            self._s_d2h.wait_event(self._ev_out_ready[b])
            flat, views = self._pl_out_host[b]
            flat.copy_(self._pl_out_dev[b], non_blocking=True)
            views["losses"].copy_(self._pl_small_dev[b][0], non_blocking=True)
            views["counters"].copy_(self._pl_small_dev[b][1], non_blocking=True)
            self._ev_out_done[b].record(self._s_d2h)
        self._pl_k = k + 1
        return k

    def collect(self, ticket):
        assert self._pl_k - 2 <= ticket < self._pl_k, "ticket already overwritten or not submitted"
        self._ev_out_done[ticket & 1].synchronize()
        return self._pl_out_host[ticket & 1][1]

    # ---- evaluation rollouts (experiment.py:372-374, 493-538) ---------------------------------------------
    def eval_rollout(self, n_eval=1):
        """get_test_rollout for `n_eval` fresh env copies at once: the task policy acts in eval mode (its mean action,
        sac.py:166-167), the Q_risk threshold and the recovery policy are the training ones (experiment.py:546-577 with
        train=False), nothing is pushed to the replay rings, no update runs and none of the training state (env copies,
        counters, sampler, Philox step) is touched: the rollout has its own env state and its own counter block.
        Returns a list of episodes, each a list of per-step info dicts (the reference's test_stats schema)."""
        k = int(n_eval)
        dev = self.device
        ev = getattr(self, "_eval", None)
        if ev is None or ev["k"] != k:
            ev = dict(k=k, state=torch.zeros(2, k, dtype=torch.float64, device=dev), ep_steps=torch.zeros(k, dtype=torch.int32, device=dev),
                      ep_return=torch.zeros(k, dtype=torch.float64, device=dev), a_task=torch.zeros(k, 2, device=dev),
                      a_real=torch.zeros(k, 2, device=dev), rec=torch.zeros(k, dtype=torch.uint8, device=dev),
                      q=torch.zeros(k, device=dev), counters=torch.zeros(native.NUM_COUNTERS, dtype=torch.int64, device=dev),
                      nxt=torch.zeros(2, k, dtype=torch.float64, device=dev), rew=torch.zeros(k, dtype=torch.float64, device=dev),
                      done=torch.zeros(k, dtype=torch.uint8, device=dev), cons=torch.zeros(k, dtype=torch.uint8, device=dev),
                      succ=torch.zeros(k, dtype=torch.uint8, device=dev), calls=0,
                      cfg=native.env_config(self.kind, k, horizon=HORIZON[self.env_name], reward_penalty=0.0,
                                            seed=self.seed + 7919, stream_id=self.rank,
                                            maze_substeps=self.env_cfg.maze_substeps))
            self._eval = ev
        ev["counters"].zero_()
        ev["counters"][native.C_VEC_STEP] = 1000003 * ev["calls"]      # fresh Philox draws for every evaluation
        ev["calls"] += 1
        native.env_reset(ev["cfg"], ev["state"], ev["ep_steps"], ev["ep_return"], ev["counters"])
        episodes = [[] for _ in range(k)]
        alive = np.ones(k, bool)
        for _ in range(HORIZON[self.env_name] + 1):
            prev = ev["state"].t().cpu().numpy().copy()
            if self.constraint_sampling:       # sac.py:139-161 ignores `eval`: the filter samples in test rollouts too
                native.sqrl_select_action(self.cfg, self.arena, k, self.safe_samples, ev["state"], ev["counters"], self._sel_ws,
                                          ev["a_task"], ev["a_real"], ev["rec"], ev["q"], start_steps=0, seed=self.seed + 7919,
                                          stream_id=self.rank)
            else:
                native.agent_act(self.cfg, self.arena, k, ev["state"], ev["counters"], ev["a_task"], ev["a_real"], ev["rec"], ev["q"],
                                 use_recovery=self.use_recovery and self.mpc is None, eval=True, start_steps=0,
                                 seed=self.seed + 7919, stream_id=self.rank)
            if self.q_sampling:
                native.qsample_recovery_action(self.cfg, self.arena, k, self.q_samples, ev["state"], ev["counters"], self._sel_ws,
                                               ev["a_real"], recovery=ev["rec"], seed=self.seed + 7919, stream_id=self.rank)
            native.env_step(ev["cfg"], ev["a_task"], ev["a_real"], ev["state"], ev["ep_steps"], ev["ep_return"], ev["counters"],
                            recovery=ev["rec"], out_next_state=ev["nxt"], out_reward=ev["rew"], out_done=ev["done"],
                            out_constraint=ev["cons"], out_success=ev["succ"])
            native.counters_advance(ev["counters"], k, 1, 1, False, False)
            ns, rw = ev["nxt"].t().cpu().numpy(), ev["rew"].cpu().numpy()
            dn, cs, su = ev["done"].cpu().numpy(), ev["cons"].cpu().numpy(), ev["succ"].cpu().numpy()
            ac, rc = ev["a_real"].cpu().numpy(), ev["rec"].cpu().numpy()
            for i in np.flatnonzero(alive):
                episodes[i].append({"constraint": int(cs[i]), "reward": float(rw[i]), "state": prev[i], "next_state": ns[i].copy(),
                                    "action": ac[i].copy(), "success": bool(su[i]), "recovery": bool(rc[i])})
            alive &= ~dn.astype(bool)
            if not alive.any():
                break
        return episodes

    # ---- state snapshots (capture warm-up, tests) -----------------------------------------------------
    def snapshot(self):
        """everything one vector step mutates, incl. the ring slots its pushes will overwrite (they hold live
        transitions once a ring has wrapped, e.g. after a resume)."""
        c = self.counters.cpu()
        ar = torch.arange(self.n, device=self.device)
        ar_t = torch.arange(2 * self.n, device=self.device) if self.add_both else ar     # + the add_both_transitions rows
        t_idx = (int(c[native.C_TASK_POS]) + ar_t) % self.task_cap
        c_idx = (int(c[native.C_CONS_POS]) + ar) % self.cons_cap
        return dict(arena=self.arena.clone(), counters=self.counters.clone(), state=self.state.clone(),
                    ep_steps=self.ep_steps.clone(), ep_return=self.ep_return.clone(), mt=self.mt_state.clone(),
                    flags=self.cons_flags.clone(), t_idx=t_idx, c_idx=c_idx, t_rows=self.task_ring[t_idx].clone(),
                    c_rows=self.cons_ring[c_idx].clone())

    def restore(self, s):
        self.arena.copy_(s["arena"]); self.counters.copy_(s["counters"]); self.state.copy_(s["state"])
        self.ep_steps.copy_(s["ep_steps"]); self.ep_return.copy_(s["ep_return"]); self.mt_state.copy_(s["mt"])
        self.cons_flags.copy_(s["flags"])
        self.task_ring[s["t_idx"]] = s["t_rows"]
        self.cons_ring[s["c_idx"]] = s["c_rows"]

    def save(self, path):
        """checkpoint (networks with the reference's state_dict names, Adam, multipliers, replay, sampler, env state)."""
        from . import checkpoint
        return checkpoint.save(path, self)

    def load(self, path):
        from . import checkpoint
        return checkpoint.load(path, self)

    def read_counters(self):
        c = self.counters.cpu().numpy()
        names = dict(total_numsteps=native.C_TOTAL_NUMSTEPS, episodes=native.C_EPISODES, num_viols=native.C_NUM_VIOLS,
                     num_successes=native.C_NUM_SUCCESSES, viol_and_recovery=native.C_VIOL_RECOVERY,
                     viol_and_no_recovery=native.C_VIOL_NO_RECOV, offline_viols=native.C_OFFLINE_VIOLS,
                     vec_steps=native.C_VEC_STEP, sac_updates=native.C_SAC_UPDATES, qrisk_updates=native.C_QRISK_UPDATES,
                     task_len=native.C_TASK_LEN, cons_len=native.C_CONS_LEN, error=native.C_ERROR)
        out = {k: int(c[v]) for k, v in names.items()}
        out["return_sum"] = float(c[native.C_RETURN_SUM_BITS:native.C_RETURN_SUM_BITS + 1].view(np.float64)[0])
        return out
