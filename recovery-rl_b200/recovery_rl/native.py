"""ctypes binding of librrl.so (include/rrl.h) -- the only door from Python to the CUDA kernels.

There is NO CPU fallback: if the library cannot be loaded, or a call is made without a CUDA device,
this module raises.  Tensors are passed as raw device pointers (torch is only the allocator and the
stream owner); every call enqueues on torch's current stream unless a stream is given.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "librrl.so")

ENV_NAV1, ENV_NAV2, ENV_MAZE = 0, 1, 2
ENV_KIND = {"navigation1": ENV_NAV1, "navigation2": ENV_NAV2, "maze": ENV_MAZE}

# counter indices (include/rrl.h)
C_TOTAL_NUMSTEPS, C_EPISODES, C_NUM_VIOLS, C_NUM_SUCCESSES, C_VIOL_RECOVERY, C_VIOL_NO_RECOV = range(6)
C_OFFLINE_VIOLS, C_VEC_STEP, C_SAC_UPDATES, C_QRISK_UPDATES = 6, 7, 8, 9
C_TASK_POS, C_TASK_LEN, C_CONS_POS, C_CONS_LEN, C_SAC_ROWS, C_QRISK_ROWS, C_ADAM_T0 = 10, 11, 12, 13, 14, 15, 16
C_EXT_VIOLS, C_RETURN_SUM_BITS, C_ERROR, NUM_COUNTERS = 20, 21, 22, 32
C_ADAM_T_ALPHA, C_ADAM_T_NU, C_ADAM_T_LAMBDA = 23, 24, 25
# comparison-algorithm branches (rrl_agent_config_t.algo_flags) and the scalar block ("scalars" scratch region)
ALGO_DGD, ALGO_UPDATE_NU, ALGO_RCPO, ALGO_AUTO_ALPHA, ALGO_DETERMINISTIC = 1, 2, 4, 8, 16
S_ALPHA, S_NU_ARG, S_LOG_ALPHA, S_G_LOG_ALPHA, S_M_ALPHA, S_V_ALPHA, S_ALPHA_LOSS, S_F64_BASE = 0, 1, 2, 3, 4, 5, 6, 8
D_G_LOG_NU, D_G_LOG_LAMBDA, D_LOG_NU, D_M_NU, D_V_NU, D_LOG_LAMBDA, D_M_LAMBDA, D_V_LAMBDA, D_LAMBDA, D_NU_LEARNED = range(10)

NET_CRITIC, NET_CRITIC_TARGET, NET_POLICY, NET_QRISK, NET_QRISK_TARGET, NET_RECOVERY = range(6)
NUM_NETS = 6
NET_NAMES = ["critic", "critic_target", "policy", "qrisk", "qrisk_target", "recovery"]


class EnvConfig(C.Structure):
    _fields_ = [("kind", C.c_int32), ("horizon", C.c_int32), ("n_envs", C.c_int64),
                ("reward_penalty", C.c_double), ("seed", C.c_uint64), ("stream_id", C.c_int32),
                ("maze_substeps", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32)]


class SampleConfig(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("batch_size", C.c_int32), ("is_constraint", C.c_int32),
                ("pos_fraction", C.c_double), ("gate_mode", C.c_int32), ("chunk", C.c_int32),
                ("gate_pos_fraction", C.c_double)]


class AgentConfig(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("max_batch", C.c_int32),
                ("gamma", C.c_float), ("alpha", C.c_float), ("tau", C.c_float),
                ("gamma_safe", C.c_float), ("tau_safe", C.c_float), ("eps_safe", C.c_float),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float),
                ("action_scale", C.c_float * 2), ("action_bias", C.c_float * 2),
                ("target_update_interval", C.c_int32), ("mf_recovery", C.c_int32),
                ("grad_scale", C.c_float), ("use_tensor_cores", C.c_int32),
                ("algo_flags", C.c_int32), ("target_entropy", C.c_float), ("nu", C.c_double), ("lambda_rcpo", C.c_double),
                ("lr64", C.c_double)]


class RRLError(RuntimeError):
    pass


_lib = None


def lib():
    """Load librrl.so once.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RRLError("librrl.so not found at %s -- run `python recovery-rl_b200/build.py` "
                           "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.rrl_last_error.restype = C.c_char_p
        if hasattr(_lib, "rrl_agent_arena_floats"):
            _lib.rrl_agent_arena_floats.restype = C.c_int64
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RRLError("%s failed (%d): %s" % (what, rc, lib().rrl_last_error().decode()))


def require_cuda():
    if not torch.cuda.is_available():
        raise RRLError("no CUDA device: the B200 path has no CPU fallback")


_DT = {"f64": torch.float64, "f32": torch.float32, "i32": torch.int32, "i64": torch.int64,
       "u8": torch.uint8, "u32": torch.int32}


def p(t, kind=None):
    """device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise RRLError("expected a contiguous CUDA tensor, got %s %s" % (t.device, t.shape))
    if kind is not None and t.dtype != _DT[kind] and not (kind == "u8" and t.dtype == torch.bool):
        raise RRLError("expected dtype %s, got %s" % (kind, t.dtype))
    return C.c_void_p(t.data_ptr())


def _stream(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def version():
    return lib().rrl_version()


def set_pdl(enabled):
    """programmatic dependent launch of the step's kernels (include/rrl.h); returns the previous setting"""
    return bool(lib().rrl_set_pdl(int(bool(enabled))))


def debug_opt_times():
    """diagnostics of the optimizer-step kernels since the previous call (include/rrl.h): dict of microseconds / counts"""
    out = (C.c_uint64 * 8)()
    _check(lib().rrl_debug_opt_times(out), "rrl_debug_opt_times")
    return dict(cta0_barrier_us=out[0] * 1e-3, cta0_grad_loads_us=out[1] * 1e-3, cta0_rest_us=out[2] * 1e-3, launches=int(out[3]),
                flag_wait_us=out[4] * 1e-3, barriers=int(out[5]), kernels_us=out[6] * 1e-3)


def pdl_enabled():
    return bool(lib().rrl_set_pdl(-1))


# ------------------------------------------------------------------------------------------------
# environments
# ------------------------------------------------------------------------------------------------
ENV_NO_AUTO_RESET = 1


def env_config(kind, n_envs, horizon=100, reward_penalty=0.0, seed=0, stream_id=0, maze_substeps=500,
               auto_reset=True):
    return EnvConfig(kind, horizon, n_envs, float(reward_penalty), seed & 0xFFFFFFFFFFFFFFFF, stream_id,
                     maze_substeps, 0 if auto_reset else ENV_NO_AUTO_RESET, 0)


def env_reset(cfg, state, ep_steps=None, ep_return=None, counters=None, mask=None, draws=None, stream=None):
    _check(lib().rrl_env_reset(C.byref(cfg), p(mask, "u8"), p(draws, "f64"), p(state, "f64"), p(ep_steps, "i32"),
                               p(ep_return, "f64"), p(counters, "i64"), _stream(stream)), "rrl_env_reset")


def env_step(cfg, action_task, action_real, state, ep_steps, ep_return, counters, recovery=None, noise=None,
             reset_draws=None, task_ring=None, task_capacity=0, cons_ring=None, cons_flags=None, cons_capacity=0,
             out_next_state=None, out_reward=None, out_done=None, out_constraint=None, out_success=None,
             action_f64=None, stream=None):
    _check(lib().rrl_env_step(C.byref(cfg), p(action_task, "f32"), p(action_real, "f32"), p(recovery, "u8"),
                              p(noise, "f64"), p(reset_draws, "f64"), p(state, "f64"), p(ep_steps, "i32"),
                              p(ep_return, "f64"), p(task_ring, "f32"), C.c_int64(task_capacity),
                              p(cons_ring, "f32"), p(cons_flags, "u8"), C.c_int64(cons_capacity), p(counters, "i64"),
                              p(out_next_state, "f64"), p(out_reward, "f64"), p(out_done, "u8"),
                              p(out_constraint, "u8"), p(out_success, "u8"), p(action_f64, "f64"), _stream(stream)),
           "rrl_env_step")


def counters_advance(counters, n, task_capacity, cons_capacity, push_task=True, push_cons=True, stream=None):
    _check(lib().rrl_counters_advance(p(counters, "i64"), C.c_int64(n), C.c_int64(task_capacity),
                                      C.c_int64(cons_capacity), int(push_task), int(push_cons), _stream(stream)),
           "rrl_counters_advance")


# ------------------------------------------------------------------------------------------------
# replay
# ------------------------------------------------------------------------------------------------
def mt19937_seed(seed):
    """CPython random.seed(int) -> uint32[625] state (host tensor, int32 storage)."""
    a = abs(int(seed))
    limbs = []
    while True:
        limbs.append(a & 0xFFFFFFFF)
        a >>= 32
        if a == 0:
            break
    key = (C.c_uint32 * len(limbs))(*limbs)
    out = (C.c_uint32 * 625)()
    _check(lib().rrl_mt19937_seed_host(key, len(limbs), out), "rrl_mt19937_seed_host")
    return torch.from_numpy(np.ctypeslib.as_array(out).copy().view(np.int32))


def replay_push(ring, capacity, rec, n, counters, cons_flags=None, stream=None):
    _check(lib().rrl_replay_push(p(ring, "f32"), p(cons_flags, "u8"), C.c_int64(capacity), p(rec, "f32"),
                                 C.c_int64(n), p(counters, "i64"), 1 if cons_flags is not None else 0,
                                 _stream(stream)), "rrl_replay_push")


def replay_flag_count(cons_flags, capacity, chunk, chunk_counts, stream=None):
    _check(lib().rrl_replay_flag_count(p(cons_flags, "u8"), C.c_int64(capacity), C.c_int32(chunk),
                                       p(chunk_counts, "i32"), _stream(stream)), "rrl_replay_flag_count")


def sample_config(capacity, batch_size, is_constraint=False, pos_fraction=None, gate_mode=0, chunk=4096,
                  gate_pos_fraction=-1.0):
    return SampleConfig(capacity, batch_size, int(is_constraint), -1.0 if pos_fraction is None else float(pos_fraction),
                        gate_mode, chunk, float(gate_pos_fraction))


def replay_sample(cfg, ring, mt_state, counters, rows_counter, out_s, out_a, out_r, out_s2, out_m, out_idx=None,
                  cons_flags=None, chunk_counts=None, stream=None):
    _check(lib().rrl_replay_sample(C.byref(cfg), p(ring, "f32"), p(cons_flags, "u8"), p(chunk_counts, "i32"),
                                   p(mt_state, "u32"), p(counters, "i64"), int(rows_counter), p(out_idx, "i64"),
                                   p(out_s, "f32"), p(out_a, "f32"), p(out_r, "f32"), p(out_s2, "f32"),
                                   p(out_m, "f32"), _stream(stream)), "rrl_replay_sample")


# ------------------------------------------------------------------------------------------------
# agent
# ------------------------------------------------------------------------------------------------
def agent_config(hidden=256, max_batch=256, gamma=0.99, alpha=0.2, tau=0.005, gamma_safe=0.5, tau_safe=0.0002,
                 eps_safe=0.1, lr=3e-4, action_scale=(1.0, 1.0), action_bias=(0.0, 0.0), target_update_interval=1,
                 mf_recovery=True, grad_scale=1.0, use_tensor_cores=0, dgd=False, update_nu=False, rcpo=False,
                 auto_alpha=False, deterministic=False, nu=0.01, lambda_rcpo=0.01, target_entropy=-2.0):
    c = AgentConfig()
    c.hidden, c.max_batch = hidden, max_batch
    c.gamma, c.alpha, c.tau = gamma, alpha, tau
    c.gamma_safe, c.tau_safe, c.eps_safe = gamma_safe, tau_safe, eps_safe
    c.lr, c.beta1, c.beta2, c.adam_eps = lr, 0.9, 0.999, 1e-8
    c.action_scale[0], c.action_scale[1] = float(action_scale[0]), float(action_scale[1])
    c.action_bias[0], c.action_bias[1] = float(action_bias[0]), float(action_bias[1])
    c.target_update_interval = target_update_interval
    c.mf_recovery = int(bool(mf_recovery))
    c.grad_scale = grad_scale
    c.use_tensor_cores = int(use_tensor_cores)
    c.algo_flags = (ALGO_DGD * bool(dgd) | ALGO_UPDATE_NU * bool(update_nu) | ALGO_RCPO * bool(rcpo) |
                    ALGO_AUTO_ALPHA * (bool(auto_alpha) and not deterministic) | ALGO_DETERMINISTIC * bool(deterministic))
    c.target_entropy = float(target_entropy)
    c.nu, c.lambda_rcpo, c.lr64 = float(nu), float(lambda_rcpo), float(lr)
    return c


def agent_arena_floats(cfg):
    return int(lib().rrl_agent_arena_floats(C.byref(cfg)))


def agent_num_tensors(net):
    return int(lib().rrl_agent_num_tensors(int(net)))


def agent_tensor_info(cfg, net, tensor):
    off, rows, cols = C.c_int64(), C.c_int64(), C.c_int64()
    _check(lib().rrl_agent_tensor_info(C.byref(cfg), int(net), int(tensor), C.byref(off), C.byref(rows),
                                       C.byref(cols)), "rrl_agent_tensor_info")
    return off.value, rows.value, cols.value


def agent_grad_range(cfg, net):
    off, cnt = C.c_int64(), C.c_int64()
    _check(lib().rrl_agent_grad_range(C.byref(cfg), int(net), C.byref(off), C.byref(cnt)), "rrl_agent_grad_range")
    return off.value, cnt.value


def agent_scratch_info(cfg, name):
    off, cnt = C.c_int64(), C.c_int64()
    _check(lib().rrl_agent_scratch_info(C.byref(cfg), name.encode(), C.byref(off), C.byref(cnt)),
           "rrl_agent_scratch_info(%s)" % name)
    return off.value, cnt.value


def agent_init_scalars(cfg, arena, stream=None):
    _check(lib().rrl_agent_init_scalars(C.byref(cfg), p(arena, "f32"), _stream(stream)), "rrl_agent_init_scalars")


def agent_refresh(cfg, arena, stream=None):
    _check(lib().rrl_agent_refresh(C.byref(cfg), p(arena, "f32"), _stream(stream)), "rrl_agent_refresh")


def agent_tc_refresh(cfg, arena, stream=None):
    _check(lib().rrl_agent_tc_refresh(C.byref(cfg), p(arena, "f32"), _stream(stream)), "rrl_agent_tc_refresh")


ACT_STAGE_POLICY, ACT_STAGE_QRISK, ACT_STAGE_RECOVERY, ACT_STAGE_ALL = 1, 2, 4, 7


def agent_act(cfg, arena, n, state, counters, action_task, action_real, recovery=None, qrisk_out=None,
              eps_task=None, eps_rec=None, rand_u=None, use_recovery=True, eval=False, start_steps=0, seed=0,
              stream_id=0, stream=None, stages=ACT_STAGE_ALL, max_ctas=0):
    """stages: ACT_STAGE_* bits (all: the fused composite action); max_ctas: SMs the launch may occupy (0: all)."""
    _check(lib().rrl_agent_act_stage(C.byref(cfg), p(arena, "f32"), C.c_int64(n), p(state, "f64"), p(eps_task, "f32"),
                                     p(eps_rec, "f32"), p(rand_u, "f32"), int(use_recovery), int(eval),
                                     C.c_int64(start_steps), C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.c_int32(stream_id),
                                     p(counters, "i64"), p(action_task, "f32"), p(action_real, "f32"), p(recovery, "u8"),
                                     p(qrisk_out, "f32"), int(stages), int(max_ctas), _stream(stream)), "rrl_agent_act_stage")


def select_workspace_floats(chunk_envs, samples):
    lib().rrl_select_workspace_floats.restype = C.c_int64
    return int(lib().rrl_select_workspace_floats(C.c_int64(chunk_envs), C.c_int32(samples)))


def sqrl_select_action(cfg, arena, n, samples, state, counters, workspace, action_task, action_real, recovery=None,
                       qrisk_out=None, eps_cand=None, cat_u=None, rand_u=None, start_steps=0, seed=0, stream_id=0, stream=None):
    """SAC.select_action with --use_constraint_sampling (sac.py:139-161) for n env copies (include/rrl.h)."""
    _check(lib().rrl_sqrl_select_action(C.byref(cfg), p(arena, "f32"), C.c_int64(n), C.c_int32(samples), p(state, "f64"),
                                        p(eps_cand, "f32"), p(cat_u, "f32"), p(rand_u, "f32"), C.c_int64(start_steps),
                                        C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.c_int32(stream_id), p(counters, "i64"),
                                        p(workspace, "f32"), C.c_int64(workspace.numel()), p(action_task, "f32"),
                                        p(action_real, "f32"), p(recovery, "u8"), p(qrisk_out, "f32"), _stream(stream)),
           "rrl_sqrl_select_action")


def qsample_recovery_action(cfg, arena, n, samples, state, counters, workspace, action_real, recovery=None, cand_u=None,
                            seed=0, stream_id=0, stream=None):
    """QRiskWrapper.select_action with --Q_sampling_recovery (qrisk.py:214-225) for the env copies whose flag is set."""
    _check(lib().rrl_qsample_recovery_action(C.byref(cfg), p(arena, "f32"), C.c_int64(n), C.c_int32(samples), p(state, "f64"),
                                             p(cand_u, "f32"), p(recovery, "u8"), C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF),
                                             C.c_int32(stream_id), p(counters, "i64"), p(workspace, "f32"),
                                             C.c_int64(workspace.numel()), p(action_real, "f32"), _stream(stream)),
           "rrl_qsample_recovery_action")


def replay_push_both(task_ring, task_capacity, n, recovery, action_real, counters, stream=None):
    """--add_both_transitions (experiment.py:446-448), after env_step + counters_advance."""
    _check(lib().rrl_replay_push_both(p(task_ring, "f32"), C.c_int64(task_capacity), C.c_int64(n), p(recovery, "u8"),
                                      p(action_real, "f32"), p(counters, "i64"), _stream(stream)), "rrl_replay_push_both")


def _upd(fn, name):
    def call(cfg, arena, counters, losses, eps_a=None, eps_b=None, seed=0, stream_id=0, stream=None):
        args = [C.byref(cfg), p(arena, "f32"), p(eps_a, "f32")]
        if name == "rrl_sac_backward":
            args.append(p(eps_b, "f32"))
        args += [C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.c_int32(stream_id), p(counters, "i64"), p(losses, "f32"),
                 _stream(stream)]
        _check(getattr(lib(), name)(*args), name)
    return call


sac_backward = _upd(None, "rrl_sac_backward")
qrisk_backward = _upd(None, "rrl_qrisk_backward")
recovery_backward = _upd(None, "rrl_recovery_backward")


def recovery_forward(cfg, arena, counters, eps_rec=None, seed=0, stream_id=0, stream=None):
    """the recovery policy's own forward pass of rrl_recovery_backward (may run next to the other updates)"""
    _check(lib().rrl_recovery_forward(C.byref(cfg), p(arena, "f32"), p(eps_rec, "f32"), C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF),
                                      C.c_int32(stream_id), p(counters, "i64"), _stream(stream)), "rrl_recovery_forward")


def recovery_backward_rest(cfg, arena, counters, losses, stream=None):
    _check(lib().rrl_recovery_backward_rest(C.byref(cfg), p(arena, "f32"), p(counters, "i64"), p(losses, "f32"), _stream(stream)),
           "rrl_recovery_backward_rest")


class Peers(C.Structure):
    """rrl_peers_t: the ranks' arenas and signal pads as mapped in this process (symmetric memory)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("arena", C.c_uint64 * 8), ("signal", C.c_uint64 * 8),
                ("epoch", C.c_uint64), ("mc_arena", C.c_uint64)]


def make_peers(rank, arena_ptrs, signal_ptrs, epoch_ptr=0, mc_ptr=0):
    """epoch_ptr != 0: the optimizer-step kernels run the flag barrier themselves (rrl.h, rrl_peers_t::epoch);
    mc_ptr != 0: and read the gradient sum through the multicast mapping (multimem.ld_reduce, rrl_peers_t::mc_arena)."""
    if not (1 <= len(arena_ptrs) <= 8 and len(arena_ptrs) == len(signal_ptrs)):
        raise RRLError("peer mode supports 1..8 ranks of one node")
    P = Peers()
    P.world, P.rank = len(arena_ptrs), int(rank)
    P.epoch = int(epoch_ptr)
    P.mc_arena = int(mc_ptr)
    for r, (a, s) in enumerate(zip(arena_ptrs, signal_ptrs)):
        P.arena[r], P.signal[r] = int(a), int(s)
    return P


def peer_barrier(peers, epoch, counters, stream=None):
    _check(lib().rrl_peer_barrier(C.byref(peers), p(epoch, "i64"), p(counters, "i64"), _stream(stream)), "rrl_peer_barrier")


def peer_sync_gate_counts(peers, epoch, counters, gate_batch=0, gate_pos_fraction=0.0, stream=None):
    _check(lib().rrl_peer_sync_gate_counts(C.byref(peers), p(epoch, "i64"), p(counters, "i64"), C.c_int32(int(gate_batch)),
                                           C.c_double(float(gate_pos_fraction)), _stream(stream)),
           "rrl_peer_sync_gate_counts")


def _apply(name):
    def call(cfg, arena, counters, stream=None, peers=None):
        if peers is None:
            _check(getattr(lib(), name)(C.byref(cfg), p(arena, "f32"), p(counters, "i64"), _stream(stream)), name)
        else:
            _check(getattr(lib(), name + "_p2p")(C.byref(cfg), p(arena, "f32"), p(counters, "i64"), C.byref(peers),
                                                 _stream(stream)), name + "_p2p")
    return call


sac_apply = _apply("rrl_sac_apply")
qrisk_apply = _apply("rrl_qrisk_apply")
recovery_apply = _apply("rrl_recovery_apply")


def twin_q_forward(cfg, arena, net, n, s, a, q1, q2, stream=None):
    _check(lib().rrl_twin_q_forward(C.byref(cfg), p(arena, "f32"), int(net), C.c_int64(n), p(s, "f32"), p(a, "f32"),
                                    p(q1, "f32"), p(q2, "f32"), _stream(stream)), "rrl_twin_q_forward")


def policy_sample(cfg, arena, net, n, s, eps, action, log_prob=None, mean_action=None, stream=None):
    _check(lib().rrl_policy_sample(C.byref(cfg), p(arena, "f32"), int(net), C.c_int64(n), p(s, "f32"), p(eps, "f32"),
                                   p(action, "f32"), p(log_prob, "f32"), p(mean_action, "f32"), _stream(stream)),
           "rrl_policy_sample")


def hard_update(cfg, arena, dst_net, src_net, stream=None):
    _check(lib().rrl_hard_update(C.byref(cfg), p(arena, "f32"), int(dst_net), int(src_net), _stream(stream)),
           "rrl_hard_update")


def soft_update(cfg, arena, dst_net, src_net, tau, stream=None):
    _check(lib().rrl_soft_update(C.byref(cfg), p(arena, "f32"), int(dst_net), int(src_net), C.c_float(float(tau)),
                                 _stream(stream)), "rrl_soft_update")


# ------------------------------------------------------------------------------------------------
# model-based recovery (PETS / CEM planner, csrc/mpc.cu)
# ------------------------------------------------------------------------------------------------
class MpcConfig(C.Structure):
    _fields_ = [("plan_hor", C.c_int32), ("popsize", C.c_int32), ("num_elites", C.c_int32), ("npart", C.c_int32),
                ("num_nets", C.c_int32), ("max_iters", C.c_int32), ("alpha", C.c_double), ("epsilon", C.c_double),
                ("ac_lb", C.c_float * 2), ("ac_ub", C.c_float * 2), ("seed", C.c_uint64), ("stream_id", C.c_int32),
                ("reserved", C.c_int32)]


def mpc_config(plan_hor, popsize, num_elites, npart=20, num_nets=5, max_iters=5, alpha=0.1, epsilon=0.001,
               ac_lb=(-1.0, -1.0), ac_ub=(1.0, 1.0), seed=0, stream_id=0):
    c = MpcConfig()
    c.plan_hor, c.popsize, c.num_elites, c.npart, c.num_nets, c.max_iters = plan_hor, popsize, num_elites, npart, num_nets, max_iters
    c.alpha, c.epsilon = float(alpha), float(epsilon)
    c.ac_lb[0], c.ac_lb[1] = float(ac_lb[0]), float(ac_lb[1])
    c.ac_ub[0], c.ac_ub[1] = float(ac_ub[0]), float(ac_ub[1])
    c.seed, c.stream_id = seed & 0xFFFFFFFFFFFFFFFF, stream_id
    return c


def dyn_image_floats():
    lib().rrl_dyn_image_floats.restype = C.c_int64
    return int(lib().rrl_dyn_image_floats())


def dyn_pack(tensors, hidden, image, stream=None):
    """tensors: the 12 fp32 CUDA tensors lin0_w, lin0_b, ..., lin3_b, inputs_mu, inputs_sigma, max_logvar, min_logvar."""
    args = [p(t.contiguous(), "f32") for t in tensors]
    _check(lib().rrl_dyn_pack(*args, C.c_int(hidden), p(image, "f32"), _stream(stream)), "rrl_dyn_pack")


def mpc_begin(cfg, n, prev_sol, mean, var, active, stream=None):
    _check(lib().rrl_mpc_begin(C.byref(cfg), C.c_int64(n), p(prev_sol, "f64"), p(mean, "f64"), p(var, "f64"), p(active, "i32"),
                               _stream(stream)), "rrl_mpc_begin")


def mpc_sample(cfg, n, it, mean, var, samples, active, z=None, counters=None, stream=None):
    _check(lib().rrl_mpc_sample(C.byref(cfg), C.c_int64(n), C.c_int(it), p(mean, "f64"), p(var, "f64"), p(z, "f64"),
                                p(counters, "i64"), p(samples, "f32"), p(active, "i32"), _stream(stream)), "rrl_mpc_sample")


def mpc_rollout(cfg, agent_cfg, arena, dyn_image, n, state, samples, row_cost, active=None, eps=None, it=0, counters=None,
                stream=None):
    _check(lib().rrl_mpc_rollout(C.byref(cfg), C.byref(agent_cfg), p(arena, "f32"), p(dyn_image, "f32"), C.c_int64(n),
                                 p(state, "f64"), p(samples, "f32"), p(eps, "f32"), p(active, "i32"), C.c_int(it),
                                 p(counters, "i64"), p(row_cost, "f32"), _stream(stream)), "rrl_mpc_rollout")


def mpc_update(cfg, n, it, samples, row_cost, active, mean, var, stream=None):
    _check(lib().rrl_mpc_update(C.byref(cfg), C.c_int64(n), C.c_int(it), p(samples, "f32"), p(row_cost, "f32"),
                                p(active, "i32"), p(mean, "f64"), p(var, "f64"), _stream(stream)), "rrl_mpc_update")


def mpc_finish(cfg, n, mean, prev_sol, action, mask=None, stream=None):
    _check(lib().rrl_mpc_finish(C.byref(cfg), C.c_int64(n), p(mean, "f64"), p(mask, "u8"), p(prev_sol, "f64"),
                                p(action, "f64"), _stream(stream)), "rrl_mpc_finish")


def dyn_train_floats():
    lib().rrl_dyn_train_floats.restype = C.c_int64
    return int(lib().rrl_dyn_train_floats())


def dyn_train_sync(params, wt, stream=None):
    _check(lib().rrl_dyn_train_sync(p(params, "f32"), p(wt, "f32"), _stream(stream)), "rrl_dyn_train_sync")


def dyn_train_step(params, adam_m, adam_v, wt, partial, mu, sigma, inputs, targets, idx, col0, rows, lr, step, ticket,
                   loss_out=None, stream=None):
    _check(lib().rrl_dyn_train_step(p(params, "f32"), p(adam_m, "f32"), p(adam_v, "f32"), p(wt, "f32"), p(partial, "f32"),
                                    p(mu, "f32"), p(sigma, "f32"), p(inputs, "f32"), p(targets, "f32"), p(idx, "i64"),
                                    C.c_int64(idx.shape[1]), C.c_int64(col0), C.c_int(rows), C.c_float(lr), p(step, "i64"),
                                    p(ticket, "i32"), p(loss_out, "f32"), _stream(stream)), "rrl_dyn_train_step")
