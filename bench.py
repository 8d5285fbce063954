#!/usr/bin/env python
"""bench.py -- env-steps/sec of the Recovery RL hot path (rollout + SAC update + Q_risk/recovery update per
vector step), N parallel env copies per GPU, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C3|C4|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one vector step of the workload named in config.workload: for each of the N env copies one
composite action (policy -> Q_risk threshold -> recovery policy), one env step, two replay pushes; plus one SAC
update and one Q_risk + recovery-policy update on sampled batches (CPython-compatible sampler).  `value` is timed
with inputs resident in HBM (Philox noise generated on device, whole step replayed as one CUDA graph); `e2e` goes
through the host-buffer face (per-step random draws uploaded from pinned host memory, per-env results + losses +
counters downloaded).  Default workload = BASELINE config C4 (Maze, 65,536 env copies per GPU, batch 256).
`--config C2|C3|C5` run the other BASELINE configs (Navigation1 4,096 x 256; Navigation2 8,192 x 1,024; Maze
model-based recovery 2,048 envs x 1,000 CEM particles) through the same code.

N > 1 adds, in the same JSON line: `strong` (the north-star form of C4: 65,536 env copies SHARDED over the N
GPUs, with the 1-GPU time measured on rank 0 in the same run), `replicas_identical` (bitwise checksum of the
six networks + Adam moments of every rank after the timed region) and `peer_equals_nccl` (the peer-memory
gradient sum against the NCCL all-reduce path over 8 steps from the same start).

`--impl reference` times the CPU restatement of the reference loop (oracle/loop.py; the reference itself is
Python + MuJoCo and cannot travel to the GPU box) on the box's host cores, N = 1 env as the reference runs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "recovery-rl_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

UNIT = "env-steps/s"
ENV_TITLE = {"maze": "Maze", "navigation1": "Navigation1", "navigation2": "Navigation2"}
# algorithmic work per unit (DESIGN.md "Measurement")
ACT_FLOP_PER_ENV = 134144 + 267264          # policy + twin Q_risk forward; + 133,120 where recovery triggers
ENV_BYTES_PER_STEP = 120                    # state/action/ep_steps read, state/flags write, two 32 B pushes + flag
# safety-critic hyper-parameters of the shipped scripts (scripts/{navigation1,navigation2,maze}.sh)
HYPER = {"navigation1": dict(gamma_safe=0.8, eps_safe=0.3, pos_fraction=-1.0),
         "navigation2": dict(gamma_safe=0.65, eps_safe=0.2, pos_fraction=-1.0),
         "maze": dict(gamma_safe=0.5, eps_safe=0.15, pos_fraction=0.3)}
CONFIGS = {"C2": dict(env_name="navigation1", envs=4096, batch=256),
           "C3": dict(env_name="navigation2", envs=8192, batch=1024),
           "C4": dict(env_name="maze", envs=65536, batch=256),
           "C5": dict(env_name="maze", envs=2048, batch=256, mb=1)}
STRONG_ENVS_TOTAL = 65536                   # BASELINE.json configs[3]: 65,536 env copies sharded over the GPUs


def metric_name(env_name):
    return "env-steps/sec (rollout+SAC+Q_risk update) on %s" % ENV_TITLE.get(env_name, env_name)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="a BASELINE.json config (default C4)")
    ap.add_argument("--envs", type=int, default=None, help="env copies PER GPU (default: the config's)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--env-name", default=None)
    ap.add_argument("--mb", type=int, default=None, help="model-based recovery (PETS/CEM planner) instead of the MF recovery policy")
    ap.add_argument("--popsize", type=int, default=50, help="--mb: CEM candidates per env (x 20 particles)")
    ap.add_argument("--pretrain", type=int, default=1000, help="Q_risk pre-training updates (untimed)")
    ap.add_argument("--demos", type=int, default=10000)
    ap.add_argument("--replay", type=int, default=1 << 23,
                    help="ring capacity per GPU (transitions). The reference's 1e6 is sized for ONE env; 65536 env copies "
                         "turn that over in 15 vector steps, and once the agent has become safe the constraint ring holds "
                         "fewer violations than one stratified batch needs (random.sample would raise).")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--tc", type=int, default=int(os.environ.get("RRL_TENSOR_CORES", "2")),
                    help="0: fp32 SIMT kernels, 1: tcgen05 kernels, 2: tcgen05 + update stages fused into the producing kernels")
    ap.add_argument("--peer-grads", type=int, default=int(os.environ.get("RRL_PEER_GRADS", "1")),
                    help="N>1: sum the ranks' gradients inside the optimizer-step kernel over NVLink peer memory (0: NCCL)")
    ap.add_argument("--e2e-pipeline", type=int, default=1, help="e2e leg through submit/collect (0: serial step_host)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-checks", action="store_true", help="N>1: skip the strong-scaling / replica / peer==NCCL blocks")
    ap.add_argument("--cpu-steps", type=int, default=300)
    args = ap.parse_args()
    base = dict(CONFIGS[args.config or "C4"])
    for k in ("env_name", "envs", "batch", "mb"):
        if getattr(args, k) is None:
            setattr(args, k, base.get(k, 0))
    args.config = args.config or ("C4" if (args.env_name, args.envs, args.batch, args.mb) == ("maze", 65536, 256, 0) else "custom")
    return args


# ----------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled from a thread
    every ~2 ms (a 20-step timed region is only ~10 ms long); nvidia-smi -lms as the fallback."""
    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.lines = [], set(), []
        self.mx = None
        self.proc = self.thread = self.h = self.nv = None
        self.stop_flag = False
        self.how = None

    def _nvml_handle(self):
        import pynvml as nv
        nv.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        return nv, h

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                for name, const in self.REASONS:
                    if r & getattr(nv, const):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            self.nv, self.h = self._nvml_handle()
            self.how = "nvml, 2 ms poll"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.how = "nvidia-smi -lms 10"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "how": self.how}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sust=float(p["bf16_tflops_sustained"]),
                    src="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def load_traffic():
    """dram bytes per launch of the dominant kernels from the committed `ncu --set full` captures (profiles/): ncu
    cannot run inside a timed bench, so this is the one roofline field that is NOT measured live."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------------------
def _oracle_experiment(args, demos, pretrain):
    from oracle import envs as oenvs
    from oracle.loop import OracleExperiment
    h = HYPER[args.env_name]
    exp = OracleExperiment(args.env_name, seed=args.seed, batch_size=args.batch, gamma_safe=h["gamma_safe"],
                           eps_safe=h["eps_safe"], pos_fraction=h["pos_fraction"])
    np.random.seed(args.seed)
    if args.env_name == "maze":
        tr = oenvs.maze_offline_data(demos, np.random.RandomState(args.seed))
    else:
        tr = oenvs.nav_offline_data(oenvs.KIND_BY_NAME[args.env_name], demos)
    exp.pretrain(tr, pretrain)
    while len(exp.memory) <= args.batch + 1:      # untimed: fill the task buffer until updates run every step
        exp.step()
    for _ in range(5):
        exp.step()
    return exp


def thread_candidates():
    n = os.cpu_count() or 1
    return sorted(set(t for t in (1, 4, 8, n) if t <= n))


def cpu_reference_sweep(args, steps):
    """the CPU restatement of the reference loop (N = 1 env) on the host cores, at 1 / 4 / 8 / all torch threads:
    (best env-steps/s, its thread count, {threads: env-steps/s}, seconds spent).  256-wide MLPs on batch 256 do not
    scale past a few threads, so "all threads" alone would handicap the baseline."""
    exp = _oracle_experiment(args, 2000, 10)
    cand = thread_candidates()
    per = max(20, steps // len(cand))
    by, spent = {}, 0.0
    for t in cand:
        torch.set_num_threads(t)
        exp.run_steps(5)
        dt = exp.run_steps(per)
        by[t] = per / dt
        spent += dt
    best = max(by, key=by.get)
    return by[best], best, by, per, spent


def run_reference(args, rank, world):
    if rank != 0:
        return
    K = max(1, args.steps)
    per_step = max(20, min(args.cpu_steps, 4000 // max(1, K + args.warmup)))
    exp = _oracle_experiment(args, min(2000, args.demos), 10)
    # thread count: the best of 1 / 4 / 8 / all on a short probe (stated in the line)
    probe = {}
    for t in thread_candidates():
        torch.set_num_threads(t)
        exp.run_steps(3)
        probe[t] = 15 / exp.run_steps(15)
    cores = max(probe, key=probe.get)
    torch.set_num_threads(cores)
    for _ in range(args.warmup):
        exp.run_steps(per_step)
    t_all = [exp.run_steps(per_step) for _ in range(K)]
    total = sum(t_all)
    value = K * per_step / total
    sample = "%d env-steps per bench step, N=1 env, update every step, %d torch threads (best of %s on a 15-step probe)" % (
        per_step, cores, {k: round(v, 1) for k, v in probe.items()})
    line = {"impl": "reference", "metric": metric_name(args.env_name), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s Recovery RL MF, CPU restatement of the reference loop (oracle/loop.py), N=1 env, batch %d"
                                   % (args.env_name, args.batch), "env_steps_per_bench_step": per_step,
                       "host_cores": os.cpu_count(), "torch_threads": cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def build_engine(args, rank, world, pg, host_inputs, dev, envs=None, peer_grads=None):
    from recovery_rl.engine import VecEngine
    torch.manual_seed(args.seed)                  # identical xavier init on every rank
    maze = args.env_name == "maze"
    cap = int(getattr(args, "replay", 1000000))
    h = HYPER[args.env_name]
    mb = bool(getattr(args, "mb", 0))
    peer = bool(args.peer_grads if peer_grads is None else peer_grads)
    eng = VecEngine(args.env_name, int(envs or args.envs), batch_size=args.batch, replay_size=cap, safe_replay_size=cap,
                    gamma_safe=h["gamma_safe"], eps_safe=h["eps_safe"], pos_fraction=h["pos_fraction"], seed=args.seed, device=dev,
                    rank=rank, world_size=world, process_group=pg, host_inputs=host_inputs, use_tensor_cores=args.tc,
                    peer_grads=peer, mf_recovery=not mb, mb_recovery=mb, mpc_popsize=getattr(args, "popsize", 50) if mb else None)
    eng.init_agent()
    rng = np.random.RandomState(args.seed + 1000 * rank)
    if maze:
        from env.maze import get_offline_data
        demos = get_offline_data(args.demos, rng=rng)
    else:
        import importlib
        np.random.seed(args.seed + 1000 * rank)
        demos = importlib.import_module("env." + args.env_name).get_offline_data(min(args.demos, 4000))
    eng.push_offline(demos)
    eng.pretrain_qrisk(args.pretrain, n_demos=len(demos))
    if mb:
        eng.train_mb(demos[:4000], epochs=5)      # experiment.py:298-305 (ensemble fit on the demos, untimed)
    eng.reset()
    return eng


def grad_mode(eng, world):
    if world == 1:
        return "none (1 GPU)"
    if eng.peer_arena is not None:
        if getattr(eng.peer_arena, "multicast", False) and eng.fused_barrier:
            return "in-switch sum inside the optimizer-step kernel (multimem.ld_reduce on the symmetric arena, flag barriers)"
        return "peer-memory sum inside the optimizer-step kernel (NVLink loads, flag barriers)"
    return "nccl, flat grad block" + (" (peer mode unavailable: %s)" % eng.peer_error if eng.peer_error else "")


def timed_steps(eng, K, flush, barrier):
    """K graph replays, each bracketed by CUDA events on the launching stream; L2 flushed between steps."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    torch.cuda.synchronize()
    for i in range(K):
        flush.add_(1.0)                            # > L2 (126 MB): evicts the previous step's lines
        ev[i][0].record()
        eng.replay()
        ev[i][1].record()
    torch.cuda.synchronize()
    barrier()
    return [a.elapsed_time(b) for a, b in ev]


def time_kernel(fn, K, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize()
    for i in range(K):
        flush.add_(1.0)
        ev[i][0].record()
        fn()
        ev[i][1].record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def ready_engine(args, rank, world, pg, dev, W, envs=None, peer_grads=None):
    """engine with the task ring past the batch size on every rank, step captured as a CUDA graph, W warm-up replays."""
    eng = build_engine(args, rank, world, pg, False, dev, envs=envs, peer_grads=peer_grads)
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    eng.capture()
    for _ in range(W):
        eng.replay()
    torch.cuda.synchronize()
    return eng


def max_over_ranks(x, dev, world):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def state_checksum(eng):
    """two 64-bit checksums over the BITS of the six networks (+ images) and the Adam moments of this rank."""
    a = eng.agent
    m_off, cnt = native_scratch(a, "adam_m")
    v_off, _ = native_scratch(a, "adam_v")
    parts = [eng.arena[:a.grad_off], eng.arena[m_off:m_off + cnt], eng.arena[v_off:v_off + cnt]]
    out = []
    for p in parts:
        bits = p.contiguous().view(torch.int32).to(torch.int64)
        w = (torch.arange(bits.numel(), device=bits.device, dtype=torch.int64) % 65521) + 1
        out += [bits.sum(), (bits * w).sum()]
    return torch.stack(out)


def native_scratch(agent, name):
    from recovery_rl import native
    return native.agent_scratch_info(agent.cfg, name)


def multi_gpu_checks(args, rank, world, pg, dev, eng, K, W, flush, barrier):
    """N > 1: replica identity, the peer-memory sum against NCCL, and C4 as BASELINE states it (65,536 envs sharded)."""
    import torch.distributed as dist
    out = {}
    # (1) every rank holds bit-identical networks + Adam state after the timed region
    cs = state_checksum(eng)
    allcs = [torch.empty_like(cs) for _ in range(world)]
    dist.all_gather(allcs, cs)
    out["replicas_identical"] = bool(all(torch.equal(c, allcs[0]) for c in allcs))
    # (2) peer-memory gradient sum vs NCCL all-reduce: 8 steps from the same start (seeded build), small shard
    try:
        res = []
        for peer in (1, 0):
            small = argparse.Namespace(**vars(args))
            small.envs, small.replay, small.pretrain, small.demos = 4096, 1 << 18, 50, 4000
            e = build_engine(small, rank, world, pg, False, dev, peer_grads=peer)
            active = e.peer_arena is not None
            for _ in range(8):
                e.step()
            torch.cuda.synchronize()
            assert e.read_counters()["error"] == 0
            res.append((e.arena[:e.agent.grad_off].clone(), active, e.read_counters()))
            del e
        (pa, p_active, pc), (na, _, nc) = res
        diff = float((pa - na).abs().max().item())
        scale = float(na.abs().max().item())
        info = {"peer_path_active": bool(p_active), "steps": 8, "envs_per_gpu": 4096, "bit_equal": bool(torch.equal(pa, na)),
                "max_abs_diff": diff, "max_abs_param": scale, "updates": [pc["sac_updates"], pc["qrisk_updates"]],
                "same_counters": pc["sac_updates"] == nc["sac_updates"] and pc["num_viols"] == nc["num_viols"],
                "note": "peer path sums in rank order; NCCL's reduction order is its own: bit-equal is guaranteed at 2 ranks only"}
        flag = torch.tensor([1 if (info["bit_equal"] or diff <= 1e-5 * max(scale, 1.0)) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        info["equal_within_1e-5_all_ranks"] = bool(flag.item())
        out["peer_equals_nccl"] = info
    except Exception as ex:
        out["peer_equals_nccl"] = {"error": repr(ex)}
    # (3) strong scaling: 65,536 env copies in total, sharded; the 1-GPU time of the same total measured on rank 0
    try:
        per = STRONG_ENVS_TOTAL // world
        es = ready_engine(args, rank, world, pg, dev, W, envs=per)
        t = timed_steps(es, K, flush, barrier)
        ms = max_over_ranks(sum(t), dev, world) / K
        assert es.read_counters()["error"] == 0
        es.graph = None
        del es
        n1_ms = None
        if rank == 0:             # the other ranks wait at the barrier below; their GPUs idle
            e1 = ready_engine(args, 0, 1, None, dev, W, envs=STRONG_ENVS_TOTAL)
            t1 = timed_steps(e1, K, flush, lambda: None)
            n1_ms = sum(t1) / K
            e1.graph = None
            del e1
        barrier()
        strong = {"envs_total": STRONG_ENVS_TOTAL, "envs_per_gpu": per, "ms_per_step": ms,
                  "value": STRONG_ENVS_TOTAL / (ms * 1e-3), "unit": UNIT}
        if n1_ms is not None:
            strong["n1_ms_per_step"] = n1_ms
            strong["n1_value"] = STRONG_ENVS_TOTAL / (n1_ms * 1e-3)
            strong["speedup_vs_n1"] = n1_ms / ms
            strong["efficiency_vs_n1"] = n1_ms / ms / world
        out["strong"] = strong
    except Exception as ex:
        out["strong"] = {"error": repr(ex)}
    return out


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from recovery_rl import native
    native.require_cuda()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD

    def barrier():
        if world > 1:
            dist.barrier()

    K, W = args.steps, max(3, args.warmup)
    eng = ready_engine(args, rank, world, pg, dev, W)
    flush = torch.zeros(64 * 1024 * 1024, device=dev)          # 256 MiB
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    native.debug_opt_times()                    # reset the optimizer-step diagnostics (include/rrl.h, rrl_debug_opt_times)
    times = timed_steps(eng, K, flush, barrier)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(sum(times), dev, world)
    od = native.debug_opt_times()               # device globaltimer sums over the timed region (tcgen05 path; zeros otherwise)
    opt_step_us = od["kernels_us"] / K          # inside the optimizer-step kernels, per step
    opt_dbg = None
    if od["launches"]:
        n = od["launches"]
        opt_dbg = {"launches": n, "barrier_us_per_launch": od["cta0_barrier_us"] / n,
                   "grad_loads_us_per_launch": od["cta0_grad_loads_us"] / n, "rest_us_per_launch": od["cta0_rest_us"] / n}
    barrier_wait = None
    if world > 1:
        # how long each rank's optimizer-step kernels waited for the slowest peer's gradient flag inside the timed region: rank
        # skew + signal latency; the rest of the multi-GPU overhead is the peer loads (optimizer_step_cta0)
        bw = torch.tensor([od["flag_wait_us"], float(od["barriers"])], dtype=torch.float64, device=dev)
        allbw = [torch.empty_like(bw) for _ in range(world)]
        dist.all_gather(allbw, bw)
        per_rank = [float(b[0]) / K for b in allbw]
        barrier_wait = {"us_per_step_by_rank": [round(x, 2) for x in per_rank], "us_per_step_mean": sum(per_rank) / world,
                        "barriers_per_step": float(allbw[0][1]) / K,
                        "note": "time CTA 0 of the optimizer-step kernels spent waiting for the peers' gradient flags (device globaltimer)"}
    c = eng.read_counters()
    assert c["error"] == 0, "device-side error %r" % (c,)
    value = world * args.envs * K / (total_ms * 1e-3)

    # ---- dominant kernels, timed alone with CUDA events on the launching stream (live) ----
    peaks = load_peaks()
    traffic = load_traffic()
    act_ms = time_kernel(lambda: native.agent_act(eng.cfg, eng.arena, eng.n, eng.state, eng.counters, eng.action_task,
                                                  eng.action_real, eng.recovery, eng.qrisk, use_recovery=True,
                                                  start_steps=eng.start_steps, seed=eng.seed, stream_id=rank), 20, flush)
    snap = eng.snapshot()
    env_ms = time_kernel(lambda: native.env_step(eng.env_cfg, eng.action_task, eng.action_real, eng.state, eng.ep_steps,
                                                 eng.ep_return, eng.counters, recovery=eng.recovery,
                                                 task_ring=eng.task_ring, task_capacity=eng.task_cap,
                                                 cons_ring=eng.cons_ring, cons_flags=eng.cons_flags,
                                                 cons_capacity=eng.cons_cap), 20, flush)
    eng.restore(snap)
    # the update chain (sample + SAC + Q_risk + recovery updates), eager launches on the stream, same L2-flushed timing
    snap = eng.snapshot()

    def updates_only():
        eng._sac_sample()
        if eng.online_qrisk:
            eng._qr_sample()
        eng._sac_compute()
        if eng.online_qrisk:
            eng._qr_compute()
    upd_ms = None
    if world == 1:
        upd_ms = time_kernel(updates_only, 20, flush)
    eng.restore(snap)
    act_tf = ACT_FLOP_PER_ENV * args.envs / (act_ms * 1e-3) / 1e12
    env_gbs = ENV_BYTES_PER_STEP * args.envs / (env_ms * 1e-3) / 1e9
    step_ms = total_ms / K
    roofline = {"kernel": "act_kernel" if not args.tc else "act_tc_kernel", "bound": "tensor", "achieved": act_tf,
                "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": act_tf / peaks["tf_burst"],
                "traffic": traffic.get("act_kernel_dram_bytes_per_launch"),
                "traffic_source": traffic.get("source", "profiles/roofline_traffic.json (ncu --set full capture; not measured in this run)"),
                "ms_per_launch": act_ms,
                "share_of_step": act_ms / step_ms, "peak_source": peaks["src"] + ", bf16 dense burst",
                "algorithmic_flop_per_launch": ACT_FLOP_PER_ENV * args.envs}
    roofline_env = {"kernel": "env_step_kernel", "bound": "hbm", "achieved": env_gbs, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": env_gbs / peaks["hbm"], "traffic": traffic.get("env_step_kernel_dram_bytes_per_launch"),
                    "ms_per_launch": env_ms, "share_of_step": env_ms / step_ms,
                    "note": ("maze: 500 dependent fp64 substeps per env -> fp64-pipe / latency bound, not bandwidth-bound (DESIGN.md)"
                             if args.env_name == "maze" else "navigation: one fp64 step per env; 120 B/env-step algorithmic")}
    breakdown = {"step_ms": step_ms, "act_ms": act_ms, "env_ms": env_ms, "updates_eager_ms": upd_ms,
                 "optimizer_step_kernels_us_per_step": opt_step_us, "optimizer_step_cta0": opt_dbg,
                 "note": "kernels timed alone, L2 flushed before each; the update chain as eager launches (the step replays them in a graph)"}

    multi = {}
    if world > 1 and not args.no_checks:
        multi = multi_gpu_checks(args, rank, world, pg, dev, eng, K, W, flush, barrier)
    if barrier_wait is not None:
        multi["barrier_wait"] = barrier_wait

    # ---- end to end through the host-buffer face (pinned H2D of the step's random draws, D2H of results) ----
    del flush
    e2e = None
    try:
        eh = build_engine(args, rank, world, pg, True, dev)
        pool = []
        rs = np.random.RandomState(args.seed + 77 + rank)
        n, B = args.envs, args.batch
        for _ in range(4):      # four pre-drawn input sets, each ONE pinned flat buffer (VecEngine.new_host_inputs)
            d = dict(reset_draws=rs.rand(2, n), eps_task=rs.randn(n, 2).astype(np.float32),
                     eps_rec=rs.randn(n, 2).astype(np.float32), rand_u=rs.rand(n, 2).astype(np.float32),
                     sac_eps_next=rs.randn(B, 2).astype(np.float32), sac_eps_cur=rs.randn(B, 2).astype(np.float32),
                     qr_eps_next=rs.randn(B, 2).astype(np.float32), qr_eps_rec=rs.randn(B, 2).astype(np.float32))
            if args.env_name != "maze":
                d["env_noise"] = rs.randn(2, n)
            hs = eh.new_host_inputs()
            for k, v in d.items():
                hs[k].copy_(torch.from_numpy(np.ascontiguousarray(v)))
            pool.append(hs)
        for i in range(3):
            eh.step_host(pool[i % 4])
        eh.capture()
        for i in range(W):
            eh.step_host(pool[i % 4])
        base_ticket = 0
        if args.e2e_pipeline:
            eh.enable_pipeline()
            for i in range(4):                      # warm the two slots
                tk = eh.submit(pool[i % 4])
                if i >= 1:
                    eh.collect(tk - 1)
            eh.collect(3)
            base_ticket = 4
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.e2e_pipeline:
            # every step still uploads its own inputs and downloads its own results inside the timed region; the
            # copies run on their own streams next to the neighbouring steps' compute (results arrive one step later)
            for i in range(K):
                tk = eh.submit(pool[i % 4])
                if i >= 1:
                    eh.collect(tk - 1)
            eh.collect(K - 1 + base_ticket)
        else:
            for i in range(K):
                eh.step_host(pool[i % 4])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e_s = max_over_ranks(t1 - t0, dev, world)
        e2e = {"value": world * args.envs * K / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": eh.h2d_bytes_per_step(), "d2h_bytes_per_step": eh.d2h_bytes_per_step(),
               "ms_per_step": 1e3 * e2e_s / K,
               "api": ("VecEngine.submit/collect: pinned H2D of the step's random draws (one flat buffer) and D2H of per-env "
                       "next_state/reward/flags/action (one flat buffer) + losses + counters on copy streams, two staging "
                       "slots, results collected one step later (host sync every step)") if args.e2e_pipeline else
                      ("VecEngine.step_host: pinned H2D of the step's random draws, graph replay, D2H of per-env "
                       "next_state/reward/flags/action + losses + counters, host sync every step")}
        assert eh.read_counters()["error"] == 0
        eh.graph = None
        del eh
    except Exception as ex:  # report, never fake
        e2e = {"value": None, "unit": UNIT, "error": repr(ex)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.mb:
        v, cores, by, per, dt = cpu_reference_sweep(args, args.cpu_steps)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "host_cores": os.cpu_count(),
                        "by_threads": {str(k): x for k, x in by.items()},
                        "sample": "%d env-steps per thread setting of the CPU restatement of the reference loop (oracle/loop.py), "
                                  "N=1 env, SAC+Q_risk update every step, batch %d, %.1f s in all; value = the best setting"
                                  % (per, args.batch, dt)}
    if rank == 0:
        recov = "model-based recovery (PETS/CEM, %d particles per env)" % (args.popsize * 20) if args.mb else "MF recovery"
        line = {"metric": metric_name(args.env_name), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (agent) / f64 (env)", "data": "synthetic",
                "config": {"workload": "%s Recovery RL (SAC + Q_risk + %s), %d env copies per GPU, batch %d, "
                                       "one SAC + one Q_risk/recovery update per vector step" % (args.env_name, recov, args.envs, args.batch),
                           "baseline_config": args.config,
                           "envs_per_gpu": args.envs, "envs_total": world * args.envs, "batch": args.batch,
                           "replay_capacity": eng.task_cap, "pretrain_updates": args.pretrain,
                           "l2": "flushed between timed steps (256 MiB write, outside the per-step event pairs)",
                           "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks",
                           "rng": "Philox4x32-10 on device (value); host draws uploaded (e2e)",
                           "cuda_graph": eng.graph is not None, "tensor_cores": bool(args.tc),
                           "grad_allreduce": grad_mode(eng, world),
                           "programmatic_dependent_launch": __import__("recovery_rl.native", fromlist=["x"]).pdl_enabled(),
                           "stage_ctas": eng.stage_ctas},
                "roofline": roofline, "roofline_env": roofline_env, "breakdown": breakdown, "cpu_baseline": cpu_baseline, "e2e": e2e,
                "gpu_launches": eng.launches_per_step * K, "launches_per_step": eng.launches_per_step, "clocks": clocks,
                "counters": {k: c[k] for k in ("total_numsteps", "episodes", "num_viols", "num_successes", "sac_updates",
                                               "qrisk_updates")}}
        line.update(multi)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # the captured graphs hold NCCL kernels: drop them before the communicator, then leave without the
        # (sometimes blocking) communicator teardown
        eng.graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517")] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
