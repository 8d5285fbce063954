#!/usr/bin/env python
"""bench.py -- env-steps/sec of the Recovery RL hot path (rollout + SAC update + Q_risk/recovery update per
vector step) on Maze, N parallel env copies per GPU, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one vector step of the workload named in config.workload: for each of the N env copies one
composite action (policy -> Q_risk threshold -> recovery policy), one env step (500 physics substeps), two
replay pushes; plus one SAC update and one Q_risk + recovery-policy update on batches of 256 sampled with
the CPython-compatible sampler.  `value` is timed with inputs resident in HBM (Philox noise generated on
device, whole step replayed as one CUDA graph); `e2e` goes through the host-buffer face (per-step random
draws uploaded from pinned host memory, per-env results + losses + counters downloaded).
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

METRIC = "env-steps/sec (rollout+SAC+Q_risk update) on Maze"
UNIT = "env-steps/s"
# algorithmic work per unit (DESIGN.md "Measurement")
ACT_FLOP_PER_ENV = 134144 + 267264          # policy + twin Q_risk forward; + 133,120 where recovery triggers
ENV_BYTES_PER_STEP = 120                    # state/action/ep_steps read, state/flags write, two 32 B pushes + flag


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="env copies PER GPU")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--env-name", default="maze")
    ap.add_argument("--pretrain", type=int, default=1000, help="Q_risk pre-training updates (untimed)")
    ap.add_argument("--demos", type=int, default=10000)
    ap.add_argument("--replay", type=int, default=1 << 23,
                    help="ring capacity per GPU (transitions). The reference's 1e6 is sized for ONE env; 65536 env copies "
                         "turn that over in 15 vector steps, and once the agent has become safe the constraint ring holds "
                         "fewer violations than one stratified batch needs (random.sample would raise).")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--tc", type=int, default=int(os.environ.get("RRL_TENSOR_CORES", "1")))
    ap.add_argument("--peer-grads", type=int, default=int(os.environ.get("RRL_PEER_GRADS", "1")),
                    help="N>1: sum the ranks' gradients inside the optimizer-step kernel over NVLink peer memory (0: NCCL)")
    ap.add_argument("--e2e-pipeline", type=int, default=1, help="e2e leg through submit/collect (0: serial step_host)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=400)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
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
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sust=float(p["bf16_tflops_sustained"]),
                    src="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------------------
def cpu_reference_run(args, steps, pretrain=10, demos=2000, threads=None):
    """the CPU restatement of the reference loop (N = 1 env) on the host cores: env-steps/s after warm-up."""
    from oracle import envs as oenvs
    from oracle.loop import OracleExperiment
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    exp = OracleExperiment(args.env_name, seed=args.seed, batch_size=args.batch, gamma_safe=0.5, eps_safe=0.15,
                           pos_fraction=0.3 if args.env_name == "maze" else -1.0)
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
    dt = exp.run_steps(steps)
    return steps / dt, cores, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    t_all = []
    K = max(1, args.steps)
    per_step = max(20, min(args.cpu_steps, 4000 // max(1, K + args.warmup)))
    cores = os.cpu_count() or 1
    from oracle import envs as oenvs
    from oracle.loop import OracleExperiment
    torch.set_num_threads(cores)
    exp = OracleExperiment(args.env_name, seed=args.seed, batch_size=args.batch, gamma_safe=0.5, eps_safe=0.15,
                           pos_fraction=0.3 if args.env_name == "maze" else -1.0)
    np.random.seed(args.seed)
    n_demo = min(2000, args.demos)
    tr = oenvs.maze_offline_data(n_demo, np.random.RandomState(args.seed)) if args.env_name == "maze" else \
        oenvs.nav_offline_data(oenvs.KIND_BY_NAME[args.env_name], n_demo)
    exp.pretrain(tr, 10)
    while len(exp.memory) <= args.batch + 1:
        exp.step()
    for _ in range(args.warmup):
        exp.run_steps(per_step)
    for _ in range(K):
        t_all.append(exp.run_steps(per_step))
    total = sum(t_all)
    value = K * per_step / total
    sample = "%d env-steps per bench step, N=1 env, update every step, %d torch threads" % (per_step, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s Recovery RL MF, CPU restatement of the reference loop (oracle/loop.py), N=1 env, batch %d"
                                   % (args.env_name, args.batch), "env_steps_per_bench_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def build_engine(args, rank, world, pg, host_inputs, dev):
    from recovery_rl.engine import VecEngine
    torch.manual_seed(args.seed)                  # identical xavier init on every rank
    maze = args.env_name == "maze"
    cap = int(getattr(args, "replay", 1000000))
    eng = VecEngine(args.env_name, args.envs, batch_size=args.batch, replay_size=cap, safe_replay_size=cap, gamma_safe=0.5 if maze else 0.8,
                    eps_safe=0.15 if maze else 0.3, pos_fraction=0.3 if maze else -1.0, seed=args.seed, device=dev,
                    rank=rank, world_size=world, process_group=pg, host_inputs=host_inputs, use_tensor_cores=args.tc,
                    peer_grads=bool(args.peer_grads))
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
    eng.reset()
    return eng


def grad_mode(eng, world):
    if world == 1:
        return "none (1 GPU)"
    if eng.peer_arena is not None:
        return "peer-memory sum inside the optimizer-step kernel (NVLink loads, 1 flag barrier per optimizer step)"
    return "nccl, flat grad block, 3 per step" + (" (peer mode unavailable: %s)" % eng.peer_error if eng.peer_error else "")


def timed_steps(eng, K, flush, world, barrier):
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
    eng = build_engine(args, rank, world, pg, False, dev)
    for _ in range(3):
        eng.step()                                 # task ring > batch on every rank before capture
    torch.cuda.synchronize()
    eng.capture()
    flush = torch.zeros(64 * 1024 * 1024, device=dev)          # 256 MiB
    for _ in range(W):
        eng.replay()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    times = timed_steps(eng, K, flush, world, barrier)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
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
    act_tf = ACT_FLOP_PER_ENV * args.envs / (act_ms * 1e-3) / 1e12
    env_gbs = ENV_BYTES_PER_STEP * args.envs / (env_ms * 1e-3) / 1e9
    roofline = {"kernel": "act_kernel" if not args.tc else "act_tc_kernel", "bound": "tensor", "achieved": act_tf,
                "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": act_tf / peaks["tf_burst"],
                "traffic": traffic.get("act_kernel_dram_bytes_per_launch"), "ms_per_launch": act_ms,
                "share_of_step": act_ms / (total_ms / K), "peak_source": peaks["src"] + ", bf16 dense burst",
                "algorithmic_flop_per_launch": ACT_FLOP_PER_ENV * args.envs}
    roofline_env = {"kernel": "env_step_kernel", "bound": "hbm", "achieved": env_gbs, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": env_gbs / peaks["hbm"], "traffic": traffic.get("env_step_kernel_dram_bytes_per_launch"),
                    "ms_per_launch": env_ms, "share_of_step": env_ms / (total_ms / K),
                    "note": "maze: 500 dependent fp64 substeps per env -> latency-bound, not bandwidth-bound (DESIGN.md)"}

    # ---- end to end through the host-buffer face (pinned H2D of the step's random draws, D2H of results) ----
    del flush
    e2e = None
    try:
        eh = build_engine(args, rank, world, pg, True, dev)
        pool = []
        rs = np.random.RandomState(args.seed + 77 + rank)
        n, B = args.envs, args.batch
        for _ in range(4):
            d = dict(reset_draws=rs.rand(2, n), eps_task=rs.randn(n, 2).astype(np.float32),
                     eps_rec=rs.randn(n, 2).astype(np.float32), rand_u=rs.rand(n, 2).astype(np.float32),
                     sac_eps_next=rs.randn(B, 2).astype(np.float32), sac_eps_cur=rs.randn(B, 2).astype(np.float32),
                     qr_eps_next=rs.randn(B, 2).astype(np.float32), qr_eps_rec=rs.randn(B, 2).astype(np.float32))
            if args.env_name != "maze":
                d["env_noise"] = rs.randn(2, n)
            pool.append({k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in d.items()})
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
        e2e_s = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e = {"value": world * args.envs * K / float(e2e_s.item()), "unit": UNIT,
               "h2d_bytes_per_step": eh.h2d_bytes_per_step(), "d2h_bytes_per_step": eh.d2h_bytes_per_step(),
               "ms_per_step": 1e3 * float(e2e_s.item()) / K,
               "api": ("VecEngine.submit/collect: pinned H2D of the step's random draws and D2H of per-env next_state/reward/"
                       "flags/action + losses + counters on copy streams, two staging slots, results collected one step "
                       "later (host sync every step)") if args.e2e_pipeline else
                      ("VecEngine.step_host: pinned H2D of the step's random draws, graph replay, D2H of per-env "
                       "next_state/reward/flags/action + losses + counters, host sync every step")}
        assert eh.read_counters()["error"] == 0
        del eh
    except Exception as ex:  # report, never fake
        e2e = {"value": None, "unit": UNIT, "error": repr(ex)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, dt = cpu_reference_run(args, args.cpu_steps)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d env-steps of the CPU restatement of the reference loop (oracle/loop.py), N=1 env, "
                                  "SAC+Q_risk update every step, batch %d, %.1f s" % (args.cpu_steps, args.batch, dt)}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (agent) / f64 (env)", "data": "synthetic",
                "config": {"workload": "%s Recovery RL MF (SAC + Q_risk + MF recovery), %d env copies per GPU, batch %d, "
                                       "one SAC + one Q_risk/recovery update per vector step" % (args.env_name, args.envs, args.batch),
                           "envs_per_gpu": args.envs, "envs_total": world * args.envs, "batch": args.batch,
                           "replay_capacity": eng.task_cap, "pretrain_updates": args.pretrain,
                           "l2": "flushed between timed steps (256 MiB write, outside the per-step event pairs)",
                           "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks",
                           "rng": "Philox4x32-10 on device (value); host draws uploaded (e2e)",
                           "cuda_graph": eng.graph is not None, "tensor_cores": bool(args.tc),
                           "grad_allreduce": grad_mode(eng, world)},
                "roofline": roofline, "roofline_env": roofline_env, "cpu_baseline": cpu_baseline, "e2e": e2e,
                "gpu_launches": eng.launches_per_step * K, "launches_per_step": eng.launches_per_step, "clocks": clocks,
                "counters": {k: c[k] for k in ("total_numsteps", "episodes", "num_viols", "num_successes", "sac_updates",
                                               "qrisk_updates")}}
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
