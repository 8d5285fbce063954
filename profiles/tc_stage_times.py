"""Per-CTA stage times of the tcgen05 update kernels (fwd_tc_kernel / bwd_tc_kernel) inside one eager vector step.

Builds a profiling variant of the library (-DRRL_TC_TIMING: globaltimer stamps written by thread 0 of every CTA) into
profiles/_timing/librrl_timing.so, loads it INSTEAD of librrl.so (this script only), runs the bench workload for a few
steps and prints, per launch: kind, grid, and for every CTA the microseconds spent in
    setup (params -> smem, barriers, TMEM alloc) | produce (layer 1 / dh2 rebuild + fp16 split, MMAs overlapped) |
    MMA drain (accumulator ready) | epilogue | TMEM release | tail (loss stage, last CTA only)
  python profiles/tc_stage_times.py [--tc 2] > gpurun_out/tc_stage_times.txt
Numbers from this build are for attribution only (the stamps cost a few hundred ns per CTA); never bench with it.
"""
import argparse
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "recovery-rl_b200")
for p in (ROOT, PKG):
    sys.path.insert(0, p)


def build_timing_lib():
    out_dir = os.path.join(ROOT, "profiles", "_timing")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "librrl_timing.so")
    csrc = os.path.join(PKG, "csrc")
    srcs = [("api.cu", []), ("env.cu", ["-fmad=false"]), ("replay.cu", ["-fmad=false"]), ("agent.cu", []),
            ("agent_tc.cu", ["-DRRL_TC_TIMING"]), ("select.cu", []), ("mpc.cu", []), ("mpc_tc.cu", [])]
    objs, procs = [], []
    for src, extra in srcs:
        obj = os.path.join(out_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
               "-I", os.path.join(ROOT, "include")] + extra + ["-c", os.path.join(csrc, src), "-o", obj]
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        assert p.wait() == 0
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tc", type=int, default=2)
    ap.add_argument("--envs", type=int, default=65536)
    args = ap.parse_args()
    lib_path = build_timing_lib()
    from recovery_rl import native
    native.LIB_PATH = lib_path
    import numpy as np
    import torch
    from recovery_rl.engine import VecEngine
    from env.maze import get_offline_data
    torch.manual_seed(1)
    eng = VecEngine("maze", args.envs, batch_size=256, gamma_safe=0.5, eps_safe=0.15, pos_fraction=0.3, seed=1,
                    use_tensor_cores=args.tc)
    eng.init_agent()
    eng.push_offline(get_offline_data(10000, rng=np.random.RandomState(1)))
    eng.pretrain_qrisk(20)
    eng.reset()
    for _ in range(6):
        eng.step()
    torch.cuda.synchronize()
    lib = native.lib()
    log = (ctypes.c_uint64 * (256 * 16 * 8))()
    kind = (ctypes.c_int32 * (256 * 4))()
    n0 = lib.rrl_debug_tc_times(log, kind)
    eng.step()
    torch.cuda.synchronize()
    n1 = lib.rrl_debug_tc_times(log, kind)
    assert n1 > n0 >= 0, (n0, n1)
    L = np.frombuffer(log, dtype=np.uint64).reshape(256, 16, 8).astype(np.int64)
    K = np.frombuffer(kind, dtype=np.int32).reshape(256, 4)
    names = {1: "fwd_tc", 2: "bwd_tc"}
    t_first = None
    print("launch  kernel  grid   cta :   setup  produce  mma_drain  epilogue  release  tail |  total   (us; start offset from the step's first launch)")
    for l in range(n0, n1):
        k = K[l & 255]
        ncta = int(k[1] * k[2])
        for c in range(min(ncta, 16)):
            s = L[l & 255, c]
            if t_first is None:
                t_first = s[0]
            st = [(s[i + 1] - s[i]) * 1e-3 for i in range(5)]
            tail = (s[6] - s[5]) * 1e-3 if k[0] == 1 else 0.0
            end = s[6] if k[0] == 1 else s[5]
            print("%5d  %7s  %dx%d  %3d : %7.2f  %7.2f  %9.2f  %8.2f  %7.2f  %5.2f | %6.2f   @%8.2f" % (
                l, names.get(int(k[0]), "?"), k[1], k[2], c, st[0], st[1], st[2], st[3], st[4], tail, (end - s[0]) * 1e-3,
                (s[0] - t_first) * 1e-3))
    print(eng.read_counters())


if __name__ == "__main__":
    main()
