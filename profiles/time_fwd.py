"""Warm latency of the grouped forward launch (rrl_twin_q_forward, 2 passes) on the SIMT and the tcgen05 path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "recovery-rl_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from recovery_rl import native
from recovery_rl.arena import AgentArena
from recovery_rl.model import build_reference_modules
torch.manual_seed(0)
ar = AgentArena(torch.device("cuda"), max_batch=256, action_scale=(0.1, 0.1))
ar.load_modules(build_reference_modules())
for n in (128, 256, 1024, 65536):
    s = torch.randn(n, 2, device="cuda") * 0.1; a = torch.randn(n, 2, device="cuda") * 0.1
    q1 = torch.zeros(n, device="cuda"); q2 = torch.zeros(n, device="cuda")
    for tc in (0, 1):
        ar.cfg.use_tensor_cores = tc
        for _ in range(20):
            native.twin_q_forward(ar.cfg, ar.arena, native.NET_QRISK, n, s, a, q1, q2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 200
        e0.record()
        for _ in range(K):
            native.twin_q_forward(ar.cfg, ar.arena, native.NET_QRISK, n, s, a, q1, q2)
        e1.record(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                native.twin_q_forward(ar.cfg, ar.arena, native.NET_QRISK, n, s, a, q1, q2)
        g.replay(); torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(); g.replay(); e3.record(); torch.cuda.synchronize()
        print("n=%6d tc=%d: eager %.1f us/launch, graph %.1f us/launch" % (n, tc, e0.elapsed_time(e1) * 1e3 / K, e2.elapsed_time(e3) * 1e3 / 20))
