"""GPU: checkpoints (recovery_rl/checkpoint.py) -- a resumed engine continues bit-identically, and the saved
networks load into the reference-shaped torch modules by their state_dict names."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(seed=9, **kw):
    from recovery_rl.engine import VecEngine
    from env.maze import get_offline_data
    torch.manual_seed(1)
    e = VecEngine("maze", 1024, batch_size=64, replay_size=4096, safe_replay_size=4096, gamma_safe=0.5, eps_safe=0.15,
                  pos_fraction=0.3, seed=seed, **kw)
    e.init_agent()
    e.push_offline(get_offline_data(1500, rng=np.random.RandomState(4)))
    e.pretrain_qrisk(5)
    e.reset()
    return e


def test_engine_resume_is_bit_identical(native, cuda, tmp_path):
    a = _engine()
    for _ in range(7):                      # rings (4096 slots) wrap: 7 * 1024 pushes
        a.step()
    path = a.save(os.path.join(str(tmp_path), "ck.pt"))
    for _ in range(5):
        a.step()
    torch.manual_seed(77)                   # a different init: everything must come from the checkpoint
    b = _engine(seed=9)
    b.load(path)
    b.capture()                             # resumed runs replay the CUDA graph
    for _ in range(5):
        b.replay()
    torch.cuda.synchronize()
    ca, cb = a.counters.clone(), b.counters.clone()
    ca[native.C_RETURN_SUM_BITS] = cb[native.C_RETURN_SUM_BITS] = 0
    assert torch.equal(ca, cb), (ca.tolist(), cb.tolist())
    assert torch.equal(a.state, b.state) and torch.equal(a.mt_state, b.mt_state)
    assert torch.equal(a.arena[:a.agent.grad_off], b.arena[:b.agent.grad_off])
    assert torch.equal(a.task_ring, b.task_ring) and torch.equal(a.cons_flags, b.cons_flags)
    assert a.read_counters()["sac_updates"] == 11


def test_checkpoint_loads_into_reference_shaped_modules(native, cuda, tmp_path):
    from recovery_rl import checkpoint
    from recovery_rl.model import build_reference_modules
    e = _engine(rcpo=True, lambda_rcpo=50.0)
    for _ in range(3):
        e.step()
    st = checkpoint.engine_state(e)
    mods = build_reference_modules(hidden=256)
    for net, mod in mods.items():
        missing, unexpected = mod.load_state_dict(st["nets"][net], strict=True)
        assert not missing and not unexpected
        for i, p in enumerate(mod.parameters()):
            assert torch.equal(p.detach(), e.agent.tensor(net, i).cpu())
    assert st["optim"]["critic"]["step"] == 2 and st["optim"]["qrisk"]["step"] >= 5
    # agent-only checkpoint through the drop-in handle; multipliers round-trip
    from recovery_rl.arena import AgentArena
    ar = AgentArena(cuda, max_batch=64, gamma_safe=0.5, eps_safe=0.15, action_scale=(0.1, 0.1), rcpo=True, lambda_rcpo=50.0)
    ar.cfg.action_scale[0] = e.cfg.action_scale[0]
    checkpoint.load_agent_state(ar, st, load_counters=False)
    assert torch.equal(ar.arena[:ar.grad_off], e.arena[:e.agent.grad_off])
    assert torch.equal(ar.scratch("scalars"), e.agent.scratch("scalars"))
    assert int(ar.counters[native.C_ADAM_T0]) == 2
