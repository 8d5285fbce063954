"""PointEnvBatch -- n copies of a point environment stepped by the CUDA kernel (csrc/env.cu) with
gym-style explicit resets (no auto-reset), numpy in / numpy out.  Host mirror used by the drop-in
Navigation1 / Navigation2 / MazeNavigation classes and by the offline-data generators; the training
engine (recovery_rl/engine.py) talks to the kernels directly and never leaves the device."""
import numpy as np
import torch

from recovery_rl import native


class PointEnvBatch(object):
    def __init__(self, env_name, n=1, device=None, horizon=100, seed=0):
        native.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.kind = native.ENV_KIND[env_name]
        self.n = n
        self.cfg = native.env_config(self.kind, n, horizon=horizon, seed=seed, auto_reset=False)
        dev = self.device
        self.state = torch.zeros(2, n, dtype=torch.float64, device=dev)
        self.ep_steps = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ep_return = torch.zeros(n, dtype=torch.float64, device=dev)
        self.counters = torch.zeros(native.NUM_COUNTERS, dtype=torch.int64, device=dev)
        self.o_next = torch.zeros(2, n, dtype=torch.float64, device=dev)
        self.o_reward = torch.zeros(n, dtype=torch.float64, device=dev)
        self.o_flags = torch.zeros(3, n, dtype=torch.uint8, device=dev)
        self.noise = torch.zeros(2, n, dtype=torch.float64, device=dev)
        self.action = torch.zeros(n, 2, dtype=torch.float32, device=dev)
        self.action64 = torch.zeros(n, 2, dtype=torch.float64, device=dev)

    def set_state(self, state, ep_steps=None):
        self.state.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(state, np.float64).reshape(self.n, 2).T)))
        if ep_steps is None:
            self.ep_steps.zero_()
        else:
            self.ep_steps.copy_(torch.from_numpy(np.asarray(ep_steps, np.int32).reshape(self.n)))

    def get_state(self):
        return self.state.cpu().numpy().T.copy()

    def reset_from_draws(self, draws):
        """draws [n,2]: N(0,1) for navigation (navigation1.py:92), U[0,1) for maze mode 'h' (maze.py:195-197)."""
        d = torch.from_numpy(np.ascontiguousarray(np.asarray(draws, np.float64).reshape(self.n, 2).T)).to(self.device)
        native.env_reset(self.cfg, self.state, self.ep_steps, self.ep_return, self.counters, draws=d)
        return self.get_state()

    def step(self, action, noise=None):
        """action [n,2] (fp32 as the policy outputs it; fp64 actions drive the dynamics in fp64); noise [n,2] N(0,1) for navigation.
        Returns next_state, reward, done (incl. the env's own horizon rule for maze), constraint, success."""
        action = np.asarray(action)
        a = np.ascontiguousarray(action.astype(np.float32).reshape(self.n, 2))
        self.action.copy_(torch.from_numpy(a))
        a64 = None
        if action.dtype == np.float64:          # keep fp64 actions exact (np.clip preserves the dtype)
            self.action64.copy_(torch.from_numpy(np.ascontiguousarray(action.reshape(self.n, 2))))
            a64 = self.action64
        nz = None
        if self.kind != native.ENV_MAZE:
            if noise is None:
                raise ValueError("navigation envs need the dynamics noise of this step")
            self.noise.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(noise, np.float64).reshape(self.n, 2).T)))
            nz = self.noise
        native.env_step(self.cfg, self.action, self.action, self.state, self.ep_steps, self.ep_return, self.counters,
                        noise=nz, out_next_state=self.o_next, out_reward=self.o_reward, out_done=self.o_flags[0],
                        out_constraint=self.o_flags[1], out_success=self.o_flags[2], action_f64=a64)
        ns = self.o_next.cpu().numpy().T.copy()
        r = self.o_reward.cpu().numpy().copy()
        f = self.o_flags.cpu().numpy().astype(bool)
        return ns, r, f[0], f[1], f[2]
