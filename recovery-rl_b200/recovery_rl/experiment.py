"""Experiment wrapper (reference recovery_rl/experiment.py:42-577): seeding, env / agent construction, offline
data, Q_risk pre-training, train / test rollouts, composite action selection, pickle logging -- same class
and method names, same log files (`args.pkl`, `run_stats.pkl` with "train_stats" / "test_stats").

Two execution modes behind the same surface:
  * --num_envs 1 (default): the reference's loop, statement for statement, on the drop-in classes
    (env.step, SAC, QRiskWrapper, ReplayMemory) whose arithmetic runs in the CUDA kernels;
  * --num_envs N > 1: the vectorised engine (recovery_rl/engine.py) -- N env copies per GPU, the whole
    step replayed as one CUDA graph, per-step host traffic only for the env copies that are logged.
Model-based recovery (PETS ensemble + CEM, recovery_rl/MPC.py) runs on the --num_envs 1 path; visual MPC, image
observations and task demos are outside this build.
"""
import datetime
import itertools
import os
import os.path as osp
import pickle

import numpy as np
import torch

from recovery_rl.sac import SAC
from recovery_rl.replay_memory import ReplayMemory, ConstraintReplayMemory
from recovery_rl.utils import linear_schedule, recovery_config_setup
from env.make_utils import register_env, make_env


def torchify(x):
    return torch.FloatTensor(x).to('cuda')


class Experiment:
    def __init__(self, exp_cfg):
        self.exp_cfg = exp_cfg
        # one process per GPU: ranks > 0 log into their own directory (same naming, "_rank<r>" appended) so that ranks
        # started in the same second never write the same args.pkl / run_stats.pkl
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        suffix = self.exp_cfg.logdir_suffix + ("_rank%d" % rank if world > 1 and rank > 0 else "")
        self.logdir = os.path.join(
            self.exp_cfg.logdir, '{}_SAC_{}_{}_{}'.format(datetime.datetime.now().strftime("%Y-%m-%d_%H-%M-%S"),
                                                          self.exp_cfg.env_name, self.exp_cfg.policy, suffix))
        if not os.path.exists(self.logdir):
            os.makedirs(self.logdir)
        print("LOGDIR: ", self.logdir)
        pickle.dump(self.exp_cfg, open(os.path.join(self.logdir, "args.pkl"), "wb"))
        self.num_envs = int(getattr(self.exp_cfg, "num_envs", 1))
        self.mb_recovery = self.exp_cfg.use_recovery and not (self.exp_cfg.MF_recovery or self.exp_cfg.Q_sampling_recovery)
        if self.mb_recovery and self.exp_cfg.vismpc_recovery:
            raise NotImplementedError("visual MPC (image observations) is outside this build")
        if self.exp_cfg.task_demos:
            raise NotImplementedError("--task_demos is outside this build (DESIGN.md, 'next')")

        self.experiment_setup()

        self.total_numsteps = 0
        self.updates = 0
        self.num_constraint_violations = 0
        self.num_unsafe_transitions = 0
        self.num_viols = 0
        self.num_successes = 0
        self.viol_and_recovery = 0
        self.viol_and_no_recovery = 0
        self.task_demos = self.exp_cfg.task_demos
        if self.num_envs == 1:
            self.memory = ReplayMemory(self.exp_cfg.replay_size, self.exp_cfg.seed)
            self.recovery_memory = ConstraintReplayMemory(self.exp_cfg.safe_replay_size, self.exp_cfg.seed)
        self.all_ep_data = []
        self.constraint_demo_data, self.task_demo_data, self.obs_seqs, self.ac_seqs, self.constraint_seqs = \
            self.get_offline_data()
        if self.exp_cfg.nu_schedule:
            self.nu_schedule = linear_schedule(self.exp_cfg.nu_start, self.exp_cfg.nu_end, self.exp_cfg.num_eps)
        else:
            self.nu_schedule = linear_schedule(self.exp_cfg.nu, self.exp_cfg.nu, 0)

    # ------------------------------------------------------------------------------------------------
    def experiment_setup(self):
        torch.manual_seed(self.exp_cfg.seed)
        np.random.seed(self.exp_cfg.seed)
        if self.mb_recovery and self.num_envs == 1:                # experiment.py:92-99
            from recovery_rl.MPC import MPC
            register_env(self.exp_cfg.env_name)
            cfg = recovery_config_setup(self.exp_cfg, self.logdir)
            env = cfg.ctrl_cfg.env
            recovery_policy = MPC(cfg.ctrl_cfg, seed=self.exp_cfg.seed)
        else:
            recovery_policy = None
            register_env(self.exp_cfg.env_name)
            env = make_env(self.exp_cfg.env_name)
        self.env = env
        self.recovery_policy = recovery_policy
        self.env.seed(self.exp_cfg.seed)
        self.env.action_space.seed(self.exp_cfg.seed)
        if self.num_envs == 1:
            self.agent = self.agent_setup(env)
            self.engine = None
            if self.mb_recovery:                                   # experiment.py:164-167
                recovery_policy.update_value_func(self.agent.safety_critic)
        else:
            self.agent = None
            self.engine = self.engine_setup()

    def agent_setup(self, env):
        return SAC(env.observation_space, env.action_space, self.exp_cfg, self.logdir,
                   tmp_env=make_env(self.exp_cfg.env_name))

    def engine_setup(self):
        from recovery_rl.engine import VecEngine
        c = self.exp_cfg
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        pg = None
        if world > 1:
            import torch.distributed as dist
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            pg = dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        if c.use_constraint_sampling and (c.use_recovery or c.policy != "Gaussian"):
            raise NotImplementedError("--use_constraint_sampling (SQRL) together with --use_recovery / --policy Deterministic runs "
                                      "on the --num_envs 1 path only (no reference script combines them)")
        eng = VecEngine(c.env_name, self.num_envs, batch_size=c.batch_size, replay_size=c.replay_size,
                        safe_replay_size=c.safe_replay_size, gamma=c.gamma, alpha=c.alpha, tau=c.tau, lr=c.lr,
                        gamma_safe=c.gamma_safe, tau_safe=c.tau_safe, eps_safe=c.eps_safe,
                        target_update_interval=c.target_update_interval, use_recovery=c.use_recovery,
                        mf_recovery=c.MF_recovery, pos_fraction=c.pos_fraction,
                        disable_online_updates=c.disable_online_updates,
                        constraint_reward_penalty=c.constraint_reward_penalty, start_steps=c.start_steps, seed=c.seed,
                        device=dev, rank=rank, world_size=world, process_group=pg,
                        log_outputs=getattr(c, "log_envs", 1) > 0, use_tensor_cores=getattr(c, "tensor_cores", 1),
                        dgd=c.DGD_constraints, update_nu=c.update_nu, rcpo=c.RCPO,
                        auto_alpha=bool(c.automatic_entropy_tuning), nu=c.nu, lambda_rcpo=c.lambda_RCPO,
                        disable_action_relabeling=c.disable_action_relabeling, mb_recovery=self.mb_recovery,
                        mpc_popsize=getattr(c, "mpc_popsize", None), constraint_sampling=c.use_constraint_sampling,
                        q_sampling_recovery=c.Q_sampling_recovery, add_both_transitions=c.add_both_transitions,
                        deterministic=c.policy != "Gaussian")
        eng.init_agent()          # same torch seed on every rank -> identical replicas
        return eng

    def get_offline_data(self):
        """experiment.py:177-246.  Navigation: env.transition_function(n).  Maze: the reference loads a pickle of
        MuJoCo demos (demos/maze/constraint_demos.pkl); it is used when present, else the demos are regenerated
        with the restated physics (env/maze.py get_offline_data)."""
        task_demo_data = None
        if 'maze' in self.exp_cfg.env_name:
            path = osp.join("demos", self.exp_cfg.env_name, "constraint_demos.pkl")
            if osp.exists(path):
                constraint_demo_data = pickle.load(open(path, "rb"))
            else:
                constraint_demo_data = self.env.transition_function(self.exp_cfg.num_unsafe_transitions)
        else:
            constraint_demo_data = self.env.transition_function(self.exp_cfg.num_unsafe_transitions)
        return constraint_demo_data, task_demo_data, [], [], []

    def pretrain_critic_recovery(self):
        """experiment.py:261-296."""
        demos = self.constraint_demo_data[:self.exp_cfg.num_unsafe_transitions]
        self.num_unsafe_transitions = len(demos)
        self.num_constraint_violations += int(sum(int(t[2]) for t in demos))
        print("Number of Constraint Transitions: ", self.num_unsafe_transitions)
        print("Number of Constraint Violations: ", self.num_constraint_violations)
        bs = min(self.exp_cfg.batch_size, len(self.constraint_demo_data))
        if self.engine is not None:
            self.engine.push_offline(demos)
            self.engine.pretrain_qrisk(self.exp_cfg.critic_safe_pretraining_steps, n_demos=len(self.constraint_demo_data))
            if self.mb_recovery:
                self.engine.train_mb(demos)                        # experiment.py:298-305
            return
        for transition in demos:
            self.recovery_memory.push(*transition)
        for i in range(self.exp_cfg.critic_safe_pretraining_steps):
            if i % 100 == 0:
                print("CRITIC SAFE UPDATE STEP: ", i)
            self.agent.safety_critic.update_parameters(memory=self.recovery_memory, policy=self.agent.policy,
                                                       batch_size=bs)
        if self.mb_recovery:                                       # experiment.py:298-305: train the PETS recovery policy
            self.train_MB_recovery(np.array([d[0] for d in demos]), np.array([d[1] for d in demos]),
                                   np.array([d[3] for d in demos]), epochs=50)

    def train_MB_recovery(self, states, actions, next_states=None, epochs=50):
        """experiment.py:251-259."""
        if next_states is not None:
            self.recovery_policy.train(states, actions, random=True, next_obs=next_states, epochs=epochs)
        else:
            self.recovery_policy.train(states, actions)

    # ------------------------------------------------------------------------------------------------
    def run(self):
        if not self.exp_cfg.disable_offline_updates and (self.exp_cfg.use_recovery or self.exp_cfg.DGD_constraints
                                                          or self.exp_cfg.RCPO):
            self.pretrain_critic_recovery()
        resume = getattr(self.exp_cfg, "resume", "")
        if self.engine is not None:
            return self.run_vectorised()
        if resume:
            self.agent.load(resume)
        train_rollouts = []
        test_rollouts = []
        ck = int(getattr(self.exp_cfg, "checkpoint_every", 0))
        for i_episode in itertools.count(1):
            train_rollouts.append(self.get_train_rollout(i_episode))
            if ck and i_episode % ck == 0:
                self.agent.save(osp.join(self.logdir, "checkpoint.pt"))
            if i_episode % 10 == 0 and self.exp_cfg.eval:
                test_rollouts.append(self.get_test_rollout(i_episode))
            if self.total_numsteps > self.exp_cfg.num_steps or i_episode > self.exp_cfg.num_eps:
                break
            self.dump_logs(train_rollouts, test_rollouts)

    def get_train_rollout(self, i_episode):
        """experiment.py:379-491, statement for statement."""
        episode_reward = 0
        episode_steps = 0
        done = False
        state = self.env.reset()
        train_rollout_info = []
        ep_states = [state]
        ep_actions = []
        ep_constraints = []
        if i_episode % 10 == 0:
            print("SEED: ", self.exp_cfg.seed)
            print("LOGDIR: ", self.logdir)
        while not done:
            if len(self.memory) > self.exp_cfg.batch_size:
                for i in range(self.exp_cfg.updates_per_step):
                    self.agent.update_parameters(self.memory, min(self.exp_cfg.batch_size, len(self.memory)),
                                                 self.updates, safety_critic=self.agent.safety_critic,
                                                 nu=self.nu_schedule(i_episode))
                    if not self.exp_cfg.disable_online_updates and len(self.recovery_memory) > self.exp_cfg.batch_size \
                            and (self.num_viols + self.num_constraint_violations) / self.exp_cfg.batch_size > \
                            self.exp_cfg.pos_fraction:
                        self.agent.safety_critic.update_parameters(memory=self.recovery_memory, policy=self.agent.policy,
                                                                   batch_size=self.exp_cfg.batch_size, plot=0)
                    self.updates += 1
            action, real_action, recovery_used = self.get_action(state)
            next_state, reward, done, info = self.env.step(real_action)
            info['recovery'] = recovery_used
            train_rollout_info.append(info)
            episode_steps += 1
            episode_reward += reward
            self.total_numsteps += 1
            if info['constraint']:
                reward -= self.exp_cfg.constraint_reward_penalty
            mask = float(not done)
            done = done or episode_steps == self.env._max_episode_steps
            if not self.exp_cfg.disable_action_relabeling:
                self.memory.push(state, action, reward, next_state, mask)
            else:
                self.memory.push(state, real_action, reward, next_state, mask)
            if self.exp_cfg.use_recovery or self.exp_cfg.DGD_constraints or self.exp_cfg.RCPO:
                self.recovery_memory.push(state, real_action, info['constraint'], next_state, mask)
                if recovery_used and self.exp_cfg.add_both_transitions:
                    self.memory.push(state, real_action, reward, next_state, mask)
            state = next_state
            ep_states.append(state)
            ep_actions.append(real_action)
            ep_constraints.append([info['constraint']])
        if info['constraint']:
            self.num_viols += 1
            if info['recovery']:
                self.viol_and_recovery += 1
            else:
                self.viol_and_no_recovery += 1
        self.num_successes += int(info['success'])
        # experiment.py:464-478: update the model-based recovery policy with the online data
        if self.exp_cfg.use_recovery and not self.exp_cfg.disable_online_updates:
            self.all_ep_data.append({'obs': np.array(ep_states), 'ac': np.array(ep_actions),
                                     'constraint': np.array(ep_constraints)})
            if i_episode % self.exp_cfg.recovery_policy_update_freq == 0 and self.mb_recovery:
                self.train_MB_recovery([ep_data['obs'] for ep_data in self.all_ep_data],
                                       [ep_data['ac'] for ep_data in self.all_ep_data])
                self.all_ep_data = []
        print("Episode: {}, total numsteps: {}, episode steps: {}, reward: {}".format(
            i_episode, self.total_numsteps, episode_steps, round(episode_reward, 2)))
        print("Num Violations So Far: %d" % self.num_viols)
        print("Violations with Recovery: %d" % self.viol_and_recovery)
        print("Violations with No Recovery: %d" % self.viol_and_no_recovery)
        print("Num Successes So Far: %d" % self.num_successes)
        return train_rollout_info

    def get_test_rollout(self, i_episode):
        """experiment.py:493-538 without the MuJoCo gif rendering."""
        test_rollout_info = []
        state = self.env.reset()
        episode_reward = 0
        episode_steps = 0
        done = False
        while not done:
            action, real_action, recovery_used = self.get_action(state, train=False)
            next_state, reward, done, info = self.env.step(real_action)
            info['recovery'] = recovery_used
            done = done or episode_steps == self.env._max_episode_steps
            test_rollout_info.append(info)
            episode_reward += reward
            episode_steps += 1
            state = next_state
        print("----------------------------------------")
        print("Avg. Reward: {}".format(round(episode_reward, 2)))
        print("----------------------------------------")
        return test_rollout_info

    def dump_logs(self, train_rollouts, test_rollouts):
        data = {"test_stats": test_rollouts, "train_stats": train_rollouts}
        with open(osp.join(self.logdir, "run_stats.pkl"), "wb") as f:
            pickle.dump(data, f)

    def get_action(self, state, train=True):
        """experiment.py:546-577."""
        def recovery_thresh(state, action):
            if not self.exp_cfg.use_recovery:
                return False
            critic_val = self.agent.safety_critic.get_value(torchify(state).unsqueeze(0), torchify(action).unsqueeze(0))
            if critic_val > self.exp_cfg.eps_safe:
                return True
            return False

        if self.exp_cfg.start_steps > self.total_numsteps and train:
            action = self.env.action_space.sample()
        elif train:
            action = self.agent.select_action(state)
        else:
            action = self.agent.select_action(state, eval=True)
        if recovery_thresh(state, action):
            recovery = True
            if self.exp_cfg.MF_recovery or self.exp_cfg.Q_sampling_recovery:
                real_action = self.agent.safety_critic.select_action(state)
            else:
                real_action = self.recovery_policy.act(state, 0)
        else:
            recovery = False
            real_action = np.copy(action)
        return action, real_action, recovery

    # ------------------------------------------------------------------------------------------------
    def run_vectorised(self, report_every=50):
        """N env copies per GPU.  Termination as in run(): total_numsteps > num_steps or episodes > num_eps.
        `train_stats` keeps the reference's schema (a list of episodes, each a list of per-step info dicts) for
        the first --log_envs env copies; the global counters go to `vec_stats`."""
        eng = self.engine
        c = self.exp_cfg
        eng.reset()
        if getattr(c, "resume", ""):
            eng.load(c.resume)             # env state, replay, sampler and counters continue where the checkpoint stopped
        eng.capture()
        ck_dir = self.logdir
        if eng.world > 1:                  # all shards of a checkpoint live next to each other, in rank 0's logdir
            import torch.distributed as dist
            box = [self.logdir]
            dist.broadcast_object_list(box, src=0, group=eng.pg)
            ck_dir = box[0]
        k = min(max(int(getattr(c, "log_envs", 1)), 0), eng.n)
        train_rollouts, open_eps, vec_stats, test_rollouts = [], [[] for _ in range(k)], [], []
        # experiment.py:372-374: one test rollout after every 10th training episode -- here after every 10 episodes PER ENV
        # COPY (10 * N * world finished episodes), max(1, log_envs) eval copies at once, rank 0 only (replicas are identical)
        eval_every = 10 * eng.n * eng.world
        next_eval = eval_every
        step = 0
        while True:
            if k:
                prev = eng.state[:, :k].t().cpu().numpy().copy()
            eng.replay()
            step += 1
            if k:
                ns = eng.out_next[:, :k].t().cpu().numpy()
                rew = eng.out_reward[:k].cpu().numpy()
                done = eng.out_done[:k].cpu().numpy()
                cons = eng.out_cons[:k].cpu().numpy()
                succ = eng.out_succ[:k].cpu().numpy()
                act = eng.action_real[:k].cpu().numpy()
                rec = eng.recovery[:k].cpu().numpy()
                for i in range(k):
                    open_eps[i].append({"constraint": int(cons[i]), "reward": float(rew[i]), "state": prev[i],
                                        "next_state": ns[i], "action": act[i], "success": bool(succ[i]),
                                        "recovery": bool(rec[i])})
                    if done[i]:
                        train_rollouts.append(open_eps[i])
                        open_eps[i] = []
            if step % report_every == 0:
                cn = eng.read_counters()
                if cn["error"]:
                    raise RuntimeError("device-side error %d: %s" % (cn["error"], {
                        1: "sampler: sample larger than population (too few positives / negatives for one stratified batch)",
                        2: "a peer GPU did not reach the gradient barrier within 60 s (peer lost or stalled)"}.get(cn["error"], "unknown")))
                self.total_numsteps = cn["total_numsteps"]
                self.num_viols, self.num_successes = cn["num_viols"], cn["num_successes"]
                self.viol_and_recovery, self.viol_and_no_recovery = cn["viol_and_recovery"], cn["viol_and_no_recovery"]
                self.updates = cn["sac_updates"]
                eng.set_nu(self.nu_schedule(1 + cn["episodes"] * eng.world))     # experiment.py:406
                if self.mb_recovery and not c.disable_online_updates and \
                        (step // report_every) % max(1, c.recovery_policy_update_freq) == 0:
                    eng.train_mb()                                               # experiment.py:464-478
                if True:
                    vec_stats.append(cn)
                    if eng.rank == 0:
                        print("Vector step: {}, total numsteps: {}, episodes: {}, violations: {}, successes: {}".format(
                            step, cn["total_numsteps"], cn["episodes"], cn["num_viols"], cn["num_successes"]))
                    if c.eval and eng.rank == 0 and cn["episodes"] * eng.world >= next_eval:
                        test_rollouts.extend(eng.eval_rollout(max(1, k)))          # experiment.py:493-538
                        next_eval += eval_every * max(1, (cn["episodes"] * eng.world - next_eval) // eval_every + 1)
                    data = {"test_stats": test_rollouts, "train_stats": train_rollouts, "vec_stats": vec_stats}
                    with open(osp.join(self.logdir, "run_stats.pkl"), "wb") as f:
                        pickle.dump(data, f)
                ck = int(getattr(c, "checkpoint_every", 0))
                if ck and (step // report_every) % ck == 0:
                    # every rank writes its own shard (checkpoint.pt.rank<r>of<w>) into rank 0's naming scheme
                    eng.save(osp.join(ck_dir, "checkpoint.pt"))
                if cn["total_numsteps"] * eng.world > c.num_steps or cn["episodes"] * eng.world > c.num_eps:
                    break
        return vec_stats
