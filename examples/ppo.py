"""A compact PPO (rollout collector + clipped-objective update) for the example vector env.

The reference trains through rl_zoo3 / stable-baselines3 (examples/train_agent.py:41-62 with
examples/ppo_tuned.yml); neither is installable offline, so this module restates the pieces
that configuration uses: frame stacking, running observation / reward normalisation
(VecNormalize), an MLP actor-critic with separate pi / vf towers, GAE, and the clipped
surrogate update. The env side is the B200 hot path: one VectorDiscreteSteps per GPU, every
step rendering all of that rank's envs in one launch. With torch.distributed initialised the
envs are sharded by index over the ranks, each rank acts on its own shard with a replicated
policy, observations are all-gathered for the buffer of the policy rank, and gradients are
all-reduced."""

import dataclasses
import time

import numpy
import torch
from torch import nn


@dataclasses.dataclass
class PPOConfig:
    """Field names follow rl_zoo3's yml keys (reference examples/ppo_tuned.yml:3-23)."""

    frame_stack: int = 5
    batch_size: int = 64
    n_steps: int = 32
    gamma: float = 0.9
    learning_rate: float = 3.338099093100241e-05
    ent_coef: float = 0.0018133869709102076
    clip_range: float = 0.2
    n_epochs: int = 20
    gae_lambda: float = 0.99
    max_grad_norm: float = 0.3
    vf_coef: float = 0.4969606569643988
    net_arch: tuple = (256, 256)
    normalize: bool = True

    @classmethod
    def from_yml(cls, path: str, env_id: str) -> "PPOConfig":
        import yaml

        with open(path) as f:
            raw = yaml.safe_load(f)[env_id]
        fields = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in raw.items() if k in fields})


class RunningMeanStd:
    """Welford-style running moments (what SB3's VecNormalize keeps)."""

    def __init__(self, shape=()):
        self.mean = numpy.zeros(shape, dtype=numpy.float64)
        self.var = numpy.ones(shape, dtype=numpy.float64)
        self.count = 1e-4

    def update(self, batch: numpy.ndarray):
        batch_mean, batch_var, batch_count = batch.mean(axis=0), batch.var(axis=0), batch.shape[0]
        delta = batch_mean - self.mean
        total = self.count + batch_count
        self.mean = self.mean + delta * batch_count / total
        m2 = self.var * self.count + batch_var * batch_count + delta**2 * self.count * batch_count / total
        self.var = m2 / total
        self.count = total


class ActorCritic(nn.Module):
    """MlpPolicy with net_arch=dict(pi=[...], vf=[...]), ReLU, no orthogonal init."""

    def __init__(self, obs_dim: int, n_actions: int, hidden=(256, 256)):
        super().__init__()

        def tower(out):
            layers, last = [], obs_dim
            for width in hidden:
                layers += [nn.Linear(last, width), nn.ReLU()]
                last = width
            return nn.Sequential(*layers, nn.Linear(last, out))

        self.pi = tower(n_actions)
        self.vf = tower(1)

    def forward(self, obs):
        return torch.distributions.Categorical(logits=self.pi(obs)), self.vf(obs).squeeze(-1)


class RolloutCollector:
    """Steps a vector env with the current policy and fills an on-policy buffer."""

    def __init__(self, env, policy: ActorCritic, config: PPOConfig, device):
        self.env, self.policy, self.cfg, self.device = env, policy, config, device
        self.n = env.num_envs
        obs_dim = env.single_observation_space.shape[0]
        self.stack = numpy.zeros((self.n, config.frame_stack, obs_dim), dtype=numpy.float32)
        self.obs_rms = RunningMeanStd((config.frame_stack * obs_dim,))
        self.ret_rms = RunningMeanStd(())
        self.returns = numpy.zeros(self.n)
        obs, _ = env.reset()
        self.stack[:] = 0
        self.stack[:, -1] = obs
        self.env_steps = 0

    def _normalised(self, update: bool):
        flat = self.stack.reshape(self.n, -1)
        if not self.cfg.normalize:
            return flat
        if update:
            self.obs_rms.update(flat)
        return numpy.clip((flat - self.obs_rms.mean) / numpy.sqrt(self.obs_rms.var + 1e-8), -10, 10).astype(
            numpy.float32)

    def collect(self):
        cfg, n, T = self.cfg, self.n, self.cfg.n_steps
        buf = {k: [] for k in ("obs", "act", "logp", "val", "rew", "done")}
        obs = self._normalised(update=True)
        for _ in range(T):
            with torch.no_grad():
                dist, value = self.policy(torch.from_numpy(obs).to(self.device))
                action = dist.sample()
                logp = dist.log_prob(action)
            act = action.cpu().numpy()
            next_obs, reward, terminated, truncated, _ = self.env.step(act)
            done = terminated | truncated
            self.env_steps += n
            # reward normalisation by the running std of the discounted return
            self.returns = self.returns * cfg.gamma + reward
            if cfg.normalize:
                self.ret_rms.update(self.returns)
                reward = numpy.clip(reward / numpy.sqrt(self.ret_rms.var + 1e-8), -10, 10)
            self.returns[done] = 0.0
            for key, value_ in (("obs", obs), ("act", act), ("logp", logp.cpu().numpy()),
                                ("val", value.cpu().numpy()), ("rew", reward.astype(numpy.float32)),
                                ("done", done)):
                buf[key].append(value_)
            # frame stack: shift, and clear the history of envs that were reset this step
            self.stack[:, :-1] = self.stack[:, 1:]
            self.stack[done, :-1] = 0
            self.stack[:, -1] = next_obs
            obs = self._normalised(update=True)
        with torch.no_grad():
            _, last_value = self.policy(torch.from_numpy(obs).to(self.device))
        data = {k: numpy.stack(v) for k, v in buf.items()}
        # GAE(lambda)
        adv = numpy.zeros((T, n), dtype=numpy.float32)
        last = numpy.zeros(n, dtype=numpy.float32)
        next_value = last_value.cpu().numpy()
        for t in reversed(range(T)):
            not_done = 1.0 - data["done"][t]
            delta = data["rew"][t] + cfg.gamma * next_value * not_done - data["val"][t]
            last = delta + cfg.gamma * cfg.gae_lambda * not_done * last
            adv[t] = last
            next_value = data["val"][t]
        data["adv"] = adv
        data["ret"] = adv + data["val"]
        return data


class DeviceRunningMeanStd:
    """RunningMeanStd on the GPU (float64 torch tensors, no host sync)."""

    def __init__(self, shape, device):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.count = 1e-4

    def update(self, batch: torch.Tensor):
        batch = batch.to(torch.float64)
        batch_mean, batch_var, batch_count = batch.mean(dim=0), batch.var(dim=0, unbiased=False), batch.shape[0]
        delta = batch_mean - self.mean
        total = self.count + batch_count
        self.mean = self.mean + delta * batch_count / total
        m2 = self.var * self.count + batch_var * batch_count + delta**2 * self.count * batch_count / total
        self.var = m2 / total
        self.count = total


class DeviceRolloutCollector:
    """RolloutCollector for a DeviceVectorEnvironment: observations, rewards, the frame stack,
    the running normalisation, the buffer and GAE all stay on the GPU; the policy's actions go
    to the env as a device tensor. Same arithmetic as RolloutCollector, in torch."""

    def __init__(self, env, policy: ActorCritic, config: PPOConfig, device):
        self.env, self.policy, self.cfg, self.device = env, policy, config, device
        self.n = env.num_envs
        obs_dim = env.single_observation_space.shape[0]
        self.stack = torch.zeros((self.n, config.frame_stack, obs_dim), dtype=torch.float32, device=device)
        self.obs_rms = DeviceRunningMeanStd((config.frame_stack * obs_dim,), device)
        self.ret_rms = DeviceRunningMeanStd((), device)
        self.returns = torch.zeros(self.n, dtype=torch.float64, device=device)
        obs, _ = env.reset()
        self.stack[:, -1] = obs
        self.env_steps = 0

    def _normalised(self, update: bool):
        flat = self.stack.reshape(self.n, -1)
        if not self.cfg.normalize:
            return flat
        if update:
            self.obs_rms.update(flat)
        return torch.clamp((flat - self.obs_rms.mean) / torch.sqrt(self.obs_rms.var + 1e-8), -10, 10).to(
            torch.float32)

    def collect(self):
        cfg, n, T = self.cfg, self.n, self.cfg.n_steps
        buf = {k: [] for k in ("obs", "act", "logp", "val", "rew", "done")}
        obs = self._normalised(update=True)
        for _ in range(T):
            with torch.no_grad():
                dist, value = self.policy(obs)
                action = dist.sample()
                logp = dist.log_prob(action)
            next_obs, reward, terminated, truncated, _ = self.env.step(action)
            done = terminated | truncated
            self.env_steps += n
            self.returns = self.returns * cfg.gamma + reward
            if cfg.normalize:
                self.ret_rms.update(self.returns)
                reward = torch.clamp(reward / torch.sqrt(self.ret_rms.var + 1e-8), -10, 10)
            self.returns = torch.where(done, torch.zeros_like(self.returns), self.returns)
            for key, value_ in (("obs", obs), ("act", action), ("logp", logp), ("val", value),
                                ("rew", reward.to(torch.float32)), ("done", done)):
                buf[key].append(value_)
            self.stack[:, :-1] = self.stack[:, 1:].clone()
            self.stack[:, :-1] *= (~done).to(torch.float32)[:, None, None]
            self.stack[:, -1] = next_obs
            obs = self._normalised(update=True)
        with torch.no_grad():
            _, last_value = self.policy(obs)
        data = {k: torch.stack(v) for k, v in buf.items()}
        adv = torch.zeros((T, n), dtype=torch.float32, device=self.device)
        last = torch.zeros(n, dtype=torch.float32, device=self.device)
        next_value = last_value
        for t in reversed(range(T)):
            not_done = 1.0 - data["done"][t].to(torch.float32)
            delta = data["rew"][t] + cfg.gamma * next_value * not_done - data["val"][t]
            last = delta + cfg.gamma * cfg.gae_lambda * not_done * last
            adv[t] = last
            next_value = data["val"][t]
        data["adv"] = adv
        data["ret"] = adv + data["val"]
        return data


def minibatch_count(local_samples: int, batch_size: int, distributed: bool) -> int:
    """Minibatches per epoch, identical on every rank: ceil(smallest shard / batch size)."""

    smallest = local_samples
    if distributed:
        t = torch.tensor([local_samples], dtype=torch.int64,
                         device="cuda" if torch.distributed.get_backend() == "nccl" else "cpu")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
        smallest = int(t.item())
    return max(1, -(-smallest // batch_size))


def ppo_update(policy: ActorCritic, optimizer, data, cfg: PPOConfig, device, max_minibatches=None):
    """Clipped-surrogate PPO epochs over one rollout; returns the last losses."""

    flat = {k: torch.as_tensor(v).reshape((-1,) + tuple(v.shape[2:])).to(device) for k, v in data.items()
            if k in ("obs", "act", "logp", "adv", "ret")}
    size = flat["obs"].shape[0]
    distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
    # every rank must run the same number of minibatches (each ends in a collective): shards
    # can differ by one env (parallel.shard_bounds), so the count follows the smallest shard
    # and a rank with a larger one spreads its samples over the same number of minibatches
    n_batches = minibatch_count(size, cfg.batch_size, distributed)
    stats, batches = {}, 0
    for _ in range(cfg.n_epochs):
        order = torch.randperm(size, device=device)
        for chunk in torch.tensor_split(order, n_batches):
            idx = chunk
            dist, value = policy(flat["obs"][idx])
            adv = flat["adv"][idx]
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)
            ratio = torch.exp(dist.log_prob(flat["act"][idx]) - flat["logp"][idx])
            pg_loss = -torch.min(ratio * adv, torch.clamp(ratio, 1 - cfg.clip_range, 1 + cfg.clip_range) * adv).mean()
            vf_loss = nn.functional.mse_loss(value, flat["ret"][idx])
            entropy = dist.entropy().mean()
            loss = pg_loss + cfg.vf_coef * vf_loss - cfg.ent_coef * entropy
            optimizer.zero_grad()
            loss.backward()
            if distributed:
                # one reduction over the flattened gradients, not one per parameter
                grads = [param.grad for param in policy.parameters() if param.grad is not None]
                flat_grads = torch.cat([g.reshape(-1) for g in grads])
                torch.distributed.all_reduce(flat_grads)
                flat_grads /= torch.distributed.get_world_size()
                offset = 0
                for g in grads:
                    g.copy_(flat_grads[offset:offset + g.numel()].view_as(g))
                    offset += g.numel()
            nn.utils.clip_grad_norm_(policy.parameters(), cfg.max_grad_norm)
            optimizer.step()
            stats = {name: float(term.detach()) for name, term in (("loss", loss), ("pg", pg_loss), ("vf", vf_loss), ("entropy", entropy))}
            batches += 1
            if max_minibatches and batches >= max_minibatches:
                return stats
    return stats


def train(num_envs: int = 8, rollouts: int = 2, config: PPOConfig | None = None, seed: int = 0,
          max_minibatches: int | None = None, log=print, device_env: bool = False):
    """Collects ``rollouts`` rollouts of n_steps on VectorDiscreteSteps(num_envs) (per rank)
    and runs the PPO update after each. Returns per-rollout statistics. ``device_env`` steps
    DeviceVectorDiscreteSteps instead: the same env sequences, with the env step, the buffer
    and the normalisation on the GPU."""

    from examples import custom_environments
    from reinfocus_b200 import parallel
    from reinfocus_b200.environments import state_initializer

    cfg = config or PPOConfig()
    rank, world, local_rank = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    torch.manual_seed(seed)  # same policy weights on every rank
    first, last = parallel.shard_bounds(num_envs, world, rank)
    initializer = state_initializer.RangedInitializer([[custom_environments.ENDS]] * 2, seed=seed + rank)
    if device_env:
        env = custom_environments.DeviceVectorDiscreteSteps(
            max_episode_steps=20, num_envs=last - first, initializer=initializer)
    else:
        env = custom_environments.VectorDiscreteSteps(
            max_episode_steps=20, num_envs=last - first, initializer=initializer)
    obs_dim = env.single_observation_space.shape[0] * cfg.frame_stack
    policy = ActorCritic(obs_dim, env.single_action_space.n, cfg.net_arch).to(device)
    optimizer = torch.optim.Adam(policy.parameters(), lr=cfg.learning_rate, eps=1e-5)
    collector = (DeviceRolloutCollector if device_env else RolloutCollector)(env, policy, cfg, device)
    torch.manual_seed(seed + 1000 + rank)  # different action samples per rank
    history = []
    for i in range(rollouts):
        t0 = time.perf_counter()
        data = collector.collect()
        torch.cuda.synchronize()
        t_collect = time.perf_counter() - t0
        if world > 1:
            # the policy rank sees every env's observations (north star: NCCL obs gather)
            parallel.gather_observations(torch.as_tensor(data["obs"][-1]).to(device), num_envs)
        t1 = time.perf_counter()
        stats = ppo_update(policy, optimizer, data, cfg, device, max_minibatches)
        torch.cuda.synchronize()
        entry = {"rollout": i, "env_steps": (last - first) * cfg.n_steps,
                 "collect_s": t_collect, "update_s": time.perf_counter() - t1,
                 "env_steps_per_s": (last - first) * cfg.n_steps / t_collect * world,
                 "mean_reward": float(data["rew"].mean()), "device_env": device_env, **stats}
        history.append(entry)
        if rank == 0:
            log(entry)
    if world > 1:
        torch.distributed.destroy_process_group()
    return history
