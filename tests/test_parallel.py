"""The multi-GPU path on CPU: env sharding arithmetic and the observation all-gather, run
as a real 2-process torch.distributed job over gloo (rendezvous on 127.0.0.1)."""

import os
import socket
import subprocess
import sys
import textwrap

import pytest

from reinfocus_b200 import parallel

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 2), (4096, 8), (13, 4), (5, 8), (0, 3)])
def test_shard_bounds_partition_the_envs(n, world):
    bounds = [parallel.shard_bounds(n, world, r) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (_, last), (first, _) in zip(bounds, bounds[1:]):
        assert last == first
    sizes = [last - first for first, last in bounds]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


WORKER = textwrap.dedent("""
    import sys
    sys.path.insert(0, {repo!r})
    import torch
    import torch.distributed as dist
    from reinfocus_b200 import parallel

    rank, world, _ = parallel.init_from_env("gloo")
    assert world == 2 and dist.get_backend() == "gloo"
    for num_envs in (8, 7):
        first, last = parallel.shard_bounds(num_envs, world, rank)
        # each rank "observes" its own envs: value = global env index (+ column offset)
        local = torch.arange(first, last, dtype=torch.float64)[:, None] + torch.tensor([[0.0, 0.5]])
        gathered = parallel.gather_observations(local, num_envs)
        want = torch.arange(num_envs, dtype=torch.float64)[:, None] + torch.tensor([[0.0, 0.5]])
        assert gathered.shape == want.shape and torch.equal(gathered, want), (rank, gathered)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 2.0
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_observation_gather_over_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for rank, proc in enumerate(procs):
        out, _ = proc.communicate(timeout=180)
        assert proc.returncode == 0, out
        assert f"rank {rank} ok" in out


PPO_WORKER = textwrap.dedent("""
    import sys
    sys.path.insert(0, {repo!r})
    import torch
    import torch.distributed as dist
    from examples import ppo
    from reinfocus_b200 import parallel

    rank, world, _ = parallel.init_from_env("gloo")
    cfg = ppo.PPOConfig()
    cfg.batch_size, cfg.n_epochs = 64, 2
    first, last = parallel.shard_bounds(5, world, rank)  # 3 envs on rank 0, 2 on rank 1
    n, steps, dim = last - first, 32, 20
    assert ppo.minibatch_count(n * steps, cfg.batch_size, True) == 1  # min(96, 64) samples -> 1
    torch.manual_seed(0)
    policy = ppo.ActorCritic(dim, 13, cfg.net_arch)
    optimizer = torch.optim.Adam(policy.parameters(), lr=1e-3)
    generator = torch.Generator().manual_seed(100 + rank)
    data = {{"obs": torch.randn(steps, n, dim, generator=generator),
            "act": torch.randint(0, 13, (steps, n), generator=generator),
            "logp": -torch.rand(steps, n, generator=generator) - 2.0,
            "adv": torch.randn(steps, n, generator=generator),
            "ret": torch.randn(steps, n, generator=generator)}}
    stats = ppo.ppo_update(policy, optimizer, data, cfg, torch.device("cpu"))
    assert all(v == v for v in stats.values())
    # averaged gradients on identical initial weights: the replicas must stay identical
    flat = torch.cat([p.detach().reshape(-1) for p in policy.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert torch.equal(gathered[0], gathered[1])
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_ppo_update_with_uneven_shards_over_gloo_world_size_2(tmp_path):
    """examples/ppo.py: ranks whose shards differ by one env (96 vs 64 samples, batch 64) must
    run the same number of minibatches - each ends in a gradient all-reduce - and keep their
    policy replicas identical."""

    script = tmp_path / "ppo_worker.py"
    script.write_text(PPO_WORKER.format(repo=REPO))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for rank, proc in enumerate(procs):
        out, _ = proc.communicate(timeout=300)
        assert proc.returncode == 0, out
        assert f"rank {rank} ok" in out
