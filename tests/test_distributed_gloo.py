"""world_size-2 gloo test of the only collective on this path (episode-statistics all-reduce) and of
the env sharding, with the CPU oracle standing in for each rank's shard."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total_envs, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200._capi import default_config
    from torchdriveenv_b200.distributed import reduce_episode_stats, shard_range
    lo, hi = shard_range(total_envs, rank, world)
    ss = S.roundabout(8)
    env = O.OracleEnvSet(default_config(num_envs=hi - lo, max_agents=8, auto_reset=1, env_index_offset=lo), ss.pack(8))
    env.reset(seed=9)
    rng = np.random.default_rng(9)
    for _ in range(30):
        a = np.stack([rng.uniform(-1, 1, total_envs), rng.uniform(-0.3, 0.3, total_envs)], 1).astype(np.float32)
        env.step(a[lo:hi], render=False)
    total = reduce_episode_stats(env.stats)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([total, env.stats]))
    dist.barrier()
    dist.destroy_process_group()


def test_episode_stats_allreduce_world2(tmp_path, oracle):
    total_envs = 13
    mp.spawn(_worker, args=(2, _free_port(), total_envs, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    np.testing.assert_array_equal(r0[:16], r1[:16])                       # every rank holds the global sum
    np.testing.assert_allclose(r0[:16], r0[16:] + r1[16:], rtol=1e-12)
    # and it equals the unsharded run
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200._capi import default_config
    env = oracle.OracleEnvSet(default_config(num_envs=total_envs, max_agents=8, auto_reset=1), S.roundabout(8).pack(8))
    env.reset(seed=9)
    rng = np.random.default_rng(9)
    for _ in range(30):
        a = np.stack([rng.uniform(-1, 1, total_envs), rng.uniform(-0.3, 0.3, total_envs)], 1).astype(np.float32)
        env.step(a, render=False)
    np.testing.assert_allclose(r0[:16], env.stats, rtol=1e-12)
    assert r0[0] > 0 and r0[8] == 30 * total_envs
