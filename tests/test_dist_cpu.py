"""Host-side logic of the multi-GPU path, on CPU with the gloo backend and world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions():
    from gym_macm.dist import shard_range
    for total in (1, 7, 8, 4096, 16385):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_macm.dist import all_gather_envs, scatter_actions, shard_range
    start, count = shard_range(total, rank, world)
    N = 3
    env_ids = torch.arange(start, start + count)
    # what a shard would hold: obs [E_local, N, 4], rewards [E_local, N], done [E_local]
    obs = (env_ids.view(-1, 1, 1) * 100 + torch.arange(N).view(1, -1, 1) * 10 + torch.arange(4).view(1, 1, -1)).float()
    local = {"obs": obs, "rewards": obs[..., 0] * 0.5, "done": (env_ids % 2).to(torch.uint8)}
    full = all_gather_envs(local, total)
    g = torch.arange(total)
    want = (g.view(-1, 1, 1) * 100 + torch.arange(N).view(1, -1, 1) * 10 + torch.arange(4).view(1, 1, -1)).float()
    ok = torch.equal(full["obs"], want) and torch.equal(full["rewards"], want[..., 0] * 0.5) and \
        torch.equal(full["done"], (g % 2).to(torch.uint8))
    # learner (rank 0) scatters actions for all envs
    acts_full = (torch.arange(total * N * 4) % 3).to(torch.uint8).view(total, N, 4)
    mine = scatter_actions(acts_full if rank == 0 else None, total, src=0,
                           like=torch.empty((total, N, 4), dtype=torch.uint8))
    ok = ok and torch.equal(mine, acts_full[start:start + count])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_and_scatter_world2_gloo(total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_peer_buffer_views_and_numa_binding_without_a_gpu():
    from gym_macm.dist import PeerBuffer, bind_to_gpu_numa
    # a mapped array of another process: slicing the env axis moves the address by whole rows
    b = PeerBuffer(4096, (10, 3, 4), torch.float32)
    v = b[2:5]
    assert v.data_ptr() == 4096 + 2 * 3 * 4 * 4 and v.shape == (3, 3, 4) and v.dtype == torch.float32
    assert b[7:].shape == (3, 3, 4) and b[7:].data_ptr() == 4096 + 7 * 48
    u = PeerBuffer(64, (5,), torch.uint8)[1:2]
    assert u.data_ptr() == 65 and u.shape == (1,)
    # no NVML device here: the binding helper declines quietly and leaves the affinity alone
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa(0) is None
    assert os.sched_getaffinity(0) == before
