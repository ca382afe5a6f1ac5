"""PeerGather: the step kernel of every rank writes its outputs into the learner rank's buffers (CUDA IPC peer
mapping).  Two processes share cuda:0 here (gloo for the rendezvous), which exercises the same IPC path as two
GPUs of one box; the result must equal the shards' own output buffers put side by side."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import gym_macm
        from gym_macm.dist import PeerGather, shard_range
        total, N = 24, 16
        start, count = shard_range(total, rank, world)
        torch.cuda.set_device(0)
        env = gym_macm.BatchedFlock(count, n_agents=[N], device="cuda:0", seed=3, env_index_base=start, start_spread=6.0)
        pg = PeerGather(env, total, learner=0, names=("obs", "nn_idx", "rewards", "collided", "done"))
        g = torch.Generator(device="cuda:0")
        g.manual_seed(5)
        ok = True
        for k in range(12):
            a_all = torch.zeros((total, N, 4), dtype=torch.uint8, device="cuda:0")
            a_all[..., :3] = torch.randint(0, 3, (total, N, 3), generator=g, device="cuda:0", dtype=torch.uint8)
            pg.step(a_all[start:start + count].contiguous())
            pg.fence()
            # every shard's own bound buffers, gathered the slow way, are the reference
            for n in pg.names:
                mine = env.state[n].cpu()
                parts = [torch.empty((shard_range(total, r, world)[1],) + tuple(mine.shape[1:]), dtype=mine.dtype)
                         for r in range(world)]
                dist.all_gather(parts, mine)
                if rank == 0:
                    ok = ok and torch.equal(pg.gathered()[n].cpu(), torch.cat(parts))
            dist.barrier()
        q.put((rank, bool(ok)))
        dist.barrier()
        env.close()
        del pg
        dist.destroy_process_group()
    except Exception as ex:   # report instead of hanging the parent
        q.put((rank, repr(ex)))


def test_peer_gather_two_processes_one_gpu():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
