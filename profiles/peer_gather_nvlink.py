"""NVLink traffic of the fused step+gather (gym_macm.dist.PeerGather), read from the GPUs' own link counters.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/peer_gather_nvlink.py

Rank 0 (the learner) reads `nvidia-smi nvlink -gt d` (data bytes received / transmitted per link, KiB) before and
after K PeerGather steps and compares the received bytes with what the shards' kernels must have stored into its
buffers: (world - 1) x (E x N x (16 + 4) + E) bytes per step.  The step kernel's peer stores are the only traffic on
the links during the window (no NCCL call inside it)."""
import json
import os
import re
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import gym_macm  # noqa: E402
from gym_macm.dist import PeerGather  # noqa: E402


def link_bytes(index):
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True).stdout
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    return rx * 1024, tx * 1024, len(re.findall(r"Data Rx", out))


rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
E, N, K = 4096, 64, 2000
env = gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=7 + rank, env_index_base=rank * E)
acts = torch.zeros((17, E, N, 4), dtype=torch.uint8, device=dev)
acts[..., :3] = torch.randint(0, 3, (17, E, N, 3), device=dev, dtype=torch.uint8)
peer = PeerGather(env, world * E, learner=0, names=("obs", "rewards", "done"))
for k in range(64):
    peer.step(acts[k % 17])
peer.fence()
b0 = link_bytes(0) if rank == 0 else None
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(K):
    peer.step(acts[k % 17])
e1.record()
peer.fence()
if rank == 0:
    b1 = link_bytes(0)
    expect = (world - 1) * (E * N * 20 + E) * K
    print(json.dumps({"gpus": world, "steps": K, "ms_per_step": e0.elapsed_time(e1) / K,
                      "links_reporting": b1[2], "rx_bytes_counted": b1[0] - b0[0], "tx_bytes_counted": b1[1] - b0[1],
                      "rx_bytes_expected_payload": expect, "rx_over_payload": (b1[0] - b0[0]) / expect if expect else None,
                      "rx_GBps": (b1[0] - b0[0]) / (e0.elapsed_time(e1) * 1e-3) / 1e9}))
peer.close()
dist.destroy_process_group()
