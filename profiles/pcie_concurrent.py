"""Host-side ceiling of the end-to-end path on a multi-GPU box: pinned D2H (5.25 MB, the e2e step's result) and H2D
(1 MB, its actions) bandwidth per GPU when 1, 2, 4, 8 GPUs copy AT THE SAME TIME (one process per GPU, torchrun;
start aligned by a barrier).  Says whether the aggregate is bounded by the host (memory / PCIe root) or scales.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/pcie_concurrent.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym-macm_b200"))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
numa = "--numa" in sys.argv
cpus = None
if numa:
    from gym_macm.dist import bind_to_gpu_numa
    cpus = bind_to_gpu_numa(local)
torch.cuda.set_device(local)
dist.init_process_group("gloo")
D2H, H2D = 5246976, 1048576
h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory()
h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory()
d_out = torch.empty(D2H, dtype=torch.uint8, device="cuda")
d_in = torch.empty(H2D, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
rows = []
n = 1
while n <= world:
    active = rank < n
    for rep in range(2):
        dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        if active:
            for _ in range(200):
                with torch.cuda.stream(s1):
                    h_out.copy_(d_out, non_blocking=True)
                with torch.cuda.stream(s2):
                    d_in.copy_(h_in, non_blocking=True)
            torch.cuda.synchronize()
        el = time.perf_counter() - t
    gb = torch.tensor([200 * (D2H + H2D) / el / 1e9 if active else 0.0], dtype=torch.float64)
    dist.all_reduce(gb)
    if rank == 0:
        rows.append({"gpus_copying": n, "aggregate_GBps": round(float(gb), 1), "per_gpu_GBps": round(float(gb) / n, 1)})
    n *= 2
if rank == 0:
    print(json.dumps({"numa_bound": bool(cpus), "copy": "5.25 MB D2H + 1 MB H2D per step, both directions in flight", "rows": rows}))
dist.destroy_process_group()
