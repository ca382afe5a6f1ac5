#!/usr/bin/env python
"""Throughput of the Flock step against batch size, block shape and the number of streams the independent
batches of the rotation are spread over (batches of a rotation share nothing, so their launches may overlap).
Device-timed like bench.py; an experiment driver, not a bench number.

    python profiles/concurrency_probe.py [quick]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import gym_macm

dev = torch.device("cuda", 0)
N, SETTLE, POOL = 64, 64, 13


def run(E, rot, streams, threads, steps):
    if threads:
        os.environ["MACM_BLOCK_THREADS"] = str(threads)
    else:
        os.environ.pop("MACM_BLOCK_THREADS", None)
    sims = [gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=1234 + r) for r in range(rot)]
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    acts = torch.zeros((POOL, E, N, 4), dtype=torch.uint8, device=dev)
    acts[..., :3] = torch.randint(0, 3, (POOL, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
    ss = [torch.cuda.Stream(device=dev) for _ in range(streams)]
    main = torch.cuda.current_stream(dev)

    def loop(n):
        for k in range(n):
            r = k % rot
            with torch.cuda.stream(ss[r % streams]):
                sims[r].engine.step(acts[(k // rot + 7 * r) % POOL])

    torch.cuda.synchronize()
    loop(SETTLE * rot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in ss:
        s.wait_stream(main)
    loop(steps)
    for s in ss:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    info = sims[0].engine.info
    out = {"envs": E, "rot": rot, "streams": streams, "threads_per_block": info.threads_per_block, "blocks": info.blocks,
           "us_per_step": 1e3 * ms, "us_per_4096_envs": 1e3 * ms * 4096 / E, "agent_steps_per_sec": E * N / (ms * 1e-3)}
    print(json.dumps(out), flush=True)
    for s in sims:
        s.close()
    del sims, acts
    torch.cuda.empty_cache()


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for threads in (896, 128):
    for streams in (1, 2, 3, 4, 8):
        run(4096, 16, streams, threads, 480)
if not quick:
    for E, rot in ((1024, 32), (2048, 32), (8192, 8), (16384, 4), (32768, 2), (65536, 2)):
        for threads in (896, 128):
            run(E, rot, 1, threads, max(64, 480 * 4096 // E))
            if E < 4096:
                run(E, rot, 4, threads, max(64, 480 * 4096 // E))
