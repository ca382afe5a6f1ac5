#!/usr/bin/env python
"""Where one warp's time goes, phase by phase (SM cycles), for the bench workload.

Needs the instrumented build:  nvcc ... -DMACM_PHASE_TRACE -o profiles/_libmacm_trace.so  (see
`build_trace_lib`), loaded through MACM_LIB.  Run with E=148 to see the latency of a lone warp per SM
and with E=4096 to see the same phases under the full load."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "profiles", "_libmacm_trace.so")


def build_trace_lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    srcs = [os.path.join(ge.CSRC, f) for f in ("macm_kernels.cu", "macm_aux.cu", "macm_api.cu")]
    subprocess.check_call(["nvcc"] + ge.NVCC_FLAGS + ["-DMACM_PHASE_TRACE", "-I", os.path.join(ROOT, "include"), "-I", ge.CSRC,
                           "-o", LIB] + srcs)


if __name__ == "__main__":
    if sys.argv[1:] == ["build"]:
        build_trace_lib()
        sys.exit(0)
    os.environ["MACM_LIB"] = LIB
    import ctypes as C
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
    import gym_macm
    from gym_macm import _lib

    N = 64
    settle = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    dev = torch.device("cuda", 0)
    sim = gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=1234)
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    POOL = 61   # like bench.py: fresh random actions every step
    acts = torch.zeros((POOL, E, N, 4), dtype=torch.uint8, device=dev)
    acts[..., :3] = torch.randint(0, 3, (POOL, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
    for k in range(settle):
        sim.engine.step(acts[k % POOL])
    tr = torch.zeros((E, 16), dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().macm_set_trace(sim.engine._h, C.c_void_p(tr.data_ptr())))
    sim.engine.step(acts[settle % POOL])
    torch.cuda.synchronize()
    t = tr.cpu().numpy()
    names = ["0 load", "1 actions", "1b tdm", "2a+2 collide", "3 integrate v", "4+5 islands, velocity solver", "6 integrate x",
             "7 position solver", "8 sleep", "9 sync fixtures", "10 find new contacts", "11-12 rewards, write-back", "13 observations"]
    tc, multi = t[:, 15] & 0xffff, (t[:, 15] >> 16) & 1
    d = np.diff(np.concatenate([np.zeros((E, 1), np.int64), t[:, :13]], axis=1), axis=1)
    print("E=%d  touching/env %.2f  multi %.3f   total cycles: mean %.0f  p99 %.0f  max %d" % (
        E, tc.mean(), multi.mean(), t[:, 12].mean(), np.percentile(t[:, 12], 99), t[:, 12].max()))
    print("%-32s %9s %9s %9s %9s" % ("phase (cycles)", "all", "tc=0", "pairs", "multi"))
    sel = [np.ones(E, bool), tc == 0, (tc > 0) & (multi == 0), multi == 1]
    for k, nm in enumerate(names):
        print("%-32s %9.0f %9.0f %9.0f %9.0f" % ((nm,) + tuple(d[s_, k].mean() if s_.any() else 0 for s_ in sel)))
    print("%-32s %9.0f %9.0f %9.0f %9.0f" % (("total",) + tuple(t[s_, 12].mean() if s_.any() else 0 for s_ in sel)))
    m = multi == 1
    if m.any():
        print("multi envs, inside 4+5: contact init + island closure %.0f, DFS %.0f, velocity lists %.0f (cycles, mean)" % (
            (t[m, 13] - t[m, 4]).mean(), (t[m, 14] - t[m, 13]).mean(), (t[m, 5] - t[m, 14]).mean()))
        for k in sorted(set(tc[m]))[:8]:
            s_ = m & (tc == k)
            print("   tc=%d: envs %d  closure %.0f  DFS %.0f  lists %.0f  position %.0f" % (k, s_.sum(), (t[s_, 13] - t[s_, 4]).mean(),
                  (t[s_, 14] - t[s_, 13]).mean(), (t[s_, 5] - t[s_, 14]).mean(), d[s_, 7].mean()))
    # the envs that set the kernel time: phase split of the 32 slowest
    worst = np.argsort(-t[:, 12])[:32]
    print("32 slowest envs: total %.0f  tc %.1f  multi %.2f" % (t[worst, 12].mean(), tc[worst].mean(), multi[worst].mean()))
    for k, nm in enumerate(names):
        print("   %-32s %9.0f" % (nm, d[worst, k].mean()))
    print("   inside 4+5: closure %.0f  DFS %.0f  lists %.0f" % ((t[worst, 13] - t[worst, 4]).mean(), (t[worst, 14] - t[worst, 13]).mean(),
                                                                  (t[worst, 5] - t[worst, 14]).mean()))
    print("   tc of the 32 slowest:", sorted(tc[worst].tolist()))
