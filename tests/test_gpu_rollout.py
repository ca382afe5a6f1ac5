"""macm_rollout (K steps in one launch, env state held on chip) against K calls of macm_step: every state buffer
and every per-step output must be bit-identical -- and the single step is what test_gpu_parity / test_golden hold
to the oracle and to the reference's own host code."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STATE = ("posvel", "angsleep", "fat", "contact_count", "env_state", "obs", "nn_idx", "rewards", "collided", "done",
         "tdm_state")


def _same_state(a, b, what):
    import torch
    for n in STATE:
        if n in a.state:
            assert torch.equal(a.state[n], b.state[n]), (what, n)
    # contact lists: only the live prefix of each env is defined
    cnt = a.state["contact_count"].cpu().numpy()
    ab_a, ab_b = a.state["contact_ab"].cpu().numpy(), b.state["contact_ab"].cpu().numpy()
    im_a, im_b = a.state["contact_imp"].cpu().numpy(), b.state["contact_imp"].cpu().numpy()
    live = np.arange(ab_a.shape[1])[None, :] < cnt[:, None]
    assert np.array_equal(ab_a[live], ab_b[live]), (what, "contact_ab")
    assert np.array_equal(im_a[live], im_b[live]), (what, "contact_imp")


def _flock_pair(E, n_agents, seed, **kw):
    import gym_macm
    return [gym_macm.BatchedFlock(E, n_agents=n_agents, device="cuda:0", seed=seed, **kw) for _ in range(2)]


@pytest.mark.parametrize("N,E,kw", [
    (64, 96, dict(reward_mode="linear")),                       # BASELINE configs[1] shape, two agents per lane
    (6, 200, dict(targets=[0, 0, 1, 1, 2, 2])),                 # configs[2]: four envs per warp
    (16, 64, dict(coord="cartesian", start_spread=6.0)),        # crowded, cartesian observations
    (33, 40, dict(start_spread=8.0)),                           # ragged: 33 agents on 64 slots, dense contacts
    (100, 30, dict(start_spread=12.0)),                         # more than 64 agents: four per lane, 128-bit sets
])
def test_rollout_equals_steps_flock(N, E, kw):
    import torch
    K = 48
    one, many = _flock_pair(E, [N], 11, **kw)
    g = torch.Generator(device="cuda:0")
    g.manual_seed(N)
    acts = torch.zeros((K, E, N, 4), dtype=torch.uint8, device="cuda:0")
    acts[..., :3] = torch.randint(0, 3, (K, E, N, 3), generator=g, device="cuda:0", dtype=torch.uint8)
    per = {n: [] for n in ("obs", "nn_idx", "rewards", "collided", "done")}
    for k in range(K):
        one.step(acts[k])
        for n in per:
            per[n].append(one.state[n].clone())
    out = many.rollout(acts)
    torch.cuda.synchronize()
    for n in per:
        assert torch.equal(out[n], torch.stack(per[n])), n
    _same_state(one, many, "flock N=%d" % N)
    assert int(many.engine.launch_count) < int(one.engine.launch_count)
    # a second rollout continues from the first (contact lists written by the launch, not just the bodies)
    for k in range(8):
        one.step(acts[k])
    many.rollout(acts[:8])
    _same_state(one, many, "flock N=%d, second launch" % N)


def test_rollout_in_chunks_and_action_repeat():
    import torch
    E, N, K = 64, 64, 30
    one, many = _flock_pair(E, [N], 5, reward_mode="linear", start_spread=12.0)
    acts = torch.zeros((K, E, N, 4), dtype=torch.uint8, device="cuda:0")
    acts[..., :3] = torch.randint(0, 3, (K, E, N, 3), device="cuda:0", dtype=torch.uint8)
    rew = []
    for k in range(K):
        one.step(acts[k])
        rew.append(one.state["rewards"].clone())
    # chunks of 7, 7, 7, 7, 2; no per-step observations (the observation pass runs after each chunk's last step)
    got = []
    for lo in range(0, K, 7):
        out = many.rollout(acts[lo:lo + 7], want=("rewards",))
        assert set(out) == {"rewards"}
        got.append(out["rewards"])
    assert torch.equal(torch.cat(got), torch.stack(rew))
    _same_state(one, many, "chunks")


def test_rollout_continuous_actions():
    import torch
    E, N, K = 32, 10, 25
    one, many = _flock_pair(E, [N], 2, action_mode="continuous", start_spread=5.0)
    acts = (torch.rand((K, E, N, 2), device="cuda:0") * 2.4 - 1.2).contiguous()
    for k in range(K):
        one.step(acts[k])
    out = many.rollout(acts)
    assert torch.equal(out["obs"][-1], one.state["obs"])
    _same_state(one, many, "continuous")


@pytest.mark.parametrize("policy", ["random", "flock", "forward", "diag"])
def test_rollout_actions_none_mode(policy):
    """actions=None (mvmnt.py:86-92): the actors inside the launch == macm_bot_actions + macm_step per step."""
    import torch
    E, N, K = 48, 64, 40
    one, many = _flock_pair(E, [N], 9, start_spread=10.0)
    obs = []
    for k in range(K):
        one.step(one.bot_actions(policy, seed=77))
        obs.append(one.state["obs"].clone())
    out = many.rollout(None, n_steps=K, policy=policy, seed=77)
    assert torch.equal(out["obs"], torch.stack(obs))
    _same_state(one, many, policy)


def test_rollout_tdm():
    import torch
    import gym_macm
    E, K = 64, 320   # an attack every 60 steps at most and four hits to die: long enough for deaths
    one, many = [gym_macm.BatchedTDM(E, n_agents=[15, 15, 15], device="cuda:0", seed=4) for _ in range(2)]
    N = one.engine.N
    g = torch.Generator(device="cuda:0")
    g.manual_seed(1)
    acts = torch.randint(0, 3, (K, E, N, 4), generator=g, device="cuda:0", dtype=torch.uint8)
    acts[..., 3] = torch.randint(0, 2, (K, E, N), generator=g, device="cuda:0", dtype=torch.uint8)
    for lo in range(0, K, 80):
        rew, done = [], []
        for k in range(lo, lo + 80):
            one.step(acts[k])
            rew.append(one.state["rewards"].clone())
            done.append(one.state["done"].clone())
        out = many.rollout(acts[lo:lo + 80], want=("rewards", "done", "obs"))
        assert torch.equal(out["rewards"], torch.stack(rew)) and torch.equal(out["done"], torch.stack(done))
        assert torch.equal(out["obs"][-1], one.state["obs"])
        _same_state(one, many, "tdm %d" % lo)
    assert int(one.state["tdm_state"][..., 3].view(torch.int32).bitwise_and(1).sum()) < E * N   # somebody died


def test_rollout_status_codes():
    import ctypes as C
    import torch
    import gym_macm
    from gym_macm import _lib
    env = gym_macm.BatchedFlock(4, n_agents=[4], device="cuda:0", seed=0, coord="cartesian")
    L = _lib.lib()
    a = torch.zeros((2, 4, 4, 4), dtype=torch.uint8, device="cuda:0")
    assert L.macm_rollout(env.engine._h, C.c_void_p(a.data_ptr()), 0, -1, 0, None, None) == -1        # n_steps < 1
    assert L.macm_rollout(env.engine._h, None, 2, 99, 0, None, None) == -1                           # unknown policy
    assert L.macm_rollout(env.engine._h, None, 2, _lib.BOTS["flock"], 0, None, None) == -6           # flock actor: polar only
    assert L.macm_rollout(env.engine._h, None, 2, _lib.BOTS["combat"], 0, None, None) == -6
    assert L.macm_rollout(env.engine._h, C.c_void_p(a.data_ptr() + 1), 2, -1, 0, None, None) == -4   # misaligned
    assert L.macm_rollout(env.engine._h, C.c_void_p(a.data_ptr()), 2, -1, 0, None, None) == 0        # out == NULL is fine
    torch.cuda.synchronize()
    assert int(env.step_count[0]) == 2
