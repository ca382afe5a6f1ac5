"""Golden vectors produced by the reference's own host code (tests/golden/gen_reference_golden.py:
/root/reference/gym_macm imported unmodified over the Box2D/gym stand-ins) against
  (a) the oracle's C restatement of that host logic            -- CPU, every round
  (b) the CUDA path through the C ABI                          -- GPU
The fixtures travel with the repo; nothing here reads /root/reference."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden", "flock_ref_*.npz")))
MODES = dict(binary=0, linear=1, discrete=0, continuous=1, polar=0, cartesian=1)


def ang_diff(a, b):
    d = np.abs(a - b) % (2 * np.pi)
    return np.minimum(d, 2 * np.pi - d)


def load(path):
    g = np.load(path)
    reward_mode, action_mode, coord, time_limit, spread = [str(x) for x in g["settings"]]
    return g, reward_mode, action_mode, coord, float(time_limit)


def test_fixtures_present():
    assert len(CASES) >= 6


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[10:-4] for p in CASES])
def test_oracle_host_logic_matches_reference_python(path, oracle_mod):
    g, reward_mode, action_mode, coord, time_limit = load(path)
    N, T = int(g["N"]), len(g["targets"])
    ref = oracle_mod.OracleBatch(1, n_agents=N, n_targets=T, reward_mode=MODES[reward_mode],
                                 action_mode=MODES[action_mode], coord=MODES[coord], time_limit=time_limit)
    ref.reset(g["pos0"][None], g["angle0"][None], targets=g["targets"][None], target_idx=g["target_idx"])
    D = g["nn"].shape[-1]
    o = ref.flock_observe()
    assert np.array_equal(o["nn_idx"][0], g["obs0_nn_id"])
    assert np.allclose(o["nn_pos"][0, :, :D], g["obs0_nn"], rtol=1e-13, atol=1e-15)
    assert np.allclose(o["tg_pos"][0, :, :D], g["obs0_tg"], rtol=1e-13, atol=1e-15)
    for k in range(len(g["actions"])):
        o = ref.flock_step(g["actions"][k][None])
        assert np.array_equal(ref.bodies()[0], g["bodies"][k]), "step %d: body state" % k
        assert np.array_equal(o["nn_idx"][0], g["nn_id"][k]), "step %d: nearest ids" % k
        assert np.array_equal(o["rewards"][0] == -1, g["rewards"][k] == -1), "step %d: collision penalty" % k
        assert np.allclose(o["rewards"][0], g["rewards"][k], rtol=1e-14, atol=0), "step %d: rewards" % k
        assert np.allclose(o["nn_pos"][0, :, :D], g["nn"][k], rtol=1e-13, atol=1e-15), "step %d: nn obs" % k
        assert np.allclose(o["tg_pos"][0, :, :D], g["tg"][k], rtol=1e-13, atol=1e-15), "step %d: target obs" % k
        assert bool(o["done"][0]) == bool(g["done"][k]), "step %d: done" % k
        assert ref.env_info(0)["contacts"] == g["n_contacts"][k]


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[10:-4] for p in CASES])
def test_cuda_path_matches_reference_python(path):
    import torch
    import gym_macm
    g, reward_mode, action_mode, coord, time_limit = load(path)
    N = int(g["N"])
    targets = None if len(g["targets"]) == 1 else [int(t) for t in g["target_idx"]]
    env = gym_macm.BatchedFlock(1, n_agents=[N], targets=targets, device="cuda:0", seed=None, reward_mode=reward_mode,
                                action_mode=action_mode, coord=coord, time_limit=time_limit)
    env.load_state(g["pos0"][None], g["angle0"][None], targets=g["targets"][None])
    D = g["nn"].shape[-1]

    def check_obs(nn_id, nn, tg, k):
        st = env.state
        assert np.array_equal(st["nn_idx"][0].cpu().numpy(), nn_id), "step %d: nearest ids" % k
        ob = st["obs"][0].cpu().numpy().astype(np.float64)
        for got, want in ((ob[:, 0:D], nn), (ob[:, D:2 * D], tg)):
            assert np.allclose(got[:, 0], want[:, 0], rtol=1e-5, atol=2e-6), "step %d: distances" % k
            if D == 2:
                assert ang_diff(got[:, 1], want[:, 1]).max() <= 3e-6, "step %d: angles" % k
            else:
                assert np.allclose(got[:, 1:], want[:, 1:], rtol=0, atol=3e-6), "step %d: cos/sin" % k

    torch.cuda.synchronize()
    check_obs(g["obs0_nn_id"], g["obs0_nn"], g["obs0_tg"], -1)
    for k in range(len(g["actions"])):
        a = g["actions"][k][None]
        if action_mode == "discrete":
            env.step(torch.as_tensor(a, device="cuda:0"))
        else:
            env.step(torch.as_tensor(a, dtype=torch.float32, device="cuda:0"))
        torch.cuda.synchronize()
        st = env.state
        body = np.concatenate([st["posvel"][0].cpu().numpy(), st["angsleep"][0].cpu().numpy(), st["fat"][0].cpu().numpy()], -1)
        assert np.array_equal(body, g["bodies"][k]), "step %d: body state" % k
        rew = st["rewards"][0].cpu().numpy().astype(np.float64)
        assert np.array_equal(rew == -1, g["rewards"][k] == -1), "step %d: collision penalty" % k
        if reward_mode == "binary":
            assert np.array_equal(rew, g["rewards"][k]), "step %d: binary rewards" % k
        else:
            assert np.allclose(rew, g["rewards"][k], rtol=0, atol=1e-6), "step %d: linear rewards" % k
        assert bool(st["done"][0]) == bool(g["done"][k]), "step %d: done" % k
        assert int(st["contact_count"][0]) == g["n_contacts"][k]
        check_obs(g["nn_id"][k], g["nn"][k], g["tg"][k], k)
    env.close()


@pytest.mark.gpu
def test_dict_api_replays_golden_run():
    """The drop-in dict API (gym_macm.make -> Flock.step) on the same draws as the reference run:
    random.seed(s) makes Flock.__init__ sample the very state the reference sampled."""
    import random
    import gym_macm
    path = [p for p in CASES if "cfg1_n4_binary" in p][0]
    g, *_ = load(path)
    random.seed(0)
    env = gym_macm.make("gym_macm:cm-flock-v0", n_agents=[4])
    assert env.done is False and sorted(env.obs) == [0, 1, 2, 3]
    assert np.allclose(np.array([env.obs[i]["nodes"][0]["position"] for i in range(4)]), g["obs0_nn"], rtol=1e-5, atol=2e-6)
    for k in range(60):
        obs, rewards = env.step({i: g["actions"][k][i] for i in range(4)})
        assert [obs[i]["nodes"][0]["id"] for i in range(4)] == g["nn_id"][k].tolist()
        assert [rewards[i] for i in range(4)] == g["rewards"][k].tolist()
        assert all(isinstance(rewards[i], int) for i in range(4))
        assert obs[0]["nodes"][1]["id"] == 4 and obs[0]["nodes"][1]["type"] == 1
        assert np.allclose(obs[2]["nodes"][1]["position"], g["tg"][k][2], rtol=1e-5, atol=3e-6)
    with pytest.raises(AssertionError):
        env.step({i: [3, 0, 0] for i in range(4)})
    env.close()
