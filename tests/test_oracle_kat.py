"""Known-answer tests that pin the CPU oracle.

The reference holds no tests and pybox2d is not importable, so these analytic vectors -- each
derivable by hand from Box2D 2.3.0's published formulas and the reference's host code
(gym_macm/envs/mvmnt.py:97-118,160-222) -- are what the restatement is checked against.
"""
import numpy as np
import pytest

f32 = np.float32
FAR = [(100.0, 100.0), (200.0, 200.0)]


def mk(oracle_mod, pos, angle, n_envs=1, **kw):
    pos = np.asarray(pos, np.float64)
    N = pos.shape[-2]
    b = oracle_mod.OracleBatch(n_envs, n_agents=N, **kw)
    T = b.T
    b.reset(pos, np.asarray(angle, np.float64), targets=np.tile([[30.0, 0.0]], (n_envs, T, 1)))
    return b


def test_kat1_free_acceleration(oracle_mod):
    # one step from rest under F = (20, 0): v = h*invMass*F*damp, x = h*v
    for model, vx, x in ((0, 0.38904545, 0.0064840913), (1, 0.391766, 0.0065294337)):
        b = mk(oracle_mod, [[0, 0], FAR[0]], [0.0, 0.0], damping_model=model)
        b.flock_step([[2, 1, 1], [1, 1, 1]])
        s = b.bodies()[0]
        assert s[0, 2] == f32(vx) and s[0, 0] == f32(x)
        assert s[0, 3] == 0 and s[0, 1] == 0
        assert np.all(s[1, :4] == f32([100, 100, 0, 0]))


def test_kat1_exact_constants(oracle_mod):
    h = f32(1.0 / 60.0)
    mass = f32(1.0) * f32(3.14159265359) * f32(0.5) * f32(0.5)
    inv_mass = f32(1.0) / mass
    assert mass == f32(0.7853982) and inv_mass == f32(1.2732395)
    v = f32(0) + h * (inv_mass * f32(20.0))
    v = v * (f32(1.0) - h * f32(5.0))
    b = mk(oracle_mod, [[0, 0], FAR[0]], [0.0, 0.0], damping_model=0)
    b.flock_step([[2, 1, 1], [1, 1, 1]])
    assert b.bodies()[0, 0, 2] == v
    # the default model (Box2D >= 2.3.1): v *= 1 / (1 + h c)
    vp = (f32(0) + h * (inv_mass * f32(20.0))) * (f32(1.0) / (f32(1.0) + h * f32(5.0)))
    b = mk(oracle_mod, [[0, 0], FAR[0]], [0.0, 0.0])
    b.flock_step([[2, 1, 1], [1, 1, 1]])
    assert b.bodies()[0, 0, 2] == vp


def test_kat2_damping_discriminator(oracle_mod):
    for model, vx in ((0, 0.9166667), (1, 0.92307687)):
        b = mk(oracle_mod, [[0, 0], FAR[0]], [0.0, 0.0], damping_model=model)
        body = b.bodies()[0]
        body[0, 2] = 1.0
        b.set_env_state(0, body, np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2)), 0.0, new_fixture=1)
        b.flock_step([[1, 1, 1], [1, 1, 1]])
        assert b.bodies()[0, 0, 2] == f32(vx)


def test_kat3_position_correction(oracle_mod):
    # A(0,0), B(0.9,0) at rest: three Baumgarte iterations push them apart; velocities stay 0
    xa, xb = f32(0.0), f32(0.9)
    m = f32(1.2732395)
    for _ in range(3):
        d = xb - xa
        n = d * (f32(1.0) / np.sqrt(d * d + f32(0) * f32(0)))  # b2Vec2::Normalize
        sep = (d * n + f32(0) * f32(0)) - f32(0.5) - f32(0.5)
        Cc = max(f32(-0.2), min(f32(0.2) * (sep + f32(0.005)), f32(0.0)))
        imp = -Cc / (m + m)
        xa = xa - m * (imp * n)
        xb = xb + m * (imp * n)
    b = mk(oracle_mod, [[0, 0], [0.9, 0]], [0.0, 0.0])
    o = b.flock_step([[1, 1, 1], [1, 1, 1]])
    s = b.bodies()[0]
    assert s[0, 0] == xa and s[1, 0] == xb
    assert abs(float(s[0, 0]) - (-0.02318)) < 1e-6 and abs(float(s[1, 0]) - 0.92318) < 1e-6
    assert np.all(s[:, 2:4] == 0) and np.all(s[:, 1] == 0)
    assert list(o["rewards"][0]) == [-1.0, -1.0]
    ab, fl, imp = b.contacts(0)
    assert ab.tolist() == [[0, 1]] and fl.tolist() == [1]


def test_kat4_proximity_penalty(oracle_mod):
    # not touching, but the fat AABBs (tight +- 0.1) overlap -> contact listed -> -1 for both
    b = mk(oracle_mod, [[0, 0], [1.15, 0]], [0.0, 0.0])
    o = b.flock_step([[1, 1, 1], [1, 1, 1]])
    assert list(o["rewards"][0]) == [-1.0, -1.0] and list(o["collided"][0]) == [1, 1]
    ab, fl, _ = b.contacts(0)
    assert ab.tolist() == [[0, 1]] and fl.tolist() == [0]
    assert np.all(b.bodies()[0, :, :2] == f32([[0, 0], [1.15, 0]]))
    b = mk(oracle_mod, [[0, 0], [1.25, 0]], [0.0, 0.0])
    o = b.flock_step([[1, 1, 1], [1, 1, 1]])
    assert list(o["collided"][0]) == [0, 0]
    assert list(o["rewards"][0]) == [0.0, 0.0]  # binary, 30 m from the target
    assert len(b.contacts(0)[0]) == 0


def test_kat5_fat_aabb_is_stateful(oracle_mod):
    b = mk(oracle_mod, [[0, 0], FAR[0]], [0.0, 0.0])
    fat0 = b.bodies()[0, 0, 6:10].copy()
    assert np.all(fat0 == f32([-0.5 - 0.1, -0.5 - 0.1, 0.5 + 0.1, 0.5 + 0.1]))
    moved_at = None
    prev_x = f32(0)
    for k in range(1, 40):
        b.flock_step([[2, 1, 1], [1, 1, 1]])
        s = b.bodies()[0, 0]
        fat = s[6:10]
        escaped = s[0] + f32(0.5) > fat0[2]
        if not escaped:
            assert np.all(fat == fat0), k
        else:
            # jump: union of tight(c0), tight(c) +- 0.1, leading edge extended by 2*(c - c0)
            c0, c = prev_x, s[0]
            lo = min(c0 - f32(0.5), c - f32(0.5)) - f32(0.1)
            hi = max(c0 + f32(0.5), c + f32(0.5)) + f32(0.1)
            hi = hi + f32(2.0) * (c - c0)
            assert fat[0] == lo and fat[2] == hi
            moved_at = k
            break
        prev_x = s[0]
    assert moved_at is not None and moved_at > 2


def test_kat6_diagonal_action(oracle_mod):
    b = mk(oracle_mod, [[0, 0], FAR[0]], [0.0, 0.0])
    b.flock_step([[2, 2, 1], [1, 1, 1]])
    s = b.bodies()[0, 0]
    # F = (cos0 + cos(pi/2), sin0 + sin(pi/2)) / sqrt(2) * 20 -> (14.142136, 14.142136) in fp32
    F = f32((np.cos(0.0) * 1 + np.cos(0.0 + np.pi / 2) * 1) * (1 / np.sqrt(2)) * 20)
    assert F == f32(14.142136)
    h = f32(1.0 / 60.0)
    v = (f32(0) + h * (f32(1.2732395) * F)) * (f32(1.0) / (f32(1.0) + h * f32(5.0)))   # default damping: Box2D >= 2.3.1
    assert s[2] == v and s[3] == v


def test_kat7_angle_wrap(oracle_mod):
    b = mk(oracle_mod, [[0, 0], FAR[0]], [3.1, 0.0])
    b.flock_step([[1, 1, 2], [1, 1, 1]])
    a = b.bodies()[0, 0, 4]
    a0 = float(f32(3.1))
    na = f32(a0 + 1 * (0.8 * 2 * np.pi) * (1 / 60.0))
    assert abs(float(na) - 3.1837757) < 1e-6
    wrapped = f32(float(na) - 2 * np.pi)
    assert a == wrapped and abs(float(a) - (-3.0994096)) < 1e-6


def test_kat8_episode_length(oracle_mod):
    b = mk(oracle_mod, [[0, 0], FAR[0]], [0.0, 0.0])
    first = None
    for k in range(1, 3700):
        o = b.flock_step([[1, 1, 1], [1, 1, 1]])
        if o["done"][0]:
            first = k
            break
    assert first == 3601


def test_kat9_nn_tie_break(oracle_mod):
    b = mk(oracle_mod, [[-1.5, 0], [0, 0], [1.5, 0]], [0.0, 0.0, 0.0])
    o = b.flock_observe()
    assert o["nn_idx"][0].tolist() == [1, 0, 1]
    assert o["nn_pos"][0, 1, 0] == 1.5
    # agent 1 looks along +x; agent 0 sits behind it: atan2(0, -1.5) - 0 = pi (not wrapped: |t| > pi is false)
    assert o["nn_pos"][0, 1, 1] == np.pi
    assert o["tg_pos"][0, 1].tolist()[:2] == [30.0, 0.0]


def test_kat10_warm_start_ratio(oracle_mod):
    # overlapping pair pushed together: step 1 runs with dtRatio 0 (fresh world), step 2 with
    # dtRatio = fl(fl(1/dt) * dt); the accumulated normal impulse must be carried
    b = mk(oracle_mod, [[0, 0], [0.95, 0]], [0.0, np.pi])
    b.flock_step([[2, 1, 1], [2, 1, 1]])
    assert b.env_info(0)["inv_dt0"] == f32(1.0) / f32(1.0 / 60.0)
    ab, fl, imp1 = b.contacts(0)
    assert fl.tolist() == [1] and imp1[0, 0] > 0
    b.flock_step([[2, 1, 1], [2, 1, 1]])
    _, _, imp2 = b.contacts(0)
    assert imp2[0, 0] > 0


def test_contact_birth_order_and_sorting(oracle_mod):
    # five mutually close agents: the first FindNewContacts creates every pair in sorted order
    pos = [[0, 0], [1.1, 0], [0, 1.1], [1.1, 1.1], [0.55, 0.55]]
    b = mk(oracle_mod, pos, [0.0] * 5)
    b.flock_step([[1, 1, 1]] * 5)
    ab, fl, _ = b.contacts(0)
    assert ab.tolist() == sorted(ab.tolist()) and len(ab) == 10


def test_continuous_bug_compat(oracle_mod):
    # mvmnt.py:124-126: (-0.9, 0.9) -> (0.7071, 0.7863): sign lost, y normalised with the NEW x
    b = mk(oracle_mod, [[0, 0], FAR[0]], [0.0, 0.0], action_mode=1)
    b.flock_step([[-0.9, 0.9], [0.0, 0.0]])
    x = np.sqrt(0.81 / 1.62)
    y = np.sqrt(0.81 / (x * x + 0.81))
    assert abs(x - 0.7071) < 1e-4 and abs(y - 0.7863) < 1e-4
    h = f32(1.0 / 60.0)
    damp = f32(1.0) / (f32(1.0) + h * f32(5.0))   # default damping: Box2D >= 2.3.1
    vx = (f32(0) + h * (f32(1.2732395) * f32(x * 20))) * damp
    vy = (f32(0) + h * (f32(1.2732395) * f32(y * 20))) * damp
    s = b.bodies()[0, 0]
    assert s[2] == vx and s[3] == vy


def test_sleep_snaps_velocity(oracle_mod):
    # a lone body coasting below the sleep tolerance falls asleep after 0.5 s: v is zeroed
    b = mk(oracle_mod, [[0, 0], FAR[0]], [0.0, 0.0])
    body = b.bodies()[0]
    body[0, 2] = 0.009
    b.set_env_state(0, body, np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2)), 0.0, new_fixture=1)
    vs = []
    for k in range(40):
        b.flock_step([[1, 1, 1], [1, 1, 1]])
        vs.append(float(b.bodies()[0, 0, 2]))
    nz = [k for k, v in enumerate(vs) if v == 0.0]
    # sleepTime reaches 0.5 on the 30th accumulated step (30 * fl(1/60) >= 0.5 in fp32)
    assert nz and nz[0] in (29, 30) and all(v > 0 for v in vs[: nz[0]])
