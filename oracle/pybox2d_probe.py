#!/usr/bin/env python
"""pybox2d probe -- TEST INFRASTRUCTURE (oracle/), run as a SUBPROCESS with a clean sys.path by tests/ and by
bench.py's reference / cpu_baseline legs.

The oracle's engine half is a restatement of Box2D 2.3.x that nothing in this image can confirm: pybox2d is not
installed and not installable (no wheel, no network).  The day a real engine IS importable -- system-wide or under
`baseline/_ref` -- this script notices and

  1. runs the discriminators of SURVEY Appendix D on the REAL engine and prints what they select:
       KAT-1  F = (20, 0) from rest, one step  ->  v.x = 0.38904545 (Box2D <= 2.3.0, damping_model "taylor")
                                                   or 0.391766 (>= 2.3.1, "pade");
       KAT-2  v0 = (1, 0), no force, one step  ->  v.x = 0.9166667 / 0.92307687 (same switch);
       KAT-3  two touching circles at rest, three position iterations -> A.x = -0.02318 (position solver);
       KAT-5  one body pushed forward: the step on which its fat AABB first changes (broadphase statefulness);
  2. writes a 100-step Flock trajectory of the real engine (positions / velocities per step, seeded, NOOP-free random
     actions) to `--trajectory out.npz`, which tests/test_pybox2d_probe.py holds the oracle to, bit for bit;
  3. with `--time S` and the reference package importable as well (`gym`, `gym_macm` of the reference under
     baseline/_ref), times the reference's own `Flock.step` loop for S seconds (kind "pybox2d").

Output: ONE JSON line.  {"available": false, "why": ...} when Box2D cannot be imported -- the callers then fall back
to the oracle port and say so ("restatement, not pybox2d").
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(HERE), "baseline", "_ref")


def _world(Box2D):
    return Box2D.b2World(gravity=(0, 0), doSleep=True)


def _body(Box2D, world, x, y, angle=0.0):
    fx = Box2D.b2FixtureDef(shape=Box2D.b2CircleShape(radius=0.5), density=1, friction=0.3)
    return world.CreateDynamicBody(fixtures=fx, linearDamping=5, fixedRotation=True, position=(x, y), angle=angle)


def discriminators(Box2D):
    import numpy as np
    out = {}
    # KAT-1: free acceleration from rest, F = (20, 0), one step: v.x = 0.38904545 (taylor) / 0.391766 (pade)
    w = _world(Box2D)
    b = _body(Box2D, w, 0.0, 0.0)
    b.ApplyForce(force=(20.0, 0.0), point=b.position, wake=True)
    w.Step(1.0 / 60.0, 8, 3)
    w.ClearForces()
    vx = float(np.float32(b.linearVelocity[0]))
    out["kat1_vx"] = vx
    out["damping_model"] = "taylor" if abs(vx - 0.38904545) < 1e-6 else ("pade" if abs(vx - 0.391766) < 1e-6 else "unknown")
    # KAT-2: v0 = (1, 0), no force, one step (needs the velocity setter)
    try:
        w = _world(Box2D)
        b = _body(Box2D, w, 0.0, 0.0)
        b.linearVelocity = (1.0, 0.0)
        w.Step(1.0 / 60.0, 8, 3)
        out["kat2_vx"] = float(np.float32(b.linearVelocity[0]))
    except Exception:
        out["kat2_vx"] = None
    # KAT-3: position correction of two overlapping circles at rest
    w = _world(Box2D)
    a, b = _body(Box2D, w, 0.0, 0.0), _body(Box2D, w, 0.9, 0.0)
    w.Step(1.0 / 60.0, 8, 3)
    out["kat3_ax"] = float(np.float32(a.position[0]))
    out["kat3_ok"] = abs(out["kat3_ax"] - (-0.02318)) < 2e-6
    # KAT-5: fat AABB statefulness
    w = _world(Box2D)
    b = _body(Box2D, w, 0.0, 0.0)
    first = None
    lo0 = None
    for k in range(60):
        b.ApplyForce(force=(20.0, 0.0), point=b.position, wake=True)
        w.Step(1.0 / 60.0, 8, 3)
        w.ClearForces()
        try:
            aabb = b.fixtures[0].GetAABB(0)
            lo = float(aabb.lowerBound[0])
        except Exception:
            break
        if lo0 is None:
            lo0 = lo
        elif lo != lo0 and first is None:
            first = k + 1
    out["kat5_first_move_step"] = first
    out["box2d_version"] = getattr(Box2D, "__version__", None)
    return out


def trajectory(Box2D, path, n_agents=16, steps=100, seed=7, spread=6.0):
    """A seeded crowded Flock world stepped by the real engine with the reference's action decode
    (mvmnt.py:97-118), saved for the oracle to replay."""
    import numpy as np
    rng = np.random.default_rng(seed)
    pos = spread * (rng.random((n_agents, 2)) - 0.5)
    ang = rng.uniform(-1, 1, n_agents) * np.pi
    acts = rng.integers(0, 3, (steps, n_agents, 3))
    w = _world(Box2D)
    bodies = [_body(Box2D, w, float(pos[i, 0]), float(pos[i, 1]), float(ang[i])) for i in range(n_agents)]
    rec = np.zeros((steps, n_agents, 5), np.float32)
    for k in range(steps):
        for i, body in enumerate(bodies):
            a = acts[k, i]
            body.angle = body.angle + (a[2] - 1) * (0.8 * 2 * np.pi) * (1 / 60.0)
            if np.abs(body.angle) > np.pi:
                body.angle -= np.sign(body.angle) * (2 * np.pi)
            c = 1 / np.sqrt(2) if (a[0] != 1 and a[1] != 1) else 1
            f = (np.array([np.cos(body.angle), np.sin(body.angle)]) * (a[0] - 1) +
                 np.array([np.cos(body.angle + np.pi / 2), np.sin(body.angle + np.pi / 2)]) * (a[1] - 1)) * c * 20
            body.ApplyForce(force=(float(f[0]), float(f[1])), point=body.position, wake=True)
        w.Step(1.0 / 60.0, 8, 3)
        w.ClearForces()
        for i, body in enumerate(bodies):
            rec[k, i] = (body.position[0], body.position[1], body.linearVelocity[0], body.linearVelocity[1], body.angle)
    np.savez(path, pos=pos, ang=ang, acts=acts, rec=rec)
    return {"trajectory": path, "steps": steps, "n_agents": n_agents}


def time_reference(seconds, n_agents=64):
    """The reference's own loop (README.md:8-14): Flock(n_agents=[N], reward_mode='linear') stepped with random
    actions on the NoRender backend, single process."""
    import numpy as np
    import gym_macm  # noqa: F401  (the REFERENCE's package: baseline/_ref comes first on sys.path here)
    from gym_macm.envs.mvmnt import Flock
    env = Flock(n_agents=[n_agents], reward_mode="linear")
    rng = np.random.default_rng(0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        env.step({a.id: rng.integers(0, 3, 3) for a in env.agents})
        n += 1
    el = time.perf_counter() - t0
    return {"agent_steps_per_sec": n * n_agents / el, "steps": n, "seconds": el, "n_agents": n_agents}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trajectory", default=None)
    ap.add_argument("--time", type=float, default=0.0)
    args = ap.parse_args()
    # a clean path: the reference install (if any) first, never this repo's own gym_macm package
    sys.path[:] = [p for p in sys.path if "gym-macm_b200" not in p and os.path.abspath(p or ".") != HERE]
    if os.path.isdir(REF):
        sys.path.insert(0, REF)
    try:
        import Box2D
    except Exception as ex:
        print(json.dumps({"available": False, "why": "import Box2D: %s: %s" % (type(ex).__name__, ex),
                          "baseline_ref": os.path.isdir(REF)}))
        return
    out = {"available": True, "baseline_ref": os.path.isdir(REF)}
    try:
        out.update(discriminators(Box2D))
        if args.trajectory:
            out.update(trajectory(Box2D, args.trajectory))
        if args.time > 0:
            try:
                out["reference_loop"] = time_reference(args.time)
            except Exception as ex:
                out["reference_loop"] = {"unavailable": "%s: %s" % (type(ex).__name__, ex)}
    except Exception as ex:
        out["error"] = "%s: %s" % (type(ex).__name__, ex)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
