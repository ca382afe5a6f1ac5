"""the state export for the reference's CPU renderer.  The check walks a Frame exactly the way
`PygletDraw.ManualDraw` / `DrawSolidCircle` do (backends/pyglet_framework.py:122-180,361-383) and records what
would be drawn; no GPU, no pyglet."""
import numpy as np


def _manual_draw(test):
    """The reads of ManualDraw, with the draw calls collected instead of rendered."""
    drawn = []
    if test.settings.drawShapes:
        for body in test.world.bodies:
            transform = body.transform
            color = body.userData.color if body.userData else None
            if not body.active:
                color = color / 3
            for fixture in body.fixtures:
                center = transform.position
                axis = (np.cos(transform.angle), np.sin(transform.angle))
                drawn.append(("solid_circle", (center[0], center[1]), fixture.shape.radius, axis, (color.r, color.g, color.b)))
        for obj in test.gui_objects.values():
            if obj["shape"] == "circle":
                c, r, col = obj["values"]
                drawn.append(("circle", (c[0], c[1]), r, None, (col.r, col.g, col.b)))
    return drawn


def test_flock_frame_is_what_the_renderer_reads():
    from gym_macm import render
    from gym_macm.settings import flockSettings
    s = flockSettings(reward_mode="binary")
    posvel = np.array([[1.0, 2.0, 0.1, 0.0], [-3.0, 0.5, 0.0, 0.0], [4.0, 4.0, 0.0, -1.0]], np.float32)
    angle = np.array([0.0, np.pi / 2, -1.0], np.float32)
    collided = np.array([0, 1, 0], np.uint8)
    targets = np.array([[30.0, 0.0], [0.0, -40.0]], np.float32)
    fr = render.build(posvel, angle, collided, s, targets=targets)
    d = _manual_draw(fr)
    assert len(d) == 5
    assert d[0][:3] == ("solid_circle", (1.0, 2.0), 0.5) and d[0][4] == (0.4, 0.4, 0.6)       # mvmnt.py:21-22
    assert d[1][4] == (1.0, 0.2, 0.2)                                                         # mvmnt.py:165-167
    assert np.allclose(d[1][3], (0.0, 1.0), atol=1e-6)                                        # heading axis
    assert d[3] == ("circle", (30.0, 0.0), 7, None, (1.0, 1.0, 1.0))                          # mvmnt.py:54-57
    assert set(fr.gui_objects) == {"target0", "target1"}
    assert [b.userData.id for b in fr.world.bodies] == [0, 1, 2]
    assert fr.world.bodies[2].linearVelocity == (0.0, -1.0)
    # ctor colours (mvmnt.py:68-69)
    fr = render.build(posvel, angle, np.zeros(3, np.uint8), s, targets=targets, colors=[(0, 1, 0)] * 3)
    assert _manual_draw(fr)[0][4] == (0.0, 1.0, 0.0)


def test_tdm_frame_dims_dead_agents():
    from gym_macm import render
    from gym_macm.settings import combatSettings
    s = combatSettings()
    posvel = np.zeros((4, 4), np.float32)
    fr = render.build(posvel, np.zeros(4), np.zeros(4, np.uint8), s, teams=[0, 0, 1, 2], alive=[1, 0, 1, 1])
    d = _manual_draw(fr)
    assert d[0][4] == (0.2, 0.2, 1.0) and d[2][4] == (1.0, 0.2, 0.2) and d[3][4] == (0.2, 1.0, 0.2)   # combat.py:37-44
    assert np.allclose(d[1][4], (0.2 / 3, 0.2 / 3, 1.0 / 3))        # `if not body.active: color = color/3`
    assert fr.gui_objects == {}
