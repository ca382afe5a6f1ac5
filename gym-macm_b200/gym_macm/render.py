"""State export for the reference's CPU renderer.

Rendering stays in the reference (gym_macm/backends/pyglet_framework.py, off the hot path); what it needs from the
simulator is, per frame, the few attributes `PygletDraw.ManualDraw` reads (pyglet_framework.py:122-180,361-383):

    for body in test.world.bodies:  body.transform.position / .angle, body.userData.color (.r .g .b, `color / 3`
                                    for an inactive body), body.active, body.fixtures[k].shape.radius
    for o in test.gui_objects.values():  o['shape'] == 'circle', o['values'] = [centre, radius, colour]
    test.settings.drawShapes / drawAABBs

`export(env, e)` builds exactly those objects for ONE env of a batch from a ~1 KB device->host read; `build(...)`
does the same from numpy arrays (what the CPU tests drive).  Colours follow the reference: an agent listed in a
world contact is b2Color(1, 0.2, 0.2) (mvmnt.py:165-167), otherwise its base colour (mvmnt.py:21-22, or `colors[i]`,
mvmnt.py:68-69); TDM agents carry their team colour (combat.py:37-44) and dead ones are inactive bodies, which the
renderer dims by 3.  Targets are the white circles of radius `reward_radius` that mvmnt.py:54-57 puts into
`gui_objects`.  (In binary reward mode the reference never turns a collided agent back from red -- reset_color is
only called on the linear branch, mvmnt.py:172-175; this export always shows the current step's contacts.)
"""
import numpy as np


class Color(object):
    """b2Color: r, g, b attributes, division by a scalar, iteration."""

    def __init__(self, r, g, b):
        self.r, self.g, self.b = float(r), float(g), float(b)

    def __truediv__(self, k):
        return Color(self.r / k, self.g / k, self.b / k)

    __div__ = __truediv__

    def __iter__(self):
        return iter((self.r, self.g, self.b))

    def __eq__(self, other):
        return tuple(self) == tuple(other)

    def __repr__(self):
        return "Color(%g, %g, %g)" % (self.r, self.g, self.b)


class Vec2(tuple):
    """b2Vec2 as the renderer uses it: indexable, .x / .y."""

    def __new__(cls, x, y):
        return tuple.__new__(cls, (float(x), float(y)))

    x = property(lambda self: self[0])
    y = property(lambda self: self[1])


class Transform(object):
    def __init__(self, position, angle):
        self.position, self.angle = position, float(angle)


class CircleShape(object):
    childCount = 1

    def __init__(self, radius):
        self.radius = float(radius)


class Fixture(object):
    def __init__(self, shape):
        self.shape = shape


class AgentData(object):
    """body.userData: the agent record (id, colour)."""

    def __init__(self, ID, color):
        self.id, self.color = ID, color


class Body(object):
    def __init__(self, transform, userData, radius, active=True, awake=True, velocity=(0.0, 0.0)):
        self.transform, self.userData = transform, userData
        self.position, self.angle = transform.position, transform.angle
        self.linearVelocity = Vec2(*velocity)
        self.active, self.awake = bool(active), bool(awake)
        self.fixtures = [Fixture(CircleShape(radius))]


class World(object):
    def __init__(self, bodies):
        self.bodies = bodies


class Frame(object):
    """What `PygletDraw(test)` reads from `test`: .world.bodies, .gui_objects, .settings."""

    def __init__(self, world, gui_objects, settings):
        self.world, self.gui_objects, self.settings = world, gui_objects, settings


FLOCK_BASE, COLLIDED, WHITE = (0.4, 0.4, 0.6), (1.0, 0.2, 0.2), (1.0, 1.0, 1.0)
TEAM_COLORS = ((0.2, 0.2, 1.0), (1.0, 0.2, 0.2), (0.2, 1.0, 0.2))   # combat.py:37-44


def build(posvel, angle, collided, settings, targets=None, colors=None, teams=None, alive=None, sleeping=None):
    """numpy state of ONE env -> Frame.  posvel [N,4], angle [N], collided [N]; Flock: targets [T,2];
    TDM: teams [N], alive [N]."""
    posvel = np.asarray(posvel, np.float64)
    N = posvel.shape[0]
    radius = settings.bodySettings["fixtures"].radius
    bodies = []
    for i in range(N):
        if teams is not None:
            t = int(teams[i])
            col = Color(*TEAM_COLORS[t]) if t < len(TEAM_COLORS) else None   # combat.py:37-44 has three teams
            ID = None
        else:
            base = colors[i] if colors else FLOCK_BASE
            col = Color(*COLLIDED) if collided[i] else Color(*base)
            ID = i
        act = True if alive is None else bool(alive[i])
        awake = True if sleeping is None else not bool(sleeping[i])
        bodies.append(Body(Transform(Vec2(posvel[i, 0], posvel[i, 1]), angle[i]), AgentData(i if ID is None else ID, col),
                           radius, active=act, awake=awake, velocity=(posvel[i, 2], posvel[i, 3])))
    gui = {}
    if targets is not None:
        for t, p in enumerate(np.asarray(targets, np.float64).reshape(-1, 2)):
            gui["target" + str(t)] = {"shape": "circle",
                                      "values": [Vec2(p[0], p[1]), settings.reward_radius, Color(*WHITE)]}
    return Frame(World(bodies), gui, settings)


def export(env, e=0):
    """One env of a BatchedFlock / BatchedTDM -> Frame (device->host read of that env's rows)."""
    t = env.engine.t
    posvel = t["posvel"][e].cpu().numpy()
    angle = t["angsleep"][e, :, 0].cpu().numpy()
    collided = t["collided"][e].cpu().numpy()
    if "tdm_state" in t:
        alive = (t["tdm_state"][e, :, 3].cpu().numpy().view(np.int32) & 1).astype(bool)
        return build(posvel, angle, collided, env.settings, teams=env.teams, alive=alive)
    return build(posvel, angle, collided, env.settings, targets=t["targets"][e].cpu().numpy(), colors=env.colors)
