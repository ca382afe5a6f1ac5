"""A stand-in for the pybox2d module, just large enough for the reference's own host code
(gym_macm/envs/mvmnt.py, combat.py, cm_framework.py, settings.py) to import and run.

TEST INFRASTRUCTURE.  The physics behind `b2World` is the CPU oracle (oracle/macm_oracle.c), so
this does NOT pin Box2D's engine semantics (pybox2d is not installable here); what it pins is
the reference's HOST logic -- action decode, force computation, reward and observation passes,
episode clock -- because that code runs unmodified from /root/reference on top of it.
Only tests/golden/gen_reference_golden.py imports this package.
"""
import numpy as np

from oracle import oracle as _oracle

b2_epsilon = float(np.finfo(np.float32).eps)
b2_dynamicBody = 2
b2_addState, b2_persistState = 1, 2
f32 = np.float32


class b2Vec2(object):
    """float32 2-vector; arithmetic rounds to float32 like the C++ struct behind the SWIG wrapper."""
    __slots__ = ("x", "y")

    def __init__(self, *a):
        if len(a) == 1:
            a = a[0]
        if len(a) == 0:
            a = (0.0, 0.0)
        self.x, self.y = float(f32(a[0])), float(f32(a[1]))

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __len__(self):
        return 2

    def __iter__(self):
        return iter((self.x, self.y))

    def __add__(self, o):
        o = o if isinstance(o, b2Vec2) else b2Vec2(o)
        return b2Vec2(f32(self.x) + f32(o.x), f32(self.y) + f32(o.y))

    __radd__ = __add__

    def __sub__(self, o):
        o = o if isinstance(o, b2Vec2) else b2Vec2(o)
        return b2Vec2(f32(self.x) - f32(o.x), f32(self.y) - f32(o.y))

    def __repr__(self):
        return "b2Vec2(%r, %r)" % (self.x, self.y)


def b2DistanceSquared(a, b):
    a = a if isinstance(a, b2Vec2) else b2Vec2(a)
    b = b if isinstance(b, b2Vec2) else b2Vec2(b)
    cx, cy = f32(a.x) - f32(b.x), f32(a.y) - f32(b.y)
    return float(f32(cx * cx) + f32(cy * cy))


class b2Color(object):
    def __init__(self, r=0, g=0, b=0):
        self.r, self.g, self.b = r, g, b


class b2CircleShape(object):
    def __init__(self, radius=0.0, pos=(0, 0)):
        self.radius, self.pos = radius, pos


class b2EdgeShape(object):
    pass


class b2PolygonShape(object):
    pass


class b2FixtureDef(object):
    def __init__(self, shape=None, density=0.0, friction=0.2, restitution=0.0):
        self.shape, self.density, self.friction, self.restitution = shape, density, friction, restitution


class b2AABB(object):
    pass


class b2Fixture(object):
    def __init__(self, body):
        self.body = body


class b2Joint(object):
    pass


class b2QueryCallback(object):
    pass


class b2DrawExtended(object):
    pass


class b2ContactListener(object):
    def __init__(self, **kw):
        pass


class b2DestructionListener(object):
    def __init__(self, **kw):
        pass


class b2RayCastCallback(object):
    def __init__(self, **kw):
        pass


def b2GetPointStates(*a):
    return (), ()


def b2Random(lo=-1.0, hi=1.0):
    return float(np.random.uniform(lo, hi))


class _Contact(object):
    def __init__(self, world, a, b, touching):
        self.fixtureA, self.fixtureB = world._bodies[a].fixtures[0], world._bodies[b].fixtures[0]
        self.touching = bool(touching)


class b2Body(object):
    def __init__(self, world, index, userData):
        self._w, self._i, self.userData = world, index, userData
        self.fixtures = [b2Fixture(self)]
        self._active = True

    def _row(self):
        return self._w._ow.bodies()[self._i]

    @property
    def position(self):
        r = self._row()
        return b2Vec2(r[0], r[1])

    @property
    def linearVelocity(self):
        r = self._row()
        return b2Vec2(r[2], r[3])

    @property
    def angle(self):   # GetAngle(): the float32 sweep angle as a Python float
        return self._w._ow.get_angle(self._i)

    @angle.setter
    def angle(self, a):   # SetTransform(position, angle): the angle is rounded to float32
        self._w._ow.set_angle(self._i, float(a))

    @property
    def active(self):
        return self._active

    @active.setter
    def active(self, flag):
        self._active = bool(flag)
        self._w._ow.set_active(self._i, flag)

    def ApplyForce(self, force, point, wake):
        assert wake
        f = force if isinstance(force, b2Vec2) else b2Vec2(force)
        self._w._ow.apply_force(self._i, f.x, f.y)


class b2World(object):
    def __init__(self, gravity=(0, 0), doSleep=True):
        assert tuple(gravity) == (0, 0) and doSleep
        self._ow, self._bodies = None, []
        self.warmStarting, self.continuousPhysics, self.subStepping = True, True, False
        self.destructionListener = None
        self.contactListener = None
        self.damping_model = int(__import__("os").environ.get("MACM_SHIM_DAMPING", "1"))   # 0 taylor, 1 pade (default, as gym_macm.settings)

    def CreateDynamicBody(self, fixtures=None, linearDamping=0.0, fixedRotation=False, position=(0, 0), angle=0.0,
                          userData=None, **kw):
        assert fixedRotation and not kw
        if self._ow is None:
            self._ow = _oracle.OracleWorld(radius=fixtures.shape.radius, density=fixtures.density,
                                           friction=fixtures.friction, linear_damping=linearDamping,
                                           damping_model=self.damping_model)
        i = self._ow.add_body(position[0], position[1], angle)
        b = b2Body(self, i, userData)
        self._bodies.append(b)
        return b

    def Step(self, timeStep, velocityIterations, positionIterations):
        self._ow.set_warm_starting(self.warmStarting)
        self._ow.step(timeStep, velocityIterations, positionIterations)

    def ClearForces(self):
        pass   # the oracle's step clears forces like b2World::Step does (auto clear) and ClearForces after it

    @property
    def contacts(self):
        ab, fl, _ = self._ow.contacts()
        # world list order: newest first
        return [_Contact(self, a, b, t) for (a, b), t in zip(ab[::-1], fl[::-1])]

    def RayCast(self, callback, point1, point2):
        p1 = point1 if isinstance(point1, b2Vec2) else b2Vec2(point1)
        p2 = point2 if isinstance(point2, b2Vec2) else b2Vec2(point2)
        hit, fr = self._ow.raycast((p1.x, p1.y), (p2.x, p2.y))
        if hit >= 0:
            pt = ((1 - fr) * p1.x + fr * p2.x, (1 - fr) * p1.y + fr * p2.y)
            callback.ReportFixture(self._bodies[hit].fixtures[0], pt, (0.0, 0.0), fr)
