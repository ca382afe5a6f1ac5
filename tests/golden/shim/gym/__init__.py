"""Minimal `gym` for running the reference's host code offline (gym is not installed).
TEST INFRASTRUCTURE; only tests/golden/gen_reference_golden.py imports it."""
from . import spaces  # noqa: F401


class Env(object):
    pass
