"""Minimal stand-in for the `gym` package (Dec-2020 API surface the reference touches).

Used ONLY by oracle/make_golden.py to import the unmodified reference `gym_ffmp`
package in a container that has no gym.  Not part of the product.
"""
from . import error, spaces, utils  # noqa: F401
from .envs.registration import make, register  # noqa: F401


class Env(object):
    metadata = {}

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError
