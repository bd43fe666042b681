"""Minimal stand-ins for the pieces of `gym` the reference environment uses (gym is not in this image).

If `gym` is importable it is used; otherwise `Env`, `spaces.Discrete`, `spaces.Box` and
`utils.seeding.np_random` below reproduce the gym 0.17-0.21 behaviour the reference relies on
(ssa_tasker_simple_2.py:14-15, 164-177, 188-191): SHA-512 hashed integer seeds feeding a numpy
`RandomState`, and `Discrete.sample()` drawing from the space's own generator.
"""
import hashlib
import os
import struct

import numpy as np

try:  # pragma: no cover - not available offline
    import gym as _gym
    from gym import spaces as _spaces
    from gym.utils import seeding as _seeding
    # only the gym 0.17-0.21 API reproduces the reference: np_random must hand back a numpy RandomState (gym >= 0.22
    # returns a Generator without randint / the legacy normal stream, and reset / step changed shape)
    HAVE_GYM = (hasattr(_seeding, "np_random") and hasattr(_gym, "Env")
                and isinstance(_seeding.np_random(0)[0], np.random.RandomState))
except Exception:  # noqa: BLE001
    HAVE_GYM = False


def _bigint_from_bytes(b):
    sizeof_int = 4
    padding = sizeof_int - len(b) % sizeof_int
    b += b"\0" * padding
    int_count = int(len(b) / sizeof_int)
    unpacked = struct.unpack("{}I".format(int_count), b)
    accum = 0
    for i, val in enumerate(unpacked):
        accum += 2 ** (sizeof_int * 8 * i) * val
    return accum


def _int_list_from_bigint(bigint):
    if bigint < 0:
        raise ValueError("Seed must be non-negative")
    if bigint == 0:
        return [0]
    ints = []
    while bigint > 0:
        bigint, mod = divmod(bigint, 2 ** 32)
        ints.append(mod)
    return ints


def create_seed(a=None, max_bytes=8):
    if a is None:
        a = _bigint_from_bytes(os.urandom(max_bytes))
    elif isinstance(a, (int, np.integer)):
        a = int(a) % 2 ** (8 * max_bytes)
    else:
        raise ValueError("Invalid type for seed: {} ({})".format(type(a), a))
    return a


def hash_seed(seed=None, max_bytes=8):
    if seed is None:
        seed = create_seed(max_bytes=max_bytes)
    h = hashlib.sha512(str(seed).encode("utf8")).digest()
    return _bigint_from_bytes(h[:max_bytes])


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, (int, np.integer)) and 0 <= seed):
        raise ValueError("Seed must be a non-negative integer or omitted, not {}".format(seed))
    seed = create_seed(seed)
    rng = np.random.RandomState()
    rng.seed(_int_list_from_bigint(hash_seed(seed)))
    return rng, seed


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random = None
        self.seed()

    def seed(self, seed=None):
        self.np_random, seed = np_random(seed)
        return [seed]


class Discrete(Space):
    def __init__(self, n):
        assert n >= 0
        self.n = n
        super().__init__((), np.int64)

    def sample(self):
        return self.np_random.randint(self.n)

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            as_int = int(x)
        elif isinstance(x, (np.generic, np.ndarray)) and (np.asarray(x).dtype.kind in "iu" and np.asarray(x).shape == ()):
            as_int = int(x)
        else:
            return False
        return 0 <= as_int < self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        low, high = np.asarray(low), np.asarray(high)
        if shape is None:
            shape = low.shape
        self.low = np.broadcast_to(low, shape).astype(dtype)
        self.high = np.broadcast_to(high, shape).astype(dtype)
        super().__init__(shape, dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

    def sample(self):
        return self.np_random.normal(size=self.shape).astype(self.dtype)

    def __repr__(self):
        return "Box" + str(self.shape)


class Env:
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        return None

    def close(self):
        return None

    def seed(self, seed=None):
        return []


class _Spaces:
    Discrete = Discrete
    Box = Box


class _Seeding:
    np_random = staticmethod(np_random)


if HAVE_GYM:  # pragma: no cover
    Env = _gym.Env  # noqa: F811
    spaces = _spaces
    seeding = _seeding
else:
    spaces = _Spaces
    seeding = _Seeding
