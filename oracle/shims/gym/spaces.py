"""TEST INFRASTRUCTURE ONLY (oracle/): the four `gym.spaces` containers the reference touches
(pypownet/environment.py:46-59 ActionSpace(MultiBinary), :277-374 ObservationSpace(Dict) with
Box/MultiBinary/Discrete leaves, read back through `.spaces`, `.shape`, `.n`)."""
from collections import OrderedDict
import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError


class MultiBinary(Space):
    def __init__(self, n):
        self.n = n
        super().__init__((self.n,), np.int8)

    def sample(self):
        return np.random.randint(low=0, high=2, size=self.n).astype(self.dtype)

    def contains(self, x):
        return ((np.asarray(x) == 0) | (np.asarray(x) == 1)).all()


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        super().__init__((), np.int64)

    def sample(self):
        return np.random.randint(self.n)

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box(Space):
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        if shape is None:
            shape = np.shape(low)
        self.low = low + np.zeros(shape)
        self.high = high + np.zeros(shape)
        super().__init__(shape, np.float32 if dtype is None else dtype)

    def sample(self):
        return np.random.uniform(size=self.shape).astype(self.dtype)

    def contains(self, x):
        return np.shape(x) == self.shape


class Dict(Space):
    def __init__(self, spaces=None, **spaces_kwargs):
        if isinstance(spaces, dict) and not isinstance(spaces, OrderedDict):
            spaces = OrderedDict(sorted(list(spaces.items())))
        if isinstance(spaces, list):
            spaces = OrderedDict(spaces)
        self.spaces = spaces
        super().__init__(None, None)

    def sample(self):
        return OrderedDict([(k, s.sample()) for k, s in self.spaces.items()])

    def contains(self, x):
        return isinstance(x, dict) and len(x) == len(self.spaces)
