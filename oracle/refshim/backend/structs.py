import dataclasses
from enum import Enum  # noqa: F401
from typing import NamedTuple  # noqa: F401

import numpy as _np


class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), _np.dtype(dtype)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(_np.prod(self.shape)) if self.shape else 1

    def __repr__(self):
        return f"ShapeDtypeStruct(shape={self.shape}, dtype={self.dtype})"


def dataclass(*args, **kwargs):
    return dataclasses.dataclass(*args, frozen=True, **kwargs)


def dataclass_field(metadata):
    return dataclasses.field(metadata=metadata)
