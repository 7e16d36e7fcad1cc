from collections.abc import Callable, Sequence  # noqa: F401
from typing import TYPE_CHECKING, Any, Generic, Literal, Protocol, TypeVar  # noqa: F401

from numpy import ndarray as Array  # noqa: F401
