import abc


class ABC(abc.ABC):  # noqa: B024
    pass


def abstractmethod(*args):
    return abc.abstractmethod(*args)
