import time


def perf_counter():
    return time.perf_counter()


def repeat(fun, /, *, repeats):
    import timeit

    return list(timeit.repeat(fun, number=1, repeat=repeats))
