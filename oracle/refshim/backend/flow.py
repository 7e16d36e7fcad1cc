"""Control flow in plain Python with the calling conventions of jax.lax."""
import numpy as _np

from oracle.refshim.backend import tree as _tree


def _index(tree, i):
    return _tree.tree_map(lambda s: _np.asarray(s)[i], tree)


def _length(xs, length):
    leaves = _tree.tree_leaves(xs)
    return int(length) if length is not None else int(_np.shape(leaves[0])[0])


def scan(step_func, /, init, xs, *, reverse=False, length=None):
    n = _length(xs, length)
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        carry, y = step_func(carry, None if xs is None else _index(xs, i))
        ys[i] = y
    if n == 0 or all(y is None for y in ys):
        return carry, None
    return carry, _tree.tree_array_stack(ys)


def fori_loop(lower, upper, step_func, /, init):
    val = init
    for i in range(int(lower), int(upper)):
        val = step_func(i, val)
    return val


def while_loop(cond_func, body_func, /, init):
    val = init
    while bool(cond_func(val)):
        val = body_func(val)
    return val


def cond(use_true_func, true_func, false_func, *operands):
    return true_func(*operands) if bool(use_true_func) else false_func(*operands)


def switch(index, options, args):
    return options[int(index)](args)
