"""A small pytree library with the semantics the reference relies on (jax.tree_util): None is an empty node; tuples,
lists, dicts (sorted keys), NamedTuples, registered dataclasses and registered node classes are containers; arrays and
scalars are leaves."""
import dataclasses
import functools

import numpy as _np

_REGISTRY = {}  # type -> (flatten(obj) -> (children, aux), unflatten(aux, children))


def register_pytree_node(node_type, /, flatten_func, unflatten_func):
    _REGISTRY[node_type] = (flatten_func, unflatten_func)


def register_pytree_node_class(node_cls, /):
    _REGISTRY[node_cls] = (lambda x: x.tree_flatten(), node_cls.tree_unflatten)
    return node_cls


def register_dataclass(datacls):
    fields = dataclasses.fields(datacls)
    data = [f.name for f in fields if not f.metadata.get("static", False)]
    meta = [f.name for f in fields if f.metadata.get("static", False)]

    def flatten(x):
        return [getattr(x, n) for n in data], tuple(getattr(x, n) for n in meta)

    def unflatten(aux, children):
        return datacls(**dict(zip(data, children)), **dict(zip(meta, aux)))

    _REGISTRY[datacls] = (flatten, unflatten)
    return datacls


class Partial(functools.partial):
    """functools.partial that is a pytree: the bound arguments are children, the function is static."""


def _partial_flatten(p):
    return [p.args, p.keywords], p.func


def _partial_unflatten(func, children):
    args, kwargs = children
    return Partial(func, *args, **kwargs)


_REGISTRY[Partial] = (_partial_flatten, _partial_unflatten)


class _Def:
    """Structure of a pytree: kind + aux + child structures (leaf: kind 'leaf')."""

    __slots__ = ("kind", "aux", "children")

    def __init__(self, kind, aux, children):
        self.kind, self.aux, self.children = kind, aux, children

    def __eq__(self, other):
        return (isinstance(other, _Def) and self.kind == other.kind and _aux_eq(self.aux, other.aux)
                and self.children == other.children)  # fmt: skip

    def __hash__(self):
        return hash((str(self.kind), len(self.children)))

    @property
    def num_leaves(self):
        return 1 if self.kind == "leaf" else sum(c.num_leaves for c in self.children)

    def __repr__(self):
        return "*" if self.kind == "leaf" else f"{getattr(self.kind, '__name__', self.kind)}{self.children}"


def _aux_eq(a, b):
    try:
        r = a == b
        return bool(r) if not hasattr(r, "all") else bool(r.all())
    except Exception:
        return a is b


_LEAF = _Def("leaf", None, ())


def _split(x):
    """(kind, aux, children) of a container, or None for a leaf."""
    if x is None:
        return ("none", None, [])
    t = type(x)
    if t in _REGISTRY:
        children, aux = _REGISTRY[t][0](x)
        return (t, aux, list(children))
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return (t, None, list(x))
    if t is tuple:
        return ("tuple", None, list(x))
    if t is list:
        return ("list", None, list(x))
    if t is dict:
        keys = sorted(x)
        return ("dict", tuple(keys), [x[k] for k in keys])
    return None


def _build(kind, aux, children):
    if kind == "none":
        return None
    if kind == "tuple":
        return tuple(children)
    if kind == "list":
        return list(children)
    if kind == "dict":
        return dict(zip(aux, children))
    if kind in _REGISTRY:
        return _REGISTRY[kind][1](aux, children)
    return kind(*children)  # NamedTuple


def tree_flatten(tree, is_leaf=None):
    leaves = []

    def rec(x):
        if is_leaf is not None and is_leaf(x):
            leaves.append(x)
            return _LEAF
        s = _split(x)
        if s is None:
            leaves.append(x)
            return _LEAF
        kind, aux, children = s
        return _Def(kind, aux, tuple(rec(c) for c in children))

    return leaves, rec(tree)


def tree_unflatten(structure, leaves, /):
    it = iter(leaves)

    def rec(d):
        if d.kind == "leaf":
            return next(it)
        return _build(d.kind, d.aux, [rec(c) for c in d.children])

    return rec(structure)


def tree_structure(tree, /):
    return tree_flatten(tree)[1]


def tree_leaves(tree, /):
    return tree_flatten(tree)[0]


def tree_map(func, tree, *rest, is_leaf=None):
    def rec(x, others):
        if is_leaf is not None and is_leaf(x):
            return func(x, *others)
        s = _split(x)
        if s is None:
            return func(x, *others)
        kind, aux, children = s
        other_children = []
        for o in others:
            so = _split(o)
            if so is None or len(so[2]) != len(children):
                raise ValueError(f"tree_map: structures differ at {type(x).__name__}: {x!r} vs {o!r}")
            other_children.append(so[2])
        return _build(kind, aux, [rec(c, [oc[i] for oc in other_children]) for i, c in enumerate(children)])

    return rec(tree, list(rest))


def tree_all(tree, /):
    return all(bool(x) for x in tree_leaves(tree))


def ravel_pytree(tree, /):
    leaves, structure = tree_flatten(tree)
    arrs = [_np.asarray(x) for x in leaves]
    shapes = [a.shape for a in arrs]
    sizes = [a.size for a in arrs]
    flat = _np.concatenate([a.reshape(-1) for a in arrs]) if arrs else _np.zeros((0,))

    def unravel(v):
        out, k = [], 0
        for shp, sz in zip(shapes, sizes):
            out.append(_np.reshape(v[k : k + sz], shp))
            k += sz
        return tree_unflatten(structure, out)

    return flat, unravel


def tree_flatten_depth_one(tree, /):
    ref = tree_structure(tree[0])
    return tree_flatten(tree, is_leaf=lambda x: tree_structure(x) == ref)


def tree_leaves_depth_one(tree, /):
    return tree_flatten_depth_one(tree)[0]


def _transpose(list_of_trees):
    return tree_map(lambda *xs: list(xs), *list_of_trees)


def _is_array_list(x):
    return isinstance(x, list) and len(x) > 0 and isinstance(x[0], (_np.ndarray, _np.generic, float, int,
                                                                     _np.random.Generator))  # fmt: skip


def _stack_leaves(xs, combine):
    if isinstance(xs[0], _np.random.Generator):  # PRNG state carried by a stochastic Jacobian: an array of generators
        out = _np.empty((len(xs),), dtype=object)
        for i, x in enumerate(xs):
            out[i] = x
        return out
    return combine([_np.asarray(x) for x in xs])


def tree_array_concatenate(list_of_trees):
    return tree_map(lambda xs: _np.concatenate([_np.asarray(x) for x in xs]), _transpose(list_of_trees),
                    is_leaf=_is_array_list)  # fmt: skip


def tree_array_stack(list_of_trees):
    return tree_map(lambda xs: _stack_leaves(xs, _np.stack), _transpose(list_of_trees), is_leaf=_is_array_list)


def tree_array_prepend(y, X, /):
    Y = tree_map(lambda s: _np.asarray(s)[None, ...], y)
    return tree_array_concatenate([Y, X])


def tree_array_append(X, y, /):
    Y = tree_map(lambda s: _np.asarray(s)[None, ...], y)
    return tree_array_concatenate([X, Y])
