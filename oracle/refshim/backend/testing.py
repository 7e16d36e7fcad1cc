"""pytest glue of the reference's own test-suite (reference: probdiffeq/backend/testing.py), so that those of the
reference's tests that do not need `pytest_cases` (not installed here) can be run on this backend -- which is how the
backend itself is validated (oracle/refshim/run_reference_tests.py). Not needed to run the step loop."""
import numpy as _np
import pytest as _pytest

from oracle.refshim.backend import tree as _tree

filterwarnings = _pytest.mark.filterwarnings
parametrize = _pytest.mark.parametrize
raises = _pytest.raises
warns = _pytest.warns
xfail = _pytest.xfail


def skip(reason):
    return _pytest.skip(reason=reason)


# --- a small stand-in for the three pytest_cases functions the reference's suite uses (pytest_cases is not installed)


def _parametrize_marks(fn):
    return [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]


def _mark_names(mark):
    return [n.strip() for n in mark.args[0].split(",")] if isinstance(mark.args[0], str) else list(mark.args[0])


def _mark_values(mark):
    """The value tuples of one parametrize mark (a `pytest.param(...)` entry contributes its values)."""
    multi = len(_mark_names(mark)) > 1
    out = []
    for v in mark.args[1]:
        v = v.values if hasattr(v, "values") and hasattr(v, "marks") else (tuple(v) if multi else (v,))
        out.append(tuple(v))
    return out


def fixture(name=None, scope="function"):
    """`pytest_cases.fixture`: a fixture that may itself carry `parametrize` (and `parametrize_with_cases`) marks --
    turned into a parametrised pytest fixture over the product of the marks' values."""
    import inspect
    import itertools

    def decorate(fn):
        marks = _parametrize_marks(fn)
        if not marks:
            return _pytest.fixture(name=name, scope=scope)(fn)
        names = [n for m in marks for n in _mark_names(m)]
        combos = [tuple(x for group in combo for x in group)
                  for combo in itertools.product(*[_mark_values(m) for m in marks])]  # fmt: skip
        sig = inspect.signature(fn)
        others = [p for n, p in sig.parameters.items() if n not in names]

        def produce(request, *args, **kwargs):
            bound = {p.name: a for p, a in zip(others, args)}
            bound.update(kwargs)
            return fn(**dict(zip(names, request.param)), **bound)

        produce.__name__ = fn.__name__
        produce.__module__ = fn.__module__
        request_param = inspect.Parameter("request", inspect.Parameter.POSITIONAL_OR_KEYWORD)
        produce.__signature__ = sig.replace(parameters=[request_param, *[p.replace(kind=inspect.Parameter.POSITIONAL_OR_KEYWORD)
                                                                         for p in others]])  # fmt: skip
        return _pytest.fixture(name=name, scope=scope, params=combos)(produce)

    return decorate


def case(*args, **_kwargs):
    """`@case` / `@case(id=..., tags=...)`: marks a case function; nothing to record for this stand-in."""
    if len(args) == 1 and callable(args[0]):
        return args[0]
    return lambda fn: fn


def parametrize_with_cases(argnames, cases=".", prefix="case_", **_kwargs):
    """Parametrise a test with the return values of the functions named `prefix*` of the test's own module (the only
    form the reference's suite uses: `cases="."`). A case function's own `parametrize` marks are expanded and the fixtures
    it asks for are requested by the test on its behalf; the case is evaluated when the test runs."""
    import inspect
    import itertools

    assert cases == ".", "only same-module cases are supported"

    def decorate(test):
        module = getattr(test, "_case_module", test.__globals__)  # (a stacked decorator sees the wrapper below)
        params, case_fixtures = [], set()
        for name, fn in module.items():
            if not (name.startswith(prefix) and inspect.isfunction(fn)):
                continue
            marks = _parametrize_marks(fn)
            names = [n for m in marks for n in _mark_names(m)]
            wanted = [n for n in inspect.signature(fn).parameters if n not in names]  # fixtures the case asks for
            case_fixtures.update(wanted)

            for combo in itertools.product(*[_mark_values(m) for m in marks]):
                flat = [x for group in combo for x in group]
                ident = name[len(prefix):] + (f"-{len(params)}" if flat else "")
                params.append(_pytest.param(_LazyCase(fn, dict(zip(names, flat)), wanted), id=ident))
        argnames_list = [a.strip() for a in argnames.split(",")] if isinstance(argnames, str) else list(argnames)
        case_arg = "_case_" + "_".join(argnames_list)  # one per decorator: a test may stack several

        import functools

        @functools.wraps(test)
        def run(*args, **kwargs):
            lazy = kwargs.pop(case_arg)
            value = lazy(kwargs)
            for extra in only_for_cases:
                kwargs.pop(extra)
            if len(argnames_list) == 1:
                kwargs[argnames_list[0]] = value
            else:
                kwargs.update(dict(zip(argnames_list, value)))
            return test(*args, **kwargs)

        # the wrapper's signature: the test's parameters minus the case arguments, plus the lazy case
        sig = inspect.signature(test)
        keep = [p for n, p in sig.parameters.items() if n not in argnames_list]
        only_for_cases = sorted(case_fixtures - set(sig.parameters))
        keep += [inspect.Parameter(n, inspect.Parameter.POSITIONAL_OR_KEYWORD) for n in only_for_cases]
        keep.append(inspect.Parameter(case_arg, inspect.Parameter.KEYWORD_ONLY))
        run.__signature__ = sig.replace(parameters=keep)
        del run.__wrapped__
        run._case_module = module
        return _pytest.mark.parametrize(case_arg, params)(run)

    return decorate


class _LazyCase:
    def __init__(self, fn, kwargs, fixtures):
        self.fn, self.kwargs, self.fixtures = fn, kwargs, fixtures

    def __call__(self, available):
        return self.fn(**self.kwargs, **{n: available[n] for n in self.fixtures})


def _allclose(a, b, /, *, atol, rtol, strict_shapes):
    a, b = _np.asarray(1.0 * _np.asarray(a)), _np.asarray(1.0 * _np.asarray(b))
    # The reference derives its default tolerances from the dtype of the leaves, and its suite runs in JAX's default
    # SINGLE precision (no x64 switch in its test configuration): sqrt(eps_float32). The same acceptance thresholds
    # are used here although the arithmetic is float64 -- several of those tests compare a solver's output with an
    # exact solution, where the threshold has to cover the discretisation error, not rounding.
    tol = _np.sqrt(_np.finfo(_np.float32).eps)
    atol = tol if atol is None else atol
    rtol = 10 * tol if rtol is None else rtol
    close = bool(_np.allclose(a, b, atol=atol, rtol=rtol))
    return close and (a.shape == b.shape or not strict_shapes)


def allclose(tree1, tree2, /, *, atol=None, rtol=None, strict_shapes=True):
    """Pytree-aware allclose with tolerances proportional to sqrt(machine epsilon), like the reference's."""
    flags = _tree.tree_map(lambda a, b: _allclose(a, b, atol=atol, rtol=rtol, strict_shapes=strict_shapes), tree1, tree2)
    return _tree.tree_all(flags)


def marginals_allclose(m1, m2, /, *, atol=None, rtol=None, strict_shapes=True):
    mean1, cov1 = m1.to_multivariate_normal()
    mean2, cov2 = m2.to_multivariate_normal()
    return (_allclose(mean1, mean2, atol=atol, rtol=rtol, strict_shapes=strict_shapes)
            and _allclose(cov1, cov2, atol=atol, rtol=rtol, strict_shapes=strict_shapes))  # fmt: skip
