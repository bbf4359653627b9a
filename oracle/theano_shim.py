"""A torch-CPU float64 stand-in for the slice of the Theano API the reference uses, so that the
reference's UNMODIFIED Python sources (``/root/reference/SCFGP/{SCFGP,Optimizer,Scaler}.py``) execute
in the build container.  TEST INFRASTRUCTURE (see oracle/scfgp_oracle.py): used only by
``tests/golden/make_reference_source_fixtures.py`` and ``tests/test_reference_source.py``.

What the reference asks of Theano (SCFGP.py:11-12,64-148; Optimizer.py:8-9,27-382):
``config``, ``shared`` (``get_value``, ``broadcastable``, ``copy``, indexing), ``function(inputs, outputs,
givens=, updates=)``, ``tensor.{dmatrices, dvector, constant, reshape, mean, std, sum, exp, log, sqrt, cos,
sin, dot, concatenate, identity_like, diagonal, maximum, grad}``, operators on variables (``+ - * / ** abs``,
``.T``, ``.dot``, ``.sum``, ``.shape``, ``[...]``) and ``sandbox.linalg.{cholesky, matrix_inverse}``.

Design: a *lazy* expression graph like Theano's -- every ``Var`` holds a closure that evaluates it with
torch in an environment mapping inputs to tensors; ``function`` binds the inputs, evaluates the outputs,
then all ``updates`` from the OLD values (Theano semantics) and only then stores them.  ``grad`` is
reverse-mode autograd through the same closures (torch.autograd), where Theano differentiates
symbolically: same mathematics, float64 round-off differs.  ``oracle/theano_graph.py`` (Theano's own
compiled graph from the shipped pickle) is the independent check that this backend computes what
Theano computes (``tests/test_reference_source.py::test_shim_run_matches_pickled_theano_graph``).

``install()`` is a context manager that puts fake ``theano`` / ``matplotlib`` modules into
``sys.modules`` for the duration of an import of the reference package and removes them afterwards.
"""
from __future__ import annotations

import contextlib
import importlib.util
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

_F64 = torch.float64


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x))


class Var:
    """A symbolic variable: ``fn(env) -> torch tensor``; results are memoised per evaluation."""

    broadcastable = ()

    def __init__(self, fn, name=None):
        self._fn = fn
        self.name = name

    def ev(self, env):
        key = id(self)
        if key not in env.cache:
            env.cache[key] = self._fn(env)
        return env.cache[key]

    # ---- arithmetic
    def _bin(self, other, op, reverse=False):
        o = as_var(other)
        a, b = (o, self) if reverse else (self, o)
        return Var(lambda env: op(a.ev(env), b.ev(env)))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.true_divide)
    def __rtruediv__(self, o): return self._bin(o, torch.true_divide, True)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __rpow__(self, o): return self._bin(o, torch.pow, True)
    def __neg__(self): return Var(lambda env: -self.ev(env))
    def __abs__(self): return Var(lambda env: torch.abs(self.ev(env)))

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)

        def run(env):
            resolved = tuple(slice(_ix(i.start, env), _ix(i.stop, env), _ix(i.step, env)) if isinstance(i, slice)
                             else _ix(i, env) for i in idx)
            return self.ev(env)[resolved]
        return Var(run)

    # ---- tensor methods
    @property
    def T(self):
        return Var(lambda env: self.ev(env).T)

    @property
    def shape(self):
        return _Shape(self)

    def dot(self, o):
        return dot(self, o)

    def sum(self, axis=None):
        return sum_(self, axis)

    def copy(self):
        # Theano: a symbolic copy node -- an alias of the expression, not a snapshot
        return Var(lambda env: self.ev(env).clone())

    __hash__ = object.__hash__


def _ix(i, env):
    if isinstance(i, Var):
        return int(i.ev(env))
    return i


class _Shape:
    def __init__(self, v):
        self.v = v

    def __getitem__(self, k):
        return Var(lambda env: torch.tensor(self.v.ev(env).shape[k], dtype=torch.int64))


class _Env:
    def __init__(self):
        self.cache = {}


def as_var(x):
    if isinstance(x, Var):
        return x
    c = _t(x)
    if c.dtype in (torch.float32, torch.float16):
        c = c.to(_F64)
    return Var(lambda env: c)


class Shared(Var):
    """``theano.shared``: a variable with persistent storage."""

    def __init__(self, value, name=None, broadcastable=None, **kw):
        self.value = np.array(value, copy=True)
        self.broadcastable = tuple(broadcastable) if broadcastable is not None else (False,) * self.value.ndim
        super().__init__(lambda env: torch.as_tensor(self.value), name)

    def get_value(self, borrow=False):
        return self.value if borrow else self.value.copy()

    def set_value(self, v, borrow=False):
        self.value = np.array(v, dtype=self.value.dtype, copy=True)


class _Input(Var):
    def __init__(self, name, ndim):
        self.ndim = ndim
        super().__init__(self._missing, name)

    def _missing(self, env):
        raise RuntimeError(f"input {self.name!r} was not bound")


# -------------------------------------------------------------------------------- theano.tensor
def _unary(fn):
    return lambda x: (lambda v: Var(lambda env: fn(v.ev(env))))(as_var(x))


exp, log, sqrt, cos, sin = map(_unary, (torch.exp, torch.log, torch.sqrt, torch.cos, torch.sin))


def dmatrices(*names):
    return [_Input(n, 2) for n in names]


def dvector(name=None):
    return _Input(name, 1)


def constant(x, **kw):
    return as_var(np.asarray(x))


def reshape(x, shape):
    x = as_var(x)
    return Var(lambda env: x.ev(env).reshape(tuple(int(s) for s in shape)))


def mean(x, axis=None):
    x = as_var(x)
    return Var(lambda env: x.ev(env).mean() if axis is None else x.ev(env).mean(axis))


def std(x, axis=None):
    x = as_var(x)                                   # theano.tensor.std: population (ddof = 0)
    return Var(lambda env: x.ev(env).std(unbiased=False) if axis is None else x.ev(env).std(axis, unbiased=False))


def sum_(x, axis=None):
    x = as_var(x)
    return Var(lambda env: x.ev(env).sum() if axis is None else x.ev(env).sum(axis))


def dot(a, b):
    a, b = as_var(a), as_var(b)
    return Var(lambda env: a.ev(env) @ b.ev(env))


def concatenate(parts, axis=0):
    parts = [as_var(p) for p in parts]
    return Var(lambda env: torch.cat([p.ev(env) for p in parts], axis))


def identity_like(x):
    x = as_var(x)
    return Var(lambda env: torch.eye(x.ev(env).shape[0], x.ev(env).shape[1], dtype=_F64))


def diagonal(x):
    x = as_var(x)
    return Var(lambda env: torch.diagonal(x.ev(env)))


def maximum(a, b):
    a, b = as_var(a), as_var(b)
    return Var(lambda env: torch.maximum(a.ev(env), b.ev(env)))


def grad(cost, wrt):
    """``theano.tensor.grad``: here reverse-mode autograd of the closures at evaluation time."""
    def run(env):
        w = wrt.ev(env)
        if not w.requires_grad:
            raise RuntimeError("grad: wrt is not a differentiable leaf in this function")
        c = cost.ev(env)
        (g,) = torch.autograd.grad(c, w, retain_graph=True)
        return g
    return Var(run)


# -------------------------------------------------------------------------------- sandbox.linalg
def cholesky(a):
    a = as_var(a)
    return Var(lambda env: torch.linalg.cholesky(a.ev(env)))          # lower, as Theano's default


def matrix_inverse(a):
    a = as_var(a)
    return Var(lambda env: torch.linalg.inv(a.ev(env)))


# -------------------------------------------------------------------------------- theano.function
class Function:
    def __init__(self, inputs, outputs, givens=None, updates=None, **kw):
        self.inputs = list(inputs)
        self.outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        self.single = not isinstance(outputs, (list, tuple))
        self.givens = list(givens.items()) if isinstance(givens, dict) else list(givens or [])
        self.updates = list(updates.items()) if updates is not None else []
        self.needs_grad = bool(self.updates)

    def __call__(self, *args):
        if len(args) != len(self.inputs):
            raise TypeError("wrong number of inputs")
        env = _Env()
        for v, a in zip(self.inputs, args):
            a = np.asarray(a, dtype=np.float64)
            if a.ndim != v.ndim:
                raise TypeError(f"Wrong number of dimensions for input {v.name!r}")
            env.cache[id(v)] = torch.as_tensor(a)
        for sym, val in self.givens:
            t = as_var(val).ev(env).detach().clone()
            env.cache[id(sym)] = t.requires_grad_(True) if t.dtype == _F64 else t
        outs = [o.ev(env) for o in self.outputs]
        new_values = [(s, e.ev(env)) for s, e in self.updates]   # all from the old state
        for s, val in new_values:
            s.set_value(val.detach().numpy())
        res = [o.detach().numpy().copy() for o in outs]
        return res[0] if self.single else res


def function(inputs, outputs, givens=None, updates=None, **kw):
    return Function(inputs, outputs, givens=givens, updates=updates, **kw)


def shared(value, name=None, **kw):
    return Shared(value, name=name, **kw)


# -------------------------------------------------------------------------------- module plumbing
def _modules():
    theano = types.ModuleType("theano")
    theano.config = types.SimpleNamespace(floatX="float64", mode=None, optimizer=None,
                                          reoptimize_unpickled_function=None)
    theano.shared, theano.function = shared, function
    tensor = types.ModuleType("theano.tensor")
    for name, obj in dict(dmatrices=dmatrices, dvector=dvector, constant=constant, reshape=reshape, mean=mean,
                          std=std, sum=sum_, exp=exp, log=log, sqrt=sqrt, cos=cos, sin=sin, dot=dot,
                          concatenate=concatenate, identity_like=identity_like, diagonal=diagonal,
                          maximum=maximum, grad=grad).items():
        setattr(tensor, name, obj)
    theano.tensor = tensor
    sandbox = types.ModuleType("theano.sandbox")
    linalg = types.ModuleType("theano.sandbox.linalg")
    linalg.cholesky, linalg.matrix_inverse = cholesky, matrix_inverse
    sandbox.linalg = linalg
    theano.sandbox = sandbox
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")          # imported at module level, used only when plotting
    mpl.pyplot = plt
    return {"theano": theano, "theano.tensor": tensor, "theano.sandbox": sandbox,
            "theano.sandbox.linalg": linalg, "matplotlib": mpl, "matplotlib.pyplot": plt}


@contextlib.contextmanager
def install():
    mods = _modules()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        yield mods
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@contextlib.contextmanager
def numpy1_names():
    """The reference predates NumPy 2: ``np.Infinity`` (SCFGP.py:224) was removed.  Restored for the
    duration of a reference call; the reference's sources stay untouched."""
    missing = {"Infinity": np.inf}
    added = [k for k in missing if not hasattr(np, k)]
    for k in added:
        setattr(np, k, missing[k])
    try:
        yield
    finally:
        for k in added:
            delattr(np, k)


def import_reference(root="/root/reference", alias="reference_scfgp_pkg"):
    """Import the reference package from its source tree under an alias, with the shim modules in place
    while its top-level imports run.  Returns the package module (``.SCFGP``, ``.Scaler``, ``.Optimizer``)."""
    import os
    pkg_dir = os.path.join(root, "SCFGP")
    with install():
        spec = importlib.util.spec_from_file_location(alias, os.path.join(pkg_dir, "__init__.py"),
                                                      submodule_search_locations=[pkg_dir])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules[alias] = pkg
        try:
            spec.loader.exec_module(pkg)
        except BaseException:
            sys.modules.pop(alias, None)
            raise
    return pkg
