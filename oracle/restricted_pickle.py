"""Restricted unpickler for the reference's checkpoint (TEST INFRASTRUCTURE, see oracle/scfgp_oracle.py).

``/root/reference/experiments/boston_housing/boston_scfgp.pkl`` was written by the reference's
``SCFGP.save`` (SCFGP/SCFGP.py:296-302).  It pickles compiled Theano functions, so ``pickle.load``
needs Theano.  This loader allows NumPy array reconstruction and a few builtins only; every other
class named by the pickle is replaced by an inert stub that just records its constructor arguments
and its ``__setstate__`` payload.  Nothing from the pickle is executed.  The stub tree keeps the
whole object graph, including the ``FunctionGraph`` of each compiled function, which
``oracle/theano_graph.py`` interprets.
"""
from __future__ import annotations

import hashlib
import io
import pickle
import sys
import threading

import numpy as np

BOSTON_PKL = "/root/reference/experiments/boston_housing/boston_scfgp.pkl"
BOSTON_SHA256 = "d37d9e99ecaea0f5bef5be71a605635e29e494cc05ec8b34c253c72484fcb1f0"


class Stub:
    """Inert stand-in for any non-NumPy class named by the pickle."""

    def __new__(cls, *a, **k):
        o = object.__new__(cls)
        o.__dict__["_args"] = a
        o.__dict__["_state"] = None
        o.__dict__["_items"] = []
        return o

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, st):
        self.__dict__["_state"] = st

    def __call__(self, *a, **k):
        return Stub()

    def append(self, x):
        self._items.append(x)

    def extend(self, xs):
        self._items.extend(xs)

    def __setitem__(self, k, v):
        self._items.append((k, v))

    def __hash__(self):
        return id(self)

    def __eq__(self, o):
        return self is o


def _make_stub_class(module, name):
    return type(name, (Stub,), {"__module__": "stub." + module})


_ALLOWED = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy", "float64"), ("numpy", "int64"),
    ("numpy", "int8"), ("numpy", "bool_"),
    ("builtins", "slice"), ("builtins", "set"), ("builtins", "frozenset"),
    ("__builtin__", "slice"), ("__builtin__", "set"), ("__builtin__", "frozenset"),
    ("collections", "OrderedDict"),
}


class RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            if module.startswith("numpy.core"):
                module = module.replace("numpy.core", "numpy._core")
            if module == "__builtin__":
                module = "builtins"
            return super().find_class(module, name)
        return _make_stub_class(module, name)


def class_name(obj) -> str:
    return type(obj).__name__


def state(obj) -> dict:
    """``__setstate__`` payload of a stub (a dict for every Theano class met here)."""
    st = obj._state
    return st if isinstance(st, dict) else {}


def load_stub_tree(path: str = BOSTON_PKL, check_sha: bool = True):
    """Stub tree of the pickle.  The object graph is deep (linked Apply/Variable chains), so the
    load runs on a thread with a large stack and a raised recursion limit."""
    with open(path, "rb") as f:
        raw = f.read()
    digest = hashlib.sha256(raw).hexdigest()
    if check_sha and path == BOSTON_PKL and digest != BOSTON_SHA256:
        raise RuntimeError(f"unexpected checkpoint content (sha256 {digest})")
    box = {}

    def work():
        try:
            box["top"] = RestrictedUnpickler(io.BytesIO(raw), encoding="latin1").load()
        except BaseException as e:  # re-raised on the caller's thread
            box["err"] = e

    old_limit = sys.getrecursionlimit()
    old_stack = threading.stack_size()
    sys.setrecursionlimit(1000000)
    threading.stack_size(512 * 1024 * 1024)
    try:
        t = threading.Thread(target=work)
        t.start()
        t.join()
    finally:
        threading.stack_size(old_stack)
        sys.setrecursionlimit(old_limit)
    if "err" in box:
        raise box["err"]
    return box["top"]


def walk_arrays(obj, out=None, seen=None):
    """Every ndarray reachable from a stub tree (iterative: the tree is deep)."""
    out = [] if out is None else out
    seen = set() if seen is None else seen
    stack = [obj]
    while stack:
        o = stack.pop()
        if id(o) in seen:
            continue
        seen.add(id(o))
        if isinstance(o, np.ndarray):
            out.append(o)
        elif isinstance(o, Stub):
            stack.extend(vars(o).values())
        elif isinstance(o, dict):
            stack.extend(o.keys())
            stack.extend(o.values())
        elif isinstance(o, (list, tuple, set, frozenset)):
            stack.extend(o)
    return out
