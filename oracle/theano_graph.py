"""NumPy interpreter of the reference's own *compiled Theano graphs* (TEST INFRASTRUCTURE).

The reference cannot run here (Theano is not installed), but its shipped checkpoint
``experiments/boston_housing/boston_scfgp.pkl`` (written by ``SCFGP.save``,
SCFGP/SCFGP.py:296-302) pickles the compiled functions themselves:

* ``train_func``  = ``theano.function([X, y], [cost, alpha, Li], givens=[(params, self.params)])``
  (SCFGP/SCFGP.py:132-135): a ``FunctionGraph`` of 85 ``Apply`` nodes after Theano's ``fast_run``
  optimiser (Dot22/Gemm, fused ``Composite`` element-wise loops, ``Cholesky``, ``MatrixInverse``,
  three ``Solve`` nodes, the 30-node Gauss-Hermite tensors with their weights/nodes as shared
  arrays);
* ``pred_func``   = ``theano.function([Xs, alpha, Li], [mu_pred, std_pred], givens=...)``
  (SCFGP/SCFGP.py:145-148): 37 nodes.

``oracle/restricted_pickle.py`` recovers both graphs as inert stubs (op class name + op state,
inputs, outputs, constants).  This module walks them in topological order and executes each node
with the NumPy/SciPy routine Theano's ``perform`` dispatches to.  The result is *the reference's
own function, evaluated on arbitrary inputs*: ``(X, y) -> cost, alpha, Li`` at any parameter vector
of the checkpoint's geometry (D=13, S=20, M=50; the shapes are constants of the graph) and
``(Xs, alpha, Li) -> mu, std``.  ``tests/golden/make_reference_graph_fixtures.py`` uses it to
produce reference outputs that pin ``oracle/scfgp_oracle.py`` and the CUDA kernels
(``tests/test_reference_graph.py``, ``tests/test_gpu_reference_graph.py``).

Only ``tests/`` and the fixture script import this module.  It needs ``/root/reference`` and so
only runs in the build container; the committed fixtures travel.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from .restricted_pickle import BOSTON_PKL, Stub, class_name, load_stub_tree, state

# --------------------------------------------------------------------------------------
# scalar ops (theano/scalar/basic.py ``impl`` methods; ScalarSoftplus from
# theano/tensor/nnet/sigm.py)
# --------------------------------------------------------------------------------------


def _softplus(x):
    # ScalarSoftplus.impl: 0 below -30, x above 30, log1p(exp(x)) in between
    x = np.asarray(x, dtype=np.float64)
    mid = np.log1p(np.exp(np.clip(x, -30.0, 30.0)))
    return np.where(x < -30.0, 0.0, np.where(x > 30.0, x, mid))


def _nary(fn):
    def run(*xs):
        out = xs[0]
        for x in xs[1:]:
            out = fn(out, x)
        return out
    return run


_SCALAR_IMPL = {
    "Add": _nary(np.add),
    "Mul": _nary(np.multiply),
    "Sub": np.subtract,
    "TrueDiv": np.true_divide,
    "IntDiv": np.floor_divide,
    "Neg": np.negative,
    "Abs": np.abs,
    "Sgn": np.sign,
    "Inv": lambda x: 1.0 / x,
    "Sqr": np.square,
    "Sqrt": np.sqrt,
    "Exp": np.exp,
    "Log": np.log,
    "Log1p": np.log1p,
    "Cos": np.cos,
    "Sin": np.sin,
    "Tanh": np.tanh,
    "Pow": np.power,
    "Maximum": np.maximum,
    "Minimum": np.minimum,
    "Identity": lambda x: x,
    "Second": lambda x, y: np.broadcast_to(y, np.broadcast(x, y).shape),
    "Switch": lambda c, a, b: np.where(c, a, b),
    "GE": np.greater_equal, "GT": np.greater, "LE": np.less_equal, "LT": np.less,
    "EQ": np.equal, "NEQ": np.not_equal,
    "ScalarSoftplus": _softplus,
}


def _owner(var):
    return state(var).get("owner")


def _constant_value(var):
    st = state(var)
    if "data" in st:
        return np.asarray(st["data"])
    raise KeyError(f"constant without data: {class_name(var)} keys {list(st)}")


def _toposort(outputs):
    order, done = [], set()
    stack = [(o, False) for o in reversed(outputs)]
    while stack:
        var, expanded = stack.pop()
        node = _owner(var)
        if node is None or id(node) in done:
            continue
        if expanded:
            done.add(id(node))
            order.append(node)
        else:
            stack.append((var, True))
            for i in reversed(state(node)["inputs"]):
                stack.append((i, False))
    return order


def _run_scalar_op(op, args):
    name = class_name(op)
    if name == "Composite":
        return _run_composite(op, args)
    if name == "Cast":
        return np.asarray(args[0]).astype(state(state(op)["o_type"])["dtype"])
    try:
        return _SCALAR_IMPL[name](*args)
    except KeyError:
        raise NotImplementedError(f"scalar op {name}") from None


def _run_composite(op, args):
    """``Composite``: a fused scalar sub-graph (theano/scalar/basic.py); evaluated with array
    arguments so that one call covers the whole element-wise loop."""
    st = state(op)
    env = {id(v): a for v, a in zip(st["inputs"], args)}

    def value(v):
        if id(v) in env:
            return env[id(v)]
        if _owner(v) is None:
            env[id(v)] = _constant_value(v)
            return env[id(v)]
        raise KeyError("unevaluated scalar variable")

    for node in _toposort(st["outputs"]):
        nst = state(node)
        res = _run_scalar_op(nst["op"], [value(i) for i in nst["inputs"]])
        outs = nst["outputs"]
        if len(outs) == 1:
            env[id(outs[0])] = res
        else:
            for o, r in zip(outs, res):
                env[id(o)] = r
    res = [env[id(o)] for o in st["outputs"]]
    return res[0] if len(res) == 1 else res


# --------------------------------------------------------------------------------------
# tensor ops (the ``perform`` of each Op class met in the two graphs)
# --------------------------------------------------------------------------------------
def _dimshuffle(op, x):
    st = state(op)
    x = np.asarray(x)
    order = st["new_order"]
    kept = [d for d in order if d != "x"]
    drop = [d for d in range(x.ndim) if d not in kept]
    for d in drop:
        if x.shape[d] != 1:
            raise ValueError("DimShuffle drops a non-broadcastable dimension")
    y = x.transpose(kept + drop).reshape([x.shape[d] for d in kept])
    shape, it = [], iter(y.shape)
    for d in order:
        shape.append(1 if d == "x" else next(it))
    return y.reshape(shape)


def _subtensor(op, x, index_inputs):
    it = iter(index_inputs)

    def entry(e):
        if isinstance(e, slice):
            return slice(entry(e.start), entry(e.stop), entry(e.step))
        if e is None or isinstance(e, (int, np.integer)):
            return e
        if isinstance(e, Stub):  # a scalar Type: the value is the next input
            return int(next(it))
        raise NotImplementedError(f"subtensor entry {e!r}")

    idx = tuple(entry(e) for e in state(op)["idx_list"])
    return np.asarray(x)[idx]


def _solve(op, A, b):
    structure = state(op)["A_structure"]
    if structure == "lower_triangular":
        return scipy.linalg.solve_triangular(A, b, lower=True)
    if structure == "upper_triangular":
        return scipy.linalg.solve_triangular(A, b, lower=False)
    return scipy.linalg.solve(A, b)


def _run_node(node, args):
    op = state(node)["op"]
    name = class_name(op)
    st = state(op)
    if name == "Elemwise":
        return _run_scalar_op(st["scalar_op"], args)
    if name == "DimShuffle":
        return _dimshuffle(op, args[0])
    if name == "Sum":
        axis = st["axis"]
        return np.sum(args[0], axis=None if axis is None else tuple(axis), dtype=st.get("acc_dtype") or None)
    if name == "Subtensor":
        return _subtensor(op, args[0], args[1:])
    if name == "Reshape":
        return np.reshape(args[0], tuple(int(v) for v in args[1]))
    if name in ("Dot22", "Dot"):
        return np.dot(args[0], args[1])
    if name == "Gemm":  # b*z + a*dot(x, y)
        z, a, x, y, b = args
        return b * z + a * np.dot(x, y)
    if name == "Join":
        return np.concatenate(args[1:], axis=int(args[0]))
    if name == "Solve":
        return _solve(op, args[0], args[1])
    if name == "Cholesky":
        return scipy.linalg.cholesky(args[0], lower=bool(st["lower"]))
    if name == "MatrixInverse":
        return np.linalg.inv(args[0])
    if name == "ExtractDiag":
        return np.diagonal(args[0]).copy()
    if name == "Shape_i":
        return np.int64(np.shape(args[0])[st["i"]])
    if name == "Shape":
        return np.asarray(np.shape(args[0]), dtype=np.int64)
    if name == "MakeVector":
        return np.asarray(args)
    raise NotImplementedError(f"tensor op {name}")


class GraphFunction:
    """One pickled ``theano.compile.function_module.Function`` made callable.

    ``explicit``: names of the inputs the caller passes (e.g. ``['X', 'y']``); ``implicit``: the
    shared-variable inputs with the values stored in the checkpoint (``train_func``: the trained
    parameter vector, the Gauss-Hermite weights / sqrt(pi) and nodes of SCFGP/SCFGP.py:118-120).
    ``__call__(*explicit, shared={position: array})`` overrides shared values by input position.
    """

    def __init__(self, function_stub):
        maker_kwargs = function_stub._args[0]._args[0]
        containers = function_stub._args[1]
        fg = state(maker_kwargs["fgraph"])
        self.name = fg.get("name")
        self.inputs = list(fg["inputs"])
        self.outputs = list(fg["outputs"])
        self.order = _toposort(self.outputs)
        self.explicit, self.implicit = [], {}
        for pos, (var, box) in enumerate(zip(self.inputs, containers)):
            bst = state(box)
            if bst.get("required"):
                self.explicit.append((pos, bst.get("name")))
            else:
                self.implicit[pos] = np.array(bst["storage"][0], copy=True)
        self.op_census = {}
        for node in self.order:
            key = class_name(state(node)["op"])
            self.op_census[key] = self.op_census.get(key, 0) + 1

    def __call__(self, *args, shared=None):
        if len(args) != len(self.explicit):
            raise TypeError(f"expected {[n for _, n in self.explicit]}")
        env = {}
        for (pos, _), a in zip(self.explicit, args):
            env[id(self.inputs[pos])] = np.asarray(a, dtype=np.float64)
        for pos, v in self.implicit.items():
            env[id(self.inputs[pos])] = v
        for pos, v in (shared or {}).items():
            if pos not in self.implicit:
                raise KeyError(pos)
            v = np.asarray(v, dtype=np.float64)
            if v.shape != self.implicit[pos].shape:
                raise ValueError(f"shared input {pos}: shape {v.shape} != {self.implicit[pos].shape}")
            env[id(self.inputs[pos])] = v

        def value(v):
            if id(v) not in env:
                if _owner(v) is not None:
                    raise KeyError("unevaluated variable")
                env[id(v)] = _constant_value(v)
            return env[id(v)]

        for node in self.order:
            nst = state(node)
            res = _run_node(node, [value(i) for i in nst["inputs"]])
            outs = nst["outputs"]
            if len(outs) != 1:
                raise NotImplementedError("multi-output node")
            env[id(outs[0])] = res
        return [np.asarray(env[id(o)]) for o in self.outputs]


class ReferenceFunctions:
    """``train_func`` / ``pred_func`` of the shipped Boston checkpoint, plus what else it holds."""

    D, S, M = 13, 20, 50

    def __init__(self, path: str = BOSTON_PKL):
        top = load_stub_tree(path)
        self.top = top
        self._train = GraphFunction(top["train_func"])
        self._pred = GraphFunction(top["pred_func"])
        # shared-variable positions (see the input lists in the module docstring)
        self._train_params_pos = self._position_of_size(self._train, 1333)
        self._pred_params_pos = self._position_of_size(self._pred, 1333)
        self.trained_params = self._train.implicit[self._train_params_pos].copy()
        self.alpha = np.asarray(top["alpha"], dtype=np.float64)
        self.Li = np.asarray(top["Li"], dtype=np.float64)

    @staticmethod
    def _position_of_size(fn, size):
        pos = [p for p, v in fn.implicit.items() if v.shape == (size,)]
        assert len(pos) == 1, pos
        return pos[0]

    def train_func(self, X, y, params=None):
        """Reference ``train_func(X, y) -> [cost, alpha, Li]`` (SCFGP/SCFGP.py:132-135) at ``params``
        (default: the trained vector bound inside the pickled function)."""
        shared = None if params is None else {self._train_params_pos: params}
        cost, alpha, Li = self._train(X, np.asarray(y, dtype=np.float64).reshape(-1, 1), shared=shared)
        return float(cost), alpha, Li

    def pred_func(self, Xs, alpha, Li, params=None):
        """Reference ``pred_func(Xs, alpha, Li) -> [mu_pred (T,1), std_pred (T,)]`` (SCFGP.py:145-148)."""
        shared = None if params is None else {self._pred_params_pos: params}
        mu, std = self._pred(Xs, alpha, Li, shared=shared)
        return mu, std

    def cost_gradient_fd(self, X, y, params, h=1e-4, index=None):
        """Gradient of the reference's own ``cost`` by Richardson-extrapolated central differences
        (4th order: error O(h^4 f^(5)) + round-off ~1e-16/h).  This is what ``TT.grad(cost, params)``
        (SCFGP/SCFGP.py:129) differentiates; used to pin the hand adjoints to the reference graph."""
        from threadpoolctl import threadpool_limits

        params = np.asarray(params, dtype=np.float64)
        index = range(params.size) if index is None else index
        g = np.zeros(len(index))
        with threadpool_limits(1):  # 140 x 140 LAPACK calls: BLAS threads only add spin time
            for k, i in enumerate(index):
                direction = np.zeros_like(params)
                direction[i] = 1.0
                g[k] = self.cost_directional_fd(X, y, params, direction, h)
        return g

    def cost_directional_fd(self, X, y, params, direction, h=1e-4):
        """d/dt cost(params + t*direction) at t=0, same Richardson scheme as ``cost_gradient_fd``."""
        params = np.asarray(params, dtype=np.float64)
        f = lambda t: self.train_func(X, y, params + t * direction)[0]
        d1 = (f(h) - f(-h)) / (2 * h)
        d2 = (f(h / 2) - f(-h / 2)) / h
        return (4 * d2 - d1) / 3
