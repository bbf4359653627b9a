"""Host-side update rules on one flat float64 parameter vector.

Restates the reference's ``Optimizer`` (SCFGP/Optimizer.py:12-382), which emits *symbolic*
Theano update dictionaries consumed inside ``train_iter_func`` (SCFGP/SCFGP.py:130-131,
136-137).  Here the gradient comes back from the CUDA evaluation and the update is plain
NumPy, so each rule is a small stateful object:

    rule = Optimizer.adam(learning_rate=0.01, beta1=0.9, beta2=0.999, epsilon=1e-8)
    rule = Optimizer.apply_nesterov_momentum(rule, momentum=0.9)
    params = rule.step(params, grad)

Reference behaviour worth knowing (kept, behind ``reference_quirk=True``, the default):
``apply_momentum``/``apply_nesterov_momentum`` pick ``list(updates.keys())[0]`` as "the
parameter" (Optimizer.py:48,89).  For ``sgd`` that is the parameter vector; for
``adam``/``adamax`` it is the first-moment slot ``m_prev`` (Optimizer.py:327,378) and for
``adagrad``/``adadelta`` the accumulator ``accu`` (Optimizer.py:162,270) -- so with the
default algorithm the wrapper perturbs Adam's first moment, not the parameters.
``rmsprop`` references an undefined name in the reference (Optimizer.py:210) and can never
have run; it is implemented here as the textbook rule its docstring describes.
"""
from __future__ import annotations

import numpy as np


class _Rule:
    """Base: ``slots`` is an ordered dict of state arrays; ``first`` names the slot the
    reference's momentum wrappers would grab (None -> the parameter vector)."""

    first = None

    def __init__(self):
        self.slots = {}
        self.wrappers = []  # [(kind, momentum, velocity)]

    def _slot(self, name, like):
        if name not in self.slots:
            self.slots[name] = np.zeros_like(like)
        return self.slots[name]

    def _update(self, params, grad):
        """returns (new_params, {slot: new_value})"""
        raise NotImplementedError

    def step(self, params, grad):
        params = np.asarray(params, dtype=np.float64)
        grad = np.asarray(grad, dtype=np.float64)
        new_p, new_slots = self._update(params, grad)
        for w in self.wrappers:
            kind, mom, quirk = w["kind"], w["momentum"], w["quirk"]
            target = self.first if quirk else None
            if target is None:
                old, new = params, new_p
            else:
                old, new = self._slot(target, params), new_slots[target]
            vel = w.setdefault("velocity", np.zeros_like(old))
            if kind == "momentum":          # Optimizer.py:27-60
                x = mom * vel + new
                w["velocity"] = x - old
                new = x
            else:                           # nesterov, Optimizer.py:62-97
                x = mom * vel + new - old
                w["velocity"] = x
                new = mom * x + new
            if target is None:
                new_p = new
            else:
                new_slots[target] = new
        for k, v in new_slots.items():
            self.slots[k] = v
        return new_p

    # state (de)serialisation for checkpoints
    def state_dict(self):
        d = {"slot_" + k: np.array(v) for k, v in self.slots.items()}
        for i, w in enumerate(self.wrappers):
            if "velocity" in w:
                d[f"vel_{i}"] = np.array(w["velocity"])
        return d

    def load_state_dict(self, d):
        for k, v in d.items():
            if k.startswith("slot_"):
                self.slots[k[5:]] = np.array(v)
            elif k.startswith("vel_"):
                self.wrappers[int(k[4:])]["velocity"] = np.array(v)


class _SGD(_Rule):
    def __init__(self, learning_rate=0.01, **_):
        super().__init__()
        self.lr = learning_rate

    def _update(self, p, g):  # Optimizer.py:99-119
        return p - self.lr * g, {}


class _Adagrad(_Rule):
    first = "accu"

    def __init__(self, learning_rate=0.01, epsilon=1e-6, **_):
        super().__init__()
        self.lr, self.eps = learning_rate, epsilon

    def _update(self, p, g):  # Optimizer.py:121-164
        accu = self._slot("accu", p) + g ** 2
        return p - self.lr * g / np.sqrt(accu + self.eps), {"accu": accu}


class _RMSProp(_Rule):
    first = None  # never ran in the reference; the momentum wrappers act on the parameters

    def __init__(self, learning_rate=0.01, rho=0.9, epsilon=1e-6, **_):
        super().__init__()
        self.lr, self.rho, self.eps = learning_rate, rho, epsilon

    def _update(self, p, g):  # Optimizer.py:166-213 (as documented; the source line 210 is broken)
        accu = self.rho * self._slot("accu", p) + (1 - self.rho) * g ** 2
        return p - self.lr * g / np.sqrt(accu + self.eps), {"accu": accu}


class _Adadelta(_Rule):
    first = "accu"

    def __init__(self, learning_rate=0.01, rho=0.95, epsilon=1e-6, **_):
        super().__init__()
        self.lr, self.rho, self.eps = learning_rate, rho, epsilon

    def _update(self, p, g):  # Optimizer.py:215-276
        accu = self.rho * self._slot("accu", p) + (1 - self.rho) * g ** 2
        dacc = self._slot("delta_accu", p)
        upd = g * np.sqrt(dacc + self.eps) / np.sqrt(accu + self.eps)
        return p - self.lr * upd, {"accu": accu, "delta_accu": self.rho * dacc + (1 - self.rho) * upd ** 2}


class _Adam(_Rule):
    first = "m"

    def __init__(self, learning_rate=0.01, beta1=0.9, beta2=0.99, epsilon=1e-8, **_):
        super().__init__()
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0.0

    def _update(self, p, g):  # Optimizer.py:278-331
        self.t += 1.0
        a_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        m = self.b1 * self._slot("m", p) + (1 - self.b1) * g
        v = self.b2 * self._slot("v", p) + (1 - self.b2) * g ** 2
        return p - a_t * m / (np.sqrt(v) + self.eps), {"m": m, "v": v}

    def state_dict(self):          # shared with _Adamax below, hence no zero-argument super()
        d = _Rule.state_dict(self); d["t"] = np.array(self.t); return d

    def load_state_dict(self, d):
        d = dict(d); self.t = float(d.pop("t", 0.0)); _Rule.load_state_dict(self, d)


class _Adamax(_Rule):
    first = "m"

    def __init__(self, learning_rate=0.01, beta1=0.9, beta2=0.999, epsilon=1e-8, **_):
        super().__init__()
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0.0

    def _update(self, p, g):  # Optimizer.py:333-382
        self.t += 1.0
        a_t = self.lr / (1 - self.b1 ** self.t)
        m = self.b1 * self._slot("m", p) + (1 - self.b1) * g
        u = np.maximum(self.b2 * self._slot("u", p), np.abs(g))
        return p - a_t * m / (u + self.eps), {"m": m, "u": u}

    state_dict = _Adam.state_dict
    load_state_dict = _Adam.load_state_dict


class Optimizer:
    """Same rule names as the reference's ``Optimizer.algos`` (Optimizer.py:14-25).
    ``norm_constraint`` / ``total_norm_constraint`` are listed there but not defined in the
    reference; they are not defined here either."""

    algos = ["apply_momentum", "apply_nesterov_momentum", "sgd", "adagrad", "rmsprop", "adadelta",
             "adam", "adamax", "norm_constraint", "total_norm_constraint"]

    sgd = staticmethod(lambda **kw: _SGD(**kw))
    adagrad = staticmethod(lambda **kw: _Adagrad(**kw))
    rmsprop = staticmethod(lambda **kw: _RMSProp(**kw))
    adadelta = staticmethod(lambda **kw: _Adadelta(**kw))
    adam = staticmethod(lambda **kw: _Adam(**kw))
    adamax = staticmethod(lambda **kw: _Adamax(**kw))

    @staticmethod
    def apply_momentum(rule, momentum=0.9, reference_quirk=True):
        rule.wrappers.append({"kind": "momentum", "momentum": momentum, "quirk": reference_quirk})
        return rule

    @staticmethod
    def apply_nesterov_momentum(rule, momentum=0.9, reference_quirk=True):
        rule.wrappers.append({"kind": "nesterov", "momentum": momentum, "quirk": reference_quirk})
        return rule

    @staticmethod
    def build(algo="adam", algo_params=None, nesterov=0.9, reference_quirk=True):
        """What ``build_theano_models`` assembles (SCFGP/SCFGP.py:130-131)."""
        rule = getattr(Optimizer, algo)(**(algo_params or {}))
        if nesterov is not None:
            Optimizer.apply_nesterov_momentum(rule, momentum=nesterov, reference_quirk=reference_quirk)
        return rule
