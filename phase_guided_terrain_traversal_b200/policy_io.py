"""Read / write policies in the layout the reference saves and deploys (training/train.py:195,266 writes
`model.save_params(...)` pickles; deploy/policy_net.py:6-33 reads them): a pickled tuple
`(RunningStatisticsState{mean, std, count, summed_variance}, PPONetworkParams{policy, value})` where
`policy["params"]["hidden_i"]["kernel" | "bias"]` are `[in, out]` / `[out]` arrays.

The shipped pickles reference jax / flax / brax classes that are not installable here; `load_policy` uses a
stand-in Unpickler that maps any unknown class to a plain attribute bag and jax arrays to numpy.
"""
from __future__ import annotations

import io
import pickle
from pathlib import Path

import numpy as np


class _Bag(dict):
    """Stands in for flax FrozenDict / brax dataclasses: item + attribute access."""

    def __init__(self, *a, **k):
        super().__init__()
        if a and isinstance(a[0], dict):
            self.update(a[0])
        self.update(k)

    def __getattr__(self, n):
        try:
            return self[n]
        except KeyError:
            raise AttributeError(n) from None

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            self.update(state[1] or {})
            if isinstance(state[0], dict):
                self.update(state[0])

    def __reduce__(self):
        return (_Bag, (dict(self),))


def _reconstruct_array(fun, args, arr_state=None, aval_state=None):
    """jax._src.array._reconstruct_array: rebuild the numpy value, drop the jax wrapper."""
    v = fun(*args)
    if arr_state is not None:
        v.__setstate__(arr_state)
    return np.asarray(v)


# The only globals a policy pickle may resolve: numpy array reconstruction and plain containers. Everything else (jax /
# flax / brax classes, but also builtins.eval, os.system, ...) becomes an inert attribute bag, so a crafted pickle cannot
# run code through this loader.
_SAFE_GLOBALS = {
    ("builtins", "tuple"), ("builtins", "list"), ("builtins", "dict"), ("builtins", "set"), ("builtins", "frozenset"),
    ("builtins", "slice"), ("builtins", "complex"), ("builtins", "bytearray"), ("builtins", "int"), ("builtins", "float"),
    ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"),
    ("collections", "OrderedDict"), ("_codecs", "encode"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy", "float32"), ("numpy", "float64"), ("numpy", "int32"), ("numpy", "int64"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer"),
}


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _SAFE_GLOBALS:
            return super().find_class(module, name)
        if name == "_reconstruct_array":
            return _reconstruct_array
        return type(name, (_Bag,), {"__module__": "pgtt_stub"})


def _layers(tree):
    p = tree["params"] if "params" in tree else tree
    names = sorted(p.keys(), key=lambda s: int(str(s).split("_")[-1]))
    return [np.asarray(p[n]["kernel"], np.float32) for n in names], [np.asarray(p[n]["bias"], np.float32) for n in names]


def load_policy(path):
    """-> dict(mean, std, count, policy=(kernels, biases), value=(kernels, biases) or None)."""
    with open(path, "rb") as f:
        params = _StubUnpickler(io.BytesIO(f.read())).load()
    norm, net = params[0], params[1]
    out = {"mean": np.asarray(norm["mean"]["state"], np.float32), "std": np.asarray(norm["std"]["state"], np.float32),
           "count": float(np.asarray(norm["count"]).reshape(-1)[0]) if "count" in norm and np.asarray(norm["count"]).size else None, "value": None}
    if len(params) == 3:                      # (normalizer, policy_params, value_params)
        out["policy"] = _layers(net)
        out["value"] = _layers(params[2])
    else:                                     # (normalizer, PPONetworkParams(policy, value))
        out["policy"] = _layers(net["policy"])
        if net.get("value") is not None:
            out["value"] = _layers(net["value"])
    if "privileged_state" in norm["mean"]:    # both layouts carry the value network's input statistics
        out["value_mean"] = np.asarray(norm["mean"]["privileged_state"], np.float32)
        out["value_std"] = np.asarray(norm["std"]["privileged_state"], np.float32)
    return out


def save_policy(path, mean, std, policy, value=None, count=0.0, value_mean=None, value_std=None):
    """Writes a pickle `deploy/policy_net.py:get_params` can read (plain dict / numpy containers). With `value_mean` / `value_std` the
    normaliser also carries the `privileged_state` statistics, like the reference's checkpoints (needed to resume training)."""
    def tree(kb):
        ks, bs = kb
        return {"params": {f"hidden_{i}": {"kernel": np.asarray(k, np.float32), "bias": np.asarray(b, np.float32)} for i, (k, b) in enumerate(zip(ks, bs))}}
    m, s = _Bag(state=np.asarray(mean, np.float32)), _Bag(state=np.asarray(std, np.float32))
    if value_mean is not None and value_std is not None:
        m["privileged_state"], s["privileged_state"] = np.asarray(value_mean, np.float32), np.asarray(value_std, np.float32)
    norm = _Bag(mean=m, std=s, count=np.float64(count))
    net = _Bag(policy=tree(policy), value=tree(value) if value is not None else None)
    Path(path).write_bytes(pickle.dumps((norm, net)))
