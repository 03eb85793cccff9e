"""[third-party, restated] the part of `flax.linen` that equiv_flows.py:36-127 and cn_flows.py:117-182 use.

Reproduces Flax's *structure* rules, which fix the parameter tree (and therefore the flat-theta layout of the C ABI):
  * Module subclasses are dataclasses over their annotated fields (+ `name`, `parent`);
  * `setup()` runs lazily; a Module (or list of Modules) assigned to `self.<attr>` there is named `<attr>` (`<attr>_<i>`);
  * a Module constructed inside an `@nn.compact` method is auto-named `<ClassName>_<n>`, n counted per class and per call;
  * `nn.Dense(features)`: params `kernel` (in, out) LeCun-normal, `bias` (out,) zeros; y = x @ kernel + bias;
  * `nn.vmap(Cls, variable_axes={'params': None}, split_rngs={'params': False}, in_axes=...)` -> class `Vmap<Cls>`
    whose call maps the inputs and SHARES one parameter set;
  * `module.init(key, *args)` -> {'params': tree}; `module.apply(variables, *args)`.
Flax itself is not installable here, so the tree (params/cnf/net/VmapNN_0/{layers_i,last_layer}/{kernel,bias}) is what
SURVEY.md 8b expects Flax to produce, not something verified against Flax.
"""
from __future__ import annotations

import torch
from jax import random as _jrandom
from jax.nn import initializers as _init

_CTX = []          # stack of {'vars': root param dict, 'init': bool, 'key': Key, 'n': int}
_FRAMES = []       # stack of (module, mode, counters) -- mode in {'compact', 'setup', 'plain'}


def compact(fn):
    fn._nn_compact = True
    return fn


def tanh(x): return torch.tanh(x)
def sigmoid(x): return torch.sigmoid(x)
def relu(x): return torch.relu(x)
def softplus(x): return torch.nn.functional.softplus(x)


def _wrap_method(fn):
    is_compact = getattr(fn, "_nn_compact", False)

    def wrapped(self, *args, **kw):
        # Flax re-binds (clones) a module on every apply, so setup() runs again each time; re-running it here also keeps
        # tensors derived in setup() (e.g. RadialMLP.nuclei) from outliving the torch.func transform level they were
        # created under
        self._ensure_setup(force=True)
        _FRAMES.append((self, "compact" if is_compact else "plain", {}))
        try:
            return fn(self, *args, **kw)
        finally:
            _FRAMES.pop()
    wrapped._nn_wrapped = True
    wrapped.__name__ = getattr(fn, "__name__", "method")
    return wrapped


class Module:
    _fields = ()

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        fields = []
        for klass in reversed(cls.__mro__):
            if klass in (object, Module):
                continue
            for k in klass.__dict__.get("__annotations__", {}):
                if not k.startswith("_") and k not in fields:
                    fields.append(k)
        cls._fields = tuple(fields)
        call = cls.__dict__.get("__call__")
        if call is not None and not getattr(call, "_nn_wrapped", False):
            cls.__call__ = _wrap_method(call)

    def __init__(self, *args, name=None, parent=None, **kw):
        cls = type(self)
        if len(args) > len(cls._fields):
            raise TypeError(f"{cls.__name__} takes {len(cls._fields)} positional fields")
        vals = dict(zip(cls._fields, args))
        for k, v in kw.items():
            if k not in cls._fields or k in vals:
                raise TypeError(f"{cls.__name__}: unexpected or duplicate field {k!r}")
            vals[k] = v
        for k in cls._fields:
            if k not in vals:
                if not hasattr(cls, k):
                    raise TypeError(f"{cls.__name__}: missing field {k!r}")
                vals[k] = getattr(cls, k)
        d = object.__getattribute__(self, "__dict__")
        d.update(vals)
        d["_setup_done"] = False
        d["name"] = name
        d["parent"] = parent
        # constructed inside a compact method: adopted by that module under an automatic name
        if parent is None and _FRAMES and _FRAMES[-1][1] == "compact":
            owner, _, counters = _FRAMES[-1]
            d["parent"] = owner
            if name is None:
                n = counters.get(cls.__name__, 0)
                counters[cls.__name__] = n + 1
                d["name"] = f"{cls.__name__}_{n}"

    def __setattr__(self, key, value):
        # assignment in setup() names the submodule(s) after the attribute
        if _FRAMES and _FRAMES[-1][0] is self and _FRAMES[-1][1] == "setup":
            if isinstance(value, Module):
                value._adopt(self, key)
            elif isinstance(value, (list, tuple)) and value and all(isinstance(v, Module) for v in value):
                for i, v in enumerate(value):
                    v._adopt(self, f"{key}_{i}")
        object.__getattribute__(self, "__dict__")[key] = value

    def _adopt(self, parent, name):
        d = object.__getattribute__(self, "__dict__")
        if d["parent"] is None:
            d["parent"] = parent
            d["name"] = name

    def _ensure_setup(self, force=False):
        d = object.__getattribute__(self, "__dict__")
        if force or not d["_setup_done"]:
            d["_setup_done"] = True
            _FRAMES.append((self, "setup", {}))
            try:
                self.setup()
            finally:
                _FRAMES.pop()

    def setup(self):
        pass

    def __getattr__(self, key):
        # attributes defined by setup() are reachable before any method ran (Flax runs setup on first access)
        d = object.__getattribute__(self, "__dict__")
        if not key.startswith("_") and not d.get("_setup_done", True):
            self._ensure_setup()
            if key in d:
                return d[key]
        raise AttributeError(f"{type(self).__name__!r} object has no attribute {key!r}")

    # ---- variables ------------------------------------------------------------------------------------------------
    def _path(self):
        names, m = [], self
        root = _CTX[-1]["root"]
        while m is not root:
            if m.parent is None:
                raise RuntimeError(f"{type(m).__name__} is not attached to the module being applied")
            names.append(m.name)
            m = m.parent
        return names[::-1]

    def param(self, name, init_fn, *shape_args):
        ctx = _CTX[-1]
        scope = ctx["vars"]
        for n in self._path():
            if ctx["init"]:
                scope = scope.setdefault(n, {})
            else:
                scope = scope[n]
        if name not in scope:
            if not ctx["init"]:
                raise KeyError(f"missing parameter {'/'.join(self._path() + [name])}")
            ctx["n"] += 1
            scope[name] = init_fn(_jrandom.Key(ctx["key"].seed + (ctx["n"],)), *shape_args)
        return scope[name]

    def init(self, key, *args, **kw):
        ctx = {"vars": {}, "init": True, "key": key, "n": 0, "root": self}
        _CTX.append(ctx)
        try:
            self(*args, **kw)
        finally:
            _CTX.pop()
        return {"params": ctx["vars"]}

    def apply(self, variables, *args, **kw):
        _CTX.append({"vars": variables["params"], "init": False, "key": None, "n": 0, "root": self})
        try:
            return self(*args, **kw)
        finally:
            _CTX.pop()


class Dense(Module):
    features: int
    use_bias: bool = True
    kernel_init: object = None
    bias_init: object = None

    @compact
    def __call__(self, inputs):
        k_init = self.kernel_init or _init.lecun_normal()
        b_init = self.bias_init or _init.zeros
        kernel = self.param("kernel", k_init, (inputs.shape[-1], self.features))
        y = inputs @ kernel
        if self.use_bias:
            y = y + self.param("bias", b_init, (self.features,))
        return y


def vmap(target, variable_axes=None, split_rngs=None, in_axes=0, out_axes=0, **_):
    """Lifted vmap over a Module class with shared parameters (variable_axes={'params': None})."""
    if variable_axes and any(v is not None for v in variable_axes.values()):
        raise NotImplementedError("only shared-parameter nn.vmap is restated")
    base_call = target.__dict__["__call__"]

    def __call__(self, *args):
        if _CTX[-1]["init"]:
            # create the (shared) parameters once on the first mapped slice
            sl = [a if ax is None else a.select(ax, 0) for a, ax in zip(args, _axes(in_axes, len(args)))]
            base_call(self, *sl)
        return torch.func.vmap(lambda *a: base_call(self, *a), in_dims=in_axes, out_dims=out_axes)(*args)

    cls = type(f"Vmap{target.__name__}", (target,), {"__call__": __call__, "__annotations__": {}})
    return cls


def _axes(in_axes, n):
    return tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * n
