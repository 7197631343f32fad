"""`mlx.core` stand-in backed by torch CPU tensors — TEST INFRASTRUCTURE ONLY.

Purpose: let the UNMODIFIED reference sources under /root/reference/f5_tts_mlx/ (dit.py, cfm.py,
rope.py, convnext_v2.py, audio.py, duration.py, utils.py, generate.py) execute in this image, where
MLX itself cannot be installed, so that the CPU oracle (oracle/f5_oracle.py) can be pinned to what
the reference's own code computes (tests/golden/make_ref_golden.py, tests/test_ref_pins.py).

Only the subset of the MLX API those files touch is provided, with MLX's documented semantics:
  * `array`: immutable-value wrapper around a torch.Tensor; python-scalar operands keep the array's
    dtype class (float32 stays float32), `arange(int)` is int32, default float is float32;
  * `pad` takes numpy-style pad_width (int | (lo, hi) for every axis | [(lo, hi), ...]);
  * `as_strided` strides are in ELEMENTS; `split(x, int)` = equal sections, `split(x, [idx])` = indices;
  * `fast.scaled_dot_product_attention` = softmax(q·kᵀ·scale [+ mask]) · v with a boolean mask
    meaning "may attend" and softmax in float32;
  * `random.seed(s)` + `random.normal(shape)`: MLX's threefry stream is NOT reproduced — the draw is
    `torch.randn(shape, generator=torch.Generator().manual_seed(s))`, which is also what the product
    path draws for `seed=` (cfm.py of this repo), so seeded runs of the reference-through-the-shim and
    of the product see the same noise.
Deliberately NOT provided: `array.expand` (it does not exist in MLX; the reference's batched mask
branch at dit.py:162 calls it and therefore cannot run upstream either).  Setting
`mlx_shim.core.ALLOW_EXPAND = True` adds it with the obviously intended broadcast semantics so that a
batch > 1 fixture of the *intended* behaviour can be produced; fixtures made that way say so.
"""
from __future__ import annotations

import math
from typing import Any, Optional, Sequence

import numpy as np
import torch

ALLOW_EXPAND = False

pi = math.pi
float32 = torch.float32
float16 = torch.float16
bfloat16 = torch.bfloat16
int32 = torch.int32
int64 = torch.int64
uint32 = torch.int64      # only ever used for keys / indices in the reference
bool_ = torch.bool
complex64 = torch.complex64


def _t(x: Any) -> Any:
    """unwrap to a torch tensor / python scalar"""
    if isinstance(x, array):
        return x._t
    return x


def _as_tensor(x: Any, dtype=None) -> torch.Tensor:
    if isinstance(x, array):
        t = x._t
    elif isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        if t.dtype == torch.float64:
            t = t.float()            # MLX has no float64 on the default path: np.float64 -> float32
        elif t.dtype == torch.int64:
            t = t.to(torch.int32)
    elif isinstance(x, (list, tuple)):
        if len(x) > 0 and all(isinstance(e, (array, torch.Tensor)) for e in x):
            t = torch.stack([_as_tensor(e) for e in x])
        else:
            t = torch.tensor(x)
            if t.dtype == torch.float64:
                t = t.float()
            elif t.dtype == torch.int64:
                t = t.to(torch.int32)
    elif isinstance(x, bool):
        t = torch.tensor(x)
    elif isinstance(x, int):
        t = torch.tensor(x, dtype=torch.int32)
    elif isinstance(x, float):
        t = torch.tensor(x, dtype=torch.float32)
    else:
        raise TypeError(f"mlx_shim: cannot make an array from {type(x)}")
    if dtype is not None:
        t = t.to(dtype)
    return t


def _ints(shape) -> tuple:
    if isinstance(shape, (int, float, array)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


class _Namespace:
    """Array-API namespace handed to einops.array_api (reshape / permute_dims / expand_dims /
    broadcast_to / stack / concat + the built-in reductions)."""

    @staticmethod
    def reshape(x, shape):
        return array(x._t.reshape(tuple(shape)))

    @staticmethod
    def permute_dims(x, axes):
        return array(x._t.permute(tuple(axes)))

    @staticmethod
    def expand_dims(x, axis):
        return array(x._t.unsqueeze(axis))

    @staticmethod
    def broadcast_to(x, shape):
        return array(x._t.broadcast_to(tuple(shape)))

    @staticmethod
    def stack(xs, axis=0):
        return array(torch.stack([_t(x) for x in xs], dim=axis))

    @staticmethod
    def concat(xs, axis=0):
        return array(torch.cat([_t(x) for x in xs], dim=axis))

    @staticmethod
    def sum(x, axis):
        return array(x._t.sum(dim=tuple(axis)))

    @staticmethod
    def mean(x, axis):
        return array(x._t.mean(dim=tuple(axis)))

    @staticmethod
    def max(x, axis):
        return array(x._t.amax(dim=tuple(axis)))

    @staticmethod
    def min(x, axis):
        return array(x._t.amin(dim=tuple(axis)))

    @staticmethod
    def prod(x, axis):
        t = x._t
        for a in sorted(axis, reverse=True):
            t = t.prod(dim=a)
        return array(t)


def _index(idx):
    if isinstance(idx, tuple):
        return tuple(_index(i) for i in idx)
    if isinstance(idx, array):
        t = idx._t
        return t.long() if t.dtype in (torch.int32, torch.int16, torch.uint8, torch.int8) else t
    if isinstance(idx, slice):
        return slice(*(int(v) if isinstance(v, array) else v for v in (idx.start, idx.stop, idx.step)))
    return idx


class array:
    __slots__ = ("_t",)
    __array_priority__ = 1000

    def __init__(self, data, dtype=None):
        self._t = _as_tensor(data, dtype)

    def __class_getitem__(cls, item):       # `mx.array["b n d"]` annotations
        return cls

    # ---- introspection ----
    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def ndim(self):
        return self._t.ndim

    @property
    def dtype(self):
        return self._t.dtype

    @property
    def size(self):
        return self._t.numel()

    @property
    def T(self):
        return array(self._t.T)

    def __array_namespace__(self, api_version=None):
        return _Namespace

    def __len__(self):
        return self._t.shape[0]

    def __iter__(self):
        for i in range(self._t.shape[0]):
            yield array(self._t[i])

    def item(self):
        return self._t.item()

    def tolist(self):
        return self._t.tolist()

    def __bool__(self):
        return bool(self._t)

    def __int__(self):
        return int(self._t.item())

    def __float__(self):
        return float(self._t.item())

    def __index__(self):
        return int(self._t.item())

    def __array__(self, dtype=None, copy=None):
        a = self._t.detach().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return f"array({self._t!r})"

    # ---- indexing ----
    def __getitem__(self, idx):
        idx = _index(idx)
        # negative-step slices ([::-1]) are legal in MLX, not in torch
        if isinstance(idx, slice) and idx.step is not None and idx.step < 0:
            assert idx.step == -1 and idx.start is None and idx.stop is None
            return array(self._t.flip(0))
        return array(self._t[idx])

    # ---- arithmetic ----
    def _bin(self, other, fn, reverse=False):
        o = _t(other)
        a, b = (o, self._t) if reverse else (self._t, o)
        return array(fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, lambda a, b: torch.as_tensor(a) - b if not isinstance(a, torch.Tensor) else a - b, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.true_divide)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: a / b, True)
    def __floordiv__(self, o): return self._bin(o, lambda a, b: torch.div(a, b, rounding_mode="floor"))
    def __mod__(self, o): return self._bin(o, torch.remainder)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __rpow__(self, o): return self._bin(o, lambda a, b: torch.pow(torch.as_tensor(float(a)) if not isinstance(a, torch.Tensor) else a, b), True)
    def __matmul__(self, o): return self._bin(o, torch.matmul)
    def __neg__(self): return array(-self._t)
    def __and__(self, o): return self._bin(o, torch.logical_and if self._t.dtype == torch.bool else torch.bitwise_and)
    def __or__(self, o): return self._bin(o, torch.logical_or if self._t.dtype == torch.bool else torch.bitwise_or)
    def __invert__(self): return array(~self._t)
    def __lt__(self, o): return self._bin(o, torch.lt)
    def __le__(self, o): return self._bin(o, torch.le)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __ge__(self, o): return self._bin(o, torch.ge)
    def __eq__(self, o): return self._bin(o, torch.eq)      # noqa: E704
    def __ne__(self, o): return self._bin(o, torch.ne)
    __hash__ = None

    # ---- methods used by the reference ----
    def astype(self, dtype):
        return array(self._t.to(dtype))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return array(self._t.reshape(_ints(shape)))

    def transpose(self, *axes):
        if len(axes) == 0:
            return array(self._t.permute(*reversed(range(self._t.ndim))))
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        return array(self._t.permute(*axes))

    def swapaxes(self, a, b):
        return array(self._t.transpose(a, b))

    def moveaxis(self, src, dst):
        return array(torch.movedim(self._t, src, dst))

    def squeeze(self, axis=None):
        return array(self._t.squeeze() if axis is None else self._t.squeeze(axis))

    def flatten(self):
        return array(self._t.flatten())

    def sin(self): return array(torch.sin(self._t))
    def cos(self): return array(torch.cos(self._t))
    def exp(self): return array(torch.exp(self._t))
    def log(self): return array(torch.log(self._t))
    def sqrt(self): return array(torch.sqrt(self._t))
    def abs(self): return array(torch.abs(self._t))
    def square(self): return array(torch.square(self._t))

    def _reduce(self, fn, axis, keepdims):
        if axis is None:
            r = fn(self._t)
            return array(r.reshape((1,) * self._t.ndim) if keepdims else r)
        ax = tuple(axis) if isinstance(axis, (tuple, list)) else (axis,)
        return array(fn(self._t, dim=ax, keepdim=keepdims))

    def sum(self, axis=None, keepdims=False):
        return self._reduce(torch.sum, axis, keepdims)

    def mean(self, axis=None, keepdims=False):
        return self._reduce(torch.mean, axis, keepdims)

    def max(self, axis=None, keepdims=False):
        return self._reduce(torch.amax, axis, keepdims) if axis is not None else array(self._t.max())

    def min(self, axis=None, keepdims=False):
        return self._reduce(torch.amin, axis, keepdims) if axis is not None else array(self._t.min())

    def __getattr__(self, name):
        if name == "expand" and ALLOW_EXPAND:
            return lambda *shape: array(self._t.expand(*_ints(shape)))
        raise AttributeError(f"'mlx.core.array' object has no attribute '{name}'")


# ---------------------------------------------------------------------------------------------
# creation
# ---------------------------------------------------------------------------------------------
def arange(*args, dtype=None):
    vals = [_t(a) if not isinstance(a, array) else a.item() for a in args]
    is_float = any(isinstance(v, float) for v in vals)
    t = torch.arange(*vals, dtype=dtype if dtype is not None else (torch.float32 if is_float else torch.int32))
    return array(t)


def linspace(start, stop, num=50, dtype=float32):
    # MLX computes linspace in the output dtype as start + i * (stop - start) / (num - 1)
    start, stop = float(_t(start)), float(_t(stop))
    i = torch.arange(num, dtype=torch.float32)
    step = torch.tensor((stop - start) / (num - 1) if num > 1 else 0.0, dtype=torch.float32)
    return array((torch.tensor(start, dtype=torch.float32) + i * step).to(dtype))


def zeros(shape, dtype=float32): return array(torch.zeros(_ints(shape), dtype=dtype))
def ones(shape, dtype=float32): return array(torch.ones(_ints(shape), dtype=dtype))
def zeros_like(x): return array(torch.zeros_like(_as_tensor(x)))
def ones_like(x): return array(torch.ones_like(_as_tensor(x)))


def full(shape, vals, dtype=None):
    v = _t(vals)
    if dtype is None:
        dtype = v.dtype if isinstance(v, torch.Tensor) else (torch.float32 if isinstance(v, float) else
                                                              torch.bool if isinstance(v, bool) else torch.int32)
    return array(torch.full(_ints(shape), v if not isinstance(v, torch.Tensor) else v.item(), dtype=dtype))


# ---------------------------------------------------------------------------------------------
# elementwise / shape ops
# ---------------------------------------------------------------------------------------------
def _un(fn):
    return lambda x: array(fn(_as_tensor(x).float() if not _as_tensor(x).is_floating_point() and not _as_tensor(x).is_complex()
                              else _as_tensor(x)))


exp = _un(torch.exp)
log = _un(torch.log)
sin = _un(torch.sin)
cos = _un(torch.cos)
sqrt = _un(torch.sqrt)
tanh = _un(torch.tanh)
erf = _un(torch.erf)
sigmoid = _un(torch.sigmoid)
square = lambda x: array(torch.square(_as_tensor(x)))                      # noqa: E731
abs = lambda x: array(torch.abs(_as_tensor(x)))                            # noqa: E731,A001
rsqrt = _un(torch.rsqrt)


def _bin(fn):
    def f(a, b):
        ta, tb = _t(a), _t(b)
        if not isinstance(ta, torch.Tensor) and not isinstance(tb, torch.Tensor):
            ta = _as_tensor(ta)
        if not isinstance(ta, torch.Tensor):
            ta = torch.as_tensor(ta, dtype=tb.dtype if (isinstance(ta, int) or tb.is_floating_point()) else None)
        if not isinstance(tb, torch.Tensor):
            tb = torch.as_tensor(tb, dtype=ta.dtype if (isinstance(tb, int) or ta.is_floating_point()) else None)
        return array(fn(ta, tb))
    return f


maximum = _bin(torch.maximum)
minimum = _bin(torch.minimum)
logaddexp = _bin(torch.logaddexp)
add = _bin(torch.add)
multiply = _bin(torch.mul)
matmul = lambda a, b: array(torch.matmul(_t(a), _t(b)))                     # noqa: E731


def clip(x, a_min, a_max):
    return array(torch.clamp(_t(x), min=_t(a_min), max=_t(a_max)))


def where(cond, a, b):
    c = _t(cond)
    ta, tb = _t(a), _t(b)
    if not isinstance(c, torch.Tensor):
        c = torch.tensor(bool(c))
    if not isinstance(ta, torch.Tensor):
        ta = torch.as_tensor(ta, dtype=tb.dtype if isinstance(tb, torch.Tensor) else None)
    if not isinstance(tb, torch.Tensor):
        tb = torch.as_tensor(tb, dtype=ta.dtype)
    return array(torch.where(c.bool(), ta, tb))


def pad(x, pad_width, mode="constant", constant_values=0):
    t = _as_tensor(x)
    if isinstance(pad_width, (int, array)):
        pw = [(int(pad_width), int(pad_width))] * t.ndim
    elif len(pad_width) == 2 and all(isinstance(p, (int, float, array)) for p in pad_width):
        pw = [(int(pad_width[0]), int(pad_width[1]))] * t.ndim        # one (before, after) pair: every axis
    else:
        pw = [(int(a), int(b)) for a, b in pad_width]
        if len(pw) == 1:
            pw = pw * t.ndim
    assert len(pw) == t.ndim and mode == "constant"
    if any(lo < 0 or hi < 0 for lo, hi in pw):
        raise ValueError(f"[pad] Invalid (negative) padding size {pw}")     # MLX rejects negative widths; torch would crop
    flat = []
    for lo, hi in reversed(pw):
        flat += [lo, hi]
    val = _t(constant_values)
    if t.dtype == torch.bool:
        return array(torch.nn.functional.pad(t.to(torch.uint8), flat, value=int(bool(val))).bool())
    return array(torch.nn.functional.pad(t, flat, value=val))


def stack(xs, axis=0): return array(torch.stack([_as_tensor(x) for x in xs], dim=axis))
def concatenate(xs, axis=0): return array(torch.cat([_as_tensor(x) for x in xs], dim=axis))
def expand_dims(x, axis): return array(_as_tensor(x).unsqueeze(axis))
def squeeze(x, axis=None): return array(_as_tensor(x).squeeze() if axis is None else _as_tensor(x).squeeze(axis))
def swapaxes(x, a, b): return array(_as_tensor(x).transpose(a, b))
def transpose(x, axes=None): return array(x).transpose(*(axes or ()))
def reshape(x, shape): return array(_as_tensor(x).reshape(_ints(shape)))
def broadcast_to(x, shape): return array(_as_tensor(x).broadcast_to(_ints(shape)))


def split(x, indices_or_sections, axis=0):
    t = _as_tensor(x)
    if isinstance(indices_or_sections, int):
        assert t.shape[axis] % indices_or_sections == 0
        return [array(p) for p in torch.split(t, t.shape[axis] // indices_or_sections, dim=axis)]
    idx = [0] + [int(i) for i in indices_or_sections] + [t.shape[axis]]
    return [array(t.narrow(axis, idx[i], idx[i + 1] - idx[i])) for i in range(len(idx) - 1)]


def einsum(subscripts, *ops):
    return array(torch.einsum(subscripts.replace(" ", ""), *[_as_tensor(o) for o in ops]))


def outer(a, b):
    ta, tb = _as_tensor(a), _as_tensor(b)
    dt = torch.result_type(ta, tb)
    return array(torch.outer(ta.to(dt), tb.to(dt)))


def as_strided(x, shape=None, strides=None, offset=0):
    t = _as_tensor(x).contiguous()
    return array(torch.as_strided(t, _ints(shape), _ints(strides), offset).clone())


def sum(x, axis=None, keepdims=False): return array(x).sum(axis, keepdims)        # noqa: A001
def mean(x, axis=None, keepdims=False): return array(x).mean(axis, keepdims)
def max(x, axis=None, keepdims=False): return array(x).max(axis, keepdims)        # noqa: A001
def softmax(x, axis=-1): return array(torch.softmax(_as_tensor(x), dim=axis))


def eval(*args, **kwargs):                                                        # noqa: A001
    return None


def compile(fn, *a, **k):                                                         # noqa: A001
    return fn


def load(path, format=None):                                                      # noqa: A002
    from safetensors.torch import load_file
    return {k: array(v) for k, v in load_file(str(path)).items()}


class _FFT:
    @staticmethod
    def rfft(a, n=None, axis=-1):
        return array(torch.fft.rfft(_as_tensor(a), n=n, dim=axis))

    @staticmethod
    def irfft(a, n=None, axis=-1):
        return array(torch.fft.irfft(_as_tensor(a), n=n, dim=axis))


class _Linalg:
    @staticmethod
    def norm(x, ord=None, axis=None, keepdims=False):                              # noqa: A002
        t = _as_tensor(x)
        assert ord in (None, 2) and isinstance(axis, int)
        return array(torch.sqrt(torch.sum(t * t, dim=axis, keepdim=keepdims)))


class _Random:
    def __init__(self):
        self._gen = torch.Generator()
        self._gen.manual_seed(0)

    def seed(self, s):
        self._gen = torch.Generator()
        self._gen.manual_seed(int(s))

    def normal(self, shape=(), dtype=float32, loc=0.0, scale=1.0, key=None):
        return array(torch.randn(_ints(shape), generator=self._gen, dtype=dtype) * scale + loc)

    def uniform(self, low=0.0, high=1.0, shape=(), dtype=float32, key=None):
        u = torch.rand(_ints(shape), generator=self._gen, dtype=dtype)
        return array(u * (_t(high) - _t(low)) + _t(low))


class _Fast:
    @staticmethod
    def scaled_dot_product_attention(q, k, v, *, scale, mask=None):
        tq, tk, tv = _as_tensor(q), _as_tensor(k), _as_tensor(v)
        s = float(_t(scale)) if not isinstance(_t(scale), torch.Tensor) else float(_t(scale).item())
        scores = torch.matmul(tq * s, tk.transpose(-1, -2)).float()
        if mask is not None:
            m = _as_tensor(mask)
            scores = scores.masked_fill(~m, float("-inf")) if m.dtype == torch.bool else scores + m
        p = torch.softmax(scores, dim=-1).to(tv.dtype)
        return array(torch.matmul(p, tv))

    @staticmethod
    def layer_norm(x, weight, bias, eps):
        t = _as_tensor(x)
        mu = t.mean(-1, keepdim=True)
        var = ((t - mu) ** 2).mean(-1, keepdim=True)
        y = (t - mu) * torch.rsqrt(var + eps)
        if weight is not None:
            y = y * _as_tensor(weight)
        if bias is not None:
            y = y + _as_tensor(bias)
        return array(y)

    @staticmethod
    def rms_norm(x, weight, eps):
        t = _as_tensor(x)
        y = t * torch.rsqrt((t * t).mean(-1, keepdim=True) + eps)
        return array(y * _as_tensor(weight) if weight is not None else y)


fft = _FFT()
linalg = _Linalg()
random = _Random()
fast = _Fast()
