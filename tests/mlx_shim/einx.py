"""`einx` stand-in (see core.py): the four calls the reference makes — less / greater_equal /
where / divide with "in, in -> out" patterns of single-letter axes (utils.py:47,58,86,90,
duration.py:238).  Each operand is transposed/unsqueezed into the output axis order and the
elementwise op broadcasts."""
from __future__ import annotations

import torch

from .core import array, _t


def _align(pattern: str, *ops):
    ins, out = pattern.split("->")
    in_axes = [p.split() for p in ins.split(",")]
    out_axes = out.split()
    assert len(in_axes) == len(ops), (pattern, len(ops))
    res = []
    for axes, op in zip(in_axes, ops):
        t = _t(op)
        if not isinstance(t, torch.Tensor):
            assert not axes
            res.append(t)
            continue
        assert t.ndim == len(axes), (pattern, t.shape)
        order = [axes.index(a) for a in out_axes if a in axes]
        t = t.permute(order) if order else t
        idx = tuple(slice(None) if a in axes else None for a in out_axes)
        res.append(t[idx])
    return res


def less(pattern, a, b):
    x, y = _align(pattern, a, b)
    return array(x < y)


def greater_equal(pattern, a, b):
    x, y = _align(pattern, a, b)
    return array(x >= y)


def divide(pattern, a, b):
    x, y = _align(pattern, a, b)
    return array(x / y)


def where(pattern, c, a, b):
    cc, x, y = _align(pattern, c, a, b)
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x, dtype=y.dtype)
    if not isinstance(y, torch.Tensor):
        y = torch.as_tensor(y, dtype=x.dtype)
    return array(torch.where(cc, x, y))
