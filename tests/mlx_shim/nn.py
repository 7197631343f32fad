"""`mlx.nn` stand-in (see core.py) — the layers the reference's hot path instantiates, with MLX's
parameter names, layouts and defaults:

  Linear.weight (out, in) · Conv1d input NLC, weight (out, k, in/groups), zero padding ·
  LayerNorm(dims, eps=1e-5, affine=True) biased variance · RMSNorm(dims, eps=1e-5) ·
  Embedding.weight (n, d) · GELU(approx: "none" exact erf | "precise"/"tanh" tanh form | "fast") ·
  Mish = x·tanh(softplus(x)) · SiLU · Softplus = logaddexp(x, 0) · Sequential.layers · Dropout (eval: id).

Module mirrors the bits of mlx.nn.Module the reference uses: the parameter tree is every public
attribute that is an array / Module / list / dict (names starting with "_" are not parameters),
`load_weights([(dotted_name, array), ...])` is strict, `eval()` returns self.
"""
from __future__ import annotations

import math
from typing import Any, Callable, List, Tuple

import torch

from . import core as mx
from .core import array


class Module:
    def __init__(self):
        self.training = True

    def __call__(self, *a, **k):
        raise NotImplementedError

    # ---- parameter tree ----
    def _children(self):
        for k, v in vars(self).items():
            if k.startswith("_") or k == "training":
                continue
            yield k, v

    def parameters(self) -> dict:
        def walk(v):
            if isinstance(v, Module):
                return v.parameters()
            if isinstance(v, array):
                return v
            if isinstance(v, (list, tuple)):
                out = [walk(e) for e in v]
                return out if any(o is not None for o in out) else None
            if isinstance(v, dict):
                out = {k: walk(e) for k, e in v.items()}
                return {k: o for k, o in out.items() if o is not None} or None
            return None

        res = {}
        for k, v in self._children():
            w = walk(v)
            if isinstance(w, array) or (w is not None and len(w) > 0):
                res[k] = w
        return res

    def flat_parameters(self) -> List[Tuple[str, array]]:
        out = []

        def rec(prefix, node):
            if isinstance(node, array):
                out.append((prefix, node))
            elif isinstance(node, dict):
                for k, v in node.items():
                    rec(f"{prefix}.{k}" if prefix else k, v)
            elif isinstance(node, list):
                for i, v in enumerate(node):
                    if v is not None:
                        rec(f"{prefix}.{i}", v)

        rec("", self.parameters())
        return out

    def load_weights(self, weights, strict: bool = True):
        if isinstance(weights, dict):
            weights = list(weights.items())
        have = dict(self.flat_parameters())
        given = dict(weights)
        if strict:
            extra = sorted(set(given) - set(have))
            missing = sorted(set(have) - set(given))
            if extra or missing:
                raise ValueError(f"load_weights: unexpected {extra[:5]} missing {missing[:5]}")
        for name, val in given.items():
            if name not in have:
                continue
            val = val if isinstance(val, array) else array(val)
            if tuple(val.shape) != tuple(have[name].shape):
                raise ValueError(f"load_weights: shape mismatch for {name}: {val.shape} vs {have[name].shape}")
            node: Any = self
            parts = name.split(".")
            for p in parts[:-1]:
                node = node[int(p)] if isinstance(node, (list, tuple)) else (node[p] if isinstance(node, dict) else getattr(node, p))
            if isinstance(node, list):
                node[int(parts[-1])] = val
            elif isinstance(node, dict):
                node[parts[-1]] = val
            else:
                setattr(node, parts[-1], val)
        return self

    def update(self, params):
        self.load_weights(_flatten(params), strict=False)
        return self

    def eval(self):
        def rec(v):
            if isinstance(v, Module):
                v.training = False
                for _, c in vars(v).items():
                    rec(c)
            elif isinstance(v, (list, tuple)):
                for e in v:
                    rec(e)
            elif isinstance(v, dict):
                for e in v.values():
                    rec(e)
        rec(self)
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    def freeze(self, *a, **k):
        return self

    def children(self):
        return {k: v for k, v in self._children() if isinstance(v, (Module, list, dict))}


def _flatten(tree, prefix=""):
    out = []
    if isinstance(tree, array):
        return [(prefix, tree)]
    if isinstance(tree, dict):
        for k, v in tree.items():
            out += _flatten(v, f"{prefix}.{k}" if prefix else k)
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            out += _flatten(v, f"{prefix}.{i}" if prefix else str(i))
    return out


def _uniform(shape, scale):
    return array((torch.rand(shape) * 2 - 1) * scale)


class Linear(Module):
    def __init__(self, input_dims: int, output_dims: int, bias: bool = True):
        super().__init__()
        s = math.sqrt(1.0 / input_dims)
        self.weight = _uniform((output_dims, input_dims), s)
        if bias:
            self.bias = _uniform((output_dims,), s)

    def __call__(self, x):
        y = torch.matmul(x._t, self.weight._t.T)
        if "bias" in vars(self):
            y = y + self.bias._t
        return array(y)


class Conv1d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        s = math.sqrt(1.0 / (in_channels * kernel_size))
        self.weight = _uniform((out_channels, kernel_size, in_channels // groups), s)
        if bias:
            self.bias = array(torch.zeros(out_channels))
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups

    def __call__(self, x):
        # NLC in, (out, k, in/groups) weights -> torch's NCL / (out, in/groups, k)
        y = torch.nn.functional.conv1d(x._t.transpose(1, 2), self.weight._t.permute(0, 2, 1),
                                       self.bias._t if "bias" in vars(self) else None, stride=self.stride,
                                       padding=self.padding, dilation=self.dilation, groups=self.groups)
        return array(y.transpose(1, 2))


class LayerNorm(Module):
    def __init__(self, dims: int, eps: float = 1e-5, affine: bool = True, bias: bool = True):
        super().__init__()
        if affine:
            self.weight = array(torch.ones(dims))
            if bias:
                self.bias = array(torch.zeros(dims))
        self.eps, self.dims = eps, dims

    def __call__(self, x):
        return mx.fast.layer_norm(x, vars(self).get("weight"), vars(self).get("bias"), self.eps)


class RMSNorm(Module):
    def __init__(self, dims: int, eps: float = 1e-5):
        super().__init__()
        self.weight = array(torch.ones(dims))
        self.eps = eps

    def __call__(self, x):
        return mx.fast.rms_norm(x, self.weight, self.eps)


class Embedding(Module):
    def __init__(self, num_embeddings: int, dims: int):
        super().__init__()
        self.weight = array(torch.randn(num_embeddings, dims) * math.sqrt(1.0 / dims))

    def __call__(self, x):
        return array(self.weight._t[x._t.long()])


class GELU(Module):
    def __init__(self, approx: str = "none"):
        super().__init__()
        if approx not in ("none", "precise", "tanh", "fast"):
            raise ValueError(f"approx must be none | precise | tanh | fast, got {approx}")
        self._approx = approx

    def __call__(self, x):
        t = x._t
        if self._approx == "none":
            return array(t * (1 + torch.erf(t / math.sqrt(2))) / 2)
        if self._approx in ("precise", "tanh"):
            return array(0.5 * t * (1 + torch.tanh(math.sqrt(2 / math.pi) * (t + 0.044715 * t ** 3))))
        return array(t * torch.sigmoid(1.702 * t))


class Mish(Module):
    def __call__(self, x):
        return array(x._t * torch.tanh(torch.logaddexp(x._t, torch.zeros((), dtype=x._t.dtype))))


class SiLU(Module):
    def __call__(self, x):
        return array(x._t * torch.sigmoid(x._t))


class Softplus(Module):
    def __call__(self, x):
        return array(torch.logaddexp(x._t, torch.zeros((), dtype=x._t.dtype)))


class Dropout(Module):
    def __init__(self, p: float = 0.5):
        super().__init__()
        self._p = p

    def __call__(self, x):
        if self._p == 0 or not self.training:
            return x
        raise NotImplementedError("mlx_shim: training-mode dropout")


class Sequential(Module):
    def __init__(self, *modules):
        super().__init__()
        self.layers = list(modules)

    def __call__(self, x):
        for m in self.layers:
            x = m(x)
        return x


class _Losses:
    @staticmethod
    def mse_loss(p, t, reduction="mean"):
        d = (p._t - t._t) ** 2
        return array(d if reduction == "none" else d.mean())

    @staticmethod
    def l1_loss(p, t, reduction="mean"):
        d = (p._t - t._t).abs()
        return array(d if reduction == "none" else d.mean())


losses = _Losses()


def quantize(model, group_size=64, bits=4, class_predicate: Callable = None):
    raise NotImplementedError("mlx_shim: nn.quantize (the 4/8-bit checkpoints are handled by the product's own "
                              "dequantiser, f5_tts_mlx_b200.weights.dequantize_mlx_affine)")
