"""Names the reference imports from timm.layers at module scope. Only
trunc_normal_ is executed on the UniDepthV2/ViT path; the rest are placeholders
that raise if instantiated (ConvNeXt path is not covered by this shim)."""
from torch.nn.init import trunc_normal_  # noqa: F401
import torch.nn as nn


class _Unavailable(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("timm is not installed; ConvNeXt path needs a real timm")


class DropPath(nn.Identity):
    def __init__(self, *a, **k):
        super().__init__()


AvgPool2dSame = GlobalResponseNormMlp = LayerNorm = LayerNorm2d = Mlp = _Unavailable


def create_conv2d(*a, **k):
    raise NotImplementedError


def get_act_layer(*a, **k):
    raise NotImplementedError


def make_divisible(*a, **k):
    raise NotImplementedError


def to_ntuple(n):
    def f(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x,) * n
    return f
