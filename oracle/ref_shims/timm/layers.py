"""Stand-ins for the `timm.layers` names the reference imports (timm is not installed here, SURVEY.md
section 8c / Appendix A).  TEST INFRASTRUCTURE, used only by oracle/make_golden.py to run the reference's
own ConvNeXt code (unidepth/models/backbones/convnext.py) when generating golden vectors.

Each class is written from its documented behaviour: LayerNorm = nn.LayerNorm with eps 1e-6; LayerNorm2d =
LayerNorm over the channel axis of an NCHW tensor; Mlp = fc1 -> act -> fc2 (modules named fc1 / act / fc2,
1x1 convolutions when use_conv); create_conv2d = nn.Conv2d with 'same'-style static padding and
groups = channels when depthwise; DropPath = identity at inference."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.init import trunc_normal_  # noqa: F401


class DropPath(nn.Identity):
    def __init__(self, *a, **k):
        super().__init__()


class LayerNorm(nn.LayerNorm):
    def __init__(self, num_channels, eps=1e-6, affine=True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)


class LayerNorm2d(nn.LayerNorm):
    def __init__(self, num_channels, eps=1e-6, affine=True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)

    def forward(self, x):
        y = F.layer_norm(x.permute(0, 2, 3, 1), self.normalized_shape, self.weight, self.bias, self.eps)
        return y.permute(0, 3, 1, 2)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                 bias=True, drop=0.0, use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        make = (lambda i, o: nn.Conv2d(i, o, kernel_size=1, bias=bias)) if use_conv else (lambda i, o: nn.Linear(i, o, bias=bias))
        self.fc1 = make(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = make(hidden_features, out_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Unavailable(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not needed by the configurations the oracle runs")


AvgPool2dSame = GlobalResponseNormMlp = _Unavailable


def create_conv2d(in_channels, out_channels, kernel_size, stride=1, dilation=1, depthwise=False, bias=True, padding="", **_):
    if padding in ("", "same", None):
        padding = ((stride - 1) + dilation * (kernel_size - 1)) // 2
    groups = out_channels if depthwise else 1
    return nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                     groups=groups, bias=bias)


def get_act_layer(name="gelu"):
    if callable(name) and not isinstance(name, str):
        return name
    return {"gelu": nn.GELU, "relu": nn.ReLU, "silu": nn.SiLU}[name]


def make_divisible(v, divisor=8, min_value=None, round_limit=0.9):
    min_value = min_value or divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


def to_ntuple(n):
    def f(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x,) * n
    return f
