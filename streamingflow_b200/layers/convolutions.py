"""Convolutional building blocks with the reference's parameter names
(reference: streamingflow/layers/convolutions.py).  ``Bottleblock`` / channels-first ``LayerNorm`` belong to the
trusting gate of the dual-GRU cells and are evaluated by the CUDA engine (stages trunk / mix); ``Block`` and
``DeepLabHead`` belong to the post-ODE refinement: at the shipped width (64 channels, two SpatialGRU blocks, one ConvNeXt
block) FuturePredictionODE runs them on the CUDA engine too (refine_engine.py); their PyTorch ``forward``
below is only reached for other widths / block counts (e.g. the 128-channel modules of BASELINE config 5)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class LayerNorm(nn.Module):
    """LayerNorm over channels for channels_last or channels_first tensors (convolutions.py:283-308)."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps, self.data_format, self.normalized_shape = eps, data_format, (normalized_shape,)

    def forward(self, x):
        if self.data_format == "channels_last":
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        mu = x.mean(1, keepdim=True)
        var = (x - mu).pow(2).mean(1, keepdim=True)
        return self.weight[:, None, None] * ((x - mu) / torch.sqrt(var + self.eps)) + self.bias[:, None, None]


class Bottleblock(nn.Module):
    """7x7 -> LN -> GELU -> 1x1 -> LN -> GELU -> 3x3 -> LN -> GELU, plus (projected) skip (convolutions.py:348-380).
    Parameter container: the arithmetic is fused into the engine's trunk7 / trunk1 / mix stages."""

    def __init__(self, in_channels, out_channels=None):
        super().__init__()
        mid = int(in_channels / 2)
        out_channels = out_channels or in_channels
        cf = dict(eps=1e-6, data_format='channels_first')
        self.layers = nn.Sequential(
            nn.Conv2d(in_channels, mid, kernel_size=7, bias=False, padding=3), LayerNorm(mid, **cf), nn.GELU(),
            nn.Conv2d(mid, mid, kernel_size=1, bias=False), LayerNorm(mid, **cf), nn.GELU(),
            nn.Conv2d(mid, out_channels, kernel_size=3, bias=False, padding=1), LayerNorm(out_channels, **cf), nn.GELU())
        self.projection = None if out_channels == in_channels else nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False), nn.GELU())

    def forward(self, *args):
        raise RuntimeError("Bottleblock.forward is fused into the CUDA ODE engine (stages trunk7/trunk1/mix); "
                           "call the owning DualGRUODECell / DualGRUCell instead.")


class Block(nn.Module):
    """ConvNeXt block: depthwise 7x7, LN, 1x1 expand, GELU, 1x1 reduce, layer scale, residual (convolutions.py:310-346)."""

    def __init__(self, dim, drop_path=0., layer_scale_init_value=1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        if drop_path > 0.:
            raise NotImplementedError("stochastic depth is a training-time feature; the inference path uses drop_path = 0")
        self.drop_path = nn.Identity()

    def forward(self, x):
        y = self.dwconv(x).permute(0, 2, 3, 1)
        y = self.pwconv2(self.act(self.pwconv1(self.norm(y))))
        if self.gamma is not None:
            y = self.gamma * y
        return x + y.permute(0, 3, 1, 2)


class ASPPConv(nn.Sequential):
    def __init__(self, in_channels, out_channels, dilation):
        super().__init__(nn.Conv2d(in_channels, out_channels, 3, padding=dilation, dilation=dilation, bias=False),
                         nn.BatchNorm2d(out_channels), nn.ReLU())


class ASPPPooling(nn.Sequential):
    def __init__(self, in_channels, out_channels):
        super().__init__(nn.AdaptiveAvgPool2d(1), nn.Conv2d(in_channels, out_channels, 1, bias=False),
                         nn.BatchNorm2d(out_channels), nn.ReLU())

    def forward(self, x):
        hw = x.shape[-2:]
        return F.interpolate(super().forward(x), size=hw, mode='bilinear', align_corners=False)


class ASPP(nn.Module):
    """Atrous spatial pyramid pooling (convolutions.py:213-240)."""

    def __init__(self, in_channels, atrous_rates, out_channels=256):
        super().__init__()
        branches = [nn.Sequential(nn.Conv2d(in_channels, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels), nn.ReLU())]
        branches += [ASPPConv(in_channels, out_channels, r) for r in tuple(atrous_rates)]
        branches.append(ASPPPooling(in_channels, out_channels))
        self.convs = nn.ModuleList(branches)
        self.project = nn.Sequential(nn.Conv2d(len(self.convs) * out_channels, out_channels, 1, bias=False),
                                     nn.BatchNorm2d(out_channels), nn.ReLU(), nn.Dropout(0.5))

    def forward(self, x):
        return self.project(torch.cat([b(x) for b in self.convs], dim=1))


class DeepLabHead(nn.Sequential):
    """ASPP([12, 24, 36]) + 3x3 conv + BN + ReLU + 1x1 conv (convolutions.py:242-250)."""

    def __init__(self, in_channels, num_classes, hidden_channel=256):
        super().__init__(ASPP(in_channels, [12, 24, 36], hidden_channel),
                         nn.Conv2d(hidden_channel, hidden_channel, 3, padding=1, bias=False), nn.BatchNorm2d(hidden_channel),
                         nn.ReLU(), nn.Conv2d(hidden_channel, num_classes, 1))
