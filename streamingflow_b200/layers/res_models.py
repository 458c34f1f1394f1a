"""Encoder / decoder / prior-network containers with the reference's parameter names
(reference: streamingflow/layers/res_models.py).

``SmallEncoder`` / ``SmallDecoder`` sit inside the module boundary but outside the ODE loop; at the shipped width (64
channels, SKIPCO off) NNFOwithBayesianJumps runs them on the CUDA engine (codec_engine.py, "next" row 1 of SURVEY.md 8f) and
their PyTorch ``forward`` is only reached for other widths / SKIPCO.  ``ConvNet`` (= p_model) and ``SELayer``
are on the hot path: they only HOLD the parameters (state_dict contract); their arithmetic runs in
the CUDA engine (engine.prior_stage_defs), so calling their ``forward`` directly is an error by design -- there is
no PyTorch fallback for the ODE path.
"""
import torch
import torch.nn as nn

_ACTS = {"relu": nn.ReLU, "lrelu": lambda: nn.LeakyReLU(0.1), "tanh": nn.Tanh, "none": None}


class ConvBlock(nn.Module):
    """3x3 (transposed) convolution, optional BatchNorm / InstanceNorm, optional activation (res_models.py:8-49)."""

    def __init__(self, in_channels, out_channels=None, kernel_size=3, stride=1, norm='bn', activation='lrelu', bias=False,
                 transpose=False):
        super().__init__()
        out_channels = out_channels or in_channels
        conv_cls = nn.ConvTranspose2d if transpose else nn.Conv2d
        self.conv = conv_cls(in_channels, out_channels, kernel_size, stride, padding=(kernel_size - 1) // 2, bias=bias)
        if norm not in ("bn", "in", "none"):
            raise ValueError('Invalid norm {}'.format(norm))
        self.norm = {"bn": nn.BatchNorm2d, "in": nn.InstanceNorm2d}[norm](out_channels) if norm != "none" else None
        if activation not in _ACTS:
            raise ValueError('Invalid activation {}'.format(activation))
        self.activation = _ACTS[activation]() if _ACTS[activation] is not None else None

    def forward(self, x):
        x = self.conv(x)
        if self.norm is not None:
            x = self.norm(x)
        return x if self.activation is None else self.activation(x)


class ResBlock(nn.Module):
    """Two ConvBlocks plus identity / 1x1-projected skip (res_models.py:52-79)."""

    def __init__(self, in_channels, out_channels=None, norm='bn', activation='lrelu', bias=False):
        super().__init__()
        out_channels = out_channels or in_channels
        body = nn.Sequential()
        body.add_module('conv_1', ConvBlock(in_channels, in_channels, 3, 1, norm, activation, bias))
        body.add_module('conv_2', ConvBlock(in_channels, out_channels, 3, 1, norm, activation, bias))
        body.add_module('dropout', nn.Dropout2d(0.25))
        self.layers = body
        self.projection = nn.Conv2d(in_channels, out_channels, 1) if out_channels != in_channels else None

    def forward(self, x):
        skip = x if self.projection is None else self.projection(x)
        return skip + self.layers(x)


class SmallEncoder(nn.Module):
    """BEV frame -> /4 latent, tanh head (res_models.py:82-109)."""

    def __init__(self, nc, nh, nf):
        super().__init__()
        widths = [(nc, nf), (nf, 2 * nf), (2 * nf, 2 * nf), (2 * nf, 2 * nf), (2 * nf, 4 * nf)]
        self.blocks = nn.ModuleList([ResBlock(a, b) for a, b in widths])
        self.last_conv = nn.Sequential(ConvBlock(4 * nf, nh, 3, stride=1, activation='tanh'))
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2, padding=0)

    def forward(self, x, return_skip=False):
        feats = []
        for i, blk in enumerate(self.blocks):
            if i == 1 or i == 2:
                x = self.maxpool(x)
            x = blk(x)
            feats.append(x)
        x = self.last_conv(x)
        return (x, feats[::-1]) if return_skip else x


class SmallDecoder(nn.Module):
    """Latent -> x4 BEV frame (res_models.py:112-147)."""

    def __init__(self, nc, nh, nf, skip):
        super().__init__()
        k = 2 if skip else 1
        self.skip = skip
        self.first_upconv = ConvBlock(nc, 4 * nf, stride=1, transpose=True)
        widths = [(4 * nf * k, 2 * nf), (2 * nf * k, 2 * nf), (2 * nf * k, 2 * nf), (2 * nf * k, nf), (nf * k, nf)]
        self.blocks = nn.ModuleList([ResBlock(a, b) for a, b in widths])
        self.last_conv = nn.Sequential(ConvBlock(nf * k, nf, 3, stride=1),
                                       ConvBlock(nf, nh, 3, stride=1, transpose=True, bias=True, norm='none'))
        self.upsample = nn.Upsample(scale_factor=2, mode='nearest')

    def forward(self, z, skip=None, sigmoid=False):
        assert skip is None and not self.skip or self.skip and skip is not None
        h = self.first_upconv(z)
        for i, blk in enumerate(self.blocks):
            if skip is not None:
                h = torch.cat([h, skip[i]], 1)
            h = blk(h)
            if i == 2 or i == 3:
                h = self.upsample(h)
        out = self.last_conv(h)
        return torch.sigmoid(out) if sigmoid else out


def _engine_only(name):
    raise RuntimeError(f"{name}.forward is not a PyTorch op in streamingflow_b200: this layer is evaluated inside the fused "
                       "CUDA stages of the ODE engine (call NNFOwithBayesianJumps.infer_state / ode_step / forward).")


class SELayer(nn.Module):
    """Squeeze-excite weights (res_models.py:150-165); evaluated by the engine's se_reduce / se_apply kernels."""

    def __init__(self, channel, reduction=8):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())

    def forward(self, x):
        _engine_only("SELayer")


class ConvNet(nn.Module):
    """p_model: ResBlock, SE, ResBlock, SE, ConvBlock(bias, no norm) (res_models.py:168-180); engine stages q1..q5."""

    def __init__(self, in_c, out_c):
        super().__init__()
        self.model = nn.Sequential(ResBlock(in_c, out_c), SELayer(out_c), ResBlock(out_c, out_c), SELayer(out_c),
                                   ConvBlock(out_c, out_c, 3, stride=1, bias=True, norm='none'))

    def forward(self, x):
        _engine_only("ConvNet")
