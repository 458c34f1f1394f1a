"""Drop-in for ``streamingflow/models/decoder.py::Decoder`` (reference :8-140), the step after the ODE head that turns its
output into occupancy logits (and the other BEV heads): same constructor, ``forward(x) -> dict`` and ``state_dict`` names (the
ResNet-18 layers come from torchvision exactly as in the reference, so a checkpoint loads with ``strict=True``).

All arithmetic runs on the CUDA engine (seg_head_engine.py -> libsf_b200.so): eval mode only, CUDA tensors only, no PyTorch
fallback.  Extra, not in the reference: ``forward(x, planes=...)`` accepts the fused refinement's output in engine layout
(no NCHW fp32 round trip), and the result dict carries ``segmentation_argmax`` -- the uint8 masks of
``segmentation.argmax(dim=2)`` (trainer.py:230-231) computed in the same kernel as the logits.
"""
import torch
import torch.nn as nn

from .. import _lib as L


def _head(channels, k, sigmoid=False):
    layers = [nn.Conv2d(channels, channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(channels), nn.ReLU(inplace=True),
              nn.Conv2d(channels, k, kernel_size=1, padding=0)]
    return nn.Sequential(*(layers + ([nn.Sigmoid()] if sigmoid else [])))


class _UpsamplingAdd(nn.Module):
    """Parameter container of the reference's UpsamplingAdd (convolutions.py:204-215): Upsample, 1x1 conv, BatchNorm."""

    def __init__(self, in_channels, out_channels, scale_factor=2):
        super().__init__()
        self.upsample_layer = nn.Sequential(nn.Upsample(scale_factor=scale_factor, mode='bilinear', align_corners=False),
                                            nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0, bias=False),
                                            nn.BatchNorm2d(out_channels))

    def forward(self, x, x_skip):
        raise RuntimeError("UpsamplingAdd is evaluated inside the CUDA Decoder engine; call Decoder.forward")


class Decoder(nn.Module):
    def __init__(self, in_channels, n_classes, n_present, n_hdmap, predict_gate):
        super().__init__()
        from torchvision.models.resnet import resnet18

        self.perceive_hdmap = predict_gate['perceive_hdmap']
        self.predict_pedestrian = predict_gate['predict_pedestrian']
        self.predict_instance = predict_gate['predict_instance']
        self.predict_future_flow = predict_gate['predict_future_flow']
        self.planning = predict_gate['planning']
        self.n_classes, self.n_present = n_classes, n_present
        if self.predict_instance is False and self.predict_future_flow is True:
            raise ValueError('flow cannot be True when not predicting instance')
        backbone = resnet18(weights=None, zero_init_residual=True)
        self.first_conv = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1, self.relu = backbone.bn1, backbone.relu
        self.layer1, self.layer2, self.layer3 = backbone.layer1, backbone.layer2, backbone.layer3
        c = in_channels
        self.up3_skip = _UpsamplingAdd(256, 128, scale_factor=2)
        self.up2_skip = _UpsamplingAdd(128, 64, scale_factor=2)
        self.up1_skip = _UpsamplingAdd(64, c, scale_factor=2)
        self.segmentation_head = _head(c, n_classes)
        if self.predict_pedestrian:
            self.pedestrian_head = _head(c, n_classes)
        if self.perceive_hdmap:
            self.hdmap_head = _head(c, 2 * n_hdmap)
        if self.predict_instance:
            self.instance_offset_head = _head(c, 2)
            self.instance_center_head = _head(c, 1, sigmoid=True)
        if self.predict_future_flow:
            self.instance_future_head = _head(c, 2)
        if self.planning:
            self.costvolume_head = _head(c, 1)
        self.precision = "bf16"
        self.__dict__["_engines"] = {}

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engines"] = {}
        return d

    def _heads(self):
        keys = ["segmentation"]
        keys += ["pedestrian"] if self.predict_pedestrian else []
        keys += ["hdmap"] if self.perceive_hdmap else []
        keys += ["instance_center", "instance_offset"] if self.predict_instance else []
        keys += ["instance_flow"] if self.predict_future_flow else []
        keys += ["costvolume"] if self.planning else []
        return keys

    def _engine_for(self, H, W, n, device):
        from ..seg_head_engine import SegHeadEngine

        if self.training:
            raise L.SfError("the CUDA Decoder engine implements inference (eval mode, BatchNorm folded); call .eval()")
        if device.type != "cuda":
            raise L.SfError("streamingflow_b200's Decoder runs on a B200 GPU only; got a tensor on " + str(device))
        key = (str(device), H, W, self.precision)
        fp = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        ent = self._engines.get(key)
        if ent is None or ent["fp"] != fp or ent["engine"].n < n:
            if len(self._engines) >= 2:
                self._engines.clear()
            ent = dict(engine=SegHeadEngine(self.state_dict(), H, W, n, self.precision, device, self._heads()), fp=fp)
            self._engines[key] = ent
        return ent["engine"]

    def forward(self, x, planes=None):
        """x: [b, s, c, h, w] fp32 (reference :91-93).  planes: optionally the same frames as (hi, lo) NHWC bf16 [b*s, h, w, c]
        in engine layout (then only x's shape is used)."""
        b, s, c, h, w = x.shape
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise L.SfError("streamingflow_b200's Decoder is inference only (no backward): call it under torch.no_grad()")
        eng = self._engine_for(h, w, b * s, x.device)
        res = eng.run(None if planes is not None else x.reshape(b * s, c, h, w), planes=planes)
        view = lambda t: t.view(b, s, *t.shape[1:])
        out = {k: None for k in ('segmentation', 'pedestrian', 'hdmap', 'instance_center', 'instance_offset', 'instance_flow', 'costvolume')}
        for k, v in res.items():
            if k == "hdmap":
                out[k] = view(v)[:, self.n_present - 1]                 # the reference evaluates this head on the present frame only
            elif k == "costvolume":
                out[k] = view(v).squeeze(2)
            else:
                out[k] = view(v)
        return out
