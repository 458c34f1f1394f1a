"""Drop-in for ``streamingflow/models/future_prediction_ode.py::FuturePredictionODE`` (reference :9-64).

Same constructor, ``forward`` signature, return value ``(x [B, T, C, H, W], 0)`` and ``state_dict`` names, so a
reference checkpoint loads with ``strict=True`` and ``streamingflow.forward`` (models/streamingflow.py:258-267) can
call it unchanged.  The per-sample Python loop of the reference (:36-51) becomes ONE batched rollout on the CUDA
engine: all observation frames of all samples are encoded together, every sample's schedule is computed on the host,
and the samples advance event by event (rollout.py).  The result is the same as the reference's loop because
samples are independent and the noise is drawn in the reference's sample-major order.
"""
import os

import torch
import torch.nn as nn

from ..layers.convolutions import Block, DeepLabHead
from ..layers.temporal import SpatialGRU
from ..layers.temporal_ode_bayes import NNFOwithBayesianJumps
from ..schedule import merge_observations


class _StackSpec:
    """Shape and device of ``torch.stack(frames)`` without the copy."""

    def __init__(self, frames):
        self.shape = (len(frames),) + tuple(frames[0].shape)
        self.device = frames[0].device


class FuturePredictionODE(nn.Module):
    def __init__(self, in_channels, latent_dim, n_future, cfg, mixture=True, n_gru_blocks=2, n_res_layers=1, delta_t=0.05):
        super().__init__()
        self.n_spatial_gru = n_gru_blocks
        self.delta_t = delta_t
        self.gru_ode = NNFOwithBayesianJumps(input_size=in_channels, hidden_size=latent_dim, cfg=cfg, mixing=int(mixture))
        grus, blocks = [], []
        for i in range(n_gru_blocks):
            grus.append(SpatialGRU(in_channels, in_channels))
            last = i == n_gru_blocks - 1
            blocks.append(DeepLabHead(in_channels, in_channels, 128) if last
                          else nn.Sequential(*[Block(in_channels) for _ in range(n_res_layers)]))
        self.spatial_grus = nn.ModuleList(grus)
        self.res_blocks = nn.ModuleList(blocks)
        self.in_channels, self.n_res_layers = in_channels, n_res_layers
        self.fused_refine = os.environ.get("SF_B200_FUSED_REFINE", "1") == "1"   # SpatialGRU / Block / DeepLabHead on the CUDA engine
        # the whole forward (layout pack, encoder, step loop, decoder, refinement: ~340 launches) captured once per schedule and
        # replayed as ONE CUDA graph; the noise draw, the gather of the caller's frames and the final layout unpack stay eager
        self.forward_graph = os.environ.get("SF_B200_FORWARD_GRAPH", "1") == "1"
        self.__dict__["_refiners"] = {}
        self.__dict__["_fwd_graphs"] = {}

    def _refine_for(self, H, W, B, T, device):
        from ..refine_engine import RefineEngine

        prec = self.gru_ode.precision
        key = (str(device), H, W, B, T, prec)
        fp = tuple((p.data_ptr(), p._version) for mod in (self.spatial_grus, self.res_blocks) for p in list(mod.parameters()) + list(mod.buffers()))
        ent = self._refiners.get(key)
        if ent is None or ent["fp"] != fp:
            if len(self._refiners) >= 2:
                self._refiners.clear()
            sd = {k: v for k, v in self.state_dict().items() if k.startswith(("spatial_grus", "res_blocks"))}
            ent = dict(engine=RefineEngine(sd, H, W, B, T, prec, device), fp=fp)
            self._refiners[key] = ent
        return ent["engine"]

    def _forward_graphed(self, frames, shape, counts, times, targets, dtypes, refine):
        """Encoder -> ODE loop -> decoder -> refinement as ONE captured CUDA graph per (shapes, schedule, weights version).
        Eager around the replay: the gather of the caller's frames into the graph's static input, the rollout's noise (one launch
        on torch's Philox stream, exactly the draws the eager path makes) and the unpack of the result into a fresh tensor."""
        ode = self.gru_ode
        n, C, H, W = shape
        dev = frames[0].device
        prep = ode.fused_prep(n, H, W, counts, times, targets, self.delta_t, dtypes, dev)
        eng, codec, ro = prep["eng"], prep["codec"], prep["ro"]
        eng.ensure_path_slots(ro.n_path)
        key = (id(prep), id(refine))
        ent = self._fwd_graphs.get(key)
        if ent is not None and (ent["prep"] is not prep or ent["refine"] is not refine or ent["gen"] != eng.alloc_gen):
            ent = None
        if ent is None:
            if len(self._fwd_graphs) >= 4:
                self._fwd_graphs.clear()
            static_in = torch.empty((n, C, H, W), dtype=torch.float32, device=dev)
            eps = torch.empty((max(ro.n_eps, 1), C, H // 4, W // 4), dtype=torch.float32, device=dev)
            torch.stack(frames, dim=0, out=static_in)
            eps.zero_()
            # one eager pass on the static buffers: builds every per-size table, binds every buffer (nothing of it is kept)
            planes, x32 = ode.fused_run(prep, static_in, eps)
            refine.run_core(planes, x32)
            torch.cuda.synchronize(dev)
            gen0 = eng.alloc_gen
            codec.launches = refine.launches = 0
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                planes, x32 = ode.fused_run(prep, static_in, eps)
                refine.run_core(planes, x32)
            if eng.alloc_gen != gen0:
                raise RuntimeError("a buffer moved during the capture of the forward graph")
            ent = self._fwd_graphs[key] = dict(graph=graph, static_in=static_in, eps=eps, prep=prep, refine=refine, gen=eng.alloc_gen,
                                               launches=ro.launches + codec.launches + refine.launches)
            codec.launches = refine.launches = 0
        torch.stack(frames, dim=0, out=ent["static_in"])
        ode._noise_into(ent["eps"], ro.n_eps, H // 4, W // 4, dev, live=ode.live_noise_slots(ro, dev, ent))
        ent["graph"].replay()
        ro.launches = ent["launches"]
        ode.last_rollout = ro
        refine.launches = 0
        return refine.unpack_output()

    def __getstate__(self):
        d = dict(self.__dict__)          # the refinement engines hold ctypes plan handles: per-instance, rebuilt on first use
        d["_refiners"], d["_fwd_graphs"] = {}, {}
        d.pop("last_output_planes", None)
        return d

    @staticmethod
    def _host_times(t):
        """Timestamps as python doubles with ONE device->host copy (the reference syncs once per dict key)."""
        return None if t is None else t.detach().to("cpu", torch.float64).tolist()

    def forward(self, future_prediction_input, camera_states, lidar_states, camera_timestamp, lidar_timestamp, target_timestamp):
        # camera_states [B, n_cam, C, H, W]; lidar_states [B, n_lidar, C, H, W] or None; timestamps [B, n] (seconds)
        B = camera_states.shape[0]
        if B == 0:                       # an empty shard (more ranks than samples): nothing to integrate
            T = target_timestamp.shape[1]
            return camera_states.new_zeros((0, T) + tuple(camera_states.shape[2:])), 0
        if torch.is_grad_enabled() and (camera_states.requires_grad or (lidar_states is not None and lidar_states.requires_grad)
                                        or any(p.requires_grad for p in self.parameters())):
            from .. import _lib as L
            raise L.SfError("streamingflow_b200's FuturePredictionODE is inference only (no backward): call it under torch.no_grad(); "
                            "with autograd recording the gradients to camera_states / lidar_states would be cut silently")
        # dtype of the stamp tensors: the reference's loop compares / subtracts in it (schedule.plan_sample).  Its per-sample
        # ``times = torch.tensor(list_of_0-dim_keys)`` is float32 only when every stamp is float32.
        obs_dtype = "float32" if (camera_timestamp.dtype == torch.float32 and
                                  (lidar_states is None or lidar_timestamp.dtype == torch.float32)) else "float64"
        tgt_dtype = "float32" if target_timestamp.dtype == torch.float32 else "float64"
        dtypes = (obs_dtype, tgt_dtype)
        cam_t = self._host_times(camera_timestamp)
        lid_t = self._host_times(lidar_timestamp) if lidar_states is not None else None
        tgt_t = self._host_times(target_timestamp)
        frames, counts, times = [], [], []
        for b in range(B):
            order = merge_observations(cam_t[b], None if lid_t is None else lid_t[b])
            frames += [camera_states[b, i] if sensor == 0 else lidar_states[b, i] for _, sensor, i in order]
            counts.append(len(order))
            times.append([t for t, _, _ in order])
        ode = self.gru_ode
        stacked = _StackSpec(frames)                  # shape / device of torch.stack(frames); the copy itself happens where it is consumed
        H, W = stacked.shape[2], stacked.shape[3]
        fused = ode.codec_available(H, W, stacked.device)
        if fused and self.fused_refine and self.n_spatial_gru == 2 and self.n_res_layers == 1 and self.in_channels in (64, 128):
            # encoder -> ODE loop -> decoder -> SpatialGRU / Block / SpatialGRU / DeepLabHead, all on the CUDA engine
            T = len(tgt_t[0])
            refine = self._refine_for(H, W, B, T, stacked.device)
            if self.forward_graph and not ode.record_all:
                x = self._forward_graphed(frames, stacked.shape, counts, times, tgt_t, dtypes, refine)
            else:
                _, (planes, x32) = ode.encode_integrate_decode(torch.stack(frames, dim=0), counts, times, tgt_t, self.delta_t, raw=True,
                                                               stamp_dtypes=dtypes)
                x = refine.run(planes, x32)
            # the same frames in engine layout ((hi, lo) NHWC bf16 [B*T, H, W, C], views of the refinement's output buffer, valid
            # until the next forward): streamingflow_b200.models.decoder.Decoder.forward(x, planes=...) consumes them directly
            self.__dict__["last_output_planes"] = refine.output_planes()
            ode.last_rollout.launches += refine.launches
            refine.launches = 0
            return x, 0
        self.__dict__["last_output_planes"] = None
        if fused:
            _, x = ode.encode_integrate_decode(torch.stack(frames, dim=0), counts, times, tgt_t, self.delta_t, stamp_dtypes=dtypes)      # encoder / loop / decoder on the CUDA engine
        else:
            hx = ode.srvp_encoder(torch.stack(frames, dim=0))
            _, sel = ode.integrate_latents(hx, counts, times, tgt_t, self.delta_t, stamp_dtypes=dtypes)
            x = ode.srvp_decode(sel)                                   # [B, T, C, H, W]
        hidden_state = x[:, 0]
        for gru, block in zip(self.spatial_grus, self.res_blocks):
            x = gru(x, hidden_state)
            b, s, c, h, w = x.shape
            x = block(x.view(b * s, c, h, w)).view(b, s, c, h, w)
        return x, 0
