"""A CHECKER backend for CPU tests of the host logic (schedule -> rollout -> event table -> selection glue).

It implements the small surface of ``streamingflow_b200.engine.OdeEngine`` that the nn.Module mirror uses, but
executes every event with the CPU oracle (oracle/sf_oracle.py).  It lives under tests/ and is only ever injected
by tests through ``NNFOwithBayesianJumps._engine_factory``; the product never constructs it.
"""
import torch

from oracle import sf_oracle as so

BUF_S0, BUF_S1, BUF_X, BUF_OBS, BUF_ZERO = 0, 1, 2, 3, 4


class OracleBackend:
    def __init__(self, sd, H, W, max_images, precision, device, dtype=torch.float64):
        self.H, self.W, self.max_images, self.dtype = H, W, max_images, dtype
        self.load_weights(sd, "")
        C = self.sd["g.gru_c.conv_decoder_2.weight"].shape[0]
        self.C = C
        z = lambda n, c=C: torch.zeros(n, c, H, W, dtype=dtype)
        self.state = [z(max_images), z(max_images)]
        self.x = z(max_images)
        self.zero = z(1)
        self.x32, self.params32 = z(max_images), z(max_images, 2 * C)
        self.state32 = self.state
        self.path = z(1)
        self.obs = z(1)
        self.events_run = 0

    def load_weights(self, sd, prefix):
        self.sd = {"g." + k: v.detach().to(self.dtype) if v.is_floating_point() else v for k, v in sd.items()}

    def bind_observations(self, hx):
        self.obs = hx.to(self.dtype)

    def bind_eps(self, eps):
        self.eps = eps.to(self.dtype)

    def zero_state(self, which=0):
        self.state[which].zero_()

    def set_state(self, which, s):
        self.state[which][: s.shape[0]] = s.to(self.dtype)

    def pack_into(self, buf, src):
        assert buf == BUF_X
        self.x[: src.shape[0]] = src.to(self.dtype)

    def ensure_path_slots(self, n):
        if self.path.shape[0] < n:
            self.path = torch.zeros(n, self.C, self.H, self.W, dtype=self.dtype)

    def unpack_path(self, slots):
        return self.path[list(slots)].to(torch.float32 if self.dtype == torch.float32 else self.dtype)

    def unpack_f32(self, t, n):
        return t[:n].clone()

    def snapshot(self, n):
        return [self.state[0][:n].clone(), self.x[:n].clone()]

    def restore(self, snap, n):
        self.state[0][:n] = snap[0]
        self.x[:n] = snap[1]

    def run_rollout(self, events):
        for e in events:
            for j, b in enumerate(e["samples"]):
                src = {BUF_X: self.x, BUF_OBS: self.obs, BUF_ZERO: self.zero}[e["x_buf"]]
                xin = src[e["x_img"][j]][None]
                s_in, s_base = self.state[e["s_in"]][b][None], self.state[e["s_base"]][b][None]
                if e.get("run_cell", 1):
                    if e["kind"] == 0:
                        dt = torch.tensor(float(torch.tensor(e["dt"][j], dtype=torch.float64).to(torch.float32)), dtype=self.dtype) \
                            if self.dtype == torch.float32 else torch.tensor(e["dt"][j], dtype=self.dtype)
                        new = s_base + dt * so.ode_derivative(self.sd, "g.gru_c", xin, s_in)
                    else:
                        new = so.observation_jump(self.sd, "g.gru_obs", s_in, xin)
                    self.state[e["s_out"]][b] = new[0]
                    if e["rec"][j] >= 0:
                        self.path[e["rec"][j]] = new[0]
                if e.get("run_prior", 1):
                    y, params = so.infer_state(self.sd, "g.p_model", self.state[e["s_out"]][b][None], self.eps[e["eps"][j]][None])
                    self.x[b] = y[0]
                    if e.get("want_f32", 0):
                        self.x32[b], self.params32[b] = y[0], params[0]
            self.events_run += 1
        return 0
