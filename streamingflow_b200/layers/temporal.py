"""Post-ODE refinement GRU with the reference's parameter names (reference: streamingflow/layers/temporal.py:11-57).
At the shipped width (64 channels) FuturePredictionODE evaluates it on the CUDA engine (refine_engine.py: the gate /
proposal stages of the ODE cell's kernels with a single pair, "next" row 2 of SURVEY.md 8f); the PyTorch ``forward`` below is
only reached for other widths (e.g. the 128-channel module of BASELINE config 5)."""
import torch
import torch.nn as nn


class SpatialGRU(nn.Module):
    """ConvGRU over a [B, T, C, H, W] sequence; each step's state goes through a 1x1 decoder conv."""

    def __init__(self, input_size, hidden_size, gru_bias_init=0.0):
        super().__init__()
        self.input_size, self.hidden_size, self.gru_bias_init = input_size, hidden_size, gru_bias_init
        cat = input_size + hidden_size
        self.conv_update = nn.Conv2d(cat, hidden_size, kernel_size=3, bias=True, padding=1)
        self.conv_reset = nn.Conv2d(cat, hidden_size, kernel_size=3, bias=True, padding=1)
        self.conv_state_tilde = nn.Conv2d(cat, hidden_size, kernel_size=3, bias=True, padding=1)
        self.conv_decoder = nn.Conv2d(hidden_size, input_size, kernel_size=1, bias=False)

    def gru_cell(self, x, state):
        xs = torch.cat([x, state], dim=1)
        update = torch.sigmoid(self.conv_update(xs) + self.gru_bias_init)
        reset = torch.sigmoid(self.conv_reset(xs) + self.gru_bias_init)
        proposal = self.conv_state_tilde(torch.cat([x, (1.0 - reset) * state], dim=1))
        return (1.0 - update) * state + update * proposal

    def forward(self, x, state=None):
        assert x.dim() == 5, 'Input tensor must be BxTxCxHxW.'
        b, steps, _, h, w = x.shape
        if state is None:
            state = torch.zeros(b, self.hidden_size, h, w, device=x.device)
        frames = []
        for t in range(steps):
            state = self.gru_cell(x[:, t], state)
            frames.append(self.conv_decoder(state))
        return torch.stack(frames, dim=1)
