"""Gaussian re-parameterisation helpers with the reference's names (reference: streamingflow/models/model_utils.py:60-109).
On the CUDA path this arithmetic lives in the q5 stage's epilogue (SF_EPI_SAMPLE); these host functions document the
formula and serve small host-side uses."""
import torch
import torch.nn.functional as F
import torch.distributions as distrib


def make_normal_from_raw_params(raw_params, scale_stddev=1, dim=2, eps=1e-8, max_log_sigma=-10000, min_log_sigma=10000):
    """Normal(loc, (softplus(raw) + eps) * scale_stddev) from a tensor holding [loc | raw] along the channel axis."""
    axis = 2 if raw_params.dim() == 5 else 1
    loc, raw = torch.chunk(raw_params, 2, axis)
    assert loc.shape[axis] == raw.shape[axis]
    return distrib.Normal(loc, (F.softplus(raw) + eps) * scale_stddev)


def rsample_normal(raw_params, scale_stddev=1, max_log_sigma=-10000, min_log_sigma=10000):
    return make_normal_from_raw_params(raw_params, scale_stddev=scale_stddev).rsample()
