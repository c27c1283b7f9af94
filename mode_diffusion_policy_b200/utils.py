"""Helpers of `mode.models.edm_diffusion.utils` the hot path touches (reference utils.py:146-203)."""
import math

import torch


def append_dims(x, target_dims):
    """Right-pad singleton dims until `x` has `target_dims` dims (reference utils.py:146-151)."""
    extra = target_dims - x.ndim
    if extra < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * extra]


def rand_log_normal(shape, loc=0.0, scale=1.0, device="cpu", dtype=torch.float32):
    return (torch.randn(shape, device=device, dtype=dtype) * scale + loc).exp()


def rand_log_logistic(shape, loc=0.0, scale=1.0, min_value=0.0, max_value=float("inf"), device="cpu",
                      dtype=torch.float32):
    """Truncated log-logistic training density, float64 inverse-CDF sampling (reference utils.py:159-166; the default
    `sigma_sample_density_type: loglogistic`, conf/model/mode_agent.yaml)."""
    lo = torch.as_tensor(min_value, device=device, dtype=torch.float64)
    hi = torch.as_tensor(max_value, device=device, dtype=torch.float64)
    cdf_lo = lo.log().sub(loc).div(scale).sigmoid()
    cdf_hi = hi.log().sub(loc).div(scale).sigmoid()
    u = torch.rand(shape, device=device, dtype=torch.float64) * (cdf_hi - cdf_lo) + cdf_lo
    return u.logit().mul(scale).add(loc).exp().to(dtype)


def rand_log_uniform(shape, min_value, max_value, device="cpu", dtype=torch.float32):
    lo, hi = math.log(min_value), math.log(max_value)
    return (torch.rand(shape, device=device, dtype=dtype) * (hi - lo) + lo).exp()


def rand_uniform(shape, min_value, max_value, device="cpu", dtype=torch.float32):
    return torch.rand(shape, device=device, dtype=dtype) * (max_value - min_value) + min_value
