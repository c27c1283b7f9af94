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


def rand_v_diffusion(shape, sigma_data=1.0, min_value=0.0, max_value=float("inf"), device="cpu", dtype=torch.float32):
    """Truncated v-diffusion timestep density (reference utils.py:176-181)."""
    cdf_lo = math.atan(min_value / sigma_data) * 2 / math.pi
    cdf_hi = math.atan(max_value / sigma_data) * 2 / math.pi
    u = torch.rand(shape, device=device, dtype=dtype) * (cdf_hi - cdf_lo) + cdf_lo
    return torch.tan(u * math.pi / 2) * sigma_data


def rand_split_log_normal(shape, loc, scale_1, scale_2, device="cpu", dtype=torch.float32):
    """Split log-normal: half-normal magnitudes to the left (scale_1) or right (scale_2) of `loc` (reference :184-191)."""
    n = torch.randn(shape, device=device, dtype=dtype).abs()
    u = torch.rand(shape, device=device, dtype=dtype)
    left, right = n * -scale_1 + loc, n * scale_2 + loc
    return torch.where(u < scale_1 / (scale_1 + scale_2), left, right).exp()


def rand_discrete(shape, values, device="cpu", dtype=torch.float32):
    """Uniform draws from a table of noise levels (reference utils.py:194-198)."""
    idx = torch.randint(0, len(values), shape, device=device)
    return torch.index_select(values, 0, idx).to(dtype)
