"""TEST INFRASTRUCTURE ONLY — CPU oracle of the FiLM-ResNet-50 token producer (reference
mode/models/perceptual_encoders/pretrained_resnets.py:25-60).

The reference builds its backbone with `timm.create_model('resnet50', pretrained=True, num_classes=0)`; timm is not
installed in this image and nothing may be downloaded. `reference_module()` therefore imports the REFERENCE's own
`FiLMResNet50Policy` with a stand-in `timm.create_model` that returns torchvision's resnet50 (the same architecture and
state_dict names: bottleneck [3, 4, 6, 3], stride on the 3x3 convolution) wearing timm's attribute names (`act1`,
`global_pool`, `fc = Identity`). The FiLM logic, the call order and the pooling are then the reference's code; only the
backbone constructor is substituted. Parity pin for this row: reference code + torchvision backbone, NOT a timm run —
"parity pinned up to the backbone constructor". `synthetic_state_dict` gives counter-based weights (incl. BatchNorm
running statistics and non-zero FiLM weights: the reference initialises FiLM to zero, which would hide it).
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch
from torch import nn


def _torchvision_as_timm(name, pretrained=False, num_classes=0, **kw):
    import torchvision

    assert name == "resnet50" and num_classes == 0
    m = torchvision.models.resnet50(weights=None)
    m.fc = nn.Identity()
    m.act1 = m.relu
    m.global_pool = nn.Sequential(m.avgpool, nn.Flatten(1))  # timm's SelectAdaptivePool2d(avg, flatten=True)
    return m


def reference_module(cond_dim: int, ref_root: str = "/root/reference"):
    """The reference's FiLMResNet50Policy (its own forward code) over a torchvision backbone; build container only."""
    stub = types.ModuleType("timm")
    stub.create_model = _torchvision_as_timm
    saved = sys.modules.get("timm")
    sys.modules["timm"] = stub
    sys.path.insert(0, ref_root)
    try:
        import importlib

        mod = importlib.import_module("mode.models.perceptual_encoders.pretrained_resnets")
        mod = importlib.reload(mod)
        return mod.FiLMResNet50Policy(cond_dim).eval()
    finally:
        sys.path.remove(ref_root)
        if saved is None:
            del sys.modules["timm"]
        else:
            sys.modules["timm"] = saved


class FiLMResNet50Oracle(nn.Module):
    """Restatement without the reference checkout (for the GPU box): torchvision resnet50 + the FiLM forward of
    pretrained_resnets.py:39-60. Checked against `reference_module` by tests/golden/make_resnet_goldens.py."""

    def __init__(self, cond_dim):
        super().__init__()
        import torchvision

        self.resnet = torchvision.models.resnet50(weights=None)
        self.resnet.fc = nn.Identity()
        for i, c in enumerate((256, 512, 1024, 2048)):
            f = nn.Module()
            f.gamma, f.beta = nn.Linear(cond_dim, c), nn.Linear(cond_dim, c)
            setattr(self, f"film{i + 1}", f)

    def forward(self, x, condition):
        if condition.dim() == 3:
            condition = condition.squeeze(1)
        r = self.resnet
        x = r.maxpool(r.relu(r.bn1(r.conv1(x))))
        for i in range(4):
            x = getattr(r, f"layer{i + 1}")(x)
            f = getattr(self, f"film{i + 1}")
            gamma, beta = f.gamma(condition)[:, :, None, None], f.beta(condition)[:, :, None, None]
            x = (1 + gamma) * x + beta  # pretrained_resnets.py:19-22
        return torch.flatten(r.avgpool(x), 1)


def synthetic_state_dict(cond_dim: int, seed: int = 77) -> dict:
    """Counter-based weights in the reference's state_dict layout (numpy generator; same arrays for oracle and engine)."""
    m = FiLMResNet50Oracle(cond_dim)
    rng = np.random.default_rng(seed)
    sd = {}
    for k, v in m.state_dict().items():
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_var"):
            sd[k] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif k.endswith("running_mean"):
            sd[k] = torch.from_numpy((0.1 * rng.standard_normal(shape)).astype(np.float32))
        elif ".bn" in k or "downsample.1" in k or k.startswith("resnet.bn1"):
            lo, hi = (0.15, 0.35) if ".bn3." in k else (0.8, 1.2)  # small last-BN gains keep the residual sums O(1)
            sd[k] = torch.from_numpy((rng.uniform(lo, hi, shape) if k.endswith("weight") else 0.05 * rng.standard_normal(shape)).astype(np.float32))
        elif k.startswith("film"):
            scale = 0.5 / np.sqrt(cond_dim) if k.endswith("weight") else 0.1
            sd[k] = torch.from_numpy((scale * rng.standard_normal(shape)).astype(np.float32))
        else:  # convolution: He-normal over the fan-in
            fan_in = int(np.prod(shape[1:]))
            sd[k] = torch.from_numpy((rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32))
    return sd


def synthetic_inputs(n: int, size: int, cond_dim: int, seed: int = 5):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((n, 3, size, size)).astype(np.float32)
    cond = rng.standard_normal((n, cond_dim)).astype(np.float32)
    return torch.from_numpy(img), torch.from_numpy(cond)
