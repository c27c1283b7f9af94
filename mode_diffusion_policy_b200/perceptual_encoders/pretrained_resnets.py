"""Drop-in for `mode.models.perceptual_encoders.pretrained_resnets.FiLMResNet50Policy` (reference
pretrained_resnets.py:25-60): same constructor, same `state_dict()` names and shapes (`resnet.*` as timm / torchvision's
resnet50 without the classifier, `film{1..4}.{gamma,beta}.{weight,bias}`), same `forward(x, condition) -> (N, 2048)`.

The module holds fp32 master parameters and buffers and contains no compute: `forward` runs the engine's FiLM-ResNet
(C ABI `mode_resnet_*`, csrc/resnet.inc) — every convolution a tcgen05 GEMM with BatchNorm folded in, shortcut add, ReLU
and FiLM in the epilogue. Inference only (eval-mode BatchNorm); in train mode `forward` raises: the reference fine-tunes
its encoders through the diffusion loss, which stays with the reference's PyTorch modules (the denoiser hands back
d loss / d state_images for them, `mode_train_input_grads`).

No pretrained weights are downloaded (the reference calls timm.create_model(pretrained=True)): load a checkpoint with
`load_state_dict` — published MoDE checkpoints carry these tensors as `static_resnet.*` / `gripper_resnet.*`.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from .. import _lib

BLOCKS = (3, 4, 6, 3)


class FiLMLayer(nn.Module):
    """Parameter holder of reference pretrained_resnets.py:5-23 (zero-initialised: identity transform at start)."""

    def __init__(self, num_features, condition_dim):
        super().__init__()
        self.num_features, self.condition_dim = num_features, condition_dim
        self.gamma = nn.Linear(condition_dim, num_features)
        self.beta = nn.Linear(condition_dim, num_features)
        for lin in (self.gamma, self.beta):
            nn.init.zeros_(lin.weight)
            nn.init.zeros_(lin.bias)


class _Bottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(planes * 4))


class _ResNet50Params(nn.Module):
    """The parameter / buffer tree of timm's `resnet50(num_classes=0)` (= torchvision's resnet50 minus `fc`)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for s, n in enumerate(BLOCKS):
            planes, blocks = 64 << s, []
            for b in range(n):
                blocks.append(_Bottleneck(inplanes, planes, 2 if (b == 0 and s > 0) else 1, downsample=b == 0))
                inplanes = planes * 4
            setattr(self, f"layer{s + 1}", nn.Sequential(*blocks))


class FiLMResNet50Policy(nn.Module):
    def __init__(self, condition_dim, max_images: int = 256):
        super().__init__()
        self.resnet = _ResNet50Params()
        self.film1 = FiLMLayer(256, condition_dim)
        self.film2 = FiLMLayer(512, condition_dim)
        self.film3 = FiLMLayer(1024, condition_dim)
        self.film4 = FiLMLayer(2048, condition_dim)
        self.condition_dim, self.max_images = condition_dim, max_images
        self._h = None
        self._key = None
        self._geom = None

    # ------------------------------------------------------------------ engine management
    def _fingerprint(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _engine(self, n, height, width):
        lib = _lib.load()
        if self._h is None or self._geom != (height, width) or n > self.max_images:
            self.close()
            self.max_images = max(self.max_images, n)
            h = C.c_void_p()
            _lib.check(lib.mode_resnet_create(self.condition_dim, self.max_images, height, width, C.byref(h)))
            self._h, self._geom, self._key = h, (height, width), None
        key = self._fingerprint()
        if key != self._key:
            for name, t in self.state_dict().items():
                t = t.detach().to(torch.float32).contiguous()
                shp = (C.c_int64 * max(t.dim(), 1))(*(tuple(t.shape) or (1,)))
                _lib.check(lib.mode_resnet_set_weight(self._h, name.encode(), t.data_ptr(), 1 if t.is_cuda else 0, shp, max(t.dim(), 1)))
            _lib.check(lib.mode_resnet_finalize(self._h, torch.cuda.current_stream().cuda_stream))
            self._key = key
        return lib

    def close(self):
        if self._h is not None:
            _lib.load().mode_resnet_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference surface
    def forward(self, x, condition):
        if self.training:
            raise NotImplementedError("MoDE engine: FiLMResNet50Policy is inference-only (eval-mode BatchNorm folded into "
                                      "the convolutions); call .eval(), or train the encoders with the reference module")
        if not x.is_cuda:
            raise _lib.ModeError("images must be CUDA tensors: the MoDE engine has no CPU path")
        if condition.dim() == 3:
            condition = condition.squeeze(1)
        n, _, height, width = x.shape
        lib = self._engine(n, height, width)
        img = x.detach().to(torch.float32).contiguous()
        cond = condition.detach().to(torch.float32).contiguous()
        if cond.shape[0] != n:  # one goal per sample, several frames per sample ('b t c h w -> (b t) c h w' upstream)
            cond = cond.repeat_interleave(n // cond.shape[0], dim=0).contiguous()
        out = torch.empty((n, 2048), dtype=torch.float32, device=x.device)
        _lib.check(lib.mode_resnet_forward(self._h, img.data_ptr(), cond.data_ptr(), out.data_ptr(), n,
                                           torch.cuda.current_stream().cuda_stream))
        return out.to(x.dtype)
