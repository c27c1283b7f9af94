"""Checkpoint I/O for the denoiser: the part of `MoDEAgent.load_pretrained_parameters` (reference
mode_agent.py:134-265) that concerns `model.*` (GCDenoiser / MoDeDiT) keys, plus the matching writer.

Published MoDE checkpoints (`model_cleaned.safetensors` / `model_cleaned.pt`, reference README.md:112-114) are keyed
as `MoDEAgent.state_dict()`: the denoiser's tensors live under `model.inner_model.*`; the image encoders
(`static_resnet.*`, `gripper_resnet.*`, legacy `img_encoder_*`) and CLIP (`*visual*`, `*clip*`) stay with the
reference and are ignored here. The reference loads non-strictly and skips shape mismatches; so does this loader, and
it reports what it did instead of printing.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import torch

DENOISER_PREFIXES = ("model.inner_model.", "inner_model.", "inner_", "")  # agent-level, GCDenoiser-level, HF-export
# (the reference's save_to_hf.py strips every "model." -> "inner_<key>", see save_to_hf.clean_key), MoDeDiT-level keys


@dataclass
class LoadReport:
    loaded: list = field(default_factory=list)
    skipped_shape: list = field(default_factory=list)  # (key, checkpoint shape, model shape)
    missing: list = field(default_factory=list)         # model keys the checkpoint does not provide
    ignored: int = 0                                    # encoder / CLIP / unrelated tensors in the file


def read_state_dict(ckpt_path: str, trust_pickle: bool = False) -> dict:
    """Directory with model_cleaned.safetensors / model_cleaned.pt, or a single .safetensors / .pt / .ckpt file
    (reference mode_agent.py:143-161)."""
    if os.path.isdir(ckpt_path):
        st = os.path.join(ckpt_path, "model_cleaned.safetensors")
        pt = os.path.join(ckpt_path, "model_cleaned.pt")
        if os.path.exists(st):
            ckpt_path = st
        elif os.path.exists(pt):
            ckpt_path = pt
        else:
            raise FileNotFoundError(f"No cleaned weights found in {ckpt_path}")
    if ckpt_path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(ckpt_path)
    try:
        data = torch.load(ckpt_path, map_location="cpu", weights_only=True)
    except Exception:
        # Lightning .ckpt files pickle hyper-parameter objects next to the tensors; the reference loads them with
        # plain torch.load (mode_agent.py:155-161). Opt in explicitly: unpickling runs code from the file.
        if not trust_pickle:
            raise
        data = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    return data["state_dict"] if isinstance(data, dict) and "state_dict" in data else data


def denoiser_state_dict(state_dict: dict, model_keys) -> tuple[dict, int]:
    """Select and rename the MoDeDiT tensors of an agent-level state dict. Returns ({MoDeDiT key: tensor}, ignored)."""
    model_keys = set(model_keys)
    out, ignored = {}, 0
    for key, tensor in state_dict.items():
        if "visual" in key or "clip" in key.lower():  # reference mode_agent.py:211-212
            ignored += 1
            continue
        for prefix in DENOISER_PREFIXES:
            if key.startswith(prefix) and key[len(prefix):] in model_keys:
                out.setdefault(key[len(prefix):], tensor)
                break
        else:
            ignored += 1
    return out, ignored


def load_pretrained_parameters(inner_model: torch.nn.Module, ckpt_path: str, strict: bool = False,
                               freeze_routers: bool = False, trust_pickle: bool = False) -> LoadReport:
    """Load the denoiser's weights from a MoDE checkpoint into `inner_model` (a MoDeDiT). Shape mismatches are skipped
    (strict=False, the reference's default) or raise (strict=True). The engine re-packs lazily on the next call.
    `freeze_routers=True` freezes the routers afterwards like the reference's fine-tuning path
    (`prepare_model_for_finetuning`, mode_agent.py:762-769); `trust_pickle=True` allows Lightning `.ckpt` files."""
    current = inner_model.state_dict()
    found, ignored = denoiser_state_dict(read_state_dict(ckpt_path, trust_pickle), current.keys())
    rep = LoadReport(ignored=ignored)
    new_state = {}
    for key, tensor in found.items():
        if tuple(tensor.shape) == tuple(current[key].shape):
            new_state[key] = tensor.to(current[key].dtype)
            rep.loaded.append(key)
        elif tensor.numel() == current[key].numel():  # same data, other view (e.g. pos_emb saved without batch dim)
            new_state[key] = tensor.reshape(current[key].shape).to(current[key].dtype)
            rep.loaded.append(key)
        else:
            rep.skipped_shape.append((key, tuple(tensor.shape), tuple(current[key].shape)))
    rep.missing = [k for k in current if k not in new_state]
    if strict and (rep.missing or rep.skipped_shape):
        raise RuntimeError(f"Failed to load weights from {ckpt_path}: missing {rep.missing[:5]}, "
                           f"shape mismatches {rep.skipped_shape[:5]}")
    inner_model.load_state_dict(new_state, strict=False)
    if freeze_routers and hasattr(inner_model, "freeze_router"):
        inner_model.freeze_router()
    return rep


def save_denoiser(inner_model: torch.nn.Module, path: str, prefix: str = "model.inner_model.") -> None:
    """Write the denoiser's tensors in the agent-level key layout (`model_cleaned.safetensors` convention)."""
    from safetensors.torch import save_file

    sd = {prefix + k: v.detach().cpu().contiguous() for k, v in inner_model.state_dict().items()}
    save_file(sd, path)
