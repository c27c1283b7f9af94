"""Hugging Face export of a MoDE checkpoint — the wire format of the reference's `mode/utils/save_to_hf.py:98-157`.

The reference script loads a Lightning `.ckpt`, strips `model.` from every key, and writes a folder with
`model.pt` (`{'state_dict': ...}`), `model.safetensors`, `config.json` (the `model_config` subset of the Hydra config,
save_to_hf.py:11-39) and a model card, then uploads the folder. This module writes the same four files from a
state dict (an agent-level one, a `GCDenoiser` / `MoDeDiT` module, or a checkpoint path), and reads them back.

Key cleaning. The reference applies `k.replace('model.', '')` (save_to_hf.py:121), which removes EVERY occurrence of the
substring — `model.inner_model.blocks.0.ln_1.g` becomes `inner_blocks.0.ln_1.g` because `inner_model.` itself ends in
`model.`. Files written by the reference therefore carry `inner_*` keys for the denoiser. `clean_key` reproduces that
(so exported files are interchangeable with the reference's), `restore_key` inverts it for the denoiser's tensors, and
`checkpoint.load_pretrained_parameters` accepts either spelling.
"""
from __future__ import annotations

import json
import os
from typing import Mapping, Optional

import torch

CONFIG_KEYS = ("latent_dim", "obs_enc_dim", "cond_dim", "resnet_type", "multistep", "sampler_type", "num_sampling_steps",
               "sigma_data", "sigma_min", "sigma_max", "noise_scheduler", "sigma_sample_density_type", "act_window_size",
               "use_proprio")  # save_to_hf.py:19-36

MODEL_CARD = """# MoDE (Mixture of Diffusion Experts) Model

Pretrained MoDE policy for language-conditioned robotic manipulation: noise-conditioned expert routing, diffusion-based
action generation with noise-conditioned self-attention, vision + language inputs.

Files: `model.safetensors` / `model.pt` (weights, keys as written by the reference's `mode/utils/save_to_hf.py`),
`config.json` (`model_config`: the constructor arguments of the agent and of `model.inner_model`).

Exported by mode_diffusion_policy_b200.save_to_hf (B200 engine for the MoDE denoising path); the files load into the
reference implementation unchanged.
"""


def clean_key(key: str) -> str:
    """The reference's key cleaning, verbatim in effect (save_to_hf.py:121)."""
    return key.replace("model.", "")


def restore_key(key: str) -> str:
    """Inverse of `clean_key` for denoiser tensors: `inner_<k>` -> `model.inner_model.<k>`; other keys unchanged."""
    return "model.inner_model." + key[len("inner_"):] if key.startswith("inner_") else key


def _as_state_dict(source) -> dict:
    if isinstance(source, (str, os.PathLike)):
        from .checkpoint import read_state_dict

        return read_state_dict(str(source))
    if isinstance(source, torch.nn.Module):
        sd = source.state_dict()
        # a bare denoiser is exported under the agent-level names the reference's checkpoints use
        if any(k.startswith("inner_model.") for k in sd):
            return {"model." + k: v for k, v in sd.items()}
        if "sigma_emb.weight" in sd:
            return {"model.inner_model." + k: v for k, v in sd.items()}
        return dict(sd)
    return dict(source)


def model_config(config: Mapping) -> dict:
    """`cleaned_config` of save_to_hf.py:18-38 from a plain mapping (an OmegaConf container works too)."""
    get = config.get if hasattr(config, "get") else lambda k, d=None: getattr(config, k, d)
    inner = None
    model = get("model")
    if model is not None:
        inner = model.get("inner_model") if hasattr(model, "get") else getattr(model, "inner_model", None)
    out = {k: get(k) for k in CONFIG_KEYS}
    out["model"] = {"inner_model": _plain(inner)}
    ordered = {k: out[k] for k in ("latent_dim", "obs_enc_dim", "cond_dim", "resnet_type")}
    ordered["model"] = out["model"]
    ordered.update({k: out[k] for k in CONFIG_KEYS[4:]})
    return {"model_config": ordered}


def _plain(x):
    if x is None or isinstance(x, (str, int, float, bool)):
        return x
    if isinstance(x, Mapping) or hasattr(x, "items"):
        return {str(k): _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    return str(x)


def export(source, config: Mapping, out_dir: str, safetensors: bool = True) -> dict:
    """Write model.pt, model.safetensors, config.json and README.md into `out_dir`. Returns {file name: path}."""
    os.makedirs(out_dir, exist_ok=True)
    cleaned = {clean_key(k): v.detach().cpu().contiguous() for k, v in _as_state_dict(source).items()}
    paths = {"model.pt": os.path.join(out_dir, "model.pt")}
    torch.save({"state_dict": cleaned}, paths["model.pt"])
    if safetensors:
        from safetensors.torch import save_file

        paths["model.safetensors"] = os.path.join(out_dir, "model.safetensors")
        save_file(cleaned, paths["model.safetensors"])
    paths["config.json"] = os.path.join(out_dir, "config.json")
    with open(paths["config.json"], "w") as f:
        json.dump(model_config(config), f, indent=2)
    paths["README.md"] = os.path.join(out_dir, "README.md")
    with open(paths["README.md"], "w") as f:
        f.write(MODEL_CARD)
    return paths


def load_export(folder: str) -> tuple[dict, dict]:
    """(state dict with agent-level keys restored for the denoiser, model_config) from an exported folder."""
    st = os.path.join(folder, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file

        sd = load_file(st)
    else:
        sd = torch.load(os.path.join(folder, "model.pt"), map_location="cpu", weights_only=True)["state_dict"]
    with open(os.path.join(folder, "config.json")) as f:
        cfg = json.load(f)["model_config"]
    return {restore_key(k): v for k, v in sd.items()}, cfg


def upload(folder: str, repo_id: str, commit_message: str = "Upload MoDE model") -> None:
    """save_to_hf.py:137-152. Needs network access and `huggingface_hub`; not exercised by the tests."""
    from huggingface_hub import HfApi, upload_folder

    HfApi().create_repo(repo_id, exist_ok=True)
    upload_folder(folder_path=folder, repo_id=repo_id, repo_type="model", commit_message=commit_message)
