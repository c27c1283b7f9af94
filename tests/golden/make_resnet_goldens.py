"""Golden vectors of the FiLM-ResNet-50 token producer, build container only:

    python tests/golden/make_resnet_goldens.py

Runs the REFERENCE's FiLMResNet50Policy.forward (pretrained_resnets.py:39-60) — imported from /root/reference with a
stand-in `timm.create_model` that returns torchvision's resnet50, see oracle/film_resnet_ref.py — on counter-based weights
and inputs, checks that the checkout-free restatement (FiLMResNet50Oracle) reproduces it bit for bit, and stores the
(N, 2048) outputs for two image sizes plus the reference's own bf16-autocast output for context."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import film_resnet_ref as R  # noqa: E402

torch.set_grad_enabled(False)
COND = 512
sd = R.synthetic_state_dict(COND)
ref = R.reference_module(COND)
assert list(ref.state_dict().keys()) == list(sd.keys()), "state_dict names/order drifted"
ref.load_state_dict(sd)
ora = R.FiLMResNet50Oracle(COND).eval()
ora.load_state_dict(sd)
out = {}
for tag, n, size in (("s64", 3, 64), ("s112", 2, 112)):
    img, cond = R.synthetic_inputs(n, size, COND)
    y = ref(img, cond)
    assert torch.equal(y, ora(img, cond)), "restatement differs from the reference forward"
    with torch.autocast("cpu", dtype=torch.bfloat16):
        yb = ref(img, cond).float()
    out[f"{tag}_out"] = y.numpy()
    out[f"{tag}_out_autocast_bf16"] = yb.numpy()
    rel = float((yb - y).norm() / y.norm())
    print(f"{tag}: out {tuple(y.shape)} |y| mean {float(y.abs().mean()):.3f}; reference autocast-bf16 vs fp32 {rel:.3e}")
np.savez_compressed(Path(__file__).resolve().parent / "film_resnet50.npz", **out)
