"""FiLM-ResNet-50 token producer (reference pretrained_resnets.py:25-60, mode_agent.py:548-567).
CPU: the drop-in's parameter tree equals the reference module's state_dict (names, shapes, order) and the checkout-free
oracle reproduces the committed reference outputs. GPU: the engine's GEMM-based forward against those goldens."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import film_resnet_ref as R
from mode_diffusion_policy_b200.perceptual_encoders.pretrained_resnets import FiLMResNet50Policy

GOLD = Path(__file__).resolve().parent / "golden"
COND = 512


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def test_drop_in_state_dict_matches_the_reference_module_and_oracle_matches_goldens():
    sd = R.synthetic_state_dict(COND)
    m = FiLMResNet50Policy(COND)
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == [(k, tuple(v.shape)) for k, v in sd.items()]  # EMA / checkpoint code zips state_dicts by position
    m.load_state_dict(sd)
    assert float(m.film1.gamma.weight.abs().sum()) > 0
    assert float(FiLMResNet50Policy(COND).film3.beta.weight.abs().sum()) == 0.0  # zero init like the reference (:14-17)
    with pytest.raises(NotImplementedError):
        m.train()(torch.zeros(1, 3, 64, 64), torch.zeros(1, COND))
    g = np.load(GOLD / "film_resnet50.npz")
    ora = R.FiLMResNet50Oracle(COND).eval()
    ora.load_state_dict(sd)
    with torch.no_grad():
        img, cond = R.synthetic_inputs(3, 64, COND)
        assert rel_l2(ora(img, cond).numpy(), g["s64_out"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n,size", [("s64", 3, 64), ("s112", 2, 112)])
def test_engine_film_resnet_matches_the_reference_forward(tag, n, size):
    g = np.load(GOLD / "film_resnet50.npz")
    sd = R.synthetic_state_dict(COND)
    m = FiLMResNet50Policy(COND, max_images=4).cuda().eval()
    m.load_state_dict(sd)
    img, cond = R.synthetic_inputs(n, size, COND)
    y = m(img.cuda(), cond.cuda()).cpu().numpy()
    assert np.isfinite(y).all() and y.shape == (n, 2048)
    e, gap = rel_l2(y, g[f"{tag}_out"]), rel_l2(g[f"{tag}_out_autocast_bf16"], g[f"{tag}_out"])
    print(f"\nFiLM-ResNet-50 {tag}: engine vs reference fp32 {e:.3e}; reference bf16-autocast vs its fp32 {gap:.3e}")
    # bf16 tensor-core operands and bf16 activations between 53 convolutions against an fp32 reference: bounded by twice
    # the reference's own bf16-autocast distance
    assert e < 2 * gap
    # a (B, 1, cond) goal and repeated calls; FiLM really acts (other goal -> other tokens)
    y2 = m(img.cuda(), cond.cuda()[:, None, :]).cpu().numpy()
    assert np.array_equal(y, y2)
    y3 = m(img.cuda(), torch.zeros_like(cond).cuda()).cpu().numpy()
    assert rel_l2(y3, y) > 1e-2
    # a weight update is picked up (fingerprint of parameters and buffers)
    with torch.no_grad():
        m.film4.beta.bias.add_(1.0)
    y4 = m(img.cuda(), cond.cuda()).cpu().numpy()
    np.testing.assert_allclose(y4 - y, 1.0, atol=0.1)  # features are O(5): one bf16 ulp of the stored activation is 0.03


@pytest.mark.gpu
def test_engine_film_resnet_batch_of_camera_frames_is_sample_independent():
    """Size-independent property at a CALVIN-like shape: encoding 16 frames at once equals encoding them in two halves."""
    sd = R.synthetic_state_dict(COND)
    m = FiLMResNet50Policy(COND, max_images=16).cuda().eval()
    m.load_state_dict(sd)
    img, cond = R.synthetic_inputs(16, 224, COND, seed=9)
    img, cond = img.cuda(), cond.cuda()
    full = m(img, cond)
    half = torch.cat([m(img[:8], cond[:8]), m(img[8:], cond[8:])])
    assert torch.isfinite(full).all() and torch.equal(full, half)
