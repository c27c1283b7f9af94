"""FiLM-ResNet-50 token producer throughput: N camera frames at 224x224 through the engine's
GEMM-based forward (`mode_resnet_forward`). ResNet-50 is 4.09 GMAC = 8.18 GFLOP per 224x224 image."""
import json
import sys

import torch

sys.path.insert(0, ".")
from oracle import film_resnet_ref as R  # noqa: E402  (synthetic weights / inputs only)
from mode_diffusion_policy_b200.perceptual_encoders.pretrained_resnets import FiLMResNet50Policy  # noqa: E402

COND = 512
sd = R.synthetic_state_dict(COND)
for n in [int(a) for a in sys.argv[1:]] or [1, 16, 256]:
    m = FiLMResNet50Policy(COND, max_images=n).cuda().eval()
    m.load_state_dict(sd)
    img, cond = R.synthetic_inputs(n, 224, COND, seed=3)
    img, cond = img.cuda(), cond.cuda()
    for _ in range(3):
        m(img, cond)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10 if n >= 64 else 50
    e0.record()
    for _ in range(reps):
        y = m(img, cond)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"images": n, "ms_per_forward": round(ms, 3), "images_per_s": round(n / ms * 1e3, 1),
                      "tflops": round(8.18e9 * n / (ms * 1e-3) / 1e12, 1), "finite": bool(torch.isfinite(y).all())}), flush=True)
    m.close()
