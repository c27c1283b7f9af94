"""Known-answer vectors for the training noise-level densities (reference mode/models/edm_diffusion/utils.py:154-203,
selected by MoDEAgent.make_sample_density, mode_agent.py:692-731). Build container only:

    python tests/golden/make_density_goldens.py

Every density is drawn with the REFERENCE's function after torch.manual_seed(1234) on CPU; the test replays the same
seed through mode_diffusion_policy_b200.utils and expects identical bits (same torch ops in the same order)."""
import math
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_goldens  # noqa: F401,E402  (stubs hydra / torchsde / ... and puts /root/reference on sys.path)
from mode.models.edm_diffusion import utils as RU  # noqa: E402

OUT = Path(__file__).resolve().parent
CASES = {
    "lognormal": lambda: RU.rand_log_normal((64,), loc=-1.2, scale=1.2),
    "loglogistic": lambda: RU.rand_log_logistic((64,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0),
    "loguniform": lambda: RU.rand_log_uniform((64,), min_value=0.001, max_value=80.0),
    "uniform": lambda: RU.rand_uniform((64,), min_value=0.001, max_value=80.0),
    "v-diffusion": lambda: RU.rand_v_diffusion((64,), sigma_data=0.5, min_value=0.001, max_value=80.0),
    "split-lognormal": lambda: RU.rand_split_log_normal((64,), loc=-1.2, scale_1=0.8, scale_2=1.4),
    "discrete": lambda: RU.rand_discrete((64,), values=torch.linspace(0.001, 80.0, 1000)),
}

if __name__ == "__main__":
    out = {}
    for name, fn in CASES.items():
        torch.manual_seed(1234)
        out[name] = fn().numpy()
    np.savez_compressed(OUT / "sample_densities.npz", **out)
    print({k: (float(v.min()), float(v.max())) for k, v in out.items()})
