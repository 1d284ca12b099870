"""tests/golden/golden_ssim.npz: outputs of the UNMODIFIED reference SSIMLoss / CombinedLoss (utils/losses.py) on seeded
inputs (build container only, needs /root/reference).  Inputs are regenerated from their seeds by the tests."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
from utils.losses import CombinedLoss, SSIMLoss  # noqa: E402  (reference)


def pair(seed, b, h, w, noise):
    g = torch.Generator().manual_seed(seed)
    t = (torch.rand(b, 3, h, w, generator=g) * 2 - 1)
    t = torch.nn.functional.avg_pool2d(t, 5, 1, 2)                 # some spatial structure
    p = (t + noise * torch.randn(b, 3, h, w, generator=g)).clamp(-1, 1)
    return p, t


out = {}
for name, (seed, b, h, w, noise) in {"a": (11, 3, 40, 56, 0.05), "b": (12, 2, 64, 64, 0.3), "c": (13, 1, 16, 16, 0.0)}.items():
    p, t = pair(seed, b, h, w, noise)
    with torch.no_grad():
        out[f"{name}_ssim_loss"] = SSIMLoss()(p, t).numpy()
        out[f"{name}_combined"] = CombinedLoss(alpha=0.5)(p, t).numpy()
        out[f"{name}_ssim_loss_per_frame"] = np.array([float(SSIMLoss()(p[i:i + 1], t[i:i + 1])) for i in range(b)],
                                                      dtype=np.float32)
    out[f"{name}_spec"] = np.array([seed, b, h, w, noise], dtype=np.float64)
np.savez(os.path.join(HERE, "golden_ssim.npz"), **out)
print({k: v for k, v in out.items() if "spec" not in k})
