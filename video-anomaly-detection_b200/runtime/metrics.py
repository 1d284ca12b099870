"""SSIM / combined loss of the reference (utils/losses.py) as alternative anomaly scores, on the GPU.

The reference only uses them for training; as scores they need the reconstruction, so pair them with
`model.score_all(x, want_recon=True)`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from models import _native as nat


def ssim_loss(pred: torch.Tensor, target: torch.Tensor, want_map: bool = False
              ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Per-frame `1 - mean(SSIM map)` of fp32 [..., 3, H, W] CUDA tensors (SSIMLoss.forward, utils/losses.py:51-93; its
    batch value is `.mean()` of the returned vector).  Returns (loss [...], ssim_map [..., 3, H, W] or None)."""
    if not (pred.is_cuda and target.is_cuda) or pred.shape != target.shape or pred.shape[-3] != 3:
        raise RuntimeError("ssim_loss expects two CUDA tensors of the same shape [..., 3, H, W]")
    p, t = pred.float().contiguous(), target.float().contiguous()
    lead, (H, W) = p.shape[:-3], p.shape[-2:]
    n = 1
    for d in lead:
        n *= int(d)
    lib = nat.load()
    loss = torch.empty(n, dtype=torch.float32, device=p.device)
    smap = torch.empty_like(p) if want_map else None
    scratch = torch.empty(lib.vad_ssim_scratch_bytes(n, H, W), dtype=torch.uint8, device=p.device)
    with torch.cuda.device(p.device):  # the library works on the current device: make it the tensors' one
        nat.check(lib.vad_ssim_loss(p.data_ptr(), t.data_ptr(), n, H, W, loss.data_ptr(), nat.ptr(smap),
                                    scratch.data_ptr(), nat.stream_ptr(p.device)), "vad_ssim_loss")
    return loss.view(lead) if lead else loss.view(()), smap


def combined_loss(pred: torch.Tensor, target: torch.Tensor, alpha: float = 0.5) -> torch.Tensor:
    """CombinedLoss.forward (utils/losses.py:116-121) per frame: (1-alpha) * MSE + alpha * (1 - SSIM)."""
    loss, _ = ssim_loss(pred, target)
    mse = ((pred.float() - target.float()) ** 2).flatten(-3).mean(-1)
    return (1 - alpha) * mse + alpha * loss
