"""Frame I/O either side of the scoring path, on the GPU.

The reference does these steps per frame on the host: torchvision `ToTensor` + `Normalize(.5, .5)` on decoded uint8
frames (utils/dataset.py:65-70, utils/video_dataset.py:62-66,356-360), `denormalize` and `create_heatmap`
(evaluate_video.py:40-66: per-frame min/max normalisation, uint8, cv2 JET colour map).  Shipping uint8 frames over PCIe
and normalising on the device moves a quarter of the bytes of the fp32 tensors the reference callers upload.
"""
from __future__ import annotations

import torch

from models import _native as nat


def normalize_u8(frames_u8: torch.Tensor) -> torch.Tensor:
    """uint8 RGB frames [..., H, W, 3] on the GPU -> fp32 [..., 3, H, W] in [-1, 1] (ToTensor + Normalize(.5, .5))."""
    if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8 or frames_u8.shape[-1] != 3:
        raise RuntimeError("normalize_u8 expects a CUDA uint8 tensor [..., H, W, 3]")
    x = frames_u8.contiguous()
    lead, (H, W) = x.shape[:-3], x.shape[-3:-1]
    n = int(torch.tensor(lead).prod()) if len(lead) else 1
    out = torch.empty(*lead, 3, H, W, dtype=torch.float32, device=x.device)
    nat.check(nat.load().vad_u8_hwc_to_f32_nchw(x.data_ptr(), n, H, W, out.data_ptr(), nat.stream_ptr()),
              "vad_u8_hwc_to_f32_nchw")
    return out


def denormalize_u8(x: torch.Tensor) -> torch.Tensor:
    """fp32 [..., 3, H, W] in [-1, 1] -> uint8 [..., H, W, 3] (reference `denormalize`)."""
    if not x.is_cuda or x.dtype != torch.float32 or x.shape[-3] != 3:
        raise RuntimeError("denormalize_u8 expects a CUDA fp32 tensor [..., 3, H, W]")
    x = x.contiguous()
    lead, (H, W) = x.shape[:-3], x.shape[-2:]
    n = int(torch.tensor(lead).prod()) if len(lead) else 1
    out = torch.empty(*lead, H, W, 3, dtype=torch.uint8, device=x.device)
    nat.check(nat.load().vad_f32_nchw_to_u8_hwc(x.data_ptr(), n, H, W, out.data_ptr(), nat.stream_ptr()),
              "vad_f32_nchw_to_u8_hwc")
    return out


def render_heatmap(heat: torch.Tensor, minmax: torch.Tensor) -> torch.Tensor:
    """per-pixel error maps fp32 [F, H, W] + per-frame (min, max) [F, 2] (both from `score_all`) -> JET-coloured RGB
    uint8 [F, H, W, 3] (reference `create_heatmap` without the optional resize)."""
    if not heat.is_cuda or heat.dtype != torch.float32 or heat.dim() != 3:
        raise RuntimeError("render_heatmap expects a CUDA fp32 tensor [F, H, W]")
    heat, minmax = heat.contiguous(), minmax.contiguous().float()
    F, H, W = heat.shape
    out = torch.empty(F, H, W, 3, dtype=torch.uint8, device=heat.device)
    nat.check(nat.load().vad_heatmap_jet_rgb(heat.data_ptr(), minmax.data_ptr(), F, H, W, out.data_ptr(),
                                             nat.stream_ptr()), "vad_heatmap_jet_rgb")
    return out
