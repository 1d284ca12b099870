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
    with torch.cuda.device(x.device):  # the library works on the current device: make it the tensors' one
        nat.check(nat.load().vad_u8_hwc_to_f32_nchw(x.data_ptr(), n, H, W, out.data_ptr(), nat.stream_ptr(x.device)),
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
    with torch.cuda.device(x.device):
        nat.check(nat.load().vad_f32_nchw_to_u8_hwc(x.data_ptr(), n, H, W, out.data_ptr(), nat.stream_ptr(x.device)),
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
    with torch.cuda.device(heat.device):
        nat.check(nat.load().vad_heatmap_jet_rgb(heat.data_ptr(), minmax.data_ptr(), F, H, W, out.data_ptr(),
                                                 nat.stream_ptr(heat.device)), "vad_heatmap_jet_rgb")
    return out


def compose_panels(x: torch.Tensor, recon: torch.Tensor, heat: torch.Tensor, minmax: torch.Tensor) -> torch.Tensor:
    """The reference's side-by-side frames, `np.hstack([denormalize(frame), denormalize(recon), create_heatmap(err)])`
    (evaluate_video.py:279-286, 355-364), for all frames of one `score_all(x, want_recon=True)` call at once:
    x, recon fp32 [F,3,H,W], heat fp32 [F,H,W], minmax [F,2] -> uint8 RGB [F, H, 3W, 3] (byte-exact when the frames have
    the size create_heatmap resizes to; the score bar / text / VideoWriter stay with cv2 on the host)."""
    if not (x.is_cuda and recon.is_cuda and heat.is_cuda) or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError("compose_panels expects CUDA fp32 tensors x, recon [F,3,H,W], heat [F,H,W], minmax [F,2]")
    x, recon, heat, minmax = x.contiguous(), recon.contiguous().float(), heat.contiguous().float(), minmax.contiguous().float()
    F, _, H, W = x.shape
    if tuple(recon.shape) != (F, 3, H, W) or tuple(heat.shape) != (F, H, W) or tuple(minmax.shape) != (F, 2):
        raise RuntimeError("compose_panels: shapes of x, recon, heat, minmax do not belong together")
    out = torch.empty(F, H, 3 * W, 3, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        nat.check(nat.load().vad_compose_panel(x.data_ptr(), recon.data_ptr(), heat.data_ptr(), minmax.data_ptr(), F, H, W,
                                               out.data_ptr(), nat.stream_ptr(x.device)), "vad_compose_panel")
    return out
