"""Synthetic workloads of the named shapes (SURVEY §8d recipes), shared by bench.py and the parity tests.

There are no datasets or checkpoints offline, so the BASELINE configs run on generated inputs: per-frame amplitude
a ~ U[0.3, 1] (spreads the scores ~10x so rankings mean something), low-pass noise (8x8 blocks, bilinear
up-sampling) + fine noise, clamped to [-1, 1]; "anomalous" frames additionally carry a +-0.5 patch of 16-48 px (scaled
with the frame size).  Everything derives from `seed`, frame by frame, so any sharding of a job sees identical data.
"""
from __future__ import annotations

from typing import Tuple

import torch


def synth_frames(n: int, H: int, W: int, device, seed: int = 1234, anomaly_fraction: float = 0.0
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (frames fp32 [n, 3, H, W] in [-1, 1] on `device`, labels int64 [n] on the CPU; 1 = carries an anomaly patch)."""
    g = torch.Generator(device=device).manual_seed(seed)
    amp = 0.3 + 0.7 * torch.rand(n, 1, 1, 1, generator=g, device=device)
    coarse = torch.rand(n, 3, max(H // 8, 1), max(W // 8, 1), generator=g, device=device) * 2 - 1
    x = amp * torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)
    x = x + 0.05 * torch.randn(n, 3, H, W, generator=g, device=device)
    labels = torch.zeros(n, dtype=torch.int64)
    if anomaly_fraction > 0:
        gc = torch.Generator().manual_seed(seed + 7919)           # patch geometry: small, drawn on the CPU
        labels = (torch.rand(n, generator=gc) < anomaly_fraction).to(torch.int64)
        scale = max(min(H, W) // 256, 1)
        size = torch.randint(16 * scale, 48 * scale + 1, (n,), generator=gc)
        fy = torch.rand(n, generator=gc)
        fx = torch.rand(n, generator=gc)
        sign = torch.where(torch.rand(n, generator=gc) < 0.5, -0.5, 0.5)
        for i in torch.nonzero(labels).flatten().tolist():
            s = min(int(size[i]), H, W)
            y0, x0 = int(fy[i] * (H - s)), int(fx[i] * (W - s))
            x[i, :, y0:y0 + s, x0:x0 + s] += float(sign[i])
    return x.clamp_(-1, 1), labels


def synth_clips(B: int, T: int, H: int, W: int, device, seed: int = 1234, anomaly_fraction: float = 0.0
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (clips fp32 [B, T, 3, H, W], per-frame labels int64 [B, T])."""
    x, labels = synth_frames(B * T, H, W, device, seed, anomaly_fraction)
    return x.view(B, T, 3, H, W), labels.view(B, T)
