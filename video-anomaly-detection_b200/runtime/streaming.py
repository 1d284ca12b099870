"""Overlap-aware streaming scoring of one video.

The reference's single-video paths (`generate_video_output`, evaluate_video.py:322-326,350-352; `on_analyze_video`,
main.py:334-338,351) slide a T-frame window with stride < T over the stream and run the whole model on every window:
each frame is re-encoded T/stride times (16x at stride 1) and each window is forwarded up to three times.  Here the
per-frame encoder features are computed once and cached; a window only costs the ConvLSTM (whose state restarts at
zero for every window, video_autoencoder.py:144-145), the decoder and the fused scoring pass, and one pass returns every
output (score, min/max, heat map, optionally the reconstruction).

Window results are bit-identical to `model.score_all(window)`: the encoder treats frames independently (eval-mode
BatchNorm), so a cached feature is the feature the full forward would compute.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from models import _engine as eng


class StreamingVideoScorer:
    """Feed frames as they arrive; get one `ScoreOutputs` per completed window.

        scorer = StreamingVideoScorer(model, seq_len=16, stride=8)
        for chunk in decoder:                      # fp32 [n, 3, H, W] in [-1, 1] on the model's device
            for win_start, out in scorer.push(chunk):
                ...                                # out.score [T], out.minmax [T, 2], out.heat [T, H, W]

    Windows start at stream indices 0, stride, 2*stride, ... exactly like the reference datasets'
    `range(0, N - T + 1, stride)` (utils/video_dataset.py:371), for any stride (also stride > seq_len: the frames
    between two windows are skipped, never encoded).
    """

    def __init__(self, model, seq_len: int = 16, stride: int = 8, want_recon: bool = False, want_heat: bool = True):
        if seq_len <= 0 or stride <= 0:
            raise ValueError("seq_len and stride must be positive")
        self.model, self.T, self.stride = model, seq_len, stride
        self.want_recon, self.want_heat = want_recon, want_heat
        self._frames: Optional[torch.Tensor] = None    # fp32 [n, 3, H, W]: frames from the next window's start on
        self._latents: Optional[torch.Tensor] = None   # bf16 [n, h, w, C]: their encoder features
        self._start = 0                                # stream index of the next window's first frame
        self._skip = 0                                 # frames still to be dropped before that window starts
        self.frames_encoded = 0                        # (a window-by-window caller would encode T per window)

    def push(self, frames: torch.Tensor) -> List[tuple]:
        if frames.dim() != 4 or frames.shape[1] != 3:
            raise RuntimeError(f"expected frames [n, 3, H, W], got {tuple(frames.shape)}")
        engine: eng.VideoEngine = self.model._get_engine(frames.device)
        if self._skip:  # stride > seq_len: frames between two windows belong to no window
            drop = min(self._skip, frames.shape[0])
            frames = frames[drop:]
            self._skip -= drop
        if frames.shape[0] == 0:
            return []
        frames = frames.float().contiguous()
        z, _ = engine.encode(frames)                   # every frame is encoded exactly once
        self.frames_encoded += frames.shape[0]
        self._frames = frames if self._frames is None else torch.cat([self._frames, frames], 0)
        self._latents = z if self._latents is None else torch.cat([self._latents, z], 0)
        out = []
        while self._frames.shape[0] >= self.T:
            out.append((self._start, self._score_window(engine)))
            drop = min(self.stride, self._frames.shape[0])
            self._frames, self._latents = self._frames[drop:], self._latents[drop:]
            self._skip = self.stride - drop            # remainder of the stride comes out of the next push(es)
            self._start += self.stride
        return out

    def _score_window(self, engine: "eng.VideoEngine"):
        T = self.T
        x = self._frames[:T].unsqueeze(0)
        lat = self._latents[:T].unsqueeze(0)
        return engine.score_latents(lat, x, self.want_recon, self.want_heat)
