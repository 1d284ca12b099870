"""Drop-in `models.video_autoencoder` for the anomaly-scoring hot path, running on libvad_b200 (sm_100a).

Mirrors the reference `models/video_autoencoder.py`: ConvLSTMCell (:24-91), ConvLSTM (:94-179), VideoEncoder
(:182-231), VideoDecoder (:234-276), VideoAutoencoder (:279-384) — same constructor arguments, sub-module tree and
`state_dict` keys, so `evaluate_video.py` loads checkpoints and scores unchanged.  The nn sub-modules only hold
parameters; compute goes through the fused kernels.  Inference only, CUDA only.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _prepare as prep
from ._engine import CellEngine, ScoreOutputs, VideoEngine
from .autoencoder import _Prepared, _xavier_like_reference


class ConvLSTMCell(_Prepared):
    """One ConvLSTM cell (reference models/video_autoencoder.py:24-91): gate convolution over cat[x, h]
    (input_dim + hidden_dim -> 4 * hidden_dim, i/f/g/o order), sigma/sigma/tanh/sigma, c' = f*c + i*g, h' = o*tanh(c')."""

    def __init__(self, input_dim: int, hidden_dim: int, kernel_size: int = 3):
        super().__init__()
        if kernel_size != 3:
            raise ValueError("vad_b200 ConvLSTM kernels implement kernel_size=3 (the only size the reference uses)")
        if input_dim % 32 or hidden_dim % 32:
            raise ValueError("ConvLSTM input_dim and hidden_dim must be multiples of 32 for the tcgen05 tile shapes")
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.conv = nn.Conv2d(input_dim + hidden_dim, 4 * hidden_dim, kernel_size=kernel_size,
                              padding=kernel_size // 2, bias=True)

    def _build(self, sd):
        return CellEngine(prep.prepare_lstm_cell(sd, prefix=""))

    @torch.no_grad()
    def forward(self, x: torch.Tensor, hidden_state: Tuple[torch.Tensor, torch.Tensor]
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        """x fp32 [B,input_dim,H,W], (h, c) fp32 [B,hidden_dim,H,W] -> (h_next, c_next) (:54-85): one gate GEMM with
        the state update in its epilogue."""
        h_cur, c_cur = hidden_state
        return self._get_engine(x.device).step(x, h_cur, c_cur)

    def init_hidden(self, batch_size: int, height: int, width: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        z = torch.zeros(batch_size, self.hidden_dim, height, width, device=device)
        return z, z.clone()


class ConvLSTM(_Prepared):
    """Stack of ConvLSTM cells (reference :94-179); `forward` takes fp32 [B,T,C,H,W] and returns the last layer's hidden
    sequence and final (h, c)."""

    def __init__(self, input_dim: int, hidden_dims, kernel_size: int = 3, num_layers: int = 2,
                 batch_first: bool = True, return_all_layers: bool = False):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dims = list(hidden_dims) if isinstance(hidden_dims, (list, tuple)) else [hidden_dims] * num_layers
        self.num_layers = len(self.hidden_dims)
        self.batch_first = batch_first
        self.return_all_layers = return_all_layers
        self.cells = nn.ModuleList(
            ConvLSTMCell(input_dim if i == 0 else self.hidden_dims[i - 1], self.hidden_dims[i], kernel_size)
            for i in range(self.num_layers))

    def _build(self, sd):
        return VideoEngine(prep.prepare_convlstm(sd, prefix=""))

    def _init_hidden(self, batch_size: int, height: int, width: int, device) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        return [cell.init_hidden(batch_size, height, width, device) for cell in self.cells]

    @torch.no_grad()
    def forward(self, x: torch.Tensor, hidden_state=None):
        if hidden_state is not None:
            raise RuntimeError("vad_b200 ConvLSTM.forward always starts from the zero state, as every reference caller "
                               "does (video_autoencoder.py:144-145, :343); step a ConvLSTMCell for custom states")
        if self.return_all_layers:
            raise RuntimeError("return_all_layers=True is not on the scoring path")
        if not self.batch_first:
            x = x.permute(1, 0, 2, 3, 4)
        out, c_last = self._get_engine(x.device).convlstm(x)
        return out, (out[:, -1], c_last)


class VideoEncoder(_Prepared):
    """4 x [conv3x3-BN-LeakyReLU-maxpool2]: 3 -> 32 -> 64 -> 128 -> latent_dim (reference :182-231); 5-D input folds T
    into the batch."""

    def __init__(self, in_channels: int = 3, latent_dim: int = 128):
        super().__init__()
        if in_channels != 3:
            raise ValueError("vad_b200 kernels are specialised for 3-channel input (as every reference dataset is)")
        if latent_dim % 32:
            raise ValueError("latent_dim must be a multiple of 32 for the tcgen05 tile shapes")
        layers = []
        cin = in_channels
        for cout in (32, 64, 128, latent_dim):
            layers += [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout),
                       nn.LeakyReLU(0.2, inplace=True), nn.MaxPool2d(2, 2)]
            cin = cout
        self.encoder = nn.Sequential(*layers)

    def _build(self, sd):
        return VideoEngine(prep.prepare_video_encoder(sd, prefix="encoder."))

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() not in (4, 5):
            raise RuntimeError(f"expected a 4-D or 5-D input, got shape {tuple(x.shape)}")
        x4 = x.reshape(-1, *x.shape[-3:])
        _, out = self._get_engine(x.device).encode(x4, want_f32=True, want_bf16=False)
        return out.view(x.shape[0], x.shape[1], *out.shape[1:]) if x.dim() == 5 else out


class VideoDecoder(_Prepared):
    """4 x convT k2 s2: latent_dim -> 128 -> 64 -> 32 -> out_channels (BN+ReLU on the first three, Tanh last; reference
    :234-276); accepts [B,C,h,w] frames or [B,T,C,h,w] sequences."""

    def __init__(self, out_channels: int = 3, latent_dim: int = 128):
        super().__init__()
        if out_channels != 3:
            raise ValueError("vad_b200 kernels are specialised for 3-channel output (as every reference dataset is)")
        if latent_dim % 32:
            raise ValueError("latent_dim must be a multiple of 32 for the tcgen05 tile shapes")
        layers = []
        cin = latent_dim
        for cout in (128, 64, 32):
            layers += [nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2), nn.BatchNorm2d(cout),
                       nn.ReLU(inplace=True)]
            cin = cout
        layers += [nn.ConvTranspose2d(cin, out_channels, kernel_size=2, stride=2), nn.Tanh()]
        self.decoder = nn.Sequential(*layers)

    def _build(self, sd):
        return VideoEngine(prep.prepare_video_decoder(sd, prefix="decoder."))

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() not in (4, 5):
            raise RuntimeError(f"expected a 4-D or 5-D input, got shape {tuple(x.shape)}")
        z4 = x.reshape(-1, *x.shape[-3:])
        out = self._get_engine(x.device).decode(z4)
        return out.view(x.shape[0], x.shape[1], *out.shape[1:]) if x.dim() == 5 else out


class VideoAutoencoder(_Prepared):
    """Encoder -> ConvLSTM -> (1x1 proj if hidden != latent) -> decoder, scored by fused sm_100a kernels.

    `forward`, `get_reconstruction_error(x, per_frame, per_pixel)` follow the reference (:329-384) including
    "per_pixel wins over per_frame" (:373-380).  `score_all` returns every output of one forward.
    """

    def __init__(self, in_channels: int = 3, latent_dim: int = 128, lstm_hidden_dim: int = 128,
                 lstm_num_layers: int = 2):
        super().__init__()
        self.encoder = VideoEncoder(in_channels, latent_dim)
        self.convlstm = ConvLSTM(input_dim=latent_dim, hidden_dims=[lstm_hidden_dim] * lstm_num_layers, kernel_size=3,
                                 num_layers=lstm_num_layers, batch_first=True, return_all_layers=False)
        self.proj = nn.Conv2d(lstm_hidden_dim, latent_dim, kernel_size=1) if lstm_hidden_dim != latent_dim \
            else nn.Identity()
        self.decoder = VideoDecoder(in_channels, latent_dim)
        _xavier_like_reference(self)

    def _build(self, sd):
        return VideoEngine(prep.prepare_video(sd))

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = self._get_engine(x.device).run(x, want_recon=True, want_heat=False)
        return out.recon.view(x.shape)

    @torch.no_grad()
    def get_reconstruction_error(self, x: torch.Tensor, per_frame: bool = False, per_pixel: bool = False
                                 ) -> torch.Tensor:
        b, t = x.shape[0], x.shape[1]
        out = self._get_engine(x.device).run(x, want_recon=False, want_heat=per_pixel)
        if per_pixel:
            return out.heat.view(b, t, 1, *x.shape[-2:])  # [B, T, 1, H, W]
        if per_frame:
            return out.score.view(b, t)  # [B, T]
        return out.score.view(b, t).mean(dim=1)  # [B]: equal-sized frames, so mean of frame means

    @torch.no_grad()
    def score_all(self, x: torch.Tensor, want_recon: bool = True, want_heat: bool = True) -> ScoreOutputs:
        return self._get_engine(x.device).run(x, want_recon=want_recon, want_heat=want_heat)

    @torch.no_grad()
    def score_frames(self, frames_u8: torch.Tensor, want_recon: bool = False, want_heat: bool = False,
                     want_heat_u8: bool = True) -> ScoreOutputs:
        """Score decoded clips directly: uint8 RGB [B,T,H,W,3] on the GPU (normalised on the device like
        utils/video_dataset.py:62-66 does on the host); `heat_u8` [B*T,H,W] is create_heatmap's uint8 normalisation of
        each frame's error map (evaluate_video.py:56-57).  Results equal `score_all(normalised fp32 clips)` bit for bit."""
        return self._get_engine(frames_u8.device).run_u8(frames_u8, want_recon, want_heat, want_heat_u8)
