"""Drop-in `models.autoencoder` for the anomaly-scoring hot path, running on libvad_b200 (sm_100a).

Same class names, constructor arguments, sub-module tree / `state_dict` keys, and method signatures as the
reference `models/autoencoder.py` (Encoder :24-86, Decoder :89-146, ConvAutoencoder :149-221), so
`evaluate.py`'s `load_model` / `compute_auroc` run unchanged.  The `nn.Conv2d` / `nn.BatchNorm2d` / ... sub-modules
are parameter containers only: `forward` never calls them.  Inference only (eval-mode BatchNorm, no autograd),
CUDA only — anything else raises.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _prepare as prep
from ._engine import ImageEngine, ScoreOutputs

_ENC_WIDTHS = (32, 64, 128)  # enc4 width is latent_dim
_DEC_WIDTHS = (128, 64, 32, 32)


def _state_signature(module: nn.Module):
    """Cheap change detector for the prepared-weight cache: (storage pointer, in-place version) of every tensor."""
    return tuple((t.data_ptr(), t._version, t.device) for t in module.state_dict(keep_vars=True).values())


def _xavier_like_reference(module: nn.Module) -> None:
    # same initialisation walk as the reference's _init_weights (autoencoder.py:170-179) so that a given
    # torch.manual_seed produces bit-identical parameters
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.xavier_normal_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


def _refuse_training(module: nn.Module) -> None:
    if module.training:
        raise RuntimeError(
            "vad_b200 implements the scoring path only (eval-mode BatchNorm folded into the convolutions); "
            "call model.eval() first — training stays with the reference implementation")


def _refuse_cpu(device: torch.device) -> None:
    if torch.device(device).type != "cuda":
        raise RuntimeError("vad_b200 scoring path is CUDA-only (sm_100a kernels, no CPU fallback); "
                           f"got a tensor on {device}")


class Encoder(nn.Module):
    """4 x [conv3x3-BN-LeakyReLU x2, maxpool2]: 3 -> 32 -> 64 -> 128 -> latent_dim, spatial / 16."""

    def __init__(self, in_channels: int = 3, latent_dim: int = 256):
        super().__init__()
        widths = _ENC_WIDTHS + (latent_dim,)
        cin = in_channels
        for i, cout in enumerate(widths, start=1):
            layers = []
            for _ in range(2):
                layers += [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout),
                           nn.LeakyReLU(0.2, inplace=True)]
                cin = cout
            layers.append(nn.MaxPool2d(2, 2))
            setattr(self, f"enc{i}", nn.Sequential(*layers))
        self._owner: Optional["ConvAutoencoder"] = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._owner is None:
            raise RuntimeError("Encoder runs through its ConvAutoencoder (weights are prepared per model)")
        return self._owner.get_latent(x)


class Decoder(nn.Module):
    """4 x [convT k2 s2-BN-ReLU, conv3x3-BN-ReLU]; the last conv goes to `out_channels` and ends in Tanh."""

    def __init__(self, out_channels: int = 3, latent_dim: int = 256):
        super().__init__()
        cin = latent_dim
        for i, cout in enumerate(_DEC_WIDTHS, start=1):
            layers = [nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2), nn.BatchNorm2d(cout),
                      nn.ReLU(inplace=True)]
            if i < 4:
                layers += [nn.Conv2d(cout, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout),
                           nn.ReLU(inplace=True)]
            else:
                layers += [nn.Conv2d(cout, out_channels, kernel_size=3, padding=1), nn.Tanh()]
            setattr(self, f"dec{i}", nn.Sequential(*layers))
            cin = cout


class ConvAutoencoder(nn.Module):
    """Image autoencoder whose scoring calls run as fused sm_100a kernels.

    API parity: `forward`, `get_latent`, `get_reconstruction_error(x, per_pixel=False)` behave like the reference
    (autoencoder.py:181-221).  `score_all` is the single-pass extension (recon + map + score from one forward).
    """

    def __init__(self, in_channels: int = 3, latent_dim: int = 256):
        super().__init__()
        if in_channels != 3:
            raise ValueError("vad_b200 kernels are specialised for 3-channel input (as every reference dataset is)")
        if latent_dim % 32 != 0:
            raise ValueError("latent_dim must be a multiple of 32 for the tcgen05 tile shapes")
        self.encoder = Encoder(in_channels, latent_dim)
        self.decoder = Decoder(in_channels, latent_dim)
        _xavier_like_reference(self)
        object.__setattr__(self.encoder, "_owner", self)
        self._engine: Optional[ImageEngine] = None
        self._engine_sig = None

    # ---- prepared-weight cache -----------------------------------------------------------------------------
    def _get_engine(self, device: torch.device) -> ImageEngine:
        _refuse_training(self)
        _refuse_cpu(device)
        sig = _state_signature(self)
        if self._engine is None or sig != self._engine_sig:
            sd: Dict[str, torch.Tensor] = {k: v.detach() for k, v in self.state_dict().items()}
            for k, v in sd.items():
                if v.device != device:
                    raise RuntimeError(f"model parameter {k} lives on {v.device} but the input is on {device}")
            self._engine = ImageEngine(prep.prepare_image(sd))
            self._engine_sig = sig
        return self._engine

    # ---- reference API -------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._get_engine(x.device).run(x, want_recon=True, want_heat=False).recon

    @torch.no_grad()
    def get_latent(self, x: torch.Tensor) -> torch.Tensor:
        return self._get_engine(x.device).latent(x)

    @torch.no_grad()
    def get_reconstruction_error(self, x: torch.Tensor, per_pixel: bool = False) -> torch.Tensor:
        out = self._get_engine(x.device).run(x, want_recon=False, want_heat=per_pixel)
        if per_pixel:
            return out.heat.unsqueeze(1)  # [B, 1, H, W]
        return out.score  # [B]

    # ---- single-pass extension --------------------------------------------------------------
    @torch.no_grad()
    def score_all(self, x: torch.Tensor, want_recon: bool = True, want_heat: bool = True) -> ScoreOutputs:
        """One forward producing score [B], min/max [B,2] and optionally the heat map [B,H,W] and recon."""
        return self._get_engine(x.device).run(x, want_recon=want_recon, want_heat=want_heat)
