"""Drop-in `models.autoencoder` for the anomaly-scoring hot path, running on libvad_b200 (sm_100a).

Same class names, constructor arguments, sub-module tree / `state_dict` keys, and method signatures as the
reference `models/autoencoder.py` (Encoder :24-86, Decoder :89-146, ConvAutoencoder :149-221), so
`evaluate.py`'s `load_model` / `compute_auroc` run unchanged.  The `nn.Conv2d` / `nn.BatchNorm2d` / ... sub-modules
are parameter containers only: `forward` never calls them.  Inference only (eval-mode BatchNorm, no autograd),
CUDA only — anything else raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _prepare as prep
from ._engine import ImageEngine, ScoreOutputs

_ENC_WIDTHS = (32, 64, 128)  # enc4 width is latent_dim
_DEC_WIDTHS = (128, 64, 32, 32)


class _PreparedCache:
    """Prepared-weight cache of one module: the engine built from its parameters plus a cheap change detector.

    Per call only (in-place version, storage pointer) of the tensors seen at build time are compared (~20 us for the
    image model; walking `state_dict()` costs 300 us).  `load_state_dict` (a post hook) and `.to()` / `.cuda()` /
    `.float()` (`_apply`) drop the cache outright, which also covers parameters being REPLACED rather than updated."""

    def __init__(self) -> None:
        self.engine = None
        self.tensors: List[torch.Tensor] = []
        self.sig: Tuple = ()

    @staticmethod
    def _signature(tensors) -> Tuple:
        return tuple((t._version, t.data_ptr()) for t in tensors)

    def get(self, module: nn.Module, device: torch.device, build):
        _refuse_training(module)
        _refuse_cpu(device)
        if self.engine is not None and self._signature(self.tensors) == self.sig:
            if self.engine.device != device:
                raise RuntimeError(f"model parameters live on {self.engine.device} but the input is on {device}")
            return self.engine
        sd: Dict[str, torch.Tensor] = {k: v.detach() for k, v in module.state_dict().items()}
        for k, v in sd.items():
            if v.device != device:
                raise RuntimeError(f"model parameter {k} lives on {v.device} but the input is on {device}")
        tensors = [t for t in module.state_dict(keep_vars=True).values()]
        self.engine, self.tensors, self.sig = build(sd), tensors, self._signature(tensors)
        return self.engine

    def drop(self) -> None:
        self.engine, self.tensors, self.sig = None, [], ()


def _drop_prepared(module: nn.Module, incompatible_keys) -> None:
    module._vad_cache.drop()


class _Prepared(nn.Module):
    """nn.Module whose compute runs on prepared weights (see `_PreparedCache`); sub-classes implement `_build(sd)`."""

    def __init__(self) -> None:
        super().__init__()
        # kept out of the module's attribute dicts that state_dict / deepcopy / pickle walk
        object.__setattr__(self, "_vad_cache", _PreparedCache())
        self.register_load_state_dict_post_hook(_drop_prepared)

    def _apply(self, fn, *args, **kwargs):
        self._vad_cache.drop()
        return super()._apply(fn, *args, **kwargs)

    def __deepcopy__(self, memo):
        cache = self.__dict__.pop("_vad_cache")
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            import copy
            new.__dict__ = copy.deepcopy(self.__dict__, memo)
        finally:
            object.__setattr__(self, "_vad_cache", cache)
        object.__setattr__(new, "_vad_cache", _PreparedCache())
        return new

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_vad_cache", None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        object.__setattr__(self, "_vad_cache", _PreparedCache())

    def _build(self, sd: Dict[str, torch.Tensor]):
        raise NotImplementedError

    def _get_engine(self, device: torch.device):
        return self._vad_cache.get(self, device, self._build)


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


class Encoder(_Prepared):
    """4 x [conv3x3-BN-LeakyReLU x2, maxpool2]: 3 -> 32 -> 64 -> 128 -> latent_dim, spatial / 16 (reference
    models/autoencoder.py:24-86).  Callable on its own or as `model.encoder`."""

    def __init__(self, in_channels: int = 3, latent_dim: int = 256):
        super().__init__()
        if in_channels != 3:
            raise ValueError("vad_b200 kernels are specialised for 3-channel input (as every reference dataset is)")
        if latent_dim % 32 != 0:
            raise ValueError("latent_dim must be a multiple of 32 for the tcgen05 tile shapes")
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

    def _build(self, sd):
        return ImageEngine(prep.prepare_image_encoder(sd, prefix=""))

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 [B,3,H,W] -> fp32 [B,latent,H/16,W/16] (autoencoder.py:81-86)."""
        return self._get_engine(x.device).latent(x)


class Decoder(_Prepared):
    """4 x [convT k2 s2-BN-ReLU, conv3x3-BN-ReLU]; the last conv goes to `out_channels` and ends in Tanh (reference
    models/autoencoder.py:89-146).  Callable on its own or as `model.decoder`."""

    def __init__(self, out_channels: int = 3, latent_dim: int = 256):
        super().__init__()
        if out_channels != 3:
            raise ValueError("vad_b200 kernels are specialised for 3-channel output (as every reference dataset is)")
        if latent_dim % 32 != 0:
            raise ValueError("latent_dim must be a multiple of 32 for the tcgen05 tile shapes")
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

    def _build(self, sd):
        return ImageEngine(prep.prepare_image_decoder(sd, prefix=""))

    @torch.no_grad()
    def forward(self, z: torch.Tensor) -> torch.Tensor:
        """fp32 [B,latent,h,w] -> fp32 [B,3,16h,16w] (autoencoder.py:141-146)."""
        return self._get_engine(z.device).decode(z)


class ConvAutoencoder(_Prepared):
    """Image autoencoder whose scoring calls run as fused sm_100a kernels.

    API parity: `forward`, `get_latent`, `get_reconstruction_error(x, per_pixel=False)` behave like the reference
    (autoencoder.py:181-221).  `score_all` is the single-pass extension (recon + map + score from one forward).
    """

    def __init__(self, in_channels: int = 3, latent_dim: int = 256):
        super().__init__()
        self.encoder = Encoder(in_channels, latent_dim)
        self.decoder = Decoder(in_channels, latent_dim)
        _xavier_like_reference(self)

    def _build(self, sd):
        return ImageEngine(prep.prepare_image(sd))

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
    def score_all(self, x: torch.Tensor, want_recon: bool = True, want_heat: bool = True,
                  want_latent: bool = False) -> ScoreOutputs:
        """One forward producing score [B], min/max [B,2] and optionally the heat map [B,H,W], recon and latent."""
        return self._get_engine(x.device).run(x, want_recon=want_recon, want_heat=want_heat, want_latent=want_latent)

    @torch.no_grad()
    def score_frames(self, frames_u8: torch.Tensor, want_recon: bool = False, want_heat: bool = False,
                     want_heat_u8: bool = True) -> ScoreOutputs:
        """Score decoded frames directly: uint8 RGB [B,H,W,3] on the GPU.  ToTensor + Normalize(.5,.5) of the reference
        datasets (utils/dataset.py:65-70) happen on the device, and `heat_u8` is create_heatmap's uint8 normalisation
        of each error map (evaluate_video.py:56-57) — a caller moves a quarter of the bytes in both directions.
        Results equal `score_all(normalised fp32 frames)` bit for bit."""
        return self._get_engine(frames_u8.device).run_u8(frames_u8, want_recon, want_heat, want_heat_u8)
