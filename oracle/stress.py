"""Seeded "stress" state_dicts.  TEST INFRASTRUCTURE ONLY.

At the reference's random init (xavier-normal, identity BN) the reconstruction is ~0 and
the score degenerates to mean(x^2) — a kernel that outputs zeros would
pass.  These generators overwrite a reference-shaped ``state_dict`` with trained-looking
values: He-scaled conv weights (O(1) activations in every layer), non-zero biases and
randomised BatchNorm affine + running statistics, so that recon / heatmap parity checks
actually exercise every layer.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Mapping

import torch


def stress_state_dict(template: Mapping[str, torch.Tensor], seed: int = 1) -> Dict[str, torch.Tensor]:
    """Return a new state_dict with the same keys/shapes as ``template``.

    Conv2d [Cout,Cin,kh,kw]: N(0, 2/(Cin*kh*kw)); ConvTranspose2d k2s2 [Cin,Cout,2,2]:
    N(0, 2/Cin) (one tap per input channel reaches each output pixel); ConvLSTM gate conv and
    the video model's Tanh-feeding layer use gain 1 instead of sqrt(2), the image model's last conv gain 0.1
    (calibrated so that recon RMS ~0.65 with ~10 % of pixels beyond |0.95|: a saturated Tanh would hide errors).  BN: gamma U[0.6,1.4],
    beta N(0,0.2), running_mean N(0,0.3), running_var U[0.5,1.5].  Biases N(0,0.1).
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    keys = list(template.keys())
    bn_prefixes = {k[: -len(".running_mean")] for k in keys if k.endswith(".running_mean")}
    for k in keys:
        v = template[k]
        prefix, _, leaf = k.rpartition(".")
        if leaf == "num_batches_tracked":
            out[k] = torch.tensor(100, dtype=v.dtype)
        elif prefix in bn_prefixes:
            n = v.shape[0]
            if leaf == "weight":
                out[k] = 0.6 + 0.8 * torch.rand(n, generator=g)
            elif leaf == "bias":
                out[k] = 0.2 * torch.randn(n, generator=g)
            elif leaf == "running_mean":
                out[k] = 0.3 * torch.randn(n, generator=g)
            else:  # running_var
                out[k] = 0.5 + torch.rand(n, generator=g)
        elif leaf == "bias":
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif leaf == "weight" and v.dim() == 4:
            is_convt = v.shape[2] == 2  # the only 2x2 kernels in either model are the ConvTranspose2d
            fan_in = v.shape[0] if is_convt else v.shape[1] * v.shape[2] * v.shape[3]
            feeds_tanh = prefix.endswith("decoder.9") or ".cells." in prefix
            gain = 1.0 if feeds_tanh else 2.0 ** 0.5
            if prefix.endswith("dec4.3"):
                gain = 0.1  # 16 layers of ~1.3x growth upstream: keeps the image model's Tanh mostly unsaturated
            out[k] = torch.randn(v.shape, generator=g) * (gain / fan_in ** 0.5)
        else:
            raise KeyError(f"unexpected state_dict entry {k} {tuple(v.shape)}")
        out[k] = out[k].to(v.dtype) if v.is_floating_point() else out[k]
    return out


def state_dict_digest(sd: Mapping[str, torch.Tensor]) -> str:
    """sha256 over key names + raw fp32 bytes — pins that two processes built identical weights."""
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        v = sd[k].detach().cpu().contiguous()
        h.update(k.encode())
        h.update(v.numpy().tobytes())
    return h.hexdigest()
