"""Second, torch-free restatement of the hot-path arithmetic in numpy.  TEST INFRASTRUCTURE ONLY.

``oracle/vad_oracle.py`` leans on ``torch.nn.functional`` for conv / convT / BN (the same
third-party library the reference uses).  This file spells the same arithmetic out with
index formulas so that the semantics the CUDA kernels must follow —
cross-correlation with zero pad 1, the k2/s2 transposed-conv scatter rule, eval-mode BN,
gate order i,f,g,o — are pinned independently of torch.  float64 throughout; meant for
SMALL cases only (the tests use <= 32x32 inputs).
"""
from __future__ import annotations

from typing import Mapping

import numpy as np

BN_EPS = 1e-5


def _np(sd: Mapping, key: str) -> np.ndarray:
    v = sd[key]
    return np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float64)


def conv3x3(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """out[n,co,y,x] = b[co] + sum_{ci,ky,kx} in[n,ci,y+ky-1,x+kx-1] * w[co,ci,ky,kx], zero pad.

    nn.Conv2d(k=3, padding=1) — reference models/autoencoder.py:39.
    """
    n, ci, h, wd = x.shape
    xp = np.zeros((n, ci, h + 2, wd + 2), dtype=np.float64)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((n, w.shape[0], h, wd), dtype=np.float64)
    for ky in range(3):
        for kx in range(3):
            out += np.einsum("nchw,oc->nohw", xp[:, :, ky:ky + h, kx:kx + wd], w[:, :, ky, kx])
    return out + b[None, :, None, None]


def convt2x2(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """out[n,co,2i+di,2j+dj] = b[co] + sum_ci in[n,ci,i,j] * w[ci,co,di,dj].

    nn.ConvTranspose2d(k=2, stride=2) — reference models/autoencoder.py:104; stride == kernel,
    so every output pixel receives exactly one tap per input channel.
    """
    n, ci, h, wd = x.shape
    co = w.shape[1]
    out = np.zeros((n, co, 2 * h, 2 * wd), dtype=np.float64)
    for di in range(2):
        for dj in range(2):
            out[:, :, di::2, dj::2] = np.einsum("nchw,co->nohw", x, w[:, :, di, dj])
    return out + b[None, :, None, None]


def bn_eval(x: np.ndarray, sd: Mapping, key: str) -> np.ndarray:
    """y = (x - running_mean) / sqrt(running_var + 1e-5) * weight + bias (eval-mode BatchNorm2d)."""
    mean, var = _np(sd, key + ".running_mean"), _np(sd, key + ".running_var")
    g, beta = _np(sd, key + ".weight"), _np(sd, key + ".bias")
    s = g / np.sqrt(var + BN_EPS)
    return (x - mean[None, :, None, None]) * s[None, :, None, None] + beta[None, :, None, None]


def leaky(x: np.ndarray, slope: float = 0.2) -> np.ndarray:
    return np.where(x >= 0, x, slope * x)


def maxpool2(x: np.ndarray) -> np.ndarray:
    n, c, h, w = x.shape
    return x.reshape(n, c, h // 2, 2, w // 2, 2).max(axis=(3, 5))


def sigmoid(x: np.ndarray) -> np.ndarray:
    return 1.0 / (1.0 + np.exp(-x))


def image_forward(sd: Mapping, x: np.ndarray) -> np.ndarray:
    """ConvAutoencoder.forward — reference models/autoencoder.py:181-193."""
    x = np.asarray(x, dtype=np.float64)
    for blk in ("enc1", "enc2", "enc3", "enc4"):
        p = f"encoder.{blk}"
        for conv_i in (0, 3):
            x = conv3x3(x, _np(sd, f"{p}.{conv_i}.weight"), _np(sd, f"{p}.{conv_i}.bias"))
            x = leaky(bn_eval(x, sd, f"{p}.{conv_i + 1}"))
        x = maxpool2(x)
    for blk in ("dec1", "dec2", "dec3", "dec4"):
        p = f"decoder.{blk}"
        x = convt2x2(x, _np(sd, f"{p}.0.weight"), _np(sd, f"{p}.0.bias"))
        x = np.maximum(bn_eval(x, sd, f"{p}.1"), 0.0)
        x = conv3x3(x, _np(sd, f"{p}.3.weight"), _np(sd, f"{p}.3.bias"))
        x = np.tanh(x) if blk == "dec4" else np.maximum(bn_eval(x, sd, f"{p}.4"), 0.0)
    return x


def video_forward(sd: Mapping, x: np.ndarray) -> np.ndarray:
    """VideoAutoencoder.forward — reference models/video_autoencoder.py:329-354."""
    x = np.asarray(x, dtype=np.float64)
    b, t = x.shape[:2]
    f = x.reshape(b * t, *x.shape[2:])
    for conv_i in (0, 4, 8, 12):
        p = f"encoder.encoder.{conv_i}"
        f = conv3x3(f, _np(sd, p + ".weight"), _np(sd, p + ".bias"))
        f = maxpool2(leaky(bn_eval(f, sd, f"encoder.encoder.{conv_i + 1}")))
    seq = f.reshape(b, t, *f.shape[1:])
    layer = 0
    while f"convlstm.cells.{layer}.conv.weight" in sd:
        w = _np(sd, f"convlstm.cells.{layer}.conv.weight")
        bias = _np(sd, f"convlstm.cells.{layer}.conv.bias")
        hid = w.shape[0] // 4
        h = np.zeros((b, hid) + seq.shape[3:], dtype=np.float64)
        c = np.zeros_like(h)
        outs = []
        for ti in range(t):
            gates = conv3x3(np.concatenate([seq[:, ti], h], axis=1), w, bias)
            gi, gf, gg, go = (gates[:, k * hid:(k + 1) * hid] for k in range(4))
            c = sigmoid(gf) * c + sigmoid(gi) * np.tanh(gg)
            h = sigmoid(go) * np.tanh(c)
            outs.append(h)
        seq = np.stack(outs, axis=1)
        layer += 1
    f = seq.reshape(b * t, *seq.shape[2:])
    if "proj.weight" in sd:
        f = np.einsum("nchw,oc->nohw", f, _np(sd, "proj.weight")[:, :, 0, 0]) + _np(sd, "proj.bias")[None, :, None, None]
    for i in (0, 3, 6):
        p = f"decoder.decoder.{i}"
        f = convt2x2(f, _np(sd, p + ".weight"), _np(sd, p + ".bias"))
        f = np.maximum(bn_eval(f, sd, f"decoder.decoder.{i + 1}"), 0.0)
    f = np.tanh(convt2x2(f, _np(sd, "decoder.decoder.9.weight"), _np(sd, "decoder.decoder.9.bias")))
    return f.reshape(b, t, *f.shape[1:])


def frame_scores(x: np.ndarray, recon: np.ndarray) -> np.ndarray:
    """mean over (C,H,W) of (x-recon)^2 — reference autoencoder.py:214-221 / video_autoencoder.py:371-380."""
    e = (np.asarray(x, dtype=np.float64) - recon) ** 2
    return e.mean(axis=(-3, -2, -1))
