"""CPU oracle for the anomaly-scoring hot path.  TEST INFRASTRUCTURE ONLY.

Nothing on the product path may import this module: only ``tests/``,
``__graft_entry__.smoke()`` and the baseline legs of ``bench.py`` (``cpu_baseline``,
``--impl reference``, and the optional ``--gpu-library-baseline``, which times these same
torch ops in eager mode on the GPU as the library bar) use it, and there only as the
checker / the reference arm being timed — never as the thing shipped or measured as ours.

What it restates
----------------
The reference (KuldeepChoksi/video-anomaly-detection) is pure Python on top of a
third-party arithmetic library: PyTorch (``requirements.txt:1`` pins ``torch>=2.0.0``;
this image has torch 2.11.0+cu128, oneDNN on CPU).  Its hot path is a fixed
composition of ``nn.Conv2d`` / ``nn.ConvTranspose2d`` / ``nn.BatchNorm2d`` (eval) /
``LeakyReLU`` / ``ReLU`` / ``Tanh`` / ``MaxPool2d`` modules plus a ConvLSTM time loop
and a squared-error reduction.  This file restates that composition as *functions of a
plain ``state_dict``* using ``torch.nn.functional`` on CPU tensors, so it needs neither
the reference's module classes nor ``/root/reference`` at run time (the GPU box has
no ``/root/reference``).  ``oracle/np_oracle.py`` restates the same arithmetic a second
time in plain numpy loops (no torch conv) for small cases.

Parity pin
----------
The reference ships no tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: the
script ``tests/golden/make_golden.py`` imports the unmodified reference classes from
``/root/reference`` in the build container, runs them on seeded weights/inputs and
commits the resulting vectors under ``tests/golden/``; ``tests/test_oracle.py`` checks
this oracle against those vectors (bit-for-bit in fp32 on the same torch build, and to
1e-6 otherwise).

All functions take ``sd``: a mapping of the reference's ``state_dict`` key names
 to tensors, and fp32 (or fp64) inputs.
"""
from __future__ import annotations

from typing import Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5  # nn.BatchNorm2d default, used by every BN in the reference


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def _conv3x3(sd: Mapping[str, Tensor], key: str, x: Tensor) -> Tensor:
    """nn.Conv2d(k=3, padding=1) — models/autoencoder.py:39,42 (and every other 3x3)."""
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=1, padding=1)


def _convt2x2(sd: Mapping[str, Tensor], key: str, x: Tensor) -> Tensor:
    """nn.ConvTranspose2d(k=2, stride=2) — models/autoencoder.py:104, video_autoencoder.py:244."""
    return F.conv_transpose2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=2)


def _bn_eval(sd: Mapping[str, Tensor], key: str, x: Tensor) -> Tensor:
    """nn.BatchNorm2d in eval mode (running statistics) — e.g. models/autoencoder.py:40."""
    return F.batch_norm(
        x,
        sd[key + ".running_mean"],
        sd[key + ".running_var"],
        sd[key + ".weight"],
        sd[key + ".bias"],
        training=False,
        eps=BN_EPS,
    )


# --------------------------------------------------------------------------------------
# image autoencoder — models/autoencoder.py
# --------------------------------------------------------------------------------------
def image_encoder(sd: Mapping[str, Tensor], x: Tensor, prefix: str = "encoder.") -> Tensor:
    """Encoder.forward — models/autoencoder.py:81-86 (layers :38-79).

    Four blocks of [conv3x3, BN, LeakyReLU(0.2)] x2 followed by MaxPool(2, 2).
    """
    for blk in ("enc1", "enc2", "enc3", "enc4"):
        p = f"{prefix}{blk}"
        x = F.leaky_relu(_bn_eval(sd, f"{p}.1", _conv3x3(sd, f"{p}.0", x)), 0.2)
        x = F.leaky_relu(_bn_eval(sd, f"{p}.4", _conv3x3(sd, f"{p}.3", x)), 0.2)
        x = F.max_pool2d(x, 2, 2)
    return x


def image_decoder(sd: Mapping[str, Tensor], z: Tensor, prefix: str = "decoder.") -> Tensor:
    """Decoder.forward — models/autoencoder.py:141-146 (layers :103-139).

    Blocks 1-3: convT(k2,s2), BN, ReLU, conv3x3, BN, ReLU.  Block 4 ends conv3x3 -> Tanh.
    """
    for blk in ("dec1", "dec2", "dec3"):
        p = f"{prefix}{blk}"
        z = F.relu(_bn_eval(sd, f"{p}.1", _convt2x2(sd, f"{p}.0", z)))
        z = F.relu(_bn_eval(sd, f"{p}.4", _conv3x3(sd, f"{p}.3", z)))
    p = f"{prefix}dec4"
    z = F.relu(_bn_eval(sd, f"{p}.1", _convt2x2(sd, f"{p}.0", z)))
    return torch.tanh(_conv3x3(sd, f"{p}.3", z))


def image_forward(sd: Mapping[str, Tensor], x: Tensor) -> Tensor:
    """ConvAutoencoder.forward — models/autoencoder.py:181-193."""
    return image_decoder(sd, image_encoder(sd, x))


def image_reconstruction_error(sd: Mapping[str, Tensor], x: Tensor, per_pixel: bool = False) -> Tensor:
    """ConvAutoencoder.get_reconstruction_error — models/autoencoder.py:199-221."""
    recon = image_forward(sd, x)
    error = ((x - recon) ** 2).mean(dim=1, keepdim=True)
    return error if per_pixel else error.mean(dim=[1, 2, 3])


# --------------------------------------------------------------------------------------
# video autoencoder — models/video_autoencoder.py
# --------------------------------------------------------------------------------------
def video_encoder(sd: Mapping[str, Tensor], x: Tensor, prefix: str = "encoder.encoder.") -> Tensor:
    """VideoEncoder.forward — models/video_autoencoder.py:217-231 (layers :191-215).

    Accepts [B,C,H,W] or [B,T,C,H,W]; T is folded into the batch (:222-229).
    """
    five_d = x.dim() == 5
    if five_d:
        b, t = x.shape[:2]
        x = x.reshape(b * t, *x.shape[2:])
    for conv_i in (0, 4, 8, 12):
        x = _conv3x3(sd, f"{prefix}{conv_i}", x)
        x = F.leaky_relu(_bn_eval(sd, f"{prefix}{conv_i + 1}", x), 0.2)
        x = F.max_pool2d(x, 2, 2)
    if five_d:
        x = x.reshape(b, t, *x.shape[1:])
    return x


def convlstm_cell(
    sd: Mapping[str, Tensor], key: str, x: Tensor, h: Tensor, c: Tensor
) -> Tuple[Tensor, Tensor]:
    """ConvLSTMCell.forward — models/video_autoencoder.py:54-85.

    gates = conv3x3(cat[x, h]); i,f,g,o = split; c' = sig(f)*c + sig(i)*tanh(g);
    h' = sig(o)*tanh(c').
    """
    gates = _conv3x3(sd, key + ".conv", torch.cat([x, h], dim=1))
    hid = h.shape[1]
    gi, gf, gg, go = torch.split(gates, hid, dim=1)
    c_next = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
    h_next = torch.sigmoid(go) * torch.tanh(c_next)
    return h_next, c_next


def convlstm(sd: Mapping[str, Tensor], x: Tensor, prefix: str = "convlstm.") -> Tensor:
    """ConvLSTM.forward — models/video_autoencoder.py:127-172 (zero init :174-179).

    Layer-outer / time-inner loop; returns the last layer's hidden sequence [B,T,hid,H,W].
    """
    n_layers = 0
    while f"{prefix}cells.{n_layers}.conv.weight" in sd:
        n_layers += 1
    b, t = x.shape[:2]
    cur = x
    for layer in range(n_layers):
        key = f"{prefix}cells.{layer}"
        hid = sd[key + ".conv.weight"].shape[0] // 4
        h = torch.zeros(b, hid, *x.shape[3:], dtype=x.dtype, device=x.device)  # `_init_hidden(..., device)` :174-179
        c = torch.zeros_like(h)
        outs = []
        for ti in range(t):
            h, c = convlstm_cell(sd, key, cur[:, ti], h, c)
            outs.append(h)
        cur = torch.stack(outs, dim=1)
    return cur


def video_decoder(sd: Mapping[str, Tensor], z: Tensor, prefix: str = "decoder.decoder.") -> Tensor:
    """VideoDecoder.forward — models/video_autoencoder.py:263-276 (layers :242-261)."""
    five_d = z.dim() == 5
    if five_d:
        b, t = z.shape[:2]
        z = z.reshape(b * t, *z.shape[2:])
    for i in (0, 3, 6):
        z = F.relu(_bn_eval(sd, f"{prefix}{i + 1}", _convt2x2(sd, f"{prefix}{i}", z)))
    z = torch.tanh(_convt2x2(sd, f"{prefix}9", z))
    if five_d:
        z = z.reshape(b, t, *z.shape[1:])
    return z


def video_forward(sd: Mapping[str, Tensor], x: Tensor) -> Tensor:
    """VideoAutoencoder.forward — models/video_autoencoder.py:329-354.

    ``proj`` is a 1x1 conv only when lstm_hidden_dim != latent_dim (:311-312).
    """
    enc = video_encoder(sd, x)
    seq = convlstm(sd, enc)
    b, t = seq.shape[:2]
    flat = seq.reshape(b * t, *seq.shape[2:])
    if "proj.weight" in sd:
        flat = F.conv2d(flat, sd["proj.weight"], sd["proj.bias"])
    proj = flat.reshape(b, t, *flat.shape[1:])
    return video_decoder(sd, proj)


def video_reconstruction_error(
    sd: Mapping[str, Tensor], x: Tensor, per_frame: bool = False, per_pixel: bool = False
) -> Tensor:
    """VideoAutoencoder.get_reconstruction_error — models/video_autoencoder.py:356-384.

    ``per_pixel`` wins over ``per_frame`` when both are set (:373-380).
    """
    recon = video_forward(sd, x)
    error = (x - recon) ** 2
    if per_pixel:
        return error.mean(dim=2, keepdim=True)
    if per_frame:
        return error.mean(dim=[2, 3, 4])
    return error.mean(dim=[1, 2, 3, 4])


# --------------------------------------------------------------------------------------
# SSIM / combined loss as an alternative score — utils/losses.py
# --------------------------------------------------------------------------------------
def ssim_window(size: int = 11, sigma: float = 1.5) -> Tensor:
    """SSIMLoss._create_gaussian_window — utils/losses.py:35-49: normalised 1-D Gaussian, outer product."""
    coords = torch.arange(size, dtype=torch.float32) - size // 2
    gauss = torch.exp(-coords ** 2 / (2 * sigma ** 2))
    gauss = gauss / gauss.sum()
    return gauss.unsqueeze(1) @ gauss.unsqueeze(0)


def ssim_map(pred: Tensor, target: Tensor, window_size: int = 11) -> Tensor:
    """SSIMLoss.forward up to `ssim_map` — utils/losses.py:65-90 (depthwise Gaussian statistics, zero padding)."""
    c = pred.shape[1]
    win = ssim_window(window_size).to(pred.device).expand(c, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    conv = lambda t: torch.nn.functional.conv2d(t, win, padding=pad, groups=c)
    mu_p, mu_t = conv(pred), conv(target)
    mu_pp, mu_tt, mu_pt = mu_p ** 2, mu_t ** 2, mu_p * mu_t
    s_pp = conv(pred ** 2) - mu_pp
    s_tt = conv(target ** 2) - mu_tt
    s_pt = conv(pred * target) - mu_pt
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu_pt + c1) * (2 * s_pt + c2)) / ((mu_pp + mu_tt + c1) * (s_pp + s_tt + c2))


def ssim_loss(pred: Tensor, target: Tensor, per_frame: bool = False) -> Tensor:
    """`1 - ssim_map.mean()` — utils/losses.py:92-93; per_frame=True: one value per leading index (the batch value is
    their mean, all frames having the same size)."""
    m = ssim_map(pred, target)
    return 1 - (m.mean(dim=[1, 2, 3]) if per_frame else m.mean())


def combined_loss(pred: Tensor, target: Tensor, alpha: float = 0.5) -> Tensor:
    """CombinedLoss.forward — utils/losses.py:116-121: (1-alpha)*MSE + alpha*(1-SSIM)."""
    return (1 - alpha) * torch.mean((pred - target) ** 2) + alpha * ssim_loss(pred, target)


# --------------------------------------------------------------------------------------
# consumers of the scores (they define the parity observables)
# --------------------------------------------------------------------------------------
def heatmap_u8(error_map: np.ndarray) -> np.ndarray:
    """create_heatmap's normalisation — evaluate_video.py:53-57 (JET LUT / resize stay host-side)."""
    e = np.asarray(error_map, dtype=np.float32).squeeze()
    norm = (e - e.min()) / (e.max() - e.min() + 1e-8)
    return (norm * 255).astype(np.uint8)


def image_flags(scores: np.ndarray, threshold: float = 0.004) -> np.ndarray:
    """UI image decision — main.py:282-283 (``score > 0.004``)."""
    return np.asarray(scores) > threshold


def video_flags(scores: np.ndarray) -> np.ndarray:
    """UI video decision — main.py:375-376 (``score > mean + 2*std``, population std)."""
    s = np.asarray(scores)
    return s > (np.mean(s) + 2 * np.std(s))


def tie_aware_rank_agreement(ref: np.ndarray, got: np.ndarray, rel_gap: float = 1e-5,
                             max_n: int = 4096, seed: int = 0) -> Tuple[int, int]:
    """Pairwise ranking agreement that ignores oracle near-ties.

    For every pair (i, j) whose ORACLE scores differ by more than ``rel_gap`` relative,
    the candidate must order the pair the same way.  Returns (pairs_checked,
    disagreements).  Above ``max_n`` scores a seeded random subset is compared.
    """
    ref = np.asarray(ref, dtype=np.float64).ravel()
    got = np.asarray(got, dtype=np.float64).ravel()
    if len(ref) > max_n:
        idx = np.random.default_rng(seed).choice(len(ref), max_n, replace=False)
        ref, got = ref[idx], got[idx]
    dr = ref[:, None] - ref[None, :]
    dg = got[:, None] - got[None, :]
    scale = np.maximum(np.abs(ref[:, None]), np.abs(ref[None, :]))
    separated = dr > rel_gap * np.maximum(scale, 1e-30)
    return int(separated.sum()), int((separated & (dg <= 0)).sum())


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench legs
# --------------------------------------------------------------------------------------
def to_dtype(sd: Mapping[str, Tensor], dtype: torch.dtype) -> dict:
    """Cast the floating tensors of a state_dict (fp64 adjudicator: SURVEY.md §4(d))."""
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def cpu_sd(sd: Mapping[str, Tensor]) -> dict:
    return {k: v.detach().to("cpu") for k, v in sd.items()}
