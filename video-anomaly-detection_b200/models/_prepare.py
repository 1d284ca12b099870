"""Host-side weight preparation: fold eval-mode BatchNorm into the preceding conv and repack weights into the
K-major bf16 GEMM operands libvad_b200 consumes.  Pure torch; runs on CPU or GPU tensors (tests exercise it on CPU).

Arithmetic restated from the reference (Appendix B of SURVEY.md):
  BatchNorm2d eval: y = (x - mean) / sqrt(var + 1e-5) * gamma + beta       (models/autoencoder.py:40 etc.)
  Conv2d 3x3 weight [Cout, Cin, 3, 3]; ConvTranspose2d k2 s2 weight [Cin, Cout, 2, 2]
  ConvLSTM gate conv over cat[x, h]; output chunks i, f, g, o                (models/video_autoencoder.py:64-75)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Mapping, Optional, Tuple

import torch

BN_EPS = 1e-5
LSTM_CH_PER_TILE = 32  # one 128-column GEMM tile = 4 gates x 32 hidden channels


@dataclass
class GemmWeights:
    """One layer's B operand: `w` bf16 [n_total, ntaps * ctap] (K-major), `bias` fp32 [n_total]."""
    w: torch.Tensor
    bias: torch.Tensor
    ntaps: int
    ctap: int      # input channels per tap (weight columns per tap)
    n_total: int
    cout: int      # real output channels
    # narrow 3x3 layers only: the same weights as [kx*Cout + co (zero rows up to a multiple of 16)][ky*Cin + ci], the
    # layout of the kernel that folds the horizontal taps into the GEMM N extent (include/vad_b200.h `weight_kx`)
    w_kx: Optional[torch.Tensor] = None
    # 3x3 layers with 32 input channels only: the pixel-pair folded form (include/vad_b200.h `pair_fold`):
    # `w_pair` bf16 [2*n_total, 9*64], `bias_pair` fp32 [2*n_total]
    w_pair: Optional[torch.Tensor] = None
    bias_pair: Optional[torch.Tensor] = None


@dataclass
class FirstConvWeights:
    """3-channel first conv: `w` fp32 [27, cout] with k = (ky*3+kx)*3+ci, `bias` fp32 [cout];
    `w_tc` bf16 [cout, 32] (K padded 27 -> 32 with zeros) for the tensor-core kernel."""
    w: torch.Tensor
    bias: torch.Tensor
    cout: int
    w_tc: Optional[torch.Tensor] = None
    # pooled layer only: bf16 [4*cout, 64] for the kernel that folds the 2x2 pooling window into the GEMM N extent
    # (include/vad_b200.h `vad_first_conv_pool`): row = (py*2+px)*cout + co, column = ((py+ky)*4 + (px+kx))*3 + ci
    w_pf: Optional[torch.Tensor] = None


def bn_scale_shift(sd: Mapping[str, torch.Tensor], bn: Optional[str], cout: int, ref: torch.Tensor
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-channel (scale, shift) of an eval-mode BatchNorm; identity when `bn` is None."""
    if bn is None:
        return torch.ones(cout, dtype=torch.float64, device=ref.device), torch.zeros(cout, dtype=torch.float64,
                                                                                     device=ref.device)
    g = sd[bn + ".weight"].double()
    b = sd[bn + ".bias"].double()
    m = sd[bn + ".running_mean"].double()
    v = sd[bn + ".running_var"].double()
    s = g / torch.sqrt(v + BN_EPS)
    return s, b - m * s


def fold_conv(sd: Mapping[str, torch.Tensor], conv: str, bn: Optional[str]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Conv2d [Cout,Cin,kh,kw] followed by eval BN -> (weight', bias') in fp64."""
    w = sd[conv + ".weight"].double()
    b = sd[conv + ".bias"].double()
    s, t = bn_scale_shift(sd, bn, w.shape[0], w)
    return w * s.view(-1, 1, 1, 1), b * s + t


def fold_convt(sd: Mapping[str, torch.Tensor], conv: str, bn: Optional[str]) -> Tuple[torch.Tensor, torch.Tensor]:
    """ConvTranspose2d [Cin,Cout,2,2] followed by eval BN -> (weight', bias') in fp64."""
    w = sd[conv + ".weight"].double()
    b = sd[conv + ".bias"].double()
    s, t = bn_scale_shift(sd, bn, w.shape[1], w)
    return w * s.view(1, -1, 1, 1), b * s + t


def pack_conv3x3(w: torch.Tensor, b: torch.Tensor, pad_n_to: int = 0) -> GemmWeights:
    """[Cout,Cin,3,3] -> [Cout, 9*Cin] with K = (ky*3+kx)*Cin + ci; optional zero rows up to `pad_n_to`."""
    cout, cin = w.shape[0], w.shape[1]
    wk = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    n_total = max(cout, pad_n_to)
    if n_total > cout:
        wk = torch.cat([wk, wk.new_zeros(n_total - cout, 9 * cin)], 0)
        b = torch.cat([b, b.new_zeros(n_total - cout)], 0)
    gw = GemmWeights(wk.to(torch.bfloat16).contiguous(), b.float().contiguous(), 9, cin, n_total, cout)
    if cout <= 64 and cin in (32, 64):
        gw.w_kx = pack_conv3x3_kx(w)
    if cin == 32 and n_total == cout and cout in (32, 64):
        gw.w_pair, gw.bias_pair = pack_conv3x3_pair(w, b)
    return gw


def pack_conv3x3_pair(w: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[Cout,32,3,3] -> the same convolution on PAIRS of horizontally adjacent pixels: bf16 [2*Cout, 9*64] with row
    p_out*Cout + co and column (ky*3 + kxp)*64 + p_in*32 + ci holding w[co][ci][ky][kx], kx = 2*(kxp-1) + p_in - p_out + 1
    (zero where that is outside 0..2: a third of the matrix), and the bias repeated for both pixels."""
    cout, cin = w.shape[0], w.shape[1]
    wp = torch.zeros(2, cout, 3, 3, 2, cin, dtype=w.dtype, device=w.device)  # [p_out][co][ky][kxp][p_in][ci]
    for p_out in range(2):
        for kxp in range(3):
            for p_in in range(2):
                kx = 2 * (kxp - 1) + p_in - p_out + 1
                if 0 <= kx <= 2:
                    wp[p_out, :, :, kxp, p_in, :] = w[:, :, :, kx].permute(0, 2, 1)  # [co][ky][ci]
    return (wp.reshape(2 * cout, 9 * 2 * cin).to(torch.bfloat16).contiguous(),
            torch.cat([b, b]).float().contiguous())


def pack_conv3x3_kx(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,3,3] -> bf16 [3*Cout padded to a multiple of 16, 3*Cin]: row = kx*Cout + co, column = ky*Cin + ci."""
    cout, cin = w.shape[0], w.shape[1]
    wk = w.permute(3, 0, 2, 1).reshape(3 * cout, 3 * cin)  # [kx, co, ky, ci]
    rows = (3 * cout + 15) // 16 * 16
    if rows > 3 * cout:
        wk = torch.cat([wk, wk.new_zeros(rows - 3 * cout, 3 * cin)], 0)
    return wk.to(torch.bfloat16).contiguous()


def pack_conv1x1(w: torch.Tensor, b: torch.Tensor) -> GemmWeights:
    cout, cin = w.shape[0], w.shape[1]
    return GemmWeights(w.reshape(cout, cin).to(torch.bfloat16).contiguous(), b.float().contiguous(), 1, cin, cout, cout)


def pack_convt2x2(w: torch.Tensor, b: torch.Tensor, pad_n_to: int = 0) -> GemmWeights:
    """[Cin,Cout,2,2] -> [4*Cout, Cin], row n = (di*2+dj)*Cout + co; bias repeated per (di,dj)."""
    cin, cout = w.shape[0], w.shape[1]
    wk = w.permute(2, 3, 1, 0).reshape(4 * cout, cin)
    bk = b.repeat(4)
    n_total = max(4 * cout, pad_n_to)
    if n_total > 4 * cout:
        wk = torch.cat([wk, wk.new_zeros(n_total - 4 * cout, cin)], 0)
        bk = torch.cat([bk, bk.new_zeros(n_total - 4 * cout)], 0)
    return GemmWeights(wk.to(torch.bfloat16).contiguous(), bk.float().contiguous(), 1, cin, n_total, cout)


def lstm_row_permutation(hid: int, device=None) -> torch.Tensor:
    """new GEMM row -> original gate-conv output channel.

    Original rows: gate*hid + j (gate order i,f,g,o — models/video_autoencoder.py:75).  New rows: tiles of 128 =
    [gate][32 channels] so that one accumulator row holds all four gates of the same hidden channels."""
    if hid % LSTM_CH_PER_TILE != 0:
        raise ValueError(f"lstm hidden dim {hid} must be a multiple of {LSTM_CH_PER_TILE}")
    n = torch.arange(4 * hid, device=device)
    tile, within = n // 128, n % 128
    gate, jj = within // LSTM_CH_PER_TILE, within % LSTM_CH_PER_TILE
    return gate * hid + tile * LSTM_CH_PER_TILE + jj


def pack_lstm(w: torch.Tensor, b: torch.Tensor, hid: int) -> GemmWeights:
    """Gate conv [4*hid, in+hid, 3, 3] -> row-permuted [4*hid, 9*(in+hid)]."""
    perm = lstm_row_permutation(hid, w.device)
    g = pack_conv3x3(w[perm], b[perm])
    g.cout = hid
    return g


def pack_first_conv_pool_folded(w: torch.Tensor) -> torch.Tensor:
    """[Cout,3,3,3] -> bf16 [4*Cout, 64]: the four conv outputs of a 2x2 pooling window as one GEMM row over the 4x4x3
    input window (zeros outside each output position's 3x3 sub-window; K = 48 padded to 64)."""
    cout = w.shape[0]
    out = torch.zeros(4, cout, 4, 4, 3, dtype=torch.float64, device=w.device)  # [pos][co][wy][wx][ci]
    for py in range(2):
        for px in range(2):
            out[py * 2 + px, :, py:py + 3, px:px + 3, :] = w.permute(0, 2, 3, 1)  # [co][ky][kx][ci]
    flat = torch.zeros(4 * cout, 64, dtype=torch.float64, device=w.device)
    flat[:, :48] = out.reshape(4 * cout, 48)
    return flat.to(torch.bfloat16).contiguous()


def pack_first_conv(w: torch.Tensor, b: torch.Tensor, pooled: bool = False) -> FirstConvWeights:
    cout = w.shape[0]
    wk = w.permute(2, 3, 1, 0).reshape(27, cout)  # [ky,kx,ci,co]
    w_tc = torch.zeros(cout, 32, dtype=torch.float64, device=w.device)
    w_tc[:, :27] = wk.t()
    return FirstConvWeights(wk.float().contiguous(), b.float().contiguous(), cout,
                            w_tc.to(torch.bfloat16).contiguous(),
                            pack_first_conv_pool_folded(w) if (pooled and cout == 32) else None)


def prepare_image_encoder(sd: Mapping[str, torch.Tensor], prefix: str = "encoder.") -> Dict[str, object]:
    """`Encoder` parameters (keys `<prefix>enc{1..4}.{0,1,3,4}.*`) -> packed layers."""
    out: Dict[str, object] = {}
    out["enc1.0"] = pack_first_conv(*fold_conv(sd, f"{prefix}enc1.0", f"{prefix}enc1.1"))
    out["enc1.3"] = pack_conv3x3(*fold_conv(sd, f"{prefix}enc1.3", f"{prefix}enc1.4"))
    for blk in ("enc2", "enc3", "enc4"):
        out[f"{blk}.0"] = pack_conv3x3(*fold_conv(sd, f"{prefix}{blk}.0", f"{prefix}{blk}.1"))
        out[f"{blk}.3"] = pack_conv3x3(*fold_conv(sd, f"{prefix}{blk}.3", f"{prefix}{blk}.4"))
    return out


def prepare_image_decoder(sd: Mapping[str, torch.Tensor], prefix: str = "decoder.") -> Dict[str, object]:
    """`Decoder` parameters (keys `<prefix>dec{1..4}.{0,1,3,4}.*`) -> packed layers."""
    out: Dict[str, object] = {}
    for blk in ("dec1", "dec2", "dec3"):
        out[f"{blk}.0"] = pack_convt2x2(*fold_convt(sd, f"{prefix}{blk}.0", f"{prefix}{blk}.1"))
        out[f"{blk}.3"] = pack_conv3x3(*fold_conv(sd, f"{prefix}{blk}.3", f"{prefix}{blk}.4"))
    out["dec4.0"] = pack_convt2x2(*fold_convt(sd, f"{prefix}dec4.0", f"{prefix}dec4.1"))
    out["dec4.3"] = pack_conv3x3(*fold_conv(sd, f"{prefix}dec4.3", None), pad_n_to=16)
    return out


def prepare_image(sd: Mapping[str, torch.Tensor]) -> Dict[str, object]:
    """ConvAutoencoder state_dict (SURVEY Appendix D keys) -> packed layers, in execution order."""
    out = prepare_image_encoder(sd)
    out.update(prepare_image_decoder(sd))
    return out


def prepare_video_encoder(sd: Mapping[str, torch.Tensor], prefix: str = "encoder.encoder.") -> Dict[str, object]:
    out: Dict[str, object] = {}
    out["enc.0"] = pack_first_conv(*fold_conv(sd, f"{prefix}0", f"{prefix}1"), pooled=True)
    for i in (4, 8, 12):
        out[f"enc.{i}"] = pack_conv3x3(*fold_conv(sd, f"{prefix}{i}", f"{prefix}{i + 1}"))
    return out


def prepare_lstm_cell(sd: Mapping[str, torch.Tensor], prefix: str) -> GemmWeights:
    """One `ConvLSTMCell` (keys `<prefix>conv.{weight,bias}`) -> row-permuted gate GEMM operand."""
    w = sd[f"{prefix}conv.weight"].double()
    b = sd[f"{prefix}conv.bias"].double()
    return pack_lstm(w, b, w.shape[0] // 4)


def prepare_convlstm(sd: Mapping[str, torch.Tensor], prefix: str = "convlstm.") -> Dict[str, object]:
    out: Dict[str, object] = {}
    layer = 0
    while f"{prefix}cells.{layer}.conv.weight" in sd:
        out[f"lstm.{layer}"] = prepare_lstm_cell(sd, f"{prefix}cells.{layer}.")
        layer += 1
    out["lstm_layers"] = layer
    return out


def prepare_video_decoder(sd: Mapping[str, torch.Tensor], prefix: str = "decoder.decoder.") -> Dict[str, object]:
    out: Dict[str, object] = {}
    for i in (0, 3, 6):
        out[f"dec.{i}"] = pack_convt2x2(*fold_convt(sd, f"{prefix}{i}", f"{prefix}{i + 1}"))
    out["dec.9"] = pack_convt2x2(*fold_convt(sd, f"{prefix}9", None), pad_n_to=16)
    return out


def prepare_video(sd: Mapping[str, torch.Tensor]) -> Dict[str, object]:
    """VideoAutoencoder state_dict -> packed layers."""
    out = prepare_video_encoder(sd)
    out.update(prepare_convlstm(sd))
    if "proj.weight" in sd:
        out["proj"] = pack_conv1x1(sd["proj.weight"].double(), sd["proj.bias"].double())
    out.update(prepare_video_decoder(sd))
    return out


def to_device(packed: Dict[str, object], device) -> Dict[str, object]:
    res: Dict[str, object] = {}
    for k, v in packed.items():
        if isinstance(v, (GemmWeights, FirstConvWeights)):
            v.w = v.w.to(device)
            v.bias = v.bias.to(device)
            if isinstance(v, FirstConvWeights) and v.w_tc is not None:
                v.w_tc = v.w_tc.to(device)
            if isinstance(v, FirstConvWeights) and v.w_pf is not None:
                v.w_pf = v.w_pf.to(device)
            if isinstance(v, GemmWeights) and v.w_kx is not None:
                v.w_kx = v.w_kx.to(device)
            if isinstance(v, GemmWeights) and v.w_pair is not None:
                v.w_pair, v.bias_pair = v.w_pair.to(device), v.bias_pair.to(device)
        res[k] = v
    return res
