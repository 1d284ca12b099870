"""Layer-by-layer schedules of the two autoencoders in Python, one per-layer C-ABI call per fused layer.

TEST / TUNING HELPERS: the product path is `_engine.py` (one model-level C call per reference method, the schedule
lives in csrc/vad_model.cu).  This module keeps the per-layer wrappers the layer tests and the tools/ scripts drive
(`_conv`, `_convt`, `_first_conv`, `_score_layer`, the fused tails, the ConvLSTM sequence calls) and the same two
schedules written out in Python, which the GPU tests use to pin the C schedule bit for bit.  Its shape-keyed buffer
cache is not stream-safe — single-stream use only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch

from . import _native as nat
from ._engine import ScoreOutputs, _check_hw, _require_cuda_input
from ._prepare import FirstConvWeights, GemmWeights

LEAKY, RELU, IDENT = 0.2, 0.0, 1.0

# bench.py sets this to a list to get (layer name, start event, end event) for every kernel launch
PROFILE: Optional[list] = None


def _timed(what: str, fn) -> None:
    if PROFILE is None:
        fn()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    PROFILE.append((what, e0, e1))


class _Buffers:
    """Shape-keyed cache of device buffers (the library itself never allocates)."""

    def __init__(self) -> None:
        self._bufs: Dict[Tuple, torch.Tensor] = {}

    def get(self, name: str, shape: Tuple[int, ...], dtype: torch.dtype, device) -> torch.Tensor:
        key = (name, tuple(shape), dtype, str(device))
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=device)
            self._bufs[key] = t
        return t


def _gemm_desc(w: GemmWeights, src: torch.Tensor, B: int, H: int, W: int, epi: int, slope: float,
                out: Optional[torch.Tensor], *, c0: int = 0, T0: int = 1, t0: int = 0, src1: Optional[torch.Tensor] = None,
                c1: int = 0, T1: int = 1, t1: int = 0, out_frame_stride: int = 0, out_cpitch: int = 0,
                out_offset_elems: int = 0, c_state: Optional[torch.Tensor] = None, lstm_first: bool = False,
                x: Optional[torch.Tensor] = None, recon: Optional[torch.Tensor] = None,
                heat: Optional[torch.Tensor] = None, partials: Optional[torch.Tensor] = None) -> "nat.ConvDesc":
    d = nat.ConvDesc()
    d.src0 = src.data_ptr()
    d.src1 = nat.ptr(src1)
    d.c0 = c0 if c0 else w.ctap  # channels read from source 0 (ConvLSTM: the x half of cat[x, h])
    d.c1 = c1 if src1 is not None else 0
    d.T0, d.T1, d.t0, d.t1 = T0, T1, t0, t1
    d.B, d.H, d.W = B, H, W
    d.ntaps = w.ntaps
    d.weight = w.w.data_ptr()
    d.weight_kx = nat.ptr(w.w_kx)
    d.bias = w.bias.data_ptr()
    d.w_ctap = w.ctap
    d.n_total = w.n_total
    d.cout = w.cout
    d.epilogue = epi
    d.slope = slope
    if out is not None:
        d.out = out.data_ptr() + out_offset_elems * out.element_size()
    d.out_frame_stride = out_frame_stride
    d.out_cpitch = out_cpitch
    d.c_state = nat.ptr(c_state)
    d.lstm_first = 1 if lstm_first else 0
    d.x, d.recon, d.heat, d.partials = nat.ptr(x), nat.ptr(recon), nat.ptr(heat), nat.ptr(partials)
    return d


def _gemm_layer(w: GemmWeights, src: torch.Tensor, B: int, H: int, W: int, epi: int, slope: float,
                out: Optional[torch.Tensor], *, what: str = "", **kw) -> None:
    d = _gemm_desc(w, src, B, H, W, epi, slope, out, **kw)
    _timed(what or "vad_conv_layer", lambda: nat.conv_layer(d, what or "vad_conv_layer"))


def _score_layer(w: GemmWeights, src: torch.Tensor, frames: int, H: int, W: int, epi: int, x: torch.Tensor,
                 want_recon: bool, want_heat: bool, Ho: int, Wo: int, bufs: "_Buffers", what: str) -> "ScoreOutputs":
    """Last decoder layer with the fused tanh + (x - recon)^2 reduction, then the per-frame finalisation."""
    dev = x.device
    recon = torch.empty(frames, 3, Ho, Wo, dtype=torch.float32, device=dev) if want_recon else None
    heat = torch.empty(frames, Ho, Wo, dtype=torch.float32, device=dev) if want_heat else None
    d = _gemm_desc(w, src, frames, H, W, epi, IDENT, None, x=x, recon=recon, heat=heat, partials=x)
    tiles = nat.layer_tiles(d)  # the tiling (and so the number of per-tile partials) is the library's choice
    partials = bufs.get("partials", (tiles, 4, 4), torch.float32, dev)  # one (sum, min, max, -) per tile and warp quarter
    d.partials = partials.data_ptr()
    _timed(what, lambda: nat.conv_layer(d, what))
    score, minmax = _finalize(partials, frames, 4 * (tiles // frames), Ho, Wo, bufs, dev)
    return ScoreOutputs(score, minmax, heat, recon)


def _fused_tail(w6: GemmWeights, w9: GemmWeights, src: torch.Tensor, frames: int, H: int, W: int, x: torch.Tensor,
                want_recon: bool, want_heat: bool, Ho: int, Wo: int, bufs: "_Buffers") -> "ScoreOutputs":
    """Video decoder.6 (ConvT 64->32 + BN + ReLU) + decoder.9 (ConvT 32->3 + Tanh) + scoring: `vad_convt2_score`."""
    dev = x.device
    recon = torch.empty(frames, 3, Ho, Wo, dtype=torch.float32, device=dev) if want_recon else None
    heat = torch.empty(frames, Ho, Wo, dtype=torch.float32, device=dev) if want_heat else None
    d = _gemm_desc(w6, src, frames, H, W, nat.EPI_CONVT, RELU, None, x=x, recon=recon, heat=heat, partials=x)
    tiles = nat.load().vad_convt2_score_tiles(C.byref(d))
    if tiles <= 0:
        nat.check(tiles if tiles < 0 else -1, "vad_convt2_score_tiles")
    partials = bufs.get("partials", (tiles, 4, 4), torch.float32, dev)
    d.partials = partials.data_ptr()
    _timed("decoder.6+9+score", lambda: nat.check(
        nat.load().vad_convt2_score(C.byref(d), w9.w.data_ptr(), w9.bias.data_ptr(), nat.stream_ptr()),
        "vad_convt2_score"))
    score, minmax = _finalize(partials, frames, 4 * (tiles // frames), Ho, Wo, bufs, dev)
    return ScoreOutputs(score, minmax, heat, recon)


def _fused_image_tail(wt: GemmWeights, wc: GemmWeights, src: torch.Tensor, frames: int, H: int, W: int,
                      x: torch.Tensor, want_recon: bool, want_heat: bool, bufs: "_Buffers") -> "ScoreOutputs":
    """Image dec4.0 (ConvT 32->32 + BN + ReLU) + dec4.3 (Conv3x3 32->3 + Tanh) + scoring: `vad_convt_conv_score`.
    src bf16 NHWC [frames,H,W,32]; x fp32 [frames,3,2H,2W]."""
    dev = x.device
    Ho, Wo = 2 * H, 2 * W
    recon = torch.empty(frames, 3, Ho, Wo, dtype=torch.float32, device=dev) if want_recon else None
    heat = torch.empty(frames, Ho, Wo, dtype=torch.float32, device=dev) if want_heat else None
    d = _gemm_desc(wt, src, frames, H, W, nat.EPI_CONVT, RELU, None, x=x, recon=recon, heat=heat, partials=x)
    tiles = nat.load().vad_convt_conv_score_tiles(C.byref(d))
    if tiles <= 0:
        nat.check(tiles if tiles < 0 else -1, "vad_convt_conv_score_tiles")
    partials = bufs.get("partials", (tiles, 4, 4), torch.float32, dev)
    d.partials = partials.data_ptr()
    _timed("dec4.0+4.3+score", lambda: nat.check(
        nat.load().vad_convt_conv_score(C.byref(d), wc.w_kx.data_ptr(), wc.bias.data_ptr(), nat.stream_ptr()),
        "vad_convt_conv_score"))
    score, minmax = _finalize(partials, frames, 4 * (tiles // frames), Ho, Wo, bufs, dev)
    return ScoreOutputs(score, minmax, heat, recon)


FIRST_CONV_TC = os.environ.get("VAD_FIRST_TC", "1") != "0"
# VAD_FUSE_DEC=0: run the decoders' last two layers one by one (vad_conv_layer) instead of the fused tail kernels
# (video: vad_convt2_score, image: vad_convt_conv_score)
FUSE_DEC_TAIL = os.environ.get("VAD_FUSE_DEC", "1") != "0"
# VAD_LSTM2=0: one launch per ConvLSTM layer (vad_convlstm_sequence) instead of the two-layer wavefront kernel
FUSE_LSTM_LAYERS = os.environ.get("VAD_LSTM2", "1") != "0"


# image enc1.0 + enc1.3 + pool as one kernel (vad_enc1_fused); tests flip this
FUSE_ENC1 = os.environ.get("VAD_FUSE_ENC1", "1") != "0"
# the pooled first conv on the pool-folded kernel (vad_first_conv_pool) when its weights were prepared; tests flip this
FIRST_CONV_POOL_FOLD = os.environ.get("VAD_FIRST_PF", "1") != "0"


def _first_conv(w: FirstConvWeights, x: torch.Tensor, B: int, H: int, W: int, pool: bool, out: torch.Tensor) -> None:
    if FIRST_CONV_TC and FIRST_CONV_POOL_FOLD and pool and w.cout == 32 and w.w_pf is not None:
        _timed("first_conv", lambda: nat.check(
            nat.load().vad_first_conv_pool(x.data_ptr(), w.w_pf.data_ptr(), w.bias.data_ptr(), LEAKY, B, H, W,
                                           out.data_ptr(), nat.stream_ptr()), "vad_first_conv_pool"))
        return
    if FIRST_CONV_TC and w.cout == 32 and w.w_tc is not None:
        _timed("first_conv", lambda: nat.check(
            nat.load().vad_first_conv_tc(x.data_ptr(), w.w_tc.data_ptr(), w.bias.data_ptr(), LEAKY, 1 if pool else 0,
                                         B, H, W, out.data_ptr(), nat.stream_ptr()), "vad_first_conv_tc"))
        return
    _timed("first_conv", lambda: nat.check(
        nat.load().vad_first_conv(x.data_ptr(), w.w.data_ptr(), w.bias.data_ptr(), w.cout, LEAKY, 1 if pool else 0,
                                  B, H, W, out.data_ptr(), nat.stream_ptr()), "vad_first_conv"))


# 3x3 layers with 32 input channels on the pixel-pair view (include/vad_b200.h `pair_fold`) when prepared; tests flip this
PAIR_FOLD = os.environ.get("VAD_PAIR_FOLD", "1") != "0"


def _conv(w: GemmWeights, src, B, H, W, out, slope, pool=False, what=""):
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    if PAIR_FOLD and w.w_pair is not None and w.ntaps == 9 and w.ctap == 32 and w.n_total <= 64 and W % 2 == 0 and \
            W >= 32 and H >= 16:
        d = _gemm_desc(w, src, B, H, W // 2, nat.EPI_POOL if pool else nat.EPI_STORE, slope, out, c0=64,
                       out_frame_stride=Ho * Wo * w.n_total, out_cpitch=w.n_total if pool else 2 * w.n_total)
        d.weight, d.weight_kx, d.bias = w.w_pair.data_ptr(), None, w.bias_pair.data_ptr()
        d.w_ctap, d.n_total, d.cout, d.pair_fold = 64, 2 * w.n_total, 2 * w.n_total, 1
        rc = [0]

        def run():
            rc[0] = nat.load().vad_conv_layer(C.byref(d), nat.stream_ptr())
            if rc[0] != nat.ERR_UNSUPPORTED:
                nat.check(rc[0], what or "vad_conv_layer(pair)")
        _timed(what or "vad_conv_layer", run)
        if rc[0] == 0:
            return
        if PROFILE is not None:
            PROFILE.pop()
    _gemm_layer(w, src, B, H, W, nat.EPI_POOL if pool else nat.EPI_STORE, slope, out,
                out_frame_stride=Ho * Wo * w.n_total, out_cpitch=w.n_total, what=what)


def _convt(w: GemmWeights, src, B, H, W, out, slope, what=""):
    _gemm_layer(w, src, B, H, W, nat.EPI_CONVT, slope, out, out_frame_stride=4 * H * W * w.cout, out_cpitch=w.cout,
                what=what)


def _finalize(partials, frames, tiles_per_frame, H, W, bufs: _Buffers, device) -> Tuple[torch.Tensor, torch.Tensor]:
    score = torch.empty(frames, dtype=torch.float32, device=device)
    minmax = torch.empty(frames, 2, dtype=torch.float32, device=device)
    nat.check(nat.load().vad_score_finalize(partials.data_ptr(), frames, tiles_per_frame, H, W, score.data_ptr(),
                                            minmax.data_ptr(), nat.stream_ptr()), "vad_score_finalize")
    return score, minmax


class ImageEngine:
    """ConvAutoencoder forward + fused scoring (reference models/autoencoder.py:181-221)."""

    def __init__(self, packed: Dict[str, object]) -> None:
        self.p = packed
        self.bufs = _Buffers()

    def encode(self, x: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
        """x fp32 [B,3,H,W] -> latent bf16 NHWC [B,H/16,W/16,latent]."""
        p, dev = self.p, x.device
        B, _, H, W = x.shape
        _check_hw(H, W)
        g = lambda name, shape: self.bufs.get(name, shape, torch.bfloat16, dev)
        w10, w13 = p["enc1.0"], p["enc1.3"]
        blocks = ("enc1", "enc2", "enc3", "enc4")
        if FUSE_ENC1 and FIRST_CONV_TC and PAIR_FOLD and w10.w_tc is not None and w10.cout == 32 and \
                w13.w_pair is not None and w13.n_total == 32:
            cur = g("enc1b", (B, H // 2, W // 2, 32))
            _timed("enc1.0+1.3", lambda: nat.check(
                nat.load().vad_enc1_fused(x.data_ptr(), w10.w_tc.data_ptr(), w10.bias.data_ptr(), w13.w_pair.data_ptr(),
                                          w13.bias_pair.data_ptr(), LEAKY, B, H, W, cur.data_ptr(), nat.stream_ptr()),
                "vad_enc1_fused"))
            h, w = H // 2, W // 2
            blocks = blocks[1:]
        else:
            a = g("e1a", (B, H, W, 32))
            _first_conv(w10, x, B, H, W, False, a)
            h, w, cur = H, W, a
        for blk in blocks:
            if blk != "enc1":
                w0: GemmWeights = p[f"{blk}.0"]
                nxt = g(f"{blk}a", (B, h, w, w0.n_total))
                _conv(w0, cur, B, h, w, nxt, LEAKY, what=f"{blk}.0")
                cur = nxt
            w3: GemmWeights = p[f"{blk}.3"]
            nxt = g(f"{blk}b", (B, h // 2, w // 2, w3.n_total))
            _conv(w3, cur, B, h, w, nxt, LEAKY, pool=True, what=f"{blk}.3")
            cur, h, w = nxt, h // 2, w // 2
        return cur, h, w

    def latent(self, x: torch.Tensor) -> torch.Tensor:
        x = _require_cuda_input(x, (4,))
        z, h, w = self.encode(x)
        B, C = x.shape[0], z.shape[-1]
        out = torch.empty(B, C, h, w, dtype=torch.float32, device=x.device)
        nat.check(nat.load().vad_nhwc_bf16_to_nchw_f32(z.data_ptr(), B, h, w, C, out.data_ptr(), nat.stream_ptr()),
                  "vad_nhwc_bf16_to_nchw_f32")
        return out

    def run(self, x: torch.Tensor, want_recon: bool, want_heat: bool) -> ScoreOutputs:
        x = _require_cuda_input(x, (4,))
        p, dev = self.p, x.device
        B, _, H, W = x.shape
        z, h, w = self.encode(x)
        g = lambda name, shape: self.bufs.get(name, shape, torch.bfloat16, dev)
        cur = z
        w40: GemmWeights = p["dec4.0"]
        w43: GemmWeights = p["dec4.3"]
        fuse_tail = FUSE_DEC_TAIL and w43.w_kx is not None and \
            (w40.ctap, w40.n_total, w40.cout, w43.ctap, w43.n_total) == (32, 128, 32, 32, 16)
        for blk in ("dec1", "dec2", "dec3", "dec4"):
            if blk == "dec4" and fuse_tail:
                # dec4.0 + dec4.3 + score in one kernel: the 32-channel full-resolution tensor never reaches HBM
                return _fused_image_tail(w40, w43, cur, B, h, w, x, want_recon, want_heat, self.bufs)
            wt: GemmWeights = p[f"{blk}.0"]
            up = g(f"{blk}a", (B, 2 * h, 2 * w, wt.cout))
            _convt(wt, cur, B, h, w, up, RELU, what=f"{blk}.0")
            h, w, cur = 2 * h, 2 * w, up
            if blk != "dec4":
                wc: GemmWeights = p[f"{blk}.3"]
                nxt = g(f"{blk}b", (B, h, w, wc.n_total))
                _conv(wc, cur, B, h, w, nxt, RELU, what=f"{blk}.3")
                cur = nxt
        return _score_layer(p["dec4.3"], cur, B, H, W, nat.EPI_TANH_SCORE, x, want_recon, want_heat, H, W, self.bufs,
                            "dec4.3+score")


class VideoEngine:
    """VideoAutoencoder forward + fused scoring (reference models/video_autoencoder.py:329-384)."""

    def __init__(self, packed: Dict[str, object]) -> None:
        self.p = packed
        self.bufs = _Buffers()

    # ---- pieces (also used by the sub-module wrappers) ---------------------------------------------------------
    def encode(self, x4: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
        """frames fp32 [F,3,H,W] -> bf16 NHWC [F,H/16,W/16,latent]."""
        p, dev = self.p, x4.device
        F, _, H, W = x4.shape
        _check_hw(H, W)
        g = lambda name, shape: self.bufs.get(name, shape, torch.bfloat16, dev)
        cur = g("e0", (F, H // 2, W // 2, 32))
        _first_conv(p["enc.0"], x4, F, H, W, True, cur)
        h, w = H // 2, W // 2
        for i in (4, 8, 12):
            wt: GemmWeights = p[f"enc.{i}"]
            nxt = g(f"e{i}", (F, h // 2, w // 2, wt.n_total))
            _conv(wt, cur, F, h, w, nxt, LEAKY, pool=True, what=f"encoder.{i}")
            cur, h, w = nxt, h // 2, w // 2
        return cur, h, w

    def _lstm_desc(self, layer: int, cur: torch.Tensor, B: int, T: int, h: int, w: int):
        p, dev = self.p, cur.device
        wt: GemmWeights = p[f"lstm.{layer}"]
        hid = wt.cout
        cin = wt.ctap - hid
        hseq = self.bufs.get(f"hseq{layer}", (B, T, h, w, hid), torch.bfloat16, dev)
        cst = self.bufs.get(f"c{layer}", (B, h, w, hid), torch.float32, dev)
        d = nat.ConvDesc()
        d.src0, d.src1, d.out = cur.data_ptr(), hseq.data_ptr(), hseq.data_ptr()
        d.c0, d.c1, d.T0, d.T1 = cin, hid, T, T
        d.B, d.H, d.W, d.ntaps = B, h, w, 9
        d.weight, d.bias, d.w_ctap = wt.w.data_ptr(), wt.bias.data_ptr(), wt.ctap
        d.n_total, d.cout, d.epilogue, d.slope = wt.n_total, hid, nat.EPI_LSTM, IDENT
        d.out_frame_stride, d.out_cpitch = T * h * w * hid, hid
        d.c_state = cst.data_ptr()
        return d, hseq

    def convlstm(self, seq: torch.Tensor, B: int, T: int, h: int, w: int) -> torch.Tensor:
        """seq bf16 [B,T,h,w,C] -> last layer's hidden sequence bf16 [B,T,h,w,hid] (zero initial state)."""
        p = self.p
        cur = seq
        layer = 0
        while layer < p["lstm_layers"]:
            d, hseq = self._lstm_desc(layer, cur, B, T, h, w)
            if FUSE_LSTM_LAYERS and layer + 1 < p["lstm_layers"]:
                # two layers as one wavefront launch (layer 2's step t runs next to layer 1's step t+1)
                d2, hseq2 = self._lstm_desc(layer + 1, hseq, B, T, h, w)
                rc = [0]

                def both():
                    rc[0] = nat.load().vad_convlstm2_sequence(C.byref(d), C.byref(d2), T, nat.stream_ptr())
                    if rc[0] != nat.ERR_UNSUPPORTED:
                        nat.check(rc[0], f"convlstm.{layer}+{layer + 1}")
                _timed(f"convlstm.{layer}+{layer + 1}", both)
                if rc[0] == 0:
                    cur = hseq2
                    layer += 2
                    continue
                if PROFILE is not None:
                    PROFILE.pop()  # nothing was launched
            _timed(f"convlstm.{layer}", lambda: nat.check(
                nat.load().vad_convlstm_sequence(C.byref(d), T, nat.stream_ptr()), f"convlstm.{layer}"))
            cur = hseq
            layer += 1
        return cur

    def project(self, seq: torch.Tensor, F: int, h: int, w: int) -> torch.Tensor:
        if "proj" not in self.p:
            return seq
        wt: GemmWeights = self.p["proj"]
        out = self.bufs.get("proj", (F, h, w, wt.n_total), torch.bfloat16, seq.device)
        _conv(wt, seq, F, h, w, out, IDENT, what="proj")
        return out

    def decode_to(self, z: torch.Tensor, F: int, h: int, w: int, layers=(0, 3, 6)) -> Tuple[torch.Tensor, int, int]:
        """bf16 NHWC [F,h,w,latent] -> input of the last ConvT, bf16 NHWC [F,8h,8w,32] (or of an earlier one)."""
        cur = z
        for i in layers:
            wt: GemmWeights = self.p[f"dec.{i}"]
            up = self.bufs.get(f"d{i}", (F, 2 * h, 2 * w, wt.cout), torch.bfloat16, z.device)
            _convt(wt, cur, F, h, w, up, RELU, what=f"decoder.{i}")
            cur, h, w = up, 2 * h, 2 * w
        return cur, h, w

    def run(self, x: torch.Tensor, want_recon: bool, want_heat: bool) -> ScoreOutputs:
        x = _require_cuda_input(x, (5,))
        B, T, Cin, H, W = x.shape
        F = B * T
        dev = x.device
        x4 = x.view(F, Cin, H, W)
        z, h, w = self.encode(x4)
        seq = self.convlstm(z.view(B, T, h, w, z.shape[-1]), B, T, h, w)
        zp = self.project(seq.view(F, h, w, seq.shape[-1]), F, h, w)
        return self.decode_and_score(zp, F, h, w, x4, want_recon, want_heat)

    def decode_and_score(self, zp: torch.Tensor, F: int, h: int, w: int, x4: torch.Tensor, want_recon: bool,
                         want_heat: bool) -> ScoreOutputs:
        """Decoder + fused scoring of F frames: zp bf16 NHWC [F,h,w,latent], x4 fp32 [F,3,16h,16w]."""
        H, W = x4.shape[-2], x4.shape[-1]
        w6: GemmWeights = self.p["dec.6"]
        w9: GemmWeights = self.p["dec.9"]
        if FUSE_DEC_TAIL and (w6.ctap, w6.n_total, w6.cout, w9.ctap, w9.n_total) == (64, 128, 32, 32, 16):
            # decoder.6 + decoder.9 + score in one kernel: the 32-channel half-resolution tensor never reaches HBM
            d, hd, wd = self.decode_to(zp, F, h, w, layers=(0, 3))
            return _fused_tail(w6, w9, d, F, hd, wd, x4, want_recon, want_heat, H, W, self.bufs)
        d, hd, wd = self.decode_to(zp, F, h, w)
        return _score_layer(w9, d, F, hd, wd, nat.EPI_CONVT_TANH_SCORE, x4, want_recon, want_heat, H, W,
                            self.bufs, "decoder.9+score")
