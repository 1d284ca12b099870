"""Binding of the model-level entry points of libvad_b200 (one C call per reference method).

`ImageEngine` / `VideoEngine` own the prepared weights of one model (or of one stand-alone sub-module) as the
`vad_image_model` / `vad_video_model` structs of include/vad_b200.h and turn each reference method into ONE call:
the whole layer schedule is enqueued by the library on the current CUDA stream of the input's device.  PyTorch only
provides device memory here: outputs are fresh tensors, and the per-call workspace (intermediate activations, score
partials, ConvLSTM step counters) comes from torch's stream-aware caching allocator — so concurrent calls on different
streams or threads never share a buffer, and nothing is cached per shape.

There is no CPU fallback: CPU tensors, a missing library or unsupported shapes raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _native as nat
from ._prepare import FirstConvWeights, GemmWeights

# VAD_FUSE_DEC=0: run the decoders' last two layers one by one instead of the fused tail kernels
# (video: vad_convt2_score, image: vad_convt_conv_score).  Tests flip the module attribute.
FUSE_DEC_TAIL = os.environ.get("VAD_FUSE_DEC", "1") != "0"
# VAD_LSTM2=0: one launch per ConvLSTM layer instead of the two-layer wavefront kernel
FUSE_LSTM_LAYERS = os.environ.get("VAD_LSTM2", "1") != "0"
# VAD_FIRST_TC=0: CUDA-core first conv (fp32 operands) instead of the tensor-core one
FIRST_CONV_TC = os.environ.get("VAD_FIRST_TC", "1") != "0"
# VAD_FUSE_ENC1=0: image enc1.0 and enc1.3 as two launches instead of the fused kernel (vad_enc1_fused)
FUSE_ENC1 = os.environ.get("VAD_FUSE_ENC1", "1") != "0"
# VAD_PAIR_FOLD=0: the 3x3 layers with 32 input channels on the ordinary view instead of the pixel-pair folded one
PAIR_FOLD = os.environ.get("VAD_PAIR_FOLD", "1") != "0"
# VAD_FIRST_PF=0: the video encoder's pooled first conv on the one-row-per-input-pixel kernel (vad_first_conv_tc)
# instead of the pool-folded one (vad_first_conv_pool)
FIRST_CONV_POOL_FOLD = os.environ.get("VAD_FIRST_PF", "1") != "0"


@dataclass
class ScoreOutputs:
    score: torch.Tensor                 # [frames] fp32: mean over (C,H,W) of (x - recon)^2
    minmax: torch.Tensor                # [frames, 2] fp32: min / max of the per-pixel map (heat-map normalisation)
    heat: Optional[torch.Tensor]        # [frames, H, W] fp32 per-pixel channel-mean squared error
    recon: Optional[torch.Tensor]       # [frames, 3, H, W] fp32
    latent: Optional[torch.Tensor] = None  # [frames, latent, H/16, W/16] fp32 (image model, on request)
    heat_u8: Optional[torch.Tensor] = None  # [frames, H, W] uint8: create_heatmap's per-frame normalisation of `heat`


def _require_cuda_input(x: torch.Tensor, ndim: Tuple[int, ...]) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("vad_b200 scoring path is CUDA-only (sm_100a kernels, no CPU fallback); "
                           f"got a tensor on {x.device}")
    if x.dim() not in ndim:
        raise RuntimeError(f"expected a {ndim}-D input, got shape {tuple(x.shape)}")
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


def _require_u8_frames(frames: torch.Tensor, ndim: Tuple[int, ...]) -> torch.Tensor:
    if not frames.is_cuda:
        raise RuntimeError("vad_b200 scoring path is CUDA-only (sm_100a kernels, no CPU fallback); "
                           f"got a tensor on {frames.device}")
    if frames.dtype != torch.uint8 or frames.dim() not in ndim or frames.shape[-1] != 3:
        raise RuntimeError(f"expected uint8 RGB frames [..., H, W, 3] with {ndim[0]} dimensions, got "
                           f"{frames.dtype} {tuple(frames.shape)}")
    return frames.contiguous()


def _check_hw(H: int, W: int) -> None:
    if H % 16 or W % 16 or H <= 0 or W <= 0:
        # the reference fails here too (x - recon shape mismatch); fail before launching anything
        raise RuntimeError(f"input height/width must be positive multiples of 16, got {H}x{W}")


def _gemm_struct(w: GemmWeights) -> "nat.GemmW":
    g = nat.GemmW()
    g.w, g.w_kx, g.bias = w.w.data_ptr(), nat.ptr(w.w_kx), w.bias.data_ptr()
    g.ntaps, g.ctap, g.n_total, g.cout = w.ntaps, w.ctap, w.n_total, w.cout
    if PAIR_FOLD and w.w_pair is not None:
        g.w_pair, g.bias_pair = w.w_pair.data_ptr(), w.bias_pair.data_ptr()
    return g


def _first_struct(w: FirstConvWeights) -> "nat.FirstW":
    f = nat.FirstW()
    f.w, f.bias, f.cout = w.w.data_ptr(), w.bias.data_ptr(), w.cout
    f.w_tc = nat.ptr(w.w_tc) if FIRST_CONV_TC else None
    f.w_pf = nat.ptr(w.w_pf) if (FIRST_CONV_TC and FIRST_CONV_POOL_FOLD) else None
    return f


def _flags() -> int:
    return (0 if FUSE_DEC_TAIL else nat.FLAG_NO_FUSED_TAIL) | (0 if FUSE_LSTM_LAYERS else nat.FLAG_NO_LSTM_WAVEFRONT) | \
        (0 if FUSE_ENC1 else nat.FLAG_NO_FUSED_ENC1)


def _packed_device(packed: Dict[str, object]) -> torch.device:
    for v in packed.values():
        if isinstance(v, (GemmWeights, FirstConvWeights)):
            return v.bias.device
    raise RuntimeError("no prepared weights")


class _Call:
    """One model-level call: makes the tensors' device current, picks its stream and allocates the workspace."""

    def __init__(self, device: torch.device, ws_bytes: int, what: str):
        if ws_bytes == 0:
            raise RuntimeError(f"libvad_b200: {what}: unsupported model / shape (workspace query returned 0)")
        self.guard = torch.cuda.device(device)
        self.device = device
        self.ws_bytes = ws_bytes

    def __enter__(self):
        self.guard.__enter__()
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.stream = torch.cuda.current_stream(self.device).cuda_stream
        return self

    def __exit__(self, *exc):
        return self.guard.__exit__(*exc)


class ImageEngine:
    """ConvAutoencoder forward + fused scoring (reference models/autoencoder.py:181-221) behind `vad_image_forward`;
    a stand-alone `Encoder` / `Decoder` builds one from its own parameters (only that half is populated)."""

    _ENC = ("enc1.3", "enc2.0", "enc2.3", "enc3.0", "enc3.3", "enc4.0", "enc4.3")
    _DEC = ("dec1.0", "dec1.3", "dec2.0", "dec2.3", "dec3.0", "dec3.3", "dec4.0", "dec4.3")

    def __init__(self, packed: Dict[str, object]) -> None:
        self.p = packed  # keeps the device tensors behind the raw pointers alive
        self.device = _packed_device(packed)
        m = nat.ImageModel()
        m.has_encoder = 1 if "enc1.0" in packed else 0
        m.has_decoder = 1 if "dec1.0" in packed else 0
        if m.has_encoder:
            m.enc1_0 = _first_struct(packed["enc1.0"])
            for i, k in enumerate(self._ENC):
                m.enc[i] = _gemm_struct(packed[k])
        if m.has_decoder:
            for i, k in enumerate(self._DEC):
                m.dec[i] = _gemm_struct(packed[k])
        self.m = m
        self.lib = nat.load()

    @property
    def latent_dim(self) -> int:
        return (self.p["enc4.3"].n_total if self.m.has_encoder else self.p["dec1.0"].ctap)

    def _ws(self, op: int, B: int, H: int, W: int) -> int:
        return int(self.lib.vad_image_workspace_bytes(C.byref(self.m), op, B, H, W))

    def run(self, x: torch.Tensor, want_recon: bool, want_heat: bool, want_latent: bool = False) -> ScoreOutputs:
        x = _require_cuda_input(x, (4,))
        B, cin, H, W = x.shape
        if cin != 3:
            raise RuntimeError(f"expected 3 input channels, got {cin}")
        _check_hw(H, W)
        dev = x.device
        self.m.flags = _flags()
        score = torch.empty(B, dtype=torch.float32, device=dev)
        minmax = torch.empty(B, 2, dtype=torch.float32, device=dev)
        heat = torch.empty(B, H, W, dtype=torch.float32, device=dev) if want_heat else None
        recon = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev) if want_recon else None
        latent = torch.empty(B, self.latent_dim, H // 16, W // 16, dtype=torch.float32, device=dev) if want_latent else None
        with _Call(dev, self._ws(nat.OP_FORWARD, B, H, W), "vad_image_forward") as c:
            nat.check(self.lib.vad_image_forward(C.byref(self.m), x.data_ptr(), B, H, W, nat.ptr(recon), nat.ptr(latent),
                                                 score.data_ptr(), minmax.data_ptr(), nat.ptr(heat), c.ws.data_ptr(),
                                                 c.ws_bytes, c.stream), "vad_image_forward")
        return ScoreOutputs(score, minmax, heat, recon, latent)

    def run_u8(self, frames: torch.Tensor, want_recon: bool, want_heat: bool, want_heat_u8: bool) -> ScoreOutputs:
        """uint8 RGB frames [B,H,W,3] (as decoded) -> scores (+ heat maps): normalisation happens on the device."""
        frames = _require_u8_frames(frames, (4,))
        B, H, W, _ = frames.shape
        _check_hw(H, W)
        dev = frames.device
        self.m.flags = _flags()
        score = torch.empty(B, dtype=torch.float32, device=dev)
        minmax = torch.empty(B, 2, dtype=torch.float32, device=dev)
        heat = torch.empty(B, H, W, dtype=torch.float32, device=dev) if want_heat else None
        heat8 = torch.empty(B, H, W, dtype=torch.uint8, device=dev) if want_heat_u8 else None
        recon = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev) if want_recon else None
        with _Call(dev, self._ws(nat.OP_FORWARD_U8, B, H, W), "vad_image_forward_u8") as c:
            nat.check(self.lib.vad_image_forward_u8(C.byref(self.m), frames.data_ptr(), B, H, W, nat.ptr(recon), None,
                                                    score.data_ptr(), minmax.data_ptr(), nat.ptr(heat), nat.ptr(heat8),
                                                    c.ws.data_ptr(), c.ws_bytes, c.stream), "vad_image_forward_u8")
        return ScoreOutputs(score, minmax, heat, recon, None, heat8)

    def latent(self, x: torch.Tensor) -> torch.Tensor:
        """Encoder.forward / get_latent: fp32 [B,3,H,W] -> fp32 [B,latent,H/16,W/16]."""
        x = _require_cuda_input(x, (4,))
        B, cin, H, W = x.shape
        if cin != 3:
            raise RuntimeError(f"expected 3 input channels, got {cin}")
        _check_hw(H, W)
        out = torch.empty(B, self.latent_dim, H // 16, W // 16, dtype=torch.float32, device=x.device)
        self.m.flags = _flags()
        with _Call(x.device, self._ws(nat.OP_ENCODE, B, H, W), "vad_image_forward(latent)") as c:
            nat.check(self.lib.vad_image_forward(C.byref(self.m), x.data_ptr(), B, H, W, None, out.data_ptr(), None, None,
                                                 None, c.ws.data_ptr(), c.ws_bytes, c.stream), "vad_image_forward")
        return out

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """Decoder.forward: fp32 [B,latent,h,w] -> fp32 [B,3,16h,16w]."""
        z = _require_cuda_input(z, (4,))
        B, cz, h, w = z.shape
        if cz != self.latent_dim:
            raise RuntimeError(f"expected {self.latent_dim} latent channels, got {cz}")
        self.m.flags = _flags()
        recon = torch.empty(B, 3, 16 * h, 16 * w, dtype=torch.float32, device=z.device)
        with _Call(z.device, self._ws(nat.OP_DECODE, B, h, w), "vad_image_decode") as c:
            nat.check(self.lib.vad_image_decode(C.byref(self.m), z.data_ptr(), B, h, w, recon.data_ptr(), c.ws.data_ptr(),
                                                c.ws_bytes, c.stream), "vad_image_decode")
        return recon


class VideoEngine:
    """VideoAutoencoder forward + fused scoring (reference models/video_autoencoder.py:329-384) behind
    `vad_video_forward`; stand-alone sub-modules populate only their part of the struct."""

    def __init__(self, packed: Dict[str, object]) -> None:
        self.p = packed
        self.device = _packed_device(packed)
        m = nat.VideoModel()
        m.has_encoder = 1 if "enc.0" in packed else 0
        m.has_decoder = 1 if "dec.0" in packed else 0
        m.lstm_layers = int(packed.get("lstm_layers", 0))
        if m.lstm_layers > nat.MAX_LSTM_LAYERS:
            raise ValueError(f"at most {nat.MAX_LSTM_LAYERS} ConvLSTM layers are supported")
        if m.has_encoder:
            m.enc0 = _first_struct(packed["enc.0"])
            for i, k in enumerate((4, 8, 12)):
                m.enc[i] = _gemm_struct(packed[f"enc.{k}"])
        for l in range(m.lstm_layers):
            m.lstm[l] = _gemm_struct(packed[f"lstm.{l}"])
        m.has_proj = 1 if "proj" in packed else 0
        if m.has_proj:
            m.proj = _gemm_struct(packed["proj"])
        if m.has_decoder:
            for i, k in enumerate((0, 3, 6, 9)):
                m.dec[i] = _gemm_struct(packed[f"dec.{k}"])
        self.m = m
        self.lib = nat.load()

    def _ws(self, op: int, B: int, T: int, H: int, W: int) -> int:
        return int(self.lib.vad_video_workspace_bytes(C.byref(self.m), op, B, T, H, W))

    @property
    def latent_dim(self) -> int:
        return self.p["enc.12"].n_total if self.m.has_encoder else self.p["dec.0"].ctap

    @property
    def hidden_dim(self) -> int:
        return self.p[f"lstm.{self.m.lstm_layers - 1}"].cout

    def _outputs(self, F: int, H: int, W: int, want_recon: bool, want_heat: bool, dev):
        score = torch.empty(F, dtype=torch.float32, device=dev)
        minmax = torch.empty(F, 2, dtype=torch.float32, device=dev)
        heat = torch.empty(F, H, W, dtype=torch.float32, device=dev) if want_heat else None
        recon = torch.empty(F, 3, H, W, dtype=torch.float32, device=dev) if want_recon else None
        return score, minmax, heat, recon

    def run(self, x: torch.Tensor, want_recon: bool, want_heat: bool) -> ScoreOutputs:
        x = _require_cuda_input(x, (5,))
        B, T, cin, H, W = x.shape
        if cin != 3:
            raise RuntimeError(f"expected 3 input channels, got {cin}")
        _check_hw(H, W)
        self.m.flags = _flags()
        score, minmax, heat, recon = self._outputs(B * T, H, W, want_recon, want_heat, x.device)
        with _Call(x.device, self._ws(nat.OP_FORWARD, B, T, H, W), "vad_video_forward") as c:
            nat.check(self.lib.vad_video_forward(C.byref(self.m), x.data_ptr(), B, T, H, W, nat.ptr(recon),
                                                 score.data_ptr(), minmax.data_ptr(), nat.ptr(heat), c.ws.data_ptr(),
                                                 c.ws_bytes, c.stream), "vad_video_forward")
        return ScoreOutputs(score, minmax, heat, recon)

    def run_u8(self, frames: torch.Tensor, want_recon: bool, want_heat: bool, want_heat_u8: bool) -> ScoreOutputs:
        """uint8 RGB clips [B,T,H,W,3] (as decoded) -> per-frame scores (+ heat maps)."""
        frames = _require_u8_frames(frames, (5,))
        B, T, H, W, _ = frames.shape
        _check_hw(H, W)
        self.m.flags = _flags()
        score, minmax, heat, recon = self._outputs(B * T, H, W, want_recon, want_heat, frames.device)
        heat8 = torch.empty(B * T, H, W, dtype=torch.uint8, device=frames.device) if want_heat_u8 else None
        with _Call(frames.device, self._ws(nat.OP_FORWARD_U8, B, T, H, W), "vad_video_forward_u8") as c:
            nat.check(self.lib.vad_video_forward_u8(C.byref(self.m), frames.data_ptr(), B, T, H, W, nat.ptr(recon),
                                                    score.data_ptr(), minmax.data_ptr(), nat.ptr(heat), nat.ptr(heat8),
                                                    c.ws.data_ptr(), c.ws_bytes, c.stream), "vad_video_forward_u8")
        return ScoreOutputs(score, minmax, heat, recon, None, heat8)

    def encode(self, x4: torch.Tensor, want_f32: bool = False, want_bf16: bool = True):
        """frames fp32 [F,3,H,W] -> (bf16 NHWC [F,h,w,latent] or None, fp32 NCHW [F,latent,h,w] or None)."""
        x4 = _require_cuda_input(x4, (4,))
        F, cin, H, W = x4.shape
        if cin != 3:
            raise RuntimeError(f"expected 3 input channels, got {cin}")
        _check_hw(H, W)
        h, w, Cz = H // 16, W // 16, self.latent_dim
        zb = torch.empty(F, h, w, Cz, dtype=torch.bfloat16, device=x4.device) if want_bf16 else None
        zf = torch.empty(F, Cz, h, w, dtype=torch.float32, device=x4.device) if want_f32 else None
        with _Call(x4.device, self._ws(nat.OP_ENCODE, F, 1, H, W), "vad_video_encode") as c:
            nat.check(self.lib.vad_video_encode(C.byref(self.m), x4.data_ptr(), F, H, W, nat.ptr(zf), nat.ptr(zb),
                                                c.ws.data_ptr(), c.ws_bytes, c.stream), "vad_video_encode")
        return zb, zf

    def score_latents(self, z: torch.Tensor, x: torch.Tensor, want_recon: bool, want_heat: bool) -> ScoreOutputs:
        """ConvLSTM -> proj -> decoder -> scoring from cached encoder features: z bf16 [B,T,h,w,latent],
        x fp32 [B,T,3,16h,16w]."""
        if z.dtype != torch.bfloat16 or z.dim() != 5 or not z.is_cuda:
            raise RuntimeError("score_latents expects a CUDA bf16 tensor [B,T,h,w,C]")
        z = z.contiguous()
        x = _require_cuda_input(x, (5,))
        B, T, h, w, _ = z.shape
        if tuple(x.shape) != (B, T, 3, 16 * h, 16 * w):
            raise RuntimeError(f"frames {tuple(x.shape)} do not match latents {tuple(z.shape)}")
        self.m.flags = _flags()
        score, minmax, heat, recon = self._outputs(B * T, 16 * h, 16 * w, want_recon, want_heat, x.device)
        with _Call(x.device, self._ws(nat.OP_SCORE_LATENTS, B, T, h, w), "vad_video_score_latents") as c:
            nat.check(self.lib.vad_video_score_latents(C.byref(self.m), z.data_ptr(), x.data_ptr(), B, T, h, w,
                                                       nat.ptr(recon), score.data_ptr(), minmax.data_ptr(), nat.ptr(heat),
                                                       c.ws.data_ptr(), c.ws_bytes, c.stream), "vad_video_score_latents")
        return ScoreOutputs(score, minmax, heat, recon)

    def decode(self, z4: torch.Tensor) -> torch.Tensor:
        """VideoDecoder.forward on frames: fp32 [F,latent,h,w] -> fp32 [F,3,16h,16w]."""
        z4 = _require_cuda_input(z4, (4,))
        F, cz, h, w = z4.shape
        if cz != self.latent_dim:
            raise RuntimeError(f"expected {self.latent_dim} latent channels, got {cz}")
        self.m.flags = _flags()
        recon = torch.empty(F, 3, 16 * h, 16 * w, dtype=torch.float32, device=z4.device)
        with _Call(z4.device, self._ws(nat.OP_DECODE, F, 1, h, w), "vad_video_decode") as c:
            nat.check(self.lib.vad_video_decode(C.byref(self.m), z4.data_ptr(), F, h, w, recon.data_ptr(),
                                                c.ws.data_ptr(), c.ws_bytes, c.stream), "vad_video_decode")
        return recon

    def convlstm(self, x5: torch.Tensor):
        """ConvLSTM.forward (zero initial state): fp32 [B,T,C,h,w] -> (out fp32 [B,T,hid,h,w], c_last fp32 [B,hid,h,w])."""
        x5 = _require_cuda_input(x5, (5,))
        B, T, cin, h, w = x5.shape
        w0: GemmWeights = self.p["lstm.0"]
        if cin != w0.ctap - w0.cout:
            raise RuntimeError(f"expected {w0.ctap - w0.cout} input channels, got {cin}")
        self.m.flags = _flags()
        hid = self.hidden_dim
        out = torch.empty(B, T, hid, h, w, dtype=torch.float32, device=x5.device)
        c_last = torch.empty(B, hid, h, w, dtype=torch.float32, device=x5.device)
        with _Call(x5.device, self._ws(nat.OP_CONVLSTM, B, T, h, w), "vad_convlstm_forward") as c:
            nat.check(self.lib.vad_convlstm_forward(C.byref(self.m), x5.data_ptr(), B, T, h, w, out.data_ptr(), None,
                                                    c_last.data_ptr(), c.ws.data_ptr(), c.ws_bytes, c.stream),
                      "vad_convlstm_forward")
        return out, c_last


class CellEngine:
    """One ConvLSTMCell step (reference models/video_autoencoder.py:54-85) behind `vad_convlstm_cell`."""

    def __init__(self, w: GemmWeights) -> None:
        self.w = w
        self.device = w.bias.device
        self.g = _gemm_struct(w)
        self.lib = nat.load()

    def step(self, x: torch.Tensor, h_cur: torch.Tensor, c_cur: torch.Tensor):
        x = _require_cuda_input(x, (4,))
        h_cur = _require_cuda_input(h_cur, (4,))
        c_cur = _require_cuda_input(c_cur, (4,))
        B, cin, h, w = x.shape
        hid = self.w.cout
        if cin != self.w.ctap - hid or tuple(h_cur.shape) != (B, hid, h, w) or tuple(c_cur.shape) != (B, hid, h, w):
            raise RuntimeError(f"ConvLSTMCell shapes: x {tuple(x.shape)}, h {tuple(h_cur.shape)}, c {tuple(c_cur.shape)} "
                               f"do not match input_dim {self.w.ctap - hid} / hidden_dim {hid}")
        h_next = torch.empty_like(h_cur)
        c_next = torch.empty_like(c_cur)
        nbytes = int(self.lib.vad_convlstm_cell_workspace_bytes(C.byref(self.g), B, h, w))
        with _Call(x.device, nbytes, "vad_convlstm_cell") as c:
            nat.check(self.lib.vad_convlstm_cell(C.byref(self.g), x.data_ptr(), h_cur.data_ptr(), c_cur.data_ptr(), B, h, w,
                                                 h_next.data_ptr(), c_next.data_ptr(), c.ws.data_ptr(), c.ws_bytes,
                                                 c.stream), "vad_convlstm_cell")
        return h_next, c_next
