"""ctypes binding of libvad_b200.so (C ABI declared in include/vad_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_DIR, "libvad_b200.so")

ERR_UNSUPPORTED = -3
EPI_STORE, EPI_POOL, EPI_CONVT, EPI_LSTM, EPI_TANH_SCORE, EPI_CONVT_TANH_SCORE = range(6)

# every symbol include/vad_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "vad_error_string", "vad_version", "vad_launch_count", "vad_debug_last_trap", "vad_debug_set_timeline", "vad_debug_set_kx", "vad_debug_set_lstm_mode", "vad_conv_layer", "vad_conv_layer_tiles", "vad_convt2_score", "vad_convt2_score_tiles", "vad_convt_conv_score", "vad_convt_conv_score_tiles", "vad_convlstm_sequence", "vad_convlstm2_sequence", "vad_conv_m_tiles", "vad_first_conv", "vad_first_conv_tc", "vad_first_conv_pool", "vad_enc1_fused",
    "vad_score_finalize", "vad_score_scratch_bytes", "vad_score", "vad_nhwc_bf16_to_nchw_f32",
    "vad_nchw_f32_to_nhwc_bf16", "vad_heatmap_u8", "vad_u8_hwc_to_f32_nchw", "vad_f32_nchw_to_u8_hwc",
    "vad_heatmap_jet_rgb", "vad_compose_panel", "vad_ssim_scratch_bytes", "vad_ssim_loss",
    # model-level entry points (one call per reference method)
    "vad_image_workspace_bytes", "vad_image_forward", "vad_image_forward_u8", "vad_image_decode",
    "vad_video_workspace_bytes", "vad_video_forward", "vad_video_forward_u8", "vad_video_encode", "vad_video_score_latents", "vad_video_decode", "vad_convlstm_forward",
    "vad_convlstm_cell_workspace_bytes", "vad_convlstm_cell", "vad_profile_enable", "vad_profile_dump",
)

FLAG_NO_FUSED_TAIL, FLAG_NO_LSTM_WAVEFRONT, FLAG_NO_FUSED_ENC1 = 1, 2, 4
OP_FORWARD, OP_ENCODE, OP_DECODE, OP_CONVLSTM, OP_SCORE_LATENTS, OP_FORWARD_U8 = range(6)
MAX_LSTM_LAYERS = 8


class ConvDesc(C.Structure):
    """Mirror of `struct vad_conv_desc` (include/vad_b200.h)."""

    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p),
        ("c0", C.c_int), ("c1", C.c_int),
        ("T0", C.c_int), ("T1", C.c_int),
        ("t0", C.c_int), ("t1", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("ntaps", C.c_int),
        ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("w_ctap", C.c_int),
        ("n_total", C.c_int), ("cout", C.c_int), ("epilogue", C.c_int),
        ("slope", C.c_float),
        ("out", C.c_void_p), ("out_frame_stride", C.c_longlong), ("out_cpitch", C.c_int),
        ("c_state", C.c_void_p), ("lstm_first", C.c_int),
        ("x", C.c_void_p), ("recon", C.c_void_p), ("heat", C.c_void_p), ("partials", C.c_void_p),
        ("weight_kx", C.c_void_p),
        ("scratch", C.c_void_p),
        ("pair_fold", C.c_int),
    ]


class GemmW(C.Structure):
    """Mirror of `struct vad_gemm_weights`."""
    _fields_ = [("w", C.c_void_p), ("w_kx", C.c_void_p), ("bias", C.c_void_p),
                ("ntaps", C.c_int), ("ctap", C.c_int), ("n_total", C.c_int), ("cout", C.c_int),
                ("w_pair", C.c_void_p), ("bias_pair", C.c_void_p)]


class FirstW(C.Structure):
    """Mirror of `struct vad_first_weights`."""
    _fields_ = [("w", C.c_void_p), ("w_tc", C.c_void_p), ("w_pf", C.c_void_p), ("bias", C.c_void_p), ("cout", C.c_int)]


class ImageModel(C.Structure):
    """Mirror of `struct vad_image_model`."""
    _fields_ = [("has_encoder", C.c_int), ("has_decoder", C.c_int), ("flags", C.c_int), ("reserved", C.c_int),
                ("enc1_0", FirstW), ("enc", GemmW * 7), ("dec", GemmW * 8)]


class VideoModel(C.Structure):
    """Mirror of `struct vad_video_model`."""
    _fields_ = [("has_encoder", C.c_int), ("lstm_layers", C.c_int), ("has_proj", C.c_int), ("has_decoder", C.c_int),
                ("flags", C.c_int), ("reserved", C.c_int),
                ("enc0", FirstW), ("enc", GemmW * 3), ("lstm", GemmW * MAX_LSTM_LAYERS), ("proj", GemmW),
                ("dec", GemmW * 4)]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the scoring path)")
    lib = C.CDLL(LIB_PATH)
    lib.vad_error_string.restype = C.c_char_p
    lib.vad_error_string.argtypes = [C.c_int]
    lib.vad_version.restype = C.c_int
    lib.vad_launch_count.restype = C.c_ulonglong
    lib.vad_debug_set_timeline.argtypes = [C.c_void_p]
    lib.vad_debug_set_kx.argtypes = [C.c_int]
    lib.vad_debug_set_lstm_mode.argtypes = [C.c_int]
    lib.vad_conv_layer.argtypes = [C.POINTER(ConvDesc), C.c_void_p]
    lib.vad_conv_layer_tiles.argtypes = [C.POINTER(ConvDesc)]
    lib.vad_convt2_score.argtypes = [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vad_convt2_score_tiles.argtypes = [C.POINTER(ConvDesc)]
    lib.vad_convt_conv_score.argtypes = [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vad_convt_conv_score_tiles.argtypes = [C.POINTER(ConvDesc)]
    lib.vad_convlstm_sequence.argtypes = [C.POINTER(ConvDesc), C.c_int, C.c_void_p]
    lib.vad_convlstm2_sequence.argtypes = [C.POINTER(ConvDesc), C.POINTER(ConvDesc), C.c_int, C.c_void_p]
    lib.vad_conv_m_tiles.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    lib.vad_first_conv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p, C.c_void_p]
    lib.vad_first_conv_tc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p]
    lib.vad_first_conv_pool.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]
    lib.vad_enc1_fused.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p, C.c_void_p]
    lib.vad_score_finalize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    lib.vad_score_scratch_bytes.restype = C.c_size_t
    lib.vad_score_scratch_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.vad_score.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]
    lib.vad_nhwc_bf16_to_nchw_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.vad_nchw_f32_to_nhwc_bf16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.vad_heatmap_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.vad_u8_hwc_to_f32_nchw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.vad_f32_nchw_to_u8_hwc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.vad_heatmap_jet_rgb.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.vad_compose_panel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p]
    lib.vad_ssim_scratch_bytes.restype = C.c_size_t
    lib.vad_ssim_scratch_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.vad_ssim_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    P, I, Z = C.c_void_p, C.c_int, C.c_size_t
    lib.vad_image_workspace_bytes.restype = Z
    lib.vad_image_workspace_bytes.argtypes = [C.POINTER(ImageModel), I, I, I, I]
    lib.vad_image_forward.argtypes = [C.POINTER(ImageModel), P, I, I, I, P, P, P, P, P, P, Z, P]
    lib.vad_image_forward_u8.argtypes = [C.POINTER(ImageModel), P, I, I, I, P, P, P, P, P, P, P, Z, P]
    lib.vad_video_forward_u8.argtypes = [C.POINTER(VideoModel), P, I, I, I, I, P, P, P, P, P, P, Z, P]
    lib.vad_image_decode.argtypes = [C.POINTER(ImageModel), P, I, I, I, P, P, Z, P]
    lib.vad_video_workspace_bytes.restype = Z
    lib.vad_video_workspace_bytes.argtypes = [C.POINTER(VideoModel), I, I, I, I, I]
    lib.vad_video_forward.argtypes = [C.POINTER(VideoModel), P, I, I, I, I, P, P, P, P, P, Z, P]
    lib.vad_video_encode.argtypes = [C.POINTER(VideoModel), P, I, I, I, P, P, P, Z, P]
    lib.vad_video_score_latents.argtypes = [C.POINTER(VideoModel), P, P, I, I, I, I, P, P, P, P, P, Z, P]
    lib.vad_video_decode.argtypes = [C.POINTER(VideoModel), P, I, I, I, P, P, Z, P]
    lib.vad_convlstm_forward.argtypes = [C.POINTER(VideoModel), P, I, I, I, I, P, P, P, P, Z, P]
    lib.vad_convlstm_cell_workspace_bytes.restype = Z
    lib.vad_convlstm_cell_workspace_bytes.argtypes = [C.POINTER(GemmW), I, I, I]
    lib.vad_convlstm_cell.argtypes = [C.POINTER(GemmW), P, P, P, I, I, I, P, P, P, Z, P]
    lib.vad_profile_enable.argtypes = [I]
    lib.vad_profile_dump.argtypes = [C.c_char_p, Z]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vad_error_string(rc).decode()
        trap = (C.c_ulonglong * 4)()
        if rc > 0 and load().vad_debug_last_trap(trap) == 0:
            msg += f" [mbarrier wait timed out: tag={trap[0]} block={trap[1]} thread={trap[2]} parity={trap[3]}]"
        raise RuntimeError(f"libvad_b200: {what} failed with code {rc}: {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def profile_enable(on: bool) -> None:
    """Bracket every layer launch of the model-level calls with CUDA events (bench.py's per-kernel roofline)."""
    load().vad_profile_enable(1 if on else 0)


def profile_dump():
    """[(layer name, ms)] in launch order since the last dump; synchronises the recorded events."""
    buf = C.create_string_buffer(1 << 20)
    n = load().vad_profile_dump(buf, len(buf))
    if n < 0:
        check(n, "vad_profile_dump")
    out = []
    for line in buf.value.decode().splitlines():
        name, _, ms = line.partition("\t")
        out.append((name, float(ms)))
    return out


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def launch_count() -> int:
    return int(load().vad_launch_count())


def conv_layer(desc: ConvDesc, what: str = "vad_conv_layer") -> None:
    check(load().vad_conv_layer(C.byref(desc), stream_ptr()), what)


def layer_tiles(desc: ConvDesc) -> int:
    """M tiles vad_conv_layer will use for `desc` (rows of the per-tile partials of a *_SCORE layer)."""
    n = load().vad_conv_layer_tiles(C.byref(desc))
    if n <= 0:
        check(n if n < 0 else -1, "vad_conv_layer_tiles")
    return n


def m_tiles(B: int, H: int, W: int, single_frame: bool) -> int:
    n = load().vad_conv_m_tiles(B, H, W, 1 if single_frame else 0)
    if n <= 0:
        check(n, "vad_conv_m_tiles")
    return n
