"""CPU: the C-ABI library builds/loads here and exports exactly what include/vad_b200.h declares; argument errors are
reported as codes (no compute calls — there is no GPU in this container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "vad_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vad_[a-z0-9_]+)\s*\(", src)) - {"vad_stream_t"})


@pytest.fixture(scope="module")
def lib():
    from models import _native as nat
    if not os.path.exists(nat.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return nat.load()


def test_header_and_binding_agree(lib):
    from models import _native as nat
    declared = _header_functions()
    assert declared == sorted(nat.EXPORTS), (declared, sorted(nat.EXPORTS))
    for name in declared:
        assert hasattr(lib, name), f"libvad_b200.so does not export {name}"


def test_struct_layout_matches_c(tmp_path, lib):
    """sizeof/offsetof of every ABI struct as the C compiler sees it == the ctypes mirror."""
    import subprocess
    from models import _native as nat
    structs = [("vad_conv_desc", nat.ConvDesc), ("vad_gemm_weights", nat.GemmW), ("vad_first_weights", nat.FirstW),
               ("vad_image_model", nat.ImageModel), ("vad_video_model", nat.VideoModel)]
    body = ""
    for cname, cls in structs:
        body += f'printf("%zu\\n", sizeof({cname}));'
        body += "".join(f'printf("%zu\\n", offsetof({cname}, {f[0]}));' for f in cls._fields_)
    prog = tmp_path / "sz.c"
    prog.write_text('#include "vad_b200.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                    f'int main(){{{body}return 0;}}')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = []
    for _, cls in structs:
        want.append(ctypes.sizeof(cls))
        want += [getattr(cls, f[0]).offset for f in cls._fields_]
    assert vals == want


def test_model_entry_points_reject_bad_arguments(lib):
    """Model-level entries: argument errors are codes, the workspace query needs no GPU."""
    from models import _native as nat
    m = nat.ImageModel()
    assert lib.vad_image_workspace_bytes(ctypes.byref(m), nat.OP_FORWARD, 4, 64, 64) == 0      # no weights
    assert lib.vad_image_forward(ctypes.byref(m), None, 4, 64, 64, None, None, None, None, None, None, 0, None) == -1
    v = nat.VideoModel()
    assert lib.vad_video_workspace_bytes(ctypes.byref(v), nat.OP_FORWARD, 1, 4, 64, 64) == 0
    assert lib.vad_video_forward(ctypes.byref(v), None, 1, 4, 64, 64, None, None, None, None, None, 0, None) == -1
    assert lib.vad_profile_enable(0) in (0, 1)
    buf = ctypes.create_string_buffer(64)
    assert lib.vad_profile_dump(buf, 64) == 0


def test_workspace_query_matches_the_schedule(lib):
    """vad_*_workspace_bytes on real (CPU-resident) prepared weights: pointers are only stored, nothing is launched.
    cfg2: two ping-pong regions and the score partials — not the sum of all layer outputs (2.9 GB); with the fused first
    block (default) the largest activation is enc1's pooled output (0.27 GB) + the 0.54 GB enc2.0 output, without it
    enc1.0's 1.07 GB full-resolution output."""
    import torch
    from models import _engine as eng, _native as nat, _prepare as prep
    from models import ConvAutoencoder
    from models.video_autoencoder import VideoAutoencoder
    torch.manual_seed(0)
    ie = eng.ImageEngine(prep.prepare_image(ConvAutoencoder().state_dict()))
    full = ie._ws(nat.OP_FORWARD, 256, 256, 256)
    assert 0.8e9 < full < 0.9e9, full
    assert eng.FUSE_ENC1
    eng.FUSE_ENC1 = False
    try:
        ie.m.flags = eng._flags()   # (run() refreshes the flags on every call; the bare workspace query does not)
        assert 1.3e9 < ie._ws(nat.OP_FORWARD, 256, 256, 256) < 1.45e9
    finally:
        eng.FUSE_ENC1 = True
        ie.m.flags = eng._flags()
    assert ie._ws(nat.OP_ENCODE, 256, 256, 256) <= full
    assert ie._ws(nat.OP_FORWARD, 1, 40, 64) == 0                                               # H not a multiple of 16
    small = ie._ws(nat.OP_FORWARD, 2, 32, 32)
    assert 0 < small < 1 << 20
    ve = eng.VideoEngine(prep.prepare_video(VideoAutoencoder().state_dict()))
    w = ve._ws(nat.OP_FORWARD, 1, 64, 720, 1280)
    assert 1.4e9 < w < 1.8e9, w
    assert 0 < ve._ws(nat.OP_SCORE_LATENTS, 1, 16, 8, 8) < ve._ws(nat.OP_FORWARD, 1, 16, 128, 128)


def test_argument_errors_are_codes_not_crashes(lib):
    assert lib.vad_version() >= 100
    assert lib.vad_error_string(0) == b"ok"
    assert b"unsupported" in lib.vad_error_string(-2).lower() or b"shape" in lib.vad_error_string(-2).lower()
    assert lib.vad_conv_layer(None, None) == -1
    assert lib.vad_conv_layer_tiles(None) == -1
    prev = lib.vad_debug_set_kx(0)
    assert lib.vad_debug_set_kx(-1) == 0 and lib.vad_debug_set_kx(-1) == prev
    assert lib.vad_conv_m_tiles(0, 16, 16, 0) == -1
    assert lib.vad_conv_m_tiles(2, 256, 256, 1) == 2 * 512
    assert lib.vad_score_scratch_bytes(0, 16, 16) == 0
    assert lib.vad_score_scratch_bytes(4, 256, 256) == 4 * 8 * 16
    assert lib.vad_first_conv(None, None, None, 32, 0.2, 0, 1, 16, 16, None, None) == -1
    assert lib.vad_score(None, None, 1, 16, 16, None, None, None, None, None) == -1
    # the fused first encoder block: null pointers, then the shape rule (H, W multiples of 16), before any CUDA call
    assert lib.vad_enc1_fused(None, None, None, None, None, 0.2, 1, 16, 16, None, None) == -1
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf) // 16 * 16 + 16
    assert lib.vad_enc1_fused(p, p, p, p, p, 0.2, 1, 24, 16, p, None) == -2
    assert lib.vad_enc1_fused(p, p, p, p, p, 0.2, 1, 16, 40, p, None) == -2
    assert lib.vad_compose_panel(None, None, None, None, 1, 16, 16, None, None) == -1


def test_fused_tail_entry_points_validate_on_the_host(lib):
    """The fused decoder-tail / two-layer ConvLSTM entry points reject bad descriptions with codes before touching the
    GPU, and their tile counts (= rows of the caller's `partials` buffer) follow the documented tilings."""
    from models._native import ConvDesc
    assert lib.vad_convt_conv_score(None, None, None, None) == -1
    assert lib.vad_convt2_score(None, None, None, None) == -1
    assert lib.vad_convlstm2_sequence(None, None, 4, None) == -1
    assert lib.vad_convt_conv_score_tiles(None) == -1 and lib.vad_convt2_score_tiles(None) == -1
    d = ConvDesc()
    d.B, d.H, d.W, d.ntaps, d.c0, d.n_total, d.cout = 3, 128, 128, 1, 32, 128, 32
    # image tail: tile (th, tw) scores output rows [14 th - 1, 14 th + 13) x columns [30 tw - 1, 30 tw + 29) of 2H x 2W
    assert lib.vad_convt_conv_score_tiles(ctypes.byref(d)) == 3 * 19 * 9
    d.H, d.W = 8, 8
    assert lib.vad_convt_conv_score_tiles(ctypes.byref(d)) == 3 * 2 * 1
    d.c0 = 64  # the image tail is the 32 -> 32 -> 3 block only
    assert lib.vad_convt_conv_score_tiles(ctypes.byref(d)) == -3
    assert lib.vad_convt_conv_score(ctypes.byref(d), None, None, None) == -1  # (null pointers are reported first)
    # video tail: 64 -> 32 -> 3, one 128-pixel tile inside one frame
    assert lib.vad_convt2_score_tiles(ctypes.byref(d)) == 3 * lib.vad_conv_m_tiles(1, 8, 8, 1)
    d.H, d.W = 180, 320
    assert lib.vad_convt2_score_tiles(ctypes.byref(d)) == lib.vad_conv_m_tiles(3, 180, 320, 1)
    d.cout, d.n_total = 16, 64
    assert lib.vad_convt2_score_tiles(ctypes.byref(d)) == -3
    d.ntaps = 9
    assert lib.vad_convt2_score_tiles(ctypes.byref(d)) == -1


def test_no_cpu_fallback():
    import torch
    from models import ConvAutoencoder
    m = ConvAutoencoder().eval()
    with pytest.raises(RuntimeError, match="CUDA-only"):
        m.get_reconstruction_error(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        m(torch.zeros(1, 3, 32, 32))


def test_jet_table_in_library_matches_golden_and_cv2():
    """The colour table compiled into vad_api.cu == tests/golden/jet_lut_rgb.npy (== cv2.COLORMAP_JET when cv2 is here)."""
    import numpy as np
    src = open(os.path.join(ROOT, "video-anomaly-detection_b200", "csrc", "vad_api.cu")).read()
    body = src[src.index("c_jet_rgb[256] = {"):]
    body = body[:body.index("};")]
    vals = [int(v, 16) for v in re.findall(r"0x([0-9a-f]{6})u", body)]
    assert len(vals) == 256
    table = np.array([[v & 255, (v >> 8) & 255, (v >> 16) & 255] for v in vals], dtype=np.uint8)
    golden = np.load(os.path.join(ROOT, "tests", "golden", "jet_lut_rgb.npy"))
    assert np.array_equal(table, golden)
    cv2 = pytest.importorskip("cv2")
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET)
    assert np.array_equal(cv2.cvtColor(lut, cv2.COLOR_BGR2RGB).reshape(256, 3), golden)
