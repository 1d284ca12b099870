"""GPU parity of each fused layer kernel against a plain torch fp32 reference of the same op.

Operands are rounded to bf16 first (that is what the kernels consume); accumulation is fp32 on both sides, so the
tolerance only has to cover the bf16 rounding of the OUTPUT (<= 2^-8 relative) plus accumulation-order noise.
Everything goes through the C ABI (ctypes -> libvad_b200.so).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

OUT_RTOL = 1.0 / 128  # bf16 output rounding (2^-8) with margin
OUT_ATOL = 2e-2


def _mods():
    from models import _layers as eng
    from models import _native as nat
    from models import _prepare as prep
    return eng, nat, prep


def _rand_nhwc(B, H, W, C, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(B, H, W, C, generator=g) * scale).to(torch.bfloat16)
    return x.to(dev)


def _nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def _nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def _assert_close(got, ref, what, rtol=OUT_RTOL, atol=OUT_ATOL):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    bad = err > bound
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} elements off; max err {float(err.max()):.4g} "
                           f"at ref {float(ref.flatten()[err.argmax()]):.4g}; ref rms {float(ref.pow(2).mean().sqrt()):.4g}")


def _dev(packed, dev):
    """Move one packed layer to the device (weights, bias and the optional kx-merged weight layout)."""
    packed.w, packed.bias = packed.w.to(dev), packed.bias.to(dev)
    if getattr(packed, "w_kx", None) is not None:
        packed.w_kx = packed.w_kx.to(dev)
    if getattr(packed, "w_pair", None) is not None:
        packed.w_pair, packed.bias_pair = packed.w_pair.to(dev), packed.bias_pair.to(dev)
    return packed


CONV_CASES = [
    # (cin, cout, B, H, W)
    (32, 32, 2, 32, 32),
    (32, 64, 2, 16, 48),
    (64, 64, 3, 16, 16),
    (64, 128, 2, 24, 40),     # partial tiles in H and W
    (128, 128, 5, 8, 8),      # two frames per tile, odd frame count
    (128, 256, 2, 16, 16),
    (256, 256, 2, 32, 32),
    (128, 128, 1, 6, 10),     # H, W not powers of two, smaller than a tile
    (32, 32, 2, 48, 80),      # kx-merged tiles: 80 = 13 x 6 + 2 (partial last tile), three tile rows
    (32, 32, 1, 16, 256),     # full-width rows: 43 tiles of 6 columns, the last one clipped to 4
    (128, 128, 3, 40, 32),    # patch + streamed-weights kernel: tile pairs, partial tile row (40 = 2*16 + 8), odd frames
    (128, 256, 2, 32, 64),    # same kernel, two N tiles of 128
    (256, 128, 1, 16, 48),    # four channel chunks
]


@pytest.mark.parametrize("cin,cout,B,H,W", CONV_CASES)
@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("kx", [True, False])
def test_conv3x3(cuda_device, cin, cout, B, H, W, pool, kx):
    eng, nat, prep = _mods()
    dev = cuda_device
    g = torch.Generator().manual_seed(cin * 1000 + cout + H)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    packed = _dev(prep.pack_conv3x3(w.double(), b.double()), dev)
    if kx and packed.w_kx is None:
        pytest.skip("layer has no kx-merged variant")
    x = _rand_nhwc(B, H, W, cin, dev, seed=H * W)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    out = torch.full((B, Ho, Wo, cout), float("nan"), dtype=torch.bfloat16, device=dev)
    prev = nat.load().vad_debug_set_kx(3 if kx else 0)  # 3: every eligible layer takes the kx-merged kernel
    try:
        eng._conv(packed, x, B, H, W, out, 0.2, pool=pool, what="test conv")
    finally:
        nat.load().vad_debug_set_kx(-1)
    torch.cuda.synchronize()
    ref = F.conv2d(_nchw(x), w.to(torch.bfloat16).float().to(dev), b.to(dev), padding=1)
    ref = F.leaky_relu(ref, 0.2)
    if pool:
        ref = F.max_pool2d(ref, 2, 2)
    _assert_close(out, _nhwc(ref), f"conv3x3 {cin}->{cout} B{B} {H}x{W} pool={pool}")


@pytest.mark.parametrize("cout,B,H,W", [(32, 2, 32, 32), (64, 2, 16, 48), (32, 1, 48, 80), (64, 1, 360, 640),
                                        (32, 3, 256, 256), (32, 2, 16, 32)])
@pytest.mark.parametrize("pool", [False, True])
def test_pixel_pair_folded_conv_equals_ordinary_view(cuda_device, cout, B, H, W, pool):
    """The 32-input-channel 3x3 layers run on the pixel-PAIR view by default (N = 2*Cout instead of the A-stream-bound
    N = 32: include/vad_b200.h `pair_fold`).  Same products, different fp32 summation order: the bf16 outputs agree
    with the ordinary view up to rare 1-ulp flips, and both agree with torch."""
    eng, nat, prep = _mods()
    dev = cuda_device
    g = torch.Generator().manual_seed(cout + H)
    w = torch.randn(cout, 32, 3, 3, generator=g) * (2.0 / 288) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    packed = _dev(prep.pack_conv3x3(w.double(), b.double()), dev)
    assert packed.w_pair is not None
    x = _rand_nhwc(B, H, W, 32, dev, seed=H + W)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    outs = {}
    prev = nat.load().vad_debug_set_kx(0)
    try:
        for pair in (True, False):
            out = torch.full((B, Ho, Wo, cout), float("nan"), dtype=torch.bfloat16, device=dev)
            eng.PAIR_FOLD = pair
            n0 = nat.launch_count()
            eng._conv(packed, x, B, H, W, out, 0.2, pool=pool, what="test conv")
            assert nat.launch_count() - n0 == 1
            torch.cuda.synchronize()
            outs[pair] = out
    finally:
        eng.PAIR_FOLD = True
        nat.load().vad_debug_set_kx(-1)
    ref = F.leaky_relu(F.conv2d(_nchw(x), w.to(torch.bfloat16).float().to(dev), b.to(dev), padding=1), 0.2)
    if pool:
        ref = F.max_pool2d(ref, 2, 2)
    _assert_close(outs[True], _nhwc(ref), f"pair-folded conv3x3 32->{cout} B{B} {H}x{W} pool={pool}")
    _assert_close(outs[False], _nhwc(ref), f"ordinary conv3x3 32->{cout} B{B} {H}x{W} pool={pool}")
    diff = (outs[True].float() - outs[False].float()).abs()
    assert float((diff > 0).float().mean()) < 0.02 and float(diff.max()) <= 0.0625 * float(ref.abs().max())


@pytest.mark.parametrize("cin,cout,B,H,W", [(256, 128, 2, 16, 16), (128, 64, 2, 8, 24), (64, 32, 3, 16, 16),
                                           (32, 32, 2, 32, 32), (128, 128, 3, 4, 4),
                                           # heights that are not a multiple of the 16-row tile (720p: 45 and 90 rows):
                                           # the pixel shuffle goes through the two split output maps, rows clipped
                                           (128, 64, 2, 45, 80), (128, 128, 1, 90, 160), (64, 32, 3, 22, 40)])
def test_convt2x2(cuda_device, cin, cout, B, H, W):
    eng, nat, prep = _mods()
    dev = cuda_device
    g = torch.Generator().manual_seed(cin + cout)
    w = torch.randn(cin, cout, 2, 2, generator=g) * (2.0 / cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    packed = _dev(prep.pack_convt2x2(w.double(), b.double()), dev)
    x = _rand_nhwc(B, H, W, cin, dev, seed=7)
    out = torch.full((B, 2 * H, 2 * W, cout), float("nan"), dtype=torch.bfloat16, device=dev)
    eng._convt(packed, x, B, H, W, out, 0.0, what="test convT")
    torch.cuda.synchronize()
    ref = F.relu(F.conv_transpose2d(_nchw(x), w.to(torch.bfloat16).float().to(dev), b.to(dev), stride=2))
    _assert_close(out, _nhwc(ref), f"convT {cin}->{cout} B{B} {H}x{W}")


@pytest.mark.parametrize("pool", [False, True, "folded"])
@pytest.mark.parametrize("B,H,W", [(2, 32, 64), (3, 16, 16), (1, 48, 80), (1, 720, 1280), (5, 128, 128), (2, 16, 48)])
def test_first_conv(cuda_device, pool, B, H, W):
    """First layer straight from the fp32 NCHW input: plain (image enc1.0), pooled on the one-row-per-input-pixel kernel
    (pool=True) and pooled on the kernel that folds the 2x2 window into the GEMM N extent ("folded": the video encoder's
    default; partial tiles at 16x16 / 16x48 / 720p)."""
    eng, nat, prep = _mods()
    dev = cuda_device
    folded = pool == "folded"
    pool = bool(pool)
    g = torch.Generator().manual_seed(5)
    w = torch.randn(32, 3, 3, 3, generator=g) * (2.0 / 27) ** 0.5
    b = torch.randn(32, generator=g) * 0.1
    fw = prep.to_device({"w": prep.pack_first_conv(w.double(), b.double(), pooled=folded)}, dev)["w"]
    assert (fw.w_pf is not None) == folded
    x = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dev)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    out = torch.full((B, Ho, Wo, 32), float("nan"), dtype=torch.bfloat16, device=dev)
    eng._first_conv(fw, x, B, H, W, pool, out)
    torch.cuda.synchronize()
    if eng.FIRST_CONV_TC:  # tensor-core path rounds both operands to bf16
        ref = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float().to(dev), b.to(dev), padding=1)
    else:
        ref = F.conv2d(x, w.to(dev), b.to(dev), padding=1)
    ref = F.leaky_relu(ref, 0.2)
    if pool:
        ref = F.max_pool2d(ref, 2, 2)
    _assert_close(out, _nhwc(ref), f"first conv pool={pool}", atol=1e-3)


@pytest.mark.parametrize("B,T,H,W", [(2, 3, 8, 8), (3, 2, 6, 10), (1, 5, 24, 40),
                                     (1, 1, 8, 8),      # a single step (x half only), odd frame count in a 2-frame tile
                                     (5, 2, 8, 8),      # 8x8 frames, last tile half empty
                                     (1, 3, 45, 80),    # the 720p latent: 8x16 tiles, partial last tile row (45 = 32+13)
                                     (2, 2, 16, 8)])    # narrowest frame the 8x16 tiling accepts
@pytest.mark.parametrize("persistent", [1, 2, 0])
def test_convlstm_sequence(cuda_device, B, T, H, W, persistent):
    """ConvLSTM over a short sequence vs the torch formulation of video_autoencoder.py:64-85 (one layer), with the
    persistent sequence kernels (1 patch variant, 2 streaming variant; cell state in registers) and with one launch per
    step (0; cell state in memory)."""
    eng, nat, prep = _mods()
    dev = cuda_device
    nat.load().vad_debug_set_lstm_mode(persistent)
    cin = hid = 128
    g = torch.Generator().manual_seed(11)
    w = torch.randn(4 * hid, cin + hid, 3, 3, generator=g) * (1.0 / (9 * (cin + hid))) ** 0.5
    b = torch.randn(4 * hid, generator=g) * 0.1
    packed = {"lstm.0": prep.pack_lstm(w.double(), b.double(), hid), "lstm_layers": 1}
    packed["lstm.0"].w, packed["lstm.0"].bias = packed["lstm.0"].w.to(dev), packed["lstm.0"].bias.to(dev)
    ve = eng.VideoEngine(packed)
    seq = _rand_nhwc(B * T, H, W, cin, dev, seed=3).view(B, T, H, W, cin)
    out = ve.convlstm(seq, B, T, H, W)
    torch.cuda.synchronize()
    wq = w.to(torch.bfloat16).float().to(dev)
    h = torch.zeros(B, hid, H, W, device=dev)
    c = torch.zeros_like(h)
    xs = seq.float().permute(0, 1, 4, 2, 3)
    for t in range(T):
        # the kernel feeds h back as bf16 (it is the next step's MMA operand)
        gates = F.conv2d(torch.cat([xs[:, t], h.to(torch.bfloat16).float()], 1), wq, b.to(dev), padding=1)
        gi, gf, gg, go = torch.split(gates, hid, 1)
        c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
        h = torch.sigmoid(go) * torch.tanh(c)
        _assert_close(out[:, t], _nhwc(h), f"convlstm h at t={t}", atol=1e-2)
    cst = ve.bufs.get("c0", (B, H, W, hid), torch.float32, dev)
    _assert_close(cst, _nhwc(c), "convlstm final c", rtol=1e-2, atol=1e-2)
    nat.load().vad_debug_set_lstm_mode(-1)


def test_convlstm_paths_agree(cuda_device):
    """The streaming sequence kernel (mode 2) runs the same MMAs in the same order as one launch per step (mode 0): the
    hidden sequence and the final cell state are bit-identical.  The patch kernel (mode 1) accumulates chunk-major
    instead of tap-major, so it agrees to fp32 summation-order / bf16 output rounding."""
    eng, nat, prep = _mods()
    dev = cuda_device
    B, T, H, W, cin, hid = 4, 6, 8, 8, 128, 128
    g = torch.Generator().manual_seed(31)
    w = torch.randn(4 * hid, cin + hid, 3, 3, generator=g) * (1.0 / (9 * (cin + hid))) ** 0.5
    b = torch.randn(4 * hid, generator=g) * 0.1
    seq = _rand_nhwc(B * T, H, W, cin, dev, seed=5).view(B, T, H, W, cin)
    outs = {}
    for mode in (0, 1, 2):
        packed = {"lstm.0": _dev(prep.pack_lstm(w.double(), b.double(), hid), dev), "lstm_layers": 1}
        ve = eng.VideoEngine(packed)
        nat.load().vad_debug_set_lstm_mode(mode)
        try:
            out = ve.convlstm(seq, B, T, H, W).clone()
            cst = ve.bufs.get("c0", (B, H, W, hid), torch.float32, dev).clone()
        finally:
            nat.load().vad_debug_set_lstm_mode(-1)
        torch.cuda.synchronize()
        outs[mode] = (out, cst)
    assert torch.equal(outs[2][0], outs[0][0])
    assert torch.equal(outs[2][1], outs[0][1])
    _assert_close(outs[1][0], outs[0][0].float(), "patch kernel h vs per-step h", atol=1e-2)
    _assert_close(outs[1][1], outs[0][1], "patch kernel c vs per-step c", rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("B,T,H,W", [(4, 6, 8, 8), (1, 1, 8, 8), (5, 2, 8, 8), (1, 5, 24, 40), (1, 3, 45, 80), (64, 16, 8, 8)])
def test_convlstm_two_layer_wavefront(cuda_device, B, T, H, W):
    """vad_convlstm2_sequence (both layers in one persistent launch, layer 2 one step behind layer 1) against two
    vad_convlstm_sequence launches: hidden sequences of both layers and both final cell states are bit-identical (same
    kernel roles, same MMA order per step)."""
    eng, nat, prep = _mods()
    dev = cuda_device
    cin = hid = 128
    g = torch.Generator().manual_seed(17)
    packed = {"lstm_layers": 2}
    for layer in range(2):
        w = torch.randn(4 * hid, cin + hid, 3, 3, generator=g) * (2.0 / (9 * (cin + hid))) ** 0.5
        b = torch.randn(4 * hid, generator=g) * 0.1
        packed[f"lstm.{layer}"] = _dev(prep.pack_lstm(w.double(), b.double(), hid), dev)
    seq = _rand_nhwc(B * T, H, W, cin, dev, seed=7).view(B, T, H, W, cin)
    res = {}
    launches = {}
    for fused in (True, False):
        ve = eng.VideoEngine(packed)
        eng.FUSE_LSTM_LAYERS = fused
        n0 = nat.launch_count()
        try:
            out = ve.convlstm(seq, B, T, H, W).clone()
        finally:
            eng.FUSE_LSTM_LAYERS = True
        torch.cuda.synchronize()
        launches[fused] = nat.launch_count() - n0
        res[fused] = (out, ve.bufs.get("hseq0", (B, T, H, W, hid), torch.bfloat16, dev).clone(),
                      ve.bufs.get("c0", (B, H, W, hid), torch.float32, dev).clone(),
                      ve.bufs.get("c1", (B, H, W, hid), torch.float32, dev).clone())
    assert launches[True] == 1 and launches[False] == 2
    for a, b, what in zip(res[True], res[False], ("h2 sequence", "h1 sequence", "c1", "c2")):
        assert torch.equal(a, b), what
    assert res[True][0].float().abs().mean().item() > 1e-3


@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (3, 16, 48), (1, 48, 256)])
@pytest.mark.parametrize("kx", [True, False])
def test_last_conv_tanh_score(cuda_device, B, H, W, kx):
    eng, nat, prep = _mods()
    dev = cuda_device
    g = torch.Generator().manual_seed(13)
    w = torch.randn(3, 32, 3, 3, generator=g) * (1.0 / 288) ** 0.5
    b = torch.randn(3, generator=g) * 0.1
    packed = _dev(prep.pack_conv3x3(w.double(), b.double(), pad_n_to=16), dev)
    assert packed.w_kx is not None and tuple(packed.w_kx.shape) == (16, 96)
    a = _rand_nhwc(B, H, W, 32, dev, seed=17)
    x = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dev)
    nat.load().vad_debug_set_kx(1 if kx else 0)
    try:
        res = eng._score_layer(packed, a, B, H, W, nat.EPI_TANH_SCORE, x, True, True, H, W, eng._Buffers(),
                               "test score")
    finally:
        nat.load().vad_debug_set_kx(-1)
    score, minmax, recon, heat = res.score, res.minmax, res.recon, res.heat
    torch.cuda.synchronize()
    ref = torch.tanh(F.conv2d(_nchw(a), w.to(torch.bfloat16).float().to(dev), b.to(dev), padding=1))
    err = ((x - ref) ** 2).mean(1)
    _assert_close(recon, ref, "recon", rtol=1e-4, atol=1e-4)
    _assert_close(heat, err, "heat", rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(score, err.mean((1, 2)), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(minmax[:, 0], err.amin((1, 2)), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(minmax[:, 1], err.amax((1, 2)), rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (3, 8, 24), (2, 8, 8)])
def test_last_convt_tanh_score(cuda_device, B, H, W):
    eng, nat, prep = _mods()
    dev = cuda_device
    g = torch.Generator().manual_seed(19)
    w = torch.randn(32, 3, 2, 2, generator=g) * (1.0 / 32) ** 0.5
    b = torch.randn(3, generator=g) * 0.1
    packed = _dev(prep.pack_convt2x2(w.double(), b.double(), pad_n_to=16), dev)
    a = _rand_nhwc(B, H, W, 32, dev, seed=23)
    Ho, Wo = 2 * H, 2 * W
    x = (torch.rand(B, 3, Ho, Wo, generator=g) * 2 - 1).to(dev)
    res = eng._score_layer(packed, a, B, H, W, nat.EPI_CONVT_TANH_SCORE, x, True, True, Ho, Wo, eng._Buffers(),
                           "test score")
    score, minmax, recon, heat = res.score, res.minmax, res.recon, res.heat
    torch.cuda.synchronize()
    ref = torch.tanh(F.conv_transpose2d(_nchw(a), w.to(torch.bfloat16).float().to(dev), b.to(dev), stride=2))
    err = ((x - ref) ** 2).mean(1)
    _assert_close(recon, ref, "recon", rtol=1e-4, atol=1e-4)
    _assert_close(heat, err, "heat", rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(score, err.mean((1, 2)), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(minmax[:, 0], err.amin((1, 2)), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(minmax[:, 1], err.amax((1, 2)), rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (3, 12, 20), (2, 4, 4), (1, 45, 80), (5, 8, 8)])
def test_fused_convt_convt_score(cuda_device, B, H, W):
    """vad_convt2_score (video decoder.6 + decoder.9 + score in one kernel) against torch and against the two layers
    run one by one: the reconstruction and the heat map must be bit-identical (same bf16 intermediate, same MMAs)."""
    eng, nat, prep = _mods()
    dev = cuda_device
    g = torch.Generator().manual_seed(31)
    w1 = torch.randn(64, 32, 2, 2, generator=g) * (2.0 / 64) ** 0.5
    b1 = torch.randn(32, generator=g) * 0.2
    w2 = torch.randn(32, 3, 2, 2, generator=g) * (1.0 / 32) ** 0.5
    b2 = torch.randn(3, generator=g) * 0.1
    p1 = _dev(prep.pack_convt2x2(w1.double(), b1.double()), dev)
    p2 = _dev(prep.pack_convt2x2(w2.double(), b2.double(), pad_n_to=16), dev)
    a = _rand_nhwc(B, H, W, 64, dev, seed=37)
    Ho, Wo = 4 * H, 4 * W
    x = (torch.rand(B, 3, Ho, Wo, generator=g) * 2 - 1).to(dev)
    res = eng._fused_tail(p1, p2, a, B, H, W, x, True, True, Ho, Wo, eng._Buffers())
    torch.cuda.synchronize()
    # the two layers one by one
    mid = torch.empty(B, 2 * H, 2 * W, 32, dtype=torch.bfloat16, device=dev)
    eng._convt(p1, a, B, H, W, mid, eng.RELU, what="test convt")
    two = eng._score_layer(p2, mid, B, 2 * H, 2 * W, nat.EPI_CONVT_TANH_SCORE, x, True, True, Ho, Wo, eng._Buffers(),
                           "test score")
    torch.cuda.synchronize()
    # (the fused kernel adds the first ConvT's bias inside its GEMM — hi + lo bf16 against a column of ones — so a few
    # intermediate values per thousand round to the neighbouring bf16)
    torch.testing.assert_close(res.recon, two.recon, rtol=0, atol=1e-2)
    assert (res.recon - two.recon).abs().mean().item() < 2e-5
    torch.testing.assert_close(res.heat, two.heat, rtol=2e-2, atol=1e-4)
    torch.testing.assert_close(res.score, two.score, rtol=1e-4, atol=1e-8)
    # torch: bf16-rounded weights, fp32 math, bf16-rounded intermediate
    m = torch.relu(F.conv_transpose2d(_nchw(a), w1.to(torch.bfloat16).float().to(dev), b1.to(dev), stride=2))
    m = m.to(torch.bfloat16).float()
    ref = torch.tanh(F.conv_transpose2d(m, w2.to(torch.bfloat16).float().to(dev), b2.to(dev), stride=2))
    err = ((x - ref) ** 2).mean(1)
    _assert_close(res.recon, ref, "recon", rtol=1e-2, atol=2e-2)  # (bf16 rounding flips of the intermediate)
    assert (res.recon - ref).abs().mean().item() < 1e-3
    torch.testing.assert_close(res.score, err.mean((1, 2)), rtol=2e-3, atol=1e-6)
    # outputs only (no recon / heat): same scores
    res2 = eng._fused_tail(p1, p2, a, B, H, W, x, False, False, Ho, Wo, eng._Buffers())
    torch.cuda.synchronize()
    assert res2.recon is None and res2.heat is None
    assert torch.equal(res2.score, res.score) and torch.equal(res2.minmax, res.minmax)


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (3, 24, 40), (1, 8, 8), (2, 128, 128), (1, 7, 15), (2, 21, 45)])
def test_fused_convt_conv_score(cuda_device, B, H, W):
    """vad_convt_conv_score (image dec4.0 + dec4.3 + score in one kernel, transposed conv recomputed per tile with a
    halo) against the two layers run one by one (rare 1-ulp bf16 flips of the intermediate: the bias is added inside the
    GEMM) and against torch."""
    eng, nat, prep = _mods()
    dev = cuda_device
    g = torch.Generator().manual_seed(41)
    w1 = torch.randn(32, 32, 2, 2, generator=g) * (2.0 / 32) ** 0.5
    b1 = torch.randn(32, generator=g) * 0.2
    w2 = torch.randn(3, 32, 3, 3, generator=g) * (1.0 / 288) ** 0.5
    b2 = torch.randn(3, generator=g) * 0.1
    p1 = _dev(prep.pack_convt2x2(w1.double(), b1.double()), dev)
    p2 = _dev(prep.pack_conv3x3(w2.double(), b2.double(), pad_n_to=16), dev)
    a = _rand_nhwc(B, H, W, 32, dev, seed=43)
    Ho, Wo = 2 * H, 2 * W
    x = (torch.rand(B, 3, Ho, Wo, generator=g) * 2 - 1).to(dev)
    res = eng._fused_image_tail(p1, p2, a, B, H, W, x, True, True, eng._Buffers())
    torch.cuda.synchronize()
    m = torch.relu(F.conv_transpose2d(_nchw(a), w1.to(torch.bfloat16).float().to(dev), b1.to(dev), stride=2))
    m = m.to(torch.bfloat16).float()
    ref = torch.tanh(F.conv2d(m, w2.to(torch.bfloat16).float().to(dev), b2.to(dev), padding=1))
    err = ((x - ref) ** 2).mean(1)
    _assert_close(res.recon, ref, "recon", rtol=1e-2, atol=2e-2)  # (bf16 rounding flips of the intermediate)
    assert (res.recon - ref).abs().mean().item() < 1e-3
    torch.testing.assert_close(res.score, err.mean((1, 2)), rtol=2e-3, atol=1e-6)
    torch.testing.assert_close(res.heat.mean((1, 2)), res.score, rtol=1e-5, atol=0)
    assert torch.equal(res.minmax[:, 0], res.heat.amin((1, 2))) and torch.equal(res.minmax[:, 1], res.heat.amax((1, 2)))
    if H >= 16 and W >= 16 and H % 8 == 0 and W % 8 == 0:  # shapes the layer-by-layer kernels take
        mid = torch.empty(B, Ho, Wo, 32, dtype=torch.bfloat16, device=dev)
        eng._convt(p1, a, B, H, W, mid, eng.RELU, what="test convt")
        two = eng._score_layer(p2, mid, B, Ho, Wo, nat.EPI_TANH_SCORE, x, True, True, Ho, Wo, eng._Buffers(),
                               "test score")
        torch.cuda.synchronize()
        # same bf16 rounding points and MMA sequence for the 3x3 conv; the fused kernel adds the transposed conv's bias
        # inside its GEMM (hi + lo bf16 against a column of ones) instead of in fp32 afterwards, so a few intermediate
        # values per thousand round to the neighbouring bf16
        torch.testing.assert_close(res.recon, two.recon, rtol=0, atol=1e-2)
        assert (res.recon - two.recon).abs().mean().item() < 2e-5
        torch.testing.assert_close(res.heat, two.heat, rtol=2e-2, atol=1e-4)
        torch.testing.assert_close(res.score, two.score, rtol=1e-4, atol=1e-8)
    res2 = eng._fused_image_tail(p1, p2, a, B, H, W, x, False, False, eng._Buffers())
    torch.cuda.synchronize()
    assert res2.recon is None and res2.heat is None
    assert torch.equal(res2.score, res.score) and torch.equal(res2.minmax, res.minmax)


def test_fused_convt_convt_score_rejects_other_widths(cuda_device):
    eng, nat, prep = _mods()
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    p1 = _dev(prep.pack_convt2x2(torch.randn(128, 64, 2, 2, generator=g).double(), torch.zeros(64).double()), dev)
    p2 = _dev(prep.pack_convt2x2(torch.randn(64, 3, 2, 2, generator=g).double(), torch.zeros(3).double(), pad_n_to=16), dev)
    a = _rand_nhwc(1, 8, 8, 128, dev, seed=1)
    x = torch.zeros(1, 3, 32, 32, device=dev)
    with pytest.raises(RuntimeError):
        eng._fused_tail(p1, p2, a, 1, 8, 8, x, False, False, 32, 32, eng._Buffers())


@pytest.mark.parametrize("N,H,W", [(5, 64, 64), (2, 256, 256), (3, 16, 48)])
def test_standalone_score_and_heatmap_u8(cuda_device, N, H, W):
    eng, nat, prep = _mods()
    dev = cuda_device
    lib = nat.load()
    g = torch.Generator().manual_seed(29)
    x = (torch.rand(N, 3, H, W, generator=g) * 2 - 1).to(dev)
    r = (torch.rand(N, 3, H, W, generator=g) * 2 - 1).to(dev)
    score = torch.empty(N, device=dev)
    minmax = torch.empty(N, 2, device=dev)
    heat = torch.empty(N, H, W, device=dev)
    scratch = torch.empty(lib.vad_score_scratch_bytes(N, H, W), dtype=torch.uint8, device=dev)
    nat.check(lib.vad_score(x.data_ptr(), r.data_ptr(), N, H, W, score.data_ptr(), minmax.data_ptr(), heat.data_ptr(),
                            scratch.data_ptr(), nat.stream_ptr()), "vad_score")
    u8 = torch.empty(N, H, W, dtype=torch.uint8, device=dev)
    nat.check(lib.vad_heatmap_u8(heat.data_ptr(), minmax.data_ptr(), N, H, W, u8.data_ptr(), nat.stream_ptr()),
              "vad_heatmap_u8")
    torch.cuda.synchronize()
    err = ((x - r) ** 2).mean(1)
    torch.testing.assert_close(heat, err, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(score, err.mean((1, 2)), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(minmax[:, 0], heat.amin((1, 2)), rtol=0, atol=0)
    torch.testing.assert_close(minmax[:, 1], heat.amax((1, 2)), rtol=0, atol=0)
    # evaluate_video.py:56-57 on the kernel's own heat map
    e = heat.cpu().numpy()
    import numpy as np
    ref_u8 = np.stack([(((m - m.min()) / (m.max() - m.min() + 1e-8)) * 255).astype(np.uint8) for m in e])
    # byte output: bit-exact (the kernel performs numpy's four fp32 operations, each correctly rounded, then truncates)
    assert np.array_equal(ref_u8, u8.cpu().numpy())


def test_layout_round_trip(cuda_device):
    eng, nat, prep = _mods()
    dev = cuda_device
    lib = nat.load()
    x = torch.randn(3, 40, 6, 10, device=dev)
    nhwc = torch.empty(3, 6, 10, 40, dtype=torch.bfloat16, device=dev)
    back = torch.empty_like(x)
    nat.check(lib.vad_nchw_f32_to_nhwc_bf16(x.data_ptr(), 3, 40, 6, 10, nhwc.data_ptr(), nat.stream_ptr()), "to nhwc")
    nat.check(lib.vad_nhwc_bf16_to_nchw_f32(nhwc.data_ptr(), 3, 6, 10, 40, back.data_ptr(), nat.stream_ptr()), "to nchw")
    torch.cuda.synchronize()
    assert torch.equal(nhwc, x.permute(0, 2, 3, 1).to(torch.bfloat16))
    assert torch.equal(back, x.to(torch.bfloat16).float())


@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (1, 16, 16), (3, 48, 80), (2, 256, 256), (1, 16, 64)])
def test_fused_enc1_equals_two_layers(cuda_device, B, H, W):
    """vad_enc1_fused (image encoder block 1: conv3x3 3->32 + LeakyReLU, conv3x3 32->32 + LeakyReLU, max-pool, with the
    32-channel full-resolution tensor kept in shared memory) against the two single-layer kernels it replaces — bit for
    bit (same MMAs in the same order, same bf16 rounding point) — and against torch."""
    eng, nat, prep = _mods()
    dev = cuda_device
    g = torch.Generator().manual_seed(H + W)
    w1 = torch.randn(32, 3, 3, 3, generator=g) * (2.0 / 27) ** 0.5
    b1 = torch.randn(32, generator=g) * 0.1
    w2 = torch.randn(32, 32, 3, 3, generator=g) * (2.0 / 288) ** 0.5
    b2 = torch.randn(32, generator=g) * 0.1
    fw = prep.to_device({"w": prep.pack_first_conv(w1.double(), b1.double())}, dev)["w"]
    pk = _dev(prep.pack_conv3x3(w2.double(), b2.double()), dev)
    x = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dev)
    fused = torch.full((B, H // 2, W // 2, 32), float("nan"), dtype=torch.bfloat16, device=dev)
    n0 = nat.launch_count()
    nat.check(nat.load().vad_enc1_fused(x.data_ptr(), fw.w_tc.data_ptr(), fw.bias.data_ptr(), pk.w_pair.data_ptr(),
                                        pk.bias_pair.data_ptr(), 0.2, B, H, W, fused.data_ptr(), nat.stream_ptr()),
              "vad_enc1_fused")
    assert nat.launch_count() - n0 == 1
    torch.cuda.synchronize()
    mid = torch.empty(B, H, W, 32, dtype=torch.bfloat16, device=dev)
    two = torch.full_like(fused, float("nan"))
    eng._first_conv(fw, x, B, H, W, False, mid)
    prev = nat.load().vad_debug_set_kx(0)
    try:
        eng._conv(pk, mid, B, H, W, two, 0.2, pool=True, what="enc1.3")
    finally:
        nat.load().vad_debug_set_kx(-1)
    torch.cuda.synchronize()
    if W >= 32 and H >= 16:                      # (narrower frames: the two-layer path uses the ordinary view of enc1.3)
        assert torch.equal(fused, two)
    m = F.leaky_relu(F.conv2d(x.to(torch.bfloat16).float(), w1.to(torch.bfloat16).float().to(dev), b1.to(dev), padding=1), 0.2)
    m = m.to(torch.bfloat16).float()
    ref = F.max_pool2d(F.leaky_relu(F.conv2d(m, w2.to(torch.bfloat16).float().to(dev), b2.to(dev), padding=1), 0.2), 2, 2)
    _assert_close(fused, _nhwc(ref), f"fused enc1 B{B} {H}x{W}")
    _assert_close(two, _nhwc(ref), f"two-layer enc1 B{B} {H}x{W}")
