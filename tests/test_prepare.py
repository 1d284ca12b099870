"""CPU: host-side weight preparation (BN folding + GEMM repacking) checked against the oracle.

A tiny torch emulator below executes the EXACT GEMM formulation the kernels implement (same K ordering, pixel-shuffle
column order, ConvLSTM gate-row permutation, zero-padded rows) on the packed operands; it must reproduce the oracle's
reference arithmetic up to the bf16 rounding of the weights.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from models import _prepare as prep
from oracle import vad_oracle
from oracle.stress import stress_state_dict


def _image_sd(stress=True):
    from models import ConvAutoencoder
    torch.manual_seed(0)
    sd = ConvAutoencoder().state_dict()
    return stress_state_dict(sd, seed=1) if stress else sd


def _video_sd(stress=True, **kw):
    from models.video_autoencoder import VideoAutoencoder
    torch.manual_seed(0)
    sd = VideoAutoencoder(**kw).state_dict()
    return stress_state_dict(sd, seed=1) if stress else sd


def im2col3x3(x):  # x NHWC fp32 -> [N,H,W,9*C] with K = (ky*3+kx)*C + c, zero padding
    n, h, w, c = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    return torch.cat([xp[:, ky:ky + h, kx:kx + w, :] for ky in range(3) for kx in range(3)], dim=-1)


def emu_conv(g, x, slope, pool=False):
    y = im2col3x3(x) @ g.w.float().t() + g.bias
    if pool:
        n, h, w, c = y.shape
        y = y.view(n, h // 2, 2, w // 2, 2, c).amax(dim=(2, 4))
    y = torch.maximum(y, y * slope)
    return y[..., :g.n_total]


def emu_convt(g, x, slope):
    n, h, w, c = x.shape
    y = x @ g.w.float().t() + g.bias                       # [n,h,w,4*cout(+pad)]
    y = torch.maximum(y, y * slope)[..., :4 * g.cout]
    y = y.view(n, h, w, 2, 2, g.cout).permute(0, 1, 3, 2, 4, 5)   # n, h, di, w, dj, co
    return y.reshape(n, 2 * h, 2 * w, g.cout)


def emu_first(fw, x_nchw, pool):
    x = x_nchw.permute(0, 2, 3, 1)
    y = im2col3x3(x) @ fw.w + fw.bias                       # fw.w is [27, cout]
    y_tc = im2col3x3(x) @ fw.w_tc.float()[:, :27].t() + fw.bias
    assert torch.allclose(y, y_tc, rtol=2e-2, atol=2e-2)    # the two packings describe the same layer
    assert float(fw.w_tc.float()[:, 27:].abs().max()) == 0.0
    if pool:
        n, h, w, c = y.shape
        y = y.view(n, h // 2, 2, w // 2, 2, c).amax(dim=(2, 4))
    return torch.maximum(y, y * 0.2)


def test_image_packing_reproduces_oracle():
    sd = _image_sd()
    p = prep.prepare_image(sd)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    a = emu_first(p["enc1.0"], x, pool=False)
    a = emu_conv(p["enc1.3"], a, 0.2, pool=True)
    for blk in ("enc2", "enc3", "enc4"):
        a = emu_conv(p[f"{blk}.0"], a, 0.2)
        a = emu_conv(p[f"{blk}.3"], a, 0.2, pool=True)
    with torch.no_grad():
        ref_lat = vad_oracle.image_encoder(sd, x)
    assert (a.permute(0, 3, 1, 2) - ref_lat).abs().max() <= 0.03 * ref_lat.abs().max()
    for blk in ("dec1", "dec2", "dec3"):
        a = emu_convt(p[f"{blk}.0"], a, 0.0)
        a = emu_conv(p[f"{blk}.3"], a, 0.0)
    a = emu_convt(p["dec4.0"], a, 0.0)
    last = p["dec4.3"]
    assert last.n_total == 16 and last.cout == 3 and float(last.w[3:].float().abs().max()) == 0.0
    rec = torch.tanh(im2col3x3(a) @ last.w.float().t() + last.bias)[..., :3].permute(0, 3, 1, 2)
    with torch.no_grad():
        ref = vad_oracle.image_forward(sd, x)
    assert (rec - ref).abs().mean() < 5e-3 and (rec - ref).abs().max() < 0.1


@pytest.mark.parametrize("kw", [{}, dict(latent_dim=128, lstm_hidden_dim=64, lstm_num_layers=1)])
def test_video_packing_reproduces_oracle(kw):
    sd = _video_sd(**kw)
    p = prep.prepare_video(sd)
    g = torch.Generator().manual_seed(5)
    b, t = 2, 3
    x = torch.rand(b, t, 3, 32, 32, generator=g) * 2 - 1
    a = emu_first(p["enc.0"], x.view(b * t, 3, 32, 32), pool=True)
    for i in (4, 8, 12):
        a = emu_conv(p[f"enc.{i}"], a, 0.2, pool=True)
    seq = a.view(b, t, *a.shape[1:])
    for layer in range(p["lstm_layers"]):
        gw = p[f"lstm.{layer}"]
        hid = gw.cout
        h = torch.zeros(b, *seq.shape[2:4], hid)
        c = torch.zeros_like(h)
        outs = []
        for ti in range(t):
            gates = im2col3x3(torch.cat([seq[:, ti], h], dim=-1)) @ gw.w.float().t() + gw.bias
            # kernel column layout: tiles of 128 = [gate][32 channels]
            gt = gates.view(*gates.shape[:-1], hid // 32, 4, 32)
            gi, gf, gg, go = (gt[..., k, :].reshape(*gates.shape[:-1], hid) for k in range(4))
            c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
            h = torch.sigmoid(go) * torch.tanh(c)
            outs.append(h)
        seq = torch.stack(outs, dim=1)
    a = seq.reshape(b * t, *seq.shape[2:])
    if "proj" in p:
        a = a @ p["proj"].w.float().t() + p["proj"].bias
    for i in (0, 3, 6):
        a = emu_convt(p[f"dec.{i}"], a, 0.0)
    last = p["dec.9"]
    assert last.n_total == 16 and last.cout == 3
    y = torch.tanh(a @ last.w.float().t() + last.bias)[..., :12]
    n, h, w, _ = y.shape
    rec = y.view(n, h, w, 2, 2, 3).permute(0, 5, 1, 3, 2, 4).reshape(n, 3, 2 * h, 2 * w).view(b, t, 3, 32, 32)
    with torch.no_grad():
        ref = vad_oracle.video_forward(sd, x)
    assert (rec - ref).abs().mean() < 5e-3 and (rec - ref).abs().max() < 0.1


def test_bn_fold_is_exact_in_fp64():
    sd = _image_sd()
    w, b = prep.fold_conv(sd, "encoder.enc2.0", "encoder.enc2.1")
    x = torch.randn(1, 32, 8, 8, dtype=torch.float64)
    sd64 = vad_oracle.to_dtype(sd, torch.float64)
    ref = vad_oracle._bn_eval(sd64, "encoder.enc2.1", vad_oracle._conv3x3(sd64, "encoder.enc2.0", x))
    got = F.conv2d(x, w, b, padding=1)
    assert torch.allclose(got, ref, rtol=1e-10, atol=1e-10)
    wt, bt = prep.fold_convt(sd, "decoder.dec2.0", "decoder.dec2.1")
    z = torch.randn(1, 128, 4, 4, dtype=torch.float64)
    ref = vad_oracle._bn_eval(sd64, "decoder.dec2.1", vad_oracle._convt2x2(sd64, "decoder.dec2.0", z))
    assert torch.allclose(F.conv_transpose2d(z, wt, bt, stride=2), ref, rtol=1e-10, atol=1e-10)


def test_lstm_row_permutation_is_a_permutation():
    for hid in (32, 64, 128):
        perm = prep.lstm_row_permutation(hid)
        assert sorted(perm.tolist()) == list(range(4 * hid))
        # first tile: gate-major, 32 channels each
        assert perm[:32].tolist() == list(range(0, 32)) and perm[32:64].tolist() == list(range(hid, hid + 32))
    with pytest.raises(ValueError):
        prep.lstm_row_permutation(48)


def test_kx_merged_weight_layout():
    """`w_kx` (the layout of the kernel that folds the horizontal taps into N, include/vad_b200.h `weight_kx`) holds the
    same weights as `w`: row kx*Cout + co, column ky*Cin + ci; zero rows up to a multiple of 16; only narrow layers."""
    g = torch.Generator().manual_seed(3)
    for cout, cin, pad in ((32, 32, 0), (3, 32, 16), (64, 64, 0)):
        w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
        pk = prep.pack_conv3x3(w, torch.zeros(cout, dtype=torch.float64), pad_n_to=pad)
        assert pk.w_kx is not None
        rows = (3 * cout + 15) // 16 * 16
        assert tuple(pk.w_kx.shape) == (rows, 3 * cin)
        wk = pk.w_kx.float().view(rows, 3, cin)
        for kx in range(3):
            for ky in range(3):
                assert torch.equal(wk[kx * cout:(kx + 1) * cout, ky], w[:, :, ky, kx].to(torch.bfloat16).float())
                # ... and it is the same number the tap-major layout holds at K = (ky*3+kx)*Cin + ci
                k0 = (ky * 3 + kx) * cin
                assert torch.equal(pk.w.float()[:cout, k0:k0 + cin], wk[kx * cout:(kx + 1) * cout, ky])
        assert float(wk[3 * cout:].abs().max() if rows > 3 * cout else 0.0) == 0.0
    wide = prep.pack_conv3x3(torch.randn(128, 128, 3, 3, generator=g, dtype=torch.float64),
                             torch.zeros(128, dtype=torch.float64))
    assert wide.w_kx is None


def test_kx_merged_gemm_reproduces_conv():
    """Emulates the kx kernel's arithmetic on the packed operand: D[q][kx][co] = sum_{ky,ci} in[q + ky rows][ci] *
    w_kx[kx*Cout+co][ky*Cin+ci], out[x] = D[x-1][0] + D[x][1] + D[x+1][2] — equals conv3x3 with zero padding."""
    g = torch.Generator().manual_seed(5)
    cout, cin, H, W = 3, 32, 6, 9
    w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(1, H, W, cin, generator=g, dtype=torch.float64)
    pk = prep.pack_conv3x3(w, torch.zeros(cout, dtype=torch.float64), pad_n_to=16)
    wq = pk.w_kx.double()[:9]                                   # [kx*3+co][ky*cin+ci]
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))                           # halo columns and rows
    rows3 = torch.cat([xp[:, ky:ky + H, :, :] for ky in range(3)], dim=-1)   # [1,H,W+2,3*cin] (q = padded column)
    D = rows3 @ wq.t()                                          # [1,H,W+2,9]
    out = D[:, :, 0:W, 0:3] + D[:, :, 1:W + 1, 3:6] + D[:, :, 2:W + 2, 6:9]
    ref = F.conv2d(x.permute(0, 3, 1, 2), w.to(torch.bfloat16).double(), padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(out, ref, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("cout", [32, 64])
def test_pixel_pair_folded_weights_are_the_same_convolution(cout):
    """pack_conv3x3_pair: the 3x3 conv on pairs of horizontally adjacent pixels (input [B,H,W/2,2*32], output
    [B,H,W/2,2*Cout]) equals the original conv, and a third of the pair matrix is structurally zero (the K steps the
    kernel skips: left-neighbour pair x first pixel, right-neighbour pair x second pixel)."""
    g = torch.Generator().manual_seed(3)
    w = torch.randn(cout, 32, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    x = torch.randn(2, 32, 6, 10, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, b, padding=1)
    saved = prep  # (bf16 rounding aside: rebuild the pair matrix in fp64 with the same index formula)
    wp, bp = saved.pack_conv3x3_pair(w.to(torch.bfloat16).double(), b)
    w4 = wp.double().reshape(2 * cout, 3, 3, 64).permute(0, 3, 1, 2)          # [n][p_in*32+ci][ky][kxp]
    xp = x.reshape(2, 32, 6, 5, 2).permute(0, 4, 1, 2, 3).reshape(2, 64, 6, 5)   # channels p_in*32 + ci, W/2 pairs
    yp = F.conv2d(xp, w4, bp.double(), padding=1)                                 # [B][p_out*cout+co][H][W/2]
    y = yp.reshape(2, 2, cout, 6, 5).permute(0, 2, 3, 4, 1).reshape(2, cout, 6, 10)
    ref_bf = F.conv2d(x, w.to(torch.bfloat16).double(), b.float().double(), padding=1)   # (the packed bias is fp32)
    assert torch.allclose(y, ref_bf, rtol=0, atol=1e-12)
    assert (ref - ref_bf).abs().max() < 0.2
    k = wp.reshape(2 * cout, 3, 3, 2, 32)                                          # [n][ky][kxp][p_in][ci]
    assert float(k[:, :, 0, 0].abs().max()) == 0.0 and float(k[:, :, 2, 1].abs().max()) == 0.0
    gw = prep.pack_conv3x3(w, b)
    assert gw.w_pair is not None and tuple(gw.w_pair.shape) == (2 * cout, 576) and tuple(gw.bias_pair.shape) == (2 * cout,)
    assert prep.pack_conv3x3(torch.randn(64, 64, 3, 3, dtype=torch.float64), torch.zeros(64, dtype=torch.float64)).w_pair is None
