"""GPU parity of the drop-in model classes (through the C ABI) against the reference's own outputs.

Goldens (tests/golden/golden_v1.npz) are outputs of the unmodified reference; larger cases are checked against the
CPU oracle (oracle/vad_oracle.py, itself pinned on the goldens by tests/test_oracle.py) and through size-independent
properties.  Tolerances (north_star): scores within 1e-3 relative at the reference's random init with bf16 operands /
fp32 accumulation; identical thresholded flags; identical ranking on pairs the oracle separates by > 1e-5 relative;
AUROC equal to 3 decimals.  Because recon ~ 0 at random init every check is repeated with the
"stress" weights, where recon / heat-map tensors are compared directly; bf16 storage of 16 layer outputs bounds
those at the stated absolute tolerances.
"""
import os

import numpy as np
import pytest
import torch

from oracle import vad_oracle
from oracle.stress import stress_state_dict

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
SCORE_RTOL_INIT = 1e-3     # north_star bar (random-init weights)
SCORE_RTOL_STRESS = 5e-3   # trained-like weights: bf16 rounding at every layer boundary (SURVEY §7.3 measured ~1e-3)
RECON_MEAN_ATOL = 1.5e-2   # stress weights, mean |recon - ref|
RECON_MAX_ATOL = 0.35      # stress weights, worst pixel (a ReLU/Tanh flank amplifies one bf16 ulp)


def image_input(seed, b, h, w):
    g = torch.Generator().manual_seed(seed)
    amp = 0.3 + 0.7 * torch.rand(b, 1, 1, 1, generator=g)
    return (amp * (2 * torch.rand(b, 3, h, w, generator=g) - 1)).clamp(-1, 1)


def video_input(seed, b, t, h, w):
    g = torch.Generator().manual_seed(seed)
    amp = 0.3 + 0.7 * torch.rand(b, t, 1, 1, 1, generator=g)
    return (amp * (2 * torch.rand(b, t, 3, h, w, generator=g) - 1)).clamp(-1, 1)


def make_image_model(dev, latent=256, stress=False):
    from models import ConvAutoencoder
    torch.manual_seed(0)
    m = ConvAutoencoder(3, latent)
    if stress:
        m.load_state_dict(stress_state_dict(m.state_dict(), seed=1))
    return m.eval().to(dev)


def make_video_model(dev, stress=False, **kw):
    from models.video_autoencoder import VideoAutoencoder
    torch.manual_seed(0)
    m = VideoAutoencoder(**kw)
    if stress:
        m.load_state_dict(stress_state_dict(m.state_dict(), seed=1))
    return m.eval().to(dev)


def rel_err(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)))


def norm_map(e):
    e = np.asarray(e, np.float32)
    mn = e.min(axis=(-2, -1), keepdims=True)
    mx = e.max(axis=(-2, -1), keepdims=True)
    return (e - mn) / (mx - mn + 1e-8)


@pytest.mark.parametrize("tag,latent", [("img", 256), ("img_l64", 64)])
@pytest.mark.parametrize("wtag", ["init", "stress"])
@pytest.mark.parametrize("shape,seed", [((2, 32, 32), 100), ((3, 48, 80), 101)])
def test_image_parity(cuda_device, tag, latent, wtag, shape, seed):
    m = make_image_model(cuda_device, latent, wtag == "stress")
    x = image_input(seed, *shape).to(cuda_device)
    key = f"{tag}.{wtag}.{shape[0]}x{shape[1]}x{shape[2]}"
    recon = m(x).cpu().numpy()
    emap = m.get_reconstruction_error(x, per_pixel=True).cpu().numpy()
    score = m.get_reconstruction_error(x).cpu().numpy()
    latent_out = m.get_latent(x).cpu().numpy()
    assert recon.shape == GOLD[key + ".recon"].shape and emap.shape == GOLD[key + ".map"].shape
    assert score.shape == GOLD[key + ".score"].shape and latent_out.shape == GOLD[key + ".latent"].shape
    d = np.abs(recon - GOLD[key + ".recon"])
    lat_ref = GOLD[key + ".latent"]
    lat_err = np.abs(latent_out - lat_ref).max() / max(np.abs(lat_ref).max(), 1e-30)
    print(f"\n{key}: score rel {rel_err(score, GOLD[key + '.score']):.3g}  recon mean|d| {d.mean():.3g} max|d| {d.max():.3g}"
          f"  latent rel-to-max {lat_err:.3g}  normmap max|d| {np.abs(norm_map(emap) - norm_map(GOLD[key + '.map'])).max():.3g}")
    assert rel_err(score, GOLD[key + ".score"]) <= (SCORE_RTOL_INIT if wtag == "init" else SCORE_RTOL_STRESS)
    assert d.mean() <= RECON_MEAN_ATOL and d.max() <= RECON_MAX_ATOL
    assert lat_err <= 0.05
    np.testing.assert_allclose(emap.mean(axis=(1, 2, 3)), score, rtol=1e-5)  # SURVEY §4 invariant 3
    assert np.abs(recon).max() <= 1.0


@pytest.mark.parametrize("tag,kw", [("vid", {}), ("vid_h64", dict(latent_dim=128, lstm_hidden_dim=64, lstm_num_layers=1))])
@pytest.mark.parametrize("wtag", ["init", "stress"])
@pytest.mark.parametrize("shape,seed", [((2, 3, 32, 32), 200), ((1, 4, 48, 80), 201)])
def test_video_parity(cuda_device, tag, kw, wtag, shape, seed):
    m = make_video_model(cuda_device, wtag == "stress", **kw)
    x = video_input(seed, *shape).to(cuda_device)
    key = f"{tag}.{wtag}." + "x".join(str(s) for s in shape)
    recon = m(x).cpu().numpy()
    emap = m.get_reconstruction_error(x, per_pixel=True).cpu().numpy()
    frame = m.get_reconstruction_error(x, per_frame=True).cpu().numpy()
    seq = m.get_reconstruction_error(x).cpu().numpy()
    both = m.get_reconstruction_error(x, per_frame=True, per_pixel=True)
    assert tuple(both.shape) == GOLD[key + ".map"].shape  # per_pixel wins (video_autoencoder.py:373-380)
    assert recon.shape == GOLD[key + ".recon"].shape and frame.shape == GOLD[key + ".frame"].shape
    assert seq.shape == GOLD[key + ".seq"].shape
    d = np.abs(recon - GOLD[key + ".recon"])
    print(f"\n{key}: frame rel {rel_err(frame, GOLD[key + '.frame']):.3g} seq rel {rel_err(seq, GOLD[key + '.seq']):.3g}"
          f"  recon mean|d| {d.mean():.3g} max|d| {d.max():.3g}"
          f"  normmap max|d| {np.abs(norm_map(emap) - norm_map(GOLD[key + '.map'])).max():.3g}")
    tol = SCORE_RTOL_INIT if wtag == "init" else SCORE_RTOL_STRESS
    assert rel_err(frame, GOLD[key + ".frame"]) <= tol and rel_err(seq, GOLD[key + ".seq"]) <= tol
    assert d.mean() <= RECON_MEAN_ATOL and d.max() <= 2 * RECON_MAX_ATOL
    np.testing.assert_allclose(emap.mean(axis=(2, 3, 4)), frame, rtol=1e-5)
    np.testing.assert_allclose(frame.mean(axis=1), seq, rtol=1e-6)


def test_cfg1_synthetic_dataset(cuda_device):
    """Config 1: the repo's 30-image synthetic test set, batch 32 -> scores, flags, ranking, AUROC vs the reference."""
    from sklearn.metrics import roc_auc_score
    m = make_image_model(cuda_device)
    x = (((torch.from_numpy(GOLD["cfg1.images_u8"]).float() / 255) - 0.5) / 0.5).to(cuda_device)
    scores = m.get_reconstruction_error(x, per_pixel=False).cpu().numpy()  # evaluate.py:63, one ragged batch of 30
    ref = GOLD["cfg1.scores"]
    labels = GOLD["cfg1.labels"]
    print(f"\ncfg1 score rel err {rel_err(scores, ref):.3g}")
    assert rel_err(scores, ref) <= SCORE_RTOL_INIT
    assert np.array_equal(vad_oracle.image_flags(scores), vad_oracle.image_flags(ref))          # main.py:282
    assert np.array_equal(vad_oracle.video_flags(scores), vad_oracle.video_flags(ref))          # main.py:375-376
    checked, bad = vad_oracle.tie_aware_rank_agreement(ref, scores, rel_gap=1e-5)
    assert checked > 300 and bad == 0
    assert round(roc_auc_score(labels, scores), 3) == round(float(GOLD["cfg1.auroc"][0]), 3) == 0.615
    # the two bit-identical test images must tie exactly
    u8 = GOLD["cfg1.images_u8"]
    dup = [(i, j) for i in range(len(u8)) for j in range(i + 1, len(u8)) if np.array_equal(u8[i], u8[j])]
    assert dup and all(scores[i] == scores[j] for i, j in dup)


@pytest.mark.parametrize("stress", [False, True])
def test_image_medium_vs_oracle(cuda_device, stress):
    """256x256 batch against the CPU oracle: scores, flags, ranking, heat map."""
    m = make_image_model(cuda_device, stress=stress)
    x = image_input(1234, 8, 256, 256)
    sd = vad_oracle.cpu_sd(m.state_dict())
    with torch.no_grad():
        ref_map = vad_oracle.image_reconstruction_error(sd, x, per_pixel=True).numpy()
    ref = ref_map.mean(axis=(1, 2, 3))
    out = m.score_all(x.to(cuda_device))
    scores = out.score.cpu().numpy()
    print(f"\nimage 8x256x256 stress={stress}: score rel {rel_err(scores, ref):.3g}")
    assert rel_err(scores, ref) <= (SCORE_RTOL_STRESS if stress else SCORE_RTOL_INIT)
    checked, bad = vad_oracle.tie_aware_rank_agreement(ref, scores, rel_gap=1e-2 if stress else 1e-5)
    assert bad == 0 and checked > 0
    heat = out.heat.cpu().numpy()
    mm = out.minmax.cpu().numpy()
    np.testing.assert_array_equal(mm[:, 0], heat.min(axis=(1, 2)))
    np.testing.assert_array_equal(mm[:, 1], heat.max(axis=(1, 2)))
    nd = np.abs(norm_map(heat) - norm_map(ref_map[:, 0]))
    print(f"normalised heat map: mean|d| {nd.mean():.3g} max|d| {nd.max():.3g}")
    assert nd.mean() <= (2e-2 if stress else 1e-4)


@pytest.mark.parametrize("stress", [False, True])
def test_video_medium_vs_oracle(cuda_device, stress):
    """cfg3-shaped clips (T=16, 128x128), a few of them, against the CPU oracle."""
    m = make_video_model(cuda_device, stress=stress)
    x = video_input(4321, 3, 16, 128, 128)
    sd = vad_oracle.cpu_sd(m.state_dict())
    with torch.no_grad():
        ref = vad_oracle.video_reconstruction_error(sd, x, per_frame=True).numpy()
    frame = m.get_reconstruction_error(x.to(cuda_device), per_frame=True).cpu().numpy()
    print(f"\nvideo 3x16x128x128 stress={stress}: frame-score rel {rel_err(frame, ref):.3g}")
    assert rel_err(frame, ref) <= (SCORE_RTOL_STRESS if stress else SCORE_RTOL_INIT)
    assert np.array_equal(vad_oracle.video_flags(frame.ravel()), vad_oracle.video_flags(ref.ravel()))
    checked, bad = vad_oracle.tie_aware_rank_agreement(ref, frame, rel_gap=1e-2 if stress else 1e-5)
    assert bad == 0 and checked > 0


def test_video_720p_window_vs_oracle(cuda_device):
    """BASELINE cfg4 shape (1280x720 frames; latent 45x80 is not a multiple of the 8-row tile: partial tiles, the
    direct-store ConvT fallback and the non-power-of-two ConvLSTM grid are all exercised).  4-frame window."""
    m = make_video_model(cuda_device, stress=True)
    x = video_input(777, 1, 4, 720, 1280)
    sd = vad_oracle.cpu_sd(m.state_dict())
    with torch.no_grad():
        ref_map = vad_oracle.video_reconstruction_error(sd, x, per_pixel=True).numpy()
    ref = ref_map.mean(axis=(2, 3, 4))
    out = m.score_all(x.to(cuda_device))
    frame = out.score.cpu().numpy().reshape(1, 4)
    print(f"\nvideo 1x4x720x1280 stress: frame-score rel {rel_err(frame, ref):.3g}")
    assert rel_err(frame, ref) <= SCORE_RTOL_STRESS
    heat = out.heat.cpu().numpy()
    nd = np.abs(norm_map(heat) - norm_map(ref_map[0, :, 0]))
    assert nd.mean() <= 2e-2
    mm = out.minmax.cpu().numpy()
    np.testing.assert_array_equal(mm[:, 0], heat.min(axis=(1, 2)))
    np.testing.assert_array_equal(mm[:, 1], heat.max(axis=(1, 2)))


@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (1, 2, 720, 1280), (3, 2, 48, 80)])
def test_video_fused_decoder_tail_equals_layerwise(cuda_device, shape):
    """The fused decoder.6 + decoder.9 + score kernel (default) and the layer-by-layer schedule (VAD_FUSE_DEC=0)
    produce the same reconstruction and heat map bit for bit; scores differ only by the summation order."""
    from models import _engine as eng
    m = make_video_model(cuda_device, stress=True)
    x = video_input(99, *shape).to(cuda_device)
    assert eng.FUSE_DEC_TAIL
    a = m.score_all(x, want_recon=True, want_heat=True)
    a_score, a_heat, a_recon, a_mm = a.score.clone(), a.heat.clone(), a.recon.clone(), a.minmax.clone()
    eng.FUSE_DEC_TAIL = False
    try:
        b = m.score_all(x, want_recon=True, want_heat=True)
    finally:
        eng.FUSE_DEC_TAIL = True
    assert torch.equal(a_recon, b.recon)
    assert torch.equal(a_heat, b.heat)
    assert torch.equal(a_mm, b.minmax)
    torch.testing.assert_close(a_score, b.score, rtol=1e-5, atol=0)


@pytest.mark.parametrize("shape", [(3, 64, 64), (2, 256, 256), (2, 48, 80)])
def test_image_fused_decoder_tail_equals_layerwise(cuda_device, shape):
    """The fused dec4.0 + dec4.3 + score kernel (default) and the layer-by-layer schedule (VAD_FUSE_DEC=0) produce the
    same reconstruction and heat map bit for bit; scores differ only by the summation order."""
    from models import _engine as eng
    m = make_image_model(cuda_device, stress=True)
    g = torch.Generator().manual_seed(77)
    x = (torch.rand(shape[0], 3, shape[1], shape[2], generator=g) * 2 - 1).to(cuda_device)
    assert eng.FUSE_DEC_TAIL
    a = m.score_all(x, want_recon=True, want_heat=True)
    a_score, a_heat, a_recon, a_mm = a.score.clone(), a.heat.clone(), a.recon.clone(), a.minmax.clone()
    eng.FUSE_DEC_TAIL = False
    try:
        b = m.score_all(x, want_recon=True, want_heat=True)
    finally:
        eng.FUSE_DEC_TAIL = True
    assert torch.equal(a_recon, b.recon)
    assert torch.equal(a_heat, b.heat)
    assert torch.equal(a_mm, b.minmax)
    torch.testing.assert_close(a_score, b.score, rtol=1e-5, atol=0)


def test_full_size_properties_cfg2(cuda_device):
    """BASELINE cfg2 (batch 256 of 256x256): determinism, batch-partition invariance, map/score consistency."""
    m = make_image_model(cuda_device, stress=True)
    g = torch.Generator(device=cuda_device).manual_seed(1234)
    x = (torch.rand(256, 3, 256, 256, generator=g, device=cuda_device) * 2 - 1) * \
        (0.3 + 0.7 * torch.rand(256, 1, 1, 1, generator=g, device=cuda_device))
    a = m.score_all(x, want_recon=False, want_heat=True)
    s1, h1 = a.score.clone(), a.heat.clone()
    s2 = m.get_reconstruction_error(x)
    assert torch.equal(s1, s2), "run-to-run results must be bitwise identical"
    halves = torch.cat([m.get_reconstruction_error(x[:128]), m.get_reconstruction_error(x[128:])])
    assert torch.equal(s1, halves), "sharding the batch must not change any score bit (multi-GPU contract)"
    torch.testing.assert_close(h1.mean(dim=(1, 2)), s1, rtol=1e-5, atol=0)
    perm = torch.randperm(256, device=cuda_device)
    assert torch.equal(m.get_reconstruction_error(x[perm]), s1[perm]), "frames are independent"


def test_full_size_properties_cfg3(cuda_device):
    """BASELINE cfg3 (64 clips x 16 frames x 128x128)."""
    m = make_video_model(cuda_device, stress=True)
    g = torch.Generator(device=cuda_device).manual_seed(1234)
    x = torch.rand(64, 16, 3, 128, 128, generator=g, device=cuda_device) * 2 - 1
    f1 = m.get_reconstruction_error(x, per_frame=True).clone()
    f2 = m.get_reconstruction_error(x, per_frame=True)
    assert torch.equal(f1, f2)
    halves = torch.cat([m.get_reconstruction_error(x[:32], per_frame=True),
                        m.get_reconstruction_error(x[32:], per_frame=True)])
    assert torch.equal(f1, halves)
    seq = m.get_reconstruction_error(x)
    torch.testing.assert_close(seq, f1.mean(dim=1), rtol=1e-6, atol=0)
    # zero initial state every forward: scoring clip 0 alone equals clip 0 inside the batch
    assert torch.equal(m.get_reconstruction_error(x[:1], per_frame=True), f1[:1])


def test_error_behaviour(cuda_device):
    m = make_image_model(cuda_device)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        m.get_reconstruction_error(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="multiples of 16"):
        m.get_reconstruction_error(torch.zeros(1, 3, 40, 32, device=cuda_device))
    m.train()
    with pytest.raises(RuntimeError, match="eval"):
        m(torch.zeros(1, 3, 32, 32, device=cuda_device))
    m.eval()
    # in-place weight update invalidates the prepared-weight cache
    x = image_input(7, 2, 32, 32).to(cuda_device)
    before = m.get_reconstruction_error(x).clone()
    with torch.no_grad():
        m.decoder.dec4[3].bias.add_(0.5)
    after = m.get_reconstruction_error(x)
    assert not torch.equal(before, after)


def test_submodule_calls(cuda_device):
    """Encoder / VideoEncoder / ConvLSTM are callable like the reference's sub-modules."""
    m = make_image_model(cuda_device, stress=True)
    x = image_input(100, 2, 32, 32).to(cuda_device)
    assert torch.equal(m.encoder(x), m.get_latent(x))
    v = make_video_model(cuda_device, stress=True)
    xv = video_input(200, 2, 3, 32, 32).to(cuda_device)
    enc = v.encoder(xv)
    assert tuple(enc.shape) == (2, 3, 128, 2, 2)
    sd = vad_oracle.cpu_sd(v.state_dict())
    with torch.no_grad():
        ref_enc = vad_oracle.video_encoder(sd, xv.cpu())
        ref_seq = vad_oracle.convlstm(sd, ref_enc)
    assert (enc.cpu() - ref_enc).abs().max() <= 0.05 * ref_enc.abs().max()
    seq, (h_last, c_last) = v.convlstm(enc)
    assert tuple(seq.shape) == tuple(ref_seq.shape) and torch.equal(h_last, seq[:, -1])
    assert (seq.cpu() - ref_seq).abs().max() <= 0.05


def test_streaming_scorer_equals_windowed_forward(cuda_device):
    """runtime/streaming.py: cached per-frame encoder features + per-window ConvLSTM/decoder/score give
    bit-identical outputs to scoring every overlapping window from scratch, with each frame encoded once."""
    from models.video_autoencoder import VideoAutoencoder
    from runtime.streaming import StreamingVideoScorer
    torch.manual_seed(0)
    m = VideoAutoencoder().eval().to(cuda_device)
    g = torch.Generator().manual_seed(21)
    n, T, stride = 22, 8, 3
    video = (torch.rand(n, 3, 64, 48, generator=g) * 2 - 1).to(cuda_device)
    sc = StreamingVideoScorer(m, seq_len=T, stride=stride, want_recon=True)
    got = []
    for chunk in (video[:5], video[5:6], video[6:19], video[19:]):   # ragged arrival
        got += sc.push(chunk)
    starts = list(range(0, n - T + 1, stride))
    assert [s for s, _ in got] == starts
    assert sc.frames_encoded == n                                    # vs len(starts) * T = 40 without the cache
    for s, out in got:
        ref = m.score_all(video[s:s + T].unsqueeze(0))
        assert torch.equal(out.score, ref.score)
        assert torch.equal(out.minmax, ref.minmax)
        assert torch.equal(out.heat, ref.heat)
        assert torch.equal(out.recon, ref.recon)


@pytest.mark.parametrize("stress", [False, True])
def test_unusual_channel_widths_vs_oracle(cuda_device, stress):
    """Widths that are multiples of 32 but not of 64 / 128 (N tiles of 32, 32-channel K chunks, the CK=32 ConvLSTM
    sequence kernel, the 1x1 projection) against the CPU oracle."""
    tol = SCORE_RTOL_STRESS if stress else SCORE_RTOL_INIT
    m = make_image_model(cuda_device, latent=96, stress=stress)
    x = image_input(77, 3, 64, 96)
    with torch.no_grad():
        ref = vad_oracle.image_reconstruction_error(vad_oracle.cpu_sd(m.state_dict()), x).numpy()
    got = m.get_reconstruction_error(x.to(cuda_device)).cpu().numpy()
    print(f"\nimage latent=96 stress={stress}: score rel {rel_err(got, ref):.3g}")
    assert rel_err(got, ref) <= tol
    for kw in (dict(latent_dim=96, lstm_hidden_dim=96, lstm_num_layers=2),
               dict(latent_dim=64, lstm_hidden_dim=160, lstm_num_layers=1)):
        mv = make_video_model(cuda_device, stress=stress, **kw)
        xv = video_input(78, 2, 5, 64, 48)
        with torch.no_grad():
            refv = vad_oracle.video_reconstruction_error(vad_oracle.cpu_sd(mv.state_dict()), xv, per_frame=True).numpy()
        gotv = mv.get_reconstruction_error(xv.to(cuda_device), per_frame=True).cpu().numpy()
        print(f"video {kw} stress={stress}: frame-score rel {rel_err(gotv, refv):.3g}")
        assert rel_err(gotv, refv) <= tol
