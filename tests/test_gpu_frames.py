"""GPU: frame I/O kernels either side of the path against the reference's host formulas."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_normalize_u8_matches_totensor_normalize(cuda_device):
    from runtime import frames
    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (2, 3, 48, 64, 3), generator=g, dtype=torch.uint8)  # [B, T, H, W, 3]
    got = frames.normalize_u8(u8.to(cuda_device))
    torch.cuda.synchronize()
    # torchvision ToTensor (HWC uint8 -> CHW float / 255) + Normalize(mean .5, std .5): utils/dataset.py:65-70
    ref = (u8.permute(0, 1, 4, 2, 3).float().div(255) - 0.5) / 0.5
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert torch.equal(got.cpu(), ref)  # bit-exact: same fp32 operations in the same order


def test_denormalize_u8_matches_reference(cuda_device):
    from runtime import frames
    g = torch.Generator().manual_seed(1)
    x = torch.rand(5, 3, 32, 40, generator=g) * 2.4 - 1.2  # outside [-1,1] too: the clamp matters
    got = frames.denormalize_u8(x.to(cuda_device))
    torch.cuda.synchronize()
    t = torch.clamp(x * 0.5 + 0.5, 0, 1)                    # evaluate_video.py:40-49
    ref = (t.permute(0, 2, 3, 1).numpy() * 255).astype(np.uint8)
    assert np.array_equal(got.cpu().numpy(), ref)


def test_render_heatmap_matches_create_heatmap(cuda_device):
    from runtime import frames
    g = torch.Generator().manual_seed(2)
    heat = torch.rand(4, 64, 48, generator=g) ** 3
    heat[3] = 0.25                                          # constant frame: (e-min)/(0+1e-8) = 0 everywhere
    hd = heat.to(cuda_device)
    minmax = torch.stack([hd.amin((1, 2)), hd.amax((1, 2))], 1)
    got = frames.render_heatmap(hd, minmax).cpu().numpy()
    lut = np.load(os.path.join(GOLDEN, "jet_lut_rgb.npy"))  # cv2.COLORMAP_JET as RGB (tests/golden/make_jet_lut.py)
    for f in range(4):
        e = heat[f].numpy()
        norm = (e - e.min()) / (e.max() - e.min() + 1e-8)   # evaluate_video.py:56-57
        idx = (norm * 255).astype(np.uint8)
        ref = lut[idx]
        assert np.array_equal(got[f], ref)                  # byte output: bit-exact, LUT index included


def test_scoring_from_u8_frames_end_to_end(cuda_device):
    """uint8 frames -> device normalisation -> scoring == scoring of the host-normalised fp32 tensor."""
    from models import ConvAutoencoder
    from runtime import frames
    torch.manual_seed(0)
    m = ConvAutoencoder().eval().to(cuda_device)
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (3, 64, 64, 3), generator=g, dtype=torch.uint8)
    x_host = (u8.permute(0, 3, 1, 2).float().div(255) - 0.5) / 0.5
    a = m.get_reconstruction_error(frames.normalize_u8(u8.to(cuda_device)))
    b = m.get_reconstruction_error(x_host.to(cuda_device))
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_ssim_loss_matches_reference_golden_and_oracle(cuda_device, name):
    """vad_ssim_loss vs the unmodified reference's SSIMLoss / CombinedLoss outputs (tests/golden/golden_ssim.npz) and the
    oracle's per-pixel map.  fp32; the separable window reorders the sums: 1e-4 relative on the loss."""
    from oracle import vad_oracle
    from runtime import metrics
    from test_oracle import _ssim_pair
    gold = np.load(os.path.join(GOLDEN, "golden_ssim.npz"))
    p, t = _ssim_pair(*gold[f"{name}_spec"])
    loss, smap = metrics.ssim_loss(p.to(cuda_device), t.to(cuda_device), want_map=True)
    comb = metrics.combined_loss(p.to(cuda_device), t.to(cuda_device))
    torch.cuda.synchronize()
    np.testing.assert_allclose(loss.cpu().numpy(), gold[f"{name}_ssim_loss_per_frame"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(float(loss.mean()), gold[f"{name}_ssim_loss"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(float(comb.mean()), gold[f"{name}_combined"], rtol=1e-4, atol=2e-6)
    ref_map = vad_oracle.ssim_map(p, t)
    assert float((smap.cpu() - ref_map).abs().max()) < 2e-4


@pytest.mark.parametrize("kind", ["image", "video"])
def test_score_frames_u8_equals_normalised_fp32_path(cuda_device, kind):
    """`model.score_frames(uint8 frames)` (one C call: device-side ToTensor + Normalize, the forward, uint8 heat maps) ==
    `model.score_all(host-normalised fp32 frames)` bit for bit; heat_u8 == create_heatmap's numpy normalisation
    (evaluate_video.py:56-57) of the fp32 map, byte for byte."""
    from oracle.stress import stress_state_dict
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(5)
    if kind == "image":
        from models import ConvAutoencoder
        m = ConvAutoencoder()
        u8 = torch.randint(0, 256, (3, 64, 96, 3), generator=g, dtype=torch.uint8)
        x = (u8.float().div(255).permute(0, 3, 1, 2) - 0.5) / 0.5               # utils/dataset.py:65-70
    else:
        from models.video_autoencoder import VideoAutoencoder
        m = VideoAutoencoder()
        u8 = torch.randint(0, 256, (2, 3, 48, 64, 3), generator=g, dtype=torch.uint8)
        x = (u8.float().div(255).permute(0, 1, 4, 2, 3) - 0.5) / 0.5
    m.load_state_dict(stress_state_dict(m.state_dict(), seed=1))
    m = m.eval().to(cuda_device)
    a = m.score_frames(u8.to(cuda_device), want_recon=True, want_heat=True, want_heat_u8=True)
    b = m.score_all(x.contiguous().to(cuda_device), want_recon=True, want_heat=True)
    assert torch.equal(a.score, b.score) and torch.equal(a.minmax, b.minmax)
    assert torch.equal(a.heat, b.heat) and torch.equal(a.recon, b.recon)
    e = b.heat.cpu().numpy()
    ref_u8 = np.stack([(((f - f.min()) / (f.max() - f.min() + 1e-8)) * 255).astype(np.uint8) for f in e])
    assert np.array_equal(a.heat_u8.cpu().numpy(), ref_u8)
    only8 = m.score_frames(u8.to(cuda_device))                               # defaults: scores + uint8 heat maps only
    assert only8.heat is None and only8.recon is None
    assert torch.equal(only8.score, b.score) and torch.equal(only8.heat_u8, a.heat_u8)
    with pytest.raises(RuntimeError, match="uint8"):
        m.score_frames(x.to(cuda_device))


def test_compose_panels_matches_reference_hstack(cuda_device):
    """runtime.frames.compose_panels == np.hstack([denormalize(frame), denormalize(recon), create_heatmap(err)]) of the
    reference (evaluate_video.py:40-66, 355-364), byte for byte, on the outputs of one score_all call (256x256 frames: the
    size create_heatmap resizes to, so its cv2.resize is the identity)."""
    from models.video_autoencoder import VideoAutoencoder
    from oracle.stress import stress_state_dict
    from runtime import frames
    torch.manual_seed(0)
    m = VideoAutoencoder()
    m.load_state_dict(stress_state_dict(m.state_dict(), seed=1))
    m = m.eval().to(cuda_device)
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(1, 3, 3, 256, 256, generator=g) * 2.4 - 1.2).to(cuda_device)       # (some values beyond [-1, 1]: clamp)
    out = m.score_all(x, want_recon=True, want_heat=True)
    panels = frames.compose_panels(x.view(3, 3, 256, 256), out.recon, out.heat, out.minmax).cpu().numpy()
    lut = np.load(os.path.join(GOLDEN, "jet_lut_rgb.npy"))

    def denorm(t):                                                  # evaluate_video.py:40-49
        t = torch.clamp(t * 0.5 + 0.5, 0, 1)
        return (t.permute(1, 2, 0).cpu().numpy() * 255).astype(np.uint8)
    for f in range(3):
        e = out.heat[f].cpu().numpy()
        norm = (e - e.min()) / (e.max() - e.min() + 1e-8)           # evaluate_video.py:56-57
        heatmap = lut[(norm * 255).astype(np.uint8)]                # applyColorMap(JET) + BGR2RGB; resize(256,256) = identity
        ref = np.hstack([denorm(x[0, f]), denorm(out.recon[f]), heatmap])
        assert panels[f].shape == ref.shape == (256, 768, 3)
        assert np.array_equal(panels[f], ref)
