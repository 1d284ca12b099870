"""CPU: pin the oracle (oracle/vad_oracle.py, oracle/np_oracle.py) against golden outputs of the reference itself.

tests/golden/golden_v1.npz was produced by tests/golden/make_golden.py running the UNMODIFIED reference classes.
The drop-in classes are used here only as parameter containers (same seeded init as the reference — pinned by the
state_dict digest stored in the goldens).
"""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle, vad_oracle
from oracle.stress import state_dict_digest, stress_state_dict

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def image_input(seed, b, h, w):
    g = torch.Generator().manual_seed(seed)
    amp = 0.3 + 0.7 * torch.rand(b, 1, 1, 1, generator=g)
    return (amp * (2 * torch.rand(b, 3, h, w, generator=g) - 1)).clamp(-1, 1)


def video_input(seed, b, t, h, w):
    g = torch.Generator().manual_seed(seed)
    amp = 0.3 + 0.7 * torch.rand(b, t, 1, 1, 1, generator=g)
    return (amp * (2 * torch.rand(b, t, 3, h, w, generator=g) - 1)).clamp(-1, 1)


def image_sd(latent=256, stress=False):
    from models import ConvAutoencoder
    torch.manual_seed(0)
    sd = ConvAutoencoder(3, latent).state_dict()
    return stress_state_dict(sd, seed=1) if stress else sd


def video_sd(stress=False, **kw):
    from models.video_autoencoder import VideoAutoencoder
    torch.manual_seed(0)
    sd = VideoAutoencoder(**kw).state_dict()
    return stress_state_dict(sd, seed=1) if stress else sd


def digest_of(key):
    return GOLD[key].tobytes().decode()


def test_seeded_init_matches_reference_bit_for_bit():
    """Same module construction order + same init walk => identical RNG stream => identical weights."""
    assert state_dict_digest(image_sd()) == digest_of("img.init_digest")
    assert state_dict_digest(image_sd(64)) == digest_of("img_l64.init_digest")
    assert state_dict_digest(video_sd()) == digest_of("vid.init_digest")
    assert state_dict_digest(video_sd(latent_dim=128, lstm_hidden_dim=64, lstm_num_layers=1)) == digest_of("vid_h64.init_digest")


def test_state_dict_keys_match_reference_layout():
    keys = set(image_sd().keys())
    assert "encoder.enc1.0.weight" in keys and "encoder.enc4.4.running_var" in keys
    assert "decoder.dec4.3.bias" in keys and "decoder.dec4.4.weight" not in keys
    assert len(keys) == 16 * 2 + 15 * 5
    vkeys = set(video_sd().keys())
    assert {"encoder.encoder.12.weight", "convlstm.cells.1.conv.bias", "decoder.decoder.9.weight"} <= vkeys
    assert not any(k.startswith("proj.") for k in vkeys)
    assert "proj.weight" in video_sd(latent_dim=128, lstm_hidden_dim=64, lstm_num_layers=1)


@pytest.mark.parametrize("tag,latent", [("img", 256), ("img_l64", 64)])
@pytest.mark.parametrize("wtag", ["init", "stress"])
@pytest.mark.parametrize("shape,seed", [((2, 32, 32), 100), ((3, 48, 80), 101)])
def test_image_oracle_matches_reference(tag, latent, wtag, shape, seed):
    sd = image_sd(latent, wtag == "stress")
    x = image_input(seed, *shape)
    key = f"{tag}.{wtag}.{shape[0]}x{shape[1]}x{shape[2]}"
    with torch.no_grad():
        recon = vad_oracle.image_forward(sd, x)
        np.testing.assert_allclose(recon.numpy(), GOLD[key + ".recon"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(vad_oracle.image_encoder(sd, x).numpy(), GOLD[key + ".latent"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(vad_oracle.image_reconstruction_error(sd, x, True).numpy(), GOLD[key + ".map"],
                                   rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(vad_oracle.image_reconstruction_error(sd, x).numpy(), GOLD[key + ".score"],
                                   rtol=1e-5, atol=1e-8)
    assert np.abs(GOLD[key + ".recon"]).max() <= 1.0  # tanh range


@pytest.mark.parametrize("tag,kw", [("vid", {}), ("vid_h64", dict(latent_dim=128, lstm_hidden_dim=64, lstm_num_layers=1))])
@pytest.mark.parametrize("wtag", ["init", "stress"])
@pytest.mark.parametrize("shape,seed", [((2, 3, 32, 32), 200), ((1, 4, 48, 80), 201)])
def test_video_oracle_matches_reference(tag, kw, wtag, shape, seed):
    sd = video_sd(wtag == "stress", **kw)
    x = video_input(seed, *shape)
    key = f"{tag}.{wtag}." + "x".join(str(s) for s in shape)
    with torch.no_grad():
        np.testing.assert_allclose(vad_oracle.video_forward(sd, x).numpy(), GOLD[key + ".recon"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(vad_oracle.video_reconstruction_error(sd, x, per_pixel=True).numpy(),
                                   GOLD[key + ".map"], rtol=1e-5, atol=1e-7)
        frame = vad_oracle.video_reconstruction_error(sd, x, per_frame=True).numpy()
        np.testing.assert_allclose(frame, GOLD[key + ".frame"], rtol=1e-5, atol=1e-8)
        seq = vad_oracle.video_reconstruction_error(sd, x).numpy()
        np.testing.assert_allclose(seq, GOLD[key + ".seq"], rtol=1e-5, atol=1e-8)
        # invariants of SURVEY §4: mean of the map == frame score; mean over T == sequence score
        np.testing.assert_allclose(GOLD[key + ".map"].mean(axis=(2, 3, 4)), GOLD[key + ".frame"], rtol=1e-5)
        np.testing.assert_allclose(GOLD[key + ".frame"].mean(axis=1), GOLD[key + ".seq"], rtol=1e-5)


def test_numpy_restatement_agrees_with_reference_small():
    """Index-formula (torch-free) restatement in float64 vs the reference's fp32 outputs."""
    sd = image_sd(256, True)
    x = image_input(100, 2, 32, 32)
    rec = np_oracle.image_forward(sd, x.numpy())
    np.testing.assert_allclose(rec, GOLD["img.stress.2x32x32.recon"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(np_oracle.frame_scores(x.numpy(), rec), GOLD["img.stress.2x32x32.score"], rtol=1e-5)
    vsd = video_sd(True)
    xv = video_input(200, 2, 3, 32, 32)
    recv = np_oracle.video_forward(vsd, xv.numpy())
    np.testing.assert_allclose(recv, GOLD["vid.stress.2x3x32x32.recon"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(np_oracle.frame_scores(xv.numpy(), recv), GOLD["vid.stress.2x3x32x32.frame"], rtol=1e-5)


def test_stress_weights_make_parity_meaningful():
    """At random init recon ~ 0; the stress weights must give O(1) reconstructions."""
    assert np.sqrt((GOLD["img.init.2x32x32.recon"] ** 2).mean()) < 1e-2
    assert np.sqrt((GOLD["img.stress.2x32x32.recon"] ** 2).mean()) > 0.2
    assert np.sqrt((GOLD["vid.stress.2x3x32x32.recon"] ** 2).mean()) > 0.2


def test_cfg1_synthetic_dataset_scores_and_auroc():
    """Config 1: the repo's synthetic 256x256 test set (30 images) through the oracle == evaluate.compute_auroc."""
    from sklearn.metrics import roc_auc_score
    u8 = torch.from_numpy(GOLD["cfg1.images_u8"])
    x = ((u8.float() / 255) - 0.5) / 0.5
    sd = image_sd()
    with torch.no_grad():
        scores = torch.cat([vad_oracle.image_reconstruction_error(sd, x[i:i + 10]) for i in range(0, len(x), 10)]).numpy()
    np.testing.assert_allclose(scores, GOLD["cfg1.scores"], rtol=2e-5)
    labels = GOLD["cfg1.labels"]
    assert labels.sum() == 20 and len(labels) == 30
    assert round(roc_auc_score(labels, scores), 3) == round(float(GOLD["cfg1.auroc"][0]), 3)
    assert np.array_equal(vad_oracle.image_flags(scores), vad_oracle.image_flags(GOLD["cfg1.scores"]))


def test_consumers_heatmap_flags_and_rank_helper():
    e = np.array([[0.0, 0.5], [1.0, 0.25]], dtype=np.float32)
    assert vad_oracle.heatmap_u8(e).tolist() == [[0, 127], [255, 63]]  # fp32: 1 + 1e-8 == 1; 127.5 truncates
    s = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 10.0])
    assert vad_oracle.video_flags(s).tolist() == [False] * 9 + [True]
    assert vad_oracle.image_flags(np.array([0.0039, 0.0041])).tolist() == [False, True]
    ref = np.array([1.0, 2.0, 3.0, 3.0000001])
    checked, bad = vad_oracle.tie_aware_rank_agreement(ref, np.array([1.0, 2.0, 3.1, 3.0]))
    assert checked == 5 and bad == 0  # the near-tie pair is not compared
    checked, bad = vad_oracle.tie_aware_rank_agreement(ref, np.array([2.0, 1.0, 3.0, 3.0]))
    assert bad == 1


# ---- SSIM / combined loss: oracle restatement vs the unmodified reference's outputs
def _ssim_pair(seed, b, h, w, noise):
    g = torch.Generator().manual_seed(int(seed))
    t = (torch.rand(int(b), 3, int(h), int(w), generator=g) * 2 - 1)
    t = torch.nn.functional.avg_pool2d(t, 5, 1, 2)
    p = (t + noise * torch.randn(int(b), 3, int(h), int(w), generator=g)).clamp(-1, 1)
    return p, t


def test_ssim_oracle_matches_reference_golden():
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ssim.npz"))
    for name in ("a", "b", "c"):
        p, t = _ssim_pair(*gold[f"{name}_spec"])
        np.testing.assert_allclose(float(vad_oracle.ssim_loss(p, t)), gold[f"{name}_ssim_loss"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(vad_oracle.ssim_loss(p, t, per_frame=True).numpy(), gold[f"{name}_ssim_loss_per_frame"],
                                   rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(float(vad_oracle.combined_loss(p, t)), gold[f"{name}_combined"], rtol=1e-5, atol=1e-6)
