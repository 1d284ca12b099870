"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no tests or fixtures for this path, so parity is pinned on outputs of the
reference classes themselves: seeded weights (random init as the reference initialises them, plus the "stress"
state_dicts of oracle/stress.py loaded INTO the reference modules) and seeded inputs.  Inputs are regenerated from
their seeds by the tests; only reference OUTPUTS (and weight digests) are stored.
"""
import hashlib
import io
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from models import ConvAutoencoder  # noqa: E402  (reference)
from models.video_autoencoder import VideoAutoencoder  # noqa: E402  (reference)

from oracle.stress import state_dict_digest, stress_state_dict  # noqa: E402


def image_input(seed, b, h, w):
    g = torch.Generator().manual_seed(seed)
    amp = 0.3 + 0.7 * torch.rand(b, 1, 1, 1, generator=g)
    return (amp * (2 * torch.rand(b, 3, h, w, generator=g) - 1)).clamp(-1, 1)


def video_input(seed, b, t, h, w):
    g = torch.Generator().manual_seed(seed)
    amp = 0.3 + 0.7 * torch.rand(b, t, 1, 1, 1, generator=g)
    return (amp * (2 * torch.rand(b, t, 3, h, w, generator=g) - 1)).clamp(-1, 1)


def main():
    torch.set_num_threads(4)
    out = {}
    # ------------------------------------------------------------------ image model
    for tag, latent in (("img", 256), ("img_l64", 64)):
        torch.manual_seed(0)
        m = ConvAutoencoder(3, latent).eval()
        out[f"{tag}.init_digest"] = np.frombuffer(state_dict_digest(m.state_dict()).encode(), dtype=np.uint8)
        for wtag in ("init", "stress"):
            if wtag == "stress":
                m.load_state_dict(stress_state_dict(m.state_dict(), seed=1))
                m.eval()
            for (b, h, w, seed) in ((2, 32, 32, 100), (3, 48, 80, 101)):
                x = image_input(seed, b, h, w)
                with torch.no_grad():
                    key = f"{tag}.{wtag}.{b}x{h}x{w}"
                    out[key + ".recon"] = m(x).numpy()
                    out[key + ".latent"] = m.get_latent(x).numpy()
                    out[key + ".map"] = m.get_reconstruction_error(x, per_pixel=True).numpy()
                    out[key + ".score"] = m.get_reconstruction_error(x, per_pixel=False).numpy()
    # ------------------------------------------------------------------ video model
    for tag, kw in (("vid", dict()), ("vid_h64", dict(latent_dim=128, lstm_hidden_dim=64, lstm_num_layers=1))):
        torch.manual_seed(0)
        m = VideoAutoencoder(**kw).eval()
        out[f"{tag}.init_digest"] = np.frombuffer(state_dict_digest(m.state_dict()).encode(), dtype=np.uint8)
        for wtag in ("init", "stress"):
            if wtag == "stress":
                m.load_state_dict(stress_state_dict(m.state_dict(), seed=1))
                m.eval()
            for (b, t, h, w, seed) in ((2, 3, 32, 32, 200), (1, 4, 48, 80, 201)):
                x = video_input(seed, b, t, h, w)
                with torch.no_grad():
                    key = f"{tag}.{wtag}.{b}x{t}x{h}x{w}"
                    out[key + ".recon"] = m(x).numpy()
                    out[key + ".map"] = m.get_reconstruction_error(x, per_pixel=True).numpy()
                    out[key + ".frame"] = m.get_reconstruction_error(x, per_frame=True).numpy()
                    out[key + ".seq"] = m.get_reconstruction_error(x).numpy()
                    both = m.get_reconstruction_error(x, per_frame=True, per_pixel=True).numpy()
                    assert both.shape == out[key + ".map"].shape  # per_pixel wins (video_autoencoder.py:373-380)
    # ------------------------------------------------------------------ cfg1: the repo's own synthetic dataset through evaluate.compute_auroc
    sys.modules.setdefault("matplotlib", MagicMock())
    sys.modules.setdefault("matplotlib.pyplot", MagicMock())
    import evaluate  # noqa: E402  (reference CLI module; compute_auroc used unmodified)
    from utils.dataset import MVTecDataset  # noqa: E402
    from utils.download_data import create_synthetic_test_data  # noqa: E402
    tmp = "/tmp/vad_golden_data"
    if not os.path.isdir(os.path.join(tmp, "synthetic")):
        create_synthetic_test_data(tmp, "synthetic")
    ds = MVTecDataset(tmp, "synthetic", "test", image_size=256)
    loader = torch.utils.data.DataLoader(ds, batch_size=32, shuffle=False)
    torch.manual_seed(0)
    m = ConvAutoencoder().eval()
    auroc, labels, scores, _types = evaluate.compute_auroc(m, loader, torch.device("cpu"))  # evaluate.py:91
    imgs = torch.stack([ds[i]["image"] for i in range(len(ds))])
    # store the dataset as uint8 (it was decoded from 8-bit PNGs: x = (u8/255 - 0.5)/0.5 exactly)
    u8 = torch.round((imgs * 0.5 + 0.5) * 255).to(torch.uint8)
    assert torch.equal(((u8.float() / 255) - 0.5) / 0.5, imgs), "dataset is not exactly 8-bit"
    out["cfg1.images_u8"] = u8.numpy()
    out["cfg1.scores"] = np.asarray(scores, dtype=np.float32)
    out["cfg1.labels"] = np.asarray(labels, dtype=np.int64)
    out["cfg1.auroc"] = np.asarray([auroc], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "golden_v1.npz"))
    print(f"wrote golden_v1.npz: {len(out)} arrays, {sz / 1e6:.2f} MB; cfg1 AUROC {auroc:.4f}")


if __name__ == "__main__":
    main()
