"""Runs the reference's OWN, unmodified evaluation code against whichever `models` package is first on PYTHONPATH.

TEST INFRASTRUCTURE (driven by tests/test_gpu_dropin.py as a subprocess; the two packages are both called `models`, so
each run needs its own interpreter):

  PYTHONPATH=baseline/_ref                                   -> the reference classes (ground truth, CPU)
  PYTHONPATH=video-anomaly-detection_b200:baseline/_ref      -> the drop-in classes; `evaluate`, `evaluate_video`, `utils`
                                                               still resolve to the reference's files

  python tests/dropin_runner.py make  <workdir>   # reference classes: checkpoints (torch.save as train.py:208-215 /
                                                  # train_video.py:241-251 do) + the synthetic image set + a synthetic
                                                  # IPAD-format video set on disk
  python tests/dropin_runner.py eval  <workdir> <out.json> [--device cuda|cpu]
      image: evaluate.load_model (evaluate.py:26-43) + evaluate.compute_auroc (:46-91) on MVTecDataset / DataLoader(16)
      video: evaluate_video.evaluate(args) (:69-248): checkpoint load, IPADDataset, the scoring loop (:138-154),
             AUROC, generate_visualizations (forward + per_pixel + per-sequence calls, create_heatmap)
The only patches are to the harness around the reference code: matplotlib (absent in this image) is stubbed, and
`roc_auc_score` inside evaluate_video is wrapped to record the (labels, scores) it is called with.
"""
import argparse
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.modules.setdefault("matplotlib", MagicMock())
sys.modules.setdefault("matplotlib.pyplot", MagicMock())

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.append(ROOT)  # oracle.stress (seeded trained-like weights) — appended, so it never shadows `models`


def _make(work: str) -> None:
    """Everything here uses the REFERENCE classes (asserted): the checkpoints are what the reference's training scripts
    would have written."""
    import models
    from models import ConvAutoencoder
    from models.video_autoencoder import VideoAutoencoder
    from oracle.stress import stress_state_dict
    from utils.download_data import create_synthetic_test_data
    from PIL import Image
    assert "baseline/_ref" in models.__file__.replace("\\", "/"), models.__file__
    os.makedirs(work, exist_ok=True)
    create_synthetic_test_data(os.path.join(work, "data"), "synthetic")  # utils/download_data.py:85-184 (cfg1's set)
    torch.manual_seed(0)
    m = ConvAutoencoder(in_channels=3, latent_dim=256)
    torch.save({"epoch": 7, "model_state_dict": m.state_dict(), "train_loss": 0.0123, "val_loss": 0.0456,
                "args": {"latent_dim": 256, "category": "synthetic", "image_size": 256}},
               os.path.join(work, "image_init.pth"))
    torch.save({"epoch": 9, "model_state_dict": stress_state_dict(m.state_dict(), seed=1), "train_loss": 0.0042,
                "args": {"latent_dim": 256, "category": "synthetic", "image_size": 256}},
               os.path.join(work, "image_stress.pth"))
    torch.manual_seed(0)
    v = VideoAutoencoder(in_channels=3, latent_dim=128, lstm_hidden_dim=128, lstm_num_layers=2)
    vargs = {"category": "S01", "sequence_length": 16, "image_size": 256, "latent_dim": 128, "lstm_hidden_dim": 128,
             "lstm_layers": 2}
    os.makedirs(os.path.join(work, "video_stress"), exist_ok=True)
    torch.save({"epoch": 5, "model_state_dict": stress_state_dict(v.state_dict(), seed=1), "train_loss": 0.0099,
                "args": vargs}, os.path.join(work, "video_stress", "best_model.pth"))
    # IPAD layout (utils/video_dataset.py:28-34): S01/{training,testing}/frames/NN/*.png + S01/test_label/NNN.npy
    rng = np.random.default_rng(7)
    base = os.path.join(work, "ipad", "S01")
    yy, xx = np.mgrid[0:64, 0:64].astype(np.float32)
    for split, n_videos in (("training", 1), ("testing", 4)):
        for vid in range(1, n_videos + 1):
            d = os.path.join(base, split, "frames", f"{vid:02d}")
            os.makedirs(d, exist_ok=True)
            labels = np.zeros(32, dtype=np.int64)
            amp = 0.35 + 0.15 * vid
            for t in range(32):
                cx, cy = 12 + 1.2 * t + 3 * vid, 32 + 10 * np.sin(0.3 * t + vid)
                blob = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 6.0 ** 2))
                img = np.stack([amp * (0.25 + 0.5 * xx / 64 + 0.4 * blob), amp * (0.2 + 0.5 * yy / 64),
                                amp * (0.3 + 0.5 * blob)], -1)
                if split == "testing" and vid % 2 == 1 and t >= 16:      # anomaly: a noisy patch in the second half
                    y0, x0 = rng.integers(4, 40, 2)
                    img[y0:y0 + 20, x0:x0 + 20] += rng.uniform(-0.5, 0.6, (20, 20, 3))
                    labels[t] = 1
                img = img + rng.normal(0, 0.01, img.shape)
                Image.fromarray((np.clip(img, 0, 1) * 255).astype(np.uint8)).save(os.path.join(d, f"{t:04d}.png"))
            if split == "testing":
                os.makedirs(os.path.join(base, "test_label"), exist_ok=True)
                np.save(os.path.join(base, "test_label", f"{vid:03d}.npy"), labels)


def _eval(work: str, out_path: str, device_name: str) -> None:
    import models
    import evaluate          # reference CLI module, unmodified
    import evaluate_video    # reference CLI module, unmodified
    from utils import MVTecDataset
    assert "baseline/_ref" in evaluate.__file__.replace("\\", "/"), evaluate.__file__
    device = torch.device(device_name)
    res = {"models_file": models.__file__, "evaluate_file": evaluate.__file__, "device": device_name}
    # ---- image: evaluate.py:231-253 (dataset, DataLoader(batch_size=16), load_model, compute_auroc)
    ds = MVTecDataset(root_dir=os.path.join(work, "data"), category="synthetic", split="test", image_size=256)
    loader = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=False, num_workers=0)
    for tag in ("image_init", "image_stress"):
        model, _args = evaluate.load_model(os.path.join(work, tag + ".pth"), device)
        auroc, labels, scores, defects = evaluate.compute_auroc(model, loader, device)
        res[tag] = {"auroc": float(auroc), "labels": [int(v) for v in labels], "scores": [float(v) for v in scores],
                    "defects": {k: float(v["mean_score"]) for k, v in defects.items()},
                    "model_class": type(model).__module__ + "." + type(model).__name__,
                    "model_file": sys.modules[type(model).__module__].__file__}
    # ---- video: evaluate_video.evaluate (:69-248) end to end
    captured = {}
    real_auc = evaluate_video.roc_auc_score

    def recording_auc(labels, scores):
        captured["labels"] = [int(v) for v in labels]
        captured["scores"] = [float(v) for v in scores]
        return real_auc(labels, scores)

    evaluate_video.roc_auc_score = recording_auc
    if device_name == "cpu":  # evaluate() picks cuda when available (:73-78); the ground-truth run must stay on the CPU
        torch.cuda.is_available = lambda: False
    args = argparse.Namespace(checkpoint=os.path.join(work, "video_stress", "best_model.pth"),
                              data_dir=os.path.join(work, "ipad"), category=None, batch_size=4)
    auroc = evaluate_video.evaluate(args)
    res["video_stress"] = {"auroc": float(auroc), **captured,
                           "results_txt": open(os.path.join(work, "video_stress", "evaluation", "results.txt")).read()}
    with open(out_path, "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["make", "eval"])
    ap.add_argument("work")
    ap.add_argument("out", nargs="?")
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args()
    if a.mode == "make":
        _make(a.work)
    else:
        _eval(a.work, a.out, a.device)
