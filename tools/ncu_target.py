"""Small fixed workload for ncu captures: a few scoring passes of one model (image: B=64 256x256; video: 8x16x128x128)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "image"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
if kind == "image":
    from models import ConvAutoencoder
    m = ConvAutoencoder().eval().to(dev)
    x = torch.rand(int(os.environ.get("NCU_B", "64")), 3, 256, 256, device=dev) * 2 - 1
else:
    from models.video_autoencoder import VideoAutoencoder
    m = VideoAutoencoder().eval().to(dev)
    x = torch.rand(int(os.environ.get("NCU_B", "8")), 16, 3, 128, 128, device=dev) * 2 - 1
for _ in range(iters):
    out = m.score_all(x, want_recon=False, want_heat=True)
torch.cuda.synchronize()
print("ok", float(out.score.mean()))
