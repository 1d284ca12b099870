"""Per-tile role timeline of CTA 0 for the first conv (bring-up tool; needs a GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
import torch  # noqa: E402
from models import _engine as eng, _native as nat, _prepare as prep  # noqa: E402

B, H, W, pool = 64, 256, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = torch.randn(32, 3, 3, 3, generator=g) * 0.2
fw = prep.to_device({"w": prep.pack_first_conv(w.double(), torch.zeros(32).double())}, dev)["w"]
x = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dev)
out = torch.empty(B, H // 2 if pool else H, W // 2 if pool else W, 32, dtype=torch.bfloat16, device=dev)
buf = torch.zeros(4, 64, 16, dtype=torch.int64, device=dev)
for _ in range(2):
    eng._first_conv(fw, x, B, H, W, bool(pool), out)
torch.cuda.synchronize()
nat.load().vad_debug_set_timeline(buf.data_ptr())
eng._first_conv(fw, x, B, H, W, bool(pool), out)
torch.cuda.synchronize()
nat.load().vad_debug_set_timeline(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
print("first conv timeline, CTA 0 (cycles): conv = converter group leader [start, patch ready, packed, A slot free, arrived];"
      " mma = [start, accE ok, full ok, done]; epi = group leader [start, accF ok, done]")
for n in range(8, 36):
    conv = [int(v) - t0 for v in t[0, n, :5]]
    mma = [int(v) - t0 for v in t[1, n, :4]]
    epi = [int(v) - t0 for v in t[2 + (n & 1), n // 2, :3]]
    print(f"tile {n:2d} | conv " + " ".join(f"{v:6d}" for v in conv) + " | mma " + " ".join(f"{v:6d}" for v in mma) +
          " | epi " + " ".join(f"{v:6d}" for v in epi))
done = [int(t[1, n, 3]) for n in range(8, 40)]
print("steady-state cycles per tile:", (done[-1] - done[0]) / (len(done) - 1))
