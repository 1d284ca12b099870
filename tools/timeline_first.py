"""Per-tile role timeline of CTA 0 for the first conv (bring-up tool; needs a GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
import torch  # noqa: E402
from models import _layers as eng, _native as nat, _prepare as prep  # noqa: E402

# needs a library built with the stamps compiled in:  python video-anomaly-detection_b200/build.py --timeline
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
print("first conv timeline, CTA 0 (cycles relative to the first stamp)")
print("conv  = converter warp of the tile [start, patch+slot ready, arrived]")
print("mma   = issuer of the tile [start, accE ok, full ok, issued]")
print("epi   = leader of the tile's epilogue group [start, acc ready, fence, idx, buf free, bar, acc in regs, staged, "
      "proxy fence, bar, done]  (groups 0/1 only)")
EPI = [0, 1, 8, 9, 10, 3, 4, 5, 6, 7, 2]
for n in range(8, 36):
    conv = [int(t[0, n, e]) - t0 for e in (0, 1, 4)]
    mma = [int(v) - t0 for v in t[1, n, :4]]
    row = f"tile {n:2d} | conv " + " ".join(f"{v:6d}" for v in conv) + " | mma " + " ".join(f"{v:6d}" for v in mma)
    g = n & 3
    if g < 2:
        epi = [int(t[2 + g, n // 4, e]) - t0 for e in EPI]
        row += " | epi g%d " % g + " ".join(f"{v:6d}" for v in epi)
    print(row)
done = [int(t[1, n, 3]) for n in range(8, 40)]
print("steady-state cycles per tile:", (done[-1] - done[0]) / (len(done) - 1))
