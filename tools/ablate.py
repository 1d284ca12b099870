"""Time one conv layer under the kernel's ablation switches (VAD_DBG bits; bring-up tool, needs a GPU).

    VAD_DBG=<bits> python tools/ablate.py [cin cout H W B pool]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
import torch  # noqa: E402
from models import _layers as eng, _prepare as prep  # noqa: E402

cin, cout, H, W, B, pool = (int(v) for v in (sys.argv[1:7] + ["32", "32", "256", "256", "256", "1"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
pk = prep.to_device({"l": prep.pack_conv3x3(w.double(), torch.zeros(cout).double())}, dev)["l"]
x = torch.randn(B, H, W, cin, generator=g).to(torch.bfloat16).to(dev)
out = torch.empty(B, H // 2 if pool else H, W // 2 if pool else W, cout, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    eng._conv(pk, x, B, H, W, out, 0.2, pool=bool(pool))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng._conv(pk, x, B, H, W, out, 0.2, pool=bool(pool))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
tiles = B * ((H + 15) // 16) * ((W + 7) // 8)
print(f"VAD_DBG={os.environ.get('VAD_DBG', '0')} conv {cin}->{cout} {H}x{W} B={B} pool={pool}: {ms:.4f} ms "
      f"(~{ms * 1e-3 * 1.965e9 / (tiles / 148):.0f} cycles per 128-pixel tile per SM)")
