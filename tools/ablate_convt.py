"""Time one transposed-convolution layer (k2 s2, streaming kernel) under the ablation switches (needs a GPU).

    VAD_DBG=<bits> [VAD_CONVT_RESIDENT=0] python tools/ablate_convt.py [cin cout H W B]
bits: 32 no epilogue math / staging | 64 no TMA store
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
import torch  # noqa: E402
from models import _layers as eng, _prepare as prep  # noqa: E402

cin, cout, H, W, B = (int(v) for v in (sys.argv[1:6] + ["128", "64", "90", "160", "64"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = torch.randn(cin, cout, 2, 2, generator=g) * 0.05
pk = prep.to_device({"l": prep.pack_convt2x2(w.double(), torch.zeros(cout).double())}, dev)["l"]
x = torch.randn(B, H, W, cin, generator=g).to(torch.bfloat16).to(dev)
out = torch.empty(B, 2 * H, 2 * W, cout, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    eng._convt(pk, x, B, H, W, out, eng.RELU, what="ablate convt")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng._convt(pk, x, B, H, W, out, eng.RELU, what="ablate convt")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
byts = x.numel() * 2 + out.numel() * 2
print(f"VAD_DBG={os.environ.get('VAD_DBG', '0')} resident={os.environ.get('VAD_CONVT_RESIDENT', '1')} convT {cin}->{cout} "
      f"{H}x{W} B={B}: {ms:.4f} ms = {byts / ms / 1e9:.2f} TB/s algorithmic")
