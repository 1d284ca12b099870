"""Time the fused image decoder tail (vad_convt_conv_score) under its ablation switches (bring-up tool, needs a GPU).

    VAD_DBG=<bits> python tools/ablate_tail.py [B H W]     (H, W: input of the transposed conv; output is 2H x 2W)
bits: 16 one MMA per sub-tile in stage 2 | 32 no patch writes | 64 no heat stores / reduction | 128 no x loads |
      256 no ep-2 math
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
import torch  # noqa: E402
from models import _layers as eng, _prepare as prep  # noqa: E402

B, H, W = (int(v) for v in (sys.argv[1:4] + ["256", "128", "128"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
p1 = prep.pack_convt2x2(torch.randn(32, 32, 2, 2, generator=g).double() * 0.2, torch.zeros(32).double())
p2 = prep.pack_conv3x3(torch.randn(3, 32, 3, 3, generator=g).double() * 0.05, torch.zeros(3).double(), pad_n_to=16)
pk = prep.to_device({"a": p1, "b": p2}, dev)
a = torch.randn(B, H, W, 32, generator=g).to(torch.bfloat16).to(dev)
x = (torch.rand(B, 3, 2 * H, 2 * W, generator=g) * 2 - 1).to(dev)
bufs = eng._Buffers()
run = lambda: eng._fused_image_tail(pk["a"], pk["b"], a, B, H, W, x, False, True, bufs)
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
tiles = B * ((2 * H + 14) // 14) * ((2 * W + 30) // 30)
print(f"VAD_DBG={os.environ.get('VAD_DBG', '0')} fused tail B={B} {H}x{W}: {ms:.4f} ms incl. finalize "
      f"(~{ms * 1e-3 * 1.965e9 / (tiles / 148):.0f} cycles per tile per SM)")
