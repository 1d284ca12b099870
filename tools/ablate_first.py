"""Time the first conv (3 -> 32, fp32 NCHW input) under the kernel's ablation switches (VAD_DBG bits; needs a GPU).

    VAD_DBG=<bits> python tools/ablate_first.py [H W B pool]      pool: 0 plain, 1 pooled, 2 pooled on the pool-folded kernel
bits (pool-folded kernel): 32 no epilogue work | 64 no output stores | 128 no input loads | 256 no im2col conversion
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
import torch  # noqa: E402
from models import _layers as eng, _prepare as prep  # noqa: E402

H, W, B, pool = (int(v) for v in (sys.argv[1:5] + ["256", "256", "256", "0"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = torch.randn(32, 3, 3, 3, generator=g) * 0.2
fw = prep.to_device({"w": prep.pack_first_conv(w.double(), torch.zeros(32).double(), pooled=(pool == 2))}, dev)["w"]
x = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dev)
out = torch.empty(B, H // 2 if pool else H, W // 2 if pool else W, 32, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    eng._first_conv(fw, x, B, H, W, bool(pool), out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng._first_conv(fw, x, B, H, W, bool(pool), out)
e1.record()
try:
    torch.cuda.synchronize()
except Exception as exc:  # report which bounded wait fired, if any
    import ctypes as C
    from models import _native as nat
    trap = (C.c_ulonglong * 4)()
    nat.load().vad_debug_last_trap(trap)
    print("FAILED:", str(exc).splitlines()[0], "| last trap {tag, block, thread, parity} =", list(trap))
    os._exit(1)
ms = e0.elapsed_time(e1) / 20
tiles = B * ((H + 7) // 8) * ((W + 15) // 16) if pool != 2 else B * ((H // 2 + 7) // 8) * ((W // 2 + 15) // 16)
print(f"VAD_DBG={os.environ.get('VAD_DBG', '0')} first conv {H}x{W} B={B} pool={pool}: {ms:.4f} ms "
      f"(~{ms * 1e-3 * 1.965e9 / (tiles / 148):.0f} cycles per tile per SM)")
