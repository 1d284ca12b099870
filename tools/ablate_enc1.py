"""Time the fused first encoder block (vad_enc1_fused) under the kernel's ablation switches (VAD_DBG bits; needs a GPU).

    VAD_DBG=<bits> python tools/ablate_enc1.py [H W B]
bits: 16 one tap of the pair-folded conv only | 32 no epilogue-A math / patch writes | 64 no output stores |
      256 no im2col conversion
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
import torch  # noqa: E402
from models import _native as nat, _prepare as prep  # noqa: E402

H, W, B = (int(v) for v in (sys.argv[1:4] + ["256", "256", "256"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w0 = torch.randn(32, 3, 3, 3, generator=g) * 0.2
w3 = torch.randn(32, 32, 3, 3, generator=g) * 0.06
fw = prep.to_device({"w": prep.pack_first_conv(w0.double(), torch.zeros(32).double())}, dev)["w"]
gw = prep.to_device({"w": prep.pack_conv3x3(w3.double(), torch.zeros(32).double())}, dev)["w"]
x = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dev)
out = torch.empty(B, H // 2, W // 2, 32, dtype=torch.bfloat16, device=dev)
lib = nat.load()


def run():
    nat.check(lib.vad_enc1_fused(x.data_ptr(), fw.w_tc.data_ptr(), fw.bias.data_ptr(), gw.w_pair.data_ptr(),
                                 gw.bias_pair.data_ptr(), 0.2, B, H, W, out.data_ptr(), nat.stream_ptr()),
              "vad_enc1_fused")


for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
try:
    torch.cuda.synchronize()
except Exception as exc:  # report which bounded wait fired, if any
    import ctypes as C
    trap = (C.c_ulonglong * 4)()
    lib.vad_debug_last_trap(trap)
    print("FAILED:", str(exc).splitlines()[0], "| last trap {tag, block, thread, parity} =", list(trap))
    os._exit(1)
ms = e0.elapsed_time(e1) / 20
tiles = B * ((H + 15) // 16) * ((W + 15) // 16)
print(f"VAD_DBG={os.environ.get('VAD_DBG', '0')} fused enc1 {H}x{W} B={B}: {ms:.4f} ms "
      f"(~{ms * 1e-3 * 1.965e9 / (tiles / 148):.0f} cycles per tile per SM at 1.965 GHz; MMA floor 1392)")
