"""Per-tile role timeline of CTA 0 for the fused first encoder block (vad_enc1_fused; bring-up tool; needs a GPU and a
library built with the stamps compiled in:  python video-anomaly-detection_b200/build.py --timeline)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
import torch  # noqa: E402
from models import _native as nat, _prepare as prep  # noqa: E402

B, H, W = 64, 256, 256
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


buf = torch.zeros(5, 64, 16, dtype=torch.int64, device=dev)
for _ in range(2):
    run()
torch.cuda.synchronize()
lib.vad_debug_set_timeline(buf.data_ptr())
run()
torch.cuda.synchronize()
lib.vad_debug_set_timeline(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
print("fused enc1 timeline, CTA 0 (cycles relative to the first stamp)")
print("conv [top, x patch ready, A slot free, done] | A [top, acc stage free, operands ready, issued] | "
      "epA(set 0, quarter 0) [top, D1 ready, in regs, patch written] | B [top, patch ready, issued] | "
      "epB [top, D2 ready, stage released, done]")
for n in range(6, 30):
    row = f"tile {n:2d}"
    for role, evs in ((0, 4), (1, 4), (2, 4), (3, 3), (4, 4)):
        row += " | " + " ".join(f"{int(t[role, n, e]) - t0:6d}" for e in range(evs))
    print(row)
done = [int(t[4, n, 3]) for n in range(8, 40)]
print("steady-state cycles per tile:", (done[-1] - done[0]) / (len(done) - 1))
