"""Per-tile role timeline of CTA 0 for one conv layer (bring-up tool; needs a GPU).

    python tools/timeline.py [cin cout H W B pool]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
import torch  # noqa: E402
from models import _layers as eng, _native as nat, _prepare as prep  # noqa: E402

cin, cout, H, W, B, pool = (int(v) for v in (sys.argv[1:7] + ["32", "32", "256", "256", "64", "1"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
b = torch.zeros(cout)
pk = prep.pack_conv3x3(w.double(), b.double()); pk = prep.to_device({"pk": pk}, dev)["pk"]
x = torch.randn(B, H, W, cin, generator=g).to(torch.bfloat16).to(dev)
out = torch.empty(B, H // 2 if pool else H, W // 2 if pool else W, cout, dtype=torch.bfloat16, device=dev)
buf = torch.zeros(4, 64, 16, dtype=torch.int64, device=dev)
for _ in range(2):
    eng._conv(pk, x, B, H, W, out, 0.2, pool=bool(pool))
torch.cuda.synchronize()
nat.load().vad_debug_set_timeline(buf.data_ptr())
eng._conv(pk, x, B, H, W, out, 0.2, pool=bool(pool))
torch.cuda.synchronize()
nat.load().vad_debug_set_timeline(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
# epilogue events: 0 start, 1 accumulator ready, 3 staging buffer free, 4 accumulator in registers, 5 staged,
# 6 proxy fence done, 7 group barrier passed, 2 tile done
names = {0: ["wait_empty>", ">got_slot"], 1: ["wait_accE>", ">wait_full>", ">issue>", ">done"]}
# 8 after tcgen05 fence, 9 epilogue_tile index math done, 10 staging buffer free (leader's bulk wait)
EPI_ORDER = [0, 1, 8, 9, 10, 3, 4, 5, 6, 7, 2]
print(f"conv {cin}->{cout} {H}x{W} B={B} pool={pool}; cycles relative to first stamp; CTA 0")
for n in range(0, 40):
    row = [f"tile {n:2d}"]
    for role in (0, 1):
        ev = [int(v) - t0 for v in t[role, n, :len(names[role])]]
        row.append(f"r{role}: " + " ".join(f"{v:6d}" for v in ev))
    for role in (2, 3):  # epilogue groups 0 and 1: their n-th OWN tile
        ev = [int(t[role, n, e]) - t0 for e in EPI_ORDER]
        row.append(f"g{role - 2}: " + " ".join(f"{v:6d}" for v in ev) + " d=" + " ".join(f"{b - a_:4d}" for a_, b in zip(ev, ev[1:])))
    print(" | ".join(row))
mma_done = [int(t[1, n, 3]) for n in range(8, 40)]
print("steady-state cycles per tile (MMA done to MMA done):", (mma_done[-1] - mma_done[0]) / (len(mma_done) - 1))
