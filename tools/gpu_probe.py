"""Bring-up probe for the GPU box: runs the kernels from simplest to most complex and prints error summaries.

Usage (on the GPU box):  python tools/gpu_probe.py [stage ...]
Each stage runs in a fresh subprocess so that a trapped kernel (sticky CUDA error) does not hide later results.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
sys.path.insert(0, ROOT)

STAGES = ["first", "score", "gemm64", "gemm32", "conv64", "conv32", "pool", "convt", "lstm", "tanh", "convt_tanh",
          "image", "video"]


def summarize(name, got, ref):
    import torch
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    rms = ref.pow(2).mean().sqrt().item()
    nan = int(torch.isnan(got).sum())
    print(f"[{name}] max_err={err.max().item():.5g} mean_err={err.mean().item():.5g} ref_rms={rms:.4g} "
          f"nan={nan}/{got.numel()} bad(>0.05)={(err > 0.05).float().mean().item():.4f}", flush=True)


def run_stage(stage):
    import torch
    import torch.nn.functional as F
    from models import _layers as eng, _native as nat, _prepare as prep
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)

    def rnd_nhwc(B, H, W, C):
        return (torch.randn(B, H, W, C, generator=g)).to(torch.bfloat16).to(dev)

    nchw = lambda t: t.float().permute(0, 3, 1, 2).contiguous()
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()

    if stage == "first":
        w = torch.randn(32, 3, 3, 3, generator=g) * 0.3
        b = torch.randn(32, generator=g) * 0.1
        fw = prep.to_device({'w': prep.pack_first_conv(w.double(), b.double())}, dev)['w']
        x = (torch.rand(2, 3, 32, 64, generator=g) * 2 - 1).to(dev)
        for pool in (False, True):
            out = torch.full((2, 16 if pool else 32, 32 if pool else 64, 32), float("nan"), dtype=torch.bfloat16, device=dev)
            eng._first_conv(fw, x, 2, 32, 64, pool, out)
            torch.cuda.synchronize()
            ref = F.leaky_relu(F.conv2d(x, w.to(dev), b.to(dev), padding=1), 0.2)
            if pool: ref = F.max_pool2d(ref, 2, 2)
            summarize(f"first pool={pool}", out, nhwc(ref))
    elif stage == "score":
        lib = nat.load()
        x = torch.rand(3, 3, 64, 64, device=dev); r = torch.rand(3, 3, 64, 64, device=dev)
        score = torch.empty(3, device=dev); mm = torch.empty(3, 2, device=dev); heat = torch.empty(3, 64, 64, device=dev)
        scr = torch.empty(lib.vad_score_scratch_bytes(3, 64, 64), dtype=torch.uint8, device=dev)
        nat.check(lib.vad_score(x.data_ptr(), r.data_ptr(), 3, 64, 64, score.data_ptr(), mm.data_ptr(), heat.data_ptr(),
                                scr.data_ptr(), nat.stream_ptr()), "score")
        torch.cuda.synchronize()
        err = ((x - r) ** 2).mean(1)
        summarize("score heat", heat, err); summarize("score", score, err.mean((1, 2)))
    elif stage in ("gemm64", "gemm32"):
        cin = 64 if stage == "gemm64" else 32
        cout = 64
        for (B, H, W) in ((1, 8, 16), (2, 16, 16)):
            w = torch.randn(cout, cin, 1, 1, generator=g) * 0.2
            b = torch.randn(cout, generator=g) * 0.1
            pk = prep.pack_conv1x1(w.double(), b.double()); pk = prep.to_device({"pk": pk}, dev)["pk"]
            x = rnd_nhwc(B, H, W, cin)
            out = torch.full((B, H, W, cout), float("nan"), dtype=torch.bfloat16, device=dev)
            eng._conv(pk, x, B, H, W, out, 1.0)
            torch.cuda.synchronize()
            ref = F.conv2d(nchw(x), w.to(torch.bfloat16).float().to(dev), b.to(dev))
            summarize(f"{stage} 1x1 {cin}->{cout} B{B} {H}x{W}", out, nhwc(ref))
            if stage == "gemm64" and B == 1:
                # diagnostics: partial-K hypotheses
                wq = w.to(torch.bfloat16).float().to(dev)
                for kk in (16, 32, 48):
                    refp = F.conv2d(nchw(x)[:, :kk], wq[:, :kk], b.to(dev))
                    summarize(f"  hypothesis first {kk} of K", out, nhwc(refp))
    elif stage in ("conv64", "conv32", "pool"):
        cases = {"conv64": [(64, 64, 2, 16, 16, False), (128, 256, 2, 16, 16, False), (64, 128, 2, 24, 40, False)],
                 "conv32": [(32, 32, 2, 32, 32, False), (32, 64, 2, 16, 48, False)],
                 "pool": [(64, 64, 2, 16, 16, True), (32, 32, 2, 32, 32, True), (128, 128, 5, 8, 8, True),
                          (256, 256, 2, 32, 32, True)]}[stage]
        for cin, cout, B, H, W, pool in cases:
            w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
            b = torch.randn(cout, generator=g) * 0.1
            pk = prep.pack_conv3x3(w.double(), b.double()); pk = prep.to_device({"pk": pk}, dev)["pk"]
            x = rnd_nhwc(B, H, W, cin)
            Ho, Wo = (H // 2, W // 2) if pool else (H, W)
            out = torch.full((B, Ho, Wo, cout), float("nan"), dtype=torch.bfloat16, device=dev)
            eng._conv(pk, x, B, H, W, out, 0.2, pool=pool)
            torch.cuda.synchronize()
            ref = F.leaky_relu(F.conv2d(nchw(x), w.to(torch.bfloat16).float().to(dev), b.to(dev), padding=1), 0.2)
            if pool: ref = F.max_pool2d(ref, 2, 2)
            summarize(f"{stage} {cin}->{cout} B{B} {H}x{W}", out, nhwc(ref))
    else:
        rc = subprocess.call([sys.executable, "-m", "pytest", "-q", "--tb=short", "-x", "-m", "gpu",
                              os.path.join(ROOT, "tests", {"convt": "test_gpu_layers.py::test_convt2x2",
                                                           "lstm": "test_gpu_layers.py::test_convlstm_sequence",
                                                           "tanh": "test_gpu_layers.py::test_last_conv_tanh_score",
                                                           "convt_tanh": "test_gpu_layers.py::test_last_convt_tanh_score",
                                                           "image": "test_gpu_models.py::test_image_parity",
                                                           "video": "test_gpu_models.py::test_video_parity"}[stage])])
        print(f"[{stage}] pytest rc={rc}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--stage":
        import time
        t0 = time.time()
        try:
            run_stage(sys.argv[2])
        except Exception as e:  # noqa: BLE001  (bring-up tool: report the trap slot, then re-raise)
            import ctypes
            from models import _native as nat
            trap = (ctypes.c_ulonglong * 4)()
            nat.load().vad_debug_last_trap(trap)
            print(f"[{sys.argv[2]}] FAILED after {time.time() - t0:.1f}s: {str(e).splitlines()[0][:200]}")
            print(f"[{sys.argv[2]}] trap slot: tag={trap[0]} block={trap[1]} thread={trap[2]} parity={trap[3]}", flush=True)
            sys.exit(1)
        sys.exit(0)
    stages = sys.argv[1:] or STAGES
    for s in stages:
        print(f"===== stage {s}", flush=True)
        try:
            rc = subprocess.call([sys.executable, os.path.abspath(__file__), "--stage", s], timeout=240)
        except subprocess.TimeoutExpired:
            rc = "timeout"
        print(f"===== stage {s} rc={rc}", flush=True)
