#!/usr/bin/env python
"""Benchmark of the anomaly-scoring hot path (recon + heat map + score) — prints ONE JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1..cfg5] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE.json configs[1]
(cfg2): image ConvAutoencoder, 256 x 3x256x256, random-init weights (torch.manual_seed(0)), fp32 input in [-1,1].
  value  : frames/s with the batch resident in HBM (CUDA events, max over ranks), all outputs produced
           (per-frame score + min/max + per-pixel heat map; the reconstruction stays on-chip).
  e2e    : frames/s through the public Python API from PINNED HOST memory, H2D + D2H inside the timed region.
  roofline: the dominant kernel of the step, timed live with CUDA events, against MEASURED_PEAKS.json.
  cpu_baseline: the oracle's restatement of the reference path on the host cores (bounded sample).
N > 1 (torchrun): each rank scores its own batch (weak scaling); per-frame scores are gathered to rank 0 with NCCL
inside the timed step.  `--impl reference` times the reference's CPU path (oracle port) instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (kind, batch, T, H, W, GFLOP per frame [SURVEY §8d], description)
    "cfg1": ("image", 32, 1, 256, 256, 8.1118, "image AE, synthetic 256x256 RGB, batch 32"),
    "cfg2": ("image", 256, 1, 256, 256, 8.1118, "image AE, MVTec-bottle-shaped synthetic 256x256, batch 256"),
    "cfg3": ("video", 64, 16, 128, 128, 0.75288, "ConvLSTM video AE, 16-frame 128x128 clips, batch 64"),
    # 720p: one 64-frame window of one stream per step and GPU (708 MB of fp32 input); under torchrun this is cfg5's
    # shape — windows sharded over the GPUs, per-frame scores gathered to rank 0 inside the step
    "cfg4": ("video", 1, 64, 720, 1280, 42.349, "ConvLSTM video AE, synthetic 720p stream, one 64-frame window"),
    "cfg5": ("video", 1, 64, 720, 1280, 42.349, "720p clips sharded over the GPUs (one 64-frame clip per GPU and step), "
                                                  "per-frame score gather"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}


def synth_input(kind, B, T, H, W, device, seed=1234):
    """SURVEY §8d recipe: per-frame amplitude a~U[0.3,1], low-passed noise + fine noise, clamp [-1,1]."""
    g = torch.Generator(device=device).manual_seed(seed)
    n = B * T
    amp = 0.3 + 0.7 * torch.rand(n, 1, 1, 1, generator=g, device=device)
    coarse = torch.rand(n, 3, H // 8, W // 8, generator=g, device=device) * 2 - 1
    x = amp * torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)
    x = (x + 0.05 * torch.randn(n, 3, H, W, generator=g, device=device)).clamp_(-1, 1)
    return x.view(B, T, 3, H, W) if kind == "video" else x


def build_model(kind, device):
    torch.manual_seed(0)
    if kind == "image":
        from models import ConvAutoencoder
        return ConvAutoencoder().eval().to(device)
    from models.video_autoencoder import VideoAutoencoder
    return VideoAutoencoder().eval().to(device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * max(mx)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None}


def run_model(model, kind, x, want_heat=True):
    out = model.score_all(x, want_recon=False, want_heat=want_heat)
    return out.score


def layer_cost(kind, name, B, T, H, W):
    """(algorithmic FLOPs, algorithmic HBM bytes) of one launch, by layer name — SURVEY Appendix A conventions:
    2*M*N*K with un-padded K, N; bytes = input + output activations (bf16 NHWC; model input fp32) + outputs."""
    F = B * T
    img = {  # name: (cin, cout, H_in divisor, taps, kind, pooled)
        "first_conv": (3, 32, 1, 9, "conv", kind == "video"),
        "enc1.3": (32, 32, 1, 9, "conv", True), "enc2.0": (32, 64, 2, 9, "conv", False),
        "enc2.3": (64, 64, 2, 9, "conv", True), "enc3.0": (64, 128, 4, 9, "conv", False),
        "enc3.3": (128, 128, 4, 9, "conv", True), "enc4.0": (128, 256, 8, 9, "conv", False),
        "enc4.3": (256, 256, 8, 9, "conv", True), "dec1.0": (256, 128, 16, 1, "convt", False),
        "dec1.3": (128, 128, 8, 9, "conv", False), "dec2.0": (128, 64, 8, 1, "convt", False),
        "dec2.3": (64, 64, 4, 9, "conv", False), "dec3.0": (64, 32, 4, 1, "convt", False),
        "dec3.3": (32, 32, 2, 9, "conv", False), "dec4.0": (32, 32, 2, 1, "convt", False),
        "dec4.3+score": (32, 3, 1, 9, "score", False),
        "encoder.4": (32, 64, 2, 9, "conv", True), "encoder.8": (64, 128, 4, 9, "conv", True),
        "encoder.12": (128, 128, 8, 9, "conv", True), "decoder.0": (128, 128, 16, 1, "convt", False),
        "decoder.3": (128, 64, 8, 1, "convt", False), "decoder.6": (64, 32, 4, 1, "convt", False),
        "decoder.9+score": (32, 3, 2, 1, "scoret", False),
    }
    if name.startswith("convlstm."):  # one launch group = the T steps of one layer (step 0 skips the h half of K)
        h, w = H // 16, W // 16
        flops = 2.0 * B * h * w * 512 * 9 * (128 + 256 * (T - 1))
        byts = B * h * w * ((128 * 2 * 2 + 128 * 2 + 128 * 4 * 2) * T - 128 * 2 - 128 * 4)
        if "+" in name:  # both layers in one wavefront launch
            return 2 * flops, 2 * byts
        return flops, byts
    if name == "decoder.6+9+score":  # fused tail: reads the 64-ch quarter-resolution tensor and x, writes the heat map
        m = F * (H // 4) * (W // 4)
        flops = 2.0 * m * (4 * 32 * 64 + 16 * 3 * 32)
        return flops, m * 64 * 2 + 16 * m * (12 + 4) + F * 12
    if name == "dec4.0+4.3+score":  # fused tail: reads the 32-ch half-resolution tensor and x, writes the heat map
        m = F * (H // 2) * (W // 2)
        flops = 2.0 * m * (4 * 32 * 32 + 4 * 3 * 9 * 32)
        return flops, m * 32 * 2 + 4 * m * (12 + 4) + F * 12
    cin, cout, div, taps, typ, pooled = img[name]
    h, w = H // div, W // div
    m = F * h * w
    in_b = m * cin * (4 if name == "first_conv" else 2)
    if typ == "conv":
        flops = 2.0 * m * cout * taps * cin
        out_b = m * cout * 2 / (4 if pooled else 1)
    elif typ == "convt":
        flops = 2.0 * m * 4 * cout * cin
        out_b = m * 4 * cout * 2
    elif typ == "score":   # + re-read of x fp32, heat map fp32, 12 B of (sum,min,max)
        flops = 2.0 * m * cout * taps * cin
        out_b = m * (12 + 4) + F * 12
    else:                  # last ConvT of the video decoder: 4 output pixels per input pixel
        flops = 2.0 * m * 4 * cout * cin
        out_b = 4 * m * (12 + 4) + F * 12
    return flops, in_b + out_b


def cpu_baseline(kind, T, H, W, budget_s=12.0):
    """The reference's CPU path (oracle restatement: same torch ops, fp32) on the host cores, bounded sample."""
    from oracle import vad_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = build_model(kind, "cpu")
    sd = vad_oracle.cpu_sd(m.state_dict())
    nb = 8 if kind == "image" else 2
    if H * W > 256 * 256:  # 720p: one window of 8 frames (the per-frame cost does not depend on T)
        nb, T = 1, min(T, 8)
    x = synth_input(kind, nb, T, H, W, "cpu")
    fn = (lambda: vad_oracle.image_reconstruction_error(sd, x)) if kind == "image" else \
         (lambda: vad_oracle.video_reconstruction_error(sd, x, per_frame=True))
    with torch.no_grad():
        fn()  # warm-up
        best, t_all, reps = float("inf"), time.perf_counter(), 0
        while reps < 3 or (time.perf_counter() - t_all < budget_s and reps < 50):
            t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0); reps += 1
    return {"value": round(nb * T / best, 2), "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{nb} x {'%dx' % T if kind == 'video' else ''}3x{H}x{W} fp32, best of {reps}, "
                      f"oracle/vad_oracle.py (torch {torch.__version__} CPU ops)"}


def gpu_library_baseline(kind, B, T, H, W, dev, x):
    """The reference's own GPU path — the same torch ops in eager mode on cuDNN (fp32, TF32 allowed as by default) —
    timed with CUDA events on this GPU: the library bar the hand-written kernels are measured against.
    Uses the oracle's restatement of the reference forward on CUDA tensors; scores only (one forward)."""
    from oracle import vad_oracle
    m = build_model(kind, "cpu")
    sd = {k: v.to(dev) for k, v in vad_oracle.cpu_sd(m.state_dict()).items()}
    nb = min(B, 64 if kind == "image" else 8)  # eager fp32 activations of the full batch would not all fit comfortably
    xs = x[:nb]
    fn = (lambda: vad_oracle.image_reconstruction_error(sd, xs)) if kind == "image" else \
         (lambda: vad_oracle.video_reconstruction_error(sd, xs, per_frame=True))
    with torch.no_grad():
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"value": round(nb * T / (ms * 1e-3), 1), "unit": "frames/s", "kind": "torch eager + cuDNN on the same GPU",
            "sample": f"{nb} x {'%dx' % T if kind == 'video' else ''}3x{H}x{W} fp32, scores only, torch {torch.__version__}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gpu-library-baseline", action="store_true",
                    help="also time the reference's torch/cuDNN eager path on this GPU (extra key, rank 0, N=1)")
    args = ap.parse_args()
    kind, B, T, H, W, gflop_per_frame, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload}: {desc}", "frames_per_step_per_gpu": B * T, "input": f"{B}x{'%dx' % T if kind == 'video' else ''}3x{H}x{W} fp32",
              "weights": "random init, torch.manual_seed(0)",
              "l2": f"batch input {B * T * 3 * H * W * 4 / 1e6:.0f} MB > 126 MB L2 (no flush needed)"
              if B * T * 3 * H * W * 4 > 126e6 else "L2 flushed between steps by a 256 MB memset",
              "outputs": "per-frame score + min/max + per-pixel heat map"}

    if args.impl == "reference":
        if rank != 0:
            return
        base = cpu_baseline(kind, T, H, W, budget_s=max(10.0, 2.0 * args.steps))
        line = {"metric": "frames/sec scored (recon+heatmap+score)", "value": base["value"], "unit": "frames/s",
                "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(1000.0 * B * T / base["value"], 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the scoring path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = None
    if world > 1:  # one process per GPU: stage this rank's uploads on the GPU's own socket
        from runtime.sharding import bind_to_gpu_numa_node
        numa_node = bind_to_gpu_numa_node(local_rank)
    dist = None
    stdout_fd = None
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created; stdout must carry the JSON
        # line and nothing else, so file descriptor 1 points at stderr until the line is printed
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from models import _native as nat
    model = build_model(kind, dev)
    x = synth_input(kind, B, T, H, W, dev, seed=1234 + rank)
    flush = None if B * T * 3 * H * W * 4 > 126e6 else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered = [torch.empty(B * T, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def step():
        s = run_model(model, kind, x)
        if world > 1:
            dist.gather(s, gathered, dst=0)   # the path's only exchange: per-frame scores to rank 0 (NCCL / NVLink)
        return s

    for _ in range(max(args.warmup, 3)):
        if flush is not None:
            flush.zero_()
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = nat.launch_count()
    torch.cuda.synchronize()
    if flush is None:
        # inputs exceed L2: one event pair around exactly K back-to-back steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        elapsed_ms = e0.elapsed_time(e1)
    else:
        # small batch: flush L2 between steps, time each step and sum (the flush itself is not counted)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in evs:
            flush.zero_()
            e0.record()
            step()
            e1.record()
        torch.cuda.synchronize()
        elapsed_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    if world > 1:
        dist.barrier()
    launches = nat.launch_count() - launches0
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())

    # ---- per-kernel timing for the roofline: the library brackets every layer launch of the model-level call with CUDA
    # events on the launching stream (vad_profile_enable / vad_profile_dump), same inputs
    reps = 5
    nat.profile_enable(True)
    for _ in range(reps):
        run_model(model, kind, x)
    torch.cuda.synchronize()
    rows = nat.profile_dump()
    nat.profile_enable(False)
    per_run = len(rows) // reps
    per = {}
    for r in range(reps):
        acc = {}
        for name, ms in rows[r * per_run:(r + 1) * per_run]:
            a = acc.setdefault(name, [0.0, 0])
            a[0] += ms
            a[1] += 1
        for name, (ms, n) in acc.items():
            per.setdefault(name, []).append((ms, n))
    # one entry per layer name; a layer launched several times per step (ConvLSTM groups of clips) keeps its launch count
    kern = {}
    small = {}
    for name, vals in per.items():
        ms = statistics.median(v[0] for v in vals)
        if name in ("finalize", "latent_out"):
            small[name] = ms
            continue
        kern[name] = {"ms": ms, "launches": vals[0][1], "rep": name}
    total_kernel_ms = sum(d["ms"] for d in kern.values())
    top_name, top = max(kern.items(), key=lambda kv: kv[1]["ms"])
    pk = peaks()
    flops, byts = layer_cost(kind, top["rep"], B, T, H, W)
    flops, byts = flops / top["launches"], byts / top["launches"]
    avg_ms = top["ms"] / top["launches"]
    ai = flops / byts
    if ai >= pk["bf16_tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9):
        roof = {"bound": "tensor", "achieved": round(flops / (avg_ms * 1e-3) / 1e12, 2), "peak": pk["bf16_tflops"],
                "unit": "TFLOP/s"}
    else:
        roof = {"bound": "hbm", "achieved": round(byts / (avg_ms * 1e-3) / 1e9, 1), "peak": pk["hbm_gbs"], "unit": "GB/s"}
    traffic = None
    for tname in ("r01d_dram_traffic.json", "r01c_dram_traffic.json"):  # newest capture that has this kernel
        tpath = os.path.join(ROOT, "profiles", tname)
        if traffic is None and os.path.exists(tpath):  # measured with ncu --set full at this exact shape (see profiles/)
            traffic = json.load(open(tpath)).get(args.workload, {}).get(top["rep"])
    # every kernel of the step against the roof its arithmetic intensity puts it under (same definitions as above)
    ridge = pk["bf16_tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9)
    per_kernel = {}
    for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"]):
        kf, kb = layer_cost(kind, v["rep"], B, T, H, W)
        kms = v["ms"]  # (whole layer: all its launches of one step)
        if kf / kb >= ridge:
            per_kernel[k] = {"ms": round(v["ms"], 4), "bound": "tensor",
                             "frac": round(kf / (kms * 1e-3) / 1e12 / pk["bf16_tflops"], 3)}
        else:
            per_kernel[k] = {"ms": round(v["ms"], 4), "bound": "hbm",
                             "frac": round(kb / (kms * 1e-3) / 1e9 / pk["hbm_gbs"], 3)}
    roof.update({"frac": round(roof["achieved"] / roof["peak"], 4), "traffic": traffic, "kernel": top_name,
                 "launch_ms": round(avg_ms, 4), "share_of_step": round(top["ms"] / total_kernel_ms, 3),
                 "peak_src": pk["src"], "alg_flops_per_launch": flops, "alg_bytes_per_launch": byts,
                 "per_kernel_ms": {k: round(v["ms"], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])},
                 "per_kernel": per_kernel})
    if top["rep"] in ("dec4.0+4.3+score",):
        roof["note"] = ("fused decoder tail: algorithmic bytes are a third of the two layers it replaces, so the HBM "
                        "fraction is low by construction; the kernel is bound by shared-memory bandwidth (N=16 MMAs "
                        "stream a 4 KB A slab each; ncu smem wavefronts in profiles/)")

    # ---- end to end through the public API from pinned host memory (double-buffered H2D on a side stream)
    xh = x.cpu().pin_memory()
    sh = torch.empty(B * T, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(x) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(n):
        main_stream = torch.cuda.current_stream()
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[b])
                bufs[b].copy_(xh, non_blocking=True)
                ready[b].record(copy_stream)
            main_stream.wait_event(ready[b])
            s = model.get_reconstruction_error(bufs[b], per_frame=True) if kind == "video" else \
                model.get_reconstruction_error(bufs[b])
            freed[b].record(main_stream)
            sh.copy_(s.reshape(-1), non_blocking=True)
        torch.cuda.synchronize()

    e2e_run(3)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None  # sampled across the device-resident and the end-to-end timed regions

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    frames = B * T * world * args.steps
    line = {"metric": "frames/sec scored (recon+heatmap+score)", "value": round(frames / (elapsed_ms * 1e-3), 1),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(elapsed_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "model_tflops": round(frames * gflop_per_frame * 1e9 / (elapsed_ms * 1e-3) / 1e12 / world, 2),
            "roofline": roof, "clocks": clocks,
            "e2e": {"value": round(frames / e2e_s, 1), "unit": "frames/s", "h2d_bytes_per_step": x.numel() * 4,
                    "d2h_bytes_per_step": B * T * 4, "note": "pinned host -> device copies double-buffered on a side stream"
                    + (f"; ranks bound to their GPU's NUMA node (rank 0: node {numa_node})" if numa_node is not None else "")},
            "gpu_launches": int(launches)}
    # extra: the same end-to-end loop fed with the decoder's uint8 HWC frames, normalised on the device
    # (a quarter of the PCIe bytes of the fp32 tensors the reference callers upload)
    if world == 1:
        from runtime import frames as fr
        u8h = ((xh.reshape(-1, 3, H, W).permute(0, 2, 3, 1) * 0.5 + 0.5).clamp_(0, 1) * 255).to(torch.uint8).contiguous().pin_memory()
        ubufs = [torch.empty_like(u8h, device=dev) for _ in range(2)]

        def e2e_u8(n):
            main_stream = torch.cuda.current_stream()
            for i in range(n):
                b = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[b])
                    ubufs[b].copy_(u8h, non_blocking=True)
                    ready[b].record(copy_stream)
                main_stream.wait_event(ready[b])
                xin = fr.normalize_u8(ubufs[b]).view(x.shape)
                s_ = model.get_reconstruction_error(xin, per_frame=True) if kind == "video" else \
                    model.get_reconstruction_error(xin)
                freed[b].record(main_stream)
                sh.copy_(s_.reshape(-1), non_blocking=True)
            torch.cuda.synchronize()

        e2e_u8(3)
        t0 = time.perf_counter()
        e2e_u8(args.steps)
        line["e2e_u8_frames"] = {"value": round(B * T * args.steps / (time.perf_counter() - t0), 1), "unit": "frames/s",
                                 "h2d_bytes_per_step": u8h.numel(),
                                 "note": "uint8 HWC frames uploaded, ToTensor+Normalize on the device (runtime/frames.py)"}
    # apples-to-apples extra: scores only (no heat map written), the one output the reference computes per forward
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run_model(model, kind, x, want_heat=False)
    e0.record()
    for _ in range(min(args.steps, 10)):
        run_model(model, kind, x, want_heat=False)
    e1.record()
    torch.cuda.synchronize()
    line["score_only"] = {"value": round(B * T * min(args.steps, 10) / (e0.elapsed_time(e1) * 1e-3), 1),
                          "unit": "frames/s", "note": "rank 0, per GPU; heat map not materialised"}
    if args.gpu_library_baseline and world == 1:
        line["gpu_library_baseline"] = gpu_library_baseline(kind, B, T, H, W, dev, x)
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(kind, T, H, W)
    if stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        os.close(stdout_fd)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)  # (anything NCCL says while shutting down goes to stderr too)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
