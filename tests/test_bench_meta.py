"""CPU: bench.py's per-layer cost table against the committed ncu DRAM-traffic captures (profiles/): every kernel name
the engine reports has a cost entry, and no kernel moves noticeably more DRAM bytes than its algorithmic bytes (traffic
above the algorithmic figure would mean wasted re-reads; below it means L2 hits between consecutive layers)."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("vad_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("workload", ["cfg2", "cfg3", "cfg4"])
def test_dram_traffic_matches_algorithmic_bytes(bench, workload):
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")))[workload]
    traffic = {k: v for k, v in traffic.items() if not k.startswith("_")}
    kind, B, T, H, W = bench.WORKLOADS[workload][:5]
    checked = 0
    for name, dram in traffic.items():
        if name == "finalize" or name.startswith("_"):
            continue
        flops, byts = bench.layer_cost(kind, name, B, T, H, W)
        assert flops > 0 and byts > 0, name
        assert dram <= 1.15 * byts, f"{name}: {dram / 1e9:.3f} GB of DRAM traffic for {byts / 1e9:.3f} GB algorithmic"
        checked += 1
    assert checked >= 8
    # the kernels that stream far more than L2 holds must also not be far BELOW their algorithmic bytes
    big = {"cfg2": ("first_conv", "enc1.3", "enc1.0+1.3", "dec4.0+4.3+score"), "cfg3": ("decoder.6+9+score",),
           "cfg4": ("first_conv", "encoder.4", "decoder.6+9+score")}[workload]
    for name in big:
        _, byts = bench.layer_cost(kind, name, B, T, H, W)
        assert traffic[name] >= 0.9 * byts, name


def test_fused_kernel_costs_are_sums_of_their_layers_minus_the_intermediate(bench):
    kind, B, T, H, W = bench.WORKLOADS["cfg2"][:5]
    f_fused, b_fused = bench.layer_cost(kind, "dec4.0+4.3+score", B, T, H, W)
    f0, b0 = bench.layer_cost(kind, "dec4.0", B, T, H, W)
    f1, b1 = bench.layer_cost(kind, "dec4.3+score", B, T, H, W)
    assert f_fused == pytest.approx(f0 + f1)
    inter = B * H * W * 32 * 2                    # the 32-channel full-resolution bf16 tensor, written once and read once
    assert b_fused == pytest.approx(b0 + b1 - 2 * inter, rel=1e-6)
    kind, B, T, H, W = bench.WORKLOADS["cfg3"][:5]
    f_fused, b_fused = bench.layer_cost(kind, "decoder.6+9+score", B, T, H, W)
    f0, b0 = bench.layer_cost(kind, "decoder.6", B, T, H, W)
    f1, b1 = bench.layer_cost(kind, "decoder.9+score", B, T, H, W)
    assert f_fused == pytest.approx(f0 + f1)
    inter = B * T * (H // 2) * (W // 2) * 32 * 2
    assert b_fused == pytest.approx(b0 + b1 - 2 * inter, rel=1e-6)
    f2, b2 = bench.layer_cost(kind, "convlstm.0+1", B, T, H, W)
    f_one, b_one = bench.layer_cost(kind, "convlstm.0", B, T, H, W)
    assert f2 == 2 * f_one and b2 == 2 * b_one
