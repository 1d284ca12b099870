"""Groundwork for DESIGN.md §9 item 1 (pixel-pair folding of the 32-channel 3x3 layers): a CPU proof that the folded
GEMM is the convolution, and the weight layout the kernel will need.  No GPU.

    python tools/pair_fold_check.py [Cin Cout H W]

Formulation: one GEMM row = two horizontally adjacent output pixels (x = 2p, 2p+1).  The A operand of row p for vertical
tap ky is the window of FOUR input pixels (2p-1 .. 2p+2) of input row y+ky-1, i.e. K = 3 ky x 4 window columns x Cin;
N = 2 x Cout (column = pix * Cout + co).  The weight of (pix, co) for window column wc is w[co][ky][wc - pix] when
0 <= wc - pix <= 2 and zero otherwise.  With the NHWC tensor viewed as [H][W/2][2*Cin] (128-byte rows for Cin = 32) the
four window pixels are: second half of pair p-1, both halves of pair p, first half of pair p+1 — three row-shifted
descriptors, the K slices chosen inside the 128-byte row.  MMAs per 256 pixels: 3 x 4 x Cin/16 of N = 2*Cout
(Cin = Cout = 32: 24 MMAs x 48 cycles = 576 cycles per 128 pixels against 18 x 40 = 720 for the tap-per-MMA form).
"""
import sys

import torch
import torch.nn.functional as F


def fold_weights(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [2*Cout, 3*4*Cin]: row = pix*Cout + co, column = (ky*4 + wc)*Cin + ci."""
    cout, cin = w.shape[:2]
    out = w.new_zeros(2, cout, 3, 4, cin)
    for pix in range(2):
        for kx in range(3):
            out[pix, :, :, pix + kx, :] = w[:, :, :, kx].permute(0, 2, 1)  # [co, ky, ci]
    return out.reshape(2 * cout, 12 * cin)


def folded_conv(x: torch.Tensor, wf: torch.Tensor, cout: int) -> torch.Tensor:
    """x [B, Cin, H, W] (W even) -> conv3x3(pad 1) via the pair-folded GEMM."""
    B, cin, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1))                       # zero padding = the conv's padding (TMA out-of-bounds fill)
    rows = []
    for ky in range(3):
        for wc in range(4):                           # window column wc <-> input x = 2p - 1 + wc
            rows.append(xp[:, :, ky:ky + H, wc:wc + W:2])          # [B, Cin, H, W/2]
    a = torch.stack(rows, 1).permute(0, 3, 4, 1, 2).reshape(B * H * (W // 2), 12 * cin)   # K = (ky*4 + wc)*Cin + ci
    d = a @ wf.t()                                    # [M pairs, 2*Cout]
    return d.view(B, H, W // 2, 2, cout).permute(0, 4, 1, 2, 3).reshape(B, cout, H, W)


def main() -> None:
    cin, cout, H, W = (int(v) for v in (sys.argv[1:5] + ["32", "32", "12", "20"][len(sys.argv) - 1:]))
    g = torch.Generator().manual_seed(0)
    w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, cin, H, W, generator=g, dtype=torch.float64)
    wf = fold_weights(w)
    ref = F.conv2d(x, w, padding=1)
    got = folded_conv(x, wf, cout)
    err = (got - ref).abs().max().item()
    nz = (wf != 0).double().mean().item()
    print(f"pair-folded conv {cin}->{cout} {H}x{W}: max |err| {err:.2e}; weight matrix {tuple(wf.shape)}, "
          f"{100 * nz:.0f} % non-zero (K padded 9 -> 12 window slots per output pixel)")
    assert err < 1e-10
    k_steps = 12 * cin // 16
    n = 2 * cout
    cyc = max(n // 2, 32 + n // 4)
    print(f"MMAs per 256 pixels: {k_steps} of N = {n} ({cyc} cycles each, DESIGN §4.2) = {k_steps * cyc // 2} cycles "
          f"per 128 pixels; tap-per-MMA form: {9 * cin // 16} x {max(cout // 2, 32 + cout // 4)} = "
          f"{9 * cin // 16 * max(cout // 2, 32 + cout // 4)}")


if __name__ == "__main__":
    main()
