"""Where does the bf16 error of the scoring path enter?  CPU emulation of the kernel pipeline (TEST / ANALYSIS TOOL).

Emulates the product path with torch on the CPU: eval-BN folded in fp64, weights rounded to bf16, fp32 accumulation,
activations rounded to bf16 at every layer boundary the kernels have (fused tails keep their intermediate in bf16 in
shared memory, so they round too), and compares per-image / per-frame scores with the fp32 oracle under the stress
weights.  Each row switches ONE rounding point off (keeps it fp32) to show its share of the error.
    python tools/bf16_error_budget.py
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-anomaly-detection_b200"))
sys.path.insert(0, ROOT)
from models import _prepare as prep  # noqa: E402
from models import ConvAutoencoder  # noqa: E402
from models.video_autoencoder import VideoAutoencoder  # noqa: E402
from oracle import vad_oracle  # noqa: E402
from oracle.stress import stress_state_dict  # noqa: E402


def bf(t, on=True):
    return t.to(torch.bfloat16).float() if on else t


def image_emulated(sd, x, keep=()):
    """keep: names of rounding points left in fp32: 'x' (input), 'w:<layer>', 'a:<layer>' (that layer's output)."""
    def W(name, conv, bn, convt=False):
        w, b = (prep.fold_convt if convt else prep.fold_conv)(sd, conv, bn)
        return bf(w.float(), f"w:{name}" not in keep), b.float()
    a = bf(x, "x" not in keep)
    for blk in ("enc1", "enc2", "enc3", "enc4"):
        for i, bn in ((0, 1), (3, 4)):
            name = f"{blk}.{i}"
            w, b = W(name, f"encoder.{blk}.{i}", f"encoder.{blk}.{bn}")
            a = F.leaky_relu(F.conv2d(a, w, b, padding=1), 0.2)
            if i == 3:
                a = F.max_pool2d(a, 2, 2)
            a = bf(a, f"a:{name}" not in keep)
    for blk in ("dec1", "dec2", "dec3", "dec4"):
        name = f"{blk}.0"
        w, b = W(name, f"decoder.{blk}.0", f"decoder.{blk}.1", convt=True)
        a = bf(F.relu(F.conv_transpose2d(a, w, b, stride=2)), f"a:{name}" not in keep)
        name = f"{blk}.3"
        if blk != "dec4":
            w, b = W(name, f"decoder.{blk}.3", f"decoder.{blk}.4")
            a = bf(F.relu(F.conv2d(a, w, b, padding=1)), f"a:{name}" not in keep)
        else:
            w, b = W(name, "decoder.dec4.3", None)
            a = torch.tanh(F.conv2d(a, w, b, padding=1))
    return ((x - a) ** 2).mean(dim=[1, 2, 3])


def video_emulated(sd, x, keep=()):
    B, T = x.shape[:2]
    def W(name, conv, bn, convt=False):
        w, b = (prep.fold_convt if convt else prep.fold_conv)(sd, conv, bn)
        return bf(w.float(), f"w:{name}" not in keep), b.float()
    a = bf(x.reshape(B * T, *x.shape[2:]), "x" not in keep)
    for i in (0, 4, 8, 12):
        name = f"enc.{i}"
        w, b = W(name, f"encoder.encoder.{i}", f"encoder.encoder.{i + 1}")
        a = bf(F.max_pool2d(F.leaky_relu(F.conv2d(a, w, b, padding=1), 0.2), 2, 2), f"a:{name}" not in keep)
    seq = a.reshape(B, T, *a.shape[1:])
    for layer in range(2):
        name = f"lstm.{layer}"
        w = bf(sd[f"convlstm.cells.{layer}.conv.weight"].float(), f"w:{name}" not in keep)
        b = sd[f"convlstm.cells.{layer}.conv.bias"].float()
        hid = w.shape[0] // 4
        h = torch.zeros(B, hid, *seq.shape[3:])
        c = torch.zeros_like(h)
        outs = []
        for t in range(T):
            g = F.conv2d(torch.cat([seq[:, t], h], 1), w, b, padding=1)
            gi, gf, gg, go = torch.split(g, hid, 1)
            c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
            h = bf(torch.sigmoid(go) * torch.tanh(c), f"a:{name}" not in keep)
            outs.append(h)
        seq = torch.stack(outs, 1)
    a = seq.reshape(B * T, *seq.shape[2:])
    for i in (0, 3, 6):
        name = f"dec.{i}"
        w, b = W(name, f"decoder.decoder.{i}", f"decoder.decoder.{i + 1}", convt=True)
        a = bf(F.relu(F.conv_transpose2d(a, w, b, stride=2)), f"a:{name}" not in keep)
    w, b = W("dec.9", "decoder.decoder.9", None, convt=True)
    a = torch.tanh(F.conv_transpose2d(a, w, b, stride=2))
    return ((x.reshape(B * T, *x.shape[2:]) - a) ** 2).mean(dim=[1, 2, 3]).reshape(B, T)


def rel(a, b):
    return float(((a - b).abs() / b.abs()).max())


def main():
    torch.set_num_threads(8)
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        m = ConvAutoencoder()
        sd = stress_state_dict(m.state_dict(), seed=1)
        x = ((0.3 + 0.7 * torch.rand(8, 1, 1, 1, generator=g)) * (2 * torch.rand(8, 3, 128, 128, generator=g) - 1)).clamp(-1, 1)
        ref = vad_oracle.image_reconstruction_error(sd, x)
        layers = [f"{b}.{i}" for b in ("enc1", "enc2", "enc3", "enc4") for i in (0, 3)] + \
                 [f"{b}.{i}" for b in ("dec1", "dec2", "dec3", "dec4") for i in (0, 3)]
        print("image model, stress weights, 8 x 128x128: max score rel err vs fp32 oracle")
        print(f"  all rounding points on (the product path)      {rel(image_emulated(sd, x), ref):.3e}")
        print(f"  input x kept fp32 (hi+lo split of the first conv) {rel(image_emulated(sd, x, ('x',)), ref):.3e}")
        print(f"  + first-conv weights fp32                        {rel(image_emulated(sd, x, ('x', 'w:enc1.0')), ref):.3e}")
        allw = tuple(f"w:{l}" for l in layers)
        alla = tuple(f"a:{l}" for l in layers)
        print(f"  all weights fp32 (activations bf16)              {rel(image_emulated(sd, x, allw), ref):.3e}")
        print(f"  all activations fp32 (weights, x bf16)           {rel(image_emulated(sd, x, alla), ref):.3e}")
        print(f"  nothing rounded (sanity)                         {rel(image_emulated(sd, x, allw + alla + ('x',)), ref):.3e}")
        for l in layers:
            print(f"  only a:{l:7s} kept fp32  {rel(image_emulated(sd, x, ('a:' + l,)), ref):.3e}"
                  f"   only w:{l:7s} kept fp32  {rel(image_emulated(sd, x, ('w:' + l,)), ref):.3e}")
        v = VideoAutoencoder()
        sdv = stress_state_dict(v.state_dict(), seed=1)
        xv = ((0.3 + 0.7 * torch.rand(2, 8, 1, 1, 1, generator=g)) * (2 * torch.rand(2, 8, 3, 64, 64, generator=g) - 1)).clamp(-1, 1)
        refv = vad_oracle.video_reconstruction_error(sdv, xv, per_frame=True)
        vl = ["enc.0", "enc.4", "enc.8", "enc.12", "lstm.0", "lstm.1", "dec.0", "dec.3", "dec.6", "dec.9"]
        print("video model, stress weights, 2 x 8 x 64x64: max frame-score rel err vs fp32 oracle")
        print(f"  all rounding points on     {rel(video_emulated(sdv, xv), refv):.3e}")
        print(f"  input x kept fp32          {rel(video_emulated(sdv, xv, ('x',)), refv):.3e}")
        print(f"  all weights fp32           {rel(video_emulated(sdv, xv, tuple('w:' + l for l in vl)), refv):.3e}")
        print(f"  all activations fp32       {rel(video_emulated(sdv, xv, tuple('a:' + l for l in vl)), refv):.3e}")
        for l in vl:
            print(f"  only a:{l:7s} kept fp32  {rel(video_emulated(sdv, xv, ('a:' + l,)), refv):.3e}"
                  f"   only w:{l:7s} kept fp32  {rel(video_emulated(sdv, xv, ('w:' + l,)), refv):.3e}")


if __name__ == "__main__":
    main()
