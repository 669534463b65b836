"""Micro-benchmark of the conv engine on the ConvGRU / GResBlock shapes of config 2 (B=64).
usage: [DVD_OPTIONS=oneacc=1,...] python profiles/conv_microbench.py [--reps R] [--only NAME] [--err] [--acc]
Prints TFLOP/s (algorithmic, 2*MAC) per shape for forward and weight-gradient, and with --err the rel-L2 error
of the forward against an fp64 CPU reference on a slice."""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvdgan_b200 import ops  # noqa: E402

SHAPES = {
    # name: (N, Cin, Cout, H, W, k)
    "s9_cell1_h_ur": (64, 256, 512, 32, 32, 5),
    "s9_cell2_h_ur": (64, 128, 256, 32, 32, 5),
    "s6_cell1_h_ur": (64, 512, 1024, 16, 16, 5),
    "s6_cell0_h_ur": (64, 256, 512, 16, 16, 3),
    "s3_cell1_h_ur": (64, 512, 1024, 8, 8, 5),
    "s0_cell1_h_ur": (64, 512, 1024, 4, 4, 5),
    "s9_cell1_x_all": (768, 128, 768, 32, 32, 5),      # a quarter of the batched x-half (B*T = 3072 frames)
    "gres_64": (512, 128, 64, 64, 64, 3),
    "gres_32": (3072, 128, 128, 32, 32, 3),           # GResBlock 3x3 convs over all B*T frames (short K)
    "gres_16": (3072, 256, 256, 16, 16, 3),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default=None)
    ap.add_argument("--err", action="store_true")
    ap.add_argument("--acc", action="store_true", help="forward in accumulate mode (y += conv(x)), the per-timestep GEMMs' epilogue")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    print("options:", os.environ.get("DVD_OPTIONS", "(defaults)"), "accumulate" if a.acc else "")
    for name, (N, Ci, Co, H, W, k) in SHAPES.items():
        if a.only and a.only != name:
            continue
        torch.manual_seed(0)
        x = torch.randn(N, Ci, H, W, device=dev)
        w = torch.randn(Co, Ci, k, k, device=dev) / (Ci * k * k) ** 0.5
        dy = torch.randn(N, Co, H, W, device=dev)
        wp = ops.pack_weight(w)
        flops = 2.0 * N * H * W * Co * Ci * k * k
        res = {}
        ybuf = torch.zeros(N, Co, H, W, device=dev)
        for what in ("fwd", "wgrad"):
            fn = (lambda: ops.conv_raw(x, wp, None, Co, (1, k, k), out=ybuf, accumulate=int(a.acc))) if what == "fwd" \
                else (lambda: ops.wgrad_raw(x, dy, (1, k, k)))
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            res[what] = (ms, flops / ms / 1e9)
        line = f"{name:16s} M={N*H*W:8d} K={Ci*k*k:6d} N={Co:5d}  fwd {res['fwd'][0]:8.3f} ms {res['fwd'][1]:7.1f} TF/s" \
               f"   wgrad {res['wgrad'][0]:8.3f} ms {res['wgrad'][1]:7.1f} TF/s"
        if a.err:
            y = ops.conv_raw(x, wp, None, Co, (1, k, k))
            ref = F.conv2d(x[:2].double().cpu(), w.double().cpu(), padding=k // 2)
            e = float((y[:2].double().cpu() - ref).norm() / ref.norm())
            line += f"   fwd rel-L2 vs fp64: {e:.2e}"
            nb = max(4, 4096 // (H * W))
            if nb <= N and Ci * Co * k * k <= 4e6:
                g = ops.unpack_wgrad(ops.wgrad_raw(x[:nb].contiguous(), dy[:nb].contiguous(), (1, k, k)), w)
                gref = torch.nn.grad.conv2d_weight(x[:nb].double().cpu(), w.shape, dy[:nb].double().cpu(), padding=k // 2)
                line += f"  wgrad: {float((g.double().cpu() - gref).norm() / gref.norm()):.2e}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
