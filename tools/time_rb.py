"""Times the whole-ResBlock1 kernel (csrc/rb_tc.cuh) at bench size for each (C, k) of the model's C <= 64 stages.

    python tools/time_rb.py [--reps 5] [--only C,k] [--sets 4] [--max-mb 0]

MMA floor printed next to each: tiles per SM x blocks x 2*n_pairs convs x k taps x (C/16) x cycles(N) at 1.9 GHz."""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visinger_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--sets", type=int, default=0, help="epilogue sets | (MMA issuer warps << 4)")
    ap.add_argument("--issuers", type=int, default=0)
    ap.add_argument("--rp", type=int, default=0, help="1: row-packed kernel (rp_tc.cuh); 2: row-packed, no Toeplitz form")
    ap.add_argument("--max-mb", type=int, default=0)
    ap.add_argument("--spb2", type=int, default=0, help="1: row-packed kernel with two epilogue warp sets per block (two blocks per set)")
    ap.add_argument("--x3", type=int, default=0, help="1: the row-packed kernel's split-bf16 instantiation (bf16x3 mode; needs --rp 1)")
    ap.add_argument("--two-cta", type=int, default=0, help="1: row-packed kernel as two CTAs per SM (8 epilogue warps, 2-block tiles)")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--frames", type=int, default=1000)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dils = (1, 3, 5)
    for C, rate in ((64, 75), (32, 150), (16, 300)):
        if args.rp and C == 64 and not args.only:
            continue
        for k in (3, 7, 11):
            if args.only and args.only != f"{C},{k}":
                continue
            L = rate * args.frames
            gen = torch.Generator().manual_seed(C + k)
            xa = F.leaky_relu(torch.randn(args.batch, L, C, generator=gen), 0.1)
            xa = (_lib.split_bf16(xa) if args.x3 else xa.to(torch.bfloat16)).to(dev).contiguous()
            ws = [torch.randn(C, C, k, generator=gen) / (C * k) ** 0.5 for _ in range(6)]
            bs = [torch.randn(C, generator=gen) * 0.1 for _ in range(6)]
            add1 = torch.randn(args.batch, L, C, generator=gen)
            add1 = (_lib.split_bf16(add1) if args.x3 else add1.to(torch.bfloat16)).to(dev).contiguous()
            out, raw, act, ms = _lib.debug_resblock_bf16(xa, ws, bs, dils, add1=add1, scale=1 / 3, max_mb=args.max_mb,
                                                         sets=args.sets | (args.issuers << 4) | (256 if args.rp else 0) | (512 if args.rp == 2 else 0) | (1024 if args.spb2 else 0) | (2048 if args.x3 else 0) | (4096 if args.two_cta else 0), want_raw=False, want_f32=False, reps=args.reps)
            mb = args.max_mb or (2 if args.x3 else 512 // (2 * C))
            H = (k - 1) // 2 * 12
            V = 128 * mb - 2 * H
            tiles = -(-L // V) * args.batch
            cyc = {16: 40, 32: 40, 64: 48}[C]
            floor_us = -(-tiles // 148) * mb * 6 * k * (C // 16) * cyc / 1.9e3
            flops = 2.0 * args.batch * L * C * C * k * 6
            print(f"C={C:3d} k={k:2d} L={L}: {ms * 1e3:8.1f} us   MMA floor {floor_us:7.1f} us   {flops / ms / 1e9:7.1f} TFLOP/s "
                  f"(tiles {tiles}, mb {mb}, V {V})", flush=True)


if __name__ == "__main__":
    main()
