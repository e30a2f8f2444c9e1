"""Times the native RelativeEncoder (vsg_relenc_forward) at bench size: B = 16 x T = 1000, the FramePriorNetwork stack
(4 layers, per-frame condition) and the PitchPredictor stack (6 layers, speaker condition), fp32 and bf16 modes, next to
the PyTorch statement of the same module on the GPU (cuDNN / cuBLAS, TF32 allowed as in round 1's throughput mode).

    python tools/time_relenc.py [--batch 16] [--frames 1000] [--reps 5]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import visinger_oracle as O  # noqa: E402  (weights generator only)
from visinger_b200.modules.rel_transformer import RelativeEncoder  # noqa: E402


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, T = a.batch, a.frames
    for name, nl, gin, g_t in (("frame prior (4 layers, g per frame)", 4, 1, True), ("pitch predictor (6 layers, speaker g)", 6, 256, False)):
        sd = O.synth_rel_encoder_state_dict(O.rel_encoder_param_shapes(192, 768, 2, nl, 9, 4, gin), 5)
        m = RelativeEncoder(192, 768, 2, nl, kernel_size=9, window_size=4, gin_channels=gin)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        gen = torch.Generator().manual_seed(1)
        x = torch.randn(B, 192, T, generator=gen).to(dev)
        g = torch.randn(B, gin, T if g_t else 1, generator=gen).to(dev)
        mask = torch.ones(B, 1, T, device=dev)
        flops = nl * (2.0 * B * T * (3 * 192 * 192 + 192 * 192 + 192 * 768 * 9 + 768 * 192) + 4.0 * B * T * T * 192)
        res = {}
        for prec in ("fp32", "bf16"):
            m.precision = prec
            res[prec] = (timeit(lambda: m(x, mask, g), a.reps), m(x, mask, g))
        m.native = False
        with torch.no_grad():
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
            t_torch = timeit(lambda: m(x, mask, g), a.reps)
            y_torch = m(x, mask, g)
        rel = float((res["bf16"][1] - res["fp32"][1]).norm() / res["fp32"][1].norm())
        print(f"{name}: native fp32 {res['fp32'][0]:.2f} ms, native bf16 {res['bf16'][0]:.2f} ms "
              f"({flops / res['bf16'][0] / 1e9:.0f} TFLOP/s), PyTorch TF32 {t_torch:.2f} ms; bf16 vs fp32 rel-L2 {rel:.2e}, "
              f"PyTorch TF32 vs native fp32 max-abs {float((y_torch - res['fp32'][1]).abs().max()):.2e}", flush=True)


if __name__ == "__main__":
    main()
