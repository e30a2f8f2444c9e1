"""Autotune the tile plans of the tensor-core convolution on a B200: for every decoder conv shape try every feasible
(blocks per tile, epilogue chunk, one/two CTAs per SM, resident weights) plan, time it with CUDA events, and write the
winners to visinger_b200/csrc/tc_plan_table.inc (compiled into the library; the launcher falls back to its cycle model
for shapes not in the table).   python tools/tune_plans.py [--x3] [--out file]"""
import argparse
import ctypes
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from visinger_b200 import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--x3", type=int, default=0)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tuned_plans.json"))
    args = ap.parse_args()
    L = _lib.lib()
    L.vsg_debug_set_plan.argtypes = [ctypes.c_int32] * 5
    L.vsg_debug_last_ms.restype = ctypes.c_float
    dev = torch.device("cuda:0")
    shapes = [(256, 5000), (128, 25000), (64, 75000), (32, 150000), (16, 300000)]
    results = []
    planes = 2 if args.x3 else 1
    for (C, Lq), k in itertools.product(shapes, (3, 7, 11)):
        B = args.batch
        x = torch.randn(B, Lq, planes * C, device=dev).to(torch.bfloat16)
        w = torch.randn(C, C, k) / (C * k) ** 0.5
        b = torch.zeros(C)
        add = torch.randn(B, Lq, planes * C, device=dev).to(torch.bfloat16)
        for (na, no) in ((0, 1), (1, 2), (1, 1), (2, 1)):
            best = None
            for mb, cw, two, res in itertools.product((4, 2, 1), (64, 32, 16), (1, 0), (1, 0)):
                if cw > C or (C <= 64 and cw > 32) or (C >= 256 and cw > 32) or (C > 64 and two) or 2 * mb * min(C, 256) > 512:
                    continue
                L.vsg_debug_set_plan(mb, cw, two, res, args.reps)
                raw = torch.empty_like(add) if no > 1 else None
                act = torch.empty_like(add)
                rc = L.vsg_debug_conv1d_bf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), add.data_ptr() if na > 0 else None,
                                             add.data_ptr() if na > 1 else None, 1.0, None,
                                             raw.data_ptr() if raw is not None else None, act.data_ptr(), B, Lq, C, C, k, 1,
                                             3 | (4 if args.x3 else 0), 0)
                if rc != 0:
                    continue
                ms = float(L.vsg_debug_last_ms())
                if best is None or ms < best[0]:
                    best = (ms, mb, cw, two, res)
            L.vsg_debug_set_plan(0, 0, -1, -1, args.reps)
            raw = torch.empty_like(add) if no > 1 else None
            act = torch.empty_like(add)
            L.vsg_debug_conv1d_bf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), add.data_ptr() if na > 0 else None,
                                    add.data_ptr() if na > 1 else None, 1.0, None, raw.data_ptr() if raw is not None else None,
                                    act.data_ptr(), B, Lq, C, C, k, 1, 3 | (4 if args.x3 else 0), 0)
            auto_ms = float(L.vsg_debug_last_ms())
            print(f"C={C} k={k} adds={na} outs={no}: best {best} | model {auto_ms:.4f} ms", flush=True)
            if best is not None:
                results.append(dict(cin=C, cout=C, k=k, n_adds=na, n_outs=no, x3=args.x3, mb=best[1], cw=best[2], two=best[3],
                                    resident=best[4], ms=best[0], model_ms=auto_ms))
        del x, add
        torch.cuda.empty_cache()
    json.dump(results, open(args.out, "w"), indent=1)
    print("sum best %.3f ms, sum model %.3f ms" % (sum(r["ms"] for r in results), sum(r["model_ms"] for r in results)))


if __name__ == "__main__":
    main()
