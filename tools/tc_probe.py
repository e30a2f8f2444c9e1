"""GPU probe: does a tcgen05 conv mode (flags) reproduce F.conv1d?  One mode per process so that a
trapping kernel cannot poison the other probes.  Usage: python tools/tc_probe.py FLAGS  -> one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F

from visinger_b200 import _lib

CASES = [  # Cin, Cout, k, dil, B, L
    (64, 64, 3, 1, 1, 128), (64, 64, 3, 1, 2, 300), (256, 256, 3, 1, 2, 300), (256, 256, 11, 5, 2, 300),
    (128, 128, 7, 3, 2, 700), (64, 64, 11, 1, 3, 129), (32, 32, 3, 3, 2, 300), (32, 32, 11, 5, 1, 1000),
    (16, 16, 7, 5, 2, 300), (16, 16, 3, 1, 1, 4000), (192, 512, 7, 1, 2, 100), (512, 256, 3, 1, 2, 200),
    (96, 192, 1, 1, 2, 300), (192, 384, 5, 1, 2, 300),
]


def main():
    flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    dev = torch.device("cuda:0")
    res = {"flags": flags, "cases": []}
    for (cin, cout, k, dil, B, L) in CASES:
        gen = torch.Generator().manual_seed(cin * 1000 + cout + k * 7 + dil)
        x = torch.randn(B, L, cin, generator=gen).to(torch.bfloat16)
        w = (torch.randn(cout, cin, k, generator=gen) / (cin * k) ** 0.5).to(torch.bfloat16).float()
        b = torch.randn(cout, generator=gen) * 0.1
        ref = F.conv1d(x.float().transpose(1, 2).double(), w.double(), b.double(), dilation=dil,
                       padding=(k - 1) * dil // 2).transpose(1, 2)
        try:
            got = _lib.debug_conv1d_bf16(x.to(dev).contiguous(), w, b, dil, flags).cpu().double()
            err = float((got - ref).abs().max())
            res["cases"].append({"case": [cin, cout, k, dil, B, L], "max_abs_err": err, "ok": err < 2e-3})
        except Exception as e:  # noqa: BLE001
            res["cases"].append({"case": [cin, cout, k, dil, B, L], "error": str(e)[:200], "ok": False})
            break
    res["all_ok"] = all(c["ok"] for c in res["cases"]) and len(res["cases"]) == len(CASES)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
