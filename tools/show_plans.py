"""Print the tile plans of every decoder convolution shape at bench size (host-only; no GPU needed)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visinger_b200 import _lib
L = _lib.lib()
L.vsg_debug_plan.argtypes = [ctypes.c_int32] * 9
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x3 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for C, Lq in ((256, 5000), (128, 25000), (64, 75000), (32, 150000), (16, 300000)):
    for k in (3, 7, 11):
        for (na, no) in ((0, 1), (1, 2), (1, 1), (2, 1)):
            L.vsg_debug_plan(C, C, k, 1, B, Lq, na, no, x3)
