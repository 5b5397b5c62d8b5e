"""Run one conv shape a few times (ncu target).  usage: one_conv.py B Cin Cout T K d [prec] [iters]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiolcm_b200 import _lib  # noqa: E402

B, Cin, Cout, T, K, d = (int(v) for v in sys.argv[1:7])
prec = sys.argv[7] if len(sys.argv) > 7 else "bf16"
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 2
lib = _lib.load()
ctx = _lib.ctx(0)
ms = C.c_float()
_lib.check(lib.alcm_bench_conv(ctx, B, Cin, Cout, T, K, d, _lib.PREC[prec], iters, 0, C.byref(ms)))
print(f"{prec} B={B} C={Cin}->{Cout} T={T} K={K} d={d}: {ms.value * 1e3:.1f} us  {2.0 * B * Cin * Cout * K * T / ms.value / 1e9:.1f} TF/s")
