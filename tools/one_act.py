"""Run one Activation1d shape a few times (ncu target).  usage: one_act.py B C T [prec] [iters]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiolcm_b200 import _lib  # noqa: E402

B, Cc, T = (int(v) for v in sys.argv[1:4])
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 2
lib = _lib.load()
ctx = _lib.ctx(0)
ms = C.c_float()
_lib.check(lib.alcm_bench_act(ctx, B, Cc, T, _lib.PREC[prec], iters, C.byref(ms)))
byt = B * Cc * T * (4 + (2 if prec == "bf16" else 4))  # algorithmic, unpadded
print(f"{prec} B={B} C={Cc} T={T}: {ms.value * 1e3:.1f} us  {byt / ms.value / 1e6:.0f} GB/s (algorithmic bytes {byt / 1e6:.0f} MB)")
