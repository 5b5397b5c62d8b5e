"""Activation1d micro-benchmark: algorithmic GB/s (fp32 in + operand out, unpadded channels) of single launches on random
operands.  ALCM_ACT_VARIANT=n forces the block size (0: 128 threads / 635 outputs, 1: 64 / 315 (default), 2: 32 / 155, 3: 64 / 315 at 96 registers)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiolcm_b200 import _lib  # noqa: E402

lib = _lib.load()
ctx = _lib.ctx(0)
SHAPES = [(1, 768, 2500), (1, 384, 10000), (1, 192, 20000), (1, 96, 40000), (1, 48, 80000), (1, 24, 160000), (8, 768, 2500), (8, 192, 20000), (8, 24, 160000), (64, 768, 2500), (64, 24, 160000)]
for prec in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16", "tf32"]):
    osz = 2 if prec == "bf16" else 4
    for (B, Cc, T) in SHAPES:
        byt = B * Cc * T * (4 + osz)   # algorithmic, unpadded channels
        ms = C.c_float()
        _lib.check(lib.alcm_bench_act(ctx, B, Cc, T, _lib.PREC[prec], 100, C.byref(ms)))
        print(f"{prec} B={B} C={Cc} T={T} ({byt / 1e6:.0f} MB): {ms.value * 1e3:8.1f} us {byt / ms.value / 1e6:7.0f} GB/s", flush=True)
