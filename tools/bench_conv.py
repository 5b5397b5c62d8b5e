"""Conv micro-benchmark: TFLOP/s of single conv shapes, optionally with the copy-skipping debug flags."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiolcm_b200 import _lib  # noqa: E402

lib = _lib.load()
ctx = _lib.ctx(0)
SHAPES = [  # B, Cin, Cout, T, K, d
    (1, 768, 768, 2500, 11, 1),
    (8, 768, 768, 2500, 11, 1),
    (8, 768, 768, 2500, 3, 1),
    (1, 1536, 1536, 312, 3, 1),
    (1, 768, 768, 312, 3, 1),
    (1, 768, 768, 624, 3, 1),
    (1, 384, 384, 624, 3, 1),
    (1, 1536, 1536, 312, 1, 1),
    (1, 384, 384, 10000, 7, 1),
    (1, 192, 192, 20000, 7, 1),
    (1, 96, 96, 40000, 7, 1),
    (1, 48, 48, 80000, 7, 1),
    (1, 32, 32, 160000, 7, 1),
    (1, 32, 32, 160000, 11, 1),
    (8, 384, 384, 10000, 7, 1),
    (8, 192, 192, 20000, 7, 1),
    (8, 96, 96, 40000, 7, 1),
    (8, 32, 32, 160000, 7, 1),
    (64, 24, 24, 160000, 3, 1),
    (64, 24, 24, 160000, 11, 5),
    (64, 48, 48, 80000, 11, 1),
    (64, 96, 96, 40000, 11, 1),
    (64, 192, 192, 20000, 11, 1),
]
if len(sys.argv) > 3:
    SHAPES = SHAPES[-int(sys.argv[3]):]
precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16"]
dbgs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3]
for prec in precs:
    for (B, Cin, Cout, T, K, d) in SHAPES:
        fl = 2.0 * B * Cin * Cout * K * T
        out = []
        for dbg in dbgs:
            ms = C.c_float()
            _lib.check(lib.alcm_bench_conv(ctx, B, Cin, Cout, T, K, d, _lib.PREC[prec], 100, dbg, C.byref(ms)))
            out.append(f"dbg{dbg}: {ms.value * 1e3:8.1f} us {fl / ms.value / 1e9:7.1f} TF/s")
        print(f"{prec} B={B} C={Cin}->{Cout} T={T} K={K}: " + " | ".join(out), flush=True)
