"""GPU diagnostic for the tcgen05 conv kernel: prints max errors for a ladder of conv shapes under
the descriptor variants, so a single GPU call tells which piece (plain GEMM, tap shift, N tiling,
k-block pipelining, phases) is wrong.  Not a test; run under a timeout."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiolcm_b200 import ops  # noqa: E402
from tests.util import round_operand  # noqa: E402

DEV = "cuda:0"
CASES = [  # B, Cin, Cout, T, K, d
    (1, 16, 16, 128, 1, 1),
    (1, 16, 16, 128, 3, 1),
    (1, 64, 32, 128, 1, 1),
    (1, 64, 128, 300, 1, 1),
    (1, 64, 128, 300, 3, 1),
    (1, 64, 128, 300, 3, 3),
    (1, 256, 256, 300, 7, 5),
    (2, 768, 768, 250, 11, 5),
]


def run(prec):
    for (B, Cin, Cout, T, K, d) in CASES:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(B, Cin, T, generator=g)
        w = torch.randn(Cout, Cin, K, generator=g) / np.sqrt(Cin * K)
        b = torch.randn(Cout, generator=g) * 0.1
        xr, wr = round_operand(x, prec), round_operand(w, prec)
        ref = F.conv1d(xr.double(), wr.double(), b.double(), dilation=d, padding=(K * d - d) // 2)
        try:
            y = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), None, dilation=d, precision=prec).cpu()
            err = float((y.double() - ref).abs().max())
            print(f"  {prec} Cin={Cin} Cout={Cout} T={T} K={K} d={d}: max err {err:.3e} (ref max {float(ref.abs().max()):.2f})", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"  {prec} Cin={Cin} Cout={Cout} T={T} K={K} d={d}: EXC {e}", flush=True)
            return False
    return True


if __name__ == "__main__":
    swap = os.environ.get("ALCM_DESC_SWAP", "0")
    print(f"== ALCM_DESC_SWAP={swap}", flush=True)
    for prec in sys.argv[1:] or ["fp32", "bf16", "tf32"]:
        if not run(prec):
            sys.exit(1)
