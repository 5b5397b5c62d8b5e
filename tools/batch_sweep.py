"""Per-clip decode time against the batch size of one call: is a 64-clip job better run as one plan or as micro-batches
whose intermediates fit the 126 MB L2?   python tools/batch_sweep.py [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_pipe, T_LAT  # noqa: E402
from audiolcm_b200 import synth  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
pipe = build_pipe(prec, "cuda:0")
z64 = torch.from_numpy(synth.synth_latent(64, T_LAT, seed=0)).to("cuda:0")
for B in (1, 2, 4, 8, 16, 32, 64):
    chunks = [z64[i:i + B].contiguous() for i in range(0, 64, B)]
    for _ in range(2):
        for c in chunks:
            pipe.decode_tensor(c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for c in chunks:
            pipe.decode_tensor(c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{prec}: 64 clips as {64 // B:2d} call(s) of batch {B:2d}: {best:7.2f} ms  ({best / 64:.3f} ms per clip)", flush=True)
