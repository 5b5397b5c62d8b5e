"""Run a few full-size decodes with an NVTX range around the last one (for ncu launch lists).

    python tools/one_decode.py [precision] [batch] [t_lat]
    ncu --nvtx --nvtx-include "alcm_decode/" --metrics gpu__time_duration.sum ... python tools/one_decode.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_pipe, T_LAT  # noqa: E402
from audiolcm_b200 import synth  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t_lat = int(sys.argv[3]) if len(sys.argv) > 3 else T_LAT
pipe = build_pipe(prec, "cuda:0")
z = torch.from_numpy(synth.synth_latent(B, t_lat, seed=0)).to("cuda:0")
for _ in range(2):
    pipe.decode_tensor(z)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("alcm_decode")
wav = pipe.decode_tensor(z)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("ok", tuple(wav.shape), float(wav.abs().max()))
