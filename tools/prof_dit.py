"""torch.profiler breakdown of one hybrid-DiT forward (B = 128 prompts, T = 312): which PyTorch ops are left around our GEMMs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiolcm_b200.denoiser import ConcatDiT2MLPB200, LCMSamplerB200  # noqa: E402
from baseline.lcm_denoiser_port import dit_state_dict  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = "cuda:0"
dit = ConcatDiT2MLPB200(dit_state_dict(seed=7), dev, prec)
B, T = 128, 312
x = torch.randn(B, 20, T, device=dev)
ctx = torch.randn(B, 154, 1024, device=dev)
t = torch.full((B,), 999, device=dev, dtype=torch.long)
w = LCMSamplerB200.guidance_embedding(torch.tensor(4.0).repeat(B)).to(dev)
for _ in range(2):
    dit(x, t, ctx, w)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as p:
    dit(x, t, ctx, w)
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
