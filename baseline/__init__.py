"""Reference-side restatements that are NOT part of the product path.

``lcm_denoiser_port`` restates the reference's LCM sampler and ConcatDiT2MLP denoiser in plain PyTorch.  The real
modules live under /root/reference, are pure Python and cannot travel to the GPU box; BASELINE.json configs[4]
("end-to-end AudioLCMBatchInfer ... denoiser left as reference PyTorch") needs a denoiser there, so bench.py's
config-5 leg and tests use this port.  It is pinned to the unmodified reference by tests/golden/lcm_denoiser.npz
(oracle/make_golden_lcm.py).  Nothing under audiolcm_b200/ imports it.
"""
