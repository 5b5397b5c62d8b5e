"""Plain-PyTorch restatement of the reference's denoiser side of BASELINE.json configs[4] - reference arithmetic,
functional form, same ATen ops (runs on CPU or CUDA; NOT the product):

  * ``concat_dit2mlp`` : ConcatDiT2MLP.forward        /root/reference/ldm/modules/diffusionmodules/concatDiT.py:285-304
        TimestepEmbedder (:35-76), ConditionEmbedder (:93-104), TemporalTransformer (:132-173),
        BasicTransformerBlock (:108-130), Conv1DFinalLayer (:79-91);
        CrossAttention / Conv1dGEGLU / Conv1dFeedForward / PositionEmbedding
                                                       /root/reference/ldm/modules/new_attention.py:48-130,212-251
  * ``LCMSchedule``    : LCMSampler.set_timesteps / step / get_scalings_for_boundary_condition_discrete /
        get_guidance_scale_embedding                  /root/reference/ldm/models/diffusion/scheduling_lcm.py:87-113,118-259,401-494
        with the linear beta schedule of DDPM.register_schedule
                                                       /root/reference/ldm/models/diffusion/ddpm.py:116-137, util.py:21-25
  * ``lcm_sample``     : LCMSampler.lcm_sampling      scheduling_lcm.py:344-382

Weights are a ``state_dict`` of the reference ``ConcatDiT2MLP`` (numpy or torch values); ``dit_state_dict`` makes
seeded synthetic ones of the shipped architecture (configs/audiolcm.yaml:39-47).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

DIT_CFG = dict(in_channels=20, context_dim=1024, hidden_size=576, num_heads=8, depth=4, max_len=1000)   # audiolcm.yaml:39-47
LCM_CFG = dict(timesteps=1000, linear_start=0.00085, linear_end=0.012, num_ddim_timesteps=50)           # audiolcm.yaml:5-9,20


def dit_tensor_shapes(cfg=DIT_CFG):
    """state_dict names and shapes of ConcatDiT2MLP(**cfg), in module order."""
    H, Cin, ctx, L = cfg["hidden_size"], cfg["in_channels"], cfg["context_dim"], cfg["max_len"]
    s = {"t_embedder.mlp.0.weight": (H, 256), "t_embedder.mlp.0.bias": (H,), "t_embedder.mlp.2.weight": (H, H), "t_embedder.mlp.2.bias": (H,),
         "t_embedder.proj_w.weight": (256, 256)}
    for c in ("c1_embedder", "c2_embedder"):
        s.update({f"{c}.mlp.0.weight": (H, ctx), f"{c}.mlp.0.bias": (H,), f"{c}.mlp.2.weight": (H, H), f"{c}.mlp.2.bias": (H,),
                  f"{c}.mlp.3.weight": (H,), f"{c}.mlp.3.bias": (H,)})
    s.update({"proj_in.weight": (H, Cin, 5), "proj_in.bias": (H,), "pos_emb.weight": (L, H)})
    for i in range(cfg["depth"]):
        p = f"blocks.{i}"
        s.update({f"{p}.norm.weight": (H,), f"{p}.norm.bias": (H,), f"{p}.proj_in.weight": (H, H, 1), f"{p}.proj_in.bias": (H,)})
        t = f"{p}.transformer_blocks.0"
        for a in ("attn1", "attn2"):
            s.update({f"{t}.{a}.to_q.weight": (H, H), f"{t}.{a}.to_k.weight": (H, H), f"{t}.{a}.to_v.weight": (H, H),
                      f"{t}.{a}.to_out.0.weight": (H, H), f"{t}.{a}.to_out.0.bias": (H,)})
        s.update({f"{t}.ff.net.0.proj.weight": (8 * H, H, 9), f"{t}.ff.net.0.proj.bias": (8 * H,),
                  f"{t}.ff.net.2.weight": (H, 4 * H, 9), f"{t}.ff.net.2.bias": (H,)})
        for n in ("norm1", "norm2", "norm3"):
            s.update({f"{t}.{n}.weight": (H,), f"{t}.{n}.bias": (H,)})
        s.update({f"{p}.proj_out.weight": (H, H, 1), f"{p}.proj_out.bias": (H,)})
    s.update({"final_layer.norm_final.weight": (H,), "final_layer.norm_final.bias": (H,),
              "final_layer.conv1d.weight": (Cin, H, 1), "final_layer.conv1d.bias": (Cin,)})
    return s


def dit_state_dict(cfg=DIT_CFG, seed=0):
    """Seeded synthetic weights (numpy PCG64): fan-in scaled matrices, norms ~ (1, 0) + noise, small biases; the
    reference's zero-initialised proj_out (new_attention.py:76-82) is randomised so that every path is exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for name, shape in dit_tensor_shapes(cfg).items():
        if name.endswith("bias"):
            v = rng.standard_normal(shape) * 0.02
        elif ".norm" in name or "norm_final" in name or name.endswith("mlp.3.weight"):
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name == "pos_emb.weight":
            v = rng.standard_normal(shape) * 0.05
        else:
            fan_in = int(np.prod(shape[1:]))
            v = rng.standard_normal(shape) / math.sqrt(fan_in)
        sd[name] = v.astype(np.float32)
    return sd


def _t(sd, device, dtype=torch.float32):
    return {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v).to(device=device, dtype=dtype) for k, v in sd.items()}


def timestep_embedding(t, dim=256, max_period=10000):            # concatDiT.py:49-69
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _attention(sd, p, x, heads):                                   # new_attention.py:107-130 (context=None: self-attention)
    B, N, C = x.shape
    d = C // heads
    q, k, v = (F.linear(x, sd[f"{p}.to_{n}.weight"]).view(B, N, heads, d).permute(0, 2, 1, 3).reshape(B * heads, N, d) for n in "qkv")
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    out = torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v)
    out = out.view(B, heads, N, d).permute(0, 2, 1, 3).reshape(B, N, C)
    return F.linear(out, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])


def _cond_embed(sd, p, c):                                         # concatDiT.py:93-104
    h = F.gelu(F.linear(c, sd[f"{p}.mlp.0.weight"], sd[f"{p}.mlp.0.bias"]), approximate="tanh")
    h = F.linear(h, sd[f"{p}.mlp.2.weight"], sd[f"{p}.mlp.2.bias"])
    return F.layer_norm(h, h.shape[-1:], sd[f"{p}.mlp.3.weight"], sd[f"{p}.mlp.3.bias"])


def concat_dit2mlp(sd, x, t, context, w_cond=None, cfg=DIT_CFG):
    """x (N,C,T), t (N,) long, context (N,2L,ctx), w_cond (N,256) -> (N,C,T).  sd: torch tensors on x's device."""
    heads = cfg["num_heads"]
    t_freq = timestep_embedding(t, 256)
    if w_cond is not None:
        t_freq = t_freq + F.linear(w_cond, sd["t_embedder.proj_w.weight"])
    temb = F.linear(F.silu(F.linear(t_freq, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                    sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"]).unsqueeze(1)
    c1, c2 = context.chunk(2, dim=1)
    c = torch.cat((_cond_embed(sd, "c1_embedder", c1), _cond_embed(sd, "c2_embedder", c2)), dim=1)
    extra = c.shape[1] + 1
    h = F.conv1d(x, sd["proj_in.weight"], sd["proj_in.bias"], padding=2).permute(0, 2, 1)
    h = torch.cat([temb, c, h], dim=1)
    h = h + sd["pos_emb.weight"][: h.shape[1]].unsqueeze(0)        # PositionEmbedding MODE_ADD, new_attention.py:245-248
    h = h.permute(0, 2, 1)                                          # (N, H, extra+T)
    for i in range(cfg["depth"]):
        p, tb = f"blocks.{i}", f"blocks.{i}.transformer_blocks.0"
        x_in = h
        y = F.group_norm(h, 32, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-6)
        y = F.conv1d(y, sd[f"{p}.proj_in.weight"], sd[f"{p}.proj_in.bias"]).permute(0, 2, 1)
        ln = lambda v, n: F.layer_norm(v, v.shape[-1:], sd[f"{tb}.{n}.weight"], sd[f"{tb}.{n}.bias"])
        y = _attention(sd, f"{tb}.attn1", ln(y, "norm1"), heads) + y
        y = _attention(sd, f"{tb}.attn2", ln(y, "norm2"), heads) + y
        f = F.conv1d(ln(y, "norm3").permute(0, 2, 1), sd[f"{tb}.ff.net.0.proj.weight"], sd[f"{tb}.ff.net.0.proj.bias"], padding=4)
        a, gate = f.chunk(2, dim=1)
        f = F.conv1d(a * F.gelu(gate), sd[f"{tb}.ff.net.2.weight"], sd[f"{tb}.ff.net.2.bias"], padding=4)
        y = (f.permute(0, 2, 1) + y).permute(0, 2, 1)
        h = F.conv1d(y, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"]) + x_in
    h = h[..., extra:]
    h = F.group_norm(h, 16, sd["final_layer.norm_final.weight"], sd["final_layer.norm_final.bias"])
    return F.conv1d(h, sd["final_layer.conv1d.weight"], sd["final_layer.conv1d.bias"])


class LCMSchedule(object):
    """Scalar side of LCMSampler (epsilon prediction, timestep_scaling 10, sigma_data 0.5)."""

    def __init__(self, cfg=LCM_CFG):
        n = cfg["timesteps"]
        betas = torch.linspace(cfg["linear_start"] ** 0.5, cfg["linear_end"] ** 0.5, n, dtype=torch.float64) ** 2   # util.py:22-25
        self.alphas_cumprod = torch.from_numpy(np.cumprod(1.0 - betas.numpy(), axis=0)).to(torch.float32)          # ddpm.py:123-136
        self.num_train = n
        self.original_inference_steps = cfg["num_ddim_timesteps"]
        self.timestep_scaling, self.sigma_data = 10.0, 0.5

    def timesteps(self, num_inference_steps, original_inference_steps=None):   # scheduling_lcm.py:158-166,247-253
        orig = original_inference_steps or self.original_inference_steps
        k = self.num_train // orig
        origin = (np.arange(1, orig + 1) * k - 1)[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
        return [int(v) for v in origin[idx]]

    def coefficients(self, ts, i):
        """step i of schedule ts -> (c_x0_sample, c_x0_eps, c_out, c_skip, a_prev_sqrt, b_prev_sqrt, last):
        x0 = c_x0_sample*sample - c_x0_eps*eps ; denoised = c_out*x0 + c_skip*sample ;
        prev = a_prev_sqrt*denoised + b_prev_sqrt*noise  (scheduling_lcm.py:441-487)."""
        t = ts[i]
        prev_t = ts[i + 1] if i + 1 < len(ts) else t
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
        st = t * self.timestep_scaling
        c_skip = self.sigma_data ** 2 / (st ** 2 + self.sigma_data ** 2)
        c_out = st / (st ** 2 + self.sigma_data ** 2) ** 0.5
        return dict(a_t_sqrt=a_t.sqrt(), b_t_sqrt=(1 - a_t).sqrt(), c_out=c_out, c_skip=c_skip, a_prev_sqrt=a_prev.sqrt(),
                    b_prev_sqrt=(1 - a_prev).sqrt(), last=(i == len(ts) - 1))


def guidance_scale_embedding(w, embedding_dim=256):               # scheduling_lcm.py:87-113
    w = w * 1000.0
    half = embedding_dim // 2
    emb = torch.log(torch.tensor(10000.0)) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -emb)
    emb = w.to(torch.float32)[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)


@torch.no_grad()
def lcm_sample(denoise_fn, cond, shape, steps=2, guidance_scale=5.0, x_T=None, schedule=None, device="cpu"):
    """LCMSampler.lcm_sampling (scheduling_lcm.py:344-382): returns (denoised, last prev_sample).
    ``denoise_fn(x, t_long, cond, w_cond)`` is the DiT.  Random draws follow the reference's order:
    x_T (if not given), then one noise tensor per non-final step, all from torch's global generator on ``device``."""
    sch = schedule or LCMSchedule()
    ts = sch.timesteps(steps)
    b = shape[0]
    img = torch.randn(shape, device=device) if x_T is None else x_T
    w_emb = guidance_scale_embedding(torch.tensor(guidance_scale - 1).repeat(b), 256).to(device=device, dtype=img.dtype)
    denoised = None
    for i, t in enumerate(ts):
        tt = torch.full((b,), t, device=device, dtype=torch.long)
        eps = denoise_fn(img, tt, cond, w_emb)
        k = sch.coefficients(ts, i)
        x0 = (img - k["b_t_sqrt"] * eps) / k["a_t_sqrt"]
        denoised = k["c_out"] * x0 + k["c_skip"] * img
        if not k["last"]:
            noise = torch.randn(eps.shape, device=eps.device)
            img = k["a_prev_sqrt"] * denoised + k["b_prev_sqrt"] * noise
        else:
            img = denoised
    return denoised, img


class PortedDenoiser(object):
    """DiT weights on a device + the 2-step sampler: the reference-PyTorch half of configs[4]."""

    def __init__(self, state_dict=None, device="cpu", cfg=DIT_CFG, seed=0):
        self.cfg, self.device = cfg, torch.device(device)
        self.sd = _t(state_dict if state_dict is not None else dit_state_dict(cfg, seed), self.device)
        self.schedule = LCMSchedule()

    def __call__(self, x, t, cond, w_cond):
        return concat_dit2mlp(self.sd, x, t, cond, w_cond, self.cfg)

    def sample(self, cond, T=312, steps=2, guidance_scale=5.0, x_T=None):
        """cond (B,154,1024) -> latents (B,20,T): what ``LCMSampler.sample(S=2, ...)`` returns first (InferAPI.py:79-86)."""
        b = cond.shape[0]
        return lcm_sample(self, cond, (b, self.cfg["in_channels"], T), steps, guidance_scale, x_T, self.schedule, self.device)[0]
