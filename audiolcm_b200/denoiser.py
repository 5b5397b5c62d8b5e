"""The denoiser side of AudioLCM inference with its heavy part on the sm_100a conv kernel.

``ConcatDiT2MLP`` (/root/reference/ldm/modules/diffusionmodules/concatDiT.py:238-304) is a 4-block transformer over
1 + 154 + T tokens whose feed-forward is a pair of 9-tap ``Conv1d`` layers (``Conv1dFeedForward`` with GEGLU,
/root/reference/ldm/modules/new_attention.py:48-74): 576 -> 4608 and 2304 -> 576 channels.  Those two convs are 93 % of
the denoiser's FLOPs and have exactly the shape ``conv_umma_kernel`` handles, so here they run on tcgen05 through
persistent handles (weights packed once, one plan per (B,T): ``alcm_ffn1d`` = conv, GEGLU, conv + residual), and so do the q/k/v and output
projections of the two self-attentions of every block (Linear over tokens = 1x1 conv); the sampler's ``step()``
(scheduling_lcm.py:411-494) is the fused ``alcm_lcm_step`` kernel.  Everything else of the DiT (timestep / condition
embedders, the softmax(QK^T)V core of the 8-head self-attentions over 467 tokens, GroupNorm) is small
and stays in PyTorch: this module is the HYBRID the scope table calls "next", not a from-scratch denoiser.

Pinned to the unmodified reference classes through tests/golden/lcm_denoiser.npz (made with the real LCM_audio,
oracle/make_golden_lcm.py); the pure-PyTorch restatement in baseline/lcm_denoiser_port.py is its test oracle.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

DIT_CFG = dict(in_channels=20, context_dim=1024, hidden_size=576, num_heads=8, depth=4, max_len=1000)   # configs/audiolcm.yaml:39-47


class Conv1dLayer(object):
    """``nn.Conv1d(Cin, Cout, K, dilation=d, padding=(K*d-d)//2)`` on ``conv_umma_kernel``: y = conv(x) + bias (+ res)."""

    def __init__(self, weight, bias=None, dilation=1, device="cuda", precision="bf16"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.AlcmError("audiolcm_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        w = (torch.from_numpy(weight) if isinstance(weight, np.ndarray) else weight).detach().to(dev, torch.float32).contiguous()
        b = None if bias is None else (torch.from_numpy(bias) if isinstance(bias, np.ndarray) else bias).detach().to(dev, torch.float32).contiguous()
        self.device, self.cout, self.cin, self.k = dev, int(w.shape[0]), int(w.shape[1]), int(w.shape[2])
        h = C.c_void_p()
        with torch.cuda.device(dev):
            torch.cuda.synchronize()
            _lib.check(_lib.load().alcm_conv1d_create(_lib.ctx(dev.index), w.data_ptr(), None if b is None else b.data_ptr(), self.cout, self.cin,
                                                      self.k, int(dilation), _lib.PREC[precision], C.byref(h)))
        self._h = h.value

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().alcm_conv1d_destroy(h)
            except Exception:
                pass
            self._h = None

    def __call__(self, x, res=None):
        x = x.to(dtype=torch.float32, device=self.device).contiguous()
        if x.dim() != 3 or x.shape[1] != self.cin:
            raise ValueError(f"expected (B,{self.cin},T), got {tuple(x.shape)}")
        B, _, T = x.shape
        if res is not None:
            res = res.to(dtype=torch.float32, device=self.device).contiguous()
        y = torch.empty((B, self.cout, T), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().alcm_conv1d_run(self._h, x.data_ptr(), None if res is None else res.data_ptr(), y.data_ptr(), B, T,
                                                   torch.cuda.current_stream().cuda_stream))
        return y


class Conv1dFeedForwardLayer(object):
    """``Conv1dFeedForward(dim, mult=4, glu=True)`` (new_attention.py:48-74) as one native plan: conv -> GEGLU -> conv (+ res)."""

    def __init__(self, w_in, b_in, w_out, b_out, device="cuda", precision="bf16"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.AlcmError("audiolcm_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        t = lambda v: None if v is None else (torch.from_numpy(v) if isinstance(v, np.ndarray) else v).detach().to(dev, torch.float32).contiguous()
        w_in, b_in, w_out, b_out = t(w_in), t(b_in), t(w_out), t(b_out)
        self.device, self.dim, self.inner, self.dim_out, self.k = dev, int(w_in.shape[1]), int(w_out.shape[1]), int(w_out.shape[0]), int(w_in.shape[2])
        if w_in.shape[0] != 2 * self.inner or w_out.shape[2] != self.k:
            raise ValueError("Conv1dFeedForward weights: expected [2*inner,dim,K] and [dim_out,inner,K]")
        h = C.c_void_p()
        ptr = lambda v: None if v is None else v.data_ptr()
        with torch.cuda.device(dev):
            torch.cuda.synchronize()
            _lib.check(_lib.load().alcm_ffn1d_create(_lib.ctx(dev.index), ptr(w_in), ptr(b_in), ptr(w_out), ptr(b_out), self.dim, self.inner,
                                                     self.dim_out, self.k, _lib.PREC[precision], C.byref(h)))
        self._h = h.value

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().alcm_ffn1d_destroy(h)
            except Exception:
                pass
            self._h = None

    def __call__(self, x, res=None):
        x = x.to(dtype=torch.float32, device=self.device).contiguous()
        if x.dim() != 3 or x.shape[1] != self.dim:
            raise ValueError(f"expected (B,{self.dim},T), got {tuple(x.shape)}")
        B, _, T = x.shape
        if res is not None:
            res = res.to(dtype=torch.float32, device=self.device).contiguous()
        y = torch.empty((B, self.dim_out, T), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().alcm_ffn1d_run(self._h, x.data_ptr(), None if res is None else res.data_ptr(), y.data_ptr(), B, T,
                                                  torch.cuda.current_stream().cuda_stream))
        return y


class ConcatDiT2MLPB200(object):
    """``ConcatDiT2MLP.forward(x, t, context, w_cond)`` with the feed-forward convs on the B200 conv kernel.

    ``state_dict``: the reference module's (``model.unet.diffusion_model.state_dict()``), numpy or torch values."""

    def __init__(self, state_dict, device="cuda", precision="bf16", cfg=DIT_CFG):
        self.cfg, self.device, self.precision = dict(cfg), torch.device(device), precision
        self.sd = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v).detach().to(self.device, torch.float32) for k, v in state_dict.items()}
        self.ff, self.qkv, self.attn_out, self.blk_in, self.blk_out = [], [], [], [], []
        for i in range(cfg["depth"]):
            tb = f"blocks.{i}.transformer_blocks.0"
            self.ff.append(Conv1dFeedForwardLayer(self.sd[f"{tb}.ff.net.0.proj.weight"], self.sd[f"{tb}.ff.net.0.proj.bias"],
                                                  self.sd[f"{tb}.ff.net.2.weight"], self.sd[f"{tb}.ff.net.2.bias"], device, precision))
            self.blk_in.append(Conv1dLayer(self.sd[f"blocks.{i}.proj_in.weight"], self.sd[f"blocks.{i}.proj_in.bias"], 1, device, precision))
            self.blk_out.append(Conv1dLayer(self.sd[f"blocks.{i}.proj_out.weight"], self.sd[f"blocks.{i}.proj_out.bias"], 1, device, precision))
            for a in ("attn1", "attn2"):   # to_q | to_k | to_v as ONE 1x1 conv (576 -> 1728), to_out as another: Linear over tokens = 1x1 conv
                wqkv = torch.cat([self.sd[f"{tb}.{a}.to_{n}.weight"] for n in "qkv"], dim=0).unsqueeze(-1).contiguous()
                self.qkv.append(Conv1dLayer(wqkv, None, 1, device, precision))
                self.attn_out.append(Conv1dLayer(self.sd[f"{tb}.{a}.to_out.0.weight"].unsqueeze(-1).contiguous(), self.sd[f"{tb}.{a}.to_out.0.bias"],
                                                 1, device, precision))

    @staticmethod
    def _timestep_embedding(t, dim=256, max_period=10000):            # concatDiT.py:49-69
        half = dim // 2
        freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
        args = t[:, None].float() * freqs[None]
        return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)

    def _attn(self, idx, xn_cf, res_cf):
        """new_attention.py:107-130 (self-attention): xn_cf (B,C,N) normalised tokens and res_cf (B,C,N) residual stream, both
        channels-first.  Returns to_out(attention) + residual; both projections run on conv_umma_kernel."""
        heads = self.cfg["num_heads"]
        B, Cc, N = xn_cf.shape
        d = Cc // heads
        qkv = self.qkv[idx](xn_cf)                                                     # (B, 3C, N)
        # one transposing copy makes q|k|v (B, heads, N, d) with unit stride in d: PyTorch's fused attention kernels
        # need that (strided heads fall back to its 3-kernel math path); bf16 mode takes the flash kernel
        dt = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(self.precision, torch.float32)
        q, k, v = qkv.reshape(B, 3, heads, d, N).permute(1, 0, 2, 4, 3).to(dt, memory_format=torch.contiguous_format)
        out = F.scaled_dot_product_attention(q, k, v, scale=d ** -0.5)                 # (B, heads, N, d)
        out_cf = out.transpose(2, 3).reshape(B, Cc, N).float()                         # channel = head*d + dd, as 'b n (h d)'
        return self.attn_out[idx](out_cf, res=res_cf)                                  # + bias + residual in the epilogue

    def _cond(self, p, c):                                              # concatDiT.py:93-104
        sd = self.sd
        h = F.gelu(F.linear(c, sd[f"{p}.mlp.0.weight"], sd[f"{p}.mlp.0.bias"]), approximate="tanh")
        h = F.linear(h, sd[f"{p}.mlp.2.weight"], sd[f"{p}.mlp.2.bias"])
        return F.layer_norm(h, h.shape[-1:], sd[f"{p}.mlp.3.weight"], sd[f"{p}.mlp.3.bias"])

    @torch.no_grad()
    def __call__(self, x, t, context, w_cond=None):
        from . import ops
        sd = self.sd
        t_freq = self._timestep_embedding(t, 256)
        if w_cond is not None:
            t_freq = t_freq + F.linear(w_cond, sd["t_embedder.proj_w.weight"])
        temb = F.linear(F.silu(F.linear(t_freq, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                        sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"]).unsqueeze(1)
        c1, c2 = context.chunk(2, dim=1)
        c = torch.cat((self._cond("c1_embedder", c1), self._cond("c2_embedder", c2)), dim=1)
        extra = c.shape[1] + 1
        h = F.conv1d(x, sd["proj_in.weight"], sd["proj_in.bias"], padding=2).permute(0, 2, 1)
        h = torch.cat([temb, c, h], dim=1)
        h = h + sd["pos_emb.weight"][: h.shape[1]].unsqueeze(0)
        h = h.permute(0, 2, 1).contiguous()                              # (N, H, extra+T)
        for i in range(self.cfg["depth"]):
            p, tb = f"blocks.{i}", f"blocks.{i}.transformer_blocks.0"
            x_in = h
            y = F.group_norm(h, 32, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], eps=1e-6)
            yc = self.blk_in[i](y)                                         # 1x1 proj_in; residual stream, channels-first from here on
            ln = lambda v_cf, n: ops.layernorm_cf(v_cf, sd[f"{tb}.{n}.weight"], sd[f"{tb}.{n}.bias"])   # LayerNorm over channels, layout kept
            yc = self._attn(2 * i, ln(yc, "norm1"), yc)
            yc = self._attn(2 * i + 1, ln(yc, "norm2"), yc)
            y = self.ff[i](ln(yc, "norm3"), res=yc)       # Conv1d 576 -> 4608 k9, GEGLU, Conv1d 2304 -> 576 k9, + residual: one plan
            h = self.blk_out[i](y, res=x_in)                              # 1x1 proj_out + the block's skip connection
        h = h[..., extra:]
        h = F.group_norm(h, 16, sd["final_layer.norm_final.weight"], sd["final_layer.norm_final.bias"])
        return F.conv1d(h, sd["final_layer.conv1d.weight"], sd["final_layer.conv1d.bias"])


class LCMSamplerB200(object):
    """``LCMSampler.sample(S=2, ...)`` (scheduling_lcm.py:298-382): the hybrid DiT + the fused step kernel.

    Scalars of the schedule follow scheduling_lcm.py:118-259,401-452 and DDPM.register_schedule (ddpm.py:116-137)."""

    def __init__(self, denoiser, timesteps=1000, linear_start=0.00085, linear_end=0.012, original_inference_steps=50):
        self.denoiser, self.device = denoiser, denoiser.device
        betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2
        self.alphas_cumprod = np.cumprod(1.0 - betas).astype(np.float32)
        self.num_train, self.original_inference_steps = timesteps, original_inference_steps
        self.timestep_scaling, self.sigma_data = 10.0, 0.5

    def timesteps(self, steps):
        k = self.num_train // self.original_inference_steps
        origin = (np.arange(1, self.original_inference_steps + 1) * k - 1)[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=steps, endpoint=False)).astype(np.int64)
        return [int(v) for v in origin[idx]]

    @staticmethod
    def guidance_embedding(w, dim=256):                                 # scheduling_lcm.py:87-113
        w = w * 1000.0
        half = dim // 2
        emb = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000.0) / (half - 1)))
        emb = w.to(torch.float32)[:, None] * emb[None, :]
        return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)

    @torch.no_grad()
    def sample(self, cond, T=312, steps=2, guidance_scale=5.0, x_T=None):
        from . import ops
        b = cond.shape[0]
        shape = (b, self.denoiser.cfg["in_channels"], T)
        img = torch.randn(shape, device=self.device) if x_T is None else x_T.to(self.device)
        w_emb = self.guidance_embedding(torch.tensor(guidance_scale - 1).repeat(b)).to(self.device)
        ts = self.timesteps(steps)
        denoised = None
        for i, t in enumerate(ts):
            eps = self.denoiser(img, torch.full((b,), t, device=self.device, dtype=torch.long), cond, w_emb)
            prev_t = ts[i + 1] if i + 1 < len(ts) else t
            a_t, a_prev = float(self.alphas_cumprod[t]), float(self.alphas_cumprod[prev_t])
            st = t * self.timestep_scaling
            c_skip = self.sigma_data ** 2 / (st ** 2 + self.sigma_data ** 2)
            c_out = st / (st ** 2 + self.sigma_data ** 2) ** 0.5
            last = i == len(ts) - 1
            noise = None if last else torch.randn(shape, device=self.device)
            img, denoised = ops.lcm_step(img, eps, noise, math.sqrt(a_t), math.sqrt(1 - a_t), c_out, c_skip, math.sqrt(a_prev),
                                         math.sqrt(1 - a_prev), last)
        return denoised
