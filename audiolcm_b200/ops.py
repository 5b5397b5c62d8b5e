"""Single-kernel entry points of the C-ABI on plain (B,C,T) float32 CUDA tensors.

These exist for parity tests and micro-benchmarks; the product path is ``VocoderBigVGAN`` /
``AutoencoderKLDecoder``.  Each cites the reference op it restates in include/audiolcm_b200.h.
"""
from __future__ import annotations

import torch

from . import _lib


def _prep(*ts):
    dev = ts[0].device
    if dev.type != "cuda":
        raise _lib.AlcmError("audiolcm_b200 ops need CUDA tensors; there is no CPU path")
    return [None if t is None else t.to(device=dev, dtype=torch.float32).contiguous() for t in ts], dev


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def activation1d(x, alpha, beta, precision="fp32"):
    """Activation1d(SnakeBeta, logscale) - alias_free_torch/act.py:23-28."""
    (x, alpha, beta), dev = _prep(x, alpha, beta)
    B, C, T = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().alcm_activation1d_fwd(_lib.ctx(dev.index), _p(x), _p(alpha), _p(beta), _p(y), B, C, T,
                                                     _lib.PREC[precision], _stream()))
    return y


def conv1d(x, w, bias=None, res=None, dilation=1, precision="fp32"):
    """Conv1d with 'same' zero padding (K*d-d)/2 (+bias, +residual) - vocoder/bigvgan/models.py:36-53."""
    (x, w, bias, res), dev = _prep(x, w, bias, res)
    B, Cin, T = x.shape
    Cout, Cin2, K = w.shape
    assert Cin2 == Cin
    y = torch.empty((B, Cout, T), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().alcm_conv1d_fwd(_lib.ctx(dev.index), _p(x), _p(w), _p(bias), _p(res), _p(y), B, Cin, Cout, T, K,
                                               int(dilation), _lib.PREC[precision], _stream()))
    return y


def conv_transpose1d(x, w, bias=None, stride=2, precision="fp32"):
    """ConvTranspose1d(k=2*stride, stride, padding=stride/2) - vocoder/bigvgan/models.py:150-155."""
    (x, w, bias), dev = _prep(x, w, bias)
    B, Cin, T = x.shape
    Cin2, Cout, K = w.shape
    assert Cin2 == Cin and K == 2 * stride
    y = torch.empty((B, Cout, T * stride), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().alcm_conv_transpose1d_fwd(_lib.ctx(dev.index), _p(x), _p(w), _p(bias), _p(y), B, Cin, Cout, T,
                                                         int(stride), _lib.PREC[precision], _stream()))
    return y


def upsample_conv3(x, w, bias=None, precision="fp32"):
    """nearest x2 then Conv1d(k=3,p=1) - ldm/models/autoencoder1d.py:291-295."""
    (x, w, bias), dev = _prep(x, w, bias)
    B, Cin, T = x.shape
    Cout = w.shape[0]
    y = torch.empty((B, Cout, 2 * T), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().alcm_upsample_conv3_fwd(_lib.ctx(dev.index), _p(x), _p(w), _p(bias), _p(y), B, Cin, Cout, T,
                                                       _lib.PREC[precision], _stream()))
    return y


def groupnorm_swish(x, gamma, beta, swish=True, groups=32, eps=1e-6):
    """GroupNorm(32, eps=1e-6) [+ x*sigmoid(x)] - ldm/models/autoencoder1d.py:169-174."""
    (x, gamma, beta), dev = _prep(x, gamma, beta)
    B, C, T = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().alcm_groupnorm_swish_fwd(_lib.ctx(dev.index), _p(x), _p(gamma), _p(beta), _p(y), B, C, T,
                                                        int(groups), float(eps), int(bool(swish)), _stream()))
    return y


def attn1d(q, k, v, precision="fp32"):
    """softmax_j(q^T k C^-0.5) applied to v - ldm/models/autoencoder1d.py:264-275."""
    (q, k, v), dev = _prep(q, k, v)
    B, C, T = q.shape
    out = torch.empty_like(q)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().alcm_attn1d_fwd(_lib.ctx(dev.index), _p(q), _p(k), _p(v), _p(out), B, C, T, _lib.PREC[precision], _stream()))
    return out


def lcm_step(sample, eps, noise, sqrt_alpha_prod_t, sqrt_beta_prod_t, c_out, c_skip, sqrt_alpha_prod_prev, sqrt_beta_prod_prev, last_step):
    """LCMSampler.step (scheduling_lcm.py:411-494, epsilon prediction) as one kernel -> (prev_sample, denoised)."""
    (sample, eps, noise), dev = _prep(sample, eps, noise)
    n = sample.numel()
    pad = (-n) % 4
    if pad:      # keep the float4 kernel: work on padded flat copies
        f = lambda t: None if t is None else torch.nn.functional.pad(t.reshape(-1), (0, pad))
        p, d = lcm_step(f(sample), f(eps), f(noise), sqrt_alpha_prod_t, sqrt_beta_prod_t, c_out, c_skip, sqrt_alpha_prod_prev,
                        sqrt_beta_prod_prev, last_step)
        return p[:n].reshape(sample.shape), d[:n].reshape(sample.shape)
    prev, den = torch.empty_like(sample), torch.empty_like(sample)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().alcm_lcm_step(_lib.ctx(dev.index), _p(sample), _p(eps), _p(noise), _p(prev), _p(den), n,
                                             float(sqrt_alpha_prod_t), float(sqrt_beta_prod_t), float(c_out), float(c_skip),
                                             float(sqrt_alpha_prod_prev), float(sqrt_beta_prod_prev), int(bool(last_step)), _stream()))
    return prev, den


def layernorm_cf(x, gamma, beta, eps=1e-5):
    """``nn.LayerNorm(C)`` over the channel axis of a channels-first (B,C,T) tensor (the DiT's norm1/2/3, new_attention.py:246-248)."""
    (x, gamma, beta), dev = _prep(x, gamma, beta)
    B, Cc, T = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().alcm_layernorm_cf(_lib.ctx(dev.index), _p(x), _p(gamma), _p(beta), _p(y), B, Cc, T, float(eps), _stream()))
    return y
