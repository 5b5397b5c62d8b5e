"""CPU restatement of the reference latent->waveform decode (torch functional ops).

TEST INFRASTRUCTURE (see oracle/__init__.py) - never imported by the product.

The reference's leaf arithmetic is PyTorch's own (F.conv1d, F.conv_transpose1d,
group_norm, softmax, ...), so this port calls the same
ATen CPU kernels on plain tensors taken from a state_dict.  That makes it both
the parity checker and an honest stand-in for the reference's CPU speed
(``cpu_baseline.kind == "port"``).  It is pinned against the unmodified
reference modules by oracle/make_golden.py -> tests/golden/.

Every function cites the reference lines it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _t(x, dtype):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(dtype)


# --------------------------------------------------------------------------- filter
def kaiser_sinc_filter(cutoff: float = 0.25, half_width: float = 0.3, kernel_size: int = 12,
                       dtype=torch.float32) -> torch.Tensor:
    """vocoder/bigvgan/alias_free_torch/filter.py:28-57 (even kernel branch)."""
    half = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    if kernel_size % 2 == 0:
        time = torch.arange(-half, half) + 0.5
    else:
        time = torch.arange(kernel_size) - half
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    filt = filt / filt.sum()
    return filt.to(dtype)


# --------------------------------------------------------------------------- Activation1d
def snake_beta(x, alpha, beta, logscale=True):
    """vocoder/bigvgan/activations.py:107-120 (SnakeBeta); Snake (:46-62) is the beta = alpha case."""
    a = (torch.exp(alpha) if logscale else alpha).view(1, -1, 1)
    b = (torch.exp(beta) if logscale else beta).view(1, -1, 1)
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a) ** 2


def upsample2x(x, filt):
    """alias_free_torch/resample.py:10-33 (ratio 2, k 12: pad 5, pad_left 15, pad_right 15)."""
    C = x.shape[1]
    k, ratio = filt.numel(), 2
    pad = k // ratio - 1
    pad_left = pad * ratio + (k - ratio) // 2
    pad_right = pad * ratio + (k - ratio + 1) // 2
    x = F.pad(x, (pad, pad), mode="replicate")
    x = ratio * F.conv_transpose1d(x, filt.view(1, 1, -1).expand(C, -1, -1), stride=ratio, groups=C)
    return x[..., pad_left:-pad_right]


def downsample2x(x, filt):
    """alias_free_torch/resample.py:36-48 + filter.py:86-95 (pad_left 5, pad_right 6, stride 2)."""
    C = x.shape[1]
    k = filt.numel()
    x = F.pad(x, (k // 2 - 1, k // 2), mode="replicate")
    return F.conv1d(x, filt.view(1, 1, -1).expand(C, -1, -1), stride=2, groups=C)


def activation1d(x, alpha, beta, filt=None, logscale=True):
    """alias_free_torch/act.py:23-28: upsample -> Snake / SnakeBeta -> downsample."""
    if filt is None:
        filt = kaiser_sinc_filter(dtype=x.dtype)
    return downsample2x(snake_beta(upsample2x(x, filt), alpha, beta, logscale), filt)


# --------------------------------------------------------------------------- weight norm
def fold_weight_norm(v, g):
    """torch.nn.utils.weight_norm, dim=0: w = g * v / ||v|| with the norm over all dims
    but 0 (vocoder/bigvgan/models.py:36-51,143,152,174).  For ConvTranspose1d dim 0 is
    the *input* channel."""
    nrm = torch.linalg.vector_norm(v, dim=(1, 2), keepdim=True)
    return v * (g / nrm)


def _wn(sd, name, dtype):
    w = fold_weight_norm(_t(sd[name + ".weight_v"], dtype), _t(sd[name + ".weight_g"], dtype))
    return w, _t(sd[name + ".bias"], dtype)


# --------------------------------------------------------------------------- BigVGAN
def _snake(sd, p, h, dtype):
    """(alpha, beta, logscale) of one activation; ``activation: snake`` has alpha only."""
    a = _t(sd[p + ".alpha"], dtype)
    b = _t(sd[p + ".beta"], dtype) if h["activation"] == "snakebeta" else a
    return a, b, bool(h["snake_logscale"])


def amp_block1(sd, h, p, x, k, dils, filt, dtype):
    """vocoder/bigvgan/models.py:72-81."""
    for l, d in enumerate(dils):
        a1, b1_, ls = _snake(sd, f"{p}.activations.{2 * l}.act", h, dtype)
        a2, b2_, _ = _snake(sd, f"{p}.activations.{2 * l + 1}.act", h, dtype)
        w1, b1 = _wn(sd, f"{p}.convs1.{l}", dtype)
        w2, b2 = _wn(sd, f"{p}.convs2.{l}", dtype)
        xt = activation1d(x, a1, b1_, filt, ls)
        xt = F.conv1d(xt, w1, b1, dilation=d, padding=(k * d - d) // 2)
        xt = activation1d(xt, a2, b2_, filt, ls)
        xt = F.conv1d(xt, w2, b2, dilation=1, padding=(k - 1) // 2)
        x = xt + x
    return x


def amp_block2(sd, h, p, x, k, dils, filt, dtype):
    """vocoder/bigvgan/models.py:119-126."""
    for l, d in enumerate(dils):
        a, b_, ls = _snake(sd, f"{p}.activations.{l}.act", h, dtype)
        w, b = _wn(sd, f"{p}.convs.{l}", dtype)
        xt = activation1d(x, a, b_, filt, ls)
        xt = F.conv1d(xt, w, b, dilation=d, padding=(k * d - d) // 2)
        x = xt + x
    return x


def bigvgan_forward(sd, h, mel, dtype=torch.float32):
    """vocoder/bigvgan/models.py:181-203.  mel (B,num_mels,T) -> (B,1,T*prod(upsample_rates))."""
    if h["activation"] not in ("snake", "snakebeta"):
        raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")
    block = amp_block1 if str(h["resblock"]) == "1" else amp_block2       # models.py:146
    x = _t(mel, dtype)
    filt = kaiser_sinc_filter(dtype=dtype).to(x.device)  # (the GPU tests also run this port as "ATen eager on the same B200")
    w, b = _wn(sd, "conv_pre", dtype)
    x = F.conv1d(x, w, b, padding=3)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        w, b = _wn(sd, f"ups.{i}.0", dtype)
        x = F.conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)
        xs = None
        for j, (kk, dd) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            y = block(sd, h, f"resblocks.{i * nk + j}", x, kk, dd, filt, dtype)
            xs = y if xs is None else xs + y
        x = xs / nk
    a, b_, ls = _snake(sd, "activation_post.act", h, dtype)
    x = activation1d(x, a, b_, filt, ls)
    w, b = _wn(sd, "conv_post", dtype)
    x = F.conv1d(x, w, b, padding=3)
    return torch.tanh(x)


def vocode(sd, h, spec, dtype=torch.float32) -> np.ndarray:
    """VocoderBigVGAN.vocode, vocoder/bigvgan/models.py:406-411 (host ndarray out, squeezed)."""
    with torch.no_grad():
        if isinstance(spec, np.ndarray):
            spec = torch.from_numpy(spec)
            if spec.dim() == 2:
                spec = spec.unsqueeze(0)
        return bigvgan_forward(sd, h, spec, dtype).squeeze().to(torch.float32).cpu().numpy()


# --------------------------------------------------------------------------- VAE decoder
def _gn_swish(sd, name, x, dtype, swish=True):
    """Normalize = GroupNorm(32, C, eps=1e-6, affine) autoencoder1d.py:169-170; swish :172-174."""
    y = F.group_norm(x, 32, _t(sd[name + ".weight"], dtype), _t(sd[name + ".bias"], dtype), eps=1e-6)
    return y * torch.sigmoid(y) if swish else y


def _conv(sd, name, x, dtype, padding):
    return F.conv1d(x, _t(sd[name + ".weight"], dtype), _t(sd[name + ".bias"], dtype), padding=padding)


def resnet_block(sd, p, x, dtype):
    """autoencoder1d.py:215-235 with temb=None, dropout 0."""
    h = _conv(sd, p + ".conv1", _gn_swish(sd, p + ".norm1", x, dtype), dtype, 1)
    h = _conv(sd, p + ".conv2", _gn_swish(sd, p + ".norm2", h, dtype), dtype, 1)
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x, dtype, 0)
    return x + h


def attn_block(sd, p, x, dtype):
    """autoencoder1d.py:257-278.  Note the reference unpacks ``b,t,c = q.shape`` on a
    (b,c,t) tensor, so the scale is C**-0.5."""
    h = _gn_swish(sd, p + ".norm", x, dtype, swish=False)
    q = _conv(sd, p + ".q", h, dtype, 0)
    k = _conv(sd, p + ".k", h, dtype, 0)
    v = _conv(sd, p + ".v", h, dtype, 0)
    C = q.shape[1]
    w = torch.bmm(q.permute(0, 2, 1), k) * (int(C) ** (-0.5))
    w = torch.softmax(w, dim=2)
    h = torch.bmm(v, w.permute(0, 2, 1))
    return x + _conv(sd, p + ".proj_out", h, dtype, 0)


def vae_decode(sd, dd, z, dtype=torch.float32):
    """AutoencoderKL.decode autoencoder1d.py:59-62 + Decoder1D.forward :484-517."""
    nl = len(dd["ch_mult"])
    nrb = dd["num_res_blocks"]
    ks = dd["kernel_size"]
    down_layers = [i + 1 for i in dd["down_layers"]]
    x = _conv(sd, "post_quant_conv", _t(z, dtype), dtype, 0)
    x = _conv(sd, "decoder.conv_in", x, dtype, ks // 2)
    x = resnet_block(sd, "decoder.mid.block_1", x, dtype)
    x = attn_block(sd, "decoder.mid.attn_1", x, dtype)
    x = resnet_block(sd, "decoder.mid.block_2", x, dtype)
    for i_level in reversed(range(nl)):
        for i_block in range(nrb + 1):
            x = resnet_block(sd, f"decoder.up.{i_level}.block.{i_block}", x, dtype)
            if i_level in dd["attn_layers"]:   # autoencoder1d.py:500-504 (never a level index in the shipped config)
                x = attn_block(sd, f"decoder.up.{i_level}.attn.{i_block}", x, dtype)
        if i_level in down_layers:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")  # autoencoder1d.py:291-295
            x = _conv(sd, f"decoder.up.{i_level}.upsample.conv", x, dtype, 1)
    x = _gn_swish(sd, "decoder.norm_out", x, dtype)
    return _conv(sd, "decoder.conv_out", x, dtype, ks // 2)


def vae_encode_moments(sd, dd, x, dtype=torch.float32):
    """AutoencoderKL.encode autoencoder1d.py:52-56 up to the posterior's parameters (mean | logvar):
    Encoder1D.forward :391-413 (ResnetBlocks with k = kernel_size, Downsample1D :296-316 = right zero pad + Conv1d k3
    stride 2) then quant_conv :34."""
    nl, nrb, ks = len(dd["ch_mult"]), dd["num_res_blocks"], dd["kernel_size"]

    def res(p, h):
        y = _conv(sd, p + ".conv1", _gn_swish(sd, p + ".norm1", h, dtype), dtype, ks // 2)
        y = _conv(sd, p + ".conv2", _gn_swish(sd, p + ".norm2", y, dtype), dtype, ks // 2)
        if (p + ".nin_shortcut.weight") in sd:
            h = _conv(sd, p + ".nin_shortcut", h, dtype, 0)
        return h + y

    with torch.no_grad():
        h = _conv(sd, "encoder.conv_in", _t(x, dtype), dtype, ks // 2)
        for i_level in range(nl):
            for i_block in range(nrb):
                h = res(f"encoder.down.{i_level}.block.{i_block}", h)
                if i_level in dd["attn_layers"]:   # autoencoder1d.py:391-396
                    h = attn_block(sd, f"encoder.down.{i_level}.attn.{i_block}", h, dtype)
            if i_level in dd["down_layers"]:
                p = f"encoder.down.{i_level}.downsample.conv"
                h = F.conv1d(F.pad(h, (0, 1)), _t(sd[p + ".weight"], dtype), _t(sd[p + ".bias"], dtype), stride=2)
        h = res("encoder.mid.block_1", h)
        h = attn_block(sd, "encoder.mid.attn_1", h, dtype)
        h = res("encoder.mid.block_2", h)
        h = _conv(sd, "encoder.conv_out", _gn_swish(sd, "encoder.norm_out", h, dtype), dtype, ks // 2)
        return _conv(sd, "quant_conv", h, dtype, 0)


def decode_first_stage(sd, dd, z, scale_factor: float = 1.0, dtype=torch.float32):
    """LCM_audio.decode_first_stage, ldm/models/diffusion/lcm_audio.py:392-406 (KL branch)."""
    with torch.no_grad():
        return vae_decode(sd, dd, (1.0 / scale_factor) * _t(z, dtype), dtype)
