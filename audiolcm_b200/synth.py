"""Deterministic synthetic configs, weights and inputs for the decode path.

Data generator only (numpy; no algorithm of the decode path lives here): used by bench.py, the smoke
test, the tests and the oracle to build identical random-init weights everywhere.

Weights are produced by numpy's PCG64 from a seed, so the *same* tensors can be
(a) loaded into the unmodified reference modules here (oracle/make_golden.py),
(b) fed to this repo's oracle, and (c) fed to the CUDA path on the GPU box,
without shipping weight files.  Key names and shapes follow the reference
state_dicts:

* BigVGAN generator  - /root/reference/vocoder/bigvgan/models.py:133-179
  (``weight_g``/``weight_v``/``bias`` per weight-normed conv, ``alpha``/``beta``
  per SnakeBeta, activations.py:89-103)
* AutoencoderKL decoder half - /root/reference/ldm/models/autoencoder1d.py:30-35,
  :415-482 (``decoder.*`` and ``post_quant_conv.*``)
"""
from __future__ import annotations

import numpy as np


class AttrDict(dict):
    """Same idea as /root/reference/vocoder/bigvgan/env.py:8-11."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


# /root/reference/vocoder/bigvgan/bigvgan_audioset16khz_80band.json (model keys only)
BIGVGAN_16K = dict(
    resblock="1",
    upsample_rates=[4, 4, 2, 2, 2, 2],
    upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
    upsample_initial_channel=1536,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    activation="snakebeta",
    snake_logscale=True,
    num_mels=80,
    sampling_rate=16000,
    hop_size=256,
)

# /root/reference/configs/audiolcm.yaml:54-70 (first_stage_config.params.ddconfig)
VAE_DDCONFIG = dict(
    double_z=True,
    in_channels=80,
    out_ch=80,
    z_channels=20,
    kernel_size=5,
    ch=384,
    ch_mult=[1, 2, 4],
    num_res_blocks=2,
    attn_layers=[3],
    down_layers=[0],
    dropout=0.0,
)
VAE_EMBED_DIM = 20


def bigvgan_config(initial_channel: int = 1536, **over) -> AttrDict:
    h = dict(BIGVGAN_16K)
    h["upsample_initial_channel"] = initial_channel
    h.update(over)
    return AttrDict(h)


def vae_config(ch: int = 384, **over) -> dict:
    c = dict(VAE_DDCONFIG)
    c["ch"] = ch
    c.update(over)
    return c


def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def _wn_conv(rng, sd, name, cout, cin, k, transposed=False):
    """weight-normed conv: Conv1d weight (cout,cin,k), g (cout,1,1);
    ConvTranspose1d weight (cin,cout,k), g (cin,1,1) (norm over dims != 0)."""
    shape = (cin, cout, k) if transposed else (cout, cin, k)
    fan_in = (cout if transposed else cin) * k
    v = _uniform(rng, shape, 1.0 / np.sqrt(fan_in))
    nrm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
    g = (nrm * (1.0 + 0.1 * rng.standard_normal(nrm.shape))).astype(np.float32)
    sd[name + ".weight_g"] = g
    sd[name + ".weight_v"] = v
    sd[name + ".bias"] = _uniform(rng, (cout,), 1.0 / np.sqrt(fan_in))


def _snake_params(rng, sd, p, ch, h):
    """alpha (and beta for snakebeta): N(0, 0.5) in log scale; U(0.6, 1.6) in linear scale, where the reference initialises
    them to 1 (activations.py:33-42,88-98) and 1/beta must stay bounded for a well-conditioned test network."""
    draw = (lambda: 0.5 * rng.standard_normal(ch)) if h["snake_logscale"] else (lambda: rng.uniform(0.6, 1.6, size=ch))
    sd[p + ".alpha"] = draw().astype(np.float32)
    if h["activation"] == "snakebeta":
        sd[p + ".beta"] = draw().astype(np.float32)


def bigvgan_state_dict(h, seed: int = 0) -> dict:
    """Synthetic generator state_dict (numpy fp32), reference key names.
    The 218 constant ``filter`` buffers are omitted (load with strict=False)."""
    rng = np.random.default_rng(seed)
    sd = {}
    c0 = h["upsample_initial_channel"]
    _wn_conv(rng, sd, "conv_pre", c0, h["num_mels"], 7)
    nk = len(h["resblock_kernel_sizes"])
    ch = c0
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cin, ch = c0 // (2 ** i), c0 // (2 ** (i + 1))
        _wn_conv(rng, sd, f"ups.{i}.0", ch, cin, k, transposed=True)
        for j, kk in enumerate(h["resblock_kernel_sizes"]):
            p = f"resblocks.{i * nk + j}"
            if str(h["resblock"]) == "1":
                for l in range(3):
                    _wn_conv(rng, sd, f"{p}.convs1.{l}", ch, ch, kk)
                for l in range(3):
                    _wn_conv(rng, sd, f"{p}.convs2.{l}", ch, ch, kk)
                nact = 6
            else:                                       # AMPBlock2: two convs, two activations
                for l in range(2):
                    _wn_conv(rng, sd, f"{p}.convs.{l}", ch, ch, kk)
                nact = 2
            for m in range(nact):
                _snake_params(rng, sd, f"{p}.activations.{m}.act", ch, h)
    _snake_params(rng, sd, "activation_post.act", ch, h)
    _wn_conv(rng, sd, "conv_post", 1, ch, 7)
    return sd


def _conv(rng, sd, name, cout, cin, k):
    b = 1.0 / np.sqrt(cin * k)
    sd[name + ".weight"] = _uniform(rng, (cout, cin, k), b)
    sd[name + ".bias"] = _uniform(rng, (cout,), b)


def _gn(rng, sd, name, c):
    sd[name + ".weight"] = (1.0 + 0.2 * rng.standard_normal(c)).astype(np.float32)
    sd[name + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)


def _resblock(rng, sd, p, cin, cout):
    _gn(rng, sd, p + ".norm1", cin)
    _conv(rng, sd, p + ".conv1", cout, cin, 3)  # decoder ResnetBlocks are k=3 (autoencoder1d.py:444-464)
    _gn(rng, sd, p + ".norm2", cout)
    _conv(rng, sd, p + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(rng, sd, p + ".nin_shortcut", cout, cin, 1)


def _attn(rng, sd, p, c):
    _gn(rng, sd, p + ".norm", c)
    for n in ("q", "k", "v", "proj_out"):
        _conv(rng, sd, f"{p}.{n}", c, c, 1)


def vae_decoder_state_dict(dd, embed_dim: int = VAE_EMBED_DIM, seed: int = 0) -> dict:
    """Synthetic state_dict for ``post_quant_conv`` + ``decoder`` (numpy fp32)."""
    rng = np.random.default_rng(seed + 1000)
    sd = {}
    ch, mult, nrb = dd["ch"], list(dd["ch_mult"]), dd["num_res_blocks"]
    nl = len(mult)
    ks = dd["kernel_size"]
    zc = dd["z_channels"]
    _conv(rng, sd, "post_quant_conv", zc, embed_dim, 1)
    block_in = ch * mult[nl - 1]
    _conv(rng, sd, "decoder.conv_in", block_in, zc, ks)
    _resblock(rng, sd, "decoder.mid.block_1", block_in, block_in)
    _gn(rng, sd, "decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        _conv(rng, sd, f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    _resblock(rng, sd, "decoder.mid.block_2", block_in, block_in)
    down_layers = [i + 1 for i in dd["down_layers"]]
    for i_level in reversed(range(nl)):
        block_out = ch * mult[i_level]
        for i_block in range(nrb + 1):
            _resblock(rng, sd, f"decoder.up.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
            if i_level in dd["attn_layers"]:
                _attn(rng, sd, f"decoder.up.{i_level}.attn.{i_block}", block_in)
        if i_level in down_layers:
            _conv(rng, sd, f"decoder.up.{i_level}.upsample.conv", block_in, block_in, 3)
    _gn(rng, sd, "decoder.norm_out", block_in)
    _conv(rng, sd, "decoder.conv_out", dd["out_ch"], block_in, ks)
    return sd


def _resblock_k(rng, sd, p, cin, cout, k):
    _gn(rng, sd, p + ".norm1", cin)
    _conv(rng, sd, p + ".conv1", cout, cin, k)
    _gn(rng, sd, p + ".norm2", cout)
    _conv(rng, sd, p + ".conv2", cout, cout, k)
    if cin != cout:
        _conv(rng, sd, p + ".nin_shortcut", cout, cin, 1)


def vae_encoder_state_dict(dd, embed_dim: int = VAE_EMBED_DIM, seed: int = 0) -> dict:
    """Synthetic state_dict for ``encoder`` + ``quant_conv`` (numpy fp32), reference key names
    (ldm/models/autoencoder1d.py:319-413,34): encoder ResnetBlocks DO get kernel_size (k = 5), unlike the decoder's."""
    rng = np.random.default_rng(seed + 2000)
    sd = {}
    ch, mult, nrb, ks = dd["ch"], list(dd["ch_mult"]), dd["num_res_blocks"], dd["kernel_size"]
    _conv(rng, sd, "encoder.conv_in", ch, dd["in_channels"], ks)
    block_in = ch
    for i_level in range(len(mult)):
        block_out = ch * mult[i_level]
        for i_block in range(nrb):
            _resblock_k(rng, sd, f"encoder.down.{i_level}.block.{i_block}", block_in, block_out, ks)
            block_in = block_out
            if i_level in dd["attn_layers"]:
                _attn(rng, sd, f"encoder.down.{i_level}.attn.{i_block}", block_in)
        if i_level in dd["down_layers"]:
            _conv(rng, sd, f"encoder.down.{i_level}.downsample.conv", block_in, block_in, 3)
    _resblock_k(rng, sd, "encoder.mid.block_1", block_in, block_in, ks)
    _gn(rng, sd, "encoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        _conv(rng, sd, f"encoder.mid.attn_1.{n}", block_in, block_in, 1)
    _resblock_k(rng, sd, "encoder.mid.block_2", block_in, block_in, ks)
    _gn(rng, sd, "encoder.norm_out", block_in)
    zc2 = 2 * dd["z_channels"] if dd.get("double_z", True) else dd["z_channels"]
    _conv(rng, sd, "encoder.conv_out", zc2, block_in, ks)
    _conv(rng, sd, "quant_conv", 2 * embed_dim, zc2, 1)
    return sd


def synth_mel(B: int, T: int, seed: int = 0, n_mels: int = 80) -> np.ndarray:
    """log10-mel shaped input: clamp(N(-2.5, 1.5^2), -5, 1.5)."""
    rng = np.random.default_rng(seed + 7)
    x = -2.5 + 1.5 * rng.standard_normal((B, n_mels, T))
    return np.clip(x, -5.0, 1.5).astype(np.float32)


def synth_latent(B: int, T: int, seed: int = 0, zc: int = 20) -> np.ndarray:
    rng = np.random.default_rng(seed + 11)
    return rng.standard_normal((B, zc, T)).astype(np.float32)
