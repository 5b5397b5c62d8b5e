"""Pins the CPU oracle (oracle/decode_oracle.py, oracle/np_closed_form.py) to outputs of the
UNMODIFIED reference modules committed under tests/golden/ (made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle as O
from oracle import np_closed_form as NP
from audiolcm_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_filter_taps_match_reference_buffer(golden_dir):
    g = _load(golden_dir, "activation1d.npz")
    f = O.kaiser_sinc_filter().numpy()
    assert np.array_equal(f, g["filter"])                      # same torch ops -> bit-equal
    np.testing.assert_allclose(NP.kaiser_sinc_12(), g["filter"].astype(np.float64), atol=5e-8)
    # SURVEY.md row a8 constants
    np.testing.assert_allclose(g["filter"][:3], [0.0020289647, 0.0093894657, -0.0255434588], atol=1e-9)


@pytest.mark.parametrize("case", ["a", "b", "c", "d", "e"])
def test_activation1d_matches_reference(golden_dir, case):
    g = _load(golden_dir, "activation1d.npz")
    x, al, be = g[f"{case}_x"], g[f"{case}_alpha"], g[f"{case}_beta"]
    y = O.activation1d(torch.from_numpy(x), torch.from_numpy(al), torch.from_numpy(be)).numpy()
    assert np.array_equal(y, g[f"{case}_y"])                   # identical op sequence
    # closed form (what the CUDA kernel computes) in float64 vs reference in float64
    y64 = NP.activation1d_closed_form(x.astype(np.float64), al.astype(np.float64), be.astype(np.float64))
    np.testing.assert_allclose(y64, g[f"{case}_y64"], atol=1e-6)  # filter taps are fp32 in the reference
    np.testing.assert_allclose(y64, g[f"{case}_y"], atol=2e-5)


@pytest.mark.parametrize("tag", ["c64", "c256"])
def test_bigvgan_small_matches_reference(golden_dir, tag):
    g = _load(golden_dir, f"bigvgan_{tag}.npz")
    h = synth.bigvgan_config(int(g["c0"]))
    sd = synth.bigvgan_state_dict(h, seed=int(g["wseed"]))
    mel = synth.synth_mel(int(g["B"]), int(g["T"]), seed=int(g["xseed"]))
    with torch.no_grad():
        wav = O.bigvgan_forward(sd, h, mel).numpy()
    assert wav.shape == g["wav"].shape
    np.testing.assert_allclose(wav, g["wav"], atol=2e-6)
    # squeeze / host-ndarray contract of vocode()
    out = O.vocode(sd, h, mel)
    assert out.dtype == np.float32 and out.shape == tuple(s for s in g["wav"].shape if s != 1)


@pytest.mark.parametrize("tag", ["rb2_snake", "rb1_linear", "rb2_snakebeta_linear"])
def test_bigvgan_config_variants_match_reference(golden_dir, tag):
    """AMPBlock2, Snake and linear-scale parameters (models.py:90-126,146,158-172; activations.py:9-62) against the
    unmodified reference BigVGAN (oracle/make_golden.py VARIANTS)."""
    from oracle.make_golden import VARIANTS
    g = _load(golden_dir, f"bigvgan_c64_{tag}.npz")
    h = synth.bigvgan_config(int(g["c0"]), **VARIANTS[tag])
    sd = synth.bigvgan_state_dict(h, seed=int(g["wseed"]))
    mel = synth.synth_mel(int(g["B"]), int(g["T"]), seed=int(g["xseed"]))
    with torch.no_grad():
        wav = O.bigvgan_forward(sd, h, mel).numpy()
    np.testing.assert_allclose(wav, g["wav"], atol=2e-6)


def test_bigvgan_full_T40_matches_reference(golden_dir):
    g = _load(golden_dir, "bigvgan_full_T40.npz")
    h = synth.bigvgan_config()
    sd = synth.bigvgan_state_dict(h, seed=int(g["wseed"]))
    mel = synth.synth_mel(1, int(g["T"]), seed=int(g["xseed"]))
    with torch.no_grad():
        wav = O.bigvgan_forward(sd, h, mel).numpy()
    np.testing.assert_allclose(wav, g["wav"], atol=2e-6)


@pytest.mark.slow
def test_bigvgan_full_10s_matches_reference(golden_dir):
    g = _load(golden_dir, "bigvgan_full_T625.npz")
    h = synth.bigvgan_config()
    sd = synth.bigvgan_state_dict(h, seed=0)
    mel = synth.synth_mel(1, 625, seed=0)
    with torch.no_grad():
        wav = O.bigvgan_forward(sd, h, mel).numpy()
    assert wav.shape == (1, 1, 160000)
    np.testing.assert_allclose(wav, g["wav"], atol=2e-6)


@pytest.mark.parametrize("tag", ["ch32", "full_T17", "full"])
def test_vae_decode_matches_reference(golden_dir, tag):
    g = _load(golden_dir, f"vae_{tag}.npz")
    dd = synth.vae_config(int(g["ch"]))
    sd = synth.vae_decoder_state_dict(dd, seed=int(g["wseed"]))
    z = synth.synth_latent(int(g["B"]), int(g["T"]), seed=int(g["xseed"]))
    mel = O.decode_first_stage(sd, dd, z).numpy()
    assert mel.shape == (int(g["B"]), 80, 2 * int(g["T"]))
    np.testing.assert_allclose(mel, g["mel"], atol=2e-5)


def test_full_path_matches_reference(golden_dir):
    g = _load(golden_dir, "path_full_T24.npz")
    dd = synth.vae_config()
    h = synth.bigvgan_config()
    z = synth.synth_latent(1, 24, seed=5)
    mel = O.decode_first_stage(synth.vae_decoder_state_dict(dd, seed=3), dd, z)
    np.testing.assert_allclose(mel.numpy(), g["mel"], atol=2e-5)
    with torch.no_grad():
        wav = O.bigvgan_forward(synth.bigvgan_state_dict(h, seed=0), h, mel).numpy()
    np.testing.assert_allclose(wav, g["wav"], atol=5e-6)


def test_polyphase_closed_forms():
    rng = np.random.default_rng(0)
    import torch.nn.functional as F
    for u in (2, 4):
        x = rng.standard_normal((2, 6, 11))
        w = rng.standard_normal((6, 4, 2 * u))
        b = rng.standard_normal(4)
        ref = F.conv_transpose1d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride=u, padding=u // 2).numpy()
        np.testing.assert_allclose(NP.conv_transpose1d_polyphase(x, w, b, u), ref, atol=1e-12)
    x = rng.standard_normal((2, 5, 9))
    w = rng.standard_normal((7, 5, 3))
    b = rng.standard_normal(7)
    ref = F.conv1d(F.interpolate(torch.from_numpy(x), scale_factor=2.0, mode="nearest"), torch.from_numpy(w), torch.from_numpy(b), padding=1).numpy()
    np.testing.assert_allclose(NP.nearest2x_conv3_polyphase(x, w, b), ref, atol=1e-12)


@pytest.mark.parametrize("tag", ["ch32", "full_T64"])
def test_oracle_vae_encode_matches_reference(golden_dir, tag):
    """oracle.vae_encode_moments == AutoencoderKL.encode(x).parameters of the unmodified reference."""
    import torch
    from audiolcm_b200 import synth
    from oracle import decode_oracle as O
    g = np.load(os.path.join(golden_dir, f"vae_enc_{tag}.npz"))
    dd = synth.vae_config(int(g["ch"]))
    sd = {k: torch.from_numpy(v) for k, v in synth.vae_encoder_state_dict(dd, seed=int(g["wseed"])).items()}
    x = synth.synth_mel(int(g["B"]), int(g["T"]), seed=int(g["xseed"]))
    mom = O.vae_encode_moments(sd, dd, x).numpy()
    assert mom.shape == g["moments"].shape
    assert np.abs(mom - g["moments"]).max() <= 2e-6


def test_vae_level_attention_matches_reference(golden_dir):
    """attn_layers holding a level index: an AttnBlock1D after every ResnetBlock1D of that level, in Decoder1D and Encoder1D
    (autoencoder1d.py:356-358,391-396,466-468,500-504) - accepted by the constructors, unused by the shipped config."""
    g = _load(golden_dir, "vae_ch32_level_attn.npz")
    dd = synth.vae_config(int(g["ch"]), attn_layers=[1])
    sd = {**synth.vae_decoder_state_dict(dd, seed=int(g["wseed"])), **synth.vae_encoder_state_dict(dd, seed=int(g["wseed"]))}
    sd = {k: torch.from_numpy(v) for k, v in sd.items()}
    with torch.no_grad():
        mel = O.vae_decode(sd, dd, torch.from_numpy(synth.synth_latent(2, 24, seed=int(g["zseed"])))).numpy()
        mom = O.vae_encode_moments(sd, dd, torch.from_numpy(synth.synth_mel(2, 48, seed=int(g["xseed"])))).numpy()
    np.testing.assert_allclose(mel, g["mel"], atol=2e-5)
    np.testing.assert_allclose(mom, g["moments"], atol=2e-5)

