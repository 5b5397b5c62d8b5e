"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs only where /root/reference
exists (the build container); the GPU box uses the committed fixtures.

    python -m oracle.make_golden            # writes tests/golden/

The reference is imported as-is with two ``sys.modules`` stubs for packages that
are absent here and unused on the forward path:
``omegaconf`` (only VocoderBigVGAN.__init__, models.py:397) and
``pytorch_lightning`` (base class only, autoencoder1d.py:18).

Weights come from audiolcm_b200/synth.py (seeded numpy), loaded with load_state_dict,
so fixtures hold only inputs' seeds and the reference's outputs.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from audiolcm_b200 import synth

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, "vocoder", "bigvgan"))


def import_reference():
    """Returns (BigVGAN, AutoencoderKL, Activation1d, SnakeBeta) classes of the reference."""
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")
        om.OmegaConf = type("OmegaConf", (), {})
        om.ListConfig = list
        sys.modules["omegaconf"] = om
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = nn.Module
        sys.modules["pytorch_lightning"] = pl
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from vocoder.bigvgan.models import BigVGAN
    from vocoder.bigvgan.activations import SnakeBeta
    from vocoder.bigvgan.alias_free_torch import Activation1d
    from ldm.models.autoencoder1d import AutoencoderKL
    return BigVGAN, AutoencoderKL, Activation1d, SnakeBeta


def _to_torch(sd):
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def ref_bigvgan(h, sd_np):
    BigVGAN, _, _, _ = import_reference()
    g = BigVGAN(h).eval()
    missing, unexpected = g.load_state_dict(_to_torch(sd_np), strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith("filter") for k in missing), [k for k in missing if not k.endswith("filter")]
    return g


def ref_vae(dd, sd_np, embed_dim=synth.VAE_EMBED_DIM, strict=False):
    _, AutoencoderKL, _, _ = import_reference()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        vae = AutoencoderKL(embed_dim=embed_dim, ddconfig=dd, lossconfig={"target": "torch.nn.Identity"}).eval()
    missing, unexpected = vae.load_state_dict(_to_torch(sd_np), strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("encoder.", "quant_conv.")) for k in missing), missing
    assert not (strict and missing), missing
    return vae


def encoder_goldens():
    """VAE encoder: AutoencoderKL.encode(x).parameters = (mean | logvar)."""
    for tag, ch, T, B in (("ch32", 32, 40, 2), ("full_T64", 384, 64, 1)):
        dd = synth.vae_config(ch)
        esd = synth.vae_encoder_state_dict(dd, seed=5)
        vae = ref_vae(dd, {**synth.vae_decoder_state_dict(dd, seed=3), **esd}, strict=True)
        x = synth.synth_mel(B, T, seed=6)
        with torch.no_grad():
            mom = vae.encode(torch.from_numpy(x)).parameters.numpy()
        np.savez(os.path.join(OUT, f"vae_enc_{tag}.npz"), ch=ch, T=T, B=B, wseed=5, xseed=6, moments=mom)
        print("vae encode", tag, mom.shape, float(np.abs(mom).max()))


VARIANTS = {   # the other generator choices BigVGAN.__init__ accepts (models.py:146,158-172; activations.py:9-120)
    "rb2_snake": dict(resblock="2", activation="snake", snake_logscale=True, resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]]),
    "rb1_linear": dict(resblock="1", activation="snakebeta", snake_logscale=False),
    "rb2_snakebeta_linear": dict(resblock="2", activation="snakebeta", snake_logscale=False,
                                 resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]]),
}


def level_attention_goldens():
    """Decoder1D / Encoder1D with an AttnBlock1D after every ResnetBlock1D of a level (attn_layers holding a level index:
    autoencoder1d.py:356-358,466-468), narrow (ch = 32) - a choice the constructors accept and the shipped config does not use."""
    dd = synth.vae_config(32, attn_layers=[1])
    vae = ref_vae(dd, {**synth.vae_decoder_state_dict(dd, seed=21), **synth.vae_encoder_state_dict(dd, seed=21)}, strict=True)
    z = synth.synth_latent(2, 24, seed=22)
    mel = vae.decode(torch.from_numpy(z)).numpy()
    x = synth.synth_mel(2, 48, seed=23)
    mom = vae.encode(torch.from_numpy(x)).parameters.numpy()
    np.savez(os.path.join(OUT, "vae_ch32_level_attn.npz"), ch=32, wseed=21, zseed=22, xseed=23, mel=mel, moments=mom)
    print("vae level attention", mel.shape, float(np.abs(mel).max()), mom.shape, float(np.abs(mom).max()))


def variant_goldens():
    """BigVGAN(h) of the unmodified reference for AMPBlock2 / Snake / linear-scale parameters, narrow (c0 = 64)."""
    for tag, over in VARIANTS.items():
        h = synth.bigvgan_config(64, **over)
        sd = synth.bigvgan_state_dict(h, seed=5)
        g = ref_bigvgan(h, sd)
        mel = synth.synth_mel(2, 21, seed=6)
        wav = g(torch.from_numpy(mel)).numpy()
        np.savez(os.path.join(OUT, f"bigvgan_c64_{tag}.npz"), c0=64, T=21, B=2, wseed=5, xseed=6, wav=wav)
        print("bigvgan variant", tag, wav.shape, float(np.abs(wav).max()))


def main():
    assert reference_available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    if "--encoder-only" in sys.argv:
        with torch.no_grad():
            encoder_goldens()
        return
    if "--variants-only" in sys.argv:
        with torch.no_grad():
            variant_goldens()
            level_attention_goldens()
        return
    torch.set_grad_enabled(False)
    _, _, Activation1d, SnakeBeta = import_reference()

    # ---- Activation1d(SnakeBeta): shapes incl. edge cases (tiny T, odd C) -----------------
    cases = {}
    for name, (B, C, T) in {"a": (2, 5, 37), "b": (1, 24, 300), "c": (1, 3, 1), "d": (1, 2, 7), "e": (2, 16, 129)}.items():
        rng = np.random.default_rng(100 + len(cases))
        x = rng.standard_normal((B, C, T)).astype(np.float32) * 1.5
        al = (0.5 * rng.standard_normal(C)).astype(np.float32)
        be = (0.5 * rng.standard_normal(C)).astype(np.float32)
        act = Activation1d(activation=SnakeBeta(C, alpha_logscale=True)).eval()
        act.act.alpha.data = torch.from_numpy(al)
        act.act.beta.data = torch.from_numpy(be)
        y = act(torch.from_numpy(x)).numpy()
        y64 = act.double()(torch.from_numpy(x).double()).numpy()
        cases.update({f"{name}_x": x, f"{name}_alpha": al, f"{name}_beta": be, f"{name}_y": y, f"{name}_y64": y64})
        if name == "a":
            cases["filter"] = act.upsample.filter.float().numpy().reshape(-1)
            assert torch.equal(act.upsample.filter, act.downsample.lowpass.filter)
    np.savez(os.path.join(OUT, "activation1d.npz"), **cases)

    # ---- BigVGAN small configs (same topology, narrower) + layer taps ---------------------
    for tag, c0, T, B in (("c64", 64, 33, 2), ("c256", 256, 20, 1)):
        h = synth.bigvgan_config(c0)
        sd = synth.bigvgan_state_dict(h, seed=1)
        g = ref_bigvgan(h, sd)
        mel = synth.synth_mel(B, T, seed=2)
        taps = {}
        hooks = [
            g.conv_pre.register_forward_hook(lambda m, i, o: taps.__setitem__("conv_pre", o.numpy().copy())),
            g.ups[0][0].register_forward_hook(lambda m, i, o: taps.__setitem__("ups0", o.numpy().copy())),
            g.resblocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("res0", o.numpy().copy())),
        ]
        wav = g(torch.from_numpy(mel)).numpy()
        for hk in hooks:
            hk.remove()
        np.savez(os.path.join(OUT, f"bigvgan_{tag}.npz"), c0=c0, T=T, B=B, wseed=1, xseed=2, wav=wav, **taps)

    # ---- BigVGAN full 16k config: 10 s clip (config 1 of BASELINE.json) + a short one ------
    h = synth.bigvgan_config()
    sd = synth.bigvgan_state_dict(h, seed=0)
    g = ref_bigvgan(h, sd)
    assert sum(p.numel() for p in g.parameters()) == 112_231_250
    for tag, T in (("T625", 625), ("T40", 40)):
        mel = synth.synth_mel(1, T, seed=0)
        wav = g(torch.from_numpy(mel)).numpy()
        np.savez(os.path.join(OUT, f"bigvgan_full_{tag}.npz"), T=T, wseed=0, xseed=0, wav=wav.astype(np.float32))
        print("bigvgan full", tag, wav.shape, float(np.abs(wav).max()))

    # ---- VAE decoder: small + full ------------------------------------------------------------
    for tag, ch, T, B in (("ch32", 32, 24, 2), ("full", 384, 312, 1), ("full_T17", 384, 17, 1)):
        dd = synth.vae_config(ch)
        sd = synth.vae_decoder_state_dict(dd, seed=3)
        vae = ref_vae(dd, sd)
        z = synth.synth_latent(B, T, seed=4)
        taps = {}
        hooks = [
            vae.decoder.mid.block_1.register_forward_hook(lambda m, i, o: taps.__setitem__("mid1", o.numpy().copy())),
            vae.decoder.mid.attn_1.register_forward_hook(lambda m, i, o: taps.__setitem__("attn", o.numpy().copy())),
        ]
        mel = vae.decode(torch.from_numpy(z)).numpy()
        for hk in hooks:
            hk.remove()
        if tag == "full":
            taps = {}
        np.savez(os.path.join(OUT, f"vae_{tag}.npz"), ch=ch, T=T, B=B, wseed=3, xseed=4, mel=mel, **taps)
        print("vae", tag, mel.shape, float(np.abs(mel).max()), float(mel.std()))

    encoder_goldens()
    variant_goldens()
    level_attention_goldens()

    # ---- full path latent -> mel -> wav (config 2), short clip to keep the fixture small -----
    dd = synth.vae_config()
    vae = ref_vae(dd, synth.vae_decoder_state_dict(dd, seed=3))
    z = synth.synth_latent(1, 24, seed=5)
    mel = vae.decode(torch.from_numpy(z))
    wav = g(mel).numpy()
    np.savez(os.path.join(OUT, "path_full_T24.npz"), T=24, mel=mel.numpy(), wav=wav)
    print("path", wav.shape, float(np.abs(wav).max()))


if __name__ == "__main__":
    main()
