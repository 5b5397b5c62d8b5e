"""CPU tests of the reference-side pieces around the decode path (BASELINE.json configs[4]):
the LCM sampler + ConcatDiT2MLP port (baseline/lcm_denoiser_port.py) against goldens made from the unmodified
reference classes (oracle/make_golden_lcm.py), and install() exercised on the REAL LCM_audio where /root/reference
exists (the build container)."""
import os

import numpy as np
import pytest
import torch

from baseline import lcm_denoiser_port as P
from oracle import reference_harness as H
from audiolcm_b200 import synth


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "lcm_denoiser.npz"))


def test_dit_port_matches_reference_forward(golden_dir):
    g = _golden(golden_dir)
    sd = P._t(P.dit_state_dict(seed=int(g["wseed"])), "cpu")
    x, ctx, t = torch.from_numpy(g["x"]), torch.from_numpy(g["ctx"]), torch.from_numpy(g["t"])
    w_emb = P.guidance_scale_embedding(torch.tensor(4.0).repeat(x.shape[0]), 256)
    with torch.no_grad():
        eps = P.concat_dit2mlp(sd, x, t, ctx, w_emb).numpy()
    assert eps.shape == g["eps"].shape
    assert np.abs(eps - g["eps"]).max() <= 2e-5 * max(1.0, np.abs(g["eps"]).max())


def test_lcm_sampler_port_matches_reference_two_step_sampling(golden_dir):
    g = _golden(golden_dir)
    den = P.PortedDenoiser(P.dit_state_dict(seed=int(g["wseed"])), "cpu")
    assert den.schedule.timesteps(2) == [int(v) for v in g["timesteps"]] == [999, 499]
    x, ctx = torch.from_numpy(g["x"]), torch.from_numpy(g["ctx"])
    torch.manual_seed(int(g["noise_seed"]))          # the reference draws the step noise from the global generator
    denoised, img = P.lcm_sample(den, ctx, tuple(x.shape), steps=2, guidance_scale=5.0, x_T=x.clone(), schedule=den.schedule)
    scale = max(1.0, float(np.abs(g["denoised"]).max()))
    assert np.abs(denoised.numpy() - g["denoised"]).max() <= 1e-4 * scale
    assert np.abs(img.numpy() - g["img"]).max() <= 1e-4 * scale
    k = den.schedule.coefficients([999, 499], 1)
    assert k["last"] and abs(float(k["c_skip"]) - 0.25 / (4990.0 ** 2 + 0.25)) < 1e-12


def test_dit_shapes_cover_the_reference_state_dict():
    shapes = P.dit_tensor_shapes()
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert 150e6 < n < 170e6          # "DiffusionWrapper has 159.69 M params" (reference log line)
    sd = P.dit_state_dict(seed=1)
    assert set(sd) == set(shapes) and all(sd[k].shape == tuple(v) for k, v in shapes.items())


@pytest.mark.skipif(not H.available(), reason="/root/reference is not present on this box")
def test_install_on_the_real_lcm_audio(monkeypatch, golden_dir):
    """install() rebinding on the real class (lcm_audio.py:392-406): after install, LCM_audio.decode_first_stage(z)
    divides by scale_factor in REFERENCE code and hands the result to our decoder, which is built from the real
    first_stage_model.state_dict() through vae_tensor_names().  No GPU here, so the CUDA decoder is replaced by a
    stand-in with the same constructor that checks names / shapes and decodes with the CPU oracle; the numerics of the
    CUDA decoder against this very decode_first_stage are the GPU test on tests/golden/lcm_decode_first_stage.npz."""
    import audiolcm_b200
    from audiolcm_b200 import autoencoder as A
    from oracle import decode_oracle as O
    model, ddconfig = H.build_lcm_audio()
    dd = synth.vae_config()
    for k in ("ch", "out_ch", "z_channels", "kernel_size", "num_res_blocks"):
        assert int(ddconfig[k]) == int(dd[k])
    assert list(ddconfig["ch_mult"]) == list(dd["ch_mult"]) and list(ddconfig["down_layers"]) == list(dd["down_layers"])
    vsd = synth.vae_decoder_state_dict(dd, seed=3)
    fsm = model.first_stage_model
    cur = fsm.state_dict()
    assert set(vsd) <= set(cur)                       # every tensor our loader asks for exists in the real module
    cur.update({k: torch.from_numpy(v) for k, v in vsd.items()})
    fsm.load_state_dict(cur)
    model.scale_factor.fill_(0.7)
    g = np.load(os.path.join(golden_dir, "lcm_decode_first_stage.npz"))
    z = torch.from_numpy(synth.synth_latent(int(g["B"]), int(g["T"]), seed=int(g["xseed"])))
    with torch.no_grad():
        want = model.decode_first_stage(z).numpy()
    np.testing.assert_allclose(want, g["mel"], atol=2e-6)     # the committed golden is this module's output
    calls = []

    class StandIn(object):
        def __init__(self, state_dict, ddconfig, embed_dim, device="cuda", precision="tf32", prefix=""):
            names = A.vae_tensor_names({k: ddconfig[k] for k in ("ch", "ch_mult", "num_res_blocks", "down_layers")}, prefix)
            assert all(n in state_dict for n in names)
            assert embed_dim == 20
            self.sd = {n: state_dict[n].detach().clone() for n in names}
            self.dd = dict(ddconfig)

        def decode(self, zz, inv_scale=1.0):
            calls.append(zz.clone())
            with torch.no_grad():
                return O.decode_first_stage(self.sd, self.dd, zz * inv_scale)

    monkeypatch.setattr(A, "AutoencoderKLDecoder", StandIn)
    dec = audiolcm_b200.install(model, ddconfig, device="cuda:0", precision="tf32")
    assert isinstance(dec, StandIn)
    with torch.no_grad():
        got = model.decode_first_stage(z).numpy()
    assert len(calls) == 1 and torch.allclose(calls[0], z / 0.7)        # scale_factor handled by the reference wrapper
    assert np.abs(got - want).max() <= 2e-5
