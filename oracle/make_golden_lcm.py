"""Golden vectors for the denoiser-side port and for decode_first_stage, from the UNMODIFIED reference classes
(run in the build container; /root/reference must exist):

    python -m oracle.make_golden_lcm

tests/golden/lcm_denoiser.npz      ConcatDiT2MLP.forward and LCMSampler.lcm_sampling (2 steps, guidance 5) of the real
                                   LCM_audio with seeded synthetic DiT weights (baseline/lcm_denoiser_port.dit_state_dict)
tests/golden/lcm_decode_first_stage.npz
                                   LCM_audio.decode_first_stage(z) (lcm_audio.py:392-406) with scale_factor = 0.7 and
                                   the seeded synthetic VAE decoder weights of audiolcm_b200/synth.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiolcm_b200 import synth  # noqa: E402
from baseline import lcm_denoiser_port as P  # noqa: E402
from oracle import reference_harness as H  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    model, ddconfig = H.build_lcm_audio()
    from ldm.models.diffusion.scheduling_lcm import LCMSampler
    out = os.path.join(ROOT, "tests", "golden")
    # ---- denoiser ----
    dit_sd = P.dit_state_dict(seed=7)
    missing, unexpected = model.unet.diffusion_model.load_state_dict({k: torch.from_numpy(v) for k, v in dit_sd.items()}, strict=True), None
    B, T = 2, 24
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 20, T, generator=g)
    ctx = torch.randn(B, 154, 1024, generator=g)
    t = torch.tensor([999, 499])
    sampler = LCMSampler(model)
    w_emb = sampler.get_guidance_scale_embedding(torch.tensor(4.0).repeat(B), embedding_dim=256)
    with torch.no_grad():
        eps = model.apply_model(x, t, ctx, model.unet, w_cond=w_emb)
    # lcm_sampling as sample() runs it (scheduling_lcm.py:326-342) minus make_schedule's hard-wired .to("cuda")
    sampler.alphas_cumprod = model.alphas_cumprod.clone().float()
    sampler.num_inference_steps = 2
    torch.manual_seed(123)
    with torch.no_grad():
        denoised, img = sampler.lcm_sampling(ctx, (B, 20, T), x_T=x.clone(), guidance_scale=5.0, original_inference_steps=50)
    np.savez(os.path.join(out, "lcm_denoiser.npz"), x=x.numpy(), ctx=ctx.numpy(), t=t.numpy(), eps=eps.numpy(), denoised=denoised.numpy(),
             img=img.numpy(), wseed=7, noise_seed=123, timesteps=np.asarray(sampler.timesteps))
    print("lcm_denoiser.npz: eps abs-max %.3f, denoised abs-max %.3f, timesteps %s" % (eps.abs().max(), denoised.abs().max(), sampler.timesteps.tolist()))
    # ---- decode_first_stage ----
    dd = synth.vae_config()
    vsd = synth.vae_decoder_state_dict(dd, seed=3)
    fsm = model.first_stage_model
    cur = fsm.state_dict()
    cur.update({k: torch.from_numpy(v) for k, v in vsd.items()})
    fsm.load_state_dict(cur)
    model.scale_factor.fill_(0.7)
    z = torch.from_numpy(synth.synth_latent(2, 12, seed=21))
    with torch.no_grad():
        mel = model.decode_first_stage(z)
    np.savez(os.path.join(out, "lcm_decode_first_stage.npz"), mel=mel.numpy(), scale_factor=0.7, wseed=3, xseed=21, B=2, T=12)
    print("lcm_decode_first_stage.npz: mel", tuple(mel.shape), "abs-max %.3f" % mel.abs().max())


if __name__ == "__main__":
    main()
