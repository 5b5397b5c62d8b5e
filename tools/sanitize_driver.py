"""Small-shape self-check workload - run plain (ALCM_GUARD=1 is set below: every device buffer is fenced by 4 KB zero
guard zones that are verified after each op and after the decodes) or under compute-sanitizer where the pool allows
it (tools/sanitize.sh).  Covers every form of the conv kernel (plain, persistent
two-accumulator, workspace split-K, cluster/DSMEM split-K, narrow operands with the zeroed K slab), both Activation1d tile sizes, GroupNorm, attention and one small end-to-end decode per mode.
Each case is also checked numerically, so a sanitizer-clean run is a correct run."""
import os
import sys

os.environ.setdefault("ALCM_GUARD", "1")

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiolcm_b200 import AutoencoderKLDecoder, LatentToWaveform, VocoderBigVGAN, ops, synth  # noqa: E402
from tests.util import round_operand  # noqa: E402

DEV = "cuda:0"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


def conv_case(name, B, Cin, Cout, T, K, d, precision="bf16", env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    x, w, b = rnd(B, Cin, T, seed=1), rnd(Cout, Cin, K, seed=2, scale=1 / np.sqrt(Cin * K)), rnd(Cout, seed=3, scale=0.1)
    ref = F.conv1d(round_operand(x, precision).double(), round_operand(w, precision).double(), b.double(), dilation=d,
                   padding=(K * d - d) // 2)
    y = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), None, dilation=d, precision=precision).cpu()
    err = float((y.double() - ref).abs().max())
    for k in (env or {}):
        os.environ.pop(k)
    print(f"  conv {name:28s} {precision} err {err:.2e}", flush=True)
    assert err < 3e-5 * max(1.0, float(ref.abs().max()))


def main():
    from oracle import decode_oracle as O
    conv_case("plain", 1, 96, 96, 129, 5, 1)
    conv_case("plain tf32", 1, 96, 96, 129, 5, 1, "tf32")
    conv_case("persistent 2-acc", 8, 32, 32, 6000, 3, 1)
    conv_case("persistent narrow C=24", 8, 24, 24, 6000, 11, 5)
    conv_case("workspace split-K", 1, 1536, 768, 100, 1, 1, env={"ALCM_CLUSTER_SPLITK": "0"})
    conv_case("cluster split-K", 1, 1536, 1536, 312, 3, 1)
    conv_case("cluster split-K tf32", 1, 1536, 1536, 100, 3, 1, "tf32")
    # every Activation1d kernel form
    for variant in (0, 1, 2):
        os.environ["ALCM_ACT_VARIANT"] = str(variant)
        for prec in ("bf16", "tf32"):
            for (B, C, T) in ((1, 8, 3), (2, 24, 1300)):
                x, al, be = rnd(B, C, T, seed=9, scale=1.5), rnd(C, seed=10, scale=0.5), rnd(C, seed=11, scale=0.5)
                ref = O.activation1d(x.double(), al.double(), be.double(), O.kaiser_sinc_filter().double())
                y = ops.activation1d(x.to(DEV), al.to(DEV), be.to(DEV), prec).cpu()
                err = float(((y.double() - ref).abs() / (ref.abs() + 1)).max())
                assert err < {"bf16": 2.0 ** -8, "tf32": 2.0 ** -11}[prec] * 1.01 + 1e-5, (variant, prec, err)
        print(f"  act1d variant {variant} ok", flush=True)
    os.environ.pop("ALCM_ACT_VARIANT")
    x = rnd(2, 64, 33, seed=12)
    ops.groupnorm_swish(x.to(DEV), torch.ones(64, device=DEV), torch.zeros(64, device=DEV))
    q = rnd(1, 64, 33, seed=13)
    ops.attn1d(q.to(DEV), q.to(DEV), q.to(DEV))
    ops.attn1d(q.to(DEV), q.to(DEV), q.to(DEV), "bf16")
    ops.attn1d(rnd(2, 256, 130, seed=19).to(DEV), rnd(2, 256, 130, seed=20).to(DEV), rnd(2, 256, 130, seed=21).to(DEV), "tf32")
    print("  groupnorm, attention ok", flush=True)
    dd, h = synth.vae_config(32), synth.bigvgan_config(64)
    vsd, gsd = synth.vae_decoder_state_dict(dd, seed=1), synth.bigvgan_state_dict(h, seed=1)
    z = synth.synth_latent(2, 16, seed=1)
    with torch.no_grad():
        ref = O.bigvgan_forward(gsd, h, O.decode_first_stage(vsd, dd, z)).squeeze(1).numpy()
    for prec, tol in (("bf16", 5e-3), ("tf32", 1e-3)):
        pipe = LatentToWaveform(AutoencoderKLDecoder(vsd, dd, synth.VAE_EMBED_DIM, DEV, prec), VocoderBigVGAN.from_state_dict(gsd, h, DEV, prec))
        err = float(np.abs(pipe.decode(z) - ref).max())
        for shape in ((1, 7), (3, 33), (2, 16)):      # more plans, replays of a cached plan
            pipe.decode(synth.synth_latent(*shape, seed=2))
        bad = pipe.vae.check_guards() + pipe.voc.check_guards()
        print(f"  decode[{prec}] err {err:.2e}; guard-zone bytes modified: {bad}", flush=True)
        assert err <= tol and bad == 0
    torch.cuda.synchronize()
    print(f"sanitize driver: all cases numerically correct, no guard zone touched (ALCM_GUARD={os.environ.get('ALCM_GUARD')})", flush=True)


if __name__ == "__main__":
    main()
