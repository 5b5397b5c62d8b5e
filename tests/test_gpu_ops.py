"""GPU parity tests of the single kernels, called through the C-ABI (audiolcm_b200.ops).

fp32 : CUDA-core path, compared with torch fp32/fp64 ops at ~1e-5.
tf32 / bf16 : tcgen05 path; operands are rounded exactly as the device does (tests/util.py) and the
reference is evaluated in float64, so what is left is accumulation order: tolerance 2e-5 relative
to the output scale.  A wrong descriptor/stride/tap shift shows up as O(1) error.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import round_operand

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from audiolcm_b200 import ops as _ops
    return _ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


# ----------------------------------------------------------------------------------- Activation1d
@pytest.mark.parametrize("case", ["a", "b", "c", "d", "e"])
def test_activation1d_golden_fp32(ops, golden_dir, case):
    g = np.load(os.path.join(golden_dir, "activation1d.npz"))
    x, al, be = (torch.from_numpy(g[f"{case}_{n}"]).to(DEV) for n in ("x", "alpha", "beta"))
    y = ops.activation1d(x, al, be, "fp32").cpu().numpy()
    np.testing.assert_allclose(y, g[f"{case}_y64"], atol=1e-5, rtol=1e-5)   # reference module output (float64 run)
    np.testing.assert_allclose(y, g[f"{case}_y"], atol=2e-5, rtol=1e-5)     # and its fp32 run


@pytest.mark.parametrize("shape", [(1, 768, 2500), (2, 24, 4099), (1, 17, 511), (1, 16, 513), (3, 40, 1030)])
@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "fp16"])
def test_activation1d_vs_oracle(ops, shape, precision):
    from oracle import decode_oracle as O
    B, C, T = shape
    x = _rand(B, C, T, seed=1, scale=1.5)
    al, be = _rand(C, seed=2, scale=0.5), _rand(C, seed=3, scale=0.5)
    ref = O.activation1d(x.double(), al.double(), be.double(), O.kaiser_sinc_filter().double())
    y = ops.activation1d(x.to(DEV), al.to(DEV), be.to(DEV), precision).cpu()
    tol = {"fp32": 1e-5, "tf32": 2.0 ** -11, "bf16": 2.0 ** -8, "fp16": 2.0 ** -11}[precision]
    err = (y.double() - ref).abs()
    assert float((err / (ref.abs() + 1.0)).max()) < tol * 1.01 + 1e-5, float(err.max())
    if precision != "fp32":  # output must be exactly representable in the operand type
        assert torch.equal(round_operand(y, precision), y)


def test_fp16_operands_saturate_instead_of_overflowing(ops):
    """precision="fp16": a value beyond the fp16 range becomes +-65504 (cvt.rn.satfinite), never inf - an out-of-range
    activation degrades the result, it cannot poison a whole accumulation with inf - inf = NaN."""
    x = torch.full((1, 8, 64), 3.0e5)
    x[:, 4:] = -3.0e5
    z = torch.zeros(8)
    y = ops.activation1d(x.to(DEV), z.to(DEV), z.to(DEV), "fp16").cpu()
    assert torch.isfinite(y).all()
    assert torch.equal(y[:, :4], torch.full((1, 4, 64), 65504.0)) and torch.equal(y[:, 4:], torch.full((1, 4, 64), -65504.0))
    w = torch.zeros(8, 8, 1)
    w[torch.arange(8), torch.arange(8), 0] = 1.0
    c = ops.conv1d(x.to(DEV), w.to(DEV), None, None, 1, "fp16").cpu()      # identity conv: operands saturate, fp32 accumulate
    assert torch.equal(c, y)


# ----------------------------------------------------------------------------------- Conv1d
CONV_CASES = [
    # B, Cin, Cout, T, K, d
    (1, 16, 16, 128, 3, 1),
    (2, 24, 24, 300, 11, 5),
    (1, 80, 192, 625, 7, 1),
    (1, 192, 192, 1000, 7, 3),
    (2, 384, 384, 257, 3, 5),
    (1, 768, 768, 250, 11, 1),
    (1, 20, 1536, 312, 5, 1),
    (1, 1536, 768, 100, 1, 1),
    (1, 48, 48, 1, 3, 1),
    (1, 5, 3, 9, 3, 1),
    (1, 96, 80, 129, 5, 1),
    (1, 1536, 1536, 312, 3, 1),   # VAE mid block: cluster split-K (DSMEM reduce-scatter), re-tiled N
    (1, 768, 768, 624, 3, 1),
    (2, 384, 384, 624, 3, 1),
    (1, 1536, 4608, 312, 1, 1),   # merged q/k/v
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "fp16"])
def test_conv1d(ops, case, precision):
    B, Cin, Cout, T, K, d = case
    x = _rand(B, Cin, T, seed=4)
    w = _rand(Cout, Cin, K, seed=5, scale=1.0 / np.sqrt(Cin * K))
    b = _rand(Cout, seed=6, scale=0.1)
    res = _rand(B, Cout, T, seed=7)
    xr, wr = round_operand(x, precision), round_operand(w, precision)
    ref = F.conv1d(xr.double(), wr.double(), b.double(), dilation=d, padding=(K * d - d) // 2) + res.double()
    y = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), res.to(DEV), dilation=d, precision=precision).cpu()
    err = float((y.double() - ref).abs().max())
    assert err < 3e-5 * max(1.0, float(ref.abs().max())), err
    y2 = ops.conv1d(x.to(DEV), w.to(DEV), None, None, dilation=d, precision=precision).cpu()
    ref2 = F.conv1d(xr.double(), wr.double(), None, dilation=d, padding=(K * d - d) // 2)
    assert float((y2.double() - ref2).abs().max()) < 3e-5 * max(1.0, float(ref2.abs().max()))


@pytest.mark.parametrize("case", [(1, 1536, 768, 625, 4), (2, 96, 48, 333, 2), (1, 48, 24, 1000, 2), (1, 8, 4, 5, 4), (1, 4, 2, 1, 2)])
@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "fp16"])
def test_conv_transpose1d(ops, case, precision):
    B, Cin, Cout, T, u = case
    x = _rand(B, Cin, T, seed=8)
    w = _rand(Cin, Cout, 2 * u, seed=9, scale=1.0 / np.sqrt(Cin * 2))
    b = _rand(Cout, seed=10, scale=0.1)
    xr, wr = round_operand(x, precision), round_operand(w, precision)
    ref = F.conv_transpose1d(xr.double(), wr.double(), b.double(), stride=u, padding=u // 2)
    y = ops.conv_transpose1d(x.to(DEV), w.to(DEV), b.to(DEV), stride=u, precision=precision).cpu()
    assert y.shape == ref.shape
    assert float((y.double() - ref).abs().max()) < 3e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("case", [(1, 768, 768, 312), (2, 64, 64, 17), (1, 32, 32, 1)])
@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "fp16"])
def test_upsample_conv3(ops, case, precision):
    B, Cin, Cout, T = case
    x = _rand(B, Cin, T, seed=11)
    w = _rand(Cout, Cin, 3, seed=12, scale=1.0 / np.sqrt(Cin * 3))
    b = _rand(Cout, seed=13, scale=0.1)
    xr = round_operand(x, precision)
    # the device sums taps in fp32 BEFORE rounding to the operand type (polyphase form)
    w0, w1, w2 = w[..., 0], w[..., 1], w[..., 2]
    r = lambda t: round_operand(t.contiguous(), precision).double()
    xp = F.pad(xr.double(), (1, 1))
    e = lambda off, ww: torch.einsum("bit,oi->bot", xp[..., 1 + off:1 + off + T], ww)
    ref = torch.zeros(B, Cout, 2 * T, dtype=torch.float64)
    ref[..., 0::2] = e(-1, r(w0)) + e(0, r(w1 + w2))
    ref[..., 1::2] = e(0, r(w0 + w1)) + e(1, r(w2))
    ref += b.double().view(1, -1, 1)
    y = ops.upsample_conv3(x.to(DEV), w.to(DEV), b.to(DEV), precision=precision).cpu()
    assert float((y.double() - ref).abs().max()) < 3e-5 * max(1.0, float(ref.abs().max()))
    if precision == "fp32":  # and against the reference formulation itself
        ref0 = F.conv1d(F.interpolate(x.double(), scale_factor=2.0, mode="nearest"), w.double(), b.double(), padding=1)
        assert float((y.double() - ref0).abs().max()) < 3e-5 * max(1.0, float(ref0.abs().max()))


# ----------------------------------------------------------------------------------- GroupNorm / attention
@pytest.mark.parametrize("shape", [(1, 1536, 312), (2, 384, 624), (1, 32, 5), (2, 64, 1), (1, 128, 48)])
@pytest.mark.parametrize("swish", [True, False])
def test_groupnorm_swish(ops, shape, swish):
    B, C, T = shape
    x = _rand(B, C, T, seed=14, scale=2.0) + 0.5
    gm, bt = 1 + _rand(C, seed=15, scale=0.2), _rand(C, seed=16, scale=0.1)
    ref = F.group_norm(x.double(), 32, gm.double(), bt.double(), eps=1e-6)
    if swish:
        ref = ref * torch.sigmoid(ref)
    y = ops.groupnorm_swish(x.to(DEV), gm.to(DEV), bt.to(DEV), swish=swish).cpu()
    assert float((y.double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("shape", [(1, 1536, 312), (2, 128, 24), (1, 32, 1), (1, 64, 33), (3, 256, 130), (1, 1536, 624)])
@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "fp16"])
def test_attn1d(ops, shape, precision):
    """fp32: CUDA-core kernels; tf32 / bf16: QK^T and PV on conv_umma_kernel with per-item operands.  The reference is
    float64 on operands rounded as the device rounds them (q, k, v; the probabilities are rounded on the device too, which the
    tolerance covers: 2^-9 relative on values <= 1 for bf16, summed with weights that add up to 1)."""
    B, C, T = shape
    q, k, v = _rand(B, C, T, seed=17), _rand(B, C, T, seed=18), _rand(B, C, T, seed=19)
    qr, kr, vr = (round_operand(t, precision).double() for t in (q, k, v))
    w = torch.softmax(torch.bmm(qr.permute(0, 2, 1), kr) * (C ** -0.5), dim=2)
    ref = torch.bmm(vr, w.permute(0, 2, 1))
    y = ops.attn1d(q.to(DEV), k.to(DEV), v.to(DEV), precision).cpu()
    tol = {"fp32": 2e-5, "tf32": 1e-3, "bf16": 6e-3, "fp16": 1e-3}[precision]
    assert float((y.double() - ref).abs().max()) < tol * max(1.0, float(ref.abs().max()))


def test_bad_arguments_raise(ops):
    from audiolcm_b200 import AlcmError
    x = torch.zeros(1, 4, 8, device=DEV)
    with pytest.raises(AlcmError):
        ops.conv1d(x, torch.zeros(4, 4, 4, device=DEV))          # even kernel
    with pytest.raises(AlcmError):
        ops.conv1d(x, torch.zeros(4, 4, 11, device=DEV), dilation=9)  # halo too large
    with pytest.raises(AlcmError):
        ops.groupnorm_swish(torch.zeros(1, 20, 8, device=DEV), torch.ones(20, device=DEV), torch.zeros(20, device=DEV))
    with pytest.raises(AlcmError):
        ops.activation1d(torch.zeros(1, 4, 8), torch.zeros(4), torch.zeros(4))  # CPU tensor: no fallback


# ----------------------------------------------------------------------------------- every Activation1d kernel form
ACT_VARIANTS = {0: "128 threads, 635 outputs per block", 1: "64 threads, 315 (72 registers)", 2: "32 threads, 155", 3: "64 threads, 315 (96 registers)"}
ACT_EDGE_SHAPES = [(1, 8, 1), (1, 8, 3), (2, 24, 4099), (1, 16, 635), (1, 16, 636), (1, 8, 630), (1, 8, 1283), (1, 8, 315), (1, 8, 316),
                   (1, 8, 311), (1, 8, 155), (1, 8, 156), (1, 8, 152), (1, 24, 6), (1, 40, 1925), (3, 16, 160000)]


@pytest.mark.parametrize("variant", sorted(ACT_VARIANTS))
@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "fp16"])
def test_activation1d_every_kernel_form(ops, monkeypatch, variant, precision):
    """The plans pick one block size of the Activation1d kernel; here each of the three is forced
    (ALCM_ACT_VARIANT) and run over tile-boundary and tiny shapes (T = 1: every tap is replicate padding;
    T = tile, tile+1, tile-3: the halo'd edges of the staged tile) against the float64 oracle."""
    from oracle import decode_oracle as O
    monkeypatch.setenv("ALCM_ACT_VARIANT", str(variant))
    tol = {"fp32": 1e-5, "tf32": 2.0 ** -11, "bf16": 2.0 ** -8, "fp16": 2.0 ** -11}[precision]
    for i, (B, C, T) in enumerate(ACT_EDGE_SHAPES):
        x = _rand(B, C, T, seed=30 + i, scale=1.5)
        al, be = _rand(C, seed=2, scale=0.5), _rand(C, seed=3, scale=0.5)
        ref = O.activation1d(x.double(), al.double(), be.double(), O.kaiser_sinc_filter().double())
        y = ops.activation1d(x.to(DEV), al.to(DEV), be.to(DEV), precision).cpu()
        err = (y.double() - ref).abs()
        assert float((err / (ref.abs() + 1.0)).max()) < tol * 1.01 + 1e-5, (ACT_VARIANTS[variant], (B, C, T), float(err.max()))


def test_narrow_operand_planes_are_not_padded_to_16(ops):
    """C = 24 / 20 / 8 operands: the conv's missing K chunk is a zeroed shared-memory slab, not an HBM plane.
    Checked against the same conv with the 16-channel padding forced in a separate process (ALCM_UNPADDED=0)
    would need a re-import; instead: results must match the float64 reference for every narrow shape in all modes
    (a stale or non-zero slab shows up as O(1) error)."""
    for (B, Cin, Cout, T, K, d) in [(2, 24, 24, 1000, 11, 5), (1, 20, 20, 77, 1, 1), (1, 8, 8, 300, 7, 1), (3, 24, 48, 513, 3, 3),
                                    (1, 40, 24, 260, 7, 1)]:
        for precision in ("tf32", "bf16", "fp32"):
            x = _rand(B, Cin, T, seed=41)
            w = _rand(Cout, Cin, K, seed=42, scale=1.0 / np.sqrt(Cin * K))
            b = _rand(Cout, seed=43, scale=0.1)
            xr, wr = round_operand(x, precision), round_operand(w, precision)
            ref = F.conv1d(xr.double(), wr.double(), b.double(), dilation=d, padding=(K * d - d) // 2)
            for _ in range(2):  # twice: persistent ring slots are reused
                y = ops.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), None, dilation=d, precision=precision).cpu()
                assert float((y.double() - ref).abs().max()) < 3e-5 * max(1.0, float(ref.abs().max())), (Cin, Cout, precision)


def test_lcm_step_matches_reference_golden(ops, golden_dir):
    """The fused LCM step kernel driven by the ported schedule reproduces the REAL LCMSampler.lcm_sampling output
    (tests/golden/lcm_denoiser.npz): DiT in PyTorch on the GPU, step() as alcm_lcm_step, same noise draw."""
    from baseline import lcm_denoiser_port as P
    g = np.load(os.path.join(golden_dir, "lcm_denoiser.npz"))
    den = P.PortedDenoiser(P.dit_state_dict(seed=int(g["wseed"])), DEV)
    x, ctx = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["ctx"]).to(DEV)
    ts = den.schedule.timesteps(2)
    w_emb = P.guidance_scale_embedding(torch.tensor(4.0).repeat(x.shape[0]), 256).to(DEV)
    torch.manual_seed(int(g["noise_seed"]))
    noise = torch.randn(x.shape)                 # the reference drew it on the CPU generator (golden made on CPU)
    img = x
    for i, t in enumerate(ts):
        with torch.no_grad():
            eps = den(img, torch.full((x.shape[0],), t, device=DEV, dtype=torch.long), ctx, w_emb)
        k = den.schedule.coefficients(ts, i)
        img, denoised = ops.lcm_step(img, eps, None if k["last"] else noise.to(DEV), k["a_t_sqrt"], k["b_t_sqrt"], k["c_out"], k["c_skip"],
                                     k["a_prev_sqrt"], k["b_prev_sqrt"], k["last"])
    scale = max(1.0, float(np.abs(g["denoised"]).max()))
    assert np.abs(denoised.cpu().numpy() - g["denoised"]).max() <= 2e-3 * scale     # GPU DiT (TF32 convs by torch default) vs CPU golden
    # and exactly the reference arithmetic on identical inputs
    s, e, z = _rand(3, 20, 50, seed=60).to(DEV), _rand(3, 20, 50, seed=61).to(DEV), _rand(3, 20, 50, seed=62).to(DEV)
    k = den.schedule.coefficients(ts, 0)
    prev, d = ops.lcm_step(s, e, z, k["a_t_sqrt"], k["b_t_sqrt"], k["c_out"], k["c_skip"], k["a_prev_sqrt"], k["b_prev_sqrt"], False)
    x0 = (s.double() - float(k["b_t_sqrt"]) * e.double()) / float(k["a_t_sqrt"])
    dref = float(k["c_out"]) * x0 + float(k["c_skip"]) * s.double()
    pref = float(k["a_prev_sqrt"]) * dref + float(k["b_prev_sqrt"]) * z.double()
    assert float((d.double() - dref).abs().max()) <= 1e-5 * float(dref.abs().max())
    assert float((prev.double() - pref).abs().max()) <= 1e-5 * float(pref.abs().max())


# ----------------------------------------------------------------------------------- DiT feed-forward convs
@pytest.mark.parametrize("precision", ["tf32", "bf16", "fp16"])
def test_conv1d_layer_handle(precision):
    """alcm_conv1d_* (persistent layer: weights packed once, a plan per (B,T)) on the DiT's feed-forward shape (k = 9)."""
    from audiolcm_b200.denoiser import Conv1dLayer
    Cin, Cout, K = 576, 1152, 9
    w = _rand(Cout, Cin, K, seed=70, scale=1.0 / np.sqrt(Cin * K))
    b = _rand(Cout, seed=71, scale=0.1)
    layer = Conv1dLayer(w, b, 1, DEV, precision)
    wr = round_operand(w, precision).double()
    for (B, T, with_res) in ((2, 467, False), (1, 50, True), (2, 467, True), (3, 129, False), (2, 467, False)):
        x = _rand(B, Cin, T, seed=72 + T)
        res = _rand(B, Cout, T, seed=73) if with_res else None
        ref = F.conv1d(round_operand(x, precision).double(), wr, b.double(), padding=4)
        if with_res:
            ref = ref + res.double()
        y = layer(x.to(DEV), None if res is None else res.to(DEV)).cpu()
        assert float((y.double() - ref).abs().max()) < 3e-5 * max(1.0, float(ref.abs().max()))


def test_layernorm_channels_first():
    """alcm_layernorm_cf = nn.LayerNorm(C) of the DiT blocks (new_attention.py:246-248) applied to the channels-first stream."""
    from audiolcm_b200 import ops
    for (B, Cc, T) in ((2, 576, 467), (1, 3, 1), (3, 130, 129), (1, 64, 1000)):
        x = _rand(B, Cc, T, seed=90 + T) + 0.5
        g, b = _rand(Cc, seed=91) + 1.0, _rand(Cc, seed=92)
        ref = F.layer_norm(x.double().permute(0, 2, 1), (Cc,), g.double(), b.double(), 1e-5).permute(0, 2, 1)
        y = ops.layernorm_cf(x.to(DEV), g.to(DEV), b.to(DEV)).cpu()
        assert float((y.double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "fp16"])
def test_conv1d_feedforward_handle(precision):
    """alcm_ffn1d_* = Conv1dFeedForward(glu=True) (new_attention.py:38-74): conv k9 -> x * gelu(gate) -> conv k9 (+ res) as one plan,
    against the same ops in float64 (operands of the second conv rounded as the kernel rounds them)."""
    from audiolcm_b200.denoiser import Conv1dFeedForwardLayer
    dim, inner, K = 64, 136, 9      # inner % 16 != 0: the gate half starts inside a 16-channel padding group
    w1 = _rand(2 * inner, dim, K, seed=80, scale=1.0 / np.sqrt(dim * K))
    b1 = _rand(2 * inner, seed=81, scale=0.1)
    w2 = _rand(dim, inner, K, seed=82, scale=1.0 / np.sqrt(inner * K))
    b2 = _rand(dim, seed=83, scale=0.1)
    layer = Conv1dFeedForwardLayer(w1, b1, w2, b2, DEV, precision)
    for (B, T, with_res) in ((2, 467, True), (1, 50, False), (3, 129, True), (2, 467, True)):
        x = _rand(B, dim, T, seed=84 + T)
        res = _rand(B, dim, T, seed=85) if with_res else None
        h = F.conv1d(round_operand(x, precision).double(), round_operand(w1, precision).double(), b1.double(), padding=K // 2)
        a, gate = h.chunk(2, dim=1)
        mid = (a * F.gelu(gate)).float()
        ref = F.conv1d(round_operand(mid, precision).double(), round_operand(w2, precision).double(), b2.double(), padding=K // 2)
        if with_res:
            ref = ref + res.double()
        y = layer(x.to(DEV), None if res is None else res.to(DEV)).cpu()
        # the intermediate is re-rounded to the operand type: a 1-ulp flip of it moves the output by ~ulp * |w|
        tol = {"fp32": 3e-5, "tf32": 1e-3, "bf16": 4e-3, "fp16": 1e-3}[precision]
        assert float((y.double() - ref).abs().max()) < tol * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("precision", ["tf32", "bf16", "fp16"])
def test_hybrid_dit_and_sampler_match_reference_golden(golden_dir, precision):
    """ConcatDiT2MLPB200 (feed-forward convs on conv_umma_kernel, the rest PyTorch) and LCMSamplerB200 (fused step kernel)
    against tests/golden/lcm_denoiser.npz = outputs of the REAL ConcatDiT2MLP / LCMSampler.lcm_sampling on CPU."""
    from baseline import lcm_denoiser_port as P
    from audiolcm_b200.denoiser import ConcatDiT2MLPB200, LCMSamplerB200
    g = np.load(os.path.join(golden_dir, "lcm_denoiser.npz"))
    dit = ConcatDiT2MLPB200(P.dit_state_dict(seed=int(g["wseed"])), DEV, precision)
    x, ctx, t = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["ctx"]).to(DEV), torch.from_numpy(g["t"]).to(DEV)
    w_emb = LCMSamplerB200.guidance_embedding(torch.tensor(4.0).repeat(x.shape[0])).to(DEV)
    eps = dit(x, t, ctx, w_emb).cpu().numpy()
    tol = {"tf32": 3e-3, "bf16": 3e-2, "fp16": 3e-3}[precision]
    err = np.abs(eps - g["eps"]).max() / np.abs(g["eps"]).max()
    print(f"\n[hybrid DiT {precision}] eps max-abs error relative to abs-max: {err:.2e}")
    assert err <= tol
    smp = LCMSamplerB200(dit)
    assert smp.timesteps(2) == [int(v) for v in g["timesteps"]]
    torch.manual_seed(int(g["noise_seed"]))
    noise = torch.randn(x.shape)                      # the golden drew its step noise on the CPU generator
    import audiolcm_b200.denoiser as D
    real_randn = torch.randn
    try:
        D.torch.randn = lambda *a, **k: noise.to(DEV)   # same draw as the reference run
        den = smp.sample(ctx, T=x.shape[-1], steps=2, guidance_scale=5.0, x_T=x)
    finally:
        D.torch.randn = real_randn
    errd = np.abs(den.cpu().numpy() - g["denoised"]).max() / np.abs(g["denoised"]).max()
    print(f"[hybrid sampler {precision}] denoised max-abs error relative to abs-max: {errd:.2e}")
    assert errd <= 4 * tol
