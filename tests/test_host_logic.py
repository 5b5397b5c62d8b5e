"""CPU-only tests: C-ABI surface, host-side tensor ordering, sharding arithmetic, gloo halo exchange."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from audiolcm_b200 import _lib
from audiolcm_b200.autoencoder import vae_tensor_names
from audiolcm_b200.pipeline import shard_range, halo_frames
from audiolcm_b200.vocoder import bigvgan_tensor_names
from audiolcm_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    from audiolcm_b200 import build
    return build.build()


def test_library_exports_every_header_symbol():
    _built()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "audiolcm_b200.h")).read()
    declared = set(re.findall(r"\b(alcm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_errors_are_codes_not_crashes():
    _built()
    lib = _lib.load()
    assert lib.alcm_vocoder_num_tensors(None) == -1
    h = ctypes.c_void_p()
    if not torch.cuda.is_available():
        rc = lib.alcm_ctx_create(ctypes.byref(h), 0)
        assert rc < 0 and len(lib.alcm_last_error()) > 0
    rc = lib.alcm_ctx_create(None, 0)
    assert rc < 0 and b"NULL" in lib.alcm_last_error()
    rc = lib.alcm_vocode(None, None, 1, 1, None, None)
    assert rc == -1


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from audiolcm_b200 import VocoderBigVGAN
    h = synth.bigvgan_config(64)
    with pytest.raises(Exception):
        VocoderBigVGAN.from_state_dict(synth.bigvgan_state_dict(h), h, device="cuda")
    with pytest.raises(_lib.AlcmError):
        VocoderBigVGAN.from_state_dict(synth.bigvgan_state_dict(h), h, device="cpu")


def test_tensor_order_matches_abi_count():
    _built()
    lib = _lib.load()
    h = synth.bigvgan_config()
    names = bigvgan_tensor_names(h)
    sd = synth.bigvgan_state_dict(h, seed=0) if False else None
    cfg = _lib.BigVGANCfg()
    cfg.num_upsamples, cfg.num_kernels = 6, 3
    assert lib.alcm_vocoder_num_tensors(ctypes.byref(cfg)) == len(names) == 3 + 6 * (3 + 3 * 30) + 5
    small = synth.bigvgan_state_dict(synth.bigvgan_config(64), seed=0)
    assert set(bigvgan_tensor_names(synth.bigvgan_config(64))) == set(small)
    dd = synth.vae_config()
    vnames = vae_tensor_names(dd)
    vsd = synth.vae_decoder_state_dict(synth.vae_config(32), seed=0)
    assert set(vae_tensor_names(synth.vae_config(32))) == set(vsd)
    vcfg = _lib.VAECfg()
    vcfg.ch, vcfg.n_levels, vcfg.num_res_blocks = 384, 3, 2
    for i, m in enumerate(dd["ch_mult"]):
        vcfg.ch_mult[i] = m
    vcfg.upsample_levels[1] = 1
    assert lib.alcm_vae_num_tensors(ctypes.byref(vcfg)) == len(vnames)


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _halo_worker(rank, world, port, T, out_q):
    import torch.distributed as dist
    from audiolcm_b200.pipeline import vocode_time_sharded
    from oracle import decode_oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    h = synth.bigvgan_config(64)
    sd = synth.bigvgan_state_dict(h, seed=5)
    mel = torch.from_numpy(synth.synth_mel(1, T, seed=9))
    s, e = shard_range(T, rank, world)
    fn = lambda m: O.bigvgan_forward(sd, h, m, torch.float64).squeeze(1)
    with torch.no_grad():
        part = vocode_time_sharded(fn, mel[..., s:e].double(), rank, world, hop=256)
    out_q.put((rank, part.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_time_sharded_vocode_matches_unsharded_gloo():
    """Config 4 host logic at world_size 2 on CPU: halo exchange + trim reproduces the un-sharded
    oracle (float64) to round-off - the 34-frame halo covers the receptive field."""
    import torch.multiprocessing as mp
    from oracle import decode_oracle as O
    T, world = 150, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    got = np.concatenate([parts[r] for r in range(world)], axis=-1)
    h = synth.bigvgan_config(64)
    sd = synth.bigvgan_state_dict(h, seed=5)
    with torch.no_grad():
        ref = O.bigvgan_forward(sd, h, torch.from_numpy(synth.synth_mel(1, T, seed=9)), torch.float64).squeeze(1).numpy()
    assert got.shape == ref.shape == (1, T * 256)
    assert np.abs(got - ref).max() < 1e-9
    assert halo_frames() == 34


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints exactly one JSON line with
    the contract's keys; on a box without a GPU the product arm must refuse instead of falling back."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    import torch
    if not torch.cuda.is_available():
        ours = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                              timeout=600, cwd=root)
        assert ours.returncode != 0 and "no CPU fallback" in (ours.stderr + ours.stdout)


def test_slaney_mel_basis_matches_torchaudio():
    """The numpy restatement of librosa.filters.mel (slaney scale and norm) used by the GPU mel front-end equals
    torchaudio's slaney filterbank (librosa itself is not in the image)."""
    import torchaudio
    from audiolcm_b200.melspec import slaney_mel_basis
    a = slaney_mel_basis(16000, 1024, 80, 0.0, 8000.0)
    b = torchaudio.functional.melscale_fbanks(513, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    assert a.shape == b.shape == (80, 513)
    assert np.abs(a - b).max() <= 1e-6
