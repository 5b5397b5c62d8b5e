"""Drop-in for the decode half of ``ldm.models.autoencoder1d.AutoencoderKL``.

Mirrors /root/reference/ldm/models/autoencoder1d.py:59-62 (``decode(z)``: post_quant_conv ->
Decoder1D.forward, :484-517) as called by ``LCM_audio.decode_first_stage``
(/root/reference/ldm/models/diffusion/lcm_audio.py:392-406).  ``install(model)`` swaps
``model.first_stage_model.decode`` in place so every reference call site keeps working; the
``z / scale_factor`` stays in reference code.  Encode / training stay reference PyTorch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .vocoder import _as_cuda_f32, _get


def vae_tensor_names(dd, prefix=""):
    """state_dict keys in the order alcm_vae_create expects (include/audiolcm_b200.h)."""
    ch, mult, nrb = int(dd["ch"]), [int(m) for m in dd["ch_mult"]], int(dd["num_res_blocks"])
    nl = len(mult)
    wb = lambda p: [prefix + p + ".weight", prefix + p + ".bias"]

    def res(p, cin, cout):
        n = wb(p + ".norm1") + wb(p + ".conv1") + wb(p + ".norm2") + wb(p + ".conv2")
        return n + (wb(p + ".nin_shortcut") if cin != cout else [])

    names = wb("post_quant_conv") + wb("decoder.conv_in")
    block_in = ch * mult[nl - 1]
    names += res("decoder.mid.block_1", block_in, block_in)
    for n in ("norm", "q", "k", "v", "proj_out"):
        names += wb(f"decoder.mid.attn_1.{n}")
    names += res("decoder.mid.block_2", block_in, block_in)
    down_layers = [i + 1 for i in dd["down_layers"]]  # autoencoder1d.py:427
    attn_layers = [int(a) for a in dd.get("attn_layers", [])]   # constructor default: no level attention
    for lv in reversed(range(nl)):
        block_out = ch * mult[lv]
        for ib in range(nrb + 1):
            names += res(f"decoder.up.{lv}.block.{ib}", block_in, block_out)
            block_in = block_out
            if lv in attn_layers:                       # autoencoder1d.py:466-468
                for n in ("norm", "q", "k", "v", "proj_out"):
                    names += wb(f"decoder.up.{lv}.attn.{ib}.{n}")
        if lv in down_layers:
            names += wb(f"decoder.up.{lv}.upsample.conv")
    return names + wb("decoder.norm_out") + wb("decoder.conv_out")


class AutoencoderKLDecoder(object):
    """``decode(z)``: (B,embed_dim,T) CUDA tensor -> (B,out_ch,T*2^n_up) float32 CUDA tensor."""

    def __init__(self, state_dict, ddconfig, embed_dim, device="cuda", precision="tf32", prefix=""):
        if precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {sorted(_lib.PREC)}")
        dd = {k: _get(ddconfig, k) for k in ("ch", "out_ch", "z_channels", "kernel_size", "ch_mult", "num_res_blocks",
                                              "attn_layers", "down_layers")}
        dd["ch_mult"] = [int(m) for m in dd["ch_mult"]]
        dd["down_layers"] = [int(i) for i in dd["down_layers"]]
        nl = len(dd["ch_mult"])
        dd["attn_layers"] = [int(a) for a in dd["attn_layers"]]
        for k in ("give_pre_end", "tanh_out"):
            try:
                if _get(ddconfig, k):
                    raise NotImplementedError(f"{k}=True is not implemented")
            except (AttributeError, KeyError):
                pass
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.AlcmError("audiolcm_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device, self.precision, self.dd, self.embed_dim = dev, precision, dd, int(embed_dim)
        down_layers = [i + 1 for i in dd["down_layers"]]
        if any(l >= nl for l in down_layers):
            raise NotImplementedError("down_layers outside the level range")
        self.up_factor = 2 ** len(down_layers)
        cfg = _lib.VAECfg()
        cfg.ch, cfg.out_ch, cfg.z_channels = int(dd["ch"]), int(dd["out_ch"]), int(dd["z_channels"])
        cfg.embed_dim, cfg.kernel_size = self.embed_dim, int(dd["kernel_size"])
        cfg.num_res_blocks, cfg.n_levels = int(dd["num_res_blocks"]), nl
        for i, m in enumerate(dd["ch_mult"]):
            cfg.ch_mult[i] = m
            cfg.upsample_levels[i] = 1 if i in down_layers else 0
            cfg.attn_levels[i] = 1 if i in dd["attn_layers"] else 0
        names = vae_tensor_names(dd, prefix)
        missing = [n for n in names if n not in state_dict]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} decoder tensors, e.g. {missing[:3]}")
        lib = _lib.load()
        with torch.cuda.device(dev):
            tensors = [_as_cuda_f32(state_dict[n], dev) for n in names]
            torch.cuda.synchronize()
            handle = C.c_void_p()
            _lib.check(lib.alcm_vae_create(_lib.ctx(dev.index), C.byref(cfg), _lib.ptr_array(tensors), len(tensors),
                                           _lib.PREC[precision], C.byref(handle)))
        self._h = handle.value

    @classmethod
    def from_module(cls, vae, device="cuda", precision="tf32", ddconfig=None):
        """From a live reference ``AutoencoderKL`` (needs its ddconfig: the module does not keep it)."""
        if ddconfig is None:
            raise ValueError("pass the ddconfig the AutoencoderKL was built with")
        return cls(vae.state_dict(), ddconfig, vae.embed_dim, device, precision)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().alcm_vae_destroy(h)
            except Exception:
                pass
            self._h = None

    def decode(self, z, inv_scale: float = 1.0):
        z = z.to(dtype=torch.float32, device=self.device).contiguous()
        if z.dim() != 3 or z.shape[1] != self.embed_dim:
            raise ValueError(f"expected a (B,{self.embed_dim},T) latent, got {tuple(z.shape)}")
        B, _, T = z.shape
        if B == 0 or T == 0:
            raise ValueError("empty latent")
        with torch.cuda.device(self.device):
            mel = torch.empty((B, int(self.dd["out_ch"]), T * self.up_factor), dtype=torch.float32, device=self.device)
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.load().alcm_vae_decode(self._h, z.data_ptr(), B, T, float(inv_scale), mel.data_ptr(), stream))
        return mel

    __call__ = decode

    def plan(self, B, T):
        """Build the (B,T) plan now on the current stream (see ``VocoderBigVGAN.plan``); returns its bytes."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().alcm_vae_plan(self._h, int(B), int(T), torch.cuda.current_stream().cuda_stream))
        return self.workspace_bytes(B, T)

    def workspace_bytes(self, B, T):
        n = C.c_size_t()
        _lib.check(_lib.load().alcm_vae_workspace_bytes(self._h, int(B), int(T), C.byref(n)))
        return int(n.value)

    def check_guards(self):
        n = C.c_longlong()
        _lib.check(_lib.load().alcm_vae_check_guards(self._h, C.byref(n)))
        return int(n.value)

    def launches(self, B, T):
        return _lib.load().alcm_vae_launches(self._h, B, T)


def vae_encoder_tensor_names(dd, prefix=""):
    """state_dict keys in the order alcm_vae_encoder_create expects (include/audiolcm_b200.h)."""
    ch, mult, nrb = int(dd["ch"]), [int(m) for m in dd["ch_mult"]], int(dd["num_res_blocks"])
    wb = lambda p: [prefix + p + ".weight", prefix + p + ".bias"]

    def res(p, cin, cout):
        n = wb(p + ".norm1") + wb(p + ".conv1") + wb(p + ".norm2") + wb(p + ".conv2")
        return n + (wb(p + ".nin_shortcut") if cin != cout else [])

    names = wb("encoder.conv_in")
    block_in = ch
    for lv in range(len(mult)):
        block_out = ch * mult[lv]
        for ib in range(nrb):
            names += res(f"encoder.down.{lv}.block.{ib}", block_in, block_out)
            block_in = block_out
            if lv in [int(a) for a in dd.get("attn_layers", [])]:   # autoencoder1d.py:356-358
                for n in ("norm", "q", "k", "v", "proj_out"):
                    names += wb(f"encoder.down.{lv}.attn.{ib}.{n}")
        if lv in [int(i) for i in dd["down_layers"]]:
            names += wb(f"encoder.down.{lv}.downsample.conv")
    names += res("encoder.mid.block_1", block_in, block_in)
    for n in ("norm", "q", "k", "v", "proj_out"):
        names += wb(f"encoder.mid.attn_1.{n}")
    names += res("encoder.mid.block_2", block_in, block_in)
    return names + wb("encoder.norm_out") + wb("encoder.conv_out") + wb("quant_conv")


class AutoencoderKLEncoder(object):
    """The encode half of ``ldm.models.autoencoder1d.AutoencoderKL`` (autoencoder1d.py:52-56, Encoder1D :319-413; the call
    site is scripts/reconstruct_audio.py:115) - SURVEY.md 8f row 4.  ``encode(x)``: (B,in_channels,T) mel ->
    ``(mean, logvar)`` of the posterior, each (B,embed_dim,T/2^n_down) float32 CUDA tensors (logvar clamped to [-30, 20] as
    ``DiagonalGaussianDistribution`` does); ``sample``/``mode`` follow distributions.py:24-37."""

    def __init__(self, state_dict, ddconfig, embed_dim, device="cuda", precision="tf32", prefix=""):
        if precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {sorted(_lib.PREC)}")
        dd = {k: _get(ddconfig, k) for k in ("ch", "in_channels", "z_channels", "kernel_size", "ch_mult", "num_res_blocks", "attn_layers",
                                              "down_layers")}
        dd["ch_mult"] = [int(m) for m in dd["ch_mult"]]
        dd["down_layers"] = [int(i) for i in dd["down_layers"]]
        nl = len(dd["ch_mult"])
        dd["attn_layers"] = [int(a) for a in dd["attn_layers"]]
        try:
            double_z = bool(_get(ddconfig, "double_z"))
        except (AttributeError, KeyError):
            double_z = True
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.AlcmError("audiolcm_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device, self.precision, self.dd, self.embed_dim = dev, precision, dd, int(embed_dim)
        self.down_factor = 2 ** len([l for l in dd["down_layers"] if l < nl])
        cfg = _lib.VAEEncCfg()
        cfg.ch, cfg.in_channels, cfg.z_channels = int(dd["ch"]), int(dd["in_channels"]), int(dd["z_channels"])
        cfg.embed_dim, cfg.kernel_size, cfg.num_res_blocks, cfg.n_levels = self.embed_dim, int(dd["kernel_size"]), int(dd["num_res_blocks"]), nl
        cfg.double_z = 1 if double_z else 0
        for i, m in enumerate(dd["ch_mult"]):
            cfg.ch_mult[i] = m
            cfg.downsample_levels[i] = 1 if i in dd["down_layers"] else 0
            cfg.attn_levels[i] = 1 if i in dd["attn_layers"] else 0
        names = vae_encoder_tensor_names(dd, prefix)
        missing = [n for n in names if n not in state_dict]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} encoder tensors, e.g. {missing[:3]}")
        lib = _lib.load()
        with torch.cuda.device(dev):
            tensors = [_as_cuda_f32(state_dict[n], dev) for n in names]
            torch.cuda.synchronize()
            handle = C.c_void_p()
            _lib.check(lib.alcm_vae_encoder_create(_lib.ctx(dev.index), C.byref(cfg), _lib.ptr_array(tensors), len(tensors),
                                                   _lib.PREC[precision], C.byref(handle)))
        self._h = handle.value

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().alcm_vae_encoder_destroy(h)
            except Exception:
                pass
            self._h = None

    def moments(self, x):
        x = x.to(dtype=torch.float32, device=self.device).contiguous()
        if x.dim() != 3 or x.shape[1] != int(self.dd["in_channels"]):
            raise ValueError(f"expected a (B,{self.dd['in_channels']},T) spectrogram, got {tuple(x.shape)}")
        B, _, T = x.shape
        if B == 0 or T == 0 or T % self.down_factor:
            raise ValueError(f"T must be a positive multiple of {self.down_factor}")
        with torch.cuda.device(self.device):
            mom = torch.empty((B, 2 * self.embed_dim, T // self.down_factor), dtype=torch.float32, device=self.device)
            _lib.check(_lib.load().alcm_vae_encode(self._h, x.data_ptr(), B, T, mom.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return mom

    def encode(self, x):
        mean, logvar = torch.chunk(self.moments(x), 2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)

    def sample(self, x):
        mean, logvar = self.encode(x)
        return mean + torch.exp(0.5 * logvar) * torch.randn(mean.shape, device=mean.device)

    def mode(self, x):
        return self.encode(x)[0]


def install(model, ddconfig, device="cuda", precision="tf32"):
    """Swap ``model.first_stage_model.decode`` (the body behind ``decode_first_stage``,
    lcm_audio.py:406) for the CUDA path, in place.  Returns the decoder object."""
    fsm = model.first_stage_model
    dec = AutoencoderKLDecoder(fsm.state_dict(), ddconfig, fsm.embed_dim, device, precision)
    fsm._alcm_decoder = dec
    fsm.decode = lambda z: dec.decode(z)
    return dec
