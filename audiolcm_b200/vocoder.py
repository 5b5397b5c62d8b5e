"""Drop-in for ``vocoder.bigvgan.models.VocoderBigVGAN`` backed by the sm_100a kernels.

Mirrors /root/reference/vocoder/bigvgan/models.py:393-414: same constructor arguments
(``ckpt_vocoder`` directory holding ``best_netG.pt`` + ``args.yml``, ``device``), same
``vocode(spec)`` contract (ndarray ``(80,T)`` or tensor ``(B,80,T)`` in, host float32 ndarray
``.squeeze()``d out) and ``__call__``.  Select it by changing the ``target:`` string of
``configs/audiolcm.yaml:90-93`` or by constructing it where the scripts construct the reference
class (``pythonscripts/InferAPI.py:121,153``).  PyTorch is used for tensor hand-off only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib

_GEN_KEYS = ("resblock", "upsample_rates", "upsample_kernel_sizes", "upsample_initial_channel",
             "resblock_kernel_sizes", "resblock_dilation_sizes", "activation", "snake_logscale", "num_mels")


def _get(h, k):
    try:
        return h[k]
    except (TypeError, KeyError, IndexError):
        return getattr(h, k)


def _as_cuda_f32(t, device):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def bigvgan_tensor_names(h):
    """state_dict keys in the order alcm_vocoder_create expects (include/audiolcm_b200.h).  ``activation: snake`` has no
    ``beta`` parameters (activations.py:9-62): its alpha is listed twice, Snake being SnakeBeta with beta = alpha."""
    names = []
    wn = lambda p: [p + ".weight_g", p + ".weight_v", p + ".bias"]
    ab = (lambda p: [p + ".alpha", p + ".beta"]) if _get(h, "activation") == "snakebeta" else (lambda p: [p + ".alpha", p + ".alpha"])
    rb2 = str(_get(h, "resblock")) == "2"
    names += wn("conv_pre")
    nk = len(_get(h, "resblock_kernel_sizes"))
    for i in range(len(_get(h, "upsample_rates"))):
        names += wn(f"ups.{i}.0")
        for j in range(nk):
            p = f"resblocks.{i * nk + j}"
            if rb2:                                   # AMPBlock2, models.py:90-126
                for l in range(2):
                    names += wn(f"{p}.convs.{l}")
                for m in range(2):
                    names += ab(f"{p}.activations.{m}.act")
                continue
            for l in range(3):
                names += wn(f"{p}.convs1.{l}")
            for l in range(3):
                names += wn(f"{p}.convs2.{l}")
            for m in range(6):
                names += ab(f"{p}.activations.{m}.act")
    names += ab("activation_post.act")
    names += wn("conv_post")
    return names


def _remove_parametrization_names(sd):
    """Accept torch>=2.1 parametrized weight_norm state_dicts (original0 = g, original1 = v)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".parametrizations.weight.original0"):
            out[k[: -len(".parametrizations.weight.original0")] + ".weight_g"] = v
        elif k.endswith(".parametrizations.weight.original1"):
            out[k[: -len(".parametrizations.weight.original1")] + ".weight_v"] = v
        else:
            out[k] = v
    return out


class VocoderBigVGAN(object):
    def __init__(self, ckpt_vocoder, device="cuda", precision="tf32"):
        sd = torch.load(os.path.join(ckpt_vocoder, "best_netG.pt"), map_location="cpu")  # models.py:395
        h = self._load_args(os.path.join(ckpt_vocoder, "args.yml"))                      # models.py:397
        self._setup(sd["generator"], h, device, precision)

    @staticmethod
    def _load_args(path):
        try:
            from omegaconf import OmegaConf
            return OmegaConf.load(path)
        except ImportError:
            import yaml
            with open(path) as f:
                return yaml.safe_load(f)

    @classmethod
    def from_state_dict(cls, state_dict, h, device="cuda", precision="tf32"):
        """Random-init / in-memory construction: ``state_dict`` of a reference ``BigVGAN`` (numpy or
        torch values), ``h`` its hyper-parameters (dict or attribute object)."""
        self = cls.__new__(cls)
        self._setup(state_dict, h, device, precision)
        return self

    @classmethod
    def from_module(cls, generator, device="cuda", precision="tf32"):
        """From a live reference ``BigVGAN`` module (weight norm kept or removed)."""
        sd = {k: v for k, v in generator.state_dict().items()}
        for name, mod in generator.named_modules():  # remove_weight_norm()'d modules: g = ||w||, v = w
            if hasattr(mod, "weight") and (name + ".weight_g") not in sd and (name + ".weight") in sd and \
                    (name + ".parametrizations.weight.original0") not in sd and mod.weight.dim() == 3:
                w = sd[name + ".weight"]
                sd[name + ".weight_v"] = w
                sd[name + ".weight_g"] = torch.linalg.vector_norm(w, dim=(1, 2), keepdim=True)
        return cls.from_state_dict(sd, generator.h, device, precision)

    # ------------------------------------------------------------------------------------------
    def _setup(self, sd, h, device, precision):
        if precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {sorted(_lib.PREC)}")
        if _get(h, "activation") not in ("snake", "snakebeta"):
            # the reference raises the same for unknown activations (models.py:70,115,172)
            raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")
        rb2 = str(_get(h, "resblock")) != "1"          # models.py:146: anything but '1' selects AMPBlock2
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.AlcmError("audiolcm_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.precision = precision
        self.h = {k: _get(h, k) for k in _GEN_KEYS}
        self.h["resblock"] = "2" if rb2 else "1"
        rates = [int(u) for u in self.h["upsample_rates"]]
        ksz = [int(k) for k in self.h["upsample_kernel_sizes"]]
        rks = [int(k) for k in self.h["resblock_kernel_sizes"]]
        rds = [[int(d) for d in dd] for dd in self.h["resblock_dilation_sizes"]]
        if len(rates) > 8 or len(rks) > 4 or any(len(d) != (2 if rb2 else 3) for d in rds):
            raise NotImplementedError("unsupported BigVGAN topology")
        self.num_mels = int(self.h["num_mels"])
        self.hop = int(np.prod(rates))
        cfg = _lib.BigVGANCfg()
        cfg.num_mels = self.num_mels
        cfg.upsample_initial_channel = int(self.h["upsample_initial_channel"])
        cfg.num_upsamples = len(rates)
        cfg.num_kernels = len(rks)
        cfg.resblock2 = int(rb2)
        cfg.snake_linear = int(not _get(h, "snake_logscale"))
        for i, (u, k) in enumerate(zip(rates, ksz)):
            cfg.upsample_rates[i] = u
            cfg.upsample_kernel_sizes[i] = k
        for j, (k, dd) in enumerate(zip(rks, rds)):
            cfg.resblock_kernel_sizes[j] = k
            for l, d in enumerate(dd):
                cfg.resblock_dilation_sizes[j][l] = d
        sd = _remove_parametrization_names(sd)
        names = bigvgan_tensor_names(self.h)
        missing = [n for n in names if n not in sd]
        if missing:
            raise KeyError(f"generator state_dict is missing {len(missing)} tensors, e.g. {missing[:3]}")
        self._check_filters(sd)
        lib = _lib.load()
        with torch.cuda.device(dev):
            tensors = [_as_cuda_f32(sd[n], dev) for n in names]
            torch.cuda.synchronize()
            handle = C.c_void_p()
            _lib.check(lib.alcm_vocoder_create(_lib.ctx(dev.index), C.byref(cfg), _lib.ptr_array(tensors), len(tensors),
                                               _lib.PREC[precision], C.byref(handle)))
        self._h = handle.value
        del tensors

    @staticmethod
    def _check_filters(sd):
        """The kernels hard-code the Kaiser-sinc taps; a checkpoint that stores different ``filter``
        buffers (alias_free_torch/resample.py:19-22) must not be silently mis-decoded."""
        ref = np.array([0.0020289647, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832, 0.4432097971], np.float64)
        for k, v in sd.items():
            if k.endswith("filter"):
                f = np.asarray(v.detach().cpu() if torch.is_tensor(v) else v, np.float64).reshape(-1)
                if f.size != 12 or np.abs(f[:6] - ref).max() > 1e-6 or np.abs(f[::-1][:6] - ref).max() > 1e-6:
                    raise ValueError(f"checkpoint filter buffer {k} differs from kaiser_sinc_filter1d(0.25, 0.3, 12)")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().alcm_vocoder_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------------------------------
    def vocode_tensor(self, spec):
        """(B,num_mels,T) tensor -> (B, T*hop) float32 CUDA tensor (no host round trip)."""
        spec = spec.to(dtype=torch.float32, device=self.device).contiguous()
        if spec.dim() != 3 or spec.shape[1] != self.num_mels:
            raise ValueError(f"expected a (B,{self.num_mels},T) spectrogram, got {tuple(spec.shape)}")
        B, _, T = spec.shape
        if B == 0 or T == 0:
            raise ValueError("empty spectrogram")
        with torch.cuda.device(self.device):
            wav = torch.empty((B, T * self.hop), dtype=torch.float32, device=self.device)
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.load().alcm_vocode(self._h, spec.data_ptr(), B, T, wav.data_ptr(), stream))
        return wav

    def vocode_pcm16_tensor(self, spec):
        """(B,num_mels,T) tensor -> (B, T*hop) int16 CUDA tensor: 16-bit PCM packed on the device,
        ``rint(wav * 32767)`` - the samples ``soundfile.write(path, wav, 16000)`` stores
        (pythonscripts/InferAPI.py:98)."""
        spec = spec.to(dtype=torch.float32, device=self.device).contiguous()
        if spec.dim() != 3 or spec.shape[1] != self.num_mels or spec.shape[0] == 0 or spec.shape[2] == 0:
            raise ValueError(f"expected a non-empty (B,{self.num_mels},T) spectrogram, got {tuple(spec.shape)}")
        B, _, T = spec.shape
        with torch.cuda.device(self.device):
            pcm = torch.empty((B, T * self.hop), dtype=torch.int16, device=self.device)
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.load().alcm_vocode_pcm16(self._h, spec.data_ptr(), B, T, pcm.data_ptr(), stream))
        return pcm

    def plan(self, B, T):
        """Build the (B,T) plan now (workspace slab, kernel list, CUDA graph) on the current stream, so that
        later ``vocode`` calls of this shape neither allocate nor synchronise.  Returns its size in bytes."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().alcm_vocoder_plan(self._h, int(B), int(T), torch.cuda.current_stream().cuda_stream))
        return self.workspace_bytes(B, T)

    def workspace_bytes(self, B, T):
        n = C.c_size_t()
        _lib.check(_lib.load().alcm_vocoder_workspace_bytes(self._h, int(B), int(T), C.byref(n)))
        return int(n.value)

    def vocode(self, spec):
        """models.py:406-411."""
        with torch.no_grad():
            if isinstance(spec, np.ndarray):
                spec = torch.from_numpy(spec).unsqueeze(0)
            return self.vocode_tensor(spec).unsqueeze(1).squeeze().cpu().numpy()

    def __call__(self, wav):
        return self.vocode(wav)

    def check_guards(self):
        """ALCM_GUARD=1 self-check: bytes of the guard zones around this handle's buffers that a kernel overwrote."""
        n = C.c_longlong()
        _lib.check(_lib.load().alcm_vocoder_check_guards(self._h, C.byref(n)))
        return int(n.value)

    def launches(self, B, T):
        return _lib.load().alcm_vocoder_launches(self._h, B, T)
