"""log10-mel front-end on the sm_100a kernels (first half).

Restates ``MelNet.forward`` (/root/reference/ldm/data/preprocess/NAT_mel.py:64-85) with the BigVGAN-16k analysis
parameters (vocoder/bigvgan/bigvgan_audioset16khz_80band.json: n_fft = win = 1024, hop 256, 80 mels, 0-8000 Hz):

    y = clamp(y, -1, 1); reflect-pad (n_fft-hop)/2 on both sides
    spec = |STFT(y, hann, center=False)| = sqrt(re^2 + im^2 + 1e-9)
    mel  = log10(clamp(mel_basis @ spec, 1e-5))          mel_basis = librosa.filters.mel (slaney scale, slaney norm)

The STFT is a GEMM: with the padded waveform folded into hop-sized rows X[n] = y[256 n : 256 n + 256], frame n is
sum_{i<4} X[n+i] . D[i] with D the windowed DFT basis cut into four 256-column blocks, i.e. a 4-tap stride-1 ``Conv1d`` with
256 input and 2*520 output channels - exactly what ``conv_umma_kernel`` runs (as a 5-tap 'same' conv whose first tap is
zero; frame n = output row n+1).  The mel projection is a 1x1 conv (513 -> 80).  One native plan per (B, L)
(``alcm_melspec_*``): fold kernel, conv, magnitude kernel, conv, log10 + unpack kernel; the bases are built here in numpy
and handed over as conv weights.  Used for the on-GPU log-mel parity metric and as the input stage of
``AutoencoderKLEncoder``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def slaney_mel_basis(sr=16000, n_fft=1024, n_mels=80, fmin=0.0, fmax=8000.0):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) (htk=False, norm='slaney'), restated in numpy."""
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


class MelSpectrogramB200(object):
    """``MelNet(hparams)(y)`` for y (B, L) with L a multiple of hop -> (B, n_mels, L/hop) float32 CUDA tensor."""

    def __init__(self, device="cuda", precision="tf32", sr=16000, n_fft=1024, hop=256, win=1024, n_mels=80, fmin=0.0, fmax=8000.0):
        if win != n_fft or n_fft % hop or (n_fft // hop) % 2:
            raise NotImplementedError("win_size == fft_size, an even multiple of hop_size, is what the shipped configs use")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.AlcmError("audiolcm_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device, self.n_fft, self.hop, self.n_mels = dev, n_fft, hop, n_mels
        self.taps = n_fft // hop                                        # 4
        nb = n_fft // 2 + 1                                             # 513 bins
        nbp = (nb + 7) // 8 * 8                                         # 520: the imaginary half starts on a 16-byte unit
        n = np.arange(n_fft, dtype=np.float64)
        window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)            # torch.hann_window (periodic)
        ang = 2.0 * np.pi * np.outer(np.arange(nb), n) / n_fft
        basis = np.zeros((2 * nbp, n_fft))
        basis[:nb], basis[nbp:nbp + nb] = np.cos(ang) * window[None, :], -np.sin(ang) * window[None, :]      # re | im
        # Conv1d weight (Cout = 2*520, Cin = hop, K = taps + 1): tap 0 zero, tap j holds columns [(j-1)*hop, j*hop)
        w = np.zeros((2 * nbp, hop, self.taps + 1), np.float32)
        for j in range(1, self.taps + 1):
            w[:, :, j] = basis[:, (j - 1) * hop:j * hop]
        melw = np.zeros((n_mels, nbp, 1), np.float32)
        melw[:, :nb, 0] = slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)
        self.nb = nb
        h = C.c_void_p()
        with torch.cuda.device(dev):
            wt, mt = torch.from_numpy(w).to(dev), torch.from_numpy(melw).to(dev)
            torch.cuda.synchronize()
            _lib.check(_lib.load().alcm_melspec_create(_lib.ctx(dev.index), wt.data_ptr(), mt.data_ptr(), hop, self.taps, nbp, n_mels,
                                                       _lib.PREC[precision], C.byref(h)))
        self._h = h.value

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().alcm_melspec_destroy(h)
            except Exception:
                pass
            self._h = None

    @torch.no_grad()
    def __call__(self, y):
        if isinstance(y, np.ndarray):
            y = torch.from_numpy(y)
        if y.dim() == 1:
            y = y.unsqueeze(0)
        y = y.to(self.device, torch.float32).contiguous()
        B, L = y.shape
        if L % self.hop or L < self.n_fft:
            raise ValueError(f"waveform length must be a multiple of hop_size {self.hop} and at least one window")
        mel = torch.empty((B, self.n_mels, L // self.hop), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().alcm_melspec_run(self._h, y.data_ptr(), mel.data_ptr(), B, L, torch.cuda.current_stream().cuda_stream))
        return mel
