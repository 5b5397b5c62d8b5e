"""Pure-numpy closed forms of the index arithmetic the CUDA kernels implement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  These are the formulas of
SURVEY.md section 8(a) rows a4/a7/a17, written as explicit loops over taps so
that the polyphase decompositions used on the device are checked against the
torch-op restatement (oracle/decode_oracle.py) and, through it, the reference.
"""
from __future__ import annotations

import numpy as np

# kaiser_sinc_filter1d(cutoff=0.25, half_width=0.3, kernel_size=12), float64 value of
# /root/reference/vocoder/bigvgan/alias_free_torch/filter.py:28-57
def kaiser_sinc_12() -> np.ndarray:
    k, half = 12, 6
    A = 2.285 * (half - 1) * np.pi * (4 * 0.3) + 7.95
    beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A < 50 else 0.1102 * (A - 8.7)
    n = np.arange(k)
    window = np.i0(beta * np.sqrt(1 - ((n - (k - 1) / 2) / ((k - 1) / 2)) ** 2)) / np.i0(beta)
    time = np.arange(-half, half) + 0.5
    f = 2 * 0.25 * window * np.sinc(2 * 0.25 * time)
    return f / f.sum()


def activation1d_closed_form(x: np.ndarray, alpha: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """x (B,C,T).  up: y[2n]   = 2*sum_q x[c(n-3+q)] f[11-2q]
                      y[2n+1] = 2*sum_q x[c(n-2+q)] f[10-2q]     (c = clamp to [0,T-1])
               act: y += sin^2(y e^alpha) / (e^beta + 1e-9)
               down: out[m] = sum_k y[clamp(2m+k-5, 0, 2T-1)] f[k]"""
    f = kaiser_sinc_12().astype(x.dtype)
    B, C, T = x.shape
    idx = np.arange(T)
    y = np.zeros((B, C, 2 * T), dtype=x.dtype)
    for q in range(6):
        y[..., 0::2] += 2 * x[..., np.clip(idx - 3 + q, 0, T - 1)] * f[11 - 2 * q]
        y[..., 1::2] += 2 * x[..., np.clip(idx - 2 + q, 0, T - 1)] * f[10 - 2 * q]
    a = np.exp(alpha).reshape(1, C, 1).astype(x.dtype)
    b = np.exp(beta).reshape(1, C, 1).astype(x.dtype)
    y = y + (1.0 / (b + 1e-9)) * np.sin(y * a) ** 2
    out = np.zeros_like(x)
    for k in range(12):
        out += y[..., np.clip(2 * idx + k - 5, 0, 2 * T - 1)] * f[k]
    return out


def conv_transpose1d_polyphase(x: np.ndarray, w: np.ndarray, b: np.ndarray, u: int) -> np.ndarray:
    """ConvTranspose1d with kernel 2u, stride u, padding u/2 (vocoder/bigvgan/models.py:150-155).
    x (B,Cin,T), w (Cin,Cout,2u).  Output phase r of block q (t = u*q + r) has two taps:
        s = r + u/2;  s <  u : x[q]*w[s]   + x[q-1]*w[s+u]
                      s >= u : x[q]*w[s-u] is wrong index -> x[q+1]*w[s-u] + x[q]*w[s]
    """
    B, Cin, T = x.shape
    Cout = w.shape[1]
    assert w.shape[2] == 2 * u and u % 2 == 0
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1)))  # xp[q+1] = x[q]
    y = np.zeros((B, Cout, u * T), dtype=x.dtype)
    for r in range(u):
        s = r + u // 2
        if s < u:
            taps = ((0, s), (-1, s + u))
        else:
            taps = ((+1, s - u), (0, s))
        acc = np.zeros((B, Cout, T), dtype=x.dtype)
        for off, kk in taps:
            acc += np.einsum("bit,io->bot", xp[..., 1 + off:1 + off + T], w[:, :, kk])
        y[..., r::u] = acc + b.reshape(1, -1, 1)
    return y


def nearest2x_conv3_polyphase(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Upsample1D (ldm/models/autoencoder1d.py:291-295): nearest x2 then Conv1d k3 p1.
    even t=2q:  W0*x[q-1] + (W1+W2)*x[q];   odd t=2q+1: (W0+W1)*x[q] + W2*x[q+1]."""
    B, Cin, T = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1)))
    y = np.zeros((B, w.shape[0], 2 * T), dtype=x.dtype)
    e = lambda off, ww: np.einsum("bit,oi->bot", xp[..., 1 + off:1 + off + T], ww)
    y[..., 0::2] = e(-1, w[:, :, 0]) + e(0, w[:, :, 1] + w[:, :, 2])
    y[..., 1::2] = e(0, w[:, :, 0] + w[:, :, 1]) + e(+1, w[:, :, 2])
    return y + b.reshape(1, -1, 1)
