"""Shared helpers for the parity tests (oracle side only - never imported by the product)."""
import numpy as np
import torch


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32: round to 10 mantissa bits, ties away from zero (sign-magnitude add)."""
    bits = x.contiguous().view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF
    return bits.view(torch.float32)


def round_operand(x: torch.Tensor, precision: str) -> torch.Tensor:
    if precision == "tf32":
        return round_tf32(x.float())
    if precision == "bf16":
        return x.float().bfloat16().float()
    if precision == "fp16":
        return x.float().clamp(-65504.0, 65504.0).half().float()
    return x.float()


def snr_db(ref: np.ndarray, got: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64)
    err = np.asarray(got, np.float64) - ref
    return float(10 * np.log10((ref ** 2).sum() / max((err ** 2).sum(), 1e-300)))


def sd_to(sd, device):
    return {k: torch.from_numpy(v).to(device) for k, v in sd.items()}


def log_mel(wav: np.ndarray, sr: int = 16000, n_fft: int = 1024, hop: int = 256, win: int = 1024, n_mels: int = 80,
            fmin: float = 0.0, fmax: float = 8000.0) -> np.ndarray:
    """log10-mel spectrogram as the reference computes it (ldm/data/preprocess/NAT_mel.py:64-85 with the
    BigVGAN-16k analysis parameters of vocoder/bigvgan/bigvgan_audioset16khz_80band.json); the slaney
    filterbank comes from torchaudio because librosa is not in the image."""
    import torchaudio
    y = torch.as_tensor(np.asarray(wav, np.float32)).reshape(1, -1).clamp(-1.0, 1.0)
    pad = (n_fft - hop) // 2
    y = torch.nn.functional.pad(y.unsqueeze(1), [pad, pad], mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win), center=False, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9)
    fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, fmin, fmax, n_mels, sr, norm="slaney", mel_scale="slaney").T
    return torch.log10(torch.clamp(fb @ mag[0], min=1e-5)).numpy()


def log_mel_l1(ref: np.ndarray, got: np.ndarray) -> float:
    """mean |log10-mel(ref) - log10-mel(got)| (log10 units)"""
    return float(np.abs(log_mel(ref) - log_mel(got)).mean())
