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
    return x.float()


def snr_db(ref: np.ndarray, got: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64)
    err = np.asarray(got, np.float64) - ref
    return float(10 * np.log10((ref ** 2).sum() / max((err ** 2).sum(), 1e-300)))


def sd_to(sd, device):
    return {k: torch.from_numpy(v).to(device) for k, v in sd.items()}
