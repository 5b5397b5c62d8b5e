"""Chained latent -> mel -> waveform decode and the two multi-GPU partitionings.

* Batch sharding (BASELINE.json configs 3/5): clips are independent (GroupNorm and attention are
  per-sample, autoencoder1d.py:169-170,267), so ``shard_range`` just splits the batch - no
  collective on the data path.
* Long-form time sharding (config 4): BigVGAN is purely convolutional; a perturbed mel frame is
  numerically visible for 8631/8517 samples (< 34 mel frames) each side (the
  analytic support is ~36.7 frames, but the Kaiser-sinc tails beyond 34 carry < 2e-7 of the signal:
  the stitched result equals the un-sharded one to rounding, not bit for bit), so each rank receives
  ``halo_frames()`` mel frames from each neighbour with one NCCL P2P exchange
  (``torch.distributed.batch_isend_irecv``), vocodes its extended chunk and drops the halo
  samples.  True sequence ends keep the reference's edge rules (zero pad for convs, replicate for
  the FIRs) because only interior edges are extended.  The VAE decoder does not shard along time
  (GroupNorm statistics and mid-attention span all of T): it runs replicated.

Replaces the serial per-item loop with its device->host->device mel bounce in
/root/reference/pythonscripts/InferAPI.py:87-96.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_HALO = 34


def halo_frames() -> int:
    return _HALO


def shard_range(n: int, rank: int, world: int):
    """Contiguous split of ``n`` items over ``world`` ranks; returns (start, stop) for ``rank``."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def exchange_halo(chunk: torch.Tensor, rank: int, world: int, halo: int = _HALO, group=None):
    """``chunk`` is this rank's contiguous (B,C,Tl) time slice.  Returns (extended, left, right):
    the chunk with ``left``/``right`` frames of the neighbours' edges attached."""
    import torch.distributed as dist

    B, C, Tl = chunk.shape
    if world > 1 and Tl < halo:
        raise ValueError(f"time shard of {Tl} frames is shorter than the {halo}-frame halo; use fewer ranks")
    ops, left_buf, right_buf = [], None, None
    keep = []
    if rank > 0:
        left_buf = torch.empty((B, C, halo), dtype=chunk.dtype, device=chunk.device)
        send_l = chunk[..., :halo].contiguous()
        keep.append(send_l)
        ops += [dist.P2POp(dist.isend, send_l, rank - 1, group), dist.P2POp(dist.irecv, left_buf, rank - 1, group)]
    if rank < world - 1:
        right_buf = torch.empty((B, C, halo), dtype=chunk.dtype, device=chunk.device)
        send_r = chunk[..., Tl - halo:].contiguous()
        keep.append(send_r)
        ops += [dist.P2POp(dist.isend, send_r, rank + 1, group), dist.P2POp(dist.irecv, right_buf, rank + 1, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    parts = ([left_buf] if left_buf is not None else []) + [chunk] + ([right_buf] if right_buf is not None else [])
    ext = torch.cat(parts, dim=-1) if len(parts) > 1 else chunk
    return ext, (halo if left_buf is not None else 0), (halo if right_buf is not None else 0)


def vocode_time_sharded(vocode_fn, mel_chunk: torch.Tensor, rank: int, world: int, hop: int, halo: int = _HALO, group=None):
    """Vocode one rank's time slice of a long clip.  ``vocode_fn`` maps (B,C,T) -> (B,T*hop)."""
    ext, left, right = exchange_halo(mel_chunk, rank, world, halo, group)
    wav = vocode_fn(ext)
    Text = ext.shape[-1]
    return wav[..., left * hop:(Text - right) * hop]


class LatentToWaveform(object):
    """``decode_first_stage`` + ``vocode`` as one call; the mel never leaves the device."""

    def __init__(self, vae_decoder, vocoder, micro_batch=None):
        """``micro_batch``: decode a batch as consecutive slices of at most this many clips through ONE smaller plan.  The
        plan's buffers scale with its batch (64 x 10 s clips: 33 GB as one plan, 8.6 GB at 16), and a 16-clip plan is within
        a few per cent of the 64-clip one either way (tools/batch_sweep.py: 64 clips in 134.7 ms vs 139.1 ms in bf16,
        202.2 vs 199.5 ms in tf32).  Default: one plan for the whole batch (what bench.py measures)."""
        if vae_decoder.device != vocoder.device:
            raise ValueError("VAE decoder and vocoder must live on the same device")
        if micro_batch is not None and int(micro_batch) < 1:
            raise ValueError("micro_batch must be a positive number of clips")
        self.vae, self.voc = vae_decoder, vocoder
        self.device = vocoder.device
        self.micro_batch = None if micro_batch is None else int(micro_batch)

    def _slices(self, B):
        mb = self.micro_batch or B
        return [(s, min(B, s + mb)) for s in range(0, B, mb)]

    def decode_tensor(self, z, scale_factor: float = 1.0, return_mel: bool = False):
        z = z.to(dtype=torch.float32, device=self.device).contiguous()
        if z.dim() != 3 or z.shape[1] != self.vae.embed_dim:
            raise ValueError(f"expected a (B,{self.vae.embed_dim},T) latent, got {tuple(z.shape)}")
        B, _, T = z.shape
        Tm = T * self.vae.up_factor
        with torch.cuda.device(self.device):
            wav = torch.empty((B, Tm * self.voc.hop), dtype=torch.float32, device=self.device)
            mel = torch.empty((B, self.voc.num_mels, Tm), dtype=torch.float32, device=self.device) if return_mel else None
            stream = torch.cuda.current_stream().cuda_stream
            for s, e in self._slices(B):      # slices of contiguous tensors: plain pointer offsets
                _lib.check(_lib.load().alcm_decode_to_wav(self.vae._h, self.voc._h, z[s:e].data_ptr(), e - s, T, 1.0 / float(scale_factor),
                                                          None if mel is None else mel[s:e].data_ptr(), wav[s:e].data_ptr(), stream))
        return (wav, mel) if return_mel else wav

    def decode_pcm16_tensor(self, z, scale_factor: float = 1.0):
        """latents (B,20,T) -> (B, 512*T) int16 CUDA tensor (16-bit PCM packed by the conv_post kernel)."""
        z = z.to(dtype=torch.float32, device=self.device).contiguous()
        if z.dim() != 3 or z.shape[1] != self.vae.embed_dim:
            raise ValueError(f"expected a (B,{self.vae.embed_dim},T) latent, got {tuple(z.shape)}")
        B, _, T = z.shape
        with torch.cuda.device(self.device):
            pcm = torch.empty((B, T * self.vae.up_factor * self.voc.hop), dtype=torch.int16, device=self.device)
            stream = torch.cuda.current_stream().cuda_stream
            for s, e in self._slices(B):
                _lib.check(_lib.load().alcm_decode_to_pcm16(self.vae._h, self.voc._h, z[s:e].data_ptr(), e - s, T, 1.0 / float(scale_factor),
                                                            None, pcm[s:e].data_ptr(), stream))
        return pcm

    def plan(self, B, T):
        """Pre-plan both models for latents of shape (B,20,T); returns the total workspace bytes."""
        return self.vae.plan(B, T) + self.voc.plan(B, T * self.vae.up_factor)

    def decode(self, z, scale_factor: float = 1.0) -> np.ndarray:
        """latents (B,20,T) (tensor or ndarray) -> host float32 waveforms (B, 512*T)."""
        if isinstance(z, np.ndarray):
            z = torch.from_numpy(z)
        return self.decode_tensor(z, scale_factor).cpu().numpy()

    __call__ = decode

    def decode_sharded(self, z_all, rank: int, world: int, scale_factor: float = 1.0):
        """Batch sharding: this rank decodes its contiguous slice of the batch."""
        s, e = shard_range(z_all.shape[0], rank, world)
        if e == s:
            return np.zeros((0, z_all.shape[-1] * self.vae.up_factor * self.voc.hop), np.float32)
        return self.decode(z_all[s:e], scale_factor)

    def launches(self, B, T):
        return self.vae.launches(B, T) + self.voc.launches(B, T * self.vae.up_factor) - 1  # mel stays packed

    def profile(self, B, T, iters=3):
        """Per-kernel-class device time (CUDA events around every kernel, eager launches) plus the
        algorithmic FLOPs / bytes of each class for one decode of shape (B, T)."""
        import ctypes as C
        prof = _lib.Profile()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.load().alcm_profile_decode(self.vae._h, self.voc._h, B, T, iters, C.byref(prof), stream))
        out = {}
        for i, name in enumerate(_lib.CLASSES):
            out[name] = dict(ms=prof.ms[i] / iters, flops=prof.flops[i], bytes=prof.bytes[i], launches=prof.launches[i])
        return out

    STAGES = ("vae", "conv_pre", "stage1", "stage2", "stage3", "stage4", "stage5", "stage6", "post")

    def profile_stages(self, B, T, iters=3):
        """The same measurement binned by pipeline stage: {stage: {class: {ms, flops, bytes, launches}}}."""
        import ctypes as C
        n = len(self.STAGES)
        arr = (_lib.Profile * n)()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.load().alcm_profile_stages(self.vae._h, self.voc._h, B, T, iters, arr, n, stream))
        out = {}
        for s, sname in enumerate(self.STAGES):
            out[sname] = {name: dict(ms=arr[s].ms[i] / iters, flops=arr[s].flops[i], bytes=arr[s].bytes[i], launches=arr[s].launches[i])
                          for i, name in enumerate(_lib.CLASSES) if arr[s].launches[i]}
        return out
