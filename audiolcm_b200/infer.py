"""Batched driver that replaces the serial per-prompt loop of the reference's inference scripts.

Reference: /root/reference/pythonscripts/InferAPI.py:63-101 (``GenSamples.gen_test_sample``) run once per prompt by
``AudioLCMBatchInfer`` (:135-163): sample one latent -> ``decode_first_stage`` -> per item ``.cpu().numpy()`` ->
``vocoder.vocode`` (host -> device -> host) -> ``soundfile.write(wav_path, wav, 16000)``.

Here a whole batch of prompts goes through the sampler at once, the latents are decoded in chunks by
``LatentToWaveform`` (the mel never leaves the device), the waveform is packed to 16-bit PCM by the conv_post kernel
itself, copied to pinned host memory asynchronously (the copy of chunk i overlaps the decode of chunk i+1) and written
as RIFF/WAVE files - the same bytes ``soundfile.write`` produces for a float waveform (PCM_16, mono, 16 kHz).

The sampler / denoiser is NOT part of this package (BASELINE.json configs[4]: "denoiser left as reference PyTorch"):
pass any callable ``sample_fn(cond) -> latents (B,20,T)``, e.g. ``lambda c: sampler.sample(S=2, conditioning=c,
batch_size=len(c), shape=[20, 312], guidance_scale=5, ...)[0]`` with the reference's ``LCMSampler``.
"""
from __future__ import annotations

import os
import wave

import numpy as np
import torch

SAMPLE_RATE = 16000


def write_wav_pcm16(path, pcm: np.ndarray, sample_rate: int = SAMPLE_RATE):
    """int16 mono samples -> RIFF/WAVE file (44-byte header + little-endian PCM)."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sample_rate)
        f.writeframes(pcm.tobytes())


class GenSamplesBatched(object):
    """``GenSamples`` (InferAPI.py:48-101) for a batch of prompts.

    ``pipe``       LatentToWaveform (VAE decoder + vocoder on one GPU)
    ``sample_fn``  cond (B, L, D) -> latents (B, 20, T) on ``pipe.device`` (reference sampler + denoiser)
    ``get_learned_conditioning``  prompts (list of dict) -> cond tensor; optional (pass cond directly otherwise)
    """

    def __init__(self, sample_fn, pipe, outpath, get_learned_conditioning=None, save_mel=False, save_wav=True, scale_factor=1.0,
                 chunk=64):
        self.sample_fn, self.pipe, self.outpath = sample_fn, pipe, outpath
        self.get_learned_conditioning = get_learned_conditioning
        self.save_mel, self.save_wav = save_mel, save_wav
        self.scale_factor, self.chunk = float(scale_factor), int(chunk)
        self._pinned = {}
        self._copy_stream = None

    def _host_buffer(self, slot, shape, dtype):
        key = (slot, tuple(shape), dtype)
        if key not in self._pinned:
            self._pinned[key] = torch.empty(shape, dtype=dtype).pin_memory()
        return self._pinned[key]

    @torch.no_grad()
    def decode_latents(self, z):
        """latents (B,20,T) on the device -> host int16 array (B, 512*T) [+ host mel], chunked and double-buffered."""
        dev = self.pipe.device
        B = z.shape[0]
        L = z.shape[-1] * self.pipe.vae.up_factor * self.pipe.voc.hop
        out = np.empty((B, L), np.int16)
        mels = [] if self.save_mel else None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        pending = []
        for i, s in enumerate(range(0, B, self.chunk)):
            zc = z[s:s + self.chunk]
            if self.save_mel:
                wavf, mel = self.pipe.decode_tensor(zc, self.scale_factor, return_mel=True)
                pcm = torch.round(wavf * 32767.0).to(torch.int16)
                mels.append(mel.cpu().numpy())
            else:
                pcm = self.pipe.decode_pcm16_tensor(zc, self.scale_factor)
            ready = torch.cuda.Event()
            ready.record(main)
            host = self._host_buffer(i % 2, pcm.shape, torch.int16)
            if len(pending) >= 2:                      # the buffer of chunk i-2 must have been drained
                ev, hb, s0 = pending.pop(0)
                ev.synchronize()
                out[s0:s0 + hb.shape[0]] = hb.numpy()
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ready)
                host.copy_(pcm, non_blocking=True)
                pcm.record_stream(self._copy_stream)
                done = torch.cuda.Event()
                done.record(self._copy_stream)
            pending.append((done, host, s))
        for ev, hb, s0 in pending:
            ev.synchronize()
            out[s0:s0 + hb.shape[0]] = hb.numpy()
        return out, (np.concatenate(mels, axis=0) if mels else None)

    @torch.no_grad()
    def gen_test_samples(self, prompts_or_cond, wav_names, mel_names=None):
        """Batched ``gen_test_sample``: returns one record dict per prompt ({'caption', 'audio_path'[, 'mel_path']})."""
        if torch.is_tensor(prompts_or_cond):
            cond, captions = prompts_or_cond, [""] * prompts_or_cond.shape[0]
        else:
            if self.get_learned_conditioning is None:
                raise ValueError("prompts given but no get_learned_conditioning")
            captions = [p["ori_caption"] if isinstance(p, dict) else str(p) for p in prompts_or_cond]
            cond = self.get_learned_conditioning(prompts_or_cond)
        if len(wav_names) != cond.shape[0]:
            raise ValueError("one wav name per prompt")
        z = self.sample_fn(cond.to(self.pipe.device))
        pcm, mel = self.decode_latents(z)
        os.makedirs(self.outpath, exist_ok=True)
        records = []
        for i, name in enumerate(wav_names):
            rec = {"caption": captions[i]}
            if self.save_mel:
                mel_path = os.path.join(self.outpath, (mel_names[i] if mel_names else name) + "_0.npy")
                np.save(mel_path, mel[i])
                rec["mel_path"] = mel_path
            if self.save_wav:
                wav_path = os.path.join(self.outpath, name + "_0.wav")
                write_wav_pcm16(wav_path, pcm[i])
                rec["audio_path"] = wav_path
            records.append(rec)
        return records


def audiolcm_batch_infer(ori_prompts, get_learned_conditioning, sample_fn, pipe, outdir="results/test", chunk=64):
    """``AudioLCMBatchInfer`` (InferAPI.py:135-163) with every prompt in ONE sampler batch and the batched decode."""
    prompts = [dict(ori_caption=p, struct_caption=f"<{p}& all>") for p in ori_prompts]
    names = [p["ori_caption"].strip().replace(" ", "-") for p in prompts]
    gen = GenSamplesBatched(sample_fn, pipe, outdir, get_learned_conditioning, save_mel=False, save_wav=True, chunk=chunk)
    return gen.gen_test_samples(prompts, names)
