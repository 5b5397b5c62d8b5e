"""audiolcm_b200 - B200-native (sm_100a) latent->waveform decode path for AudioLCM.

Public API (drop-ins for the two reference call sites):

* ``VocoderBigVGAN``        - ``vocoder.bigvgan.models.VocoderBigVGAN`` (``.vocode(spec)``)
* ``AutoencoderKLDecoder``  - ``AutoencoderKL.decode`` behind ``decode_first_stage``; ``install(model, ddconfig)``
* ``LatentToWaveform``      - both chained with the mel kept on the device; batch / time sharding helpers
* ``GenSamplesBatched`` / ``audiolcm_batch_infer`` - batched replacement of the per-prompt loop of
  ``pythonscripts/InferAPI.py`` (latents -> 16-bit PCM packed on the GPU -> WAV files)

Everything runs through csrc/libaudiolcm_b200.so (C-ABI: include/audiolcm_b200.h).  There is no
CPU, PyTorch-eager or Triton fallback: a missing library or a non-sm_100 device raises.
"""
from ._lib import AlcmError, LIB_PATH  # noqa: F401
from .vocoder import VocoderBigVGAN  # noqa: F401
from .autoencoder import AutoencoderKLDecoder, AutoencoderKLEncoder, install  # noqa: F401
from .pipeline import LatentToWaveform, shard_range, halo_frames  # noqa: F401
from .infer import GenSamplesBatched, audiolcm_batch_infer, write_wav_pcm16  # noqa: F401

__all__ = ["VocoderBigVGAN", "AutoencoderKLDecoder", "AutoencoderKLEncoder", "install", "LatentToWaveform", "shard_range", "halo_frames", "AlcmError",
           "GenSamplesBatched", "audiolcm_batch_infer", "write_wav_pcm16"]
