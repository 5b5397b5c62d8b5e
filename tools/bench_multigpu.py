"""Multi-GPU measurements of BASELINE.json configs[2] and configs[3] (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_multigpu.py [--precision bf16] [--steps 3]

configs[2]: 64 x 10 s clips, batch-sharded (64/N clips per rank, no data-path collective), strong scaling.
configs[3]: one 5-minute clip (mel [1,80,18750]) vocoded time-sharded over the N ranks with one NCCL P2P
            halo exchange (34 frames per side, inside the timed region); rank 0 also vocodes the whole clip
            un-sharded and the gathered shards are compared with it.
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
Rank 0 prints one JSON line per config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audiolcm_b200 import synth  # noqa: E402
from audiolcm_b200.pipeline import shard_range, vocode_time_sharded, halo_frames  # noqa: E402
from bench import build_pipe, T_LAT, audio_seconds  # noqa: E402


def timed(fn, steps, world, device):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--frames", type=int, default=18750)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    pipe = build_pipe(args.precision, device)

    # ---- configs[2]: batch sharding ------------------------------------------------------------
    s, e = shard_range(args.clips, rank, world)
    z = torch.from_numpy(synth.synth_latent(e - s, T_LAT, seed=100 + rank)).to(device)
    ms, _ = timed(lambda: pipe.decode_tensor(z), args.steps, world, device)
    if rank == 0:
        print(json.dumps(dict(config="configs[2]: %d x 10 s clips, batch-sharded" % args.clips, n_gpus=world, precision=args.precision,
                              clips_per_rank=e - s, ms_per_step=round(ms, 3), scaling="strong",
                              value=round(audio_seconds(args.clips, T_LAT) / (ms * 1e-3), 1), unit="audio-s/s")), flush=True)
    del z
    torch.cuda.empty_cache()

    # ---- configs[3]: long-form clip, time-sharded vocoder ----------------------------------------
    T = args.frames
    mel_all = torch.from_numpy(synth.synth_mel(1, T, seed=7))
    s, e = shard_range(T, rank, world)
    chunk = mel_all[..., s:e].contiguous().to(device)
    hop = pipe.voc.hop
    ms, wav = timed(lambda: vocode_time_sharded(pipe.voc.vocode_tensor, chunk, rank, world, hop), args.steps, world, device)
    # gather the shards on rank 0 and compare with the un-sharded vocode of the whole clip
    err = None
    if world > 1:
        sizes = [(shard_range(T, r, world)[1] - shard_range(T, r, world)[0]) * hop for r in range(world)]
        if rank == 0:
            parts = [wav] + [torch.empty((1, n), dtype=torch.float32, device=device) for n in sizes[1:]]
            for r in range(1, world):
                dist.recv(parts[r], src=r)
            full = torch.cat(parts, dim=-1)
        else:
            dist.send(wav.contiguous(), dst=0)
    else:
        full = wav
    if rank == 0:
        ref = pipe.voc.vocode_tensor(mel_all.to(device))
        err = float((full - ref).abs().max())
        ms1 = None
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            pipe.voc.vocode_tensor(mel_all.to(device))
            e1.record()
            torch.cuda.synchronize()
            ms1 = round(e0.elapsed_time(e1), 3)
        print(json.dumps(dict(config="configs[3]: %.0f s clip (mel [1,80,%d]), vocoder time-sharded, %d-frame NCCL P2P halo" %
                                     (T * hop / 16000.0, T, halo_frames()), n_gpus=world, precision=args.precision,
                              ms_per_step=round(ms, 3), value=round(T * hop / 16000.0 / (ms * 1e-3), 1), unit="audio-s/s",
                              ms_unsharded_one_gpu=ms1, max_abs_vs_unsharded=err, abs_max=float(ref.abs().max()))), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
