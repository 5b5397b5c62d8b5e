#!/usr/bin/env python
"""Headline benchmark: audio-seconds per second of latent->waveform decode (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|tf32|fp32] [--batch B]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

Workload (BASELINE.json configs[1]): full latent->waveform decode - autoencoder1d VAE decoder +
BigVGAN-16k - of one 10 s clip (z [B,20,312] -> wav [B,159744]), batch 1 per GPU, random-init
weights of the shipped architectures (audiolcm_b200/synth.py), synthetic latents.  A step = one decode.
N>1 (torchrun): every rank decodes its own clip, no data-path collective (weak scaling).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA events around each step on
the launching stream, L2 flushed between steps, max over ranks.  `e2e`: the public API
(`LatentToWaveform.decode`) with a pinned host latent in and a host waveform out every step.
`roofline`: dominant kernel class (tcgen05 conv GEMMs) - algorithmic FLOPs / CUDA-event time per
launch measured in this process (eager launches with an event pair per kernel), against the
measured peaks in MEASURED_PEAKS.json.  `cpu_baseline`: the oracle port of the reference's CPU path
on this box's host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_LAT = 312                      # configs/audiolcm.yaml:13 mel_length -> 10 s clip
SR, HOP, VAE_UP = 16000, 256, 2
METRIC = "audio_seconds_per_second_latent_to_waveform_decode"


def audio_seconds(B, t_lat):
    return B * t_lat * VAE_UP * HOP / SR


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------- CPU arm
def cpu_decode_fn(threads):
    import torch
    from oracle import decode_oracle as O
    from audiolcm_b200 import synth
    torch.set_num_threads(threads)
    dd, h = synth.vae_config(), synth.bigvgan_config()
    vsd = {k: torch.from_numpy(v) for k, v in synth.vae_decoder_state_dict(dd, seed=3).items()}
    gsd = {k: torch.from_numpy(v) for k, v in synth.bigvgan_state_dict(h, seed=0).items()}

    def run(z):
        with torch.no_grad():
            mel = O.decode_first_stage(vsd, dd, z)
            return O.bigvgan_forward(gsd, h, mel)
    return run


def cpu_baseline(budget_s=25.0):
    """Oracle port of the reference CPU decode on all host cores: full 10 s clip, best of <=2 runs
    (bounded to ~budget_s of CPU work)."""
    from audiolcm_b200 import synth
    import torch
    cores = os.cpu_count() or 1
    run = cpu_decode_fn(cores)
    run(torch.from_numpy(synth.synth_latent(1, 8, seed=1)))      # touch the code paths / thread pool
    z = torch.from_numpy(synth.synth_latent(1, T_LAT, seed=0))
    times = []
    t_start = time.perf_counter()
    for _ in range(2):
        t0 = time.perf_counter()
        run(z)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    return dict(value=round(audio_seconds(1, T_LAT) / min(times), 4), unit="audio-s/s", cores=cores, kind="port",
                sample=f"{len(times)} full decode(s) of the same 10 s clip (VAE+BigVGAN-16k, fp32, batch 1), best; "
                       f"oracle/decode_oracle.py = same ATen CPU ops as the reference modules")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is pure
    Python/PyTorch under /root/reference (absent on the GPU box, nothing to compile into
    oracle/_ref), so this times the oracle port (same torch CPU ops, pinned to the reference by
    tests/golden) with all host threads.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from audiolcm_b200 import synth
    cores = os.cpu_count() or 1
    run = cpu_decode_fn(cores)
    # bound the whole run to a few minutes: calibrate on a 1 s clip, then pick the clip length
    zc = torch.from_numpy(synth.synth_latent(1, 32, seed=1))
    run(zc)
    t0 = time.perf_counter()
    run(zc)
    per_lat = (time.perf_counter() - t0) / 32
    n_runs = args.steps + args.warmup
    t_lat = T_LAT
    while t_lat > 32 and per_lat * t_lat * n_runs > 200.0:
        t_lat //= 2
    z = torch.from_numpy(synth.synth_latent(1, t_lat, seed=0))
    for _ in range(args.warmup):
        run(z)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(z)
    dt = time.perf_counter() - t0
    val = audio_seconds(1, t_lat) * args.steps / dt
    sample = (f"each step = full decode of a {audio_seconds(1, t_lat):.2f} s clip (z [1,20,{t_lat}]), fp32, batch 1, "
              f"{cores} host threads; oracle port of the reference modules (reference is Python, cannot travel)")
    line = dict(metric=METRIC, value=round(val, 4), unit="audio-s/s", impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=round(1e3 * dt / args.steps, 3), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"latent->waveform decode (VAE decoder + BigVGAN-16k), batch 1, "
                                     f"{audio_seconds(1, t_lat):.2f} s clip, CPU"),
                cpu_baseline=dict(value=round(val, 4), unit="audio-s/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=round(val, 4), unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


# --------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        hi = sorted(sm)[len(sm) // 2:]  # upper half = samples taken under load
        return dict(sm_mhz=statistics.median(hi), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def kernel_rooflines(precision, peaks):
    """The two kernel classes at sizes that fill the GPU (the batch-1 decode cannot): CUDA-event time of
    back-to-back launches inside the library (alcm_bench_conv / alcm_bench_act), algorithmic work / time."""
    import ctypes as C
    from audiolcm_b200 import _lib
    lib, ctx, prec = _lib.load(), _lib.ctx(0), _lib.PREC[precision]
    out = {}
    ms = C.c_float()
    B, Cc, T, K = 8, 768, 2500, 11          # stage-1 AMP conv of the 10 s clip, batch 8
    _lib.check(lib.alcm_bench_conv(ctx, B, Cc, Cc, T, K, 1, prec, 20, 0, C.byref(ms)))
    tf = 2.0 * B * Cc * Cc * K * T / (ms.value * 1e-3) / 1e12
    out["conv_gemm"] = dict(kernel="conv_umma_kernel", shape=f"Conv1d {Cc}->{Cc} k{K}, T={T}, batch {B}", bound="tensor",
                            achieved=round(tf, 1), peak=peaks["tf_sustained"], unit="TFLOP/s", frac=round(tf / peaks["tf_sustained"], 4),
                            us_per_launch=round(ms.value * 1e3, 1))
    B, Cc, T = 64, 24, 160000               # last-stage Activation1d, batch 64: 2-2.6 GB, far beyond L2
    for key, p, osz in (("activation1d", precision, 2 if precision == "bf16" else 4), ("activation1d_fp32_out", "tf32", 4)):
        _lib.check(lib.alcm_bench_act(ctx, B, Cc, T, _lib.PREC[p], 10, C.byref(ms)))
        byt = B * ((Cc + 15) // 16 * 16) * T * (4 + osz)
        gbs = byt / (ms.value * 1e-3) / 1e9
        out[key] = dict(kernel="act1d_kernel", shape=f"C={Cc} T={T} batch {B}, fp32 in / {'bf16' if osz == 2 else 'fp32'} out "
                                                      f"({byt / 1e6:.0f} MB algorithmic)", bound="hbm",
                        achieved=round(gbs, 1), peak=peaks["hbm"], unit="GB/s", frac=round(gbs / peaks["hbm"], 4),
                        us_per_launch=round(ms.value * 1e3, 1))
    return out


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(p))["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def build_pipe(precision, device):
    from audiolcm_b200 import AutoencoderKLDecoder, LatentToWaveform, VocoderBigVGAN
    from audiolcm_b200 import synth  # seeded synthetic weights/inputs only (data generator, not the checker)
    dd, h = synth.vae_config(), synth.bigvgan_config()
    vae = AutoencoderKLDecoder(synth.vae_decoder_state_dict(dd, seed=3), dd, synth.VAE_EMBED_DIM, device, precision)
    voc = VocoderBigVGAN.from_state_dict(synth.bigvgan_state_dict(h, seed=0), h, device, precision)
    return LatentToWaveform(vae, voc)


def time_steps(pipe, z_dev, steps, warmup, flush):
    import torch
    for _ in range(warmup):
        pipe.decode_tensor(z_dev)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        flush.zero_()                       # > L2 (126 MB): every step starts cold in L2
        e0.record()
        pipe.decode_tensor(z_dev)
        e1.record()
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for e0, e1 in evs) / 1e3  # seconds of device time over `steps`


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from audiolcm_b200 import synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: audiolcm_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    B = args.batch
    pipe = build_pipe(args.precision, device)
    z_host = torch.from_numpy(synth.synth_latent(B, T_LAT, seed=rank)).pin_memory()
    z_dev = z_host.to(device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    pipe.decode_tensor(z_dev)               # plan + CUDA graph
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    dev_s = time_steps(pipe, z_dev, args.steps, args.warmup, flush)
    barrier()
    clocks = sampler.stop() if sampler else None
    # end to end through the public API: pinned host latent in, host waveform out, every step
    barrier()
    for _ in range(max(1, args.warmup)):
        pipe.decode(z_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wav = pipe.decode(z_host)
    e2e_s = time.perf_counter() - t0
    barrier()
    if world > 1:
        t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        peaks = measured_peaks()
        asec = audio_seconds(B, T_LAT)
        value = world * asec * args.steps / dev_s
        prof = pipe.profile(B, T_LAT, iters=3)
        conv, act = prof["conv"], prof["act"]
        tf = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
        act_gbs = act["bytes"] / (act["ms"] * 1e-3) / 1e9
        total_ms = sum(c["ms"] for c in prof.values())
        launches = pipe.launches(B, T_LAT)
        line = dict(
            metric=METRIC, value=round(value, 2), unit="audio-s/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=round(1e3 * dev_s / args.steps, 4), higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype={"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.precision], data="synthetic",
            config=dict(workload=f"configs[1]: full latent->waveform decode (autoencoder1d VAE decoder + BigVGAN-16k), batch {B} "
                                 f"per GPU, 10 s clip (z [{B},20,{T_LAT}] -> wav [{B},{T_LAT * VAE_UP * HOP}])",
                        precision=args.precision, l2="flushed between timed steps (256 MiB write)", graph=True,
                        weights="random-init (seeded), shipped architectures"),
            e2e=dict(value=round(world * asec * args.steps / e2e_s, 2), unit="audio-s/s", h2d_bytes_per_step=int(z_host.numel() * 4),
                     d2h_bytes_per_step=int(wav.size * 4), api="LatentToWaveform.decode(pinned host latent) -> host ndarray"),
            gpu_launches=launches * args.steps,
            clocks=clocks,
            roofline=dict(bound="tensor", achieved=round(tf, 2), peak=peaks["tf_sustained"], unit="TFLOP/s",
                          frac=round(tf / peaks["tf_sustained"], 4), traffic=ncu_traffic(), kernel="conv_umma_kernel (all conv GEMM launches)",
                          peak_source=f"{peaks['source']} bf16 sustained", flops_per_step=conv["flops"],
                          ms_per_step=round(conv["ms"], 4), launches=conv["launches"],
                          flops_per_launch=round(conv["flops"] / max(conv["launches"], 1)),
                          us_per_launch=round(1e3 * conv["ms"] / max(conv["launches"], 1), 2),
                          note="per-launch averages over the conv GEMM launches of one batch-1 decode; `traffic` = ncu DRAM bytes per "
                               "launch (profiles/r1_traffic.json); the same kernel at a GPU-filling size is in kernel_rooflines"),
            roofline_act=dict(bound="hbm", achieved=round(act_gbs, 1), peak=peaks["hbm"], unit="GB/s",
                              frac=round(act_gbs / peaks["hbm"], 4), kernel="act1d_kernel (all Activation1d launches)",
                              bytes_per_step=act["bytes"], ms_per_step=round(act["ms"], 4), launches=act["launches"],
                              note="batch-1 tensors (<=15 MB) are L2-resident; see DESIGN.md for the HBM-sized run"),
            kernel_rooflines=kernel_rooflines(args.precision, peaks) if args.precision != "fp32" else None,
            class_ms={k: round(v["ms"], 4) for k, v in prof.items()},
            class_ms_total_eager=round(total_ms, 4),
        )
        if world == 1 and not args.no_batch64:
            # BASELINE.json configs[2] on this GPU (64 x 10 s clips in one call) - context for the batch-1 headline
            zb = torch.from_numpy(synth.synth_latent(64, T_LAT, seed=11)).to(device)
            pipe.decode_tensor(zb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            e0.record()
            for _ in range(3):
                pipe.decode_tensor(zb)
            e1.record()
            torch.cuda.synchronize()
            ms64 = e0.elapsed_time(e1) / 3
            line["configs2_batch64"] = dict(value=round(audio_seconds(64, T_LAT) / (ms64 * 1e-3), 1), unit="audio-s/s", ms_per_step=round(ms64, 2),
                                            workload="configs[2]: 64 x 10 s clips in one call on one GPU (inputs in HBM, 3 steps)")
            del zb
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def protect_stdout():
    """The contract is ONE JSON line on stdout: libraries that print to fd 1 (NCCL's version banner under
    torchrun) are sent to stderr; emit() writes the line to the real stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("ALCM_BENCH_PRECISION", "bf16"), choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-batch64", action="store_true", help="skip the configs[2] (batch 64) context measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
