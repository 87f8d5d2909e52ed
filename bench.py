#!/usr/bin/env python
"""bench.py — the reference's headline metric on its named configuration, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mel|cluster] [--impl ours|reference]

Workload `mel` (default; BASELINE.json configs[1], the configuration the metric is quoted on):
    log-mel of 1 h of synthetic 16 kHz mono audio, 25 ms / 10 ms frames, 512-point FFT, 80 mels -> [360 001 x 80].
    A step = one pass over the hour.  `value` = audio-hours/s with audio and output resident in HBM (CUDA events on
    the launching stream); `e2e` = the same through fa_mel_compute with pinned HOST buffers (H2D and D2H inside).
Workload `cluster` (configs[2]): 10 000 x 256 embeddings -> normalise + AHC + cut + VBx + centroids + assignment.
    Its numbers are ALSO attached to the default line under "cluster" so that one run reports both halves of the metric.
With N > 1 (torchrun) every rank runs the same per-GPU workload on its own GPU: units are independent, there is no
data-path collective, scaling is weak; timing is barrier + device sync on both sides, MAX over ranks.

--impl reference: the reference's CPU implementation of the path timed on the host cores — the oracle port of
AudioMelSpectrogram (no Swift toolchain exists) for `mel`, the UNMODIFIED FastClusterWrapper.cpp (oracle/_ref, when it
was built) plus the oracle port of the Swift stages for `cluster`.  Rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MEL_SAMPLES = 57_600_000
MEL_FRAMES = 360_001
N_MELS = 80
MEL_BYTES_PER_HOUR = 4 * MEL_SAMPLES + 4 * MEL_FRAMES * N_MELS            # 345 600 320 B
CLUSTER_N, CLUSTER_D, CLUSTER_R, CLUSTER_K = 10_000, 256, 128, 8
AHC_BYTES = 8.0 * CLUSTER_D * CLUSTER_N * CLUSTER_N                       # 2.048e11 B


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_mel(audio: np.ndarray, threads: int, repeats: int = 1):
    """Oracle port of AudioMelSpectrogram on `threads` host threads: the audio cut into 30 s clips, every clip
    processed `repeats` times, each thread working through its own share of the clips with its own output buffer
    (one AudioMelSpectrogram-like instance per thread; ctypes releases the GIL inside the C++ call)."""
    import ctypes as C
    from oracle import oracle as O
    L = O.lib()
    L.oracle_tune_allocator()
    cfg = O.mel_config(n_mels=N_MELS)
    clip = 480_000
    pieces = [np.ascontiguousarray(audio[i:i + clip]) for i in range(0, audio.size, clip)] * repeats
    shares = [pieces[t::threads] for t in range(threads)]
    shares = [s for s in shares if s]
    O.mel_flat_transposed(cfg, pieces[0][:16000])
    ml0, nf0 = C.c_int64(), C.c_int64()
    cap = max(int(L.oracle_mel_compute_flat_transposed(C.byref(cfg), p.ctypes.data, p.size, 0.0, 0, -1, None, 0,
                                                       C.byref(ml0), C.byref(nf0))) for p in {p.size: p for p in pieces}.values())

    def work(share):
        out = np.empty(cap, np.float32)
        ml, nf = C.c_int64(), C.c_int64()
        frames = 0
        for p in share:
            L.oracle_mel_compute_flat_transposed(C.byref(cfg), p.ctypes.data, p.size, 0.0, 0, -1, out.ctypes.data, cap,
                                                 C.byref(ml), C.byref(nf))
            frames += ml.value
        return frames

    t0 = time.perf_counter()
    with ThreadPoolExecutor(len(shares)) as ex:
        list(ex.map(work, shares))
    dt = time.perf_counter() - t0
    return (repeats * audio.size / 16000.0 / 3600.0) / dt, dt


def cpu_cluster(emb, rho, psi):
    from oracle import oracle as O
    t0 = time.perf_counter()
    res = O.diarize_cluster(emb, rho, psi, use_ref=O.ref_available())
    dt = time.perf_counter() - t0
    return emb.shape[0] / dt, dt, ("reference" if O.ref_available() else "port"), res


def host_threads() -> int:
    return max(1, min(os.cpu_count() or 1, 64))


# ------------------------------------------------------------------------------------------------ GPU arms
def bench_mel(args, dist):
    from fluidaudio_b200 import _lib, sharding, synth
    from fluidaudio_b200.mel import AudioMelSpectrogram
    audio = synth.tone_noise_audio(MEL_SAMPLES, seed=7 + dist.rank)
    mel = AudioMelSpectrogram(n_mels=N_MELS)
    pin_in = _lib.PinnedArray(MEL_SAMPLES, np.float32)
    pin_in.array[:] = audio
    pin_out = _lib.PinnedArray(MEL_FRAMES * N_MELS, np.float32)
    d_in = _lib.DeviceBuffer(MEL_SAMPLES * 4 + 64)
    d_out = _lib.DeviceBuffer(MEL_FRAMES * N_MELS * 4)
    d_in.upload(audio)
    # ---- kernel-only: inputs resident in HBM (345.6 MB touched per step > 126 MB L2: nothing survives a step) ----
    clocks = ClockSampler(dist.local_rank)
    if dist.is_root:
        clocks.start()
    for _ in range(args.warmup):
        mel.compute_device(d_in, MEL_SAMPLES, d_out)
    _lib.synchronize()
    sharding.barrier(dist)
    launches0 = _lib.kernel_launch_count()
    mel.timer_start()
    for _ in range(args.steps):
        mel.compute_device(d_in, MEL_SAMPLES, d_out)
    dev_ms = mel.timer_stop_ms()
    _lib.synchronize()
    launches = _lib.kernel_launch_count() - launches0
    sharding.barrier(dist)
    dev_ms = sharding.all_reduce_max(dist, dev_ms)
    # ---- end to end through the C ABI with host buffers --------------------------------------------------------
    for _ in range(args.warmup):
        mel.compute_flat_transposed(pin_in.array, out=pin_out.array)
    sharding.barrier(dist)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, ml, nf = mel.compute_flat_transposed(pin_in.array, out=pin_out.array)
    _lib.synchronize()
    e2e_s = time.perf_counter() - t0
    sharding.barrier(dist)
    e2e_s = sharding.all_reduce_max(dist, e2e_s)
    clock_info = clocks.stop() if dist.is_root else None
    assert ml == MEL_FRAMES
    ms_per_step = dev_ms / args.steps
    hours = dist.world * 1.0
    peak, peak_src = measured_peaks()
    achieved = MEL_BYTES_PER_HOUR / (ms_per_step * 1e-3) / 1e9
    out = {
        "metric": "audio-hours/s", "value": hours / (ms_per_step * 1e-3), "unit": "audio-hours/s",
        "ms_per_step": ms_per_step, "dtype": "f32",
        "e2e": {"value": hours / (e2e_s / args.steps), "unit": "audio-hours/s", "ms_per_step": e2e_s / args.steps * 1e3,
                "h2d_bytes_per_step": 4 * MEL_SAMPLES, "d2h_bytes_per_step": 4 * MEL_FRAMES * N_MELS,
                "host_buffers": "pinned (fa_host_alloc)", "api": "fa_mel_compute"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": 317.5e6, "traffic_source": "ncu --set full, profiles/r01c_summary.txt: dram read 230.5 MB + "
                     "write 87.0 MB per launch (the tail of the output is still in L2 at kernel end)",
                     "kernel": "mel512_kernel", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": MEL_BYTES_PER_HOUR},
        "config": {"workload": "log-mel STFT, 1 h synthetic 16 kHz mono, 25 ms/10 ms frames, nFFT 512, 80 mels, per GPU",
                   "samples": MEL_SAMPLES, "frames": MEL_FRAMES, "n_mels": N_MELS,
                   "l2": "inputs+outputs 345.6 MB per step exceed the 126 MB L2 (no flush needed)",
                   "parallelism": f"dp{dist.world}: one process per GPU, independent clips, no data-path collective"},
        "clocks": clock_info,
    }
    return out, audio


def bench_cluster(args, dist, steps=None):
    from fluidaudio_b200 import _lib, sharding, synth
    from fluidaudio_b200.clustering import OfflineClusterer
    steps = steps or args.steps
    emb, _ = synth.speaker_embeddings(CLUSTER_N, CLUSTER_D, CLUSTER_K, sigma=0.02, seed=42 + dist.rank)
    rho, psi = synth.synthetic_plda(emb, CLUSTER_R)
    pin_e = _lib.PinnedArray(emb.shape, np.float32); pin_e.array[:] = emb
    pin_r = _lib.PinnedArray(rho.shape, np.float64); pin_r.array[:] = rho
    c = OfflineClusterer(psi=psi)
    for _ in range(max(1, min(args.warmup, 3))):
        res = c.cluster(pin_e.array, pin_r.array)
    sharding.barrier(dist)
    launches0 = _lib.kernel_launch_count()
    t0 = time.perf_counter()
    dev_ms, ahc_ms = 0.0, 0.0
    for _ in range(steps):
        res = c.cluster(pin_e.array, pin_r.array)
        i = res.info
        dev_ms += i["ms_normalize"] + i["ms_ahc"] + i["ms_cut"] + i["ms_vbx"] + i["ms_assign"]
        ahc_ms += i["ms_ahc"]
    e2e_s = time.perf_counter() - t0
    launches = _lib.kernel_launch_count() - launches0
    sharding.barrier(dist)
    e2e_s = sharding.all_reduce_max(dist, e2e_s)
    dev_ms = sharding.all_reduce_max(dist, dev_ms)
    peak, peak_src = measured_peaks()
    achieved = AHC_BYTES / (ahc_ms / steps * 1e-3) / 1e9
    n_total = dist.world * CLUSTER_N
    out = {
        "metric": "embeddings/s", "value": n_total / (dev_ms / steps * 1e-3), "unit": "embeddings/s",
        "ms_per_step": dev_ms / steps, "dtype": "f64",
        "e2e": {"value": n_total / (e2e_s / steps), "unit": "embeddings/s", "ms_per_step": e2e_s / steps * 1e3,
                "h2d_bytes_per_step": int(emb.nbytes + rho.nbytes), "d2h_bytes_per_step": int(4 * CLUSTER_N + 32 * (CLUSTER_N - 1)),
                "host_buffers": "pinned (fa_host_alloc)", "api": "fa_diarize_cluster"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": 42.3e6, "traffic_source": "ncu --set full, profiles/r01c_summary.txt: ahc_merge_kernel dram "
                     "read 41.3 MB + write 1.0 MB per launch (+ 20.5 MB read by ahc_init_nn_kernel)",
                     "kernel": "ahc_merge_kernel (+ ahc_init_nn_kernel)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": AHC_BYTES,
                     "note": "node vectors are resident in shared memory, so algorithmic bytes are served on-chip; "
                             "the loop is bound by N-1 dependent steps (latency), see DESIGN.md"},
        "stages_ms": {k: res.info[k] for k in ("ms_normalize", "ms_ahc", "ms_cut", "ms_vbx", "ms_assign", "ms_total")},
        "config": {"workload": "offline diarization backend: 10 000 x 256-d embeddings (8 speakers), cosine-normalise + "
                               "centroid AHC (thr 0.6) + cut + VBx (Fa 0.07, Fb 0.8, <=20 it) + centroids + argmax, per GPU",
                   "n": CLUSTER_N, "dim": CLUSTER_D, "rho_dim": CLUSTER_R,
                   "parallelism": f"dp{dist.world}: one process per GPU, independent embedding sets"},
    }
    return out, (emb, rho, psi, res)


def bench_batched_shares():
    """Per-GPU shares of BASELINE configs[3] and configs[4] (secondary numbers, N = 1 only): 64 clips x 30 s through
    fa_mel_compute_batch from pinned host memory, and 8 meetings x 5 000 x 256 through fa_diarize_cluster_batch
    (three meetings side by side on disjoint SM partitions)."""
    from fluidaudio_b200 import _lib, synth
    from fluidaudio_b200.clustering import OfflineClusterer
    from fluidaudio_b200.mel import AudioMelSpectrogram
    out = {}
    mel = AudioMelSpectrogram(n_mels=N_MELS)
    n_clip, count = 480_000, 64
    pin_in = _lib.PinnedArray((count * n_clip,), np.float32)
    for i in range(count):
        pin_in.array[i * n_clip:(i + 1) * n_clip] = synth.tone_noise_audio(n_clip, seed=i)
    offsets = np.arange(count + 1, dtype=np.int64) * n_clip
    T = mel.frame_count(n_clip)
    pin_out = _lib.PinnedArray((count * T * N_MELS,), np.float32)
    for _ in range(2):
        mel.compute_batch(None, packed_audio=pin_in.array, offsets=offsets, out=pin_out.array)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        mel.compute_batch(None, packed_audio=pin_in.array, offsets=offsets, out=pin_out.array)
    dt = (time.perf_counter() - t0) / reps
    out["mel_c4_share"] = {"workload": "64 clips x 30 s (configs[3] / 8 GPUs), pinned host buffers, fa_mel_compute_batch",
                           "e2e_ms": dt * 1e3, "e2e_audio_hours_per_s": count * 30 / 3600 / dt}
    sets = [synth.speaker_embeddings(5000, CLUSTER_D, 4, weights=(0.4, 0.3, 0.2, 0.1), sigma=0.02, seed=m)[0] for m in range(8)]
    emb = np.concatenate(sets)
    rho, psi = synth.synthetic_plda(emb, CLUSTER_R)
    offs = np.arange(9, dtype=np.int64) * 5000
    c = OfflineClusterer(psi=psi)
    c.cluster_batch(emb, rho, offs)
    t0 = time.perf_counter()
    labels, infos = c.cluster_batch(emb, rho, offs)
    dt = time.perf_counter() - t0
    out["cluster_c5_share"] = {"workload": "8 meetings x 5 000 x 256 (configs[4] / 8 GPUs), fa_diarize_cluster_batch",
                               "e2e_ms": dt * 1e3, "e2e_embeddings_per_s": emb.shape[0] / dt,
                               "ahc_ms_per_meeting": [round(i["ms_ahc"], 2) for i in infos]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=["mel", "cluster"], default="mel")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from fluidaudio_b200 import sharding, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = host_threads()
        steps = max(1, min(args.steps, 3))
        if args.workload == "mel":
            sample_s, repeats = 3600, 4                                          # 4 audio-hours (~20 s of CPU work) per step
            audio = synth.tone_noise_audio(16000 * sample_s)
            for _ in range(min(args.warmup, 1)):
                cpu_mel(audio[: 16000 * 120], threads)
            vals = [cpu_mel(audio, threads, repeats) for _ in range(steps)]
            v = float(np.mean([x[0] for x in vals])); dt = float(np.mean([x[1] for x in vals]))
            line = {"impl": "reference", "metric": "audio-hours/s", "value": v, "unit": "audio-hours/s", "dtype": "f32",
                    "config": {"workload": "log-mel STFT, 1 h synthetic 16 kHz mono, 25 ms/10 ms frames, nFFT 512, 80 mels"},
                    "cpu_baseline": {"value": v, "unit": "audio-hours/s", "cores": threads, "kind": "port",
                                     "sample": f"the workload's {sample_s} s of audio x {repeats} per step, 30 s clips over {threads} threads "
                                               "(oracle port of AudioMelSpectrogram.swift; no Swift toolchain)"}}
        else:
            emb, _ = synth.speaker_embeddings(CLUSTER_N, CLUSTER_D, CLUSTER_K, seed=42)
            rho, psi = synth.synthetic_plda(emb, CLUSTER_R)
            v, dt, kind, _ = cpu_cluster(emb, rho, psi)
            steps = 1
            line = {"impl": "reference", "metric": "embeddings/s", "value": v, "unit": "embeddings/s", "dtype": "f64",
                    "config": {"workload": "offline diarization backend: 10 000 x 256-d embeddings, AHC + VBx + assignment"},
                    "cpu_baseline": {"value": v, "unit": "embeddings/s", "cores": 1, "kind": kind,
                                     "sample": "the full 10 000 x 256 problem once; fastcluster is single-threaded as shipped"}}
        line.update({"n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                     "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
                     "e2e": {"value": v, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line))
        return 0

    from fluidaudio_b200 import _lib
    dist = sharding.init_distributed()
    if _lib.device_count() < 1:
        raise SystemExit("bench.py needs a B200: " + "no sm_100a device visible (there is no CPU fallback)")
    _lib.set_device(dist.local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = sharding.bind_to_gpu_numa(dist.local_rank)

    if args.workload == "mel":
        line, audio = bench_mel(args, dist)
        extra_steps = max(3, min(args.steps, 5))
        cluster_line, cluster_data = bench_cluster(args, dist, steps=extra_steps)
        line["cluster"] = {k: cluster_line[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline",
                                                         "stages_ms", "gpu_launches", "config")}
        line["cluster"]["steps"] = extra_steps
    else:
        line, cluster_data = bench_cluster(args, dist)
        audio = None

    if dist.is_root and world == 1 and args.workload == "mel" and not args.no_cpu_baseline:
        line["batched_shares"] = bench_batched_shares()
    os.sched_setaffinity(0, all_cpus)      # the CPU baseline may use every host core again
    if dist.is_root and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        if args.workload == "mel":
            repeats = 4
            v, dt = cpu_mel(audio, threads, repeats)
            line["cpu_baseline"] = {"value": v, "unit": "audio-hours/s", "cores": threads, "kind": "port",
                                    "sample": f"the workload's hour of audio x {repeats}, 30 s clips over {threads} host threads, "
                                              f"{dt:.2f} s wall (oracle port of AudioMelSpectrogram.swift)"}
            v1, dt1 = cpu_mel(audio[: 16000 * 300], 1)
            line["cpu_baseline"]["single_thread_value"] = v1
            emb, rho, psi, res = cluster_data
            cv, cdt, kind, ores = cpu_cluster(emb, rho, psi)
            line["cluster"]["cpu_baseline"] = {"value": cv, "unit": "embeddings/s", "cores": 1, "kind": kind,
                                               "sample": f"the full 10 000 x 256 problem once, {cdt:.1f} s "
                                                         "(fastcluster is single-threaded as shipped)"}
            line["cluster"]["labels_equal_cpu"] = bool(np.array_equal(res.labels, ores.labels))
        else:
            emb, rho, psi, res = cluster_data
            cv, cdt, kind, ores = cpu_cluster(emb, rho, psi)
            line["cpu_baseline"] = {"value": cv, "unit": "embeddings/s", "cores": 1, "kind": kind,
                                    "sample": f"the full 10 000 x 256 problem once, {cdt:.1f} s"}
            line["labels_equal_cpu"] = bool(np.array_equal(res.labels, ores.labels))

    line["host_binding"] = numa
    line.update({"n_gpus": dist.world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                 "scaling": "weak", "vs_baseline": None, "data": "synthetic"})
    if dist.is_root:
        print(json.dumps(line))
    sharding.finalize(dist)
    return 0


if __name__ == "__main__":
    sys.exit(main())
