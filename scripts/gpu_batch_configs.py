"""BASELINE configs[3] / configs[4] per-GPU shares: 64 clips x 30 s (mel batch) and 8 meetings x 5 000 (cluster batch)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from fluidaudio_b200 import _lib, synth, clustering as cl
from fluidaudio_b200.mel import AudioMelSpectrogram

# ---- C5 share: 8 meetings x 5000 x 256
sets = []
for m in range(8):
    emb, _ = synth.speaker_embeddings(5000, 256, 4, weights=(0.4, 0.3, 0.2, 0.1), sigma=0.02, seed=m)
    sets.append(emb)
emb = np.concatenate(sets)
rho, psi = synth.synthetic_plda(emb)
offs = np.arange(9, dtype=np.int64) * 5000
c = cl.OfflineClusterer(psi=psi)
for _ in range(2):
    labels, infos = c.cluster_batch(emb, rho, offs)
t0 = time.perf_counter()
labels, infos = c.cluster_batch(emb, rho, offs)
dt = time.perf_counter() - t0
print(f"C5 share: 8 x 5000 batch {dt*1e3:.1f} ms -> {40000/dt:.0f} emb/s; per-set ahc ms {[round(i['ms_ahc'],1) for i in infos]}")
t0 = time.perf_counter()
for m in range(8):
    r = c.cluster(emb[m*5000:(m+1)*5000], rho[m*5000:(m+1)*5000])
    assert np.array_equal(r.labels, labels[m*5000:(m+1)*5000])
dt1 = time.perf_counter() - t0
print(f"          one at a time {dt1*1e3:.1f} ms -> {40000/dt1:.0f} emb/s (labels identical)")

r = c.cluster(emb[:5000], rho[:5000])
print("          single-set stages:", {k: round(v, 2) if isinstance(v, float) else v for k, v in r.info.items()})
t0 = time.perf_counter(); r = c.cluster(emb[:5000], rho[:5000]); print(f"          single call wall {1e3*(time.perf_counter()-t0):.1f} ms")

# ---- C4 share: 64 clips x 480000 samples, packed in pinned host memory (what a loader would hand over)
mel = AudioMelSpectrogram(n_mels=80)
n_clip, count = 480_000, 64
pin_in = _lib.PinnedArray((count * n_clip,), np.float32)
for i in range(count):
    pin_in.array[i * n_clip:(i + 1) * n_clip] = synth.tone_noise_audio(n_clip, seed=i)
offsets = np.arange(count + 1, dtype=np.int64) * n_clip
T = mel.frame_count(n_clip)
pin_out = _lib.PinnedArray((count * T * 80,), np.float32)
for _ in range(2):
    outs = mel.compute_batch(None, packed_audio=pin_in.array, offsets=offsets, out=pin_out.array)
t0 = time.perf_counter()
for _ in range(10):
    outs = mel.compute_batch(None, packed_audio=pin_in.array, offsets=offsets, out=pin_out.array)
dt = (time.perf_counter() - t0) / 10
hours = count * 30 / 3600
print(f"C4 share: 64 x 30 s batch, pinned host buffers, e2e {dt*1e3:.2f} ms -> {hours/dt:.0f} audio-h/s; mel lengths {outs[2][:3]}")
d_a = _lib.DeviceBuffer(pin_in.array.nbytes + 64); d_a.upload(pin_in.array)
d_o = _lib.DeviceBuffer(pin_out.array.nbytes)
out_offsets = np.arange(count + 1, dtype=np.int64) * T * 80
for _ in range(3): mel.compute_batch_device(d_a, offsets, d_o, out_offsets)
mel.timer_start()
for _ in range(20): mel.compute_batch_device(d_a, offsets, d_o, out_offsets)
ms = mel.timer_stop_ms() / 20
print(f"          device-resident {ms:.4f} ms -> {hours/(ms*1e-3):.0f} audio-h/s")
