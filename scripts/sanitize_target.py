"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from fluidaudio_b200 import synth, clustering as cl
import os
from fluidaudio_b200.mel import AudioMelSpectrogram, LSEENDMelFrontend, UnifiedMelExtractor, PaddingMode, Precision
from fluidaudio_b200.audio_converter import AudioConverter

a = synth.tone_noise_audio(16000 * 3 + 77)
for nm in (80, 128):
    m = AudioMelSpectrogram(n_mels=nm)
    m.compute_flat_transposed(a)
    m.compute_flat(a[:20001])
    m.compute_flat_transposed(a[:9000], padding_mode=PaddingMode.pre_padded)
    m.compute(a[:4000])
# round 2: float32-pair transform, any-nFFT kernel, converter stage (sinc / linear / mixdown), fused PCM -> mel
m = AudioMelSpectrogram(n_mels=80, precision=Precision.f32)
m.compute_flat_transposed(a)
m.compute_flat(a[:20001])
m.compute_flat_transposed(a[:161])
for kw in (dict(n_fft=256, win_length=200, hop_length=80, n_mels=23), dict(n_fft=1024, win_length=800, hop_length=321, n_mels=64)):
    g = AudioMelSpectrogram(**kw)
    g.compute_flat_transposed(a[:30000])
    g.compute_flat(a[:5000])
conv = AudioConverter()
t = np.arange(48000) / 48000.0
st = np.stack([np.sin(2 * np.pi * 440 * t), np.sin(2 * np.pi * 550 * t)]).astype(np.float32)
conv.resample(st[0], 48000)
conv.resample(st[0][:22050], 22050)
conv.resample_buffer(np.round(st * 32767).astype(np.int16), 44100)
conv.resample_buffer(np.stack([st[0], st[1], st[0], st[1]])[:, :8000], 8000)
conv.resample_buffer(st[:, :16000], 16000)
m.compute_from_pcm(np.ascontiguousarray(np.round(st.T * 32767).astype(np.int16)), 48000, interleaved=True)
m.compute_from_pcm(np.round(st[0] * 32767).astype(np.int16)[:16000], 16000)
UnifiedMelExtractor(24000).features(np.concatenate([a[:20000], np.zeros(4000, np.float32)]), 20000)
LSEENDMelFrontend().process(a[:16000])
for n in (2, 3, 50, 400):
    emb, _ = synth.speaker_embeddings(n, 256, 4, seed=n)
    rho, psi = synth.synthetic_plda(emb)
    chunk = (np.arange(n) // 2).astype(np.int32)
    cl.OfflineClusterer(psi=psi).cluster(emb, rho)
    cl.OfflineClusterer(psi=psi).cluster(emb, rho, chunk_indices=chunk)
    if n >= 50:
        cl.OfflineClusterer(cl.OfflineDiarizerConfig().with_speakers(exactly=6), psi=psi).cluster(emb, rho)
# hundreds of speakers: E-step reads alpha through L2; batch entry point: several sets side by side
rng = np.random.default_rng(0)
emb, _ = synth.speaker_embeddings(320, 256, 4, seed=5)
rho, psi = synth.synthetic_plda(emb)
init = np.concatenate([np.arange(260), rng.integers(0, 260, 60)]).astype(np.int32)
cl.VBxClustering(psi=psi).refine(rho, init)
emb, _ = synth.speaker_embeddings(900, 256, 4, seed=6)
rho, psi = synth.synthetic_plda(emb)
cl.OfflineClusterer(psi=psi).cluster_batch(emb, rho, np.array([0, 300, 600, 900], np.int64))
m = AudioMelSpectrogram(n_mels=80)
m.compute_batch([a[:30000], a[:1000], a[:48077]])
# the float32 filter of the AHC nearest-neighbour pass is on from N = 2048 (FA_AHC_FILTER_MIN_N lowers it for this run)
emb, _ = synth.speaker_embeddings(700, 64, 3, seed=8)
cl.centroid_linkage(emb.astype(np.float64))
print("sanitize target done")
