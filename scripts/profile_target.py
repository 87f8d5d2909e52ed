"""Small, fixed workloads for ncu captures (never a bench value).  Usage: profile_target.py mel|cluster [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fluidaudio_b200 import _lib, synth
what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if what in ("mel", "mel32"):
    from fluidaudio_b200.mel import AudioMelSpectrogram, Precision
    n = 57_600_000
    a = synth.tone_noise_audio(n)
    m = AudioMelSpectrogram(n_mels=80, precision=Precision.f32 if what == "mel32" else Precision.f64)
    d_a = _lib.DeviceBuffer(n * 4 + 64); d_a.upload(a)
    d_o = _lib.DeviceBuffer(360001 * 80 * 4)
    for _ in range(reps):
        m.compute_device(d_a, n, d_o)
    _lib.synchronize()
elif what == "pcm":
    # converter stage in the fused pipeline: 600 s of 48 kHz stereo int16 and of 44.1 kHz mono float32 -> mel
    from fluidaudio_b200.mel import AudioMelSpectrogram, Precision
    m = AudioMelSpectrogram(n_mels=80, precision=Precision.f32)
    st = np.stack([synth.tone_noise_audio(48000 * 600, sample_rate=48000), synth.tone_noise_audio(48000 * 600, seed=9, sample_rate=48000)])
    i16 = np.ascontiguousarray(np.round(st.T * 32767).astype(np.int16))
    mono = synth.tone_noise_audio(44100 * 600, sample_rate=44100)
    for _ in range(reps):
        m.compute_from_pcm(i16, 48000.0, interleaved=True)
        m.compute_from_pcm(mono, 44100.0)
    _lib.synchronize()
else:
    from fluidaudio_b200.clustering import OfflineClusterer
    emb, _ = synth.speaker_embeddings(10000, 256, 8, seed=42)
    rho, psi = synth.synthetic_plda(emb)
    c = OfflineClusterer(psi=psi)
    for _ in range(reps):
        r = c.cluster(emb, rho)
    print(r.info)
print("profile target done")
