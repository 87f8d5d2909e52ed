"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from fluidaudio_b200 import synth, clustering as cl
from fluidaudio_b200.mel import AudioMelSpectrogram, LSEENDMelFrontend, UnifiedMelExtractor, PaddingMode

a = synth.tone_noise_audio(16000 * 3 + 77)
for nm in (80, 128):
    m = AudioMelSpectrogram(n_mels=nm)
    m.compute_flat_transposed(a)
    m.compute_flat(a[:20001])
    m.compute_flat_transposed(a[:9000], padding_mode=PaddingMode.pre_padded)
    m.compute(a[:4000])
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
print("sanitize target done")
