"""Randomised parity sweep against the oracle: mel configurations / lengths / modes, clustering shapes and odd inputs.
Usage: python scripts/gpu_fuzz.py [seed] [mel_cases] [cluster_cases]"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from fluidaudio_b200 import synth, clustering as cl, _lib
from fluidaudio_b200.mel import AudioMelSpectrogram, PaddingMode, LogFloorMode
from oracle import oracle as O

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_mel = int(sys.argv[2]) if len(sys.argv) > 2 else 60
n_cl = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rng = np.random.default_rng(seed)
os.makedirs("gpurun_out", exist_ok=True)
bad = 0
t0 = time.time()
base = synth.speech_like_audio(16000 * 8, seed=seed) if hasattr(synth, "speech_like_audio") else synth.tone_noise_audio(16000 * 8)
for case in range(n_mel):
    hop = int(rng.choice([80, 128, 160, 200, 256, 320]))
    win = int(rng.choice([200, 256, 320, 400, 480, 512]))
    nm = int(rng.choice([23, 40, 64, 80, 128, 257]))
    pre = float(rng.choice([0.0, 0.97, 0.5]))
    periodic = bool(rng.integers(0, 2))
    clamped = bool(rng.integers(0, 2))
    floor = float(rng.choice([2.0 ** -24, 1e-10, 1e-5]))
    pad_to = int(rng.choice([0, 1, 8, 16]))
    n = int(rng.choice([0, 1, 100, 399, 400, 511, 512, 513, 1000, 5000, 16000, 40001, 127999]))
    mode = int(rng.integers(0, 2))
    last = float(rng.choice([0.0, 0.25, -0.5]))
    scale = float(rng.choice([1.0, 1e-3, 30.0]))
    a = (base[:n] * scale).astype(np.float32)
    kw = dict(sample_rate=16000, n_mels=nm, n_fft=512, hop_length=hop, win_length=win, preemph=pre, pad_to=pad_to,
              log_floor=floor, window_periodic=periodic)
    m = AudioMelSpectrogram(log_floor_mode=LogFloorMode.clamped if clamped else LogFloorMode.additive, **kw)
    cfg = O.mel_config(log_floor_mode=int(clamped), **kw)
    tm = bool(rng.integers(0, 2))
    try:
        if tm:
            got, ml, nf = m.compute_flat_transposed(a, last_audio_sample=last, padding_mode=PaddingMode(mode))
            ref, rml, rnf = O.mel_flat_transposed(cfg, a, last, mode, None)
        else:
            if mode == 1:
                continue
            got, ml, nf = m.compute_flat(a, last_audio_sample=last)
            ref, rml, rnf = O.mel_flat(cfg, a, last)
        got, ref = np.asarray(got).ravel(), np.asarray(ref).ravel()
        ok = (ml, nf) == (rml, rnf) and got.shape == ref.shape and (got.size == 0 or np.abs(got - ref).max() <= 1e-4)
    except Exception as e:
        ok = False
        print("EXC", type(e).__name__, e)
    if not ok:
        bad += 1
        d = np.abs(got - ref).max() if got.shape == ref.shape and got.size else -1
        print("MEL MISMATCH", got.shape, ref.shape, dict(hop=hop, win=win, nm=nm, pre=pre, periodic=periodic, clamped=clamped, floor=floor,
                                   pad_to=pad_to, n=n, mode=mode, last=last, scale=scale, tm=tm), (ml, nf), (rml, rnf), d)
    m.close()
print(f"mel: {n_mel} cases, {bad} bad, {time.time()-t0:.1f} s", flush=True)

cbad = 0
for case in range(n_cl):
    n = int(rng.choice([1, 2, 3, 5, 17, 100, 333, 1000, 2500]))
    d = int(rng.choice([16, 64, 192, 255, 256, 257]))
    r = int(rng.choice([32, 127, 128]))
    k = int(rng.integers(1, 9))
    emb, _ = synth.speaker_embeddings(n, d, k, seed=int(rng.integers(0, 1 << 30)), sigma=float(rng.choice([0.02, 0.2])))
    kind = rng.integers(0, 5)
    if kind == 1 and n > 4:
        emb[rng.integers(0, n, n // 3)] = emb[0]                    # exact duplicates
    if kind == 2 and n > 3:
        emb[rng.integers(0, n)] = np.nan                            # filtered rows
        emb[rng.integers(0, n)] = np.inf
    if kind == 3:
        emb = np.round(emb * 4) / 4                                 # lattice: many exact ties
    if kind == 4 and n > 2:
        emb[rng.integers(0, n)] = 0.0                               # zero-norm row
    rho, psi = synth.synthetic_plda(np.nan_to_num(emb, nan=0.0, posinf=0.0, neginf=0.0), r)
    chunk = np.sort(rng.integers(0, max(1, n // 2), n)).astype(np.int32) if rng.integers(0, 2) else None
    spk = {}
    if rng.integers(0, 3) == 0:
        spk = dict(num_speakers=int(rng.integers(1, 7)))
    cfg = cl.OfflineDiarizerConfig()
    if spk:
        cfg = cfg.with_speakers(exactly=spk["num_speakers"])
    try:
        got = cl.OfflineClusterer(cfg, psi=psi).cluster(emb, rho, chunk_indices=chunk)
        ref = O.diarize_cluster(emb, rho, psi, use_ref=O.ref_available(), chunk_indices=chunk, **spk)
        ok = np.array_equal(got.labels, ref.labels)
        diag = ""
        if not ok:
            gi = got.initial[got.initial >= 0] if got.initial.size == n else got.initial
            diag = (f"init_equal={np.array_equal(gi, ref.initial)} S={got.info['initial_clusters']}/{len(set(ref.initial.tolist()))} "
                    f"K={got.info['centroid_count']}/{ref.centroids.shape[0]} it={got.info['vbx_iterations']}/{ref.vbx.elbos.size} "
                    f"ndiff={(got.labels != ref.labels).sum()} adj={got.info['was_adjusted']}/{ref.was_adjusted} "
                    f"det={got.info['detected_clusters']}/{ref.detected_clusters} "
                    f"cent_maxdiff={np.abs(got.centroids - ref.centroids).max() if got.centroids.shape == ref.centroids.shape else 'shape'}")
    except Exception as e:
        ok = False
        diag = f"EXC {type(e).__name__} {e}"
    if not ok:
        cbad += 1
        print("CLUSTER MISMATCH", dict(n=n, d=d, r=r, k=k, kind=int(kind), chunks=chunk is not None, **spk), diag)
        np.savez(f"gpurun_out/fuzz_cluster_{seed}_{case}.npz", emb=emb, rho=rho, psi=psi, chunk=chunk if chunk is not None else np.zeros(0), got=got.labels if 'got' in dir() and hasattr(got, "labels") else np.zeros(0), ref=ref.labels if 'ref' in dir() and hasattr(ref, "labels") else np.zeros(0))
print(f"cluster: {n_cl} cases, {cbad} bad, {time.time()-t0:.1f} s")
sys.exit(1 if (bad or cbad) else 0)
