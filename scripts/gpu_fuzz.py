"""Randomised parity sweep against the oracle: mel configurations / lengths / modes, clustering shapes and odd inputs.
Usage: python scripts/gpu_fuzz.py [seed] [mel_cases] [cluster_cases]"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from fluidaudio_b200 import synth, clustering as cl, _lib
from fluidaudio_b200.mel import AudioMelSpectrogram, PaddingMode, LogFloorMode, Precision
from fluidaudio_b200.audio_converter import AudioConverter
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
    n_fft = int(rng.choice([512, 512, 512, 256, 1024, 128]))
    hop = int(rng.choice([80, 128, 160, 200, 256, 320, 77, 161]))
    win = int(rng.choice([w for w in (100, 128, 200, 256, 320, 400, 480, 512, 800, 1024) if w <= n_fft]))
    f32 = bool(rng.integers(0, 2))
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
    nm = min(nm, n_fft // 2 + 1)
    kw = dict(sample_rate=16000, n_mels=nm, n_fft=n_fft, hop_length=hop, win_length=win, preemph=pre, pad_to=pad_to,
              log_floor=floor, window_periodic=periodic)
    m = AudioMelSpectrogram(log_floor_mode=LogFloorMode.clamped if clamped else LogFloorMode.additive,
                            precision=Precision.f32 if f32 else Precision.f64, **kw)
    # float32 transform: 1e-4 with pre-emphasis, unit scale and the standard floor (the documented envelope of the option).
    # Outside it the float32 noise floor (2^-24 of the frame's loudest bin - the reference's own vDSP arithmetic has it
    # too) is no longer hidden by the log floor: a 1e-10 floor or 30x over-scale audio without pre-emphasis show 1e-3.
    if not f32:
        tol = 1e-4
    elif pre > 0 and scale <= 1.0 and floor >= 2.0 ** -24:
        tol = 1e-4
    elif floor >= 2.0 ** -24 and scale <= 1.0:
        tol = 5e-4
    else:
        tol = 5e-3
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
        ok = (ml, nf) == (rml, rnf) and got.shape == ref.shape and (got.size == 0 or np.abs(got - ref).max() <= tol)
    except Exception as e:
        ok = False
        print("EXC", type(e).__name__, e)
    if not ok:
        bad += 1
        d = np.abs(got - ref).max() if got.shape == ref.shape and got.size else -1
        print("MEL MISMATCH", got.shape, ref.shape, dict(n_fft=n_fft, f32=f32, hop=hop, win=win, nm=nm, pre=pre, periodic=periodic, clamped=clamped, floor=floor,
                                   pad_to=pad_to, n=n, mode=mode, last=last, scale=scale, tm=tm), (ml, nf), (rml, rnf), d)
    m.close()
print(f"mel: {n_mel} cases, {bad} bad, {time.time()-t0:.1f} s", flush=True)

# converter stage and the fused PCM -> mel entry against the float64 evaluation of the documented filter / the oracle's linear path
vbad = 0
conv = AudioConverter()
for case in range(max(10, n_mel // 2)):
    rate = int(rng.choice([8000, 11025, 12000, 22050, 24000, 32000, 44100, 48000, 88200, 96000, 16000, 15999]))
    ch = int(rng.choice([1, 1, 2, 2, 3, 5]))
    frames = int(rng.choice([0, 1, 7, 160, 1000, 4410, 30001]))
    i16 = bool(rng.integers(0, 2))
    inter = bool(rng.integers(0, 2))
    x = (rng.standard_normal((ch, frames)) * 0.2).astype(np.float32)
    if i16:
        x = np.round(np.clip(x, -1, 1) * 32767).astype(np.int16)
    try:
        arg = np.ascontiguousarray(x.T) if inter else x
        got = conv.resample_buffer(arg, rate, interleaved=inter)
        mono = O.mixdown(x)
        if ch > 2:
            ref = O.linear_resample(x.astype(np.float32) / (32768.0 if i16 else 1.0), rate, 16000) if i16 else O.linear_resample(x, rate, 16000)
            ok = got.shape == ref.shape and (np.array_equal(got, ref) if not i16 else np.abs(got - ref).max() <= 1e-6 if got.size else True)
        elif rate == 16000:
            ok = got.shape == mono.shape and np.array_equal(got, mono)
        else:
            ref = O.sinc_resample(mono, rate, 16000)
            ok = got.shape == ref.shape and (got.size == 0 or np.abs(got - ref).max() <= (3e-5 if rate == 15999 else 4e-6))
    except Exception as e:
        ok = False
        print("EXC", type(e).__name__, e)
    if not ok:
        vbad += 1
        print("CONVERTER MISMATCH", dict(rate=rate, ch=ch, frames=frames, i16=i16, inter=inter), got.shape if 'got' in dir() else None)
print(f"converter: {max(10, n_mel // 2)} cases, {vbad} bad, {time.time()-t0:.1f} s", flush=True)
bad += vbad

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
        # labels must agree wherever the decision is not a rounding-level tie: VBx on duplicate / lattice inputs returns
        # IDENTICAL centroids (cosine 1.0 between them), and which of two equal scores wins depends on the last bit of the
        # centroid sums (same criterion as tests/test_gpu_parity.py::test_pipeline_odd_shapes_and_tiny_inputs)
        decided = np.ones(n, bool)
        if ref.centroids.shape[0] > 1 and chunk is None and not ref.was_adjusted:
            okr = np.isfinite(emb).all(axis=1)
            cn = ref.centroids / np.maximum(np.linalg.norm(ref.centroids, axis=1, keepdims=True), 1e-300)
            e64 = np.where(okr[:, None], emb, 0.0).astype(np.float64)
            sc = (e64 / np.maximum(np.linalg.norm(e64, axis=1, keepdims=True), 1e-300)) @ cn.T
            srt = np.sort(sc, axis=1)
            decided = okr & (srt[:, -1] - srt[:, -2] > 1e-9)
        ok = np.array_equal(got.labels[decided], ref.labels[decided])
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
