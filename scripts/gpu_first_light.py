"""First-light GPU check: parity vs oracle + rough timings.  Usage: python scripts/gpu_first_light.py [mel|ahc|pipe|all]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fluidaudio_b200 import synth, _lib
from fluidaudio_b200.mel import AudioMelSpectrogram, PaddingMode, LogFloorMode
from fluidaudio_b200 import clustering as cl
from oracle import oracle as O

what = sys.argv[1] if len(sys.argv) > 1 else "all"
print("devices", _lib.device_count(), flush=True)

def mel_part():
    for nm in (80, 128):
        m = AudioMelSpectrogram(n_mels=nm)
        cfg = O.mel_config(n_mels=nm)
        assert np.array_equal(m.get_hann_window(), O.hann_window())
        assert np.array_equal(m.get_filterbank(), O.mel_filterbank(512, nm))
        for n in (1, 159, 400, 4000, 16000 * 3 + 137, 16000 * 30):
            a = synth.tone_noise_audio(n)
            got, ml, nf = m.compute_flat_transposed(a)
            ref, rml, rnf = O.mel_flat_transposed(cfg, a)
            d = np.abs(got.reshape(nf, nm) - ref).max()
            print(f"mel nm={nm} n={n} T={ml}/{rml} nf={nf}/{rnf} max|d|={d:.3e}", flush=True)
        a = synth.tone_noise_audio(16000 * 5 + 77)
        got, ml, nf = m.compute_flat(a, last_audio_sample=0.25)
        ref, rml, rnf = O.mel_flat(cfg, a, last=0.25)
        print(f"computeFlat nm={nm} max|d|={np.abs(got.reshape(nm, nf) - ref).max():.3e}", flush=True)
        got, ml = m.compute(a)
        ref, rml = O.mel_legacy(cfg, a)
        print(f"compute legacy nm={nm} T={ml}/{rml} max|d|={np.abs(got[0] - ref).max():.3e}", flush=True)
        got, ml, nf = m.compute_flat_transposed(a, padding_mode=PaddingMode.pre_padded, expected_frame_count=None)
        ref, rml, rnf = O.mel_flat_transposed(cfg, a, padding_mode=1)
        print(f"prePadded nm={nm} T={ml}/{rml} max|d|={np.abs(got.reshape(nf, nm) - ref).max():.3e}", flush=True)
        sp = synth.speech_like_audio(16000 * 20)
        got, ml, nf = m.compute_flat_transposed(sp)
        ref, _, _ = O.mel_flat_transposed(cfg, sp)
        print(f"speech-like nm={nm} max|d|={np.abs(got.reshape(nf, nm) - ref).max():.3e}", flush=True)
    # batch
    m = AudioMelSpectrogram(n_mels=80)
    cfg = O.mel_config(n_mels=80)
    clips = [synth.tone_noise_audio(n, seed=i) for i, n in enumerate([480000, 1000, 33333, 480000, 7])]
    out, offs, ml, nf = m.compute_batch(clips)
    worst = 0
    for i, c in enumerate(clips):
        ref, rml, rnf = O.mel_flat_transposed(cfg, c)
        g = out[offs[i]:offs[i + 1]].reshape(-1, 80)
        worst = max(worst, np.abs(g - ref).max()); assert ml[i] == rml and nf[i] == rnf
    print(f"batch max|d|={worst:.3e}", flush=True)
    # timings: 1 hour
    n = 57_600_000
    a = synth.tone_noise_audio(n)
    pin = _lib.PinnedArray(n, np.float32); pin.array[:] = a
    T = m.frame_count(n)
    pout = _lib.PinnedArray(T * 80, np.float32)
    for rep in range(3):
        t = time.time(); got, ml, nf = m.compute_flat_transposed(pin.array, out=pout.array); dt = time.time() - t
        print(f"e2e 1h pinned: {dt*1e3:.2f} ms -> {1/dt:.1f} audio-h/s", flush=True)
    t = time.time(); got2, _, _ = m.compute_flat_transposed(a); dt = time.time() - t
    print(f"e2e 1h pageable: {dt*1e3:.2f} ms", flush=True)
    d_a = _lib.DeviceBuffer(n * 4 + 64); d_a.upload(a)
    d_o = _lib.DeviceBuffer(T * 80 * 4)
    L = _lib.load()
    import ctypes as C
    for rep in range(4):
        L.fa_timer_start()
        m.compute_device(d_a, n, d_o)
        ms = C.c_float(); L.fa_timer_stop_ms(C.byref(ms))
        print(f"kernel-only 1h: {ms.value:.3f} ms -> {1e3/ms.value:.0f} audio-h/s, {345600320/ms.value/1e6:.0f} GB/s", flush=True)
    dev = d_o.download((T, 80), np.float32)
    print("device path == host path:", np.array_equal(dev, pout.array.reshape(T, 80)), flush=True)
    t = time.time(); ref, _, _ = O.mel_flat_transposed(cfg, a[:16000 * 600]); dt = time.time() - t
    print(f"oracle 10 min: {dt:.2f} s; max|d| first 10 min = {np.abs(dev[:ref.shape[0]-3] - ref[:-3]).max():.3e}", flush=True)

def ahc_part():
    for N, K in ((2, 1), (3, 2), (100, 4), (1000, 8), (3000, 8)):
        emb, _ = synth.speaker_embeddings(N, 256, K, seed=N)
        x = O.l2_normalize_rows(emb.astype(np.float64))
        t = time.time(); st, z = cl.centroid_linkage(x); dt = time.time() - t
        t = time.time(); st2, z2 = O.centroid_linkage(x, use_ref=O.ref_available()); dt2 = time.time() - t
        ms = np.zeros(4, np.float32); _lib.load().fa_ahc_last_stage_ms(ms.ctypes.data)
        print("   stages", ms)
        print(f"AHC N={N} status={st}/{st2} bit-exact={np.array_equal(z, z2)} gpu={dt*1e3:.1f} ms cpu={dt2*1e3:.1f} ms", flush=True)
        if st != 0: print("   last error:", _lib.load().fa_last_error())
        if not np.array_equal(z, z2) and N <= 100: print(z[:5], z2[:5])
    rng = np.random.default_rng(0)
    base = rng.standard_normal((50, 8)); x = np.repeat(base, 4, axis=0)[rng.permutation(200)]
    x = O.l2_normalize_rows(x)
    print("dups bit-exact", np.array_equal(cl.centroid_linkage(x)[1], O.centroid_linkage(x, use_ref=O.ref_available())[1]), flush=True)
    g = np.array([[i, j, k] for i in range(6) for j in range(6) for k in range(6)], float) + 1
    print("lattice bit-exact", np.array_equal(cl.centroid_linkage(g)[1], O.centroid_linkage(g, use_ref=O.ref_available())[1]), flush=True)
    xx = x.copy(); xx[3, 2] = np.nan
    print("nan status", cl.centroid_linkage(xx)[0], flush=True)
    emb, _ = synth.speaker_embeddings(10000, 256, 8, seed=42)
    x = O.l2_normalize_rows(emb.astype(np.float64))
    for rep in range(2):
        t = time.time(); st, z = cl.centroid_linkage(x); dt = time.time() - t
        ms = np.zeros(4, np.float32); _lib.load().fa_ahc_last_stage_ms(ms.ctypes.data)
        print(f"AHC N=10000 gpu={dt*1e3:.1f} ms status={st} stages(init,heapify,merge,total)={ms}", flush=True)
    np.save("gpurun_out/z10000.npy", z)

def pipe_part():
    for N, K in ((500, 4), (2000, 8)):
        emb, _ = synth.speaker_embeddings(N, 256, K, seed=N + 1)
        emb[5, 3] = np.nan
        rho, psi = synth.synthetic_plda(np.nan_to_num(emb))
        r = cl.OfflineClusterer(psi=psi).cluster(emb, rho)
        o = O.diarize_cluster(emb, rho, psi, use_ref=O.ref_available())
        print(f"pipeline N={N}: labels equal={np.array_equal(r.labels, o.labels)} initial equal={np.array_equal(r.initial[o.training_indices], o.initial)} "
              f"K={r.info['centroid_count']}/{o.centroids.shape[0]} iters={r.info['vbx_iterations']}/{len(o.vbx.elbos)} "
              f"cent max|d|={np.abs(r.centroids - o.centroids).max() if r.centroids.shape == o.centroids.shape else 'shape'} info={r.info}", flush=True)
        v = cl.VBxClustering(psi=psi).refine(rho[o.training_indices], o.initial)
        print(f"  vbx gamma max|d|={np.abs(v.gamma - o.vbx.gamma).max():.3e} pi={np.abs(v.pi - o.vbx.pi).max():.3e} "
              f"elbo rel={np.abs((v.elbos - o.vbx.elbos) / o.vbx.elbos).max():.3e} hard equal={np.array_equal(v.hard_clusters, o.vbx.hard)}", flush=True)
    emb, _ = synth.speaker_embeddings(10000, 256, 8, seed=42)
    rho, psi = synth.synthetic_plda(emb)
    c = cl.OfflineClusterer(psi=psi)
    for rep in range(2):
        t = time.time(); r = c.cluster(emb, rho); dt = time.time() - t
        print(f"C3 pipeline N=10000: {dt*1e3:.1f} ms -> {10000/dt:.0f} emb/s info={r.info}", flush=True)

os.makedirs("gpurun_out", exist_ok=True)
if what in ("mel", "all"): mel_part()
if what in ("ahc", "all"): ahc_part()
if what in ("pipe", "all"): pipe_part()
print("DONE", flush=True)
