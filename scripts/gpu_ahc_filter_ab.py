"""A/B of the float32 filter's kernels (FA_AHC_FILTER_IMPL, read once per process): C3 single call stage times and the
C5 batch (64 meetings x 5 000 on the batch lanes).  Usage: python scripts/gpu_ahc_filter_ab.py (spawns one process per variant)."""
import ctypes as C, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child():
    import numpy as np
    from fluidaudio_b200 import _lib, synth
    from fluidaudio_b200.clustering import OfflineClusterer
    out = {}
    for n in (10000, 5000, 2500):
        e, _ = synth.speaker_embeddings(n, 256, 8 if n == 10000 else 4, sigma=0.02, seed=3)
        r, psi = synth.synthetic_plda(e, 128)
        c = OfflineClusterer(psi=psi)
        ahc, init = [], []
        st = (C.c_float * 4)()
        for _ in range(6):
            res = c.cluster(e, r)
            _lib.load().fa_ahc_last_stage_ms(st)
            ahc.append(res.info["ms_ahc"]); init.append((st[0], st[1], st[2]))
        out[f"n{n}"] = {"ahc_ms_min": min(ahc[1:]), "init_heap_merge_ms": [round(v, 3) for v in min(init[1:])]}
    M, N = 64, 5000
    embs, rhos = [], []
    for m in range(M):
        e, _ = synth.speaker_embeddings(N, 256, 4, weights=(0.4, 0.3, 0.2, 0.1), sigma=0.02, seed=m)
        r, psi = synth.synthetic_plda(e, 128)
        embs.append(e); rhos.append(r)
    pe = _lib.PinnedArray((M * N, 256), np.float32); pe.array[:] = np.concatenate(embs)
    pr = _lib.PinnedArray((M * N, 128), np.float64); pr.array[:] = np.concatenate(rhos)
    offs = np.arange(M + 1, dtype=np.int64) * N
    c = OfflineClusterer(psi=psi)
    c.cluster_batch(pe.array, pr.array, offs)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); _, infos = c.cluster_batch(pe.array, pr.array, offs); ts.append((time.perf_counter() - t0) * 1e3)
    a = [i["ms_ahc"] for i in infos]
    out["c5"] = {"ms": sorted(ts), "ahc_min": min(a), "ahc_max": max(a)}
    print(json.dumps(out))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for rep in range(2):
            for impl in ("0", "1", "2", "3"):
                env = dict(os.environ, FA_AHC_FILTER_IMPL=impl)
                r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=600)
                print("impl", impl, (r.stdout.strip().splitlines() or [r.stderr[-400:]])[-1], flush=True)
