import sys, glob
import numpy as np
sys.path.insert(0, ".")
from fluidaudio_b200 import synth, clustering as cl
from oracle import oracle as O

def compare(tag, rho, psi, init):
    g = cl.VBxClustering(psi=psi).refine(rho, init)
    o = O.vbx_refine(rho, psi, init)
    n = min(g.elbos.size, o.elbos.size)
    de = np.abs(g.elbos[:n] - o.elbos[:n])
    first = int(np.argmax(de > 1e-9)) if (de > 1e-9).any() else -1
    print(tag, "standalone elbos", g.elbos[:8], "pi", g.pi[:8], "rho0", rho[0,:2], "psi0", psi[0])
    print(tag, "T,D,S", rho.shape, len(set(init.tolist())), "iters", g.elbos.size, o.elbos.size, "first elbo diff at", first,
          "max|dgamma|", np.abs(g.gamma - o.gamma).max() if g.gamma.shape == o.gamma.shape else "shape",
          "elbo0", g.elbos[0], o.elbos[0])
    if first >= 0:
        print("    gpu", g.elbos[max(0, first - 1):first + 3], "\n    ora", o.elbos[max(0, first - 1):first + 3])

for n in (5, 40):
    emb, _ = synth.speaker_embeddings(n, 64, 3, seed=3)
    rho, psi = synth.synthetic_plda(emb, 64)
    init = O.ahc_cluster(emb.astype(np.float64), 0.6)
    compare(f"clean tiny n={n}", rho, psi, init)
for f in sorted(glob.glob("scratch_cases/*.npz")):
    z = np.load(f)
    emb, rho, psi = z["emb"], z["rho"], z["psi"]
    if emb.shape[0] > 400: continue
    ok = np.isfinite(emb).all(axis=1)
    train = emb[ok].astype(np.float64)
    init = O.ahc_cluster(train, 0.6) if train.shape[0] >= 2 else np.zeros(train.shape[0], np.int32)
    compare(f.split("/")[-1], rho[ok], psi, init)
    got = cl.OfflineClusterer(psi=psi).cluster(emb, rho)
    ref = O.diarize_cluster(emb, rho, psi)
    print("    pipeline: labels equal", np.array_equal(got.labels, ref.labels), "info", {k: got.info[k] for k in ("training_count", "initial_clusters", "vbx_iterations", "centroid_count")},
          "oracle iters", ref.vbx.elbos.size, "K", ref.centroids.shape[0], "init equal", np.array_equal(got.initial[got.initial >= 0], ref.initial))
