"""Golden labels for BASELINE configs[4] (64 meetings x 5 000 x 256, seed = meeting index) and configs[2].

Run HERE (the container that has /root/reference and therefore oracle/_ref/liboracle_fc.so):

    python tests/golden/make_c5_golden.py [workers]

Every meeting goes through the UNMODIFIED reference FastClusterWrapper.cpp (centroid linkage) and the oracle
restatement of the Swift stages either side of it (normalise, cut, VBx, centroids, argmax) — the same function the GPU
tests use as their checker.  Stored per meeting: SHA-256 of the reference dendrogram bytes, of the AHC labels and of the
final int32 labels, plus the cluster count.  bench.py compares every rank's labels against these hashes after the NCCL
gather (`labels_equal_ref`), so the multi-GPU lines carry a parity field that does not need the oracle at run time.
"""
import hashlib
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

MEETINGS, N, D, K = 64, 5000, 256, 4
WEIGHTS = (0.4, 0.3, 0.2, 0.1)


def one(m: int) -> dict:
    from fluidaudio_b200 import synth
    from oracle import oracle as O
    emb, _ = synth.speaker_embeddings(N, D, K, weights=WEIGHTS, sigma=0.02, seed=m)
    rho, psi = synth.synthetic_plda(emb)
    x = O.l2_normalize_rows(emb.astype(np.float64))
    st, z = O.centroid_linkage(x, use_ref=True)
    assert st == 0
    ahc = O.dendrogram_cut(z, N, 0.6)
    pipe = O.diarize_cluster(emb, rho, psi, use_ref=True)
    return {"meeting": m, "z_sha256": hashlib.sha256(z.tobytes()).hexdigest(),
            "ahc_labels_sha256": hashlib.sha256(np.ascontiguousarray(ahc, np.int32).tobytes()).hexdigest(),
            "final_labels_sha256": hashlib.sha256(np.ascontiguousarray(pipe.labels, np.int32).tobytes()).hexdigest(),
            "ahc_clusters": int(ahc.max() + 1), "centroids": int(pipe.centroids.shape[0]),
            "vbx_iterations": int(len(pipe.vbx.elbos))}


def main():
    from oracle import oracle as O
    O.build()
    assert O.ref_available(), "oracle/_ref/liboracle_fc.so missing: run `make -C oracle ref` where /root/reference exists"
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else max(1, (os.cpu_count() or 2) - 1)
    with ProcessPoolExecutor(workers) as ex:
        rows = list(ex.map(one, range(MEETINGS)))
    out = {"config": {"meetings": MEETINGS, "n": N, "dim": D, "speakers": K, "weights": WEIGHTS, "sigma": 0.02,
                      "seed": "meeting index", "rho": "synth.synthetic_plda(emb) per meeting", "threshold": 0.6,
                      "linkage": "unmodified reference FastClusterWrapper.cpp (oracle/_ref)"},
           "meetings": rows}
    with open(os.path.join(HERE, "c5_meetings.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("c5_meetings.json written:", len(rows), "meetings;", sum(r["centroids"] for r in rows), "centroids in total")


if __name__ == "__main__":
    main()
