"""Host-side mirror of the offline clustering backend's Swift surface.

* ``AHCClustering.cluster``      Sources/FluidAudio/Diarizer/Offline/Clustering/AHCClustering.swift:20-67
* ``VBxClustering.refine``       Sources/FluidAudio/Diarizer/Offline/Clustering/VBxClustering.swift:41-165
* ``VBxOutput``                  Sources/FluidAudio/Diarizer/Offline/Core/OfflineDiarizerTypes.swift:629-702
* ``OfflineDiarizerConfig``      OfflineDiarizerTypes.swift:33-454 (only the clustering knobs)
* ``OfflineClusterer.cluster``   OfflineDiarizerManager.cluster(_:) lines 286-375 of
                                 Sources/FluidAudio/Diarizer/Offline/Core/OfflineDiarizerManager.swift
* ``centroid_linkage``           the reference's C symbol fastcluster_compute_centroid_linkage
                                 (Sources/FastClusterWrapper/include/FastClusterWrapper.h:34-40)

Every computation runs in libfluidaudio_b200.so on the GPU; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib


@dataclass
class ClusteringConfig:           # OfflineDiarizerConfig.Clustering.community
    threshold: float = 0.6
    warm_start_fa: float = 0.07
    warm_start_fb: float = 0.8
    num_speakers: int | None = None      # exact count, overrides min / max (OfflineDiarizerTypes.swift)
    min_speakers: int | None = None
    max_speakers: int | None = None


@dataclass
class VBxConfig:                  # OfflineDiarizerConfig.VBx.community
    max_iterations: int = 20
    convergence_tolerance: float = 1e-4


@dataclass
class OfflineDiarizerConfig:
    clustering: ClusteringConfig = field(default_factory=ClusteringConfig)
    vbx: VBxConfig = field(default_factory=VBxConfig)

    @property
    def clustering_threshold(self) -> float:
        return self.clustering.threshold

    def _c_vbx(self) -> _lib.VbxConfig:
        return _lib.VbxConfig(self.clustering.warm_start_fa, self.clustering.warm_start_fb, self.vbx.max_iterations,
                              self.vbx.convergence_tolerance, 7.0)

    def _c_cluster(self) -> _lib.ClusterConfig:
        opt = lambda v: _lib.NO_VALUE if v is None else int(v)
        c = self.clustering
        return _lib.ClusterConfig(c.threshold, self._c_vbx(), opt(c.num_speakers), opt(c.min_speakers),
                                  opt(c.max_speakers), 0)

    def with_speakers(self, min: int | None = None, max: int | None = None, exactly: int | None = None):
        """OfflineDiarizerConfig.withSpeakers(min:max:) / withSpeakers(exactly:) (OfflineDiarizerTypes.swift:731-):
        a copy with the constraints applied; `exactly` takes precedence, min/max clear a previous exact count."""
        import copy
        out = copy.deepcopy(self)
        if exactly is not None:
            out.clustering.num_speakers, out.clustering.min_speakers, out.clustering.max_speakers = exactly, None, None
        else:
            out.clustering.num_speakers, out.clustering.min_speakers, out.clustering.max_speakers = None, min, max
        return out


def centroid_linkage(normalized_rows: np.ndarray):
    """Calls the drop-in C symbol.  Returns (status, Z [(N-1) x 4])."""
    x = np.ascontiguousarray(normalized_rows, np.float64)
    n, d = x.shape
    z = np.zeros((max(n - 1, 0), 4), np.float64)
    zbuf = z if z.size else np.zeros(4, np.float64)
    st = _lib.load().fastcluster_compute_centroid_linkage(x.ctypes.data, n, d, zbuf.ctypes.data, z.size)
    return int(st), z


def l2_normalize_rows(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros_like(x)
    _lib.check(_lib.load().fa_l2_normalize_rows(x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data),
               "fa_l2_normalize_rows")
    return out


def dendrogram_cut(z: np.ndarray, count: int, threshold: float) -> np.ndarray:
    labels = np.zeros(count, np.int32)
    zz = np.ascontiguousarray(z, np.float64).reshape(-1)
    _lib.check(_lib.load().fa_dendrogram_cut(zz.ctypes.data if zz.size else None, count, float(threshold),
                                             labels.ctypes.data if count else None), "fa_dendrogram_cut")
    return labels


class AHCClustering:
    def cluster(self, embedding_features, threshold: float) -> np.ndarray:
        rows = list(embedding_features) if not isinstance(embedding_features, np.ndarray) else embedding_features
        count = len(rows)
        if count == 0:
            return np.zeros(0, np.int32)
        x = np.asarray(rows, np.float64)
        if x.ndim < 2 or x.shape[1] == 0:
            return np.zeros(count, np.int32)
        x = np.ascontiguousarray(x)
        labels = np.zeros(count, np.int32)
        _lib.check(_lib.load().fa_ahc_cluster(x.ctypes.data, count, x.shape[1], float(threshold), labels.ctypes.data),
                   "fa_ahc_cluster")
        return labels


@dataclass
class VBxOutput:
    gamma: np.ndarray
    pi: np.ndarray
    hard_clusters: np.ndarray
    num_clusters: int
    elbos: np.ndarray
    centroids: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))
    was_adjusted: bool = False

    active_cluster_epsilon = 1e-7

    @property
    def active_cluster_count(self) -> int:
        if self.pi.size == 0:
            return self.num_clusters
        return int((self.pi > self.active_cluster_epsilon).sum())

    @property
    def assigned_cluster_count(self) -> int:
        if self.gamma.size == 0:
            return self.active_cluster_count
        return int(np.unique(self.gamma.argmax(axis=1)).size)


class VBxClustering:
    def __init__(self, config: OfflineDiarizerConfig | None = None, psi: np.ndarray | None = None):
        self.config = config or OfflineDiarizerConfig()
        self.psi = None if psi is None else np.ascontiguousarray(psi, np.float64)

    def refine(self, rho_features, initial_clusters) -> VBxOutput:
        rho = np.asarray(rho_features, np.float64)
        if rho.size == 0 or rho.ndim < 2 or rho.shape[1] == 0:
            return VBxOutput(np.zeros((0, 0)), np.zeros(0), np.zeros(0, np.int32), 0, np.zeros(0))
        rho = np.ascontiguousarray(rho)
        T, D = rho.shape
        init = np.ascontiguousarray(initial_clusters, np.int32)
        S = max(1, len(set(init.tolist())))
        cfg = self.config._c_vbx()
        cap = max(cfg.max_iterations, 1)
        gamma = np.zeros((T, S), np.float64)
        pi = np.zeros(S, np.float64)
        elbos = np.zeros(cap, np.float64)
        hard = np.zeros(T, np.int32)
        its = C.c_int32()
        psi = self.psi
        _lib.check(_lib.load().fa_vbx_refine(rho.ctypes.data, T, D, _lib.ptr(psi), 0 if psi is None else psi.size,
                                             init.ctypes.data if init.size else None, S, C.byref(cfg),
                                             gamma.ctypes.data, pi.ctypes.data, elbos.ctypes.data, hard.ctypes.data,
                                             C.byref(its)), "fa_vbx_refine")
        return VBxOutput(gamma, pi, hard, S, elbos[: its.value].copy())


def compute_centroids(training_embeddings: np.ndarray, vbx: VBxOutput) -> np.ndarray:
    emb = np.ascontiguousarray(training_embeddings, np.float64)
    T, dim = emb.shape
    S = vbx.pi.size
    cents = np.zeros((S, dim), np.float64)
    k = C.c_int32()
    _lib.check(_lib.load().fa_compute_centroids(emb.ctypes.data, T, dim,
                                                np.ascontiguousarray(vbx.gamma).ctypes.data,
                                                np.ascontiguousarray(vbx.pi).ctypes.data, S, cents.ctypes.data,
                                                C.byref(k)), "fa_compute_centroids")
    return cents[: k.value].copy()


def assign_embeddings(embedding_features: np.ndarray, centroids: np.ndarray, want_scores: bool = False):
    emb = np.ascontiguousarray(embedding_features, np.float64)
    cen = np.ascontiguousarray(centroids, np.float64)
    N, dim = emb.shape
    K = cen.shape[0]
    labels = np.zeros(N, np.int32)
    scores = np.zeros((N, max(K, 1)), np.float64) if want_scores else None
    _lib.check(_lib.load().fa_assign_embeddings(emb.ctypes.data, N, dim, cen.ctypes.data if K else None, K,
                                                labels.ctypes.data, _lib.ptr(scores)), "fa_assign_embeddings")
    return (labels, scores) if want_scores else labels


@dataclass
class ClusterResult:
    labels: np.ndarray
    initial: np.ndarray
    centroids: np.ndarray
    info: dict


class OfflineClusterer:
    """The clustering phase of OfflineDiarizerManager.cluster(_:) (embeddings -> per-embedding speaker labels)."""

    def __init__(self, config: OfflineDiarizerConfig | None = None, psi: np.ndarray | None = None):
        self.config = config or OfflineDiarizerConfig()
        self.psi = None if psi is None else np.ascontiguousarray(psi, np.float64)

    def cluster(self, embedding256: np.ndarray, rho128: np.ndarray, max_centroids: int = 64,
                chunk_indices=None) -> ClusterResult:
        """chunk_indices (TimedEmbedding.chunkIndex per embedding) switches on the reference's default constrained
        assignment (labels -2 where a chunk has more local speakers than clusters)."""
        emb = np.ascontiguousarray(embedding256, np.float32)
        rho = np.ascontiguousarray(rho128, np.float64)
        N, E = emb.shape
        R = rho.shape[1]
        psi = self.psi if self.psi is not None and self.psi.size == R else None   # VBxClustering.swift:71-76: identity
        labels = np.zeros(N, np.int32)
        initial = np.zeros(N, np.int32)
        cents = np.zeros((max_centroids, E), np.float64)
        info = _lib.ClusterInfo()
        cfg = self.config._c_cluster()
        if chunk_indices is None:
            _lib.check(_lib.load().fa_diarize_cluster(emb.ctypes.data, rho.ctypes.data, N, E, R, _lib.ptr(psi),
                                                      C.byref(cfg), labels.ctypes.data, initial.ctypes.data,
                                                      cents.ctypes.data, max_centroids, C.byref(info)),
                       "fa_diarize_cluster")
        else:
            chunk = np.ascontiguousarray(chunk_indices, np.int32)
            _lib.check(_lib.load().fa_diarize_cluster_chunks(emb.ctypes.data, rho.ctypes.data, N, E, R,
                                                             _lib.ptr(psi), C.byref(cfg), chunk.ctypes.data,
                                                             labels.ctypes.data, initial.ctypes.data, cents.ctypes.data,
                                                             max_centroids, C.byref(info)), "fa_diarize_cluster_chunks")
        d = {f: getattr(info, f) for f, _ in _lib.ClusterInfo._fields_}
        return ClusterResult(labels, initial, cents[: min(info.centroid_count, max_centroids)].copy(), d)

    def cluster_batch(self, embedding256: np.ndarray, rho128: np.ndarray, set_offsets,
                      chunk_indices=None) -> tuple[np.ndarray, list]:
        """chunk_indices: TimedEmbedding.chunkIndex per row, numbered inside its own set -> the reference's default
        constrained assignment in every set (as `cluster(..., chunk_indices=...)` does for one)."""
        emb = np.ascontiguousarray(embedding256, np.float32)
        rho = np.ascontiguousarray(rho128, np.float64)
        offs = np.ascontiguousarray(set_offsets, np.int64)
        count = offs.size - 1
        labels = np.zeros(emb.shape[0], np.int32)
        infos = (_lib.ClusterInfo * max(count, 1))()
        cfg = self.config._c_cluster()
        psi = self.psi if self.psi is not None and self.psi.size == rho.shape[1] else None
        if chunk_indices is None:
            _lib.check(_lib.load().fa_diarize_cluster_batch(emb.ctypes.data, rho.ctypes.data, offs.ctypes.data, count,
                                                            emb.shape[1], rho.shape[1], _lib.ptr(psi), C.byref(cfg),
                                                            labels.ctypes.data, infos), "fa_diarize_cluster_batch")
        else:
            chunk = np.ascontiguousarray(chunk_indices, np.int32)
            if chunk.shape[0] != emb.shape[0]:
                raise ValueError("chunk_indices needs one entry per embedding")
            _lib.check(_lib.load().fa_diarize_cluster_batch_chunks(emb.ctypes.data, rho.ctypes.data, offs.ctypes.data,
                                                                   count, emb.shape[1], rho.shape[1], _lib.ptr(psi),
                                                                   C.byref(cfg), chunk.ctypes.data, labels.ctypes.data,
                                                                   infos), "fa_diarize_cluster_batch_chunks")
        out = [{f: getattr(infos[i], f) for f, _ in _lib.ClusterInfo._fields_} for i in range(count)]
        return labels, out


# ---- HungarianAssignment / ConstrainedClusterAssignment (Sources/FluidAudio/Diarizer/HungarianAssignment.swift,
#      Diarizer/Offline/Clustering/ConstrainedClusterAssignment.swift) ---------------------------------------------
class HungarianAssignment:
    @staticmethod
    def solve(cost_square, n: int) -> list[int]:
        if n == 0:
            return []
        cost = np.ascontiguousarray(cost_square, np.int64).reshape(-1)
        out = np.zeros(n, np.int32)
        _lib.check(_lib.load().fa_hungarian_solve(cost.ctypes.data, n, out.ctypes.data), "fa_hungarian_solve")
        return out.tolist()

    @staticmethod
    def max_score_assignment(scores) -> list[int]:
        rows = len(scores)
        if rows == 0:
            return []
        cols = len(scores[0])
        s = np.ascontiguousarray(scores, np.float64).reshape(rows, cols) if cols else np.zeros((rows, 0))
        out = np.zeros(rows, np.int32)
        _lib.check(_lib.load().fa_max_score_assignment(s.ctypes.data if s.size else None, rows, cols, out.ctypes.data),
                   "fa_max_score_assignment")
        return out.tolist()


class ConstrainedClusterAssignment:
    @staticmethod
    def assign(scores, chunk_indices) -> list[int]:
        n = len(chunk_indices)
        if n == 0:
            return []
        s = np.ascontiguousarray(scores, np.float64)
        k = s.shape[1] if s.ndim == 2 else 0
        chunk = np.ascontiguousarray(chunk_indices, np.int32)
        out = np.zeros(n, np.int32)
        _lib.check(_lib.load().fa_constrained_assign(s.ctypes.data if s.size else None, n, k, chunk.ctypes.data,
                                                     out.ctypes.data), "fa_constrained_assign")
        return out.tolist()


def build_chunk_assignments(chunk_indices, speaker_indices, assignments, num_chunks: int, num_speakers: int,
                            cluster_count: int) -> np.ndarray:
    """OfflineDiarizerManager.buildChunkAssignments (:885-911): [numChunks x numSpeakers], -2 where unassigned."""
    chunk = np.ascontiguousarray(chunk_indices, np.int32)
    spk = np.ascontiguousarray(speaker_indices, np.int32)
    asg = np.ascontiguousarray(assignments, np.int32)
    m = np.zeros((num_chunks, num_speakers), np.int32)
    _lib.check(_lib.load().fa_build_chunk_assignments(chunk.ctypes.data, spk.ctypes.data, asg.ctypes.data, chunk.size,
                                                      num_chunks, num_speakers, cluster_count, m.ctypes.data),
               "fa_build_chunk_assignments")
    return m


# ---- speaker-count constraints + K-Means re-clustering (SpeakerCountConstraints.swift, KMeansClustering.swift) -----
@dataclass
class SpeakerCountConstraints:
    num_speakers: int | None
    min_speakers: int
    max_speakers: int

    @staticmethod
    def resolve(num_embeddings: int, num_speakers=None, min_speakers=None, max_speakers=None) -> "SpeakerCountConstraints":
        opt = lambda v: _lib.NO_VALUE if v is None else int(v)
        lo, hi = C.c_int64(), C.c_int64()
        _lib.check(_lib.load().fa_speaker_constraints_resolve(int(num_embeddings), opt(num_speakers), opt(min_speakers),
                                                              opt(max_speakers), C.byref(lo), C.byref(hi)),
                   "fa_speaker_constraints_resolve")
        num = lo.value if lo.value == hi.value else num_speakers
        return SpeakerCountConstraints(num, lo.value, hi.value)

    def needs_adjustment(self, detected_count: int) -> bool:
        return detected_count < self.min_speakers or detected_count > self.max_speakers

    def target_count(self, detected_count: int) -> int:
        return min(max(detected_count, self.min_speakers), self.max_speakers)


class KMeansClustering:
    """KMeansClustering.swift:39-130; all arithmetic on the GPU (`fa_kmeans_cluster`)."""

    @staticmethod
    def cluster_with_centroids_n_init(embeddings, num_clusters: int, max_iterations: int = 300, n_init: int = 10,
                                      base_seed: int = 0):
        emb = np.ascontiguousarray(embeddings, np.float64)
        if emb.ndim != 2 or emb.shape[0] == 0:
            return np.zeros(0, np.int32), np.zeros((0, 0), np.float64), 0
        n, d = emb.shape
        cap = max(1, min(int(num_clusters), n))
        labels = np.zeros(n, np.int32)
        cents = np.zeros((cap, d), np.float64)
        rows, best = C.c_int32(), C.c_int32()
        _lib.check(_lib.load().fa_kmeans_cluster(emb.ctypes.data, n, d, int(num_clusters), int(max_iterations),
                                                 int(n_init), int(base_seed), labels.ctypes.data, cents.ctypes.data, cap,
                                                 C.byref(rows), C.byref(best)), "fa_kmeans_cluster")
        return labels, cents[: rows.value].copy(), best.value

    @staticmethod
    def cluster_with_centroids(embeddings, num_clusters: int, max_iterations: int = 300, seed: int | None = None):
        labels, cents, _ = KMeansClustering.cluster_with_centroids_n_init(embeddings, num_clusters, max_iterations, 1,
                                                                          seed or 0)
        return labels, cents

    @staticmethod
    def cluster(embeddings, num_clusters: int, max_iterations: int = 300, seed: int | None = None):
        return KMeansClustering.cluster_with_centroids(embeddings, num_clusters, max_iterations, seed)[0]


# ---- timeline reconstruction (Diarizer/Offline/Utils/OfflineReconstruction.swift) -----------------------------------
@dataclass
class TimedSpeakerSegment:            # Diarizer/Core/DiarizerTypes.swift:191-213 (embedding = centroid of `cluster`)
    speaker_id: str
    cluster: int
    start_time_seconds: float
    end_time_seconds: float
    quality_score: float


class OfflineReconstruction:
    """buildSegments (:24-253) without the optional zero-vote re-embed pass; host code inside the library."""

    def __init__(self, frame_duration: float, window_duration: float = 10.0, min_gap_duration: float = 0.1,
                 segmentation_min_duration_off: float = 0.0, segmentation_min_duration_on: float = 0.0,
                 min_segment_duration: float = 1.0, exclusive_segments: bool = True):
        self.cfg = _lib.ReconstructConfig(frame_duration, window_duration, min_gap_duration, segmentation_min_duration_off,
                                          segmentation_min_duration_on, min_segment_duration, int(exclusive_segments), 0)

    def build_segments(self, speaker_weights, hard_clusters, centroid_count: int, chunk_offsets=None) -> list:
        w = np.ascontiguousarray(speaker_weights, np.float32)
        chunks, frames, speakers = (w.shape if w.ndim == 3 else (0, 0, 0))
        hard = np.ascontiguousarray(hard_clusters, np.int32).reshape(-1, max(speakers, 1)) if np.size(hard_clusters) else \
            np.zeros((0, max(speakers, 1)), np.int32)
        offs = np.ascontiguousarray(chunk_offsets if chunk_offsets is not None else [], np.float64)
        cap = max(16, chunks * max(speakers, 1) * 4)
        while True:
            cl, st, en, q = (np.zeros(cap, np.int32), np.zeros(cap, np.float32), np.zeros(cap, np.float32),
                             np.zeros(cap, np.float32))
            n = C.c_int32()
            status = _lib.load().fa_build_segments(w.ctypes.data if w.size else None, chunks, frames, speakers,
                                                   offs.ctypes.data if offs.size else None, offs.size,
                                                   hard.ctypes.data if hard.size else None, hard.shape[0], int(centroid_count),
                                                   C.byref(self.cfg), cl.ctypes.data, st.ctypes.data, en.ctypes.data,
                                                   q.ctypes.data, cap, C.byref(n))
            if status == 3 and n.value > cap:      # FA_STATUS_OUTPUT_TOO_SMALL: retry with the reported size
                cap = n.value
                continue
            _lib.check(status, "fa_build_segments")
            break
        return [TimedSpeakerSegment(f"S{int(cl[i]) + 1}", int(cl[i]), float(st[i]), float(en[i]), float(q[i]))
                for i in range(n.value)]

    @staticmethod
    def build_speaker_database(segments, centroids) -> dict:
        """buildSpeakerDatabase (:296-357): {"S<k+1>": float32 mean embedding} for the speakers that own a segment."""
        cen = np.ascontiguousarray(centroids, np.float64)
        K, dim = cen.shape if cen.ndim == 2 else (0, 0)
        cl = np.ascontiguousarray([s.cluster for s in segments], np.int32)
        db = np.zeros((K, dim), np.float32)
        counts = np.zeros(max(K, 1), np.int32)
        _lib.check(_lib.load().fa_build_speaker_database(cl.ctypes.data if cl.size else None, cl.size,
                                                         cen.ctypes.data if cen.size else None, K, dim,
                                                         db.ctypes.data if db.size else None, counts.ctypes.data),
                   "fa_build_speaker_database")
        return {f"S{k + 1}": db[k].copy() for k in range(K) if counts[k] > 0}
