"""Deterministic synthetic workloads for tests, smoke() and bench.py.

numpy only; no dependency on the CUDA library or on oracle/.  The shapes mirror BASELINE.json's configs:

* ``tone_noise_audio``   the reference's own mel fixture, Tests/FluidAudioTests/Diarizer/Sortformer/
  SortformerStreamingMelTests.swift:17-25: 0.3 sin(2 pi 220 t) + 0.15 sin(2 pi 517 t) + 0.05 (u - 0.5),
  computed in float32.  The noise term is mandatory: without broadband energy most mel bins sit at the float32
  FFT noise floor and no two float32 implementations agree to 1e-4 in the log domain.  The Swift test draws
  ``u`` from drand48 seeded with 7; here ``u`` comes from a counter-based hash (any length, any offset, no state)
  so that a 1-hour signal can be generated in chunks on any rank.
* ``speaker_embeddings``  K unit-norm speaker centres + isotropic noise, stored float32 (TimedEmbedding.embedding256).
* ``synthetic_plda``      a stand-in for the PLDA rho transform, whose real weights live in a CoreML model that is
  not part of the reference repository ("parity unpinned"): rho = (unit(e) - mean) W, psi decaying.
"""
from __future__ import annotations

import numpy as np


def _hash_uniform(idx: np.ndarray, seed: int) -> np.ndarray:
    """Counter-based U[0,1) in float64 from 64-bit indices (splitmix64 finaliser)."""
    z = idx.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(seed + 1)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def tone_noise_audio(n: int, seed: int = 7, start: int = 0, sample_rate: int = 16000) -> np.ndarray:
    """float32 samples [start, start+n) of the tone+noise signal."""
    with np.errstate(over="ignore"):
        idx = np.arange(start, start + n, dtype=np.int64)
        t = (idx.astype(np.float32) / np.float32(sample_rate)).astype(np.float32)
        two_pi = np.float32(2.0) * np.float32(np.pi)
        tone = np.float32(0.3) * np.sin(two_pi * np.float32(220.0) * t, dtype=np.float32) + np.float32(0.15) * np.sin(
            two_pi * np.float32(517.0) * t, dtype=np.float32)
        u = _hash_uniform(idx, seed)
        noise = ((u - 0.5).astype(np.float32)) * np.float32(0.05)
        return (tone + noise).astype(np.float32)


def speech_like_audio(n: int, seed: int = 11, sample_rate: int = 16000) -> np.ndarray:
    """Harder mel fixture: amplitude-modulated harmonics with pauses + low-level noise (60 dB dynamic range)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sample_rate
    f0 = 110.0 + 40.0 * np.sin(2 * np.pi * 0.7 * t)
    phase = 2 * np.pi * np.cumsum(f0) / sample_rate
    sig = sum((0.5 / k) * np.sin(k * phase + 0.3 * k) for k in range(1, 12))
    env = np.clip(np.sin(2 * np.pi * 1.3 * t), 0, None) ** 2
    sig = sig * env * 0.4 + 1e-3 * rng.standard_normal(n)
    return sig.astype(np.float32)


def speaker_embeddings(n: int, dim: int = 256, speakers: int = 8, weights=None, sigma: float = 0.02,
                       seed: int = 42) -> tuple[np.ndarray, np.ndarray]:
    """Returns (embeddings float32 [n x dim], true speaker id int32 [n])."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((speakers, dim))
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    if weights is None:
        weights = np.arange(speakers, 0, -1, dtype=np.float64)
    w = np.asarray(weights, np.float64)
    w = w / w.sum()
    who = rng.choice(speakers, size=n, p=w).astype(np.int32)
    emb = centres[who] + sigma * rng.standard_normal((n, dim))
    return emb.astype(np.float32), who


def synthetic_plda(emb: np.ndarray, rho_dim: int = 128, seed: int = 1234) -> tuple[np.ndarray, np.ndarray]:
    """Returns (rho float64 [n x rho_dim], psi float64 [rho_dim]) from float32 embeddings."""
    e = emb.astype(np.float64)
    dim = e.shape[1]
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((dim, dim)))
    W = q[:, :rho_dim] * np.sqrt(float(rho_dim))
    unit = e / np.maximum(np.linalg.norm(e, axis=1, keepdims=True), 1e-30)
    rho = (unit - unit.mean(axis=0, keepdims=True)) @ W
    psi = (0.1 + 10.0 * np.exp(-np.arange(rho_dim) / 32.0)).astype(np.float32).astype(np.float64)
    return np.ascontiguousarray(rho), psi
