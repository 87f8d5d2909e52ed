"""Consumers of goldens produced by the real reference on a Mac (swift/Tools/DumpGoldens.swift ->
tests/golden/swift_fixtures.py pack).  Absent here (no Swift toolchain, no Accelerate): every test then SKIPS with
"parity unpinned", which is the honest status DESIGN.md records.  When present they pin the oracle (CPU, here) and the
CUDA path (tests/test_gpu_parity.py::test_swift_goldens_when_present) to Apple's actual numbers."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"parity unpinned: {name} absent (needs swift/Tools/DumpGoldens.swift run on a Mac)")
    return np.load(path)


def test_oracle_mel_against_swift_values(oracle):
    g = _load("swift_mel.npz")
    assert np.array_equal(oracle.hann_window(400, False), g["hann_400"])
    assert np.abs(oracle.mel_filterbank(512, 80).reshape(-1) - g["filterbank_80"]).max() <= 1e-7
    for name in ("tone_noise", "speech_like"):
        a = g[f"audio_{name}"]
        for nm in (80, 128):
            cfg = oracle.mel_config(n_mels=nm)
            ref, ml, nf = oracle.mel_flat_transposed(cfg, a)
            assert [ml, nf] == g[f"{name}_{nm}_center_shape"].tolist()
            assert np.abs(ref.reshape(-1) - g[f"{name}_{nm}_center"]).max() <= 2e-4     # two float32 FFTs: north_star 1e-4 + vDSP's own noise
            ref, ml, nf = oracle.mel_flat_transposed(cfg, a, last=0.25, padding_mode=1)
            assert np.abs(ref.reshape(-1) - g[f"{name}_{nm}_prepadded"]).max() <= 2e-4
            ref, ml, nf = oracle.mel_flat(cfg, a)
            assert np.abs(ref.reshape(-1) - g[f"{name}_{nm}_flat"]).max() <= 2e-4


def test_oracle_vbx_against_swift_values(oracle):
    g = _load("swift_vbx.npz")
    out = oracle.vbx_refine(g["rho"], g["psi"], g["initial"])
    assert np.array_equal(np.asarray(out.hard, np.int32).reshape(-1), g["hard"].reshape(-1))   # labels: bit-exact
    assert len(out.elbos) == g["elbos"].size
    assert np.abs(out.gamma - g["gamma"]).max() <= 1e-6 and np.abs(out.pi - g["pi"]).max() <= 1e-6


def test_resampler_against_avaudioconverter(oracle):
    """Sample values of the documented Kaiser-sinc filter vs AVAudioConverter (Mastering): different filters, so only the
    pass band is comparable — length within 1 %, and the 440 Hz + 3 kHz tones agree within -40 dB after alignment."""
    g = _load("swift_resample.npz")
    for rate in (48000, 44100, 8000):
        y = g[f"resampled_{rate}"]
        ours = oracle.sinc_resample(g[f"pcm_{rate}"], rate, 16000)
        assert abs(y.size - ours.size) <= 0.01 * ours.size + 1
        n = min(y.size, ours.size) - 4000
        lags = range(-64, 65)
        best = min(np.abs(ours[2000:2000 + n] - y[2000 + lag:2000 + lag + n]).max() for lag in lags)
        assert best <= 1e-2
