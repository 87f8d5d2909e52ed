"""Filterbank-stage schedule: warp 0's handicap in the LPT deal (FA_MEL_ISSUE_HANDICAP, read at plan creation)."""
import os, sys, numpy as np
sys.path.insert(0, '.')
from fluidaudio_b200 import _lib, synth
from fluidaudio_b200.mel import AudioMelSpectrogram, Precision
n = 57_600_000
a = synth.tone_noise_audio(n)
d_a = _lib.DeviceBuffer(n * 4 + 64); d_a.upload(a)
ref = None
for h in (6, 0, 3, 9, 12, 16, 24, 6):
    os.environ["FA_MEL_ISSUE_HANDICAP"] = str(h)
    m = AudioMelSpectrogram(n_mels=80, precision=Precision.f32)
    T = m.frame_count(n)
    d_o = _lib.DeviceBuffer(T * 80 * 4)
    for _ in range(3): m.compute_device(d_a, n, d_o)
    m.timer_start()
    for _ in range(30): m.compute_device(d_a, n, d_o)
    ms = m.timer_stop_ms() / 30
    out = d_o.download((T * 80,), np.float32)
    if ref is None: ref = out
    print(f"handicap {h:2d}: {ms:.4f} ms per audio-hour  identical={np.array_equal(out, ref)}", flush=True)
    m.close()
