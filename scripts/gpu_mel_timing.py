"""Kernel-only timing of the mel kernel (device-resident hour) at 80 and 128 mels, both value types, both layouts."""
import sys, numpy as np
sys.path.insert(0, '.')
from fluidaudio_b200 import _lib, synth
from fluidaudio_b200.mel import AudioMelSpectrogram, Precision
n = 57_600_000
a = synth.tone_noise_audio(n)
d_a = _lib.DeviceBuffer(n * 4 + 64); d_a.upload(a)
for nm in (80, 128):
    for prec in (Precision.f32, Precision.f64):
        for tm in (True, False):
            m = AudioMelSpectrogram(n_mels=nm, precision=prec)
            T = m.frame_count(n)
            d_o = _lib.DeviceBuffer(T * nm * 4)
            kw = {} if tm else dict(time_major=False)
            try:
                for _ in range(3): m.compute_device(d_a, n, d_o, **kw)
                m.timer_start()
                for _ in range(20): m.compute_device(d_a, n, d_o, **kw)
                print(f"n_mels {nm:3d} {prec.name} {'time-major' if tm else 'mel-major '}: {m.timer_stop_ms()/20:.4f} ms per audio-hour", flush=True)
            except TypeError as e:
                print("skip", e)
            m.close()
