#!/bin/bash
# timeline of the int16 host pipeline (FA_MEL_TRACE_PIPELINE): where do the 0.7 ms above the bare-copy floor go?
mkdir -p gpurun_out
FA_MEL_TRACE_PIPELINE=1 timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/pipeline_trace.log
import sys, time, numpy as np
sys.path.insert(0, '.')
from fluidaudio_b200 import _lib, synth
from fluidaudio_b200.mel import AudioMelSpectrogram, Precision
n = 57_600_000
a = synth.tone_noise_audio(n)
T = 360001
pin_16 = _lib.PinnedArray(n, np.int16); pin_16.array[:] = np.round(a * 32767).astype(np.int16)
pin_out = _lib.PinnedArray(T * 80, np.float32)
m = AudioMelSpectrogram(n_mels=80, precision=Precision.f32)
L = _lib.load()
for chunks in (4, 12):
    _lib.check(L.fa_mel_set_pipeline_chunks(m._h, chunks), "chunks")
    for rep in range(4):
        t0 = time.perf_counter()
        m.compute_from_pcm(pin_16.array, 16000.0, out=pin_out.array)
        print(f"chunks {chunks} rep {rep}: wall {(time.perf_counter()-t0)*1e3:.3f} ms", flush=True)
PY
