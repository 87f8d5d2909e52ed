#!/bin/bash
# mel experiments: pipeline depth of the host-buffer paths, ncu capture of the float32-pair kernel
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/mel_exp.log
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.')
from fluidaudio_b200 import _lib, synth
from fluidaudio_b200.mel import AudioMelSpectrogram, Precision
n = 57_600_000
a = synth.tone_noise_audio(n)
T = 360001
pin_in = _lib.PinnedArray(n, np.float32); pin_in.array[:] = a
pin_16 = _lib.PinnedArray(n, np.int16); pin_16.array[:] = np.round(a * 32767).astype(np.int16)
pin_out = _lib.PinnedArray(T * 80, np.float32)
m = AudioMelSpectrogram(n_mels=80, precision=Precision.f32)
L = _lib.load()
def t(fn, reps=10):
    for _ in range(3): fn()
    _lib.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    _lib.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
d_a = _lib.DeviceBuffer(n * 4 + 64); d_a.upload(a)
d_o = _lib.DeviceBuffer(T * 80 * 4)
for _ in range(3): m.compute_device(d_a, n, d_o)
m.timer_start()
for _ in range(20): m.compute_device(d_a, n, d_o)
print(f"kernel-only f32: {m.timer_stop_ms()/20:.4f} ms/h", flush=True)
ref = None
for zc in ((1, 0) if "ZC" in __import__("os").environ else (0,)):
    _lib.check(L.fa_mel_set_zero_copy_output(m._h, zc), "zc")
    for chunks in (2, 4, 8, 12, 16, 24, 32, 48, 96):
        _lib.check(L.fa_mel_set_pipeline_chunks(m._h, chunks), "chunks")
        f = t(lambda: m.compute_flat_transposed(pin_in.array, out=pin_out.array))
        if ref is None: ref = pin_out.array.copy()
        assert np.array_equal(ref, pin_out.array)
        i = t(lambda: m.compute_from_pcm(pin_16.array, 16000.0, out=pin_out.array))
        print(f"zero_copy={zc} chunks={chunks:4d}  e2e f32 {f:.3f} ms   e2e i16 {i:.3f} ms", flush=True)
_lib.check(L.fa_mel_set_zero_copy_output(m._h, 1), "zc")
ms = C.c_float()
for nb_in, nb_out in ((4*n, 4*T*80), (2*n, 4*T*80), (4*n, 0), (2*n, 0), (0, 4*T*80)):
    L.fa_memcpy_probe(pin_in.array.ctypes.data, nb_in, pin_out.array.ctypes.data, nb_out, 10, C.byref(ms))
    print(f"probe h2d {nb_in/1e6:.0f} MB d2h {nb_out/1e6:.0f} MB: {ms.value:.3f} ms", flush=True)
# 48 kHz stereo int16 -> mel (sinc resampler in the pipeline)
st = np.stack([synth.tone_noise_audio(48000 * 600, sample_rate=48000), synth.tone_noise_audio(48000 * 600, seed=9, sample_rate=48000)])
i16 = np.ascontiguousarray(np.round(st.T * 32767).astype(np.int16))
pin_st = _lib.PinnedArray(i16.shape, np.int16); pin_st.array[:] = i16
_lib.check(L.fa_mel_set_pipeline_chunks(m._h, 24), "chunks")
ms48 = t(lambda: m.compute_from_pcm(pin_st.array, 48000.0, interleaved=True), reps=5)
print(f"600 s of 48 kHz stereo int16 -> mel: {ms48:.3f} ms  ({600/3600/(ms48*1e-3):.1f} audio-h/s, {i16.nbytes/1e6:.0f} MB in)")
PY
if [ "$1" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel_f32 python scripts/profile_target.py mel32 3 > gpurun_out/ncu_mel32_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel python scripts/profile_target.py mel 3 > gpurun_out/ncu_mel_full.log 2>&1
fi
ls -la gpurun_out | head -30
