#!/bin/bash
# mel-only GPU check: parity tests + kernel timing + optional ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel" 2>&1 | tail -15
timeout 300 python - <<'PY'
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
from fluidaudio_b200 import _lib, synth
from fluidaudio_b200.mel import AudioMelSpectrogram
from oracle import oracle as O
n = 57_600_000
a = synth.tone_noise_audio(n)
from fluidaudio_b200.mel import Precision
for nm, prec in ((80, Precision.f64), (80, Precision.f32), (128, Precision.f64), (128, Precision.f32)):
    m = AudioMelSpectrogram(n_mels=nm, precision=prec)
    T = m.frame_count(n)
    d_a = _lib.DeviceBuffer(n * 4 + 64); d_a.upload(a)
    d_o = _lib.DeviceBuffer(T * nm * 4)
    for _ in range(3): m.compute_device(d_a, n, d_o)
    m.timer_start()
    for _ in range(20): m.compute_device(d_a, n, d_o)
    ms = m.timer_stop_ms() / 20
    print(f"nm={nm} {prec.name} kernel-only {ms:.4f} ms/h -> {1e3/ms:.0f} audio-h/s  {(4*n+4*T*nm)/ms/1e6:.0f} GB/s", flush=True)
    got = d_o.download((T, nm), np.float32)
    ref, rml, _ = O.mel_flat_transposed(O.mel_config(n_mels=nm), a[:16000*120])
    d = np.abs(got[:rml-3] - ref[:rml-3])
    print(f"   vs oracle (first 120 s): max|d|={d.max():.3e} mismatches={(d>0).mean():.4f}", flush=True)
PY
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel python scripts/profile_target.py mel 3 > gpurun_out/ncu_mel_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel_f32 python scripts/profile_target.py mel32 3 > gpurun_out/ncu_mel32_full.log 2>&1
fi
