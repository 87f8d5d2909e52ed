#!/bin/bash
# Round-2 GPU session: full GPU test-suite, smoke, mel timing (both transforms), bench line, ncu captures.
# Everything lands in gpurun_out/.  Usage: bash scripts/gpu_r02.sh [ncu]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/mel_quick.log
import sys, numpy as np
sys.path.insert(0, '.')
from fluidaudio_b200 import _lib, synth
from fluidaudio_b200.mel import AudioMelSpectrogram, Precision
from oracle import oracle as O
n = 57_600_000
a = synth.tone_noise_audio(n)
ref80 = None
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
timeout 900 python bench.py 2>gpurun_out/bench.err > gpurun_out/bench.json; tail -c 600 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench.json'))
    c = d['cluster']
    print('mel', d['value'], d['ms_per_step'], 'sustained', d['sustained']['ms_per_step'], 'f64', d['f64_transform']['ms_per_step'], 'frac', d['roofline']['frac'])
    print('parity', d['parity'])
    print('e2e', d['e2e']['ms_per_step'], d['e2e']['copy_floor_ms'], 'i16', d['e2e_i16']['ms_per_step'], d['e2e_i16']['copy_floor_ms'])
    print('cluster', c['value'], c['ms_per_step'], c['stages_ms'], c.get('labels_equal_ref'), c.get('labels_equal_cpu'))
    print('c4', d.get('c4')); print('c5', d.get('c5')); print('streaming', d.get('streaming')); print('clocks', d['clocks'])
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench.err > gpurun_out/bench_reference.json; cat gpurun_out/bench_reference.json | head -c 900; echo
if [ "$1" == "ncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --only-main > gpurun_out/ncu_bench_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel_f32 python scripts/profile_target.py mel32 3 > gpurun_out/ncu_mel32_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel python scripts/profile_target.py mel 3 > gpurun_out/ncu_mel_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cluster.csv python scripts/profile_target.py cluster 2 > gpurun_out/ncu_cluster_list.log 2>&1
fi
ls gpurun_out | head -50
