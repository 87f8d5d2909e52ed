"""Bridge to goldens produced by the REAL reference on a Mac (swift/Tools/DumpGoldens.swift).

    python tests/golden/swift_fixtures.py inputs <dir>    writes the synthetic inputs the Swift tool reads
    python tests/golden/swift_fixtures.py pack <dir>      packs the tool's outputs into tests/golden/swift_*.npz

No Swift toolchain exists in the build image, so the npz files are absent until a FluidAudio maintainer runs the tool;
tests that consume them skip with "parity unpinned" until then.  Inputs are regenerated from seeds (nothing binary is
committed for them); every array the tests need to re-run the same case is stored in the npz beside Apple's output.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fluidaudio_b200 import synth  # noqa: E402


def fixtures():
    audio = {"tone_noise": synth.tone_noise_audio(16000 * 4 + 137), "speech_like": synth.speech_like_audio(16000 * 4)}
    t = {r: np.arange(int(r * 1.5)) / r for r in (48000, 44100, 8000)}
    pcm = {r: (0.4 * np.sin(2 * np.pi * 440.0 * t[r]) + 0.2 * np.sin(2 * np.pi * 3000.0 * t[r])).astype(np.float32) for r in t}
    emb, who = synth.speaker_embeddings(400, 256, 4, weights=(0.4, 0.3, 0.2, 0.1), seed=17)
    rho, psi = synth.synthetic_plda(emb)
    return audio, pcm, rho, psi, who.astype(np.int32)


def write_inputs(out):
    os.makedirs(out, exist_ok=True)
    audio, pcm, rho, psi, init = fixtures()
    for k, v in audio.items():
        v.tofile(os.path.join(out, f"audio_{k}.f32"))
    for r, v in pcm.items():
        v.tofile(os.path.join(out, f"pcm_{r}.f32"))
    np.array(rho.shape, np.int32).tofile(os.path.join(out, "vbx_shape.i32"))
    rho.tofile(os.path.join(out, "vbx_rho.f64"))
    psi.tofile(os.path.join(out, "vbx_psi.f64"))
    init.tofile(os.path.join(out, "vbx_initial.i32"))
    print("inputs written to", out)


def pack(src):
    man = json.load(open(os.path.join(src, "manifest.json")))
    audio, pcm, rho, psi, init = fixtures()
    mel = {"host": np.array(man.get("host", ""))}
    for name, a in audio.items():
        mel[f"audio_{name}"] = a
        for nm in (80, 128):
            for kind in ("center", "prepadded", "flat"):
                mel[f"{name}_{nm}_{kind}"] = np.fromfile(os.path.join(src, f"mel_{name}_{nm}_{kind}.f32"), np.float32)
                mel[f"{name}_{nm}_{kind}_shape"] = np.array(man[f"mel_{name}_{nm}"][kind])
        mel[f"{name}_128_legacy"] = np.fromfile(os.path.join(src, f"mel_{name}_128_legacy.f32"), np.float32)
    mel["filterbank_80"] = np.fromfile(os.path.join(src, "filterbank_80.f32"), np.float32)
    mel["hann_400"] = np.fromfile(os.path.join(src, "hann_400.f32"), np.float32)
    np.savez_compressed(os.path.join(HERE, "swift_mel.npz"), **mel)
    rs = {f"pcm_{r}": v for r, v in pcm.items()}
    for r in pcm:
        rs[f"resampled_{r}"] = np.fromfile(os.path.join(src, f"resampled_{r}.f32"), np.float32)
    np.savez_compressed(os.path.join(HERE, "swift_resample.npz"), **rs)
    S = int(man["vbx"]["S"])
    np.savez_compressed(os.path.join(HERE, "swift_vbx.npz"), rho=rho, psi=psi, initial=init,
                        gamma=np.fromfile(os.path.join(src, "vbx_gamma.f64"), np.float64).reshape(rho.shape[0], S),
                        pi=np.fromfile(os.path.join(src, "vbx_pi.f64"), np.float64),
                        elbos=np.fromfile(os.path.join(src, "vbx_elbos.f64"), np.float64),
                        hard=np.fromfile(os.path.join(src, "vbx_hard.i32"), np.int32))
    print("packed swift_mel.npz, swift_resample.npz, swift_vbx.npz")


if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in ("inputs", "pack"):
        raise SystemExit(__doc__)
    (write_inputs if sys.argv[1] == "inputs" else pack)(sys.argv[2])
