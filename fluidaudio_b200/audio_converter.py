"""Host-side mirror of ``AudioConverter`` (Sources/FluidAudio/Shared/AudioConverter.swift:14-71, 299-442).

The conversion itself (mixdown, int16 widening, sample-rate conversion) runs on the GPU behind ``fa_audio_resample`` /
``fa_audio_to_mel``; nothing here computes.  One or two channels take the library's documented Kaiser-windowed-sinc
polyphase filter in place of Apple's closed AVAudioConverter ("parity unpinned" for sample values, output length as the
reference's tests demand); more than two channels take ``linearResample`` exactly as the reference does.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _lib


class Algorithm(enum.IntEnum):
    auto = 0      # AudioConverter.convertBuffer's rule: > 2 channels -> linear, else the converter
    sinc = 1
    linear = 2


class AudioConverterError(RuntimeError):
    """AudioConverterError (AudioConverter.swift:535-560)."""


def _format(in_rate, out_rate, channels, dtype, interleaved, algorithm) -> _lib.AudioFormat:
    if dtype == np.float32:
        fmt = 0
    elif dtype == np.int16:
        fmt = 1
    else:
        raise AudioConverterError(f"unsupported sample format {dtype}: float32 or int16")
    return _lib.AudioFormat(float(in_rate), float(out_rate), int(channels), fmt, int(bool(interleaved)), int(algorithm))


def _as_pcm(pcm, channels, interleaved):
    a = np.asarray(pcm)
    if a.dtype not in (np.float32, np.int16):
        a = a.astype(np.float32)
    a = np.ascontiguousarray(a)
    if a.ndim == 1:
        if channels is None:
            channels = 1
        frames = a.size // channels
    else:
        if interleaved:
            frames, ch = a.shape
        else:
            ch, frames = a.shape
        if channels is not None and channels != ch:
            raise AudioConverterError("channel count does not match the array shape")
        channels = ch
    return a, int(frames), int(channels)


class AudioConverter:
    def __init__(self, sample_rate: float = 16000.0, algorithm: Algorithm = Algorithm.auto):
        self.sample_rate = float(sample_rate)      # targetFormat.sampleRate (:22-52)
        self.algorithm = Algorithm(algorithm)
        self._L = _lib.load()

    def output_count(self, frames: int, input_rate: float) -> int:
        fmt = _format(input_rate, self.sample_rate, 1, np.dtype(np.float32), False, self.algorithm)
        return int(self._L.fa_resample_output_count(C.byref(fmt), int(frames)))

    def resample(self, samples, input_rate: float) -> np.ndarray:
        """``resample(_:from:)`` (:60-71): mono Float32 in, mono Float32 at the target rate out."""
        samples = np.ascontiguousarray(samples, np.float32).reshape(-1)
        if samples.size == 0:
            return np.zeros(0, np.float32)
        if float(input_rate) == self.sample_rate:
            return samples                                   # "return as-is"
        return self.resample_buffer(samples, input_rate, channels=1)

    def resample_buffer(self, pcm, input_rate: float, channels: int | None = None, interleaved: bool = False) -> np.ndarray:
        """``resampleBuffer`` (:77-85): any PCM layout (float32 / int16; planar [channels x frames] or interleaved
        [frames x channels]) -> mono Float32 at the target rate."""
        a, frames, channels = _as_pcm(pcm, channels, interleaved)
        fmt = _format(input_rate, self.sample_rate, channels, a.dtype, interleaved, self.algorithm)
        n = C.c_int64()
        _lib.check(self._L.fa_audio_resample(None, frames, C.byref(fmt), None, 0, C.byref(n)), "fa_audio_resample")
        out = np.zeros(n.value, np.float32)
        if n.value:
            _lib.check(self._L.fa_audio_resample(a.ctypes.data, frames, C.byref(fmt), out.ctypes.data, out.size, C.byref(n)),
                       "fa_audio_resample")
        return out
