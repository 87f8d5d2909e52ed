// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatements of the small pieces either side of the mel hot path:
//   R1  Sources/FluidAudio/Shared/AudioConverter.swift:388-442  linearResample (>2-channel fallback: mean
//       mixdown, src = i*ratio, 2-tap lerp, outCount = Int(inCount/ratio)).  The main AVAudioConverter path
//       (:299-375) is closed-source Apple code: "parity unpinned" for its sample values.
//   M5  Sources/FluidAudio/ASR/Parakeet/Unified/UnifiedMelExtractor.swift:88-113  normalizePerFeature
//       Sources/FluidAudio/ASR/Parakeet/Streaming/Nemotron/NemotronMelExtractor.swift:44-67  [T x M] -> [1,M,T]

#include <cmath>
#include <malloc.h>
#include <cstdint>

extern "C" {

// channels: planar [channel][frame] laid out channel-major (channelData[c][f] = in[c*frames + f]).
// Returns the output count; writes only when out != nullptr.
int64_t oracle_linear_resample(const float *in, int64_t frames, int32_t channels, double in_rate, double out_rate,
                               float *out) {
    if (in_rate == out_rate) {
        if (out) {
            const float w = 1.0f / (float)channels;
            for (int64_t f = 0; f < frames; ++f) {
                float s = 0;
                for (int32_t c = 0; c < channels; ++c) s += in[(int64_t)c * frames + f];
                out[f] = s * w;
            }
        }
        return frames;
    }
    const double ratio = in_rate / out_rate;
    const int64_t out_count = (int64_t)((double)frames / ratio);
    if (!out) return out_count;
    const float w = 1.0f / (float)channels;
    auto mono = [&](int64_t f) {
        float s = 0;
        for (int32_t c = 0; c < channels; ++c) s += in[(int64_t)c * frames + f];
        return s * w;
    };
    for (int64_t i = 0; i < out_count; ++i) {
        const double src = (double)i * ratio;
        const int64_t idx = (int64_t)src;
        const float frac = (float)(src - (double)idx);
        if (idx < frames - 1) out[i] = mono(idx) * (1.0f - frac) + mono(idx + 1) * frac;
        else if (idx < frames) out[i] = mono(idx);
        else out[i] = 0.0f;
    }
    return out_count;
}

// In-place NeMo per-feature normalisation on a time-major [frames x n_mels] buffer.
void oracle_normalize_per_feature(float *x, int64_t frames, int32_t n_mels, int64_t valid_frames) {
    if (valid_frames <= 0) {
        for (int64_t i = 0; i < frames * n_mels; ++i) x[i] = 0;
        return;
    }
    const float denom = (float)(valid_frames > 1 ? valid_frames - 1 : 1);
    for (int32_t m = 0; m < n_mels; ++m) {
        float mean = 0;
        for (int64_t t = 0; t < valid_frames; ++t) mean += x[t * n_mels + m];
        mean /= (float)valid_frames;
        float var_sum = 0;
        for (int64_t t = 0; t < valid_frames; ++t) {
            const float d = x[t * n_mels + m] - mean;
            var_sum += d * d;
        }
        const float sd = sqrtf(var_sum / denom) + 1e-5f;
        for (int64_t t = 0; t < frames; ++t)
            x[t * n_mels + m] = t < valid_frames ? (x[t * n_mels + m] - mean) / sd : 0.0f;
    }
}

// LS-EEND feature scaling + cumulative mean normalisation (LSEENDPreprocessor.swift:259-279), in place on a time-major
// [frames x n_mels] log-mel buffer, carrying (cmn_mean[n_mels], cmn_count) across calls exactly like the preprocessor's
// state.  Per frame k: count += 1; alpha = 1 / Float(count); mean = mean + alpha * (x - mean)  (vDSP_vintb: one
// subtract, one multiply, one add, each rounded to float32); x = x - mean  (vDSP_vsub).  The scale is
// Float(1) / logf(10) (:36) applied with vDSP_vsmul first.
void oracle_lseend_scale_cmn(float *x, int64_t frames, int32_t n_mels, float *cmn_mean, int64_t *cmn_count) {
    const float scale = 1.0f / logf(10.0f);
    int64_t count = *cmn_count;
    for (int64_t t = 0; t < frames; ++t) {
        count += 1;
        const float alpha = 1.0f / (float)count;
        float *row = x + t * n_mels;
        for (int32_t m = 0; m < n_mels; ++m) {
            const float v = row[m] * scale;
            const float d = v - cmn_mean[m];
            const float ad = alpha * d;
            const float mu = cmn_mean[m] + ad;
            cmn_mean[m] = mu;
            row[m] = v - mu;
        }
    }
    *cmn_count = count;
}

// Benchmark hygiene for the multi-threaded CPU baseline (bench.py): keep multi-megabyte scratch vectors on the
// per-thread heaps instead of mmap/munmap-ing them on every call, which serialises 64 threads on the process's
// address-space lock and on page faults.  Has no effect on results.
void oracle_tune_allocator(void) {
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
}

// time-major [T x M] -> mel-major [M x T]
void oracle_transpose_tm(const float *in, int64_t T, int32_t M, float *out) {
    for (int64_t t = 0; t < T; ++t)
        for (int32_t m = 0; m < M; ++m) out[(int64_t)m * T + t] = in[t * M + m];
}

} // extern "C"
