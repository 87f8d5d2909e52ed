// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// The CPU arm bench.py times beside the GPU (`cpu_baseline`, `--impl reference`): AudioMelSpectrogram's
// computeFlatTransposed (Sources/FluidAudio/Shared/AudioMelSpectrogram.swift:325-456) as the reference actually runs
// it — float32 throughout, a float32 512-point complex FFT of the real frame (:459-481, vDSP_DFT_zop), a DENSE
// [nMels x 257] mat-vec (:434-447, vDSP_mmul), logf — written the way a vector library would run it on x86: sixteen
// frames at a time in structure-of-arrays form so that every butterfly, multiply-add and dot product is one SIMD
// operation across frames (built -O3 -march=x86-64-v3: AVX2 + FMA, present on every host these boxes use).
// oracle_mel.cpp stays the parity oracle (DFT rounded once from float64, pinned flags); this file is only the *timed*
// CPU implementation, and tests check it against the oracle within the float32-FFT spread (2e-4).
// No Swift toolchain and no Accelerate exist here, so this is still `kind: "port"` — but of the reference's
// arithmetic (float32 FFT), not of the oracle's (float64 DFT), and vectorised as vDSP is.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {
struct oracle_mel_config;   // layout in oracle_mel.cpp
void oracle_mel_hann(int32_t length, int32_t periodic, float *out);
void oracle_mel_filterbank(int32_t n_fft, int32_t n_mels, int32_t sample_rate, float *out);
}

namespace {

constexpr int kLanes = 16;   // frames processed together (two AVX2 registers per value)
constexpr int kN = 512, kBins = 257;

struct FastMel {
    int n_mels, hop, win, off;
    float preemph, log_floor;
    int clamped;
    std::vector<float> window, fb, cs, sn;
    std::vector<int> rev;
    // SoA work buffers: [index][lane]
    std::vector<float> re, im, power;

    FastMel(int n_mels_, int hop_, int win_, float preemph_, float log_floor_, int clamped_, int periodic, int sample_rate)
        : n_mels(n_mels_), hop(hop_), win(win_), off((kN - win_) / 2), preemph(preemph_), log_floor(log_floor_),
          clamped(clamped_) {
        window.resize(win);
        oracle_mel_hann(win, periodic, window.data());
        fb.resize((size_t)n_mels * kBins);
        oracle_mel_filterbank(kN, n_mels, sample_rate, fb.data());
        cs.resize(kN / 2);
        sn.resize(kN / 2);
        for (int k = 0; k < kN / 2; ++k) {
            const double a = 2.0 * M_PI * (double)k / (double)kN;
            cs[k] = (float)std::cos(a);
            sn[k] = (float)-std::sin(a);
        }
        rev.resize(kN);
        for (int i = 0; i < kN; ++i) {
            int r = 0;
            for (int b = 0; b < 9; ++b)
                if (i & (1 << b)) r |= 1 << (8 - b);
            rev[i] = r;
        }
        re.assign((size_t)kN * kLanes, 0.0f);
        im.assign((size_t)kN * kLanes, 0.0f);
        power.assign((size_t)kBins * kLanes, 0.0f);
    }

    // p: pre-emphasised, zero-padded signal; frames f0 .. f0+count-1 (count <= kLanes); out time-major
    void run(const float *p, int64_t padded, int64_t f0, int count, float *out) {
        // window product into bit-reversed order (imag = 0)
        std::fill(re.begin(), re.end(), 0.0f);
        std::fill(im.begin(), im.end(), 0.0f);
        for (int l = 0; l < count; ++l) {
            const int64_t start = (f0 + l) * hop + off;
            const int64_t avail = std::min<int64_t>(win, padded - start);
            for (int64_t k = 0; k < avail; ++k) re[(size_t)rev[off + k] * kLanes + l] = p[start + k] * window[k];
        }
        float *R = re.data(), *I = im.data();
        for (int len = 2; len <= kN; len <<= 1) {
            const int half = len >> 1, step = kN / len;
            for (int base = 0; base < kN; base += len) {
                for (int j = 0; j < half; ++j) {
                    const float wr = cs[j * step], wi = sn[j * step];
                    float *ar = R + (size_t)(base + j) * kLanes, *ai = I + (size_t)(base + j) * kLanes;
                    float *br = ar + (size_t)half * kLanes, *bi = ai + (size_t)half * kLanes;
#pragma GCC ivdep
                    for (int l = 0; l < kLanes; ++l) {
                        const float tr = br[l] * wr - bi[l] * wi;
                        const float ti = br[l] * wi + bi[l] * wr;
                        br[l] = ar[l] - tr;
                        bi[l] = ai[l] - ti;
                        ar[l] += tr;
                        ai[l] += ti;
                    }
                }
            }
        }
        float *P = power.data();
        for (int b = 0; b < kBins; ++b)
#pragma GCC ivdep
            for (int l = 0; l < kLanes; ++l) P[b * kLanes + l] = R[b * kLanes + l] * R[b * kLanes + l] + I[b * kLanes + l] * I[b * kLanes + l];
        // dense mat-vec as vDSP_mmul: four filters at a time so that eight independent SIMD accumulators hide the FMA latency
        typedef float v8 __attribute__((vector_size(32), aligned(4)));
        for (int m0 = 0; m0 < n_mels; m0 += 4) {
            const int mc = std::min(4, n_mels - m0);
            v8 acc[4][2];
            for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = v8{0, 0, 0, 0, 0, 0, 0, 0};
            const float *rows[4];
            for (int q = 0; q < 4; ++q) rows[q] = &fb[(size_t)(m0 + std::min(q, mc - 1)) * kBins];
            for (int b = 0; b < kBins; ++b) {
                const v8 p0 = *reinterpret_cast<const v8 *>(P + b * kLanes), p1 = *reinterpret_cast<const v8 *>(P + b * kLanes + 8);
                for (int q = 0; q < 4; ++q) {
                    const float w = rows[q][b];
                    acc[q][0] += w * p0;
                    acc[q][1] += w * p1;
                }
            }
            for (int q = 0; q < mc; ++q)
                for (int l = 0; l < count; ++l) {
                    const float a = l < 8 ? acc[q][0][l] : acc[q][1][l - 8];
                    const float v = clamped ? std::max(a, log_floor) : a + log_floor;
                    out[(f0 + l) * n_mels + m0 + q] = logf(v);
                }
        }
    }
};

} // namespace

extern "C" {

// computeFlatTransposed, .center, pad_to 1, nFFT 512.  cfg fields are read through this mirror of oracle_mel_config.
struct fast_cfg {
    int32_t sample_rate, n_mels, n_fft, hop_length, win_length;
    float preemph;
    int32_t pad_to;
    float log_floor;
    int32_t log_floor_mode, window_periodic, precision;
};

int64_t oracle_mel_fast_flat_transposed(const void *cfg_, const float *audio, int64_t n, float last, float *out,
                                        int64_t out_cap, int64_t *mel_length) {
    const fast_cfg &c = *reinterpret_cast<const fast_cfg *>(cfg_);
    if (c.n_fft != kN || n <= 0) return -1;
    const int64_t pad = kN / 2, padded = n + 2 * pad;
    const int64_t T = 1 + (padded - c.win_length) / c.hop_length;
    if (mel_length) *mel_length = T;
    const int64_t need = T * c.n_mels;
    if (!out || out_cap < need) return need;
    std::vector<float> p((size_t)padded + kN, 0.0f);
    if (c.preemph == 0.0f) {
        std::memcpy(&p[pad], audio, (size_t)n * sizeof(float));
    } else {
        p[pad] = audio[0] - c.preemph * last;
        const float neg = -c.preemph;
        for (int64_t i = 1; i < n; ++i) p[pad + i] = fmaf(audio[i - 1], neg, audio[i]);
    }
    thread_local FastMel *mel = nullptr;
    thread_local fast_cfg held{};
    if (!mel || std::memcmp(&held, &c, sizeof(c)) != 0) {
        delete mel;
        mel = new FastMel(c.n_mels, c.hop_length, c.win_length, c.preemph, c.log_floor, c.log_floor_mode,
                          c.window_periodic, c.sample_rate);
        held = c;
    }
    for (int64_t f = 0; f < T; f += kLanes) mel->run(p.data(), padded, f, (int)std::min<int64_t>(kLanes, T - f), out);
    return need;
}

} // extern "C"
