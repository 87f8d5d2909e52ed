// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement (plain C++, single thread) of FluidAudio's log-mel frontend.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product (fluidaudio_b200/) never does.
//
// Follows, line by line:
//   Sources/FluidAudio/Shared/AudioMelSpectrogram.swift
//     :59-121   init (window, filterbank, flat filterbank)
//     :132-178  compute            (legacy: no preemph, no centre pad, window at offset 0)
//     :185-292  computeFlat        (mel-major  [nMels x Tp])
//     :325-456  computeFlatTransposed (time-major [Tp x nMels]; .center / .prePadded / expectedFrameCount)
//     :459-481  computePowerSpectrumInPlace (512-pt complex DFT of a real frame, re^2+im^2, bins 0..nFFT/2)
//     :542-549  logValue (additive / clamped)
//     :553-562  createHannWindow (Float32, symmetric or periodic)
//     :564-642  createMelFilterbank (Slaney scale + Slaney norm, Float32 throughout)
//
// Parity status: the reference has NO golden numeric vectors for mel values and its DFT,
// mat-vec, cos/log/exp come from Apple Accelerate / libm (closed).  This restatement is therefore the
// oracle by construction: every operation is done in IEEE float32 in the order the Swift source states;
// where the Swift delegates to a closed library the mathematically defined operation is used
// (mat-vec = sequential float32 accumulation in bin order; vDSP_vsma = fused multiply-add).  The DFT is the one
// place where "the reference's float32 algorithm" cannot be restated (vDSP_DFT_zop is closed): the oracle
// therefore uses the implementation-independent definition — the DFT of the float32 frame evaluated in
// float64 and rounded ONCE to float32 (precision = 0), which every float32 FFT, vDSP's included, approximates
// to within its own rounding noise.  Two more evaluations are provided so that tests can report the spread:
// precision = 2 runs a float32 radix-2 Cooley-Tukey instead (a second, independent float32 FFT), precision = 1
// runs the whole pipeline in float64 (distance from exact arithmetic).  "Parity pinned" only for structure: frame counts, shapes, window/filterbank
// properties (AudioMelSpectrogramTests.swift, EouChunkSizeFrameCountTests.swift).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

extern "C" {

struct oracle_mel_config {
    int32_t sample_rate;     // 16000
    int32_t n_mels;          // 128 default, 80 for BASELINE config 2
    int32_t n_fft;           // 512
    int32_t hop_length;      // 160
    int32_t win_length;      // 400
    float   preemph;         // 0.97
    int32_t pad_to;          // 0 -> treated as 1 (AudioMelSpectrogram.swift:72)
    float   log_floor;       // 2^-24
    int32_t log_floor_mode;  // 0 additive, 1 clamped
    int32_t window_periodic; // 0 symmetric, 1 periodic
    int32_t precision;       // 0 = float32 pipeline, DFT correctly rounded (THE oracle); 1 = all float64;
                             // 2 = float32 pipeline with a float32 radix-2 FFT (spread indicator)
};

// Swift's Float.pi is pi rounded TOWARD ZERO (0x40490FDA), not to nearest.
static inline float swift_float_pi() {
    uint32_t bits = 0x40490FDAu;
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

// AudioMelSpectrogram.swift:553-562
void oracle_mel_hann(int32_t length, int32_t periodic, float *out) {
    const float divisor = periodic ? (float)length : (float)(length - 1);
    const float pi = swift_float_pi();
    for (int32_t i = 0; i < length; ++i) {
        const float phase = 2.0f * pi * (float)i / divisor;
        out[i] = 0.5f * (1.0f - cosf(phase));
    }
}

// AudioMelSpectrogram.swift:564-642 — every intermediate is Float.
static float hz_to_mel(float hz) {
    const float fSp = 200.0f / 3.0f;
    const float minLogHz = 1000.0f;
    const float minLogMel = minLogHz / fSp;
    const float logStep = logf(6.4f) / 27.0f;
    if (hz >= minLogHz) return minLogMel + logf(hz / minLogHz) / logStep;
    return hz / fSp;
}
static float mel_to_hz(float mel) {
    const float fSp = 200.0f / 3.0f;
    const float minLogHz = 1000.0f;
    const float minLogMel = minLogHz / fSp;
    const float logStep = logf(6.4f) / 27.0f;
    if (mel >= minLogMel) return minLogHz * expf(logStep * (mel - minLogMel));
    return fSp * mel;
}

// out: row-major [n_mels x (n_fft/2+1)]  (melFilterbankFlat, :96-104)
void oracle_mel_filterbank(int32_t n_fft, int32_t n_mels, int32_t sample_rate, float *out) {
    const int32_t bins = n_fft / 2 + 1;
    const float fMin = 0.0f;
    const float fMax = (float)sample_rate / 2.0f;
    const float melMin = hz_to_mel(fMin);
    const float melMax = hz_to_mel(fMax);
    std::vector<float> pts(n_mels + 2);
    for (int32_t i = 0; i < n_mels + 2; ++i) {
        const float mel = melMin + (float)i * (melMax - melMin) / (float)(n_mels + 1);
        pts[i] = mel_to_hz(mel);
    }
    std::vector<float> freqs(bins);
    for (int32_t i = 0; i < bins; ++i) freqs[i] = (float)i * (float)sample_rate / (float)n_fft;
    std::fill(out, out + (size_t)n_mels * bins, 0.0f);
    for (int32_t m = 0; m < n_mels; ++m) {
        const float fl = pts[m], fc = pts[m + 1], fr = pts[m + 2];
        const float norm = 2.0f / (fr - fl);
        for (int32_t b = 0; b < bins; ++b) {
            const float f = freqs[b];
            if (f >= fl && f < fc) {
                out[(size_t)m * bins + b] = norm * (f - fl) / (fc - fl);
            } else if (f >= fc && f <= fr) {
                out[(size_t)m * bins + b] = norm * (fr - f) / (fr - fc);
            }
        }
    }
}

// Frame count rules.  mode: 0 = .center (:338-343), 1 = .prePadded (:344-346), 2 = legacy compute() (:133).
// expected < 0 means "nil".
int64_t oracle_mel_frame_count(const oracle_mel_config *c, int64_t n, int32_t mode, int64_t expected) {
    int64_t computed;
    if (mode == 0) {
        const int64_t padded = n + 2 * (int64_t)(c->n_fft / 2);
        // Swift Int division truncates toward zero
        computed = 1 + (padded - c->win_length) / c->hop_length;
    } else if (mode == 1) {
        computed = std::max<int64_t>(0, (n - c->n_fft) / c->hop_length + 1);
    } else {
        computed = 1 + (n - c->win_length) / c->hop_length;
    }
    return expected >= 0 ? expected : computed;
}

} // extern "C"

namespace {

template <typename T>
struct Fft {
    int n = 0;
    bool pow2 = false;
    std::vector<T> cs, sn;  // twiddles e^{-2 pi i k / n}: cos, -sin handled below
    std::vector<int> rev;
    void init(int n_) {
        n = n_;
        pow2 = n > 0 && (n & (n - 1)) == 0;
        cs.resize(n);
        sn.resize(n);
        for (int k = 0; k < n; ++k) {
            const double a = 2.0 * M_PI * (double)k / (double)n;
            cs[k] = (T)std::cos(a);
            sn[k] = (T)std::sin(a);
        }
        if (pow2) {
            rev.resize(n);
            int lg = 0;
            while ((1 << lg) < n) ++lg;
            for (int i = 0; i < n; ++i) {
                int r = 0;
                for (int b = 0; b < lg; ++b)
                    if (i & (1 << b)) r |= 1 << (lg - 1 - b);
                rev[i] = r;
            }
        }
    }
    // forward, unnormalised DFT of a REAL input (imag = 0), as vDSP_DFT_zop with imagIn cleared (:462-471)
    void forward_real(const T *in, T *re, T *im) const {
        if (pow2) {
            for (int i = 0; i < n; ++i) {
                re[rev[i]] = in[i];
                im[rev[i]] = (T)0;
            }
            for (int len = 2; len <= n; len <<= 1) {
                const int half = len >> 1, step = n / len;
                for (int base = 0; base < n; base += len) {
                    for (int j = 0; j < half; ++j) {
                        const T wr = cs[j * step], wi = -sn[j * step];
                        const int a = base + j, b = a + half;
                        const T tr = re[b] * wr - im[b] * wi;
                        const T ti = re[b] * wi + im[b] * wr;
                        re[b] = re[a] - tr;
                        im[b] = im[a] - ti;
                        re[a] = re[a] + tr;
                        im[a] = im[a] + ti;
                    }
                }
            }
        } else {
            for (int k = 0; k < n; ++k) {
                T sr = 0, si = 0;
                for (int t = 0; t < n; ++t) {
                    const int idx = (int)(((int64_t)k * t) % n);
                    sr += in[t] * cs[idx];
                    si -= in[t] * sn[idx];
                }
                re[k] = sr;
                im[k] = si;
            }
        }
    }
};

template <typename T>
struct MelCore {
    const oracle_mel_config &c;
    int bins;
    std::vector<float> window;
    std::vector<float> fb;
    Fft<T> fft;
    Fft<double> fft_exact;
    std::vector<double> dframe, dre, dim;
    std::vector<T> frame, re, im, power;

    explicit MelCore(const oracle_mel_config &cfg) : c(cfg) {
        bins = c.n_fft / 2 + 1;
        window.resize(c.win_length);
        oracle_mel_hann(c.win_length, c.window_periodic, window.data());
        fb.resize((size_t)c.n_mels * bins);
        oracle_mel_filterbank(c.n_fft, c.n_mels, c.sample_rate, fb.data());
        fft.init(c.n_fft);
        if (sizeof(T) == 4 && c.precision == 0) {
            fft_exact.init(c.n_fft);
            dframe.assign(c.n_fft, 0);
            dre.assign(c.n_fft, 0);
            dim.assign(c.n_fft, 0);
        }
        frame.assign(c.n_fft, 0);
        re.assign(c.n_fft, 0);
        im.assign(c.n_fft, 0);
        power.assign(bins, 0);
    }
    T log_value(T v) const {
        const T fl = (T)c.log_floor;
        if (c.log_floor_mode == 0) return (T)std::log(v + fl);
        return (T)std::log(std::max(v, fl));
    }
    // frame[] already filled.  Writes n_mels log values through `store(m, value)`.
    template <typename Store>
    void finish_frame(Store store) {
        if (sizeof(T) == 4 && c.precision == 0) {
            for (int i = 0; i < c.n_fft; ++i) dframe[i] = (double)frame[i];
            fft_exact.forward_real(dframe.data(), dre.data(), dim.data());
            for (int b = 0; b < bins; ++b) {
                re[b] = (T)dre[b];   // single rounding to float32
                im[b] = (T)dim[b];
            }
        } else {
            fft.forward_real(frame.data(), re.data(), im.data());
        }
        for (int b = 0; b < bins; ++b) power[b] = re[b] * re[b] + im[b] * im[b];
        for (int m = 0; m < c.n_mels; ++m) {
            T acc = 0;
            const float *row = &fb[(size_t)m * bins];
            for (int b = 0; b < bins; ++b) acc += (T)row[b] * power[b];
            store(m, log_value(acc));
        }
    }
};

// computeFlat / computeFlatTransposed (:185-292, :325-456).  layout 0 = time-major, 1 = mel-major.
template <typename T>
int64_t run_flat(const oracle_mel_config &c, const float *audio, int64_t n, float last, int32_t mode,
                 int64_t expected, int32_t layout, float *out, int64_t out_cap, int64_t *mel_length,
                 int64_t *num_frames) {
    const int64_t T_frames = oracle_mel_frame_count(&c, n, mode, layout == 1 ? -1 : expected);
    const int pad_to = std::max(1, c.pad_to);
    if (T_frames <= 0 || n <= 0) {
        // guard (:199-201, :349-351): mel = [padValue] * nMels, melLength 0, numFrames 1
        if (mel_length) *mel_length = 0;
        if (num_frames) *num_frames = 1;
        if (out && out_cap >= c.n_mels)
            for (int m = 0; m < c.n_mels; ++m) out[m] = 0.0f;
        return c.n_mels;
    }
    const int64_t Tp = ((T_frames + pad_to - 1) / pad_to) * pad_to;
    const int64_t need = Tp * c.n_mels;
    if (mel_length) *mel_length = T_frames;
    if (num_frames) *num_frames = Tp;
    if (!out || out_cap < need) return need;

    const int64_t pad = (mode == 0) ? c.n_fft / 2 : 0;
    const int64_t padded = n + 2 * pad;
    std::vector<T> p((size_t)padded, (T)0);
    if (c.preemph == 0.0f) {
        for (int64_t i = 0; i < n; ++i) p[pad + i] = (T)audio[i];
    } else {
        if (sizeof(T) == 4) {
            p[pad] = (T)(audio[0] - c.preemph * last);
            const float neg = -c.preemph;
            for (int64_t i = 1; i < n; ++i) p[pad + i] = (T)fmaf(audio[i - 1], neg, audio[i]);  // vDSP_vsma
        } else {
            p[pad] = (T)((double)audio[0] - (double)c.preemph * (double)last);
            for (int64_t i = 1; i < n; ++i) p[pad + i] = (T)((double)audio[i] - (double)c.preemph * (double)audio[i - 1]);
        }
    }
    std::fill(out, out + need, 0.0f);
    MelCore<T> core(c);
    const int64_t off = (c.n_fft - c.win_length) / 2;
    for (int64_t f = 0; f < T_frames; ++f) {
        const int64_t start = f * c.hop_length + off;
        const int64_t avail = std::min<int64_t>(c.win_length, padded - start);
        std::fill(core.frame.begin(), core.frame.end(), (T)0);
        for (int64_t k = 0; k < avail; ++k) core.frame[off + k] = p[start + k] * (T)core.window[k];
        if (layout == 0) {
            core.finish_frame([&](int m, T v) { out[f * c.n_mels + m] = (float)v; });
        } else {
            core.finish_frame([&](int m, T v) { out[(int64_t)m * Tp + f] = (float)v; });
        }
    }
    return need;
}

// compute(audio:) (:132-178): returns [nMels x T] (the [1] batch dim is implicit).
template <typename T>
int64_t run_legacy(const oracle_mel_config &c, const float *audio, int64_t n, float *out, int64_t out_cap,
                   int64_t *mel_length) {
    const int64_t T_frames = oracle_mel_frame_count(&c, n, 2, -1);
    if (T_frames <= 0) {
        if (mel_length) *mel_length = 0;
        return 0;
    }
    if (mel_length) *mel_length = T_frames;
    const int64_t need = T_frames * c.n_mels;
    if (!out || out_cap < need) return need;
    MelCore<T> core(c);
    for (int64_t f = 0; f < T_frames; ++f) {
        const int64_t start = f * c.hop_length;
        std::fill(core.frame.begin(), core.frame.end(), (T)0);
        for (int64_t i = 0; i < c.win_length; ++i) {
            const int64_t idx = start + i;
            if (idx < n) core.frame[i] = (T)audio[idx] * (T)core.window[i];
        }
        core.finish_frame([&](int m, T v) { out[(int64_t)m * T_frames + f] = (float)v; });
    }
    return need;
}

} // namespace

extern "C" {

// Returns the number of floats the output needs; writes only if out_cap is sufficient.
int64_t oracle_mel_compute_flat_transposed(const oracle_mel_config *c, const float *audio, int64_t n,
                                           float last_sample, int32_t padding_mode, int64_t expected_frames,
                                           float *out, int64_t out_cap, int64_t *mel_length,
                                           int64_t *num_frames) {
    if (c->precision == 1)
        return run_flat<double>(*c, audio, n, last_sample, padding_mode, expected_frames, 0, out, out_cap,
                                mel_length, num_frames);
    return run_flat<float>(*c, audio, n, last_sample, padding_mode, expected_frames, 0, out, out_cap, mel_length,
                           num_frames);
}

int64_t oracle_mel_compute_flat(const oracle_mel_config *c, const float *audio, int64_t n, float last_sample,
                                float *out, int64_t out_cap, int64_t *mel_length, int64_t *num_frames) {
    if (c->precision == 1)
        return run_flat<double>(*c, audio, n, last_sample, 0, -1, 1, out, out_cap, mel_length, num_frames);
    return run_flat<float>(*c, audio, n, last_sample, 0, -1, 1, out, out_cap, mel_length, num_frames);
}

int64_t oracle_mel_compute_legacy(const oracle_mel_config *c, const float *audio, int64_t n, float *out,
                                  int64_t out_cap, int64_t *mel_length) {
    if (c->precision == 1) return run_legacy<double>(*c, audio, n, out, out_cap, mel_length);
    return run_legacy<float>(*c, audio, n, out, out_cap, mel_length);
}

// UnifiedMelExtractor.normalizePerFeature (Sources/FluidAudio/ASR/Parakeet/Unified/UnifiedMelExtractor.swift:88-113)
// is restated in oracle_adapters.cpp.

} // extern "C"
