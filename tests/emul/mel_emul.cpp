// Host lane-emulator for fluidaudio_b200/csrc/mel_core.cuh (CPU test-suite only).
// Runs the exact per-lane functions the CUDA kernel runs, one lane at a time, phase by phase, for every
// frame of a clip, so the index math / twiddles / bank layout of the device code is checked without a GPU.
//   extern "C" int mel_emul(audio, n, last, hop, win, off, pad, preemph, n_mels, fb[n_mels*257],
//                           window[win], log_floor, clamped, T, out[T*n_mels])         FP64 transform, one frame per warp
//   extern "C" int mel_emul_f32x2(... same ...)                                        float32 pairs, two frames per warp
#include "../../fluidaudio_b200/csrc/mel_core.cuh"
#include <cmath>
#include <vector>
#include <cstring>

using namespace fa::mel;

template <typename V>
static int mel_emul_t(const float *audio, long long n, float last, int hop, int win, int off, int pad, float preemph,
                      int n_mels, const float *fb, const float *window, float log_floor, int clamped, long long T,
                      float *out) {
    constexpr int kF = vtraits<V>::kFrames;
    if (hop & 1) return 1;
    std::vector<float> win_tab(kNfft, 0.0f);
    std::vector<uint8_t> in_tab(kNfft, 0);
    for (int j = 0; j < kNfft; ++j)
        if (j >= off && j < off + win) {
            win_tab[j] = window[j - off];
            in_tab[j] = 1;
        }
    std::vector<LaneTables<V>> tabs(32);
    for (int l = 0; l < 32; ++l) load_lane_tables(l, win_tab.data(), in_tab.data(), tabs[l]);
    const bool mid_full = off <= 64 && off + win >= 448;
    std::vector<float> fbq((size_t)n_mels * kBins);   // the kernel's weights: filterbank / 4 (its power tile holds 4|X|^2)
    for (size_t i = 0; i < fbq.size(); ++i) fbq[i] = 0.25f * fb[i];
    // sparse filterbank ranges
    std::vector<int> lo(n_mels), hi(n_mels);
    for (int m = 0; m < n_mels; ++m) {
        int a = kBins, b = 0;
        for (int k = 0; k < kBins; ++k)
            if (fb[(size_t)m * kBins + k] != 0.0f) {
                a = std::min(a, k);
                b = k + 1;
            }
        if (b == 0) a = 0;
        lo[m] = a;
        hi[m] = b;
    }
    alignas(16) cpxv<V> buf[kFftPad];
    std::vector<float> pfv((size_t)kNfft + hop + 8);
    float *pf = pfv.data();
    std::vector<float> prow2(kPairStride), prow(2 * 260);   // the device's pair row, de-interleaved for the host dot product
    for (long long f = 0; f < T; f += kF) {
        for (int j = 0; j < kNfft + (kF - 1) * hop; ++j) {
            const long long i = f * hop + j - pad;
            float v = 0.0f;
            if (i >= 0 && i < n) {
                if (preemph == 0.0f) v = audio[i];
                else if (i == 0) v = preemph_first(audio[0], last, preemph);
                else v = preemph_rest(audio[i], audio[i - 1], preemph);
            }
            pf[j] = v;
        }
        std::memset((void *)buf, 0, sizeof(buf));
        V re[32][8], im[32][8];
        for (int l = 0; l < 32; ++l) {
            if (mid_full) pass1<true>(l, pf, hop, tabs[l], buf); else pass1<false>(l, pf, hop, tabs[l], buf);
        }
        for (int l = 0; l < 32; ++l) pass2_load(l, buf, re[l], im[l]);
        for (int l = 0; l < 32; ++l) pass2_store(l, tabs[l], re[l], im[l], buf);
        for (int l = 0; l < 32; ++l) pass3_post(l, buf, tabs[l], prow2.data());
        for (int b = 0; b < kBins; ++b) {
            prow[b] = prow2[2 * pow_pos(b)];
            prow[260 + b] = prow2[2 * pow_pos(b) + 1];
        }
        for (int k = 0; k < kF && f + k < T; ++k)
            for (int m = 0; m < n_mels; ++m) {
                const float acc = mel_dot(prow.data() + k * 260, fbq.data() + (size_t)m * kBins + lo[m], lo[m], hi[m]);
                out[(f + k) * n_mels + m] = log_value(acc, log_floor, clamped);
            }
    }
    return 0;
}

extern "C" int mel_emul(const float *audio, long long n, float last, int hop, int win, int off, int pad, float preemph,
                        int n_mels, const float *fb, const float *window, float log_floor, int clamped, long long T,
                        float *out) {
    return mel_emul_t<double>(audio, n, last, hop, win, off, pad, preemph, n_mels, fb, window, log_floor, clamped, T, out);
}
extern "C" int mel_emul_f32x2(const float *audio, long long n, float last, int hop, int win, int off, int pad,
                              float preemph, int n_mels, const float *fb, const float *window, float log_floor,
                              int clamped, long long T, float *out) {
    return mel_emul_t<f32x2>(audio, n, last, hop, win, off, pad, preemph, n_mels, fb, window, log_floor, clamped, T, out);
}
