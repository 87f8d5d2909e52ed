// The thin per-caller adapters that sit directly behind AudioMelSpectrogram in the reference,
// as device epilogues of the mel kernel: the log-mel never leaves HBM between the STFT and the adapter.
//
//   unified:  UnifiedMelExtractor.features(window:validCount:)   ASR/Parakeet/Unified/UnifiedMelExtractor.swift:52-113
//             center-padded log-mel with a fixed frame count, NeMo per-feature (per mel bin) mean / unbiased-std
//             normalisation over the valid frames, pad frames zeroed, packed [1, nMels, T].
//   lseend:   LSEENDPreprocessor.processAudioQueue            Diarizer/LS-EEND/LSEENDPreprocessor.swift:249-283
//             .prePadded log-mel (preemph 0, periodic Hann, clamped floor) * 1/ln(10), cumulative mean normalisation
//             with state (mean per mel, frame count) carried from call to call.
//
// Both normalisations are sequential in time by definition (the reference's loops) and independent across mel bins:
// one thread per mel bin walks the frames in order with individually rounded float32 operations, so the result is the
// reference's arithmetic exactly; loads are coalesced across bins.  Sizes are tiny (<= 512 bins x a few thousand
// frames), the point is fusion with the producer, not throughput.
#include "fa_common.cuh"
#include "mel_plan.h"

#include <algorithm>
#include <cmath>
#include <cuda_runtime.h>

namespace fa {
namespace mel {

#define FA_CUDA_TRY(expr)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            fa::set_error("CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return FA_CUDA_ERROR;                                                              \
        }                                                                                      \
    } while (0)

constexpr int kBinsPerCta = 128;
constexpr int kTile = 32;

// x: time-major [T x M]; out: mel-major [M x T].  valid >= 1.
__global__ void __launch_bounds__(kBinsPerCta) per_feature_norm_kernel(const float *__restrict__ x, long long T, int M,
                                                                       long long valid, float *__restrict__ out) {
    __shared__ float tile[kBinsPerCta][kTile + 1];
    const int m0 = blockIdx.x * kBinsPerCta, m = m0 + threadIdx.x;
    const bool live = m < M;
    float mean = 0.0f, sd = 1.0f;
    if (live) {
        for (long long t = 0; t < valid; ++t) mean = __fadd_rn(mean, x[t * M + m]);
        mean = __fdiv_rn(mean, (float)valid);
        float var_sum = 0.0f;
        for (long long t = 0; t < valid; ++t) {
            const float d = __fsub_rn(x[t * M + m], mean);
            var_sum = __fadd_rn(var_sum, __fmul_rn(d, d));
        }
        const float denom = (float)(valid > 1 ? valid - 1 : 1);
        sd = __fadd_rn(__fsqrt_rn(__fdiv_rn(var_sum, denom)), 1e-5f);
    }
    const int rows = min(kBinsPerCta, M - m0);
    for (long long t0 = 0; t0 < T; t0 += kTile) {
        const int nt = (int)min((long long)kTile, T - t0);
        if (live)
            for (int i = 0; i < nt; ++i) {
                const long long t = t0 + i;
                tile[threadIdx.x][i] = t < valid ? __fdiv_rn(__fsub_rn(x[t * M + m], mean), sd) : 0.0f;
            }
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int r = warp; r < rows; r += kBinsPerCta / 32)
            if (lane < nt) out[(long long)(m0 + r) * T + t0 + lane] = tile[r][lane];
        __syncthreads();
    }
}

// Same statistics, time-major in place (the standalone UnifiedMelExtractor.normalizePerFeature entry point of the C ABI).
__global__ void per_feature_norm_inplace_kernel(float *x, long long T, int M, long long valid) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float mean = 0.0f;
    for (long long t = 0; t < valid; ++t) mean = __fadd_rn(mean, x[t * M + m]);
    mean = __fdiv_rn(mean, (float)valid);
    float var_sum = 0.0f;
    for (long long t = 0; t < valid; ++t) {
        const float d = __fsub_rn(x[t * M + m], mean);
        var_sum = __fadd_rn(var_sum, __fmul_rn(d, d));
    }
    const float denom = (float)(valid > 1 ? valid - 1 : 1);
    const float sd = __fadd_rn(__fsqrt_rn(__fdiv_rn(var_sum, denom)), 1e-5f);
    for (long long t = 0; t < T; ++t) x[t * M + m] = t < valid ? __fdiv_rn(__fsub_rn(x[t * M + m], mean), sd) : 0.0f;
}

// host buffer in, host buffer out (x: [T x M] time-major, normalised in place); valid >= 1
int normalize_per_feature_host(float *x, long long T, int M, long long valid) {
    struct Buf {
        float *d = nullptr;
        ~Buf() { if (d) cudaFree(d); }
    } b;
    const size_t bytes = sizeof(float) * (size_t)T * M;
    FA_CUDA_TRY(cudaMalloc(&b.d, bytes));
    FA_CUDA_TRY(cudaMemcpy(b.d, x, bytes, cudaMemcpyHostToDevice));
    per_feature_norm_inplace_kernel<<<(M + kBinsPerCta - 1) / kBinsPerCta, kBinsPerCta>>>(b.d, T, M, valid);
    FA_CUDA_TRY(cudaGetLastError());
    FA_CUDA_TRY(cudaMemcpy(x, b.d, bytes, cudaMemcpyDeviceToHost));
    return FA_OK;
}

// x: time-major [T x M], in place; state: mean[M] (in/out), count (in: frames seen before this call)
__global__ void lseend_scale_cmn_kernel(float *x, long long T, int M, float *mean_io, long long count0, float scale) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float mean = mean_io[m];
    for (long long t = 0; t < T; ++t) {
        const float alpha = __fdiv_rn(1.0f, (float)(count0 + t + 1));
        const float v = __fmul_rn(x[t * M + m], scale);
        mean = __fadd_rn(mean, __fmul_rn(alpha, __fsub_rn(v, mean)));
        x[t * M + m] = __fsub_rn(v, mean);
    }
    mean_io[m] = mean;
}

int unified_features(MelPlan &p, const float *window, long long n, long long valid_count, float *out, long long out_len,
                     long long *total_frames, int *valid_frames) {
    const int M = p.cfg.n_mels, hop = p.cfg.hop_length;
    const long long T = n / hop + 1;                                   // UnifiedMelExtractor.swift:30
    const long long valid = std::min<long long>(valid_count / hop, T); // :71
    if (total_frames) *total_frames = T;
    if (valid_frames) *valid_frames = (int)valid;
    if (out_len < T * M) {
        fa::set_error("unified mel features need %lld floats, buffer has %lld", T * M, out_len);
        return FA_OUTPUT_TOO_SMALL;
    }
    int st = p.ensure_staging((size_t)n + 16, (size_t)(2 * T * M));
    if (st != FA_OK) return st;
    cudaStream_t s = p.streams[1];
    float *d_flat = p.d_out, *d_pack = p.d_out + T * M;
    if (n) FA_CUDA_TRY(cudaMemcpyAsync(p.d_audio, window, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    long long ml = 0, nf = 0;
    st = p.compute_device(p.d_audio, n, 0.0f, 0, T, 0, d_flat, T * M, &ml, &nf, s);
    if (st != FA_OK) return st;
    if (valid <= 0) {
        FA_CUDA_TRY(cudaMemsetAsync(d_pack, 0, sizeof(float) * T * M, s));
    } else {
        per_feature_norm_kernel<<<(M + kBinsPerCta - 1) / kBinsPerCta, kBinsPerCta, 0, s>>>(d_flat, T, M, valid, d_pack);
        FA_CUDA_TRY(cudaGetLastError());
        ++p.launches;
    }
    FA_CUDA_TRY(cudaMemcpyAsync(out, d_pack, sizeof(float) * T * M, cudaMemcpyDeviceToHost, s));
    FA_CUDA_TRY(cudaStreamSynchronize(s));
    return FA_OK;
}

int lseend_features(MelPlan &p, const float *chunk, long long n, float *cmn_mean, long long *cmn_count, float *out,
                    long long out_len, long long *frames) {
    const int M = p.cfg.n_mels;
    const long long T = p.frame_count(n, 1, -1);
    if (frames) *frames = T;
    if (T <= 0) return FA_OK;
    if (out_len < T * M) {
        fa::set_error("LS-EEND features need %lld floats, buffer has %lld", T * M, out_len);
        return FA_OUTPUT_TOO_SMALL;
    }
    int st = p.ensure_staging((size_t)n + 16, (size_t)(T * M + M));
    if (st != FA_OK) return st;
    cudaStream_t s = p.streams[1];
    float *d_flat = p.d_out, *d_mean = p.d_out + T * M;
    FA_CUDA_TRY(cudaMemcpyAsync(p.d_audio, chunk, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    FA_CUDA_TRY(cudaMemcpyAsync(d_mean, cmn_mean, sizeof(float) * M, cudaMemcpyHostToDevice, s));
    long long ml = 0, nf = 0;
    st = p.compute_device(p.d_audio, n, 0.0f, 1, -1, 0, d_flat, T * M, &ml, &nf, s);
    if (st != FA_OK) return st;
    const float scale = 1.0f / logf(10.0f);   // LSEENDPreprocessor.swift:36, Float arithmetic
    lseend_scale_cmn_kernel<<<(M + 127) / 128, 128, 0, s>>>(d_flat, T, M, d_mean, *cmn_count, scale);
    FA_CUDA_TRY(cudaGetLastError());
    ++p.launches;
    FA_CUDA_TRY(cudaMemcpyAsync(out, d_flat, sizeof(float) * T * M, cudaMemcpyDeviceToHost, s));
    FA_CUDA_TRY(cudaMemcpyAsync(cmn_mean, d_mean, sizeof(float) * M, cudaMemcpyDeviceToHost, s));
    FA_CUDA_TRY(cudaStreamSynchronize(s));
    *cmn_count += T;
    return FA_OK;
}

} // namespace mel
} // namespace fa
