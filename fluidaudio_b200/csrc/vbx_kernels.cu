// VBx refinement, gamma-weighted centroids and cosine assignment in FP64 on sm_100a.
//
// Re-implements
//   Sources/FluidAudio/Diarizer/Offline/Clustering/VBxClustering.swift:167-664   (runVBx)
//   Sources/FluidAudio/Diarizer/Offline/Core/OfflineDiarizerManager.swift:613-691 (computeCentroids)
//   OfflineDiarizerManager.swift:789-883                                          (centroidScores/assignEmbeddings)
//
// All reductions over frames use a FIXED two-level order (sequential inside a chunk of frames, chunks folded in
// ascending order by one CTA), so results are bit-reproducible run to run (the reference asserts its cluster
// phase is bit-identical across repeats, OfflineDiarizerTwoPhaseTests.swift:20-33).  The reference's own sums
// come from closed-source BLAS/vDSP with unknown association, so gamma/pi/ELBO parity is to 1e-9-ish, hard
// labels exact.  The whole EM loop runs without a host round trip: a device-side `done` flag turns the kernels
// of the remaining iterations into no-ops once |dELBO| < epsilon.
#include "vbx_plan.h"

#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

namespace fa {
namespace vbx {

#define FA_CUDA_TRY(expr)                                                                               \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            fa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return e__ == cudaErrorMemoryAllocation ? FA_ALLOCATION_FAILURE : FA_CUDA_ERROR;            \
        }                                                                                               \
    } while (0)

constexpr int kChunks = 128;      // frame chunks for the two-level reductions (fewer when there are > 1024 speakers)
constexpr int kEThreads = 128;    // threads per CTA in the E-step

// Partial sums are [chunks x S x D]: keep them bounded when AHC hands over thousands of clusters (degenerate input:
// every embedding its own speaker).  The chunk count only changes the (fixed) summation order.
static int chunks_for(int S) { return S <= 1024 ? kChunks : std::max(1, (kChunks * 1024) / S); }

struct Dev {
    int T, D, S, Tp, chunks;
    const double *x;      // [T x D] features
    const double *phi_c;  // [D] clamped psi
    double *rho;          // [T x D]
    double *rhoT;         // [D x Tp]
    double *G;            // [T]
    double *gamma;        // [T x S]
    double *pi;           // [S]
    double *invL, *alpha; // [S x D]
    double *phiTerm, *logPi, *gsum; // [S]
    double *pA;           // [chunks x S x D]
    double *pG;           // [chunks x S]
    double *pLL;          // [blocks]
    double *pPi;          // [blocks x S]
    double *elbos;        // [max_it]
    double *scal;         // [0]=sumLogInv [1]=sumInv [2]=sumAlphaSq [3]=prevElbo
    int *state;           // [0]=done [1]=iterations
    int eblocks;
    double Fa, Fb, eps;
};

// gamma0 = softmax(7 * onehot) row-wise, then renormalise (VBxClustering.swift:190-235); rho = x * sqrt(phi);
// G = -0.5 (|x|^2 + D ln 2 pi) (:239-282)
__global__ void vbx_init_kernel(Dev d, const int *__restrict__ init, double smoothing) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.T) return;
    const int S = d.S;
    double *g = d.gamma + (size_t)t * S;
    const int label = init ? max(0, min(init[t], S - 1)) : -1;
    for (int s = 0; s < S; ++s) g[s] = init ? (s == label ? 1.0 : 0.0) : 1.0 / (double)S;
    if (smoothing >= 0.0) {
        double mx = -1.7976931348623157e308;
        for (int s = 0; s < S; ++s) mx = fmax(mx, g[s] * smoothing);
        double sum = 0.0;
        for (int s = 0; s < S; ++s) {
            const double e = exp(g[s] * smoothing - mx);
            g[s] = e;
            sum += e;
        }
        if (sum <= 0.0 || !isfinite(sum)) {
            for (int s = 0; s < S; ++s) g[s] = 1.0 / (double)S;
        } else {
            const double inv = 1.0 / sum;
            for (int s = 0; s < S; ++s) g[s] *= inv;
        }
    }
    double sum = 0.0;
    for (int s = 0; s < S; ++s) sum += g[s];
    if (sum <= 0.0 || !isfinite(sum)) {
        for (int s = 0; s < S; ++s) g[s] = 1.0 / (double)S;
    } else {
        const double inv = 1.0 / sum;
        for (int s = 0; s < S; ++s) g[s] *= inv;
    }
    const double *x = d.x + (size_t)t * d.D;
    double ss = 0.0;
    for (int k = 0; k < d.D; ++k) {
        const double v = x[k];
        const double r = v * sqrt(d.phi_c[k]);
        d.rho[(size_t)t * d.D + k] = r;
        d.rhoT[(size_t)k * d.Tp + t] = r;
        ss += v * v;
    }
    d.G[t] = -0.5 * (ss + (double)d.D * log(2.0 * 3.14159265358979323846));
}

// per chunk of frames: pG[c][s] = sum_t gamma[t][s], pA[c][s][k] = sum_t gamma[t][s] rho[t][k]   (:304-360)
__global__ void vbx_accumulate_kernel(Dev d) {
    if (d.state[0]) return;
    const int c = blockIdx.x;
    const int per = (d.T + d.chunks - 1) / d.chunks;
    const int t0 = c * per, t1 = min(d.T, t0 + per);
    const int SD = d.S * d.D;
    for (int o = threadIdx.x; o < SD; o += blockDim.x) {
        const int s = o / d.D, k = o % d.D;
        double acc = 0.0;
        for (int t = t0; t < t1; ++t) acc += d.gamma[(size_t)t * d.S + s] * d.rho[(size_t)t * d.D + k];
        d.pA[(size_t)c * SD + o] = acc;
    }
    for (int s = threadIdx.x; s < d.S; s += blockDim.x) {
        double acc = 0.0;
        for (int t = t0; t < t1; ++t) acc += d.gamma[(size_t)t * d.S + s];
        d.pG[(size_t)c * d.S + s] = acc;
    }
}

// single CTA: fold chunks, invL, alpha, phiTerm, log pi and the three ELBO sums   (:330-436, :496-516, :623-644)
__global__ void vbx_update_kernel(Dev d) {
    if (d.state[0]) return;
    const int S = d.S, D = d.D, SD = S * D;
    const double ratio = d.Fa / d.Fb;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < d.chunks; ++c) acc += d.pG[(size_t)c * S + s];
        d.gsum[s] = acc;
        d.logPi[s] = log(fmax(d.pi[s], 1e-8));
    }
    __syncthreads();
    for (int o = threadIdx.x; o < SD; o += blockDim.x) {
        const int s = o / D, k = o % D;
        double acc = 0.0;
        for (int c = 0; c < d.chunks; ++c) acc += d.pA[(size_t)c * SD + o];
        const double il = 1.0 / fmax(1.0 + (ratio * d.gsum[s]) * d.phi_c[k], 1e-12);
        d.invL[o] = il;
        d.alpha[o] = (acc * il) * ratio;
    }
    __syncthreads();
    __shared__ double sh[3][256];
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        double p = 0.0;
        for (int k = 0; k < D; ++k) {
            const double a = d.alpha[s * D + k], il = d.invL[s * D + k];
            p += (a * a + il) * d.phi_c[k];
        }
        d.phiTerm[s] = p;
    }
    // ELBO sums: thread-strided partials in a fixed order, then a sequential fold by thread 0
    double l = 0.0, i2 = 0.0, a2 = 0.0;
    for (int o = threadIdx.x; o < SD; o += blockDim.x) {
        const double il = d.invL[o], a = d.alpha[o];
        l += log(il);
        i2 += il;
        a2 += a * a;
    }
    sh[0][threadIdx.x] = l;
    sh[1][threadIdx.x] = i2;
    sh[2][threadIdx.x] = a2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double x0 = 0, x1 = 0, x2 = 0;
        for (int i = 0; i < (int)blockDim.x; ++i) {
            x0 += sh[0][i];
            x1 += sh[1][i];
            x2 += sh[2][i];
        }
        d.scal[0] = x0;
        d.scal[1] = x1;
        d.scal[2] = x2;
    }
}

// thread per frame: log-likelihood row, soft-max -> gamma, per-CTA partial LL and partial pi   (:438-602)
// kAlphaSmem: alpha [S x D], -phiTerm/2 and log pi staged in shared memory (the normal case, S x D x 8 <= 200 KB);
// otherwise read through L2 (hundreds of speakers: same arithmetic, every warp reads the same addresses).
template <bool kAlphaSmem>
__global__ void __launch_bounds__(kEThreads) vbx_estep_kernel(Dev d) {
    if (d.state[0]) return;
    extern __shared__ double sm[];
    const int S = d.S, D = d.D;
    double *red = sm;                                        // [kEThreads]
    const double *alpha = d.alpha;
    if (kAlphaSmem) {
        double *a_s = sm + kEThreads;                        // [S x D]
        double *off_s = a_s + (size_t)S * D;                 // [S]  -0.5 phiTerm
        double *lpi_s = off_s + S;                           // [S]
        for (int o = threadIdx.x; o < S * D; o += kEThreads) a_s[o] = d.alpha[o];
        for (int s = threadIdx.x; s < S; s += kEThreads) {
            off_s[s] = d.phiTerm[s] * -0.5;
            lpi_s[s] = d.logPi[s];
        }
        __syncthreads();
        alpha = a_s;
    }
    const double *off = kAlphaSmem ? sm + kEThreads + (size_t)S * D : nullptr;
    const double *lpi = kAlphaSmem ? off + S : nullptr;
    const int t = blockIdx.x * kEThreads + threadIdx.x;
    double ll = 0.0;
    if (t < d.T) {
        double *g = d.gamma + (size_t)t * S;
        const double Gt = d.G[t];
        double mx = -1.7976931348623157e308;
        for (int s = 0; s < S; ++s) {
            double acc = 0.0;
            const double *a = alpha + (size_t)s * D;
            for (int k = 0; k < D; ++k) acc += d.rhoT[(size_t)k * d.Tp + t] * a[k];
            const double o_s = kAlphaSmem ? off[s] : d.phiTerm[s] * -0.5;
            const double l_s = kAlphaSmem ? lpi[s] : d.logPi[s];
            const double v = ((acc + o_s) + Gt) * d.Fa + l_s;
            g[s] = v;
            mx = fmax(mx, v);
        }
        double sum = 0.0;
        for (int s = 0; s < S; ++s) {
            const double e = exp(g[s] - mx);
            g[s] = e;
            sum += e;
        }
        if (sum <= 0.0 || !isfinite(sum)) {
            for (int s = 0; s < S; ++s) g[s] = 1.0 / (double)S;
            ll = mx;
        } else {
            const double inv = 1.0 / sum;
            for (int s = 0; s < S; ++s) g[s] *= inv;
            ll = mx + log(sum);
        }
    }
    red[threadIdx.x] = ll;
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int i = 0; i < kEThreads; ++i) acc += red[i];
        d.pLL[blockIdx.x] = acc;
    }
    // partial pi: column sums of this CTA's rows, sequential over rows
    // (the barrier above also makes every gamma row of this CTA visible to all of its threads)
    const int t0 = blockIdx.x * kEThreads, t1 = min(d.T, t0 + kEThreads);
    for (int s = threadIdx.x; s < S; s += kEThreads) {
        double acc = 0.0;
        for (int tt = t0; tt < t1; ++tt) acc += d.gamma[(size_t)tt * S + s];
        d.pPi[(size_t)blockIdx.x * S + s] = acc;
    }
}

// single CTA: LL, pi, ELBO, convergence   (:578-661)
__global__ void vbx_finish_kernel(Dev d, int iteration) {
    if (d.state[0]) return;
    const int S = d.S;
    __shared__ double piSum;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        double acc = 0.0;
        for (int b = 0; b < d.eblocks; ++b) acc += d.pPi[(size_t)b * S + s];
        d.pi[s] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ps = 0.0;
        for (int s = 0; s < S; ++s) ps += d.pi[s];
        piSum = ps;
    }
    __syncthreads();
    const double ps = piSum;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        if (ps > 0.0 && isfinite(ps)) d.pi[s] *= (1.0 / ps);
        else d.pi[s] = 1.0 / (double)S;
    }
    if (threadIdx.x == 0) {
        double ll = 0.0;
        for (int b = 0; b < d.eblocks; ++b) ll += d.pLL[b];
        const double count = (double)(S * d.D);
        const double elbo = ll + d.Fb * 0.5 * (d.scal[0] - d.scal[1] - d.scal[2] + count);
        d.elbos[iteration] = elbo;
        d.state[1] = iteration + 1;
        if (iteration > 0 && fabs(elbo - d.scal[3]) < d.eps) d.state[0] = 1;
        d.scal[3] = elbo;
    }
}

// first maximum wins (VBxClustering.swift:144-146)
__global__ void vbx_hard_kernel(const double *__restrict__ gamma, int T, int S, int *__restrict__ hard) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const double *g = gamma + (size_t)t * S;
    int best = 0;
    for (int s = 1; s < S; ++s)
        if (g[best] < g[s]) best = s;
    hard[t] = best;
}

// ---- centroids ----------------------------------------------------------------------------------------
// per chunk: num[c][s][k] = sum_t (gamma>0) gamma[t][s] e[t][k], den[c][s] = sum_t gamma   (:642-674)
__global__ void centroid_accumulate_kernel(const double *__restrict__ emb, const double *__restrict__ gamma, int T,
                                           int E, int S, int chunks, double *__restrict__ pnum,
                                           double *__restrict__ pden) {
    const int c = blockIdx.x;
    const int per = (T + chunks - 1) / chunks;
    const int t0 = c * per, t1 = min(T, t0 + per);
    const int SE = S * E;
    for (int o = threadIdx.x; o < SE; o += blockDim.x) {
        const int s = o / E, k = o % E;
        double acc = 0.0;
        for (int t = t0; t < t1; ++t) {
            const double w = gamma[(size_t)t * S + s];
            if (w > 0.0) acc += w * emb[(size_t)t * E + k];
        }
        pnum[(size_t)c * SE + o] = acc;
    }
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        double acc = 0.0;
        for (int t = t0; t < t1; ++t) {
            const double w = gamma[(size_t)t * S + s];
            if (w > 0.0) acc += w;
        }
        pden[(size_t)c * S + s] = acc;
    }
}

// single CTA: speakers with pi > 1e-7 in ascending order become centroids 0..K-1; also L2-normalised copies
__global__ void centroid_finish_kernel(const double *__restrict__ pnum, const double *__restrict__ pden,
                                       const double *__restrict__ pi, int E, int S, int chunks, int *__restrict__ map,
                                       double *__restrict__ cent, double *__restrict__ cent_n, int *__restrict__ count) {
    __shared__ int K;
    if (threadIdx.x == 0) {
        int k = 0;
        for (int s = 0; s < S; ++s) map[s] = (pi[s] > 1e-7) ? k++ : -1;
        K = k;
        *count = k;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < S * E; o += blockDim.x) {
        const int s = o / E, k = o % E;
        if (map[s] < 0) continue;
        double num = 0.0, den = 0.0;
        for (int c = 0; c < chunks; ++c) {
            num += pnum[(size_t)c * S * E + o];
            den += pden[(size_t)c * S + s];
        }
        cent[(size_t)map[s] * E + k] = den > 0.0 ? num / den : 0.0;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < K; c += blockDim.x) {   // normalize (:824-860): unchanged if |c|^2 <= 0
        const double *v = cent + (size_t)c * E;
        double ss = 0.0;   // individually rounded operations, like the reference's scalar loops (no FMA contraction)
        for (int k = 0; k < E; ++k) ss = __dadd_rn(ss, __dmul_rn(v[k], v[k]));
        const double sc = ss > 0.0 ? __ddiv_rn(1.0, __dsqrt_rn(ss)) : 1.0;
        for (int k = 0; k < E; ++k) cent_n[(size_t)c * E + k] = __dmul_rn(v[k], sc);
    }
}

// thread per embedding: cosine against every centroid, strict '>' (:800-822).  scores optional [N x K].
__global__ void assign_kernel(const double *__restrict__ emb, int N, int E, const double *__restrict__ cent_n,
                              const int *__restrict__ count_ptr, int K_fixed, int *__restrict__ labels,
                              double *__restrict__ scores) {
    const int K = count_ptr ? *count_ptr : K_fixed;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (K <= 0) {
        labels[n] = 0;
        return;
    }
    const double *e = emb + (size_t)n * E;
    double ss = 0.0;   // individually rounded operations in the reference's order: exact ties resolve the same way
    for (int k = 0; k < E; ++k) ss = __dadd_rn(ss, __dmul_rn(e[k], e[k]));
    const double sc = ss > 0.0 ? __ddiv_rn(1.0, __dsqrt_rn(ss)) : 1.0;
    int best = 0;
    double best_score = -INFINITY;
    for (int c = 0; c < K; ++c) {
        const double *v = cent_n + (size_t)c * E;
        double dot = 0.0;
        for (int k = 0; k < E; ++k) dot = __dadd_rn(dot, __dmul_rn(__dmul_rn(e[k], sc), v[k]));
        if (scores) scores[(size_t)n * K + c] = dot;
        if (dot > best_score) {
            best_score = dot;
            best = c;
        }
    }
    labels[n] = best;
}

__global__ void onehot_kernel(const int *__restrict__ labels, int T, int S, double *__restrict__ gamma,
                              double *__restrict__ pi) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < S) pi[t] = 1.0;
    if (t >= T) return;
    for (int s = 0; s < S; ++s) gamma[(size_t)t * S + s] = (labels[t] == s) ? 1.0 : 0.0;
}

__global__ void finite_rows_kernel(const float *__restrict__ emb, int N, int E, unsigned char *__restrict__ ok) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    bool fin = true;
    for (int k = 0; k < E; ++k) fin = fin && isfinite(emb[(size_t)n * E + k]);
    ok[n] = fin ? 1 : 0;
}

__global__ void gather_rows_kernel(const double *__restrict__ src, const int *__restrict__ idx, int rows, int dim,
                                   double *__restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * dim) return;
    const int r = (int)(i / dim), k = (int)(i % dim);
    dst[i] = src[(size_t)idx[r] * dim + k];
}

__global__ void mean_rows_kernel(const double *__restrict__ src, int rows, int dim, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= dim) return;
    double acc = 0.0;
    for (int r = 0; r < rows; ++r) acc += src[(size_t)r * dim + k];
    out[k] = acc * (1.0 / (double)rows);
}

// ------------------------------------------------------------------------------------------------ host side
Workspace::~Workspace() { release(); }
void Workspace::release() {
    if (pool) cudaFree(pool);
    pool = nullptr;
    pool_bytes = 0;
}

int Workspace::reserve(size_t bytes) {
    if (bytes <= pool_bytes) return FA_OK;
    if (pool) cudaFree(pool);
    pool = nullptr;
    pool_bytes = 0;
    FA_CUDA_TRY(cudaMalloc(&pool, bytes));
    pool_bytes = bytes;
    return FA_OK;
}

namespace {
struct Carver {
    char *base;
    size_t off = 0;
    template <typename T> T *take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T *p = reinterpret_cast<T *>(base + off);
        off += count * sizeof(T);
        return p;
    }
};
} // namespace

size_t refine_bytes(int T, int D, int S, int max_it) {
    Carver c{nullptr};
    const int Tp = (T + 31) & ~31;
    const int eblocks = (T + kEThreads - 1) / kEThreads;
    c.take<double>((size_t)T * D);
    c.take<double>((size_t)D * Tp);
    c.take<double>(T);
    c.take<double>(D);
    c.take<double>((size_t)S * D);
    c.take<double>((size_t)S * D);
    c.take<double>(S);
    c.take<double>(S);
    c.take<double>(S);
    c.take<double>((size_t)chunks_for(S) * S * D);
    c.take<double>((size_t)chunks_for(S) * S);
    c.take<double>(eblocks);
    c.take<double>((size_t)eblocks * S);
    c.take<double>(std::max(max_it, 1));
    c.take<double>(8);
    c.take<int>(8);
    return c.off + 512;
}

// d_x: [T x D] device, h_psi: [D] HOST (already identity-substituted by the caller if lengths mismatch),
// d_init: [T] device labels (or nullptr), d_gamma [T x S], d_pi [S], d_elbos [max(max_it,1)], d_hard [T].
int refine_device(Workspace &ws, const double *d_x, int T, int D, const double *h_psi, const int *d_init, int S,
                  const Config &cfg, double *d_gamma, double *d_pi, double *d_elbos, int *d_hard, int *iterations_host,
                  cudaStream_t stream, long long *launches) {
    if (T <= 0 || D <= 0 || S <= 0) return FA_INVALID_ARGUMENT;
    const int max_it = cfg.max_iterations;
    int st = ws.reserve(refine_bytes(T, D, S, max_it));
    if (st != FA_OK) return st;
    Carver c{static_cast<char *>(ws.pool)};
    Dev d{};
    d.T = T;
    d.D = D;
    d.S = S;
    d.Tp = (T + 31) & ~31;
    d.chunks = chunks_for(S);
    d.eblocks = (T + kEThreads - 1) / kEThreads;
    d.x = d_x;
    d.rho = c.take<double>((size_t)T * D);
    d.rhoT = c.take<double>((size_t)D * d.Tp);
    d.G = c.take<double>(T);
    double *phi_c = c.take<double>(D);
    d.phi_c = phi_c;
    d.invL = c.take<double>((size_t)S * D);
    d.alpha = c.take<double>((size_t)S * D);
    d.phiTerm = c.take<double>(S);
    d.logPi = c.take<double>(S);
    d.gsum = c.take<double>(S);
    d.pA = c.take<double>((size_t)d.chunks * S * D);
    d.pG = c.take<double>((size_t)d.chunks * S);
    d.pLL = c.take<double>(d.eblocks);
    d.pPi = c.take<double>((size_t)d.eblocks * S);
    (void)c.take<double>(std::max(max_it, 1));
    d.scal = c.take<double>(8);
    d.state = c.take<int>(8);
    d.gamma = d_gamma;
    d.pi = d_pi;
    d.elbos = d_elbos;
    d.Fa = cfg.Fa;
    d.Fb = cfg.Fb;
    d.eps = cfg.epsilon;

    // phi clamp (:239) and pi = 1/S (:237) via tiny host staging
    std::vector<double> h_pi(S, 1.0 / (double)S), h_scal(8, 0.0);
    h_scal[3] = -1.7976931348623157e308;
    FA_CUDA_TRY(cudaMemcpyAsync(d_pi, h_pi.data(), S * sizeof(double), cudaMemcpyHostToDevice, stream));
    FA_CUDA_TRY(cudaMemcpyAsync(d.scal, h_scal.data(), 8 * sizeof(double), cudaMemcpyHostToDevice, stream));
    FA_CUDA_TRY(cudaMemsetAsync(d.state, 0, 8 * sizeof(int), stream));
    FA_CUDA_TRY(cudaMemsetAsync(d_elbos, 0, std::max(max_it, 1) * sizeof(double), stream));
    {
        std::vector<double> tmp(D);
        for (int k = 0; k < D; ++k) tmp[k] = std::max(h_psi[k], 1e-12);
        FA_CUDA_TRY(cudaMemcpyAsync(phi_c, tmp.data(), D * sizeof(double), cudaMemcpyHostToDevice, stream));
        FA_CUDA_TRY(cudaStreamSynchronize(stream));   // tmp, h_pi and h_scal are pageable stack/heap buffers
    }
    vbx_init_kernel<<<(T + 127) / 128, 128, 0, stream>>>(d, d_init, cfg.init_smoothing);
    FA_CUDA_TRY(cudaGetLastError());
    long long n_launch = 1;
    const size_t esmem_full = sizeof(double) * ((size_t)S * D + 2 * S + kEThreads);
    const bool alpha_smem = esmem_full <= 200 * 1024;
    const size_t esmem = alpha_smem ? esmem_full : sizeof(double) * kEThreads;
    {
        static std::once_flag once;   // per-function attribute shared by concurrent callers: set once to the maximum
        static cudaError_t attr_err = cudaSuccess;
        std::call_once(once, [&]() {
            attr_err = cudaFuncSetAttribute(vbx_estep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        });
        FA_CUDA_TRY(attr_err);
    }
    for (int it = 0; it < max_it; ++it) {
        vbx_accumulate_kernel<<<d.chunks, 256, 0, stream>>>(d);
        vbx_update_kernel<<<1, 256, 0, stream>>>(d);
        if (alpha_smem) vbx_estep_kernel<true><<<d.eblocks, kEThreads, esmem, stream>>>(d);
        else vbx_estep_kernel<false><<<d.eblocks, kEThreads, esmem, stream>>>(d);
        vbx_finish_kernel<<<1, 256, 0, stream>>>(d, it);
        n_launch += 4;
    }
    FA_CUDA_TRY(cudaGetLastError());
    vbx_hard_kernel<<<(T + 127) / 128, 128, 0, stream>>>(d_gamma, T, S, d_hard);
    FA_CUDA_TRY(cudaGetLastError());
    ++n_launch;
    if (launches) *launches += n_launch;
    if (iterations_host) {
        int h_state[2] = {0, 0};
        FA_CUDA_TRY(cudaMemcpyAsync(h_state, d.state, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        FA_CUDA_TRY(cudaStreamSynchronize(stream));
        *iterations_host = h_state[1];
    }
    return FA_OK;
}

size_t centroid_bytes(int E, int S) {
    return ((size_t)chunks_for(S) * S * E + (size_t)chunks_for(S) * S) * sizeof(double) + (size_t)S * sizeof(int) + 2048;
}

int centroids_device(Workspace &ws, const double *d_emb, int T, int E, const double *d_gamma, const double *d_pi, int S,
                     double *d_cent, double *d_cent_n, int *d_count, cudaStream_t stream, long long *launches) {
    const int chunks = chunks_for(S);
    int st = ws.reserve(centroid_bytes(E, S));
    if (st != FA_OK) return st;
    Carver c{static_cast<char *>(ws.pool)};
    double *pnum = c.take<double>((size_t)chunks * S * E);
    double *pden = c.take<double>((size_t)chunks * S);
    int *map = c.take<int>(S);
    centroid_accumulate_kernel<<<chunks, 256, 0, stream>>>(d_emb, d_gamma, T, E, S, chunks, pnum, pden);
    centroid_finish_kernel<<<1, 256, 0, stream>>>(pnum, pden, d_pi, E, S, chunks, map, d_cent, d_cent_n, d_count);
    FA_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 2;
    return FA_OK;
}

int assign_device(const double *d_emb, int N, int E, const double *d_cent_n, const int *d_count, int K_fixed,
                  int *d_labels, double *d_scores, cudaStream_t stream, long long *launches) {
    if (N <= 0) return FA_OK;
    assign_kernel<<<(N + 127) / 128, 128, 0, stream>>>(d_emb, N, E, d_cent_n, d_count, K_fixed, d_labels, d_scores);
    FA_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return FA_OK;
}

int onehot_device(const int *d_labels, int T, int S, double *d_gamma, double *d_pi, cudaStream_t stream) {
    const int n = std::max(T, S);
    onehot_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_labels, T, S, d_gamma, d_pi);
    FA_CUDA_TRY(cudaGetLastError());
    return FA_OK;
}

int finite_rows_device(const float *d_emb, int N, int E, unsigned char *d_ok, cudaStream_t stream) {
    finite_rows_kernel<<<(N + 127) / 128, 128, 0, stream>>>(d_emb, N, E, d_ok);
    FA_CUDA_TRY(cudaGetLastError());
    return FA_OK;
}

int gather_rows_device(const double *d_src, const int *d_idx, int rows, int dim, double *d_dst, cudaStream_t stream) {
    const long long total = (long long)rows * dim;
    if (total <= 0) return FA_OK;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_src, d_idx, rows, dim, d_dst);
    FA_CUDA_TRY(cudaGetLastError());
    return FA_OK;
}

int mean_rows_device(const double *d_src, int rows, int dim, double *d_out, cudaStream_t stream) {
    mean_rows_kernel<<<(dim + 127) / 128, 128, 0, stream>>>(d_src, rows, dim, d_out);
    FA_CUDA_TRY(cudaGetLastError());
    return FA_OK;
}

} // namespace vbx
} // namespace fa
