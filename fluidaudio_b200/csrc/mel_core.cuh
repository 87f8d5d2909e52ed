// Per-lane math of the fused log-mel kernel (nFFT = 512), written so that the SAME source compiles for the
// device (mel_kernels.cu) and for the host lane-emulator used by the CPU test-suite (tests/emul/mel_emul.cpp):
// every function takes the lane id explicitly, touches "shared memory" only through the pointers it is given,
// and keeps no state between phases other than the read-only LaneTables.
//
// Reference being re-implemented: Sources/FluidAudio/Shared/AudioMelSpectrogram.swift:404-453 (frame loop),
// :459-481 (512-point DFT of a real frame, power of bins 0..256), :432-452 (filterbank mat-vec + log).
//
// Algorithm (one warp per frame):
//   real frame x[0..512) (float32: window * pre-emphasised sample, exactly the reference's vDSP_vmul)
//   z[m] = x[2m] + i x[2m+1], m < 256                                     (even/odd packing)
//   Z = FFT256(z): radix 8 x 8 x 4 decimation in frequency in FP64, through shared memory, __syncwarp between passes
//   X[b] = E[b] + W512^b O[b],  E = (Z[b] + conj Z[256-b]) / 2,  O = (Z[b] - conj Z[256-b]) / (2i)      (FP64)
//   power[b] = fl32(Re X)^2 + fl32(Im X)^2 in float32, b = 0..256
//   The last radix-4 pass and the recombination are ONE register-resident step: a lane computes the two radix-4
//   butterflies whose outputs are each other's mirror images (k and 256-k), so Z never goes back to shared memory.
//
// Why FP64 for the transform: the reference's DFT (vDSP_DFT_zop) is float32, and ANY float32 FFT carries rounding
// noise of ~0.5 ulp of the frame's largest spectral line in every bin.  On input with 60 dB of dynamic range that
// alone moves weak log-mel bins by up to ~1e-4, so two correct float32 implementations cannot be compared to
// 1e-4.  Evaluating the transform in FP64 and rounding once reproduces the implementation-independent value (the
// oracle's definition) to ~1e-6 in the log domain.  The kernel is instruction-issue bound, the FP64 pipe is not the
// limiter (profiles/), so this costs little.
//
// Shared-memory layouts of the 256 complex doubles (16-byte elements; a warp-wide 16-byte access is 4 wavefronts
// when every 8 consecutive lanes hit 8 distinct 16-byte banks):
//   A  (pass 1 out / pass 2 in):  element idx            at idx + 4*(idx>>5)          = l + 36 q
//   B  (pass 2 out / pass 3 in):  z_{q,q2}[h]            at 74 h + 9 q + q2
// Twiddles are produced by recurrence in FP64 from one per-lane root per pass (W256^l, W32^(l&3), W512^l).  The
// kernel is bound by shared-memory wavefronts (profiles/r01c_mel.md: the three transposes of 256 complex doubles cost
// ~200 of ~330 wavefronts per frame), the FP64 pipe has headroom: twiddle tables in shared memory were measured and
// bought nothing (80 fewer FP64 instructions, 67 more wavefronts per frame).
// The recombination handles the bin pair (b, 256-b) together — both need only Z[b] and Z[256-b]:
//   2 X[b] = S + T,  2 X[256-b] = conj(S - T),  S = Z[b] + conj Z[256-b],  T = W512^b * (D.y, -D.x),  D = Z[b] - conj Z[256-b]
// — works on 2X (no multiplications by 1/2) and stores 4*|X|^2; the plan hands the kernel the filterbank weights
// times 1/4, which gives bit-identical products (power-of-two scalings commute with rounding).
#pragma once

#include "fa_common.cuh"
#include <math.h>
#include <stdint.h>

namespace fa {
namespace mel {

constexpr int kNfft = 512;
constexpr int kHalf = 256;
constexpr int kBins = 257;
constexpr int kFftPad = 304;      // complex doubles per warp buffer (max layout extent 296 + Z[0] mirror at 288)
constexpr int kTileFrames = 16;   // frames per CTA tile; the mel stage maps 32 / kTileFrames mel bins onto one warp
constexpr int kPowStride = 260;   // rows 16-byte aligned; stride = 4 (mod 32) floats: 8 lanes x 16 B of one bin quad hit distinct banks

struct alignas(8) cpx {
    float x, y;
};
struct alignas(16) cpxd {
    double x, y;
};

// Constants a lane needs for every frame it processes; loaded once per kernel.
struct LaneTables {
    float win[16];     // window coefficient at buffer positions j = 2(l+32r) [slot 2r] and j+1 [slot 2r+1]
    uint32_t in_win;   // bit s set  <=>  slot s lies inside [off, off+win)
    cpxd w256;         // W256^l        (pass 1 root)
    cpxd w32;          // W32^(l & 3)   (pass 2 root)
    cpxd wk0;          // W512^k0       (recombination root of this lane's first butterfly)
    int a1, a2, k0;    // layout-B addresses of the lane's two mirror-image radix-4 butterflies; first output index
};

// Pass 3 + recombination: which two of the 64 radix-4 butterflies a lane owns.  Butterfly c = 8a + b reads
// z_{c}[h] (layout B) and produces Z[k0 + 64 k2], k0 = a + 8b.  The mirror bins 256 - k come out of the butterfly
// with k0' = 64 - k0: c' = 8(8-a) + (7-b) for a >= 1, c' = (8-b) mod 8 for a = 0.  Lanes 8..31 own (l, mirror(l)),
// lanes 1..3 own (1,7) (2,6) (3,5), lanes 4..7 own (32,39) .. (35,36), lane 0 owns the two self-mirrored butterflies
// 0 and 4.  Both 16-byte loads of a quarter-warp hit eight distinct banks ((a + b + 2h) mod 8).
FA_HD void lane_butterflies(int l, int &c1, int &c2) {
    if (l >= 8) {
        c1 = l;
        c2 = 8 * (8 - (l >> 3)) + (7 - (l & 7));
    } else if (l >= 4) {
        c1 = 32 + (l - 4);
        c2 = 39 - (l - 4);
    } else if (l >= 1) {
        c1 = l;
        c2 = 8 - l;
    } else {
        c1 = 0;
        c2 = 4;
    }
}

FA_HD cpxd unit_root(int k, int n) {   // exp(-2 pi i k / n) in FP64
    const double a = -6.283185307179586476925286766559 * (double)k / (double)n;
    cpxd r;
    r.x = cos(a);
    r.y = sin(a);
    return r;
}

// win_tab[512]: window value per buffer position (0 outside the window), in_tab[512]: 1 inside the window.
FA_HD void load_lane_tables(int l, const float *win_tab, const uint8_t *in_tab, LaneTables &T) {
    T.in_win = 0;
    for (int r = 0; r < 8; ++r) {
        const int j = 2 * (l + 32 * r);
        T.win[2 * r] = win_tab[j];
        T.win[2 * r + 1] = win_tab[j + 1];
        if (in_tab[j]) T.in_win |= 1u << (2 * r);
        if (in_tab[j + 1]) T.in_win |= 1u << (2 * r + 1);
    }
    T.w256 = unit_root(l, 256);
    T.w32 = unit_root(l & 3, 32);
    int c1, c2;
    lane_butterflies(l, c1, c2);
    T.a1 = 9 * (c1 >> 3) + (c1 & 7);
    T.a2 = 9 * (c2 >> 3) + (c2 & 7);
    T.k0 = (c1 >> 3) + 8 * (c1 & 7);
    T.wk0 = unit_root(T.k0, 512);
}

FA_HD cpxd cmul(cpxd a, cpxd b) {
    cpxd r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}

// forward 4-point DFT, in place, natural order
FA_HD void dft4(double &r0, double &i0, double &r1, double &i1, double &r2, double &i2, double &r3, double &i3) {
    const double ar = r0 + r2, ai = i0 + i2;
    const double br = r0 - r2, bi = i0 - i2;
    const double cr = r1 + r3, ci = i1 + i3;
    const double dr = r1 - r3, di = i1 - i3;
    r0 = ar + cr;
    i0 = ai + ci;
    r2 = ar - cr;
    i2 = ai - ci;
    r1 = br + di;   // d1 + (-i)(c1 - c3)
    i1 = bi - dr;
    r3 = br - di;
    i3 = bi + dr;
}

// forward 8-point DFT, in place, natural order (decimation in frequency: 4 radix-2 + two 4-point DFTs)
FA_HD void dft8(double (&re)[8], double (&im)[8]) {
    const double kS = 0.70710678118654752440;
    double er0 = re[0] + re[4], ei0 = im[0] + im[4];
    double er1 = re[1] + re[5], ei1 = im[1] + im[5];
    double er2 = re[2] + re[6], ei2 = im[2] + im[6];
    double er3 = re[3] + re[7], ei3 = im[3] + im[7];
    double or0 = re[0] - re[4], oi0 = im[0] - im[4];
    const double xr1 = re[1] - re[5], xi1 = im[1] - im[5];
    const double xr2 = re[2] - re[6], xi2 = im[2] - im[6];
    const double xr3 = re[3] - re[7], xi3 = im[3] - im[7];
    double or1 = (xr1 + xi1) * kS, oi1 = (xi1 - xr1) * kS;     // * (1 - i)/sqrt2
    double or2 = xi2, oi2 = -xr2;                              // * (-i)
    double or3 = (xi3 - xr3) * kS, oi3 = -(xr3 + xi3) * kS;    // * (-1 - i)/sqrt2
    dft4(er0, ei0, er1, ei1, er2, ei2, er3, ei3);
    dft4(or0, oi0, or1, oi1, or2, oi2, or3, oi3);
    re[0] = er0; im[0] = ei0;
    re[2] = er1; im[2] = ei1;
    re[4] = er2; im[4] = ei2;
    re[6] = er3; im[6] = ei3;
    re[1] = or0; im[1] = oi0;
    re[3] = or1; im[3] = oi1;
    re[5] = or2; im[5] = oi2;
    re[7] = or3; im[7] = oi3;
}

// multiply element q by root^q, q = 1..7, and hand the product to `emit(q, value)`.  The powers are built as a
// depth-3 tree (w2 = w*w, w3 = w2*w, w4 = w2*w2, w5 = w4*w, w6 = w4*w2, w7 = w4*w3) to keep the dependent chain short.
template <typename Emit>
FA_HD void twiddle_emit(double (&re)[8], double (&im)[8], cpxd root, Emit emit) {
    cpxd p[8];
    p[1] = root;
    p[2] = cmul(root, root);
    p[3] = cmul(p[2], root);
    p[4] = cmul(p[2], p[2]);
    p[5] = cmul(p[4], root);
    p[6] = cmul(p[4], p[2]);
    p[7] = cmul(p[4], p[3]);
    cpxd v;
    v.x = re[0];
    v.y = im[0];
    emit(0, v);
#pragma unroll
    for (int q = 1; q < 8; ++q) {
        v.x = re[q] * p[q].x - im[q] * p[q].y;
        v.y = re[q] * p[q].y + im[q] * p[q].x;
        emit(q, v);
    }
}

// Pass 1.  pf -> pre-emphasised sample at buffer position j = 0 of this frame (8-byte aligned, hop even).
// Window product in float32 (the reference's vDSP_vmul), then widened.  Output layout A.
// kMidFull: the window covers buffer positions [64, 448), so slots r = 1..6 of every lane are inside it and only the
// first and last slot need the in-window select (win 400 centred: positions 56..455).
template <bool kMidFull>
FA_HD void pass1(int l, const float *pf, const LaneTables &T, cpxd *buf) {
    double re[8], im[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int j = 2 * (l + 32 * r);
#if defined(__CUDA_ARCH__)
        const float2 v = *reinterpret_cast<const float2 *>(pf + j);   // one 64-bit shared load (LDS.64)
#else
        const cpx v = *reinterpret_cast<const cpx *>(pf + j);
#endif
        float a = T.win[2 * r] * v.x, b = T.win[2 * r + 1] * v.y;
        if (!kMidFull || r == 0 || r == 7) {   // outside the window the reference's buffer holds 0, whatever the sample
            a = (T.in_win >> (2 * r)) & 1u ? a : 0.0f;
            b = (T.in_win >> (2 * r + 1)) & 1u ? b : 0.0f;
        }
        re[r] = (double)a;
        im[r] = (double)b;
    }
    dft8(re, im);
    twiddle_emit(re, im, T.w256, [&](int q, cpxd v) { buf[l + 36 * q] = v; });
}

// Pass 2 is split: its output layout (B) differs from its input layout (A), every lane must finish loading first.
FA_HD void pass2_load(int l, const cpxd *buf, double (&re)[8], double (&im)[8]) {
    const int base = 36 * (l >> 2) + (l & 3);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const cpxd v = buf[base + 4 * r];
        re[r] = v.x;
        im[r] = v.y;
    }
}
FA_HD void pass2_store(int l, const LaneTables &T, double (&re)[8], double (&im)[8], cpxd *buf) {
    dft8(re, im);
    const int base = 74 * (l & 3) + 9 * (l >> 2);
    twiddle_emit(re, im, T.w32, [&](int q2, cpxd v) { buf[base + q2] = v; });
}

// Real-FFT recombination + power, one bin PAIR (b, 256 - b) per step (see the header).  prow -> this frame's row of
// the power tile; receives 4 |X[b]|^2 for b = 0..256.
FA_HD void pair_power(cpxd zb, cpxd zc, cpxd w, float &pb, float &pc) {
    const double sr = zb.x + zc.x, si = zb.y - zc.y;             // S
    const double dr = zb.y + zc.y, di = zc.x - zb.x;             // (D.y, -D.x)
    const double tr = w.x * dr - w.y * di, ti = w.x * di + w.y * dr;
    const float xr = (float)(sr + tr), xi = (float)(si + ti);    // single rounding of the exact-arithmetic DFT (x2)
    const float yr = (float)(sr - tr), yi = (float)(si - ti);
#if defined(__CUDA_ARCH__)
    pb = __fadd_rn(__fmul_rn(xr, xr), __fmul_rn(xi, xi));
    pc = __fadd_rn(__fmul_rn(yr, yr), __fmul_rn(yi, yi));
#else
    const float a = xr * xr, b = xi * xi, c = yr * yr, d = yi * yi;
    pb = a + b;
    pc = c + d;
#endif
}
// Pass 3 fused with the recombination (see lane_butterflies): A = butterfly c1 -> Z[k0 + 64 k2], B = butterfly c2 ->
// Z[(64 - k0) + 64 k2], so bin b = k0 + 64 k2 pairs A[k2] with B[3 - k2] and W512^b = W512^k0 * W8^k2.
// Lane 0 (k0 = 0; A = Z[0], Z[64], Z[128], Z[192]; B = Z[32], Z[96], Z[160], Z[224]) pairs inside its butterflies:
// (0,256 = Z[0]) (64,192) (128,128) in slots 0..2, (32,224) in slot 3 and (96,160) in one extra step.
FA_HD void pass3_post(int l, const cpxd *buf, const LaneTables &T, float *prow) {
    double ar[4], ai[4], br[4], bi[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const cpxd u = buf[T.a1 + 74 * h], v = buf[T.a2 + 74 * h];
        ar[h] = u.x;
        ai[h] = u.y;
        br[h] = v.x;
        bi[h] = v.y;
    }
    dft4(ar[0], ai[0], ar[1], ai[1], ar[2], ai[2], ar[3], ai[3]);
    dft4(br[0], bi[0], br[1], bi[1], br[2], bi[2], br[3], bi[3]);
    const bool z = l == 0;
    const double hh = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
    const cpxd w0 = T.wk0;
    cpxd w1, w2, w3;
    w1.x = hh * (w0.x + w0.y);   // * W8   = (1 - i)/sqrt2
    w1.y = hh * (w0.y - w0.x);
    w2.x = w0.y;                 // * W8^2 = -i
    w2.y = -w0.x;
    w3.x = w1.y;                 // * W8^3 = W8 * (-i)
    w3.y = -w1.x;
    cpxd zb, zc, w;
    float pb, pc;
    // slot 0
    zb.x = ar[0]; zb.y = ai[0];
    zc.x = z ? ar[0] : br[3]; zc.y = z ? ai[0] : bi[3];
    pair_power(zb, zc, w0, pb, pc);
    prow[T.k0] = pb;
    prow[kHalf - T.k0] = pc;
    // slot 1
    zb.x = ar[1]; zb.y = ai[1];
    zc.x = z ? ar[3] : br[2]; zc.y = z ? ai[3] : bi[2];
    pair_power(zb, zc, w1, pb, pc);
    prow[T.k0 + 64] = pb;
    prow[kHalf - 64 - T.k0] = pc;
    // slot 2
    zb.x = ar[2]; zb.y = ai[2];
    zc.x = z ? ar[2] : br[1]; zc.y = z ? ai[2] : bi[1];
    pair_power(zb, zc, w2, pb, pc);
    prow[T.k0 + 128] = pb;
    prow[kHalf - 128 - T.k0] = pc;
    // slot 3 (lane 0: bins 32 / 224, W512^32 = W16)
    zb.x = z ? br[0] : ar[3]; zb.y = z ? bi[0] : ai[3];
    zc.x = z ? br[3] : br[0]; zc.y = z ? bi[3] : bi[0];
    w.x = z ? c1 : w3.x;
    w.y = z ? -s1 : w3.y;
    pair_power(zb, zc, w, pb, pc);
    const int b3 = z ? 32 : T.k0 + 192;
    prow[b3] = pb;
    prow[kHalf - b3] = pc;
    if (z) {                     // bins 96 / 160, W512^96 = W16^3
        zb.x = br[1]; zb.y = bi[1];
        zc.x = br[2]; zc.y = bi[2];
        w.x = s1;
        w.y = -c1;
        pair_power(zb, zc, w, pb, pc);
        prow[96] = pb;
        prow[160] = pc;
    }
}

// Filterbank dot product over the filter's non-zero band.  Host (emulator / oracle order): float32 accumulate in bin
// order with separate multiply and add roundings.  Device: the plan aligns the band to bin quads (explicit zero
// weights), rows of the power tile are 16-byte aligned, so one step is two 16-byte shared loads and four FMAs;
// fusing the multiply-add moves the result by <= 1 ulp of the sum (the parity bar on log-mel is 1e-4).
FA_HD float mel_dot(const float *prow, const float *w, int lo, int hi) {
    float acc = 0.0f;
    const float *p = prow + lo;
    const int n = hi - lo;
    for (int b = 0; b < n; ++b) {
        const float t = w[b] * p[b];
        acc = acc + t;
    }
    return acc;
}
#if defined(__CUDACC__)
// p4 / w4: first quad of the band in the power row / in the packed weights; nq quads.  Not unrolled: bands are 1..6
// quads long and the unrolled remainder ladder cost ~100 instructions per dot product (profiles/r01c_mel.md).
__device__ __forceinline__ float mel_dot_quads(const float4 *p4, const float4 *w4, int nq) {
    float acc = 0.0f;
#pragma unroll 1
    for (int b = 0; b < nq; ++b) {
        const float4 x = p4[b], c = w4[b];
        acc = fmaf(c.x, x.x, acc);
        acc = fmaf(c.y, x.y, acc);
        acc = fmaf(c.z, x.z, acc);
        acc = fmaf(c.w, x.w, acc);
    }
    return acc;
}
#endif

FA_HD float log_value(float v, float floor_, int clamped) {
    return clamped ? logf(v > floor_ ? v : floor_) : logf(v + floor_);
}

// Pre-emphasis y[i] = x[i] - a x[i-1].  i == 0 uses Swift scalar arithmetic (two roundings,
// AudioMelSpectrogram.swift:373); i > 0 is vDSP_vsma = fused multiply-add (:381-387).
FA_HD float preemph_first(float x0, float last, float a) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(x0, __fmul_rn(a, last));
#else
    const float t = a * last;
    return x0 - t;
#endif
}
FA_HD float preemph_rest(float x, float xprev, float a) { return fmaf(xprev, -a, x); }

} // namespace mel
} // namespace fa
