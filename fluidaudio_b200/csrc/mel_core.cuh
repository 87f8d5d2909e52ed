// Per-lane math of the fused log-mel kernel (nFFT = 512), written so that the SAME source compiles for the
// device (mel_kernels.cu) and for the host lane-emulator used by the CPU test-suite (tests/emul/mel_emul.cpp):
// every function takes the lane id explicitly, touches "shared memory" only through the pointers it is given,
// and keeps no state between phases other than the read-only LaneTables.
//
// Reference being re-implemented: Sources/FluidAudio/Shared/AudioMelSpectrogram.swift:404-453 (frame loop),
// :459-481 (512-point DFT of a real frame, power of bins 0..256), :432-452 (filterbank mat-vec + log).
//
// Algorithm (one warp per frame):
//   real frame x[0..512)  ->  z[m] = x[2m] + i x[2m+1], m < 256           (even/odd packing)
//   Z = FFT256(z): radix 8 x 8 x 4 decimation in frequency, in shared memory, __syncwarp between passes
//   X[b] = E[b] + W512^b O[b],  E = (Z[b] + conj Z[256-b]) / 2,  O = (Z[b] - conj Z[256-b]) / (2i)
//   power[b] = |X[b]|^2, b = 0..256
//
// Shared-memory layout of the 256-point complex buffer (separate re/im arrays): element idx lives at
// idx + 4*(idx >> 5)  (288 floats per array).  With it every pass is bank-conflict free:
//   pass 1 stores  l + 32q            -> bank (l + 4q)      mod 32, distinct over lanes l
//   pass 2 ld/st   32q + h + 4r       -> bank (4q+h + 4r)   mod 32, lane = 4q + h
//   pass 3 loads   4C .. 4C+3 as one 128-bit access; stores natural order k = q + 8 q2 + 64 k2
//                                      -> bank q + 4(q2>>2) + 8(q2&3) + 8 k2, distinct over C = 8q + q2
//   post   loads   b = l + 32 j       -> bank (l + 4j)      mod 32
#pragma once

#include "fa_common.cuh"
#include <math.h>
#include <stdint.h>

namespace fa {
namespace mel {

constexpr int kNfft = 512;
constexpr int kHalf = 256;
constexpr int kBins = 257;
constexpr int kFftPad = 288;      // floats per re / im array
constexpr int kTileFrames = 32;   // frames per CTA tile == lanes of the mel stage
constexpr int kPowStride = 257;   // odd: lane-per-frame reads of a fixed bin hit 32 distinct banks

struct alignas(8) cpx {
    float x, y;
};
struct alignas(16) vec4 {
    float a, b, c, d;
};

FA_HD int pad_addr(int idx) { return idx + 4 * (idx >> 5); }

// Constants a lane needs for every frame it processes; loaded once per kernel.
struct LaneTables {
    float win[16];     // window coefficient at buffer positions j = 2(l+32r) [even slot 2r] and j+1 [odd slot 2r+1]
    uint32_t in_win;   // bit s set  <=>  slot s lies inside [off, off+win)
    float t1r[8], t1i[8];   // W256^(l q)
    float t2r[8], t2i[8];   // W32^((l&3) q2)
    float pwr[8], pwi[8];   // W512^(l + 32 j)
};

// win_tab[512]: window value per buffer position (0 outside the window), in_tab[512]: 1 inside the window.
// tw256[k] = (cos 2 pi k/256, -sin 2 pi k/256), tw512[k] = (cos 2 pi k/512, -sin 2 pi k/512), k < 256.
FA_HD void load_lane_tables(int l, const float *win_tab, const uint8_t *in_tab, const cpx *tw256, const cpx *tw512,
                            LaneTables &T) {
    T.in_win = 0;
    for (int r = 0; r < 8; ++r) {
        const int j = 2 * (l + 32 * r);
        T.win[2 * r] = win_tab[j];
        T.win[2 * r + 1] = win_tab[j + 1];
        if (in_tab[j]) T.in_win |= 1u << (2 * r);
        if (in_tab[j + 1]) T.in_win |= 1u << (2 * r + 1);
    }
    for (int q = 0; q < 8; ++q) {
        const cpx a = tw256[(l * q) & 255];
        T.t1r[q] = a.x;
        T.t1i[q] = a.y;
        const cpx b = tw256[(8 * (l & 3) * q) & 255];
        T.t2r[q] = b.x;
        T.t2i[q] = b.y;
        const cpx c = tw512[l + 32 * q];
        T.pwr[q] = c.x;
        T.pwi[q] = c.y;
    }
}

// forward 4-point DFT, in place, natural order
FA_HD void dft4(float &r0, float &i0, float &r1, float &i1, float &r2, float &i2, float &r3, float &i3) {
    const float ar = r0 + r2, ai = i0 + i2;
    const float br = r0 - r2, bi = i0 - i2;
    const float cr = r1 + r3, ci = i1 + i3;
    const float dr = r1 - r3, di = i1 - i3;
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
FA_HD void dft8(float (&re)[8], float (&im)[8]) {
    const float kS = 0.70710678118654752440f;
    float er0 = re[0] + re[4], ei0 = im[0] + im[4];
    float er1 = re[1] + re[5], ei1 = im[1] + im[5];
    float er2 = re[2] + re[6], ei2 = im[2] + im[6];
    float er3 = re[3] + re[7], ei3 = im[3] + im[7];
    float or0 = re[0] - re[4], oi0 = im[0] - im[4];
    const float xr1 = re[1] - re[5], xi1 = im[1] - im[5];
    const float xr2 = re[2] - re[6], xi2 = im[2] - im[6];
    const float xr3 = re[3] - re[7], xi3 = im[3] - im[7];
    float or1 = (xr1 + xi1) * kS, oi1 = (xi1 - xr1) * kS;     // * (1 - i)/sqrt2
    float or2 = xi2, oi2 = -xr2;                              // * (-i)
    float or3 = (xi3 - xr3) * kS, oi3 = -(xr3 + xi3) * kS;    // * (-1 - i)/sqrt2
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

// Pass 1.  pf -> pre-emphasised sample at buffer position j = 0 of this frame (8-byte aligned, hop even).
FA_HD void pass1(int l, const float *pf, const LaneTables &T, float *sre, float *sim) {
    float re[8], im[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int j = 2 * (l + 32 * r);
        const cpx v = *reinterpret_cast<const cpx *>(pf + j);   // one 64-bit shared load
        const float x0 = v.x, x1 = v.y;
        re[r] = (T.in_win >> (2 * r)) & 1u ? T.win[2 * r] * x0 : 0.0f;
        im[r] = (T.in_win >> (2 * r + 1)) & 1u ? T.win[2 * r + 1] * x1 : 0.0f;
    }
    dft8(re, im);
    sre[l] = re[0];
    sim[l] = im[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
        const float wr = T.t1r[q], wi = T.t1i[q];
        sre[l + 36 * q] = re[q] * wr - im[q] * wi;
        sim[l + 36 * q] = re[q] * wi + im[q] * wr;
    }
}

FA_HD void pass2(int l, const LaneTables &T, float *sre, float *sim) {
    const int base = 36 * (l >> 2) + (l & 3);
    float re[8], im[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        re[r] = sre[base + 4 * r];
        im[r] = sim[base + 4 * r];
    }
    dft8(re, im);
    sre[base] = re[0];
    sim[base] = im[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
        const float wr = T.t2r[q], wi = T.t2i[q];
        sre[base + 4 * q] = re[q] * wr - im[q] * wi;
        sim[base + 4 * q] = re[q] * wi + im[q] * wr;
    }
}

// Pass 3 is split: every lane must finish loading before any lane stores (the store layout differs).
FA_HD void pass3_load(int l, const float *sre, const float *sim, float (&re)[8], float (&im)[8]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int c = l + 32 * u;
        const int a = 4 * c + 4 * (c >> 3);
        const vec4 vr = *reinterpret_cast<const vec4 *>(sre + a);   // 128-bit shared loads
        const vec4 vi = *reinterpret_cast<const vec4 *>(sim + a);
        re[4 * u] = vr.a; re[4 * u + 1] = vr.b; re[4 * u + 2] = vr.c; re[4 * u + 3] = vr.d;
        im[4 * u] = vi.a; im[4 * u + 1] = vi.b; im[4 * u + 2] = vi.c; im[4 * u + 3] = vi.d;
    }
}

FA_HD void pass3_store(int l, float (&re)[8], float (&im)[8], float *sre, float *sim) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int c = l + 32 * u;
        dft4(re[4 * u], im[4 * u], re[4 * u + 1], im[4 * u + 1], re[4 * u + 2], im[4 * u + 2], re[4 * u + 3],
             im[4 * u + 3]);
        const int k0 = (c >> 3) + 8 * (c & 7);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            const int a = pad_addr(k0 + 64 * k2);
            sre[a] = re[4 * u + k2];
            sim[a] = im[4 * u + k2];
        }
    }
}

// Real-FFT recombination + power.  prow -> this frame's row of the power tile (257 floats).
FA_HD void post_power(int l, const float *sre, const float *sim, const LaneTables &T, float *prow) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int b = l + 32 * j;
        const int kb = pad_addr(b), kc = pad_addr((kHalf - b) & (kHalf - 1));
        const float ar = sre[kb], ai = sim[kb], cr = sre[kc], ci = sim[kc];
        const float er = 0.5f * (ar + cr), ei = 0.5f * (ai - ci);
        const float orr = 0.5f * (ai + ci), oi = 0.5f * (cr - ar);
        const float wr = T.pwr[j], wi = T.pwi[j];
        const float xr = er + (wr * orr - wi * oi);
        const float xi = ei + (wr * oi + wi * orr);
        prow[b] = xr * xr + xi * xi;
    }
    if (l == 0) {
        const float x = sre[0] - sim[0];   // X[256] = E[0] - O[0]
        prow[kHalf] = x * x;
    }
}

// float32 accumulate in bin order with separate multiply and add roundings (the oracle's mat-vec order).
FA_HD float mel_dot(const float *prow, const float *w, int lo, int hi) {
    float acc = 0.0f;
    for (int b = lo; b < hi; ++b) {
#if defined(__CUDA_ARCH__)
        acc = __fadd_rn(acc, __fmul_rn(w[b - lo], prow[b]));
#else
        const float t = w[b - lo] * prow[b];
        acc = acc + t;
#endif
    }
    return acc;
}

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
