// Packed two-transform FFT core for the fused STFT kernel (sm_100a).
//
// Same index algebra as fft_core.cuh (N = 16*16*Q, three register-resident passes, 16 complex
// points per thread, M = N/16 threads per group), with two changes measured to matter on B200
// (profiles/r1_notes.md):
//   * every thread carries TWO independent transforms in SoA register pairs
//     (re = {re_A, re_B}, im = {im_A, im_B}), so each butterfly is one FADD2/FMUL2/FFMA2
//     (add/mul/fma.rn.f32x2): half the issue slots per transform at the same FLOP rate; scalar
//     twiddles and window values enter as broadcast operands (ptxas folds {s, s} into `R.F32`),
//     and every shared-memory exchange is one 128-bit access per point for both transforms;
//   * the exchanges are IN PLACE: slot(d1, d2, d3) = d1*P1 + d2*Q + d3 holds, in turn,
//       after pass 1   A[k1; n2, n3]      (thread r = n2*Q + n3 wrote its 16 k1 values)
//       after pass 2   B[k1, k2; n3]      (thread k1 + 16*n3 read n2 = 0..15, wrote k2 = 0..15)
//       after pass 3   Z[k1 + 16 k2 + 256 k3]   (thread (k1, k2) read n3 < Q, keeps k3 in registers)
//     so a thread always overwrites exactly the slots it has just read and one barrier per pass
//     boundary suffices (4 per transform pair instead of 6 per transform).  P1 = 16*Q + 1 makes
//     every 128-bit access of a quarter-warp hit 8 distinct 16-byte bank groups.
// The Hermitian split needs Z[k] and Z[N-k]: each thread keeps its k3 < Q/2 values, stores the
// k3 >= Q/2 ones and reads the mirror slots of its kept bins (slot_of()).
//
// Everything is __host__ __device__ so tests/host_fft_harness.cu runs the same pass code on the CPU.
#pragma once
#include <cuda_runtime.h>

#include "fft_core.cuh"

namespace ta {
namespace p2 {

// ---- packed fp32x2 primitives ---------------------------------------------------------------
TA_HD float2 padd(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
TA_HD float2 psub(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
TA_HD float2 pmul(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
TA_HD float2 pfma(float2 a, float2 b, float2 c) {
#ifdef __CUDA_ARCH__
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7};"
        " fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
// scalar broadcast forms: a * s and a * s + c
TA_HD float2 pmuls(float2 a, float s) { return pmul(a, make_float2(s, s)); }
TA_HD float2 pfmas(float2 a, float s, float2 c) { return pfma(a, make_float2(s, s), c); }

// Two complex numbers (transform A in .x, transform B in .y).
struct C2 {
    float2 re, im;
};
TA_HD C2 cadd(const C2& a, const C2& b) { return {padd(a.re, b.re), padd(a.im, b.im)}; }
TA_HD C2 csub(const C2& a, const C2& b) { return {psub(a.re, b.re), psub(a.im, b.im)}; }
// times the scalar twiddle (wr, wi)
TA_HD C2 cmuls(const C2& a, float wr, float wi) {
    C2 o;
    o.re = pfmas(a.im, -wi, pmuls(a.re, wr));
    o.im = pfmas(a.im, wr, pmuls(a.re, wi));
    return o;
}
// times W8^1 = (1 - i)/sqrt2, -i, W8^3 = (-1 - i)/sqrt2
TA_HD C2 mul_w8_1(const C2& a) {
    const float R2 = 0.70710678118654752440f;
    return {pmuls(padd(a.re, a.im), R2), pmuls(psub(a.im, a.re), R2)};
}
TA_HD C2 mul_mi(const C2& a) { return {a.im, pmuls(a.re, -1.0f)}; }
TA_HD C2 mul_w8_3(const C2& a) {
    const float R2 = 0.70710678118654752440f;
    return {pmuls(psub(a.im, a.re), R2), pmuls(padd(a.re, a.im), -R2)};
}

// In-place 4-point forward DFT, natural order out.  The -i rotation of (a1 - a3) is folded into the
// last four add/subs.
TA_HD void dft4(C2& a0, C2& a1, C2& a2, C2& a3) {
    const C2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1.re = padd(t1.re, d.im);
    a1.im = psub(t1.im, d.re);
    a3.re = psub(t1.re, d.im);
    a3.im = padd(t1.im, d.re);
}

// In-place 16-point forward DFT, natural order in and out (same factorisation as ta::dft16).
TA_HD void dft16(C2 (&v)[16]) {
    const float C1 = 0.92387953251128673848f;  // cos(pi/8)
    const float S1 = 0.38268343236508978178f;  // sin(pi/8)
#pragma unroll
    for (int n0 = 0; n0 < 4; ++n0) dft4(v[n0], v[n0 + 4], v[n0 + 8], v[n0 + 12]);
    // v[n0 + 4*k1] = y[n0][k1]; apply W16^{n0*k1}
    v[1 + 4] = cmuls(v[1 + 4], C1, -S1);
    v[2 + 4] = mul_w8_1(v[2 + 4]);
    v[3 + 4] = cmuls(v[3 + 4], S1, -C1);
    v[1 + 8] = mul_w8_1(v[1 + 8]);
    v[2 + 8] = mul_mi(v[2 + 8]);
    v[3 + 8] = mul_w8_3(v[3 + 8]);
    v[1 + 12] = cmuls(v[1 + 12], S1, -C1);
    v[2 + 12] = mul_w8_3(v[2 + 12]);
    v[3 + 12] = cmuls(v[3 + 12], -C1, S1);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // X[k1 + 4*k0] sits in v[4*k1 + k0]: transpose (register renaming only)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            const C2 t = v[4 * a + b];
            v[4 * a + b] = v[4 * b + a];
            v[4 * b + a] = t;
        }
}

TA_HD void dft8(C2& a0, C2& a1, C2& a2, C2& a3, C2& a4, C2& a5, C2& a6, C2& a7) {
    dft4(a0, a2, a4, a6);
    dft4(a1, a3, a5, a7);
    const C2 w1 = mul_w8_1(a3), w2 = mul_mi(a5), w3 = mul_w8_3(a7);
    const C2 e0 = a0, e1 = a2, e2 = a4, e3 = a6, o0 = a1;
    a0 = cadd(e0, o0);
    a4 = csub(e0, o0);
    a1 = cadd(e1, w1);
    a5 = csub(e1, w1);
    a2 = cadd(e2, w2);
    a6 = csub(e2, w2);
    a3 = cadd(e3, w3);
    a7 = csub(e3, w3);
}

template <int Q>
TA_HD void dftq(C2 (&v)[16]);
template <>
TA_HD void dftq<16>(C2 (&v)[16]) { dft16(v); }
template <>
TA_HD void dftq<8>(C2 (&v)[16]) {
    dft8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    dft8(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
}
template <>
TA_HD void dftq<4>(C2 (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(v[4 * b], v[4 * b + 1], v[4 * b + 2], v[4 * b + 3]);
}

// ---- exchange slots: float4 {re_A, re_B, im_A, im_B} ---------------------------------------------
TA_HD float4 pack(const C2& c) { return make_float4(c.re.x, c.re.y, c.im.x, c.im.y); }
TA_HD C2 unpack(const float4& f) { return {make_float2(f.x, f.y), make_float2(f.z, f.w)}; }

template <int N>
struct Ex {
    using C = FftCfg<N>;
    static constexpr int SLOTS = 16 * C::P1;  // float4 slots per group
    TA_HD static int slot(int d1, int d2, int d3) { return d1 * C::P1 + d2 * C::Q + d3; }
    // slot that holds spectrum bin k after pass 3
    TA_HD static int slot_of(int k) { return slot(k & 15, (k >> 4) & 15, k >> 8); }
};

// tw1[(k1-1)*M + r] = W_N^{r*k1} (k1 = 1..15);  tw2[k2*Q + n3] = W_M^{n3*k2} (k2 = 0..15)
template <int N>
TA_HD void pass1(C2 (&v)[16], int r, const float2* __restrict__ tw1, float4* ex) {
    using C = FftCfg<N>;
    dft16(v);
    ex[r] = pack(v[0]);
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {
        const float2 w = tw1[(k1 - 1) * C::M + r];
        ex[k1 * C::P1 + r] = pack(cmuls(v[k1], w.x, w.y));
    }
}

// thread tid = k1 + 16*n3: DFT over n2, twiddle, back into the same 16 slots
template <int N>
TA_HD void pass2(C2 (&v)[16], int tid, const float2* __restrict__ tw2, float4* ex) {
    using C = FftCfg<N>;
    const int k1 = tid & 15, n3 = tid >> 4;
    float4* base = ex + k1 * C::P1 + n3;
    // loads in the order the first radix-4 level consumes them (n0, n0+4, n0+8, n0+12), so that the
    // butterflies of column n0 can issue while the later columns are still in flight
#pragma unroll
    for (int n0 = 0; n0 < 4; ++n0)
#pragma unroll
        for (int q = 0; q < 4; ++q) v[n0 + 4 * q] = unpack(base[(n0 + 4 * q) * C::Q]);
    dft16(v);
    base[0] = pack(v[0]);
#pragma unroll
    for (int k2 = 1; k2 < 16; ++k2) {
        const float2 w = tw2[k2 * C::Q + n3];
        base[k2 * C::Q] = pack(cmuls(v[k2], w.x, w.y));
    }
}

// thread tid: butterflies j = tid + M*b (k1 = j & 15, k2 = j >> 4); v[b*Q + k3] = Z[j + 256*k3].
// STORE = 1: the upper half (k3 >= Q/2) goes back to its slots for the mirror reads of the Hermitian split
// of a half spectrum; STORE = 2: every value goes back (full-spectrum mirror reads); STORE = 0: registers only.
template <int N, int STORE = 1>
TA_HD void pass3(C2 (&v)[16], int tid, float4* ex) {
    using C = FftCfg<N>;
#pragma unroll
    for (int b = 0; b < C::NB; ++b) {
        const int j = tid + C::M * b;
        const float4* src = ex + Ex<N>::slot(j & 15, j >> 4, 0);
#pragma unroll
        for (int n3 = 0; n3 < C::Q; ++n3) v[b * C::Q + n3] = unpack(src[n3]);
    }
    dftq<C::Q>(v);
    if (STORE == 0) return;
#pragma unroll
    for (int b = 0; b < C::NB; ++b) {
        const int j = tid + C::M * b;
        float4* dst = ex + Ex<N>::slot(j & 15, j >> 4, 0);
#pragma unroll
        for (int k3 = (STORE == 2 ? 0 : C::Q / 2); k3 < C::Q; ++k3) dst[k3] = pack(v[b * C::Q + k3]);
    }
}

// ---- paired pass 3 (Q <= 8): no exchange for the Hermitian split --------------------------------------
// A thread transforms NB/2 butterflies A_b = (k1, k2a) with k2a < 8 together with the butterflies B_b that hold
// the mirror bins: N - (k1 + 16 k2a + 256 k3) = k1m + 16 k2m + 256 (Q-1-k3) with k1m = (16 - k1) & 15 and
// k2m = (15 - k2a + [k1 == 0]) & 15.  (0, 0) and (0, 8) mirror themselves and are paired with each other
// (thread 0, pair 0).  Registers: v[b*Q + k3] = A_b, v[(NB/2 + b)*Q + k3] = B_b.
template <int N>
struct Pair3 {
    using C = FftCfg<N>;
    static_assert(C::NB >= 2, "pairing needs at least two butterflies per thread");
    static constexpr int HB = C::NB / 2;
    TA_HD static void ids(int tid, int b, int& k1, int& k2a, int& k1m, int& k2m) {
        k1 = tid & 15;
        k2a = (tid >> 4) + (C::M / 16) * b;
        if (k1 == 0 && k2a == 0) {
            k1m = 0;
            k2m = 8;
        } else {
            k1m = (16 - k1) & 15;
            k2m = (15 - k2a + (k1 == 0 ? 1 : 0)) & 15;
        }
    }
    // bins of pair b: side 0 = butterfly A, side 1 = butterfly B
    TA_HD static int bin(int tid, int b, int side, int k3) {
        int k1, k2a, k1m, k2m;
        ids(tid, b, k1, k2a, k1m, k2m);
        return (side ? k1m + 16 * k2m : k1 + 16 * k2a) + 256 * k3;
    }
    // i = 0..7 enumerates the thread's bins below N/2: i = b*Q + side*(Q/2) + k3
    TA_HD static int owned_bin(int tid, int i) {
        return bin(tid, i / C::Q, (i % C::Q) / (C::Q / 2), i % (C::Q / 2));
    }
};

template <int N>
TA_HD void pass3_paired(C2 (&v)[16], int tid, const float4* ex) {
    using C = FftCfg<N>;
    using P = Pair3<N>;
#pragma unroll
    for (int b = 0; b < P::HB; ++b) {
        int k1, k2a, k1m, k2m;
        P::ids(tid, b, k1, k2a, k1m, k2m);
        const float4* sa = ex + Ex<N>::slot(k1, k2a, 0);
        const float4* sb = ex + Ex<N>::slot(k1m, k2m, 0);
#pragma unroll
        for (int n3 = 0; n3 < C::Q; ++n3) {
            v[b * C::Q + n3] = unpack(sa[n3]);
            v[(P::HB + b) * C::Q + n3] = unpack(sb[n3]);
        }
    }
    dftq<C::Q>(v);
}

// Lower-half bins kept by thread tid after pass 3: i = b*(Q/2) + k3 (k3 < Q/2), 8 per thread.
template <int N>
TA_HD int kept_bin(int tid, int i) {
    using C = FftCfg<N>;
    const int b = i / (C::Q / 2), k3 = i % (C::Q / 2);
    return tid + C::M * b + 256 * k3;
}
template <int N>
TA_HD int kept_reg(int i) {
    using C = FftCfg<N>;
    return (i / (C::Q / 2)) * C::Q + i % (C::Q / 2);
}

// Hermitian split of Z (spectrum of a + i*b, inputs pre-scaled by 1/2): Xa = Z[k] + conj(Z[N-k]),
// Xb = -i (Z[k] - conj(Z[N-k])); packed over the two transforms.
TA_HD void split_pair(const C2& zk, const C2& zn, C2& xa, C2& xb) {
    xa.re = padd(zk.re, zn.re);
    xa.im = psub(zk.im, zn.im);
    xb.re = padd(zk.im, zn.im);
    xb.im = psub(zn.re, zk.re);
}

}  // namespace p2
}  // namespace ta
