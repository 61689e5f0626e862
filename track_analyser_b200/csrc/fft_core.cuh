// Shared-memory FFT core for the fused STFT / tempogram kernels (sm_100a).
//
// One complex N-point transform (N = 16*16*Q, Q in {4,8,16} -> N in
// {1024,2048,4096}) is executed by a *group* of M = N/16 threads, 16 complex
// points per thread, in three register-resident passes (radix 16, 16, Q) with
// two shared-memory exchanges.  Index algebra, forward sign exp(-2*pi*i*nk/N):
//
//   n = n1*M + n2*Q + n3          k = k1 + 16*k2 + 256*k3
//   pass 1 (thread r = n2*Q+n3):  A[k1]   = DFT16_{n1} x[n1*M + r],  * W_N^{r*k1}
//   pass 2 (thread n3*16 + k1):   B[k2]   = DFT16_{n2} A[k1; n2*Q+n3], * W_M^{n3*k2}
//   pass 3 (thread j = k1+16*k2): Z[j+256*k3] = DFT_Q_{n3} B[j; n3]
//
// The exchange layouts are chosen so that every 64-bit shared access of a
// half-warp touches 16 distinct bank pairs (row pitch P1 = M+1 is odd).
//
// Everything here is __host__ __device__ so tests/host_fft_harness.cu can run
// the exact pass code on the CPU (threads emulated by loops) and compare it
// with a float64 DFT; there is no GPU in the build container.
#pragma once
#include <cuda_runtime.h>

#define TA_HD __host__ __device__ __forceinline__

namespace ta {

template <int N>
struct FftCfg {
    static constexpr int M = N / 16;    // threads per transform, stride of pass-1 inputs
    static constexpr int Q = N / 256;   // last-pass radix
    static constexpr int P1 = M + 1;    // padded row pitch of exchange 1 (complex units)
    static constexpr int EX = (16 * P1 > N) ? 16 * P1 : N;  // complex slots per group buffer
    static constexpr int NB = 16 / Q;   // last-pass butterflies per thread
    static_assert(N == 1024 || N == 2048 || N == 4096, "N must be 16*16*{4,8,16}");
};

TA_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
TA_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
TA_HD float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by -i (forward) : (x,y) -> (y,-x)
TA_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

// In-place 4-point forward DFT on a0..a3 (natural order out).
TA_HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    float2 t2 = cadd(a1, a3), t3 = mul_mi(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

// In-place 16-point forward DFT, natural order in and out.
//   n = n0 + 4*n1, k = k1 + 4*k0:  X[k1+4k0] = sum_n0 W4^{n0 k0} W16^{n0 k1} sum_n1 W4^{n1 k1} x[n0+4n1]
TA_HD void dft16(float2 (&v)[16]) {
    const float C1 = 0.92387953251128673848f;  // cos(pi/8)
    const float S1 = 0.38268343236508978178f;  // sin(pi/8)
    const float R2 = 0.70710678118654752440f;  // sqrt(1/2)
#pragma unroll
    for (int n0 = 0; n0 < 4; ++n0) dft4(v[n0], v[n0 + 4], v[n0 + 8], v[n0 + 12]);
    // now v[n0 + 4*k1] = y[n0][k1]; apply W16^{n0*k1}
    // k1 = 1: exponents 1,2,3 ; k1 = 2: 2,4,6 ; k1 = 3: 3,6,9
    v[1 + 4] = cmul(v[1 + 4], make_float2(C1, -S1));
    v[2 + 4] = make_float2(R2 * (v[2 + 4].x + v[2 + 4].y), R2 * (v[2 + 4].y - v[2 + 4].x));
    v[3 + 4] = cmul(v[3 + 4], make_float2(S1, -C1));
    v[1 + 8] = make_float2(R2 * (v[1 + 8].x + v[1 + 8].y), R2 * (v[1 + 8].y - v[1 + 8].x));
    v[2 + 8] = mul_mi(v[2 + 8]);
    v[3 + 8] = make_float2(R2 * (v[3 + 8].y - v[3 + 8].x), -R2 * (v[3 + 8].x + v[3 + 8].y));
    v[1 + 12] = cmul(v[1 + 12], make_float2(S1, -C1));
    v[2 + 12] = make_float2(R2 * (v[2 + 12].y - v[2 + 12].x), -R2 * (v[2 + 12].x + v[2 + 12].y));
    v[3 + 12] = cmul(v[3 + 12], make_float2(-C1, S1));
    // second radix-4 level over n0 for each k1; results X[k1 + 4*k0] land in v[4*k1 + k0]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // transpose 4x4 so that v[k] = X[k]  (v[4*k1+k0] -> v[k1+4*k0])
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            float2 t = v[4 * a + b];
            v[4 * a + b] = v[4 * b + a];
            v[4 * b + a] = t;
        }
}

// In-place 8-point forward DFT on v[o..o+7], natural order.
TA_HD void dft8(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4, float2& a5, float2& a6,
                float2& a7) {
    const float R2 = 0.70710678118654752440f;
    // n = n0 + 2*n1 : even samples a0,a2,a4,a6 ; odd a1,a3,a5,a7
    dft4(a0, a2, a4, a6);  // y0[k1] in a0,a2,a4,a6
    dft4(a1, a3, a5, a7);  // y1[k1] in a1,a3,a5,a7
    float2 w1 = make_float2(R2 * (a3.x + a3.y), R2 * (a3.y - a3.x));    // * W8^1
    float2 w2 = mul_mi(a5);                                            // * W8^2
    float2 w3 = make_float2(R2 * (a7.y - a7.x), -R2 * (a7.x + a7.y));  // * W8^3
    float2 e0 = a0, e1 = a2, e2 = a4, e3 = a6, o0 = a1;
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
TA_HD void dftq(float2 (&v)[16]);
template <>
TA_HD void dftq<16>(float2 (&v)[16]) { dft16(v); }
template <>
TA_HD void dftq<8>(float2 (&v)[16]) {
    dft8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    dft8(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
}
template <>
TA_HD void dftq<4>(float2 (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(v[4 * b], v[4 * b + 1], v[4 * b + 2], v[4 * b + 3]);
}

// ---- the three passes -------------------------------------------------------
// tw1[(k1-1)*M + r] = W_N^{r*k1} (k1 = 1..15);  tw2[k2*Q + n3] = W_M^{n3*k2} (k2 = 0..15)

template <int N>
TA_HD void pass1(float2 (&v)[16], int r, const float2* __restrict__ tw1, float2* ex) {
    using C = FftCfg<N>;
    dft16(v);
    ex[r] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) ex[k1 * C::P1 + r] = cmul(v[k1], tw1[(k1 - 1) * C::M + r]);
}

template <int N>
TA_HD void pass2_load(float2 (&v)[16], int tid, const float2* ex) {
    using C = FftCfg<N>;
    const int k1 = tid & 15, n3 = tid >> 4;
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = ex[k1 * C::P1 + n2 * C::Q + n3];
}

template <int N>
TA_HD void pass2_store(float2 (&v)[16], int tid, const float2* __restrict__ tw2, float2* ex) {
    using C = FftCfg<N>;
    const int k1 = tid & 15, n3 = tid >> 4;
    dft16(v);
    ex[n3 * 256 + k1] = v[0];
#pragma unroll
    for (int k2 = 1; k2 < 16; ++k2)
        ex[n3 * 256 + k1 + 16 * k2] = cmul(v[k2], tw2[k2 * C::Q + n3]);
}

template <int N>
TA_HD void pass3_load(float2 (&v)[16], int tid, const float2* ex) {
    using C = FftCfg<N>;
#pragma unroll
    for (int b = 0; b < C::NB; ++b)
#pragma unroll
        for (int n3 = 0; n3 < C::Q; ++n3) v[b * C::Q + n3] = ex[n3 * 256 + tid + C::M * b];
}

// natural-order spectrum Z[k] into ex[k]
template <int N>
TA_HD void pass3_store(float2 (&v)[16], int tid, float2* ex) {
    using C = FftCfg<N>;
    dftq<C::Q>(v);
#pragma unroll
    for (int b = 0; b < C::NB; ++b)
#pragma unroll
        for (int k3 = 0; k3 < C::Q; ++k3) ex[tid + C::M * b + 256 * k3] = v[b * C::Q + k3];
}

// Split the spectrum Z of z = a + i*b (a, b real) into the two half spectra.
// With the inputs pre-scaled by 1/2:  Xa[k] = Z[k] + conj(Z[N-k]),  Xb[k] = -i (Z[k] - conj(Z[N-k])).
TA_HD void split_pair(float2 zk, float2 zn, float2& xa, float2& xb) {
    xa = make_float2(zk.x + zn.x, zk.y - zn.y);
    xb = make_float2(zk.y + zn.y, zn.x - zk.x);
}

}  // namespace ta
