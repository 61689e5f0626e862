// Geometry of the shared-memory FFT used by the fused STFT / tempogram kernels (sm_100a).
//
// One complex N-point transform (N = 16*16*Q, Q in {4,8,16} -> N in {1024,2048,4096}) is executed by a
// *group* of M = N/16 threads, 16 complex points per thread, in three register-resident passes
// (radix 16, 16, Q).  Index algebra, forward sign exp(-2*pi*i*nk/N):
//
//   n = n1*M + n2*Q + n3          k = k1 + 16*k2 + 256*k3
//   pass 1 (thread r = n2*Q+n3):  A[k1]   = DFT16_{n1} x[n1*M + r],  * W_N^{r*k1}
//   pass 2 (thread n3*16 + k1):   B[k2]   = DFT16_{n2} A[k1; n2*Q+n3], * W_M^{n3*k2}
//   pass 3 (thread j = k1+16*k2): Z[j+256*k3] = DFT_Q_{n3} B[j; n3]
//
// The passes themselves (packed two-transform form, in-place exchanges) are in fft2_core.cuh.
#pragma once
#include <cuda_runtime.h>

#define TA_HD __host__ __device__ __forceinline__

namespace ta {

template <int N>
struct FftCfg {
    static constexpr int M = N / 16;    // threads per transform, stride of pass-1 inputs
    static constexpr int Q = N / 256;   // last-pass radix
    static constexpr int P1 = M + 1;    // padded pitch of the k1 digit in the exchange buffer (slots)
    static constexpr int NB = 16 / Q;   // last-pass butterflies per thread
    static_assert(N == 1024 || N == 2048 || N == 4096, "N must be 16*16*{4,8,16}");
};

}  // namespace ta
