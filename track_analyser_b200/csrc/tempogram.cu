// K4b: windowed autocorrelation tempogram with the packed shared-memory FFT core (fft2_core.cuh).
//
// Replaces librosa.feature.tempogram(win_length=384, center=True, window="hann", norm=inf) as
// called from report.py:260.  For output frame t: the onset envelope padded by win/2 on both
// sides with a linear ramp to 0, 384 samples starting at t, times a periodic Hann window,
// autocorrelated (librosa pads to 768; any length >= 2*win-1 gives the same linear
// autocorrelation, here 1024), first `win` lags, divided by the frame's max |.|.
// FOUR frames share one packed transform pair: transform A = frame f + i*frame f+2, transform B =
// frame f+1 + i*frame f+3 (so the packed real parts are frames f, f+1 and the imaginary parts f+2,
// f+3) -> Z -> Hermitian split -> (|X_re|^2 - i |X_im|^2) -> second forward FFT = (N*ac_re, -N*ac_im)
// because both power spectra are real and even.  After pass 3 a thread holds exactly the bins
// k = r + 64*n1 that pass 1 of the second transform consumes, so the power spectrum never goes
// through shared memory; the second transform ends in registers.
// Output (win, T) row-major float32, rows written 32 frames (128 B) at a time.
#include <algorithm>

#include "common.cuh"
#include "fft2_core.cuh"

namespace ta {

static constexpr int TG_N = 1024;
static constexpr int TG_TF = 32;       // frames per tile
static constexpr int TG_TFP = TG_TF + 2;  // tile row pitch, TFP/2 odd: conflict-free 64-bit column-pair stores

struct TgParams {
    const TrackDesc* tracks;
    int n_tracks;
    int total_tiles;
    int win;
    const float2* tw1;
    const float2* tw2;
    const float* window;  // [win]
    const float* env;     // packed per-frame series
    float* out;           // [win * P]
};

__device__ __forceinline__ void tg_barrier(int g, int n) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(n) : "memory"); }

__device__ __forceinline__ float padded_env(const float* __restrict__ x, int T, int half, int j) {
    // j indexes np.pad(x, (half, half), mode="linear_ramp", end_values=0)
    const int i = j - half;
    if (i < 0) return x[0] * (float(j) / float(half));
    if (i >= T) {
        const int d = i - T;  // 0 .. half-1
        return (d < half) ? x[T - 1] * (float(half - 1 - d) / float(half)) : 0.f;
    }
    return x[i];
}

// WIN > 0: the window length is a compile-time constant, so the transform is pruned by the compiler: pass 1 of the
// first transform sees the rows n1 >= WIN/64 as literal zeros and pass 3 of the second one only computes the lags
// below WIN.  WIN == 0: window length from the parameters (any even length <= 512).
template <int WIN>
__global__ void __launch_bounds__(512, 1) tempogram_kernel(const TgParams p) {
    using namespace p2;
    using C = FftCfg<TG_N>;
    using E = Ex<TG_N>;
    constexpr int M = C::M, NG = 512 / M, TFP = TG_TFP, N = TG_N;
    static_assert(NG * 4 == TG_TF, "one round of slots (4 frames each) fills a tile");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);  // [win][TFP]
    const size_t tile_bytes = ((size_t(WIN > 0 ? WIN : p.win) * TFP * 4 + 15) / 16) * 16;
    float4* ex_all = reinterpret_cast<float4*>(smem_raw + tile_bytes);
    float2* tw1s = reinterpret_cast<float2*>(ex_all + size_t(NG) * E::SLOTS);
    float2* tw2s = tw1s + 15 * M;
    float* red = reinterpret_cast<float*>(tw2s + 16 * C::Q);  // [16 warps][4]

    const int win = WIN > 0 ? WIN : p.win;
    const int tid = threadIdx.x, g = tid / M, r = tid % M, lane = tid & 31, warp = tid >> 5;
    float4* ex = ex_all + size_t(g) * E::SLOTS;
    for (int i = tid; i < 15 * M; i += 512) tw1s[i] = p.tw1[i];
    for (int i = tid; i < 16 * C::Q; i += 512) tw2s[i] = p.tw2[i];
    float wreg[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const int n = n1 * M + r;
        wreg[n1] = (n < win) ? 0.5f * p.window[n] : 0.f;  // 1/2: Hermitian split scaling
    }
    __syncthreads();
    const int half = win / 2;

    for (int w = blockIdx.x; w < p.total_tiles; w += gridDim.x) {
        int lo = 0, hi = p.n_tracks - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.tracks[mid].pitch_off <= (long long)w * TG_TF) lo = mid; else hi = mid - 1;
        }
        const TrackDesc td = p.tracks[lo];
        const int T = td.n_frames;
        const int t0 = int((long long)w * TG_TF - td.pitch_off);  // frame pitches are multiples of 32: tiles = pitch / 32
        const int nf = min(TG_TF, T - t0);
        const float* __restrict__ x = p.env + td.pitch_off;
        const int f = 4 * g, t = t0 + f;  // this group's four frames
        if (f < nf) {                     // group-uniform
            C2 v[16];
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const int n = n1 * M + r;
                if ((WIN > 0 && WIN % M == 0) ? (n1 < WIN / M) : (n < win)) {
                    const float a0 = padded_env(x, T, half, t + n), a1 = padded_env(x, T, half, t + 1 + n);
                    const float a2 = padded_env(x, T, half, t + 2 + n), a3 = padded_env(x, T, half, t + 3 + n);
                    v[n1].re = pmuls(make_float2(a0, a1), wreg[n1]);
                    v[n1].im = pmuls(make_float2(a2, a3), wreg[n1]);
                } else {
                    v[n1].re = v[n1].im = make_float2(0.f, 0.f);
                }
            }
            pass1<N>(v, r, tw1s, ex);
            tg_barrier(g, M);
            pass2<N>(v, r, tw2s, ex);
            tg_barrier(g, M);
            pass3<N, 2>(v, r, ex);
            tg_barrier(g, M);
            // power spectra of the four frames in pass-1 register order of the second transform:
            // register b*Q + k3 holds bin k = r + M*(b + 4*k3), i.e. input row n1 = b + 4*k3
            C2 u[16];
#pragma unroll
            for (int b = 0; b < C::NB; ++b)
#pragma unroll
                for (int k3 = 0; k3 < C::Q; ++k3) {
                    const int n1 = b + 4 * k3, k = r + M * n1;
                    const C2 zk = v[b * C::Q + k3];
                    C2 zn = unpack(ex[E::slot_of((N - k) & (N - 1))]);
                    if (n1 == 0 && r == 0) zn = zk;
                    C2 xa, xb;
                    split_pair(zk, zn, xa, xb);
                    u[n1].re = pfma(xa.re, xa.re, pmul(xa.im, xa.im));
                    u[n1].im = pmuls(pfma(xb.re, xb.re, pmul(xb.im, xb.im)), -1.0f);
                }
            tg_barrier(g, M);  // mirror reads done before pass 1 overwrites the slots
            pass1<N>(u, r, tw1s, ex);
            tg_barrier(g, M);
            pass2<N>(u, r, tw2s, ex);
            tg_barrier(g, M);
            pass3<N, 0>(u, r, ex);
            // u[b*Q + k3] = (N*ac of frames f, f+1 ; -N*ac of frames f+2, f+3) at lag r + M*(b + 4*k3)
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
            for (int b = 0; b < C::NB; ++b)
#pragma unroll
                for (int k3 = 0; k3 < C::Q; ++k3) {
                    const int lag = r + M * (b + 4 * k3);
                    if ((WIN > 0 && WIN % M == 0) ? (b + 4 * k3 < WIN / M) : (lag < win)) {
                        const C2 z = u[b * C::Q + k3];
                        m0 = fmaxf(m0, fabsf(z.re.x));
                        m1 = fmaxf(m1, fabsf(z.re.y));
                        m2 = fmaxf(m2, fabsf(z.im.x));
                        m3 = fmaxf(m3, fabsf(z.im.y));
                    }
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
                m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
                m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
                m3 = fmaxf(m3, __shfl_xor_sync(0xffffffffu, m3, o));
            }
            if (lane == 0) *reinterpret_cast<float4*>(red + warp * 4) = make_float4(m0, m1, m2, m3);
            tg_barrier(g, M);
            {
                const float4 ra = *reinterpret_cast<const float4*>(red + (2 * g) * 4);
                const float4 rb = *reinterpret_cast<const float4*>(red + (2 * g + 1) * 4);
                m0 = fmaxf(ra.x, rb.x); m1 = fmaxf(ra.y, rb.y); m2 = fmaxf(ra.z, rb.z); m3 = fmaxf(ra.w, rb.w);
            }
            const float s0 = (m0 > 0.f) ? 1.0f / m0 : 1.0f, s1 = (m1 > 0.f) ? 1.0f / m1 : 1.0f;
            const float s2 = (m2 > 0.f) ? -1.0f / m2 : -1.0f, s3 = (m3 > 0.f) ? -1.0f / m3 : -1.0f;
#pragma unroll
            for (int b = 0; b < C::NB; ++b)
#pragma unroll
                for (int k3 = 0; k3 < C::Q; ++k3) {
                    const int lag = r + M * (b + 4 * k3);
                    if ((WIN > 0 && WIN % M == 0) ? (b + 4 * k3 < WIN / M) : (lag < win)) {
                        const C2 z = u[b * C::Q + k3];
                        *reinterpret_cast<float2*>(tile + lag * TFP + f) = make_float2(z.re.x * s0, z.re.y * s1);
                        *reinterpret_cast<float2*>(tile + lag * TFP + f + 2) = make_float2(z.im.x * s2, z.im.y * s3);
                    }
                }
        }
        __syncthreads();
        {   // rows of 32 frames -> global, 64-bit accesses, 4 rows per warp instruction? no: 16 lanes per row
            const int hw = tid >> 4, fp = tid & 15;
            const bool ok0 = 2 * fp < nf, ok1 = 2 * fp + 1 < nf;
            float* dst = p.out + size_t(td.pitch_off) * win + t0 + 2 * fp;
            if (ok0)
                for (int lag = hw; lag < win; lag += 32) {
                    const float2 a = *reinterpret_cast<const float2*>(tile + lag * TFP + 2 * fp);
                    if (ok1) *reinterpret_cast<float2*>(dst + size_t(lag) * td.ld) = a;
                    else dst[size_t(lag) * td.ld] = a.x;
                }
        }
        __syncthreads();
    }
}

int run_tempogram(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* env, float* out,
                  cudaStream_t stream) {
    using C = FftCfg<TG_N>;
    const int win = plan->desc.tempogram_win;
    TA_REQUIRE(win >= 2 && win <= 512 && win % 2 == 0, "tempogram window must be even and <= 512 frames");
    TA_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "tempogram output must be 16-byte aligned");
    TgParams p{};
    p.tracks = d_tracks;
    p.n_tracks = hb.n_tracks;
    p.total_tiles = int(hb.total_pitch / TG_TF);
    p.win = win;
    p.tw1 = plan->d_tg_tw1;
    p.tw2 = plan->d_tg_tw2;
    p.window = plan->d_tg_window;
    p.env = env;
    p.out = out;
    const size_t smem = ((size_t(win) * TG_TFP * 4 + 15) / 16) * 16 + size_t(8) * p2::Ex<TG_N>::SLOTS * 16 + size_t(15) * C::M * 8 +
                        size_t(16) * C::Q * 8 + 16 * 4 * 4;
    const int grid = std::max(1, std::min(plan->sm_count, p.total_tiles));
    if (win == 384) {  // librosa's default: pruned instantiation
        TA_CUDA(cudaFuncSetAttribute(tempogram_kernel<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tempogram_kernel<384><<<grid, 512, smem, stream>>>(p);
    } else {
        TA_CUDA(cudaFuncSetAttribute(tempogram_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tempogram_kernel<0><<<grid, 512, smem, stream>>>(p);
    }
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
