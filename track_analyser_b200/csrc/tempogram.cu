// K4b: windowed autocorrelation tempogram with the shared-memory FFT core.
//
// Replaces librosa.feature.tempogram(win_length=384, center=True, window="hann", norm=inf) as
// called from report.py:260.  For output frame t: the onset envelope padded by win/2 on both
// sides with a linear ramp to 0, 384 samples starting at t, times a periodic Hann window,
// autocorrelated (librosa pads to 768; any length >= 2*win-1 gives the same linear
// autocorrelation, here 1024), first `win` lags, divided by the frame's max |.|.
// Two frames share one complex transform pair: z = a + i*b -> Z -> (|A|^2, |B|^2) by
// Hermitian split -> FFT(|A|^2 - i|B|^2) = (N*ac_a, -N*ac_b) because both power spectra
// are real and even.  The split lands exactly on the register layout pass 1 of the second
// transform needs (k = r + 64*i), so the power spectrum never goes through shared memory.
// Output (win, T) row-major float32, rows written 32 frames (128 B) at a time.
#include <algorithm>

#include "common.cuh"
#include "fft_core.cuh"

namespace ta {

static constexpr int TG_N = 1024;
static constexpr int TG_TF = 32;

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

__global__ void __launch_bounds__(512, 1) tempogram_kernel(const TgParams p) {
    using C = FftCfg<TG_N>;
    constexpr int M = C::M, NG = 512 / M, TFP = TG_TF + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);  // [win][TFP]
    const size_t tile_bytes = ((size_t(p.win) * TFP * 4 + 15) / 16) * 16;
    float2* ex_all = reinterpret_cast<float2*>(smem_raw + tile_bytes);
    float2* tw1s = ex_all + size_t(NG) * C::EX;
    float2* tw2s = tw1s + 15 * M;
    float* red = reinterpret_cast<float*>(tw2s + 16 * C::Q);  // [NG][2 warps][2]

    const int tid = threadIdx.x, g = tid / M, r = tid % M, lane = tid & 31, warp = tid >> 5;
    float2* ex = ex_all + size_t(g) * C::EX;
    for (int i = tid; i < 15 * M; i += 512) tw1s[i] = p.tw1[i];
    for (int i = tid; i < 16 * C::Q; i += 512) tw2s[i] = p.tw2[i];
    float wreg[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const int n = n1 * M + r;
        wreg[n1] = (n < p.win) ? 0.5f * p.window[n] : 0.f;  // 1/2: Hermitian split scaling
    }
    __syncthreads();
    const int half = p.win / 2;
    const int nlag_i = (p.win + M - 1) / M;

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
        const int slots = (nf + 1) / 2;
        for (int s = g; s < slots; s += NG) {
            const int f = 2 * s, t = t0 + f;
            float2 v[16];
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const int n = n1 * M + r;
                if (n < p.win) {
                    const float a = padded_env(x, T, half, t + n), b = padded_env(x, T, half, t + 1 + n);
                    v[n1] = make_float2(a * wreg[n1], b * wreg[n1]);
                } else {
                    v[n1] = make_float2(0.f, 0.f);
                }
            }
            pass1<TG_N>(v, r, tw1s, ex);
            tg_barrier(g, M);
            pass2_load<TG_N>(v, r, ex);
            tg_barrier(g, M);
            pass2_store<TG_N>(v, r, tw2s, ex);
            tg_barrier(g, M);
            pass3_load<TG_N>(v, r, ex);
            tg_barrier(g, M);
            pass3_store<TG_N>(v, r, ex);
            tg_barrier(g, M);
            // power spectra of both frames, directly in pass-1 register order (k = r + M*i)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = r + M * i;
                float2 xa, xb;
                split_pair(ex[k], ex[(TG_N - k) & (TG_N - 1)], xa, xb);
                v[i] = make_float2(fmaf(xa.x, xa.x, xa.y * xa.y), -fmaf(xb.x, xb.x, xb.y * xb.y));
            }
            tg_barrier(g, M);
            pass1<TG_N>(v, r, tw1s, ex);
            tg_barrier(g, M);
            pass2_load<TG_N>(v, r, ex);
            tg_barrier(g, M);
            pass2_store<TG_N>(v, r, tw2s, ex);
            tg_barrier(g, M);
            pass3_load<TG_N>(v, r, ex);
            tg_barrier(g, M);
            pass3_store<TG_N>(v, r, ex);
            tg_barrier(g, M);
            // ex[lag] = (N*ac_a, -N*ac_b); inf-norm over the first `win` lags
            float ma = 0.f, mb = 0.f;
            for (int i = 0; i < nlag_i; ++i) {
                const int lag = r + M * i;
                if (lag < p.win) {
                    ma = fmaxf(ma, fabsf(ex[lag].x));
                    mb = fmaxf(mb, fabsf(ex[lag].y));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
                mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
            }
            if (lane == 0) {
                red[warp * 2 + 0] = ma;
                red[warp * 2 + 1] = mb;
            }
            tg_barrier(g, M);
            const int w0 = g * (M / 32);
            ma = fmaxf(red[w0 * 2], red[(w0 + 1) * 2]);
            mb = fmaxf(red[w0 * 2 + 1], red[(w0 + 1) * 2 + 1]);
            const float sa = (ma > 0.f) ? 1.0f / ma : 1.0f, sb = (mb > 0.f) ? -1.0f / mb : -1.0f;
            for (int i = 0; i < nlag_i; ++i) {
                const int lag = r + M * i;
                if (lag < p.win) {
                    const float2 z = ex[lag];
                    tile[lag * TFP + f] = z.x * sa;
                    if (f + 1 < nf) tile[lag * TFP + f + 1] = z.y * sb;
                }
            }
            tg_barrier(g, M);
        }
        __syncthreads();
        float* dst = p.out + size_t(td.pitch_off) * p.win + t0 + lane;
        if (lane < nf)
            for (int lag = warp; lag < p.win; lag += 16) dst[size_t(lag) * td.ld] = tile[lag * TFP + lane];
        __syncthreads();
    }
}

int run_tempogram(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* env, float* out,
                  cudaStream_t stream) {
    using C = FftCfg<TG_N>;
    const int win = plan->desc.tempogram_win;
    TA_REQUIRE(win >= 2 && win <= 512 && win % 2 == 0, "tempogram window must be even and <= 512 frames");
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
    const size_t smem = ((size_t(win) * (TG_TF + 1) * 4 + 15) / 16) * 16 + size_t(8) * C::EX * 8 + size_t(15) * C::M * 8 +
                        size_t(16) * C::Q * 8 + 16 * 2 * 4;
    TA_CUDA(cudaFuncSetAttribute(tempogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::max(1, std::min(plan->sm_count, p.total_tiles));
    tempogram_kernel<<<grid, 512, smem, stream>>>(p);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
