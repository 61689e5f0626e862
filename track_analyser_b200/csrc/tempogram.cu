// K4b: windowed autocorrelation tempogram.
//
// Replaces librosa.feature.tempogram(win_length=384, center=True, window="hann", norm=inf) as
// called from report.py:260.  For output frame t: the onset envelope padded by win/2 on both
// sides with a linear ramp to 0, 384 samples starting at t, times a periodic Hann window,
// autocorrelated, first `win` lags, divided by the frame's max |.|.
// Two implementations:
//  * tempogram_sliding_kernel (default): O(1) float64 work per (frame, lag) from the closed form of the product
//    of two shifted Hann windows -- see the comment above it;
//  * tempogram_kernel (TA_TEMPOGRAM=fft): two 1024-point transforms per frame with the packed shared-memory FFT
//    core (fft2_core.cuh), the first version, kept as a cross-check.  FOUR frames share one packed transform pair:
//    transform A = frame f + i*frame f+2, transform B = frame f+1 + i*frame f+3 -> Z -> Hermitian split ->
//    (|X_re|^2 - i |X_im|^2) -> second forward FFT = (N*ac_re, -N*ac_im) because both power spectra are real and
//    even.  After pass 3 a thread holds exactly the bins k = r + 64*n1 that pass 1 of the second transform
//    consumes, so the power spectrum never goes through shared memory; the second transform ends in registers.
// Output (win, T) row-major float32, rows written 32 frames (128 B) at a time.
#include <algorithm>
#include <cstdlib>
#include <string>
#include <type_traits>

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

// ---------------------------------------------------------------------------------------------------------------
// Sliding formulation (default).  Consecutive tempogram frames are the same envelope shifted by ONE sample, and the
// product of two shifted periodic Hann windows is a three-term trigonometric polynomial:
//     w[n] w[n+l] = c0(l) + Re(A1(l) e^{ian}) + Re(A2(l) e^{2ian}),   a = 2 pi / win,
//     c0 = 1/4 + cos(al)/8,  A1 = -(1 + e^{ial})/4,  A2 = e^{ial}/8.
// With q_l[s] = y[s] y[s+l] (y = the padded envelope) the autocorrelation of frame t at lag l is
//     ac_t[l] = c0 E0 + Re(A1 E1) + Re(A2 E2),   E_k(t) = sum_{n < win-l} e^{ikan} q_l[t+n],
// and E_k(t+1) = e^{-ika} (E_k(t) - q_l[t] + e^{-ikal} q_l[t+win-l]): O(1) work per (frame, lag) instead of two
// 1024-point transforms per frame -- ~30 float64 operations against ~130 float32 ones plus six shared-memory
// exchanges, and no transform at all.  One thread owns one lag and walks the frames of its chunk; every 32 frames the
// block forms the per-frame maxima over the lags (inf-norm), scales and writes 128-byte rows.
// Round-off: the running sums are float64; a bound on their accumulated error (16 eps * steps * max window L1 norm)
// is compared per 32-frame tile with the smallest frame maximum of the tile, and when it exceeds 1e-7 of it (a loud
// passage followed by a much quieter one, or by digital silence, where the reference is exactly 0) the whole block
// recomputes its sums from scratch at the tile start and redoes the tile.  That keeps every output within 1e-7 of
// the exact windowed autocorrelation after normalisation.
static constexpr int TS_TILE = 32;
#ifndef TS_MINB
#define TS_MINB 3
#endif

struct TsRot {  // e^{ia}, e^{2ia}: uniform, read straight from the constant bank
    double c1, s1, c2, s2;
};

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) tempogram_sliding_kernel(const TrackDesc* __restrict__ tracks,
                                                                          const float* __restrict__ env, float* __restrict__ out,
                                                                          int win, int chunk, const TsRot rot) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* ys = reinterpret_cast<double*>(smem_raw);              // [TS_TILE + win] padded envelope of the current tile
    float* tile = reinterpret_cast<float*>(ys + TS_TILE + win);    // [win][TS_TILE + 1]
    float* mraw = tile + size_t(win) * (TS_TILE + 1);              // [1] smallest frame maximum of the tile
    unsigned* nzw = reinterpret_cast<unsigned*>(mraw + 1);         // [(TS_TILE + win) / 32 + 1] bit j: ys[j] != 0
    unsigned* zflag = nzw + (TS_TILE + win) / 32 + 1;              // [1] bit j: frame j of the tile is exactly zero
    unsigned* errmax = zflag + 1;                                  // [1] float bits of the largest error bound in the block
    const TrackDesc td = tracks[blockIdx.y];
    const int T = td.n_frames;
    const int ts = blockIdx.x * chunk;
    if (ts >= T) return;
    const int nt = min(chunk, T - ts);
    const int half = win / 2;
    const float* __restrict__ x = env + td.pitch_off;
    const int l = threadIdx.x, lane = l & 31, warp = l >> 5, nwarps = blockDim.x >> 5;
    const bool act = l < win;
    const int L = win - l;
    double sl, cl, s2l, c2l;
    sincospi(2.0 * l / win, &sl, &cl);
    sincospi(4.0 * l / win, &s2l, &c2l);
    const double c1 = rot.c1, s1 = rot.s1, c2 = rot.c2, s2 = rot.s2;
    const double c0 = 0.25 + 0.125 * cl, A1r = -0.25 * (1.0 + cl), A1i = -0.25 * sl, A2r = 0.125 * cl, A2i = 0.125 * sl;
    double e0 = 0, e1r = 0, e1i = 0, e2r = 0, e2i = 0, dabs = 0, dmax = 0;
    int since = 0;
    int neg_age = 0;  // tiles since the envelope last held a negative value (block-uniform); the frontend's never does
    auto init = [&](int j0) {  // sums of the window starting at ys[j0], Horner in e^{+ika}
        e0 = e1r = e1i = e2r = e2i = dabs = 0.0;
        if (act)
            for (int n = L - 1; n >= 0; --n) {
                const double q = ys[j0 + n] * ys[j0 + n + l];
                const double r1 = fma(c1, e1r, fma(-s1, e1i, q)), i1 = fma(c1, e1i, s1 * e1r);
                const double r2 = fma(c2, e2r, fma(-s2, e2i, q)), i2 = fma(c2, e2i, s2 * e2r);
                e1r = r1; e1i = i1; e2r = r2; e2i = i2;
                e0 += q;
                dabs += fabs(q);
            }
        dmax = dabs;
        since = 0;
    };
    auto emit = [&](int j) {
        const double ac = fma(c0, e0, fma(A1r, e1r, fma(-A1i, e1i, fma(A2r, e2r, -A2i * e2i))));
        tile[l * (TS_TILE + 1) + j] = float(ac);
    };
    // smallest lag-0 value among the frames of the tile that are not exactly zero (warp 0, after its rows are written)
    auto publish_min = [&](int nj) {
        if (warp == 0) {  // row 0 was written by lane 0 of this warp
            __syncwarp();
            float m = (lane < nj && !((zflag[0] >> lane) & 1u)) ? fabsf(tile[lane]) : 3.0e38f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) mraw[0] = m;
        }
    };
    // 32 frames: emit ac, slide the sums.  |ac_t[l]| <= ac_t[0] (Cauchy-Schwarz on the windowed frame), so the
    // inf-norm over the lags is the lag-0 value, row 0 of the tile.  ABS: keep the L1 norm of the window terms
    // separately (needed for the error bound only when the envelope has negative values; otherwise it equals e0).
    auto run_tile = [&](int nj, auto abs_tag) {
        constexpr bool ABS = decltype(abs_tag)::value;
        for (int j = 0; j < nj; ++j) {
            emit(j);
            const double qo = ys[j] * ys[j + l], qi = ys[j + L] * ys[j + win];
            const double u1r = fma(cl, qi, e1r - qo), u1i = fma(-sl, qi, e1i);    // + e^{-ial} q_in
            const double u2r = fma(c2l, qi, e2r - qo), u2i = fma(-s2l, qi, e2i);  // + e^{-2ial} q_in
            e1r = fma(c1, u1r, s1 * u1i);  // times e^{-ia}
            e1i = fma(c1, u1i, -s1 * u1r);
            e2r = fma(c2, u2r, s2 * u2i);
            e2i = fma(c2, u2i, -s2 * u2r);
            e0 += qi - qo;
            if (ABS) dabs += fabs(qi) - fabs(qo);
        }
        if (!ABS) dabs = e0;
        dmax = fmax(dmax, dabs);  // sampled per tile: the window (win frames) is much longer than a tile
        since += nj;
    };
    const int nwords = (TS_TILE + win + 31) / 32;
    for (int tb = 0; tb < nt; tb += TS_TILE) {
        const int nj = min(TS_TILE, nt - tb);
        int neg = 0;
        for (int j0 = warp * 32; j0 < nwords * 32; j0 += nwarps * 32) {  // whole warps, so the ballot is complete
            const int j = j0 + lane;
            const float v = (j < nj + win) ? padded_env(x, T, half, ts + tb + j) : 0.f;
            neg |= v < 0.f;
            if (j < TS_TILE + win) ys[j] = double(v);
            const unsigned word = __ballot_sync(0xffffffffu, v != 0.f);
            if (lane == 0) nzw[j0 >> 5] = word;
        }
        if (l == 0) errmax[0] = 0u;
        // also orders the previous tile's row stores (which read `tile`) before this tile's writes
        neg_age = __syncthreads_or(neg) ? 0 : neg_age + 1;
        const bool track_abs = neg_age <= win / TS_TILE + 1;
        if (warp == 0) {
            // frame j is exactly zero in the reference when every sample that meets a non-zero window weight is zero:
            // n = 1 .. win-1 (the periodic Hann window starts with w[0] = 0)
            int cnt = 0;
            const int a = lane + 1, b = lane + win;
            for (int k = 0; k < nwords; ++k) {
                const int lo = max(a - 32 * k, 0), hi = min(b - 32 * k, 32);
                if (lo < hi) {
                    const unsigned mask = (hi == 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                    cnt += __popc(nzw[k] & mask);
                }
            }
            const unsigned z = __ballot_sync(0xffffffffu, cnt == 0);
            if (lane == 0) zflag[0] = z;
        }
        if (tb == 0) init(0);
        // level 0: slide on; level 1: exact sums at the tile start, slide through the tile; level 2: exact sums per frame
        for (int level = 0;; ++level) {
            if (act) {
                if (level == 2) {
                    for (int j = 0; j < nj; ++j) { init(j); emit(j); }
                    init(nj);  // state for the next tile (ys holds nj + win samples)
                } else if (track_abs) {
                    run_tile(nj, std::true_type{});
                } else {
                    run_tile(nj, std::false_type{});
                }
            }
            if (level < 2) {
                float eb = act ? float(16.0 * 2.3e-16 * double(since + win) * dmax) : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) eb = fmaxf(eb, __shfl_xor_sync(0xffffffffu, eb, o));
                if (lane == 0) atomicMax(errmax, __float_as_uint(eb));
                publish_min(nj);
            }
            __syncthreads();  // tile, error bound and frame minimum complete
            if (level == 2 || !(__uint_as_float(errmax[0]) > 1e-7f * mraw[0])) break;
            __syncthreads();  // (rare) everybody has read the verdict
            if (l == 0) errmax[0] = 0u;
            if (level == 0) init(0);  // exact sums at the tile start; the tile is recomputed from them
            __syncthreads();
        }
        float* __restrict__ dst = out + size_t(td.pitch_off) * win + ts + tb;
        if (lane < nj) {
            const float mj = fabsf(tile[lane]);
            const float sc = ((zflag[0] >> lane) & 1u) ? 0.f : ((mj > 0.f) ? 1.0f / mj : 1.0f);
            const float* src = tile + warp * (TS_TILE + 1) + lane;
            float* d = dst + size_t(warp) * td.ld + lane;
            const size_t dstep = size_t(nwarps) * td.ld;
            const int sstep = nwarps * (TS_TILE + 1);
            int r = warp;
            for (; r + 3 * nwarps < win; r += 4 * nwarps, src += 4 * sstep, d += 4 * dstep) {
                const float v0 = src[0], v1 = src[sstep], v2 = src[2 * sstep], v3 = src[3 * sstep];
                d[0] = v0 * sc;
                d[dstep] = v1 * sc;
                d[2 * dstep] = v2 * sc;
                d[3 * dstep] = v3 * sc;
            }
            for (; r < win; r += nwarps, src += sstep, d += dstep) d[0] = src[0] * sc;
        }
    }
}

static int sliding_chunk(int max_frames, int n_tracks, int slots) {
    // frames per block: whole waves of resident blocks, >= 512 frames so the start-up sums (~75 frame-steps) stay small
    int best = (max_frames + 31) & ~31;
    double best_cost = 1e30;
    for (int nc = 1; nc <= std::max(1, max_frames / 512); ++nc) {
        const int chunk = (((max_frames + nc - 1) / nc) + 31) & ~31;
        const long long blocks = (long long)n_tracks * ((max_frames + chunk - 1) / chunk);
        const double cost = double((blocks + slots - 1) / slots) * (chunk + 75);
        if (cost < best_cost) { best_cost = cost; best = chunk; }
    }
    return best;
}

int run_tempogram(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* env, float* out,
                  cudaStream_t stream) {
    using C = FftCfg<TG_N>;
    const int win = plan->desc.tempogram_win;
    TA_REQUIRE(win >= 2 && win <= 512 && win % 2 == 0, "tempogram window must be even and <= 512 frames");
    TA_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "tempogram output must be 16-byte aligned");
    static const bool use_fft = [] { const char* e = getenv("TA_TEMPOGRAM"); return e && std::string(e) == "fft"; }();
    if (!use_fft) {
        TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
        const int threads = (win + 31) & ~31;
        const size_t smem = sizeof(double) * (TS_TILE + win) + sizeof(float) * (size_t(win) * (TS_TILE + 1) + 1) +
                            sizeof(unsigned) * ((TS_TILE + win) / 32 + 3);
        const double pi = 3.14159265358979323846;
        const TsRot rot{std::cos(2.0 * pi / win), std::sin(2.0 * pi / win), std::cos(4.0 * pi / win), std::sin(4.0 * pi / win)};
        auto* kern = threads <= 384 ? tempogram_sliding_kernel<384, TS_MINB> : tempogram_sliding_kernel<512, 2>;
        TA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        TA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        const int chunk = sliding_chunk(hb.max_frames, hb.n_tracks, std::max(1, per_sm) * plan->sm_count);
        dim3 grid((hb.max_frames + chunk - 1) / chunk, hb.n_tracks);
        kern<<<grid, threads, smem, stream>>>(d_tracks, env, out, win, chunk, rot);
        count_launch();
        TA_CUDA(cudaGetLastError());
        return TA_OK;
    }
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
