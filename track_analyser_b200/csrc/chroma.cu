// K2b: chroma_stft = tuning estimate + chroma filterbank + projection + per-frame inf-norm.
//
// Replaces librosa.feature.chroma_stft(y, sr) (norm=inf, tuning=None) as called from
// harmony.py:108,149.  Stages, all per track and all on the device (no host round trip
// for the data-dependent tuning):
//   pip_peaks   librosa.piptrack on the POWER spectrogram S = |X|^2 (chroma_stft passes S):
//               bins in [150 Hz, min(4000 Hz, sr/2)), local maxima of S*(S > 0.1*max_f S),
//               parabolic shift, interpolated magnitude; peaks are appended to a per-track
//               list (magnitude, 0.01-semitone histogram bin of the tuning residual).
//   tuning      estimate_tuning: median of the peak magnitudes by radix select on the float
//               bit patterns, 100-bin histogram of the residuals of the peaks >= median,
//               left edge of the arg-max bin.
//   chroma_fb   librosa.filters.chroma(tuning): Gaussian bumps over log-frequency, per-column
//               L2 normalisation, Gaussian octave weighting, roll by -3, float32.
//   project     raw[c,t] = sum_f fb[c,f] * S[f,t], then divide each frame by its max.
// The projection is HBM-bound (reads the magnitude once, 4*B*T bytes per track, 13 flop per
// element) on CUDA cores; a tcgen05 version cannot beat an HBM-bound kernel and TF32 would
// break the 1e-4 parity bar (DESIGN.md section "filterbank contraction").
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "fft2_core.cuh"

namespace ta {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

struct PeakList {
    float* mag;            // [cap_total]
    unsigned char* bin;    // [cap_total]
    unsigned int* count;   // [n_tracks]
    const size_t* offset;  // [n_tracks] start of each track's region
};

__device__ __forceinline__ int residual_bin(float pitch, float bpo) {
    // pitch_tuning: residual = mod(bpo*log2(f/27.5), 1), folded to [-0.5, 0.5); np.histogram over linspace(-0.5,0.5,101)
    float res = bpo * log2f(pitch / 27.5f);
    res = res - floorf(res);
    if (res >= 0.5f) res -= 1.0f;
    const double x = double(res);
    int i = int(floor((x + 0.5) * 100.0));
    i = max(0, min(99, i));
    while (i > 0 && x < -0.5 + i * 0.01) --i;
    while (i < 99 && x >= ((i + 1 == 100) ? 0.5 : -0.5 + (i + 1) * 0.01)) ++i;
    return i;
}

// POWER: piptrack on |X|^2 (chroma_stft hands estimate_tuning the power spectrogram); otherwise on |X| (estimate_tuning(y=...)
// as reached from librosa.cqt: _spectrogram(power=1)).  bpo: bins per octave of the tuning residual (12 or 36).
template <bool POWER>
__global__ void __launch_bounds__(256) pip_peaks_kernel(const TrackDesc* __restrict__ tracks, const float* __restrict__ mag,
                                                       const float* __restrict__ frame_max, PeakList pl, int n_bins, int kmin,
                                                       int kmax, float bin_hz, float bpo) {
    const TrackDesc td = tracks[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= td.n_frames) return;
    const float* __restrict__ col = mag + size_t(td.pitch_off) * n_bins + t;
    const float mx = frame_max[td.pitch_off + t];
    const float ref = 0.1f * (POWER ? mx * mx : mx);
    float* out_mag = pl.mag + pl.offset[blockIdx.y];
    unsigned char* out_bin = pl.bin + pl.offset[blockIdx.y];
    auto S = [&](int k) {
        const float m = col[size_t(k) * td.ld];
        return POWER ? m * m : m;
    };
    float sm = S(kmin - 1), s0 = S(kmin);
    auto visit = [&](int k, float sp) {
        const float xm = sm > ref ? sm : 0.f, x0 = s0 > ref ? s0 : 0.f, xp = sp > ref ? sp : 0.f;
        if (x0 > xm && x0 >= xp) {
            // parabolic interpolation (float64 from float32 inputs, result stored as float32)
            const double a = double(sp) + double(sm) - 2.0 * double(s0);
            const double b = (double(sp) - double(sm)) / 2.0;
            const float shift = (fabs(b) >= fabs(a)) ? 0.f : float(-b / a);
            const float avg = (sp - sm) / 2.0f;
            const float dskew = 0.5f * avg * shift;
            const float pitch = float((double(k) + double(shift)) * double(bin_hz));
            const float m = s0 + dskew;
            if (pitch > 0.f) {
                const unsigned slot = atomicAdd(&pl.count[blockIdx.y], 1u);
                out_mag[slot] = m;
                out_bin[slot] = (unsigned char)residual_bin(pitch, bpo);
            }
        }
        sm = s0;
        s0 = sp;
    };
    int k = kmin;
    for (; k + 8 <= kmax; k += 8) {  // eight row reads in flight per thread: the loop is latency-bound otherwise
        float nx[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) nx[u] = S(k + 1 + u);
#pragma unroll
        for (int u = 0; u < 8; ++u) visit(k + u, nx[u]);
    }
    for (; k < kmax; ++k) visit(k, S(k + 1));
}

__device__ __forceinline__ unsigned ordered_key(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest (0-based) of n floats by 4 x 8-bit radix select; all threads of the CTA call this.
__device__ float block_select(const float* __restrict__ v, unsigned n, unsigned k, unsigned* hist /* smem[256] */,
                              unsigned* bcast /* smem[2] */) {
    unsigned prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned key = ordered_key(v[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned acc = 0, d = 0;
            for (; d < 256; ++d) {
                if (acc + hist[d] > k) break;
                acc += hist[d];
            }
            bcast[0] = d;
            bcast[1] = k - acc;
        }
        __syncthreads();
        prefix |= bcast[0] << shift;
        mask |= 0xffu << shift;
        k = bcast[1];
        __syncthreads();
    }
    return key_to_float(prefix);
}

__global__ void __launch_bounds__(1024) tuning_kernel(PeakList pl, double* __restrict__ tuning, int* __restrict__ tuning_idx) {
    __shared__ unsigned hist[256];
    __shared__ unsigned bc[2];
    __shared__ unsigned counts[100];
    const int trk = blockIdx.x;
    const unsigned n = pl.count[trk];
    const float* v = pl.mag + pl.offset[trk];
    const unsigned char* b = pl.bin + pl.offset[trk];
    if (n == 0) {  // pitch_tuning on an empty set returns 0.0
        if (threadIdx.x == 0) {
            tuning[trk] = 0.0;
            if (tuning_idx) tuning_idx[trk] = 50;
        }
        return;
    }
    const float lo = block_select(v, n, (n - 1) / 2, hist, bc);
    const float hi = block_select(v, n, n / 2, hist, bc);
    const float thr = __fadd_rn(lo, hi) * 0.5f;  // np.median of float32: mean of the two middle values
    for (int i = threadIdx.x; i < 100; i += blockDim.x) counts[i] = 0;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
        if (v[i] >= thr) atomicAdd(&counts[b[i]], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 0;
        for (int i = 1; i < 100; ++i)
            if (counts[i] > counts[best]) best = i;
        tuning[trk] = -0.5 + best * 0.01;
        if (tuning_idx) tuning_idx[trk] = best;
    }
}

// librosa.filters.chroma for one track; thread per FFT bin (column), 12 chroma rows.
__global__ void __launch_bounds__(256) chroma_fb_kernel(const double* __restrict__ tuning, float* __restrict__ fb /* [trk][B][12] */,
                                                       int n_bins, int n_fft, double sr) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_bins) return;
    const int trk = blockIdx.y;
    const double a440 = 440.0 * exp2(tuning[trk] / 12.0);
    const double step = sr / double(n_fft);
    auto frq = [&](int j) {  // frqbins[j]
        if (j == 0) return 12.0 * log2((1.0 * step) / (a440 / 16.0)) - 18.0;
        return 12.0 * log2((double(j) * step) / (a440 / 16.0));
    };
    const double fk = frq(k);
    const double width = (k + 1 < n_fft) ? fmax(frq(k + 1) - fk, 1.0) : 1.0;
    double w[12], ss = 0.0;
#pragma unroll
    for (int c = 0; c < 12; ++c) {
        double D = fk - double(c) + 6.0 + 120.0;
        D = D - 12.0 * floor(D / 12.0) - 6.0;  // np.remainder(., 12) - 6
        const double z = 2.0 * D / width;
        w[c] = exp(-0.5 * z * z);
        ss += w[c] * w[c];
    }
    double len = sqrt(ss);
    if (len < 2.2250738585072014e-308) len = 1.0;
    const double oct = (fk / 12.0 - 5.0) / 2.0;
    const double octw = exp(-0.5 * oct * oct);
    float* dst = fb + (size_t(trk) * n_bins + k) * 12;
#pragma unroll
    for (int c = 0; c < 12; ++c) dst[c] = float((w[(c + 3) % 12] / len) * octw);  // roll(-3) then float32
}

// raw[c, t] = sum_f fb[c, f] * |X[f, t]|^2, then each frame divided by its max.  Two frames per thread (one 64-bit read of
// the magnitude row per bin, one packed FFMA2 per chroma bin with the filterbank weight as broadcast operand), eight rows
// in flight per thread, 256 threads per CTA sharing the track's whole filterbank in shared memory (12 x n_bins floats,
// 49 KB at n_fft 2048, staged once): ~64 registers.  (The four-frames-per-thread version needed 98 registers: 13 warps
// per SM, 71 % long-scoreboard stalls at 48 % of the DRAM peak.)
//
// ROLL: the same walk down the frequency axis also carries librosa.feature.spectral_rolloff (features.py:116) the way
// numpy evaluates it -- `total = np.cumsum(S, axis=-2)` is a SEQUENTIAL float32 chain per frame, the threshold is
// float32(roll_percent) * total[-1], the answer the first bin whose running sum is not below it.  An integer decision,
// so the chain is walked in numpy's order on the very float32 magnitudes K1 wrote: one extra FADD per bin and frame here
// (hidden: the kernel is HBM-bound); see RollState below for how the answer is found without a second pass.
#ifndef CP_ROWS
#define CP_ROWS 8  // magnitude rows whose loads are in flight per thread
#endif

// Roll-off state of one thread's two frames.  K1 hands over sum_f |X| per frame, evaluated in another order (float32 chunk
// sums combined in double): within ~7e-5 of numpy's sequential float32 total in the worst case.  While the projection walks
// down the bins it keeps, per frame, the first bin whose running sum reaches (1 - ROLL_MARGIN) of the threshold that
// estimate predicts, and the running sum just before it; once the exact total is known the answer is found by walking on
// from there, typically zero to two bins.  If the exact threshold turns out below the margin (never observed; the bound
// above excludes it) the walk restarts at bin 0, so the result is numpy's in every case.
static constexpr float ROLL_MARGIN = 2e-4f;
struct RollState {
    float2 run, lo, before;   // running sums, early thresholds, running sum before the candidate bin
    int kx, ky;               // candidate bins (-1: none yet)
};
__device__ __forceinline__ void roll_init(RollState& r, const float* frame_sum, size_t col, bool ok, bool ok1, float roll_percent) {
    r.run = r.before = make_float2(0.f, 0.f);
    const float s0 = ok ? frame_sum[col] : 0.f, s1 = ok1 ? frame_sum[col + 1] : 0.f;
    r.lo = make_float2(roll_percent * s0 * (1.0f - ROLL_MARGIN), roll_percent * s1 * (1.0f - ROLL_MARGIN));
    r.kx = r.ky = -1;
}
__device__ __forceinline__ void roll_step(RollState& r, const float2 m, int k) {
    const float2 prev = r.run;
    r.run.x = __fadd_rn(r.run.x, m.x);
    r.run.y = __fadd_rn(r.run.y, m.y);
    if (r.kx < 0 && !(r.run.x < r.lo.x)) { r.kx = k; r.before.x = prev.x; }
    if (r.ky < 0 && !(r.run.y < r.lo.y)) { r.ky = k; r.before.y = prev.y; }
}
// first bin whose sequential float32 running sum is not below float32(roll_percent) * total, for frame column `colh`
__device__ __forceinline__ int roll_finish(const float* __restrict__ colh, int ld, int n_bins, float roll_percent, float total,
                                           float lo, int kc, float before) {
    const float thr = __fmul_rn(roll_percent, total);
    if (kc < 0 || thr < lo) { kc = 0; before = 0.f; }   // estimate too high: walk from the first bin
    float r = before;
    for (int k0 = kc; k0 < n_bins; k0 += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(colh + size_t(min(k0 + u, n_bins - 1)) * ld);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            r = __fadd_rn(r, v[u]);
            if (k0 + u >= n_bins - 1 || !(r < thr)) return min(k0 + u, n_bins - 1);
        }
    }
    return n_bins - 1;
}
static constexpr int CP_THREADS = 256;


template <bool CHROMA, bool ROLL>
__global__ void __launch_bounds__(CP_THREADS, 3) chroma_project_kernel(const TrackDesc* __restrict__ tracks,
                                                                       const float* __restrict__ mag, const float* __restrict__ fb,
                                                                       float* __restrict__ out, int32_t* __restrict__ rolloff_bin,
                                                                       const float* __restrict__ frame_sum, float roll_percent,
                                                                       int n_bins) {
    using namespace p2;
    extern __shared__ __align__(16) float wsm[];  // CHROMA: [n_bins * 12] filterbank
    const TrackDesc td = tracks[blockIdx.y];
    if (blockIdx.x * CP_THREADS * 2 >= td.n_frames) return;
    const int t = (blockIdx.x * CP_THREADS + threadIdx.x) * 2;
    const bool ok = t < td.n_frames;  // rows are padded to a multiple of 32 frames, so t, t+1 stay inside the row
    const float* __restrict__ col = mag + size_t(td.pitch_off) * n_bins + (ok ? t : 0);
    const float* __restrict__ w = fb + size_t(blockIdx.y) * n_bins * 12;
    float2 acc[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) acc[c] = make_float2(0.f, 0.f);
    RollState rs;
    if (ROLL) roll_init(rs, frame_sum, size_t(td.pitch_off) + t, ok, ok && t + 1 < td.n_frames, roll_percent);
    auto accumulate = [&](const float2 m, const float* wk, int k) {
        if (ROLL) roll_step(rs, m, k);
        if (CHROMA) {
            const float2 s = pmul(m, m);
            const float4 w0 = *reinterpret_cast<const float4*>(wk);
            const float4 w1 = *reinterpret_cast<const float4*>(wk + 4);
            const float4 w2 = *reinterpret_cast<const float4*>(wk + 8);
            const float ww[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
            for (int c = 0; c < 12; ++c) acc[c] = pfmas(s, ww[c], acc[c]);
        }
    };
    if (CHROMA) {
        for (int i = threadIdx.x; i < n_bins * 12; i += CP_THREADS) wsm[i] = w[i];
        __syncthreads();
    }
    int kk = 0;
    for (; kk + CP_ROWS <= n_bins; kk += CP_ROWS) {
        float2 m[CP_ROWS];
#pragma unroll
        for (int u = 0; u < CP_ROWS; ++u) m[u] = __ldg(reinterpret_cast<const float2*>(col + size_t(kk + u) * td.ld));
#pragma unroll
        for (int u = 0; u < CP_ROWS; ++u) accumulate(m[u], wsm + (kk + u) * 12, kk + u);
    }
    for (; kk < n_bins; ++kk) accumulate(__ldg(reinterpret_cast<const float2*>(col + size_t(kk) * td.ld)), wsm + kk * 12, kk);
    if (!ok) return;
    if (ROLL) {
        rolloff_bin[td.pitch_off + t] = roll_finish(col, td.ld, n_bins, roll_percent, rs.run.x, rs.lo.x, rs.kx, rs.before.x);
        if (t + 1 < td.n_frames)
            rolloff_bin[td.pitch_off + t + 1] = roll_finish(col + 1, td.ld, n_bins, roll_percent, rs.run.y, rs.lo.y, rs.ky, rs.before.y);
    }
    if (CHROMA) {
        float2 mx = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            mx.x = fmaxf(mx.x, fabsf(acc[c].x));
            mx.y = fmaxf(mx.y, fabsf(acc[c].y));
        }
        const float l0 = (mx.x < 1.1754943508222875e-38f) ? 1.0f : mx.x, l1 = (mx.y < 1.1754943508222875e-38f) ? 1.0f : mx.y;
        float* dst = out + size_t(td.pitch_off) * 12 + t;
#pragma unroll
        for (int c = 0; c < 12; ++c) *reinterpret_cast<float2*>(dst + size_t(c) * td.ld) = make_float2(acc[c].x / l0, acc[c].y / l1);
    }
}

#ifndef TA_TP_KT
#define TA_TP_KT 8
#endif
#ifndef TA_TP_STAGES
#define TA_TP_STAGES 3
#endif
static constexpr int TP_KT = TA_TP_KT;  // magnitude rows per pipeline stage
static constexpr int TP_BOX = 256;      // frames per TMA box (the hardware limit per box dimension); a stage is two boxes
static constexpr int TP_STAGES = TA_TP_STAGES;
static constexpr int TP_STAGE_FLOATS = 2 * TP_KT * TP_BOX;

__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1 << 24)) __trap();   // a transfer that never lands must abort the kernel, not hang the device
    }
}

// The same walk as chroma_project_kernel (kept below as the reference formulation and for configurations whose filterbank
// does not fit next to the pipeline), fed by the TMA unit: thread 0 issues, per stage, two cp.async.bulk.tensor loads of
// {256 frames, 8 bins} boxes of the track's (bins, T) matrix into shared memory and the CTA consumes them behind an mbarrier
// -- TP_STAGES * 16 KB in flight per CTA without a register or a scoreboard slot held, instead of eight 64-bit LDGs per
// thread (75 % long-scoreboard stalls at 57 % of the HBM peak).  Frames past the end of the track and bins past the last one
// arrive as zeros (the tensor map's extents are T and B).
template <bool CHROMA, bool ROLL>
__global__ void __launch_bounds__(CP_THREADS, 2) chroma_project_tma_kernel(const TrackDesc* __restrict__ tracks,
                                                                           const CUtensorMap* __restrict__ maps,
                                                                           const float* __restrict__ mag, const float* __restrict__ fb,
                                                                           float* __restrict__ out, int32_t* __restrict__ rolloff_bin,
                                                                           const float* __restrict__ frame_sum, float roll_percent,
                                                                           int n_bins) {
    using namespace p2;
    extern __shared__ __align__(128) float tsm[];   // [TP_STAGES][2][TP_KT][TP_BOX] tiles | filterbank | barriers
    float* wsm = tsm + TP_STAGES * TP_STAGE_FLOATS;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(wsm + (CHROMA ? ((n_bins * 12 + 1) & ~1) : 0));
    const TrackDesc td = tracks[blockIdx.y];
    const int f0 = blockIdx.x * CP_THREADS * 2;      // first frame of this CTA
    if (f0 >= td.n_frames) return;
    const int tid = threadIdx.x;
    const int t = f0 + 2 * tid;
    const bool ok = t < td.n_frames;
    const unsigned tile0 = static_cast<unsigned>(__cvta_generic_to_shared(tsm));
    const unsigned bar0 = static_cast<unsigned>(__cvta_generic_to_shared(bars));
    const unsigned long long map = reinterpret_cast<unsigned long long>(maps + blockIdx.y);
    const int n_groups = (n_bins + TP_KT - 1) / TP_KT;
    auto issue = [&](int grp) {   // thread 0: both boxes of row group `grp` into stage grp % TP_STAGES
        const int st = grp % TP_STAGES;
        const unsigned bar = bar0 + 8 * st, dst = tile0 + st * TP_STAGE_FLOATS * 4;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TP_STAGE_FLOATS * 4) : "memory");
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst + bx * TP_KT * TP_BOX * 4), "l"(map), "r"(f0 + bx * TP_BOX), "r"(grp * TP_KT), "r"(bar)
                         : "memory");
    };
    if (tid == 0) {
        for (int s = 0; s < TP_STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int gq = 0; gq < TP_STAGES && gq < n_groups; ++gq) issue(gq);
    }
    const float* __restrict__ w = fb + size_t(blockIdx.y) * n_bins * 12;
    if (CHROMA)
        for (int i = tid; i < n_bins * 12; i += CP_THREADS) wsm[i] = w[i];
    __syncthreads();   // barriers initialised and the filterbank staged
    float2 acc[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) acc[c] = make_float2(0.f, 0.f);
    RollState rs;
    if (ROLL) roll_init(rs, frame_sum, size_t(td.pitch_off) + t, ok, ok && t + 1 < td.n_frames, roll_percent);
    const int col = (2 * tid) % TP_BOX + ((2 * tid) / TP_BOX) * TP_KT * TP_BOX;   // this thread's frame pair inside a stage
    for (int grp = 0; grp < n_groups; ++grp) {
        const int st = grp % TP_STAGES;
        mbar_wait(bar0 + 8 * st, (grp / TP_STAGES) & 1);
        const float* tile = tsm + st * TP_STAGE_FLOATS + col;
        const int kk = grp * TP_KT;
        float2 m[TP_KT];
#pragma unroll
        for (int u = 0; u < TP_KT; ++u) m[u] = *reinterpret_cast<const float2*>(tile + u * TP_BOX);
#pragma unroll
        for (int u = 0; u < TP_KT; ++u) {
            if (kk + u >= n_bins) break;   // rows past the last bin are zero-filled; they have no filterbank row
            if (ROLL) roll_step(rs, m[u], kk + u);
            if (CHROMA) {
                const float* wk = wsm + (kk + u) * 12;
                const float2 s2 = pmul(m[u], m[u]);
                const float4 w0 = *reinterpret_cast<const float4*>(wk);
                const float4 w1 = *reinterpret_cast<const float4*>(wk + 4);
                const float4 w2 = *reinterpret_cast<const float4*>(wk + 8);
                const float ww[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
                for (int c = 0; c < 12; ++c) acc[c] = pfmas(s2, ww[c], acc[c]);
            }
        }
        __syncthreads();   // every thread has read this stage: it may be refilled
        if (tid == 0 && grp + TP_STAGES < n_groups) issue(grp + TP_STAGES);
    }
    if (!ok) return;
    const float* __restrict__ colg = mag + size_t(td.pitch_off) * n_bins + t;
    if (ROLL) {
        rolloff_bin[td.pitch_off + t] = roll_finish(colg, td.ld, n_bins, roll_percent, rs.run.x, rs.lo.x, rs.kx, rs.before.x);
        if (t + 1 < td.n_frames)
            rolloff_bin[td.pitch_off + t + 1] = roll_finish(colg + 1, td.ld, n_bins, roll_percent, rs.run.y, rs.lo.y, rs.ky, rs.before.y);
    }
    if (CHROMA) {
        float2 mx = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            mx.x = fmaxf(mx.x, fabsf(acc[c].x));
            mx.y = fmaxf(mx.y, fabsf(acc[c].y));
        }
        const float l0 = (mx.x < 1.1754943508222875e-38f) ? 1.0f : mx.x, l1 = (mx.y < 1.1754943508222875e-38f) ? 1.0f : mx.y;
        float* dst = out + size_t(td.pitch_off) * 12 + t;
#pragma unroll
        for (int c = 0; c < 12; ++c) *reinterpret_cast<float2*>(dst + size_t(c) * td.ld) = make_float2(acc[c].x / l0, acc[c].y / l1);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// A/B variant for the north star's "tensor-core filterbank projection, kept only if ncu shows it wins" clause
// (profiles/r2_row_g_tensor_core_evidence.md): the same contraction chroma[c, t] = sum_k fb[k, c] * |X[k, t]|^2 issued as
// tcgen05.mma kind::tf32 with the accumulator in tensor memory.  Selected by TA_PROJECT=umma for chroma-only requests; it
// is NOT the product path: one TF32 pass keeps 10 mantissa bits of every operand, which does not meet the rtol 1e-4 parity
// bar (a parity-grade version needs two more passes over residual tiles that CUDA cores would have to produce).
//   D[128 frames, 16] (TMEM, fp32)  +=  A[128 frames, 8 bins] (smem, MN-major, 128-byte swizzle)  x  B[8 bins, 16] (smem, K-major)
// A CTA owns 128 frames of one track.  Per stage, TMA brings four boxes of {32 frames, 64 bins} with the 128-byte swizzle
// the UMMA descriptor names (one box = eight 1 KB atoms of 8 bins x 32 frames); 128 threads square the stage in
// place (the projection runs on the power spectrum), and one thread issues eight MMAs and commits them to the stage's
// "empty" mbarrier, behind which the next TMA loads are issued.  Epilogue: tcgen05.ld, inf-norm, store.
static constexpr size_t SMEM_OPTIN_LIMIT = 232448;   // 227 KB per CTA on sm_100
static constexpr int UM_FRAMES = 128, UM_KCH = 64, UM_STAGES = 4, UM_N = 16;
static constexpr int UM_TILES = 8;   // 128-frame tiles per CTA
static constexpr int UM_CONSUMERS = 256, UM_THREADS = UM_CONSUMERS + 32;   // eight warps square the stages, one feeds the TMA unit
static constexpr int UM_STAGE_BYTES = UM_FRAMES * UM_KCH * 4;   // 32 KB

__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes, unsigned layout) {
    // cute::UMMA::SmemDescriptor: start >> 4 in [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46), version 1 in [46,48), layout in [61,64)
    return (unsigned long long)((smem_addr >> 4) & 0x3fff) | ((unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46) | ((unsigned long long)layout << 61);
}

__global__ void __launch_bounds__(UM_THREADS, 1) chroma_project_umma_kernel(const TrackDesc* __restrict__ tracks,
                                                                     const CUtensorMap* __restrict__ maps,
                                                                     const float* __restrict__ fb, float* __restrict__ out, int n_bins) {
    extern __shared__ __align__(1024) unsigned char usm[];
    const TrackDesc td = tracks[blockIdx.y];
    const int tile0 = blockIdx.x * UM_TILES, n_tiles_track = (td.n_frames + UM_FRAMES - 1) / UM_FRAMES;
    if (tile0 >= n_tiles_track) return;
    const int n_tiles = min(UM_TILES, n_tiles_track - tile0);   // this CTA's run of 128-frame tiles (one filterbank staging)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_chunks = (n_bins + UM_KCH - 1) / UM_KCH, k_pad = n_chunks * UM_KCH;
    const unsigned base = (static_cast<unsigned>(__cvta_generic_to_shared(usm)) + 1023u) & ~1023u;   // 1 KB atoms
    unsigned char* gen = usm + (base - static_cast<unsigned>(__cvta_generic_to_shared(usm)));
    float* bsm = reinterpret_cast<float*>(gen + UM_STAGES * UM_STAGE_BYTES);       // [k_pad / 4][2][8][4]
    const unsigned b_base = base + UM_STAGES * UM_STAGE_BYTES;
    const unsigned bar0 = b_base + k_pad * UM_N * 4;   // full[UM_STAGES], empty[UM_STAGES], done, then the TMEM address slot
    unsigned* slot = reinterpret_cast<unsigned*>(gen + UM_STAGES * UM_STAGE_BYTES + size_t(k_pad) * UM_N * 4 + (2 * UM_STAGES + 1) * 8);
    const unsigned long long map = reinterpret_cast<unsigned long long>(maps + blockIdx.y);
    auto full = [&](int s) { return bar0 + 8 * s; };
    auto empty = [&](int s) { return bar0 + 8 * (UM_STAGES + s); };
    const unsigned done = bar0 + 8 * 2 * UM_STAGES;
    if (tid == 0) {
        for (int s = 0; s < 2 * UM_STAGES + 1; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {   // 32 columns of tensor memory for the 128 x 16 fp32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(slot))) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // B: the track's filterbank, K-major core matrices of 8 (chroma) x 4 (bins) floats, two per 4-bin group (chroma 0-7, 8-15)
    const float* __restrict__ w = fb + size_t(blockIdx.y) * n_bins * 12;
    for (int i = tid; i < k_pad * UM_N; i += UM_THREADS) {
        const int kk = i / UM_N, n = i % UM_N;
        bsm[(kk >> 2) * 64 + (n >> 3) * 32 + (n & 7) * 4 + (kk & 3)] = (n < 12 && kk < n_bins) ? w[kk * 12 + n] : 0.f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *slot;
    const int total = n_tiles * n_chunks;   // stages are numbered through the CTA's whole run: g = tile * n_chunks + chunk
    if (warp == UM_CONSUMERS / 32) {   // producer warp: one thread keeps UM_STAGES loads in flight and never joins the consumers' barriers
        if (lane == 0)
            for (int g = 0; g < total; ++g) {
                const int st = g % UM_STAGES, f0 = (tile0 + g / n_chunks) * UM_FRAMES, chunk = g % n_chunks;
                if (g >= UM_STAGES) mbar_wait(empty(st), ((g / UM_STAGES) - 1) & 1);   // the MMAs that read the stage are done
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full(st)), "r"(UM_STAGE_BYTES) : "memory");
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(base + st * UM_STAGE_BYTES + q * (UM_STAGE_BYTES / 4)), "l"(map), "r"(f0 + 32 * q),
                                   "r"(chunk * UM_KCH), "r"(full(st)) : "memory");
            }
        return;
    }
    // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, A MN-major, B K-major, N = 16, M = 128
    constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((UM_N >> 3) << 17) | ((128u >> 4) << 24);
    for (int g = 0; g < total; ++g) {
        const int st = g % UM_STAGES, c = g % n_chunks, ti = g / n_chunks;
        mbar_wait(full(st), (g / UM_STAGES) & 1);
        float4* tile = reinterpret_cast<float4*>(gen + st * UM_STAGE_BYTES);
#pragma unroll 4
        for (int i = tid; i < UM_STAGE_BYTES / 16; i += UM_CONSUMERS) {   // |X| -> |X|^2 in place (element-wise: the swizzle does not matter)
            float4 v = tile[i];
            v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w;
            tile[i] = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(UM_CONSUMERS) : "memory");   // also: the previous tile's tcgen05.ld are behind this
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < UM_KCH / 8; ++j) {
                // MN-major 32-bit operands: 128-byte swizzle with 32-byte atoms (cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B = 1,
                // the layout CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes; with the 16-byte-atom swizzle the MMA returns zeros),
                // K atoms of 4 rows (512 B): an 8-bin MMA spans two of them (SBO), the 32-frame boxes are the M atoms (LBO)
                const unsigned long long ad = umma_desc(base + st * UM_STAGE_BYTES + j * 1024, UM_STAGE_BYTES / 4, 512, 1);
                const unsigned long long bd = umma_desc(b_base + (c * (UM_KCH / 8) + j) * 512, 256, 128, 0);
                const unsigned acc = (c | j) ? 1u : 0u;   // the first MMA of a tile overwrites the accumulator
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty(st)) : "memory");
            if (c == n_chunks - 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
        }
        if (c == n_chunks - 1 && warp < 4) {   // tile finished: warp w reads TMEM lanes 32 w .. 32 w + 31, one frame per thread
            mbar_wait(done, ti & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(tmem + ((unsigned(warp) * 32u) << 16)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int t = (tile0 + ti) * UM_FRAMES + warp * 32 + lane;
            if (t < td.n_frames) {
                float mx = 0.f;
#pragma unroll
                for (int q = 0; q < 12; ++q) mx = fmaxf(mx, fabsf(__uint_as_float(r[q])));
                const float l0 = (mx < 1.1754943508222875e-38f) ? 1.0f : mx;
                float* dst = out + size_t(td.pitch_off) * 12 + t;
#pragma unroll
                for (int q = 0; q < 12; ++q) dst[size_t(q) * td.ld] = __uint_as_float(r[q]) / l0;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(UM_CONSUMERS) : "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

// tensor maps for the variant above: boxes of {32 frames, 64 bins}, 128-byte swizzle
static int build_magnitude_maps_umma(const ta_plan* plan, const HostBatch& hb, const float* mag, void* d_maps, cudaStream_t stream) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return TA_ERR_UNSUPPORTED;
    std::vector<CUtensorMap> maps(hb.n_tracks);
    const int B = plan->n_bins;
    for (int i = 0; i < hb.n_tracks; ++i) {
        const TrackDesc& t = hb.tracks[i];
        const cuuint64_t dims[2] = {cuuint64_t(t.n_frames), cuuint64_t(B)};
        const cuuint64_t strides[1] = {cuuint64_t(t.ld) * sizeof(float)};
        const cuuint32_t box[2] = {32, UM_KCH};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(mag) + size_t(t.pitch_off) * B, dims,
                               strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return TA_ERR_UNSUPPORTED;
    }
    TA_CUDA(cudaMemcpyAsync(d_maps, maps.data(), sizeof(CUtensorMap) * hb.n_tracks, cudaMemcpyHostToDevice, stream));
    return TA_OK;
}

// workspace layout helpers ----------------------------------------------------------------
static void pip_range(const ta_plan* plan, int& kmin, int& kmax) {
    const double sr = plan->desc.sample_rate, fmax = std::min(4000.0, sr / 2.0), fmin = 150.0;
    kmin = plan->n_bins;
    kmax = 0;
    for (int k = 0; k < plan->n_bins; ++k)
        if (plan->h_freqs[k] >= fmin && plan->h_freqs[k] < fmax) {
            kmin = std::min(kmin, k);
            kmax = std::max(kmax, k + 1);
        }
    kmin = std::max(kmin, 1);
    kmax = std::min(kmax, plan->n_bins - 1);
}

static size_t peaks_per_frame(const ta_plan* plan) {
    int kmin, kmax;
    pip_range(plan, kmin, kmax);
    return size_t(std::max(1, (kmax - kmin + 1) / 2 + 1));
}

// scratch of one tuning estimate: peak magnitudes, residual bins, counts, offsets
size_t tuning_scratch_bytes(const ta_plan* plan, const HostBatch& hb) {
    const size_t per_frame = peaks_per_frame(plan);
    size_t peaks = 0;
    for (auto& t : hb.tracks) peaks += per_frame * size_t(t.n_frames);
    size_t bytes = 0;
    auto add = [&](size_t b) { bytes = (bytes + b + 255) / 256 * 256; };
    add(peaks * sizeof(float));
    add(peaks);
    add(sizeof(unsigned) * hb.n_tracks);
    add(sizeof(size_t) * hb.n_tracks);
    return bytes;
}

size_t chroma_scratch_bytes(const ta_plan* plan, const HostBatch& hb) {
    return tuning_scratch_bytes(plan, hb) + (sizeof(float) * size_t(hb.n_tracks) * plan->n_bins * 12 + 255) / 256 * 256;
}

// librosa.estimate_tuning on an existing magnitude spectrogram: piptrack on |X|^2 (power) or |X|, median gate,
// 0.01-bin histogram of the residuals at `bpo` bins per octave.  tuning_idx (optional): the arg-max histogram bin.
int run_tuning(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mag, const float* frame_max,
               bool power, int bpo, void* scratch, double* tuning, int* tuning_idx, cudaStream_t stream) {
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    int kmin, kmax;
    pip_range(plan, kmin, kmax);
    const size_t per_frame = peaks_per_frame(plan);
    std::vector<size_t> off(hb.n_tracks);
    size_t peaks = 0;
    for (int i = 0; i < hb.n_tracks; ++i) {
        off[i] = peaks;
        peaks += per_frame * size_t(hb.tracks[i].n_frames);
    }
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    size_t cur = 0;
    auto take = [&](size_t b) {
        unsigned char* r = base + cur;
        cur = (cur + b + 255) / 256 * 256;
        return r;
    };
    PeakList pl;
    pl.mag = reinterpret_cast<float*>(take(peaks * sizeof(float)));
    pl.bin = take(peaks);
    pl.count = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * hb.n_tracks));
    size_t* d_off = reinterpret_cast<size_t*>(take(sizeof(size_t) * hb.n_tracks));
    pl.offset = d_off;
    TA_CUDA(cudaMemcpyAsync(d_off, off.data(), sizeof(size_t) * hb.n_tracks, cudaMemcpyHostToDevice, stream));
    TA_CUDA(cudaMemsetAsync(pl.count, 0, sizeof(unsigned) * hb.n_tracks, stream));
    const float bin_hz = float(double(plan->desc.sample_rate) / double(plan->desc.n_fft));
    dim3 gf((hb.max_frames + 255) / 256, hb.n_tracks);
    if (kmax > kmin) {
        if (power) pip_peaks_kernel<true><<<gf, 256, 0, stream>>>(d_tracks, mag, frame_max, pl, plan->n_bins, kmin, kmax, bin_hz, float(bpo));
        else pip_peaks_kernel<false><<<gf, 256, 0, stream>>>(d_tracks, mag, frame_max, pl, plan->n_bins, kmin, kmax, bin_hz, float(bpo));
        count_launch();
        TA_CUDA(cudaGetLastError());
    }
    tuning_kernel<<<hb.n_tracks, 1024, 0, stream>>>(pl, tuning, tuning_idx);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

// One 2-d tensor map per track over its (bins, T) float32 magnitude matrix (row pitch ld), box = {256 frames, 8 bins}: what one
// TMA load of the projection's pipeline fetches.  Extents are T and B, so partial tiles arrive zero-filled.
static int build_magnitude_maps(const ta_plan* plan, const HostBatch& hb, const float* mag, void* d_maps, cudaStream_t stream) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return TA_ERR_UNSUPPORTED;
    std::vector<CUtensorMap> maps(hb.n_tracks);
    const int B = plan->n_bins;
    for (int i = 0; i < hb.n_tracks; ++i) {
        const TrackDesc& t = hb.tracks[i];
        const cuuint64_t dims[2] = {cuuint64_t(t.n_frames), cuuint64_t(B)};
        const cuuint64_t strides[1] = {cuuint64_t(t.ld) * sizeof(float)};
        const cuuint32_t box[2] = {TP_BOX, TP_KT};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(mag) + size_t(t.pitch_off) * B, dims,
                               strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return TA_ERR_UNSUPPORTED;
    }
    TA_CUDA(cudaMemcpyAsync(d_maps, maps.data(), sizeof(CUtensorMap) * hb.n_tracks, cudaMemcpyHostToDevice, stream));
    return TA_OK;
}

template <bool CHROMA, bool ROLL>
static int launch_project(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mag, const float* fb,
                          float* chroma, int32_t* rolloff_bin, const float* frame_sum, void* d_maps, cudaStream_t stream) {
    if (ROLL && !frame_sum) {
        set_error("the roll-off walk needs the per-frame magnitude sums of the STFT stage");
        return TA_ERR_INVALID;
    }
    static const bool use_tma = [] { const char* e = std::getenv("TA_PROJECT"); return !(e && e[0] == 'l'); }();  // TA_PROJECT=ldg
    static const bool use_umma = [] { const char* e = std::getenv("TA_PROJECT"); return e && e[0] == 'u'; }();    // TA_PROJECT=umma (A/B)
    if (use_umma && CHROMA && !ROLL && d_maps) {
        const int n_chunks = (plan->n_bins + UM_KCH - 1) / UM_KCH;
        const size_t smem = 1024 + size_t(UM_STAGES) * UM_STAGE_BYTES + size_t(n_chunks) * UM_KCH * UM_N * 4 + (2 * UM_STAGES + 2) * 8;
        if (smem <= SMEM_OPTIN_LIMIT && build_magnitude_maps_umma(plan, hb, mag, d_maps, stream) == TA_OK) {
            TA_CUDA(cudaFuncSetAttribute(chroma_project_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            chroma_project_umma_kernel<<<dim3((hb.max_frames + UM_FRAMES * UM_TILES - 1) / (UM_FRAMES * UM_TILES), hb.n_tracks), UM_THREADS, smem, stream>>>(
                d_tracks, reinterpret_cast<const CUtensorMap*>(d_maps), fb, chroma, plan->n_bins);
            count_launch();
            TA_CUDA(cudaGetLastError());
            return TA_OK;
        }
    }
    const size_t tma_smem = size_t(TP_STAGES) * TP_STAGE_FLOATS * 4 + (CHROMA ? size_t(plan->n_bins) * 12 * sizeof(float) : 0) +
                            TP_STAGES * 8 + 128;
    if (use_tma && d_maps && tma_smem <= 113 * 1024 && build_magnitude_maps(plan, hb, mag, d_maps, stream) == TA_OK) {
        auto kern = chroma_project_tma_kernel<CHROMA, ROLL>;
        TA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem));
        kern<<<dim3((hb.max_frames + 2 * CP_THREADS - 1) / (2 * CP_THREADS), hb.n_tracks), CP_THREADS, tma_smem, stream>>>(
            d_tracks, reinterpret_cast<const CUtensorMap*>(d_maps), mag, fb, chroma, rolloff_bin, frame_sum,
            float(plan->desc.roll_percent), plan->n_bins);
        count_launch();
        TA_CUDA(cudaGetLastError());
        return TA_OK;
    }
    const size_t smem = CHROMA ? size_t(plan->n_bins) * 12 * sizeof(float) : 0;
    auto kern = chroma_project_kernel<CHROMA, ROLL>;
    TA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((hb.max_frames + 2 * CP_THREADS - 1) / (2 * CP_THREADS), hb.n_tracks), CP_THREADS, smem, stream>>>(
        d_tracks, mag, fb, chroma, rolloff_bin, frame_sum, float(plan->desc.roll_percent), plan->n_bins);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

// chroma (may be NULL) and / or the roll-off bins (may be NULL) from an existing magnitude spectrogram
int run_chroma(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mag, const float* frame_max,
               float* chroma, double* tuning, int32_t* rolloff_bin, const float* frame_sum, void* scratch, size_t scratch_bytes,
               void* d_maps, cudaStream_t stream) {
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    TA_REQUIRE(mag && (reinterpret_cast<uintptr_t>(mag) & 15) == 0, "the magnitude buffer must be present and 16-byte aligned");
    if (!chroma) {
        if (!rolloff_bin) return TA_OK;
        return launch_project<false, true>(plan, hb, d_tracks, mag, nullptr, nullptr, rolloff_bin, frame_sum, d_maps, stream);
    }
    TA_REQUIRE(plan->desc.n_chroma == 12, "only n_chroma = 12 is implemented");
    TA_REQUIRE(frame_max && tuning, "chroma needs the frame_max and tuning buffers");
    TA_REQUIRE(scratch && scratch_bytes >= chroma_scratch_bytes(plan, hb), "chroma scratch too small");
    int rc = run_tuning(plan, hb, d_tracks, mag, frame_max, true, 12, scratch, tuning, nullptr, stream);
    if (rc != TA_OK) return rc;
    float* fb = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + tuning_scratch_bytes(plan, hb));
    chroma_fb_kernel<<<dim3((plan->n_bins + 255) / 256, hb.n_tracks), 256, 0, stream>>>(tuning, fb, plan->n_bins, plan->desc.n_fft,
                                                                                      double(plan->desc.sample_rate));
    count_launch();
    TA_CUDA(cudaGetLastError());
    TA_REQUIRE((reinterpret_cast<uintptr_t>(chroma) & 15) == 0, "the chroma buffer must be 16-byte aligned");
    return rolloff_bin ? launch_project<true, true>(plan, hb, d_tracks, mag, fb, chroma, rolloff_bin, frame_sum, d_maps, stream)
                       : launch_project<true, false>(plan, hb, d_tracks, mag, fb, chroma, nullptr, nullptr, d_maps, stream);
}

}  // namespace ta
