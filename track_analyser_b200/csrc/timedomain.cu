// K5 + K6: one pass over the PCM for everything that is computed in the time domain.
//
//   K5  BS.1770 K-weighting (pyloudnorm.Meter.integrated_loudness, loudness.py:60-61):
//       two cascaded biquads evaluated as a parallel linear-recurrence scan in
//       float64 -- each thread runs the direct-form-II-transposed recurrence over
//       8 consecutive samples from zero state, the (z0,z1) end states are combined
//       by a Kogge-Stone warp scan with precomputed powers of the 2x2 transition
//       matrix, warp totals are chained through shared memory, and the zero-input
//       response of the true start state is added back.  The stage-1 output is
//       rounded to float32 before stage 2 and the stage-2 output before squaring,
//       exactly where scipy.signal.lfilter's float64 result is stored back into
//       pyloudnorm's float32 working copy.  Chunks of `cs` samples are independent
//       CTAs: each warms the filters up over the preceding HALO samples from zero
//       state (the slowest pole pair has |p| = 0.9946; after 8192 samples the
//       truncation error is below 1e-15 of the state).
//   K6  mid/side/L/R moments (stereo.py:62-83, loudness.py:118) and hop-granule
//       sums of mono^2 for the centred RMS frames (loudness.py:30-42).
//   finalize: gating-block energies z_j, absolute/relative gates -> LUFS, RMS frames.
// HBM-bound target: algorithmic bytes 4*C*N read per track; FP64 work ~33 DFMA/sample.
#include <cmath>
#include <numeric>

#include "common.cuh"

namespace ta {

static constexpr int TD_THREADS = 256;
static constexpr int TD_SEG = 8;                       // samples per thread per iteration
static constexpr int TD_ITER = TD_THREADS * TD_SEG;    // 2048 samples per iteration
#ifndef TD_MIN_BLOCKS
#define TD_MIN_BLOCKS 2
#endif
static constexpr int TD_MAXG = 1024;                   // granule accumulators per set per CTA

struct Mat2 {
    double m00, m01, m10, m11;
};

struct StageConst {
    double b0, b1, b2, a1, a2;
    Mat2 P[5];      // A^(8*2^d), d = 0..4 : Kogge-Stone steps inside a warp
    Mat2 W;         // A^256 : one warp
    Mat2 AL[32];    // A^(8*lane)
};

struct TdParams {
    const TrackDesc* tracks;
    int n_tracks;
    int total_chunks;
    int cs;        // chunk samples (multiple of TD_ITER)
    int halo;      // warm-up samples before each chunk (multiple of TD_ITER)
    int stereo;
    int g_k, g_m, g_s;            // granule sizes: K-weighted, momentary hop, short-term hop
    int pitch_k, pitch_m, pitch_s;
    double* gran_k;               // [n_tracks][pitch_k]
    double* gran_m;
    double* gran_s;
    double* moments;              // [n_tracks][8]
    StageConst stage[2];          // by value (kernel parameter space): no cross-plan races
};

__host__ __device__ inline Mat2 matmul(const Mat2& a, const Mat2& b) {
    return {a.m00 * b.m00 + a.m01 * b.m10, a.m00 * b.m01 + a.m01 * b.m11, a.m10 * b.m00 + a.m11 * b.m10,
            a.m10 * b.m01 + a.m11 * b.m11};
}

__device__ __forceinline__ void matvec_add(const Mat2& M, double x0, double x1, double& y0, double& y1) {
    y0 = fma(M.m00, x0, fma(M.m01, x1, y0));
    y1 = fma(M.m10, x0, fma(M.m11, x1, y1));
}

// Stage constants as laid out in shared memory (copied once per CTA: reading the by-value kernel
// parameter block through LDC at ~70 distinct addresses thrashed the constant cache -- profiles/r1).
struct StageSm {
    double b0, b1, b2, a1, a2, pad;
    Mat2 P[5];
    Mat2 W;
    double al[4][32];  // A^(8*lane), one row per matrix element: conflict-free lane-indexed reads
};

// One biquad stage over the thread's TD_SEG samples; x is replaced by the filter output (double).
// carry0/carry1: filter state at the start of this iteration (updated to the state at its end).
__device__ __forceinline__ void biquad_stage(const StageSm& c, double (&x)[TD_SEG], double& carry0,
                                             double& carry1, double2* wt /* [8] warp totals */, int lane, int warp) {
    // 1. zero-state response
    const double b0 = c.b0, b1 = c.b1, b2 = c.b2, na1 = -c.a1, na2 = -c.a2;
    double z0 = 0.0, z1 = 0.0;
#pragma unroll
    for (int i = 0; i < TD_SEG; ++i) {
        const double xi = x[i];
        const double y = fma(b0, xi, z0);
        z0 = fma(na1, y, fma(b1, xi, z1));
        z1 = fma(na2, y, b2 * xi);
        x[i] = y;
    }
    // 2. inclusive scan of end states inside the warp
    double e0 = z0, e1 = z1;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        const int o = 1 << d;
        const double p0 = __shfl_up_sync(0xffffffffu, e0, o);
        const double p1 = __shfl_up_sync(0xffffffffu, e1, o);
        if (lane >= o) matvec_add(c.P[d], p0, p1, e0, e1);
    }
    if (lane == 31) wt[warp] = make_double2(e0, e1);
    __syncthreads();
    // 3. chain warp totals from the iteration carry; remember the state at this warp's start
    double s0 = carry0, s1 = carry1, w0 = 0.0, w1 = 0.0;
#pragma unroll
    for (int ww = 0; ww < TD_THREADS / 32; ++ww) {
        if (ww == warp) { w0 = s0; w1 = s1; }
        const double2 t = wt[ww];
        double n0 = t.x, n1 = t.y;
        matvec_add(c.W, s0, s1, n0, n1);
        s0 = n0;
        s1 = n1;
    }
    carry0 = s0;
    carry1 = s1;
    // 4. true state at the start of this thread's segment
    double i0 = __shfl_up_sync(0xffffffffu, e0, 1);
    double i1 = __shfl_up_sync(0xffffffffu, e1, 1);
    if (lane == 0) { i0 = 0.0; i1 = 0.0; }
    {
        const Mat2 al{c.al[0][lane], c.al[1][lane], c.al[2][lane], c.al[3][lane]};
        matvec_add(al, w0, w1, i0, i1);
    }
    // 5. add the zero-input response
#pragma unroll
    for (int i = 0; i < TD_SEG; ++i) {
        const double yi = i0;
        x[i] += yi;
        i0 = fma(na1, yi, i1);
        i1 = na2 * yi;
    }
}

// n / g for n < 2^31 without the ~20-instruction integer divide: double reciprocal + one correction.
__device__ __forceinline__ unsigned fast_div(unsigned n, unsigned g, double inv_g) {
    unsigned q = unsigned(double(n) * inv_g);
    if (q * g > n) --q;
    else if ((q + 1) * g <= n) ++q;
    return q;
}

// Running granule sum kept in registers: all lanes of a warp sit in the same granule most of the
// time, so each thread adds its 8 values locally and the warp reduces + publishes (one shared-memory
// atomic) only when its granule changes.
struct GranAcc {
    double acc = 0.0;
    unsigned gid = 0xffffffffu;
};

__device__ __forceinline__ void gran_flush(GranAcc& a, double* sm_acc, int first_gid, int lane) {
    if (a.gid == 0xffffffffu) return;  // warp-uniform
    double v = a.acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) atomicAdd(&sm_acc[int(a.gid) - first_gid], v);
    a.acc = 0.0;
    a.gid = 0xffffffffu;
}

// Adds the thread's TD_SEG values q[] (sample indices n_first..n_first+7).  `all_ok`: every sample of
// the warp is inside the chunk (warp-uniform).
__device__ __forceinline__ void granule_add(GranAcc& a, double* sm_acc, int first_gid, unsigned g, double inv_g, unsigned n_first,
                                            const float (&q)[TD_SEG], const bool (&ok)[TD_SEG], bool all_ok, int lane) {
    const unsigned n_warp = n_first - unsigned(lane) * TD_SEG;  // first sample of this warp
    const unsigned gid_w = fast_div(n_warp, g, inv_g);          // warp-uniform
    if (all_ok && n_warp + 32 * TD_SEG <= (gid_w + 1) * g) {
        if (gid_w != a.gid) {
            gran_flush(a, sm_acc, first_gid, lane);
            a.gid = gid_w;
        }
        a.acc += (double(q[0]) + double(q[1])) + (double(q[2]) + double(q[3])) +
                 ((double(q[4]) + double(q[5])) + (double(q[6]) + double(q[7])));
        return;
    }
    gran_flush(a, sm_acc, first_gid, lane);
    const unsigned gid = fast_div(n_first, g, inv_g);
    const unsigned boundary = (gid + 1) * g;  // first sample index of the next granule
    double lo = 0.0, hi = 0.0;
    bool any = false;
#pragma unroll
    for (int i = 0; i < TD_SEG; ++i) {
        if (ok[i]) {
            any = true;
            if (n_first + i < boundary) lo += double(q[i]); else hi += double(q[i]);
        }
    }
    if (any) {
        atomicAdd(&sm_acc[int(gid) - first_gid], lo);
        if (n_first + TD_SEG > boundary) atomicAdd(&sm_acc[int(gid) + 1 - first_gid], hi);
    }
}

__global__ void __launch_bounds__(TD_THREADS, TD_MIN_BLOCKS) time_domain_kernel(const __grid_constant__ TdParams p) {
    __shared__ double2 wt[2][TD_THREADS / 32];
    __shared__ double gacc[3][TD_MAXG];
    __shared__ double red[TD_THREADS / 32][7];
    __shared__ StageSm cst[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2) {
        const StageConst& c = p.stage[tid];
        StageSm& d = cst[tid];
        d.b0 = c.b0; d.b1 = c.b1; d.b2 = c.b2; d.a1 = c.a1; d.a2 = c.a2; d.pad = 0.0;
        for (int i = 0; i < 5; ++i) d.P[i] = c.P[i];
        d.W = c.W;
    }
    if (tid < 64) {  // lane-indexed powers: a divergent constant-bank read, done once per CTA
        const Mat2 m = p.stage[tid >> 5].AL[tid & 31];
        StageSm& d = cst[tid >> 5];
        d.al[0][tid & 31] = m.m00; d.al[1][tid & 31] = m.m01; d.al[2][tid & 31] = m.m10; d.al[3][tid & 31] = m.m11;
    }
    __syncthreads();

    for (int w = blockIdx.x; w < p.total_chunks; w += gridDim.x) {
        // locate track
        int lo_t = 0, hi_t = p.n_tracks - 1;
        while (lo_t < hi_t) {
            const int mid = (lo_t + hi_t + 1) >> 1;
            if (p.tracks[mid].chunk_begin <= w) lo_t = mid; else hi_t = mid - 1;
        }
        const int trk = lo_t;
        const TrackDesc td = p.tracks[trk];
        const long long cs0 = (long long)(w - td.chunk_begin) * p.cs;
        const long long ce = min(cs0 + (long long)p.cs, (long long)td.n_samples);
        const long long ws = max(0ll, cs0 - (long long)p.halo);
        const int fg_k = int(cs0 / p.g_k), fg_m = int(cs0 / p.g_m), fg_s = int(cs0 / p.g_s);
        for (int i = tid; i < 3 * TD_MAXG; i += TD_THREADS) (&gacc[0][0])[i] = 0.0;
        __syncthreads();

        const float* __restrict__ L = td.ch0;
        const float* __restrict__ R = td.ch1;
        const bool vec = ((reinterpret_cast<uintptr_t>(L) & 15) == 0) && (!p.stereo || (reinterpret_cast<uintptr_t>(R) & 15) == 0);
        double c10 = 0, c11 = 0, c20 = 0, c21 = 0;  // carries of stage 1 / stage 2
        GranAcc ga_k, ga_m, ga_s;
        const double inv_k = 1.0 / double(p.g_k), inv_m = 1.0 / double(p.g_m), inv_s = 1.0 / double(p.g_s);
        double sL = 0, sR = 0, sLL = 0, sRR = 0, sLR = 0, sMM = 0, sSS = 0;

        for (long long n0 = ws; n0 < ce; n0 += TD_ITER) {
            const long long nf = n0 + (long long)tid * TD_SEG;
            float l[TD_SEG], r[TD_SEG];
            if (vec && nf + TD_SEG <= td.n_samples) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(L + nf));
                const float4 b = __ldg(reinterpret_cast<const float4*>(L + nf) + 1);
                l[0] = a.x; l[1] = a.y; l[2] = a.z; l[3] = a.w; l[4] = b.x; l[5] = b.y; l[6] = b.z; l[7] = b.w;
                if (p.stereo) {
                    const float4 c = __ldg(reinterpret_cast<const float4*>(R + nf));
                    const float4 d = __ldg(reinterpret_cast<const float4*>(R + nf) + 1);
                    r[0] = c.x; r[1] = c.y; r[2] = c.z; r[3] = c.w; r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = d.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < TD_SEG; ++i) {
                    const bool in = nf + i < td.n_samples;
                    l[i] = in ? __ldg(L + nf + i) : 0.f;
                    r[i] = (in && p.stereo) ? __ldg(R + nf + i) : 0.f;
                }
            }
            // ---- raw-sample statistics first, so l/r/mono are dead before the filter stages ----
            const bool warm = n0 + TD_ITER <= cs0;  // warm-up iteration: nothing is accumulated (CTA-uniform)
            const long long nw = n0 + (long long)warp * 32 * TD_SEG;
            const bool all_ok = nw >= cs0 && nw + 32 * TD_SEG <= ce;  // warp-uniform
            const unsigned n32 = unsigned(nf);
            bool ok[TD_SEG];
            double x[TD_SEG];
            {
                float qm[TD_SEG];
#pragma unroll
                for (int i = 0; i < TD_SEG; ++i) {
                    const float mono = p.stereo ? 0.5f * (l[i] + r[i]) : l[i];
                    ok[i] = all_ok || ((nf + i >= cs0) && (nf + i < ce));
                    x[i] = double(mono);
                    qm[i] = mono * mono;
                }
                if (!warm) {
#pragma unroll
                    for (int i = 0; i < TD_SEG; ++i) {
                        if (!all_ok && !ok[i]) continue;
                        const double dm = x[i], dl = l[i];
                        sMM = fma(dm, dm, sMM);
                        sL += dl;
                        sLL = fma(dl, dl, sLL);
                        if (p.stereo) {
                            const double dr = r[i], sd = double(0.5f * (l[i] - r[i]));
                            sR += dr;
                            sRR = fma(dr, dr, sRR);
                            sLR = fma(dl, dr, sLR);
                            sSS = fma(sd, sd, sSS);
                        }
                    }
                    if (p.gran_m) granule_add(ga_m, gacc[1], fg_m, unsigned(p.g_m), inv_m, n32, qm, ok, all_ok, lane);
                    if (p.gran_s) granule_add(ga_s, gacc[2], fg_s, unsigned(p.g_s), inv_s, n32, qm, ok, all_ok, lane);
                }
            }
            // ---- K-weighting: shelf, float32 round trip, high-pass, float32 round trip ----
            biquad_stage(cst[0], x, c10, c11, wt[0], lane, warp);
#pragma unroll
            for (int i = 0; i < TD_SEG; ++i) x[i] = double(float(x[i]));
            biquad_stage(cst[1], x, c20, c21, wt[1], lane, warp);
            if (!warm && p.gran_k) {
                float qk[TD_SEG];
#pragma unroll
                for (int i = 0; i < TD_SEG; ++i) {
                    const float y = float(x[i]);
                    qk[i] = y * y;
                }
                granule_add(ga_k, gacc[0], fg_k, unsigned(p.g_k), inv_k, n32, qk, ok, all_ok, lane);
            }
        }
        gran_flush(ga_k, gacc[0], fg_k, lane);
        gran_flush(ga_m, gacc[1], fg_m, lane);
        gran_flush(ga_s, gacc[2], fg_s, lane);
        __syncthreads();
        // flush granules (a granule is shared by at most two chunks -> a + b is order independent)
        const int ng_k = int((ce - 1) / p.g_k) - fg_k + 1, ng_m = int((ce - 1) / p.g_m) - fg_m + 1,
                  ng_s = int((ce - 1) / p.g_s) - fg_s + 1;
        if (ce > cs0) {
            if (p.gran_k)
                for (int i = tid; i < ng_k; i += TD_THREADS) atomicAdd(&p.gran_k[size_t(trk) * p.pitch_k + fg_k + i], gacc[0][i]);
            if (p.gran_m)
                for (int i = tid; i < ng_m; i += TD_THREADS) atomicAdd(&p.gran_m[size_t(trk) * p.pitch_m + fg_m + i], gacc[1][i]);
            if (p.gran_s)
                for (int i = tid; i < ng_s; i += TD_THREADS) atomicAdd(&p.gran_s[size_t(trk) * p.pitch_s + fg_s + i], gacc[2][i]);
        }
        // moments
        if (p.moments) {
            double v[7] = {sL, sR, sLL, sRR, sLR, sMM, sSS};
#pragma unroll
            for (int k = 0; k < 7; ++k) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
                if (lane == 0) red[warp][k] = v[k];
            }
            __syncthreads();
            if (tid < 7) {
                double s = 0.0;
                for (int ww = 0; ww < TD_THREADS / 32; ++ww) s += red[ww][tid];
                atomicAdd(&p.moments[size_t(trk) * 8 + tid], s);
            }
        }
        __syncthreads();
    }
}

// ---- finalize: one CTA per track --------------------------------------------------
struct FinParams {
    const TrackDesc* tracks;
    const double* gran_k;
    const double* gran_m;
    const double* gran_s;
    int pitch_k, pitch_m, pitch_s;
    int g_k, g_m, g_s;
    int stereo;
    double T_g, rate;
    double* kw_blocks; int kw_pitch;
    double* lufs;
    double* rms_m; double* rms_s; int rms_pitch;
    int frame_m, frame_s;
    double* moments;
};

__device__ double block_sum_d(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    return s;
}

__global__ void __launch_bounds__(256) time_finalize_kernel(const FinParams p) {
    __shared__ double sh[8];
    const int trk = blockIdx.x;
    const TrackDesc td = p.tracks[trk];
    const double ns = double(td.n_samples);
    if (p.moments && threadIdx.x == 0) p.moments[size_t(trk) * 8 + 7] = ns;
    // RMS frames: frame j covers hop granules j-1 and j
    for (int s = 0; s < 2; ++s) {
        double* dst = s ? p.rms_s : p.rms_m;
        const double* gr = s ? p.gran_s : p.gran_m;
        if (!dst || !gr) continue;
        const int hop = s ? p.g_s : p.g_m, frame = s ? p.frame_s : p.frame_m, pitch = s ? p.pitch_s : p.pitch_m;
        const long long nfr = 1 + td.n_samples / hop;
        for (long long j = threadIdx.x; j < nfr && j < p.rms_pitch; j += blockDim.x) {
            const double a = j > 0 ? gr[size_t(trk) * pitch + j - 1] : 0.0;
            const double b = gr[size_t(trk) * pitch + j];
            dst[size_t(trk) * p.rms_pitch + j] = (a + b) / double(frame);
        }
    }
    if (!p.gran_k) return;
    // gating blocks (pyloudnorm: z_j = 1/(T_g*rate) * sum x^2 over [l_j, u_j))
    const double T = ns / p.rate;
    long long nb = 0;
    if (ns >= p.T_g * p.rate) nb = (long long)(rint((T - p.T_g) / (p.T_g * 0.25)) + 1);
    const double inv = 1.0 / (p.T_g * p.rate);
    double sum_abs = 0.0, cnt_abs = 0.0;
    for (long long j = threadIdx.x; j < nb; j += blockDim.x) {
        const long long l = (long long)(p.T_g * (double(j) * 0.25) * p.rate);
        const long long u = (long long)(p.T_g * (double(j) * 0.25 + 1) * p.rate);
        double acc = 0.0;
        for (long long q = l / p.g_k; q < u / p.g_k; ++q) acc += p.gran_k[size_t(trk) * p.pitch_k + q];
        const double z = inv * acc;
        if (p.kw_blocks && j < p.kw_pitch) p.kw_blocks[size_t(trk) * p.kw_pitch + j] = z;
        const double lj = -0.691 + 10.0 * log10(z);
        if (lj >= -70.0) { sum_abs += z; cnt_abs += 1.0; }
    }
    sum_abs = block_sum_d(sum_abs, sh);
    cnt_abs = block_sum_d(cnt_abs, sh);
    const double gamma_r = -0.691 + 10.0 * log10(sum_abs / cnt_abs) - 10.0;  // NaN when nothing passes, like numpy
    double sum_rel = 0.0, cnt_rel = 0.0;
    for (long long j = threadIdx.x; j < nb; j += blockDim.x) {
        const long long l = (long long)(p.T_g * (double(j) * 0.25) * p.rate);
        const long long u = (long long)(p.T_g * (double(j) * 0.25 + 1) * p.rate);
        double acc = 0.0;
        for (long long q = l / p.g_k; q < u / p.g_k; ++q) acc += p.gran_k[size_t(trk) * p.pitch_k + q];
        const double z = inv * acc;
        const double lj = -0.691 + 10.0 * log10(z);
        if (lj > gamma_r && lj > -70.0) { sum_rel += z; cnt_rel += 1.0; }
    }
    sum_rel = block_sum_d(sum_rel, sh);
    cnt_rel = block_sum_d(cnt_rel, sh);
    if (threadIdx.x == 0 && p.lufs) {
        double zavg = (cnt_rel > 0.0) ? sum_rel / cnt_rel : 0.0;  // nan_to_num(mean([])) == 0
        if (nb == 0) p.lufs[trk] = nan("");                       // reference raises ValueError for short input
        else p.lufs[trk] = -0.691 + 10.0 * log10(zavg);
    }
}

// ---- host side ------------------------------------------------------------------

static Mat2 mat_pow(Mat2 a, long long e) {
    Mat2 r{1, 0, 0, 1};
    while (e) {
        if (e & 1) r = matmul(r, a);
        a = matmul(a, a);
        e >>= 1;
    }
    return r;
}

static StageConst make_stage(const Biquad& b) {
    StageConst c{};
    c.b0 = b.b0; c.b1 = b.b1; c.b2 = b.b2; c.a1 = b.a1; c.a2 = b.a2;
    const Mat2 A{-b.a1, 1.0, -b.a2, 0.0};
    for (int d = 0; d < 5; ++d) c.P[d] = mat_pow(A, (long long)TD_SEG << d);
    c.W = mat_pow(A, (long long)TD_SEG * 32);
    for (int l = 0; l < 32; ++l) c.AL[l] = mat_pow(A, (long long)TD_SEG * l);
    return c;
}

void kw_block_bounds(const ta_plan* plan, int64_t n_samples, std::vector<int64_t>& lo, std::vector<int64_t>& hi);

// Largest granule that divides every gating-block bound of a track of `max_samples` samples.
int kw_granule_for(const ta_plan* plan, int64_t max_samples) {
    std::vector<int64_t> lo, hi;
    kw_block_bounds(plan, max_samples, lo, hi);
    int64_t g = 0;
    for (size_t j = 0; j < lo.size(); ++j) {
        g = std::gcd(g, lo[j]);
        g = std::gcd(g, hi[j]);
    }
    if (g == 0) g = int64_t(std::nearbyint(double(plan->desc.meter_block) * plan->desc.sample_rate * 0.25));
    return int(std::max<int64_t>(g, 1));
}

static void td_pitches(const ta_plan* plan, const HostBatch& hb, int& g_k, int& pk, int& pm, int& ps) {
    int64_t max_samples = 0;
    for (auto& t : hb.tracks) max_samples = std::max(max_samples, t.n_samples);
    g_k = kw_granule_for(plan, max_samples);
    pk = int(max_samples / g_k + 2);
    pm = int(max_samples / plan->rms_m_hop + 2);
    ps = int(max_samples / plan->rms_s_hop + 2);
}

size_t td_granule_doubles(const ta_plan* plan, const HostBatch& hb) {
    int g, pk, pm, ps;
    td_pitches(plan, hb, g, pk, pm, ps);
    return size_t(hb.n_tracks) * (size_t(pk) + pm + ps);
}

int time_chunk_samples(const ta_plan* plan, int64_t total_samples) {
    // aim for >= 4 chunks per SM, chunk in [8192, 65536], multiple of TD_ITER
    const int64_t target = total_samples / (int64_t(plan->sm_count) * 4) + 1;
    int64_t cs = ((target + TD_ITER - 1) / TD_ITER) * TD_ITER;
    cs = std::min<int64_t>(std::max<int64_t>(cs, 8192), 65536);
    return int(cs);
}

int run_time_domain(const ta_plan* plan, const HostBatch& hb, const Workspace& ws, const ta_frontend_out* out,
                    cudaStream_t stream) {
    int64_t max_samples = 0;
    for (auto& t : hb.tracks) max_samples = std::max(max_samples, t.n_samples);
    TA_REQUIRE(max_samples < (int64_t(1) << 31), "tracks longer than 2^31 samples are not supported");
    const bool want_k = out->kw_blocks || out->lufs;
    const bool want_m = out->rms_momentary != nullptr, want_s = out->rms_short != nullptr;
    const int cs = time_chunk_samples(plan, hb.total_samples);
    TdParams p{};
    p.tracks = ws.d_tracks;
    p.n_tracks = hb.n_tracks;
    p.total_chunks = hb.total_chunks;
    p.cs = cs;
    p.stereo = hb.channels == 2;
    td_pitches(plan, hb, p.g_k, p.pitch_k, p.pitch_m, p.pitch_s);
    p.g_m = plan->rms_m_hop;
    p.g_s = plan->rms_s_hop;
    if (want_k && (p.g_k < 8 || cs / p.g_k + 2 > TD_MAXG)) {
        set_error("gating-block bounds at this sample rate / block size have no common granule >= 8 samples");
        return TA_ERR_UNSUPPORTED;
    }
    TA_REQUIRE(p.g_m >= 8 && p.g_s >= 8, "RMS hop too small");
    TA_REQUIRE(cs / p.g_m + 2 <= TD_MAXG && cs / p.g_s + 2 <= TD_MAXG, "RMS hop too small for the chunk size");
    const size_t need = size_t(hb.n_tracks) * (size_t(p.pitch_k) + p.pitch_m + p.pitch_s);
    TA_REQUIRE(need <= ws.gran_doubles, "granule workspace too small");
    p.gran_k = want_k ? ws.d_granules : nullptr;
    p.gran_m = want_m ? ws.d_granules + size_t(hb.n_tracks) * p.pitch_k : nullptr;
    p.gran_s = want_s ? ws.d_granules + size_t(hb.n_tracks) * (size_t(p.pitch_k) + p.pitch_m) : nullptr;
    p.moments = out->moments;
    TA_CUDA(cudaMemsetAsync(ws.d_granules, 0, sizeof(double) * need, stream));
    if (p.moments) TA_CUDA(cudaMemsetAsync(p.moments, 0, sizeof(double) * 8 * hb.n_tracks, stream));

    // warm-up length: the slowest K-weighting pole pair is the high-pass double pole of radius r = sqrt(a2);
    // a zero-input response decays like n*r^n, so take the first multiple of TD_ITER with n*r^n < 1e-10.
    {
        const double r = std::sqrt(std::fabs(plan->highpass.a2));
        int n = TD_ITER;
        while (n < (1 << 22) && double(n) * std::pow(r, double(n)) > 1e-10) n += TD_ITER;
        p.halo = n;
    }
    p.stage[0] = make_stage(plan->shelf);
    p.stage[1] = make_stage(plan->highpass);

    const int grid = std::max(1, std::min(hb.total_chunks, plan->sm_count * 8));
    time_domain_kernel<<<grid, TD_THREADS, 0, stream>>>(p);
    count_launch();
    TA_CUDA(cudaGetLastError());

    FinParams f{};
    f.tracks = ws.d_tracks;
    f.gran_k = p.gran_k; f.gran_m = p.gran_m; f.gran_s = p.gran_s;
    f.pitch_k = p.pitch_k; f.pitch_m = p.pitch_m; f.pitch_s = p.pitch_s;
    f.g_k = p.g_k; f.g_m = p.g_m; f.g_s = p.g_s;
    f.stereo = p.stereo;
    f.T_g = double(plan->desc.meter_block);
    f.rate = double(plan->desc.sample_rate);
    f.kw_blocks = out->kw_blocks; f.kw_pitch = out->kw_pitch;
    f.lufs = out->lufs;
    f.rms_m = out->rms_momentary; f.rms_s = out->rms_short; f.rms_pitch = out->rms_pitch;
    f.frame_m = plan->rms_m_frame; f.frame_s = plan->rms_s_frame;
    f.moments = out->moments;
    time_finalize_kernel<<<hb.n_tracks, 256, 0, stream>>>(f);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
