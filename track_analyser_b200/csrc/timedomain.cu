// K5 + K6: one pass over the PCM for everything that is computed in the time domain.
//
//   K5  BS.1770 K-weighting (pyloudnorm.Meter.integrated_loudness, loudness.py:60-61):
//       two cascaded biquads evaluated as a parallel linear-recurrence scan in float64.
//       Every WARP owns a contiguous chunk of one track and streams through it 256 samples
//       at a time (8 per lane): each lane runs the direct-form-II-transposed recurrence over
//       its 8 samples from zero state, the (z0,z1) end states are combined by a Kogge-Stone
//       warp scan with precomputed powers of the 2x2 transition matrix, and the zero-input
//       response of the true start state (previous lanes' end state + A^(8*lane) * carry) is
//       added back with precomputed response vectors -- 16 independent DFMAs instead of a
//       second dependent recurrence.  No shared memory, no block barrier: the carry lives in
//       registers.  The stage-1 output is rounded to float32 precision before stage 2 and the
//       stage-2 output before squaring, exactly where scipy.signal.lfilter's float64 result is
//       stored back into pyloudnorm's float32 working copy; the rounding is a Veltkamp split
//       (3 DFMA-class ops) because float<->double conversions run at 1/8 of the DFMA rate.
//       Chunks are independent: each warms the filters up over the preceding HALO samples
//       from zero state (the slowest pole pair has |p| = 0.9946; n*r^n < 1e-10 after 6144).
//   K6  mid/side/L/R moments (stereo.py:62-83, loudness.py:118) and hop-granule sums of mono^2
//       for the centred RMS frames (loudness.py:30-42): float32 partial sums over a lane's 8
//       samples, accumulated in float64.
//   finalize: gating-block energies z_j, absolute/relative gates -> LUFS, RMS frames.
// HBM-bound target: algorithmic bytes 4*C*N read per track; FP64 work ~27 DFMA/sample.
#include <cmath>
#include <numeric>

#include "common.cuh"

namespace ta {

static constexpr int TD_SEG = 8;                 // samples per lane per step
static constexpr int TD_STEP = 32 * TD_SEG;      // samples per warp per step
static constexpr int TD_WARPS = 4;               // warps per CTA
static constexpr int TD_THREADS = 32 * TD_WARPS;

struct Mat2 {
    double m00, m01, m10, m11;
};

// Uniform per-stage constants; passed by value so that they are read as constant-bank operands.
struct StageU {
    double b0, b1, b2, na1, na2;
    Mat2 P[5];            // A^(8*2^d), d = 0..4: Kogge-Stone steps inside the warp
    Mat2 W;               // A^256: one warp step
    double H0[TD_SEG];    // zero-input response y_i for start state (1, 0): (A^i)[0][0]
    double H1[TD_SEG];    //                          and for (0, 1): (A^i)[0][1]
};

struct TdParams {
    const TrackDesc* tracks;
    int n_tracks;
    int total_chunks;
    int cs;        // chunk samples per warp (multiple of TD_STEP)
    int halo;      // warm-up samples before each chunk (multiple of TD_STEP)
    int stereo;
    int g_k, g_m, g_s;            // granule sizes: K-weighted, momentary hop, short-term hop
    int pitch_k, pitch_m, pitch_s;
    double* gran_k;               // [n_tracks][pitch_k]
    double* gran_m;
    double* gran_s;
    double* moments;              // [n_tracks][TA_N_MOMENTS]
    const double* lane_pow;       // [2 stages][32 lanes][4]: A^(8*lane)
    float* blk_absmax;            // [n_tracks][blk_pitch] max |mono| of each 256-sample step (nullable)
    uint32_t* absmax_bits;        // [n_tracks] float bits of max |mono|
    int blk_pitch;
    StageU stage[2];
};

__host__ __device__ inline Mat2 matmul(const Mat2& a, const Mat2& b) {
    return {a.m00 * b.m00 + a.m01 * b.m10, a.m00 * b.m01 + a.m01 * b.m11, a.m10 * b.m00 + a.m11 * b.m10,
            a.m10 * b.m01 + a.m11 * b.m11};
}

__device__ __forceinline__ void matvec_add(const Mat2& M, double x0, double x1, double& y0, double& y1) {
    y0 = fma(M.m00, x0, fma(M.m01, x1, y0));
    y1 = fma(M.m10, x0, fma(M.m11, x1, y1));
}

// y rounded to 24 significant bits (round to nearest) without leaving the FP64 pipe: Veltkamp split with
// s = 29.  Equals double(float(y)) for every y in float32's normal range except exact ties, which cannot
// change any sum at the tolerances of this path.
__device__ __forceinline__ double round_f32(double y) {
    const double c = y * 536870913.0;  // 2^29 + 1
    return c - (c - y);
}

// One biquad stage over the lane's TD_SEG samples, part 1: zero-state response and the warp-inclusive scan
// of end states.  x is replaced by the zero-state output; (e0, e1) is the state after this lane's segment
// given a zero state at the start of the warp's 256-sample block.
__device__ __forceinline__ void stage_zero_state(const StageU& c, double (&x)[TD_SEG], double& e0, double& e1, int lane) {
    double z0 = 0.0, z1 = 0.0;
#pragma unroll
    for (int i = 0; i < TD_SEG; ++i) {
        const double xi = x[i];
        const double y = fma(c.b0, xi, z0);
        z0 = fma(c.na1, y, fma(c.b1, xi, z1));
        z1 = fma(c.na2, y, c.b2 * xi);
        x[i] = y;
    }
    e0 = z0;
    e1 = z1;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        const int o = 1 << d;
        const double p0 = __shfl_up_sync(0xffffffffu, e0, o);
        const double p1 = __shfl_up_sync(0xffffffffu, e1, o);
        if (lane >= o) matvec_add(c.P[d], p0, p1, e0, e1);
    }
}

// Part 2: add the zero-input response of the true start state and advance the carry (state at the start of
// the warp's block -> state at its end).
__device__ __forceinline__ void stage_correct(const StageU& c, const Mat2& al, double (&x)[TD_SEG], double e0, double e1,
                                              double& carry0, double& carry1, int lane) {
    double s0 = __shfl_up_sync(0xffffffffu, e0, 1);
    double s1 = __shfl_up_sync(0xffffffffu, e1, 1);
    if (lane == 0) { s0 = 0.0; s1 = 0.0; }
    matvec_add(al, carry0, carry1, s0, s1);
#pragma unroll
    for (int i = 0; i < TD_SEG; ++i) x[i] = fma(c.H0[i], s0, fma(c.H1[i], s1, x[i]));
    double n0 = __shfl_sync(0xffffffffu, e0, 31), n1 = __shfl_sync(0xffffffffu, e1, 31);
    matvec_add(c.W, carry0, carry1, n0, n1);
    carry0 = n0;
    carry1 = n1;
}

// n / g for n < 2^31 without the ~20-instruction integer divide: double reciprocal + one correction.
__device__ __forceinline__ unsigned fast_div(unsigned n, unsigned g, double inv_g) {
    unsigned q = unsigned(double(n) * inv_g);
    if (q * g > n) --q;
    else if ((q + 1) * g <= n) ++q;
    return q;
}

// Running granule sum of a warp: all lanes sit in the same granule most of the time, so each lane adds its
// partial sum locally and the warp reduces + publishes (one global atomic) only when its granule changes.
struct GranAcc {
    double acc = 0.0;
    unsigned gid = 0xffffffffu;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void gran_flush(GranAcc& a, double* gran, int lane) {
    if (a.gid == 0xffffffffu) return;  // warp-uniform
    const double v = warp_sum(a.acc);
    if (lane == 0) atomicAdd(&gran[a.gid], v);
    a.acc = 0.0;
    a.gid = 0xffffffffu;
}

// Adds the lane's TD_SEG values q[] (sample indices n_first .. n_first+7; all lanes call this).
// `full`: every sample of the warp's block is a real sample (warp-uniform); otherwise n_valid counts them.
__device__ __forceinline__ void granule_add(GranAcc& a, double* gran, unsigned g, double inv_g, unsigned n_block,
                                            const double (&q)[TD_SEG], bool full, long long n_samples, int lane) {
    const unsigned gid_w = fast_div(n_block, g, inv_g);  // granule of the block's first sample (warp-uniform)
    if (full && n_block + TD_STEP <= (gid_w + 1) * g) {
        if (gid_w != a.gid) {
            gran_flush(a, gran, lane);
            a.gid = gid_w;
        }
        a.acc += ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
        return;
    }
    // slow path (a granule boundary or the end of the track inside this block): per-sample granule index,
    // one atomic per distinct granule per warp (a block spans at most 1 + TD_STEP/g boundaries)
    gran_flush(a, gran, lane);
    const unsigned n_first = n_block + unsigned(lane) * TD_SEG;
    unsigned gid = gid_w;
    while (true) {
        const unsigned lo = gid * g, hi = lo + g;
        double part = 0.0;
#pragma unroll
        for (int i = 0; i < TD_SEG; ++i) {
            const unsigned n = n_first + i;
            if (n >= lo && n < hi && (long long)n < n_samples) part += q[i];
        }
        part = warp_sum(part);
        if (lane == 0 && part != 0.0) atomicAdd(&gran[gid], part);
        if (hi >= n_block + TD_STEP) break;  // warp-uniform
        ++gid;
    }
}

#ifndef TD_MIN_BLOCKS
#define TD_MIN_BLOCKS 3  // 158 registers, no spills: 4.0 ms vs 4.6 ms at 4 blocks/SM with spills (profiles/r1_notes.md)
#endif
__global__ void __launch_bounds__(TD_THREADS, TD_MIN_BLOCKS) time_domain_kernel(const __grid_constant__ TdParams p) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * TD_WARPS + (threadIdx.x >> 5), nw = gridDim.x * TD_WARPS;
    Mat2 al[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const double* t = p.lane_pow + (s * 32 + lane) * 4;
        al[s] = {t[0], t[1], t[2], t[3]};
    }
    const double inv_k = 1.0 / double(p.g_k), inv_m = 1.0 / double(p.g_m), inv_s = 1.0 / double(p.g_s);

    for (int w = gw; w < p.total_chunks; w += nw) {
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
        double* gk = p.gran_k ? p.gran_k + size_t(trk) * p.pitch_k : nullptr;
        double* gm = p.gran_m ? p.gran_m + size_t(trk) * p.pitch_m : nullptr;
        double* gs = p.gran_s ? p.gran_s + size_t(trk) * p.pitch_s : nullptr;

        const float* __restrict__ L = td.ch0;
        const float* __restrict__ R = td.ch1;
        const bool vec = ((reinterpret_cast<uintptr_t>(L) & 15) == 0) && (!p.stereo || (reinterpret_cast<uintptr_t>(R) & 15) == 0);
        double c10 = 0, c11 = 0, c20 = 0, c21 = 0;  // carries of stage 1 / stage 2
        GranAcc ga_k, ga_m, ga_s;
        double sL = 0, sR = 0, sLL = 0, sRR = 0, sLR = 0, sMM = 0, sSS = 0, sAL = 0, sAR = 0;
        float amax = 0.f;  // max |mono| of this chunk (true-peak screening)

        // raw samples of the lane's segment at step n0 (zero past the end of the track)
        auto load_step = [&](long long n0, float (&l)[TD_SEG], float (&r)[TD_SEG]) {
            const long long nf = n0 + (long long)lane * TD_SEG;
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
        };
        // Software pipeline, two levels deep:
        //   * the raw samples of the next step are in flight while this step is processed (ln / rn);
        //   * the two biquad stages are skewed by one step: iteration `it` runs the shelf on step `it` and the
        //     high-pass on step `it - 1` (y1 = the shelf's rounded output of the previous iteration).  The two
        //     recurrences + warp scans are independent instruction streams in one basic block, so their DFMA
        //     dependency chains interleave.  The first iteration feeds the high-pass zeros (state stays zero), the
        //     last one feeds the shelf zeros.
        float ln[TD_SEG], rn[TD_SEG];
        load_step(ws, ln, rn);
        double y1[TD_SEG];
#pragma unroll
        for (int i = 0; i < TD_SEG; ++i) y1[i] = 0.0;
        for (long long n0 = ws; n0 < ce + TD_STEP; n0 += TD_STEP) {
            const bool have = n0 < ce;                                 // a real step enters stage 1 this iteration
            const long long nprev = n0 - TD_STEP;                      // the step whose stage 2 runs now
            // in the extra last iteration ln / rn hold whatever follows the chunk (or zeros past the track end): the
            // shelf output computed from it is never used and nothing of it is accumulated (`have`)
            float l[TD_SEG], r[TD_SEG];
#pragma unroll
            for (int i = 0; i < TD_SEG; ++i) {
                l[i] = ln[i];
                r[i] = rn[i];
            }
            if (have) load_step(n0 + TD_STEP, ln, rn);
            const bool warm = n0 < cs0;                              // warm-up step: nothing is accumulated
            const bool full = n0 + TD_STEP <= td.n_samples;          // no sample of this step is past the end
            const unsigned n32 = unsigned(n0);
            double x[TD_SEG];
            {
                // float32 partial sums over the lane's 8 samples (zero padding past the end adds exact zeros)
                float mono[TD_SEG];
                float fL = 0.f, fR = 0.f, fLL = 0.f, fRR = 0.f, fLR = 0.f, fMM = 0.f, fSS = 0.f, fAL = 0.f, fAR = 0.f;
#pragma unroll
                for (int i = 0; i < TD_SEG; ++i) {
                    mono[i] = p.stereo ? 0.5f * (l[i] + r[i]) : l[i];
                    x[i] = double(mono[i]);
                }
                if (have && !warm && p.blk_absmax) {
                    float bm = 0.f;
#pragma unroll
                    for (int i = 0; i < TD_SEG; ++i) bm = fmaxf(bm, fabsf(mono[i]));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
                    if (lane == 0) p.blk_absmax[size_t(trk) * p.blk_pitch + size_t(n0 / TD_STEP)] = bm;
                    amax = fmaxf(amax, bm);
                }
                if (have && !warm) {
#pragma unroll
                    for (int i = 0; i < TD_SEG; ++i) {
                        fMM = fmaf(mono[i], mono[i], fMM);
                        fL += l[i];
                        fLL = fmaf(l[i], l[i], fLL);
                        fAL += fabsf(l[i]);
                        if (p.stereo) {
                            fAR += fabsf(r[i]);
                            const float sd = 0.5f * (l[i] - r[i]);
                            fR += r[i];
                            fRR = fmaf(r[i], r[i], fRR);
                            fLR = fmaf(l[i], r[i], fLR);
                            fSS = fmaf(sd, sd, fSS);
                        }
                    }
                    sMM += double(fMM);
                    sL += double(fL);
                    sLL += double(fLL);
                    sAL += double(fAL);
                    if (p.stereo) {
                        sAR += double(fAR);
                        sR += double(fR);
                        sRR += double(fRR);
                        sLR += double(fLR);
                        sSS += double(fSS);
                    }
                    if (gm || gs) {
                        double qm[TD_SEG];
                        const unsigned gid_m = fast_div(n32, unsigned(p.g_m), inv_m), gid_s = fast_div(n32, unsigned(p.g_s), inv_s);
                        const bool fast_m = full && n32 + TD_STEP <= (gid_m + 1) * unsigned(p.g_m);
                        const bool fast_s = full && n32 + TD_STEP <= (gid_s + 1) * unsigned(p.g_s);
                        if (fast_m && fast_s) {   // common case: one float->double conversion serves both
                            const double v = double(fMM);
                            if (gm) {
                                if (gid_m != ga_m.gid) { gran_flush(ga_m, gm, lane); ga_m.gid = gid_m; }
                                ga_m.acc += v;
                            }
                            if (gs) {
                                if (gid_s != ga_s.gid) { gran_flush(ga_s, gs, lane); ga_s.gid = gid_s; }
                                ga_s.acc += v;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < TD_SEG; ++i) qm[i] = double(mono[i] * mono[i]);
                            if (gm) granule_add(ga_m, gm, unsigned(p.g_m), inv_m, n32, qm, full, td.n_samples, lane);
                            if (gs) granule_add(ga_s, gs, unsigned(p.g_s), inv_s, n32, qm, full, td.n_samples, lane);
                        }
                    }
                }
            }
            // ---- K-weighting: shelf on this step || high-pass on the previous step (straight-line, no branches) ----
            double ea0, ea1, eb0, eb1;
            stage_zero_state(p.stage[0], x, ea0, ea1, lane);
            stage_zero_state(p.stage[1], y1, eb0, eb1, lane);
            stage_correct(p.stage[0], al[0], x, ea0, ea1, c10, c11, lane);
            stage_correct(p.stage[1], al[1], y1, eb0, eb1, c20, c21, lane);
            // y1 now holds the K-weighted samples of the previous step; x the shelf output of this one
            if (nprev >= cs0 && gk) {
#pragma unroll
                for (int i = 0; i < TD_SEG; ++i) {
                    const double y = round_f32(y1[i]);
                    y1[i] = y * y;
                }
                granule_add(ga_k, gk, unsigned(p.g_k), inv_k, unsigned(nprev), y1, nprev + TD_STEP <= td.n_samples, td.n_samples, lane);
            }
#pragma unroll
            for (int i = 0; i < TD_SEG; ++i) y1[i] = round_f32(x[i]);  // float32 round trip between the stages
        }
        if (gk) gran_flush(ga_k, gk, lane);
        if (gm) gran_flush(ga_m, gm, lane);
        if (gs) gran_flush(ga_s, gs, lane);
        if (p.blk_absmax && lane == 0 && ce > cs0) atomicMax(&p.absmax_bits[trk], __float_as_uint(amax));
        if (p.moments && ce > cs0) {
            // slot 7 is the sample count (written by the finalize kernel)
            double v[9] = {sL, sR, sLL, sRR, sLR, sMM, sSS, sAL, sAR};
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                v[k] = warp_sum(v[k]);
                if (lane == 0) atomicAdd(&p.moments[size_t(trk) * TA_N_MOMENTS + (k < 7 ? k : k + 1)], v[k]);
            }
        }
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
    if (p.moments && threadIdx.x == 0) p.moments[size_t(trk) * TA_N_MOMENTS + 7] = ns;
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

static StageU make_stage(const Biquad& b, double* lane_pow /* [32][4] */) {
    StageU c{};
    c.b0 = b.b0; c.b1 = b.b1; c.b2 = b.b2; c.na1 = -b.a1; c.na2 = -b.a2;
    const Mat2 A{-b.a1, 1.0, -b.a2, 0.0};
    for (int d = 0; d < 5; ++d) c.P[d] = mat_pow(A, (long long)TD_SEG << d);
    c.W = mat_pow(A, (long long)TD_SEG * 32);
    for (int i = 0; i < TD_SEG; ++i) {
        const Mat2 ai = mat_pow(A, i);
        c.H0[i] = ai.m00;
        c.H1[i] = ai.m01;
    }
    for (int l = 0; l < 32; ++l) {
        const Mat2 m = mat_pow(A, (long long)TD_SEG * l);
        lane_pow[l * 4 + 0] = m.m00; lane_pow[l * 4 + 1] = m.m01; lane_pow[l * 4 + 2] = m.m10; lane_pow[l * 4 + 3] = m.m11;
    }
    return c;
}

void kw_block_bounds(const ta_plan* plan, int64_t n_samples, std::vector<int64_t>& lo, std::vector<int64_t>& hi);
int run_true_peak(const ta_plan* plan, const HostBatch& hb, const Workspace& ws, float* true_peak, int oversample, cudaStream_t stream);

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

// granule sums (three areas) followed by the 2 x 32 x 4 lane powers of the transition matrices
size_t td_granule_doubles(const ta_plan* plan, const HostBatch& hb) {
    int g, pk, pm, ps;
    td_pitches(plan, hb, g, pk, pm, ps);
    return size_t(hb.n_tracks) * (size_t(pk) + pm + ps) + 256;
}

int time_chunk_samples(const ta_plan* plan, int64_t total_samples) {
    // one chunk per warp; aim for >= 4 chunks per resident warp (16 per SM), chunk in [16384, 131072]
    const int64_t target = total_samples / (int64_t(plan->sm_count) * 16 * 4) + 1;
    int64_t cs = ((target + TD_STEP - 1) / TD_STEP) * TD_STEP;
    cs = std::min<int64_t>(std::max<int64_t>(cs, 16384), 131072);
    return int(cs);
}

int run_time_domain(const ta_plan* plan, const HostBatch& hb, const Workspace& ws, const ta_frontend_out* out,
                    cudaStream_t stream) {
    int64_t max_samples = 0;
    for (auto& t : hb.tracks) max_samples = std::max(max_samples, t.n_samples);
    TA_REQUIRE(max_samples < (int64_t(1) << 31) - TD_STEP, "tracks longer than 2^31 samples are not supported");
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
    TA_REQUIRE(p.g_k >= 1 && p.g_m >= 1 && p.g_s >= 1, "granule sizes must be positive");
    const size_t gran = size_t(hb.n_tracks) * (size_t(p.pitch_k) + p.pitch_m + p.pitch_s);
    TA_REQUIRE(gran + 256 <= ws.gran_doubles, "granule workspace too small");
    p.gran_k = want_k ? ws.d_granules : nullptr;
    p.gran_m = want_m ? ws.d_granules + size_t(hb.n_tracks) * p.pitch_k : nullptr;
    p.gran_s = want_s ? ws.d_granules + size_t(hb.n_tracks) * (size_t(p.pitch_k) + p.pitch_m) : nullptr;
    p.moments = out->moments;
    if (out->true_peak) {
        p.blk_absmax = ws.d_blk_absmax;
        p.absmax_bits = ws.d_absmax_bits;
        p.blk_pitch = ws.blk_pitch;
        TA_CUDA(cudaMemsetAsync(ws.d_absmax_bits, 0, sizeof(uint32_t) * hb.n_tracks, stream));
        TA_CUDA(cudaMemsetAsync(out->true_peak, 0, sizeof(float) * hb.n_tracks, stream));
    }
    TA_CUDA(cudaMemsetAsync(ws.d_granules, 0, sizeof(double) * gran, stream));
    if (p.moments) TA_CUDA(cudaMemsetAsync(p.moments, 0, sizeof(double) * TA_N_MOMENTS * hb.n_tracks, stream));

    // warm-up length: the slowest K-weighting pole pair is the high-pass double pole of radius r = sqrt(a2);
    // a zero-input response decays like n*r^n, so take the first multiple of 2048 with n*r^n < 1e-10.
    {
        const double r = std::sqrt(std::fabs(plan->highpass.a2));
        int n = 2048;
        while (n < (1 << 22) && double(n) * std::pow(r, double(n)) > 1e-10) n += 2048;
        p.halo = n;
    }
    double lane_pow[256];
    p.stage[0] = make_stage(plan->shelf, lane_pow);
    p.stage[1] = make_stage(plan->highpass, lane_pow + 128);
    double* d_lane_pow = ws.d_granules + gran;
    TA_CUDA(cudaMemcpyAsync(d_lane_pow, lane_pow, sizeof(lane_pow), cudaMemcpyHostToDevice, stream));
    p.lane_pow = d_lane_pow;

    const int ctas = (hb.total_chunks + TD_WARPS - 1) / TD_WARPS;
    const int grid = std::max(1, std::min(ctas, plan->sm_count * TD_MIN_BLOCKS * 2));
    time_domain_kernel<<<grid, TD_THREADS, 0, stream>>>(p);
    count_launch();
    TA_CUDA(cudaGetLastError());

    if (out->true_peak) {
        int rc = run_true_peak(plan, hb, ws, out->true_peak, out->true_peak_oversample > 0 ? out->true_peak_oversample : 8, stream);
        if (rc != TA_OK) return rc;
    }

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
