// K13: MFCC self-similarity novelty for the structure stage ("MFCC/self-similarity novelty").
//
// Replaces, on the (13, T) float64 cepstrum K10 leaves in HBM, the chain at analysis/structure.py:199-210:
//     mfcc = scipy.ndimage.gaussian_filter1d(mfcc, sigma=1.0, axis=1)       9 taps, mode="reflect"
//     for frame in range(context, frames - context):                       context = round(2 s * sr / hop) = 172
//         left  = mean(mfcc[:, frame-context:frame], axis=1);  left  /= norm(left)  + 1e-9
//         right = mean(mfcc[:, frame:frame+context], axis=1);  right /= norm(right) + 1e-9
//         self_similarity[frame] = 1 - dot(left, right)
// in float64 like the reference, with its orders of evaluation: scipy's symmetric correlate1d adds the centre tap first and
// then the tap pairs from the outermost inwards; numpy's mean over 172 contiguous doubles is the pairwise sum (two halves of
// 80 and 92, eight interleaved partial sums each); norm and dot run sequentially over the 13 coefficients.  No fused
// multiply-adds (the x86 wheels of scipy / numpy do not contract).
// Three small kernels (smooth, window means -> unit vectors, dot); the curve is 8 * T bytes instead of the 13 x T matrix
// the host needed before, and the host loses its 13 ms per three-minute track of sliding-window means.
#include <cmath>

#include "common.cuh"

namespace ta {

static constexpr int NV_RADIUS = 4;   // int(4.0 * sigma + 0.5) with sigma = 1
struct NvWeights {
    double w[2 * NV_RADIUS + 1];
};

// scipy.ndimage mode="reflect" (d c b a | a b c d | d c b a), valid for any n >= 1
__device__ __forceinline__ int nv_reflect(int i, int n) {
    while (i < 0 || i >= n) i = (i < 0) ? -i - 1 : 2 * n - i - 1;
    return i;
}

__global__ void __launch_bounds__(128) nv_smooth_kernel(const TrackDesc* __restrict__ tracks, const double* __restrict__ mfcc,
                                                        double* __restrict__ smooth, const NvWeights wt) {
    const TrackDesc td = tracks[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x, T = td.n_frames;
    if (t >= T) return;
    int idx[2 * NV_RADIUS + 1];
#pragma unroll
    for (int d = -NV_RADIUS; d <= NV_RADIUS; ++d) idx[d + NV_RADIUS] = nv_reflect(t + d, T);
    for (int c = 0; c < TA_N_MFCC; ++c) {
        const double* __restrict__ row = mfcc + (size_t(td.pitch_off) * TA_N_MFCC + size_t(c) * td.ld);
        double tmp = __dmul_rn(row[t], wt.w[NV_RADIUS]);
#pragma unroll
        for (int d = NV_RADIUS; d >= 1; --d)
            tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(row[idx[NV_RADIUS - d]], row[idx[NV_RADIUS + d]]), wt.w[NV_RADIUS - d]));
        smooth[size_t(td.pitch_off) * TA_N_MFCC + size_t(c) * td.ld + t] = tmp;
    }
}

// numpy's pairwise sum of n contiguous doubles (numpy/core/src/umath/loops_utils.h.src: pairwise_sum)
__device__ double nv_pairwise(const double* __restrict__ a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(nv_pairwise(a, n2), nv_pairwise(a + n2, n - n2));
}

// unit[c][j] = mean(smooth[c][j : j + context]) / (norm + 1e-9) for window starts j = 0 .. T - context
__global__ void __launch_bounds__(128) nv_unit_kernel(const TrackDesc* __restrict__ tracks, const double* __restrict__ smooth,
                                                      double* __restrict__ unit, int context) {
    const TrackDesc td = tracks[blockIdx.y];
    const int j = blockIdx.x * blockDim.x + threadIdx.x, T = td.n_frames;
    if (T <= 2 * context || j > T - context) return;
    double m[TA_N_MFCC], ss = 0.0;
    for (int c = 0; c < TA_N_MFCC; ++c) {
        m[c] = nv_pairwise(smooth + size_t(td.pitch_off) * TA_N_MFCC + size_t(c) * td.ld + j, context) / double(context);
        ss = __dadd_rn(ss, __dmul_rn(m[c], m[c]));
    }
    const double den = __dadd_rn(sqrt(ss), 1e-9);
    for (int c = 0; c < TA_N_MFCC; ++c) unit[size_t(td.pitch_off) * TA_N_MFCC + size_t(c) * td.ld + j] = m[c] / den;
}

__global__ void __launch_bounds__(128) nv_dot_kernel(const TrackDesc* __restrict__ tracks, const double* __restrict__ unit,
                                                     double* __restrict__ out, int context) {
    const TrackDesc td = tracks[blockIdx.y];
    const int f = blockIdx.x * blockDim.x + threadIdx.x, T = td.n_frames;
    if (f >= T) return;
    double v = 0.0;
    if (T > 2 * context && f >= context && f < T - context) {
        double s = 0.0;
        for (int c = 0; c < TA_N_MFCC; ++c) {
            const double* __restrict__ row = unit + size_t(td.pitch_off) * TA_N_MFCC + size_t(c) * td.ld;
            s = __dadd_rn(s, __dmul_rn(row[f - context], row[f]));
        }
        v = __dadd_rn(1.0, -s);
    }
    out[td.pitch_off + f] = v;
}

// scratch: 2 * TA_N_MFCC * P doubles
int run_self_similarity(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const double* mfcc, double* scratch,
                        double* out, cudaStream_t stream) {
    TA_REQUIRE(mfcc && scratch && out, "self-similarity needs the mfcc buffer");
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    // scipy.ndimage._gaussian_kernel1d(sigma=1, order=0, radius=4): exp(-0.5 x^2), normalised by its sum (numpy's sum of nine
    // doubles is sequential)
    NvWeights wt;
    double sum = 0.0;
    for (int i = 0; i <= 2 * NV_RADIUS; ++i) {
        const double x = double(i - NV_RADIUS);
        wt.w[i] = std::exp(-0.5 / (1.0 * 1.0) * x * x);
        sum += wt.w[i];
    }
    for (double& w : wt.w) w /= sum;
    // context = max(2, int(round(2.0 * sr / hop))) with Python's round-half-even
    const int context = std::max(2, int(std::nearbyint(2.0 * double(plan->desc.sample_rate) / double(plan->desc.hop))));
    double* smooth = scratch;
    double* unit = scratch + size_t(TA_N_MFCC) * size_t(hb.total_pitch);
    dim3 grid((hb.max_frames + 127) / 128, hb.n_tracks);
    nv_smooth_kernel<<<grid, 128, 0, stream>>>(d_tracks, mfcc, smooth, wt);
    nv_unit_kernel<<<grid, 128, 0, stream>>>(d_tracks, smooth, unit, context);
    nv_dot_kernel<<<grid, 128, 0, stream>>>(d_tracks, unit, out, context);
    count_launch(3);
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
