// K8: true peak -- max |y| of the 8x oversampled mono signal (true_peak_dbtp, analysis/loudness.py:81-97:
// scipy.signal.resample_poly(samples, 8, 1) followed by max(abs(.))).
//
// resample_poly(x, 8, 1) is y[8q + ph] = sum_{i=-10..10} c[ph][i] x[q - i] with c[ph][i] = 8 h[80 + ph + 8 i],
// h = firwin(161, 1/8, window=("kaiser", 5.0)) in float32 (plan.cu).  Only the maximum is wanted, and
// |y[8q + ph]| <= G * max_{|i| <= 10} |x[q - i]| with G = max_ph sum_i |c[ph][i]|, while the output at the sample of
// largest magnitude M is at least g0 * M (g0 = |c[0][0]| - sum_{i != 0} |c[0][i]|, ~1: branch 0 of a Nyquist filter is
// an impulse).  So a 256-sample step whose neighbourhood maximum is below (g0 / G) * M cannot contain the true peak
// and is skipped: the result is exact, and the 161 MAC per sample are spent only near the loud passages.  The
// per-step maxima and M come for free from the time-domain pass (timedomain.cu).
// One warp per step, 8 consecutive input samples (64 outputs) per lane.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace ta {

static constexpr int TP_MAX_UP = 32;

struct TpParams {
    const TrackDesc* tracks;
    int n_tracks;
    int stereo;
    int blk_pitch;
    long long total_blocks;       // sum over tracks of ceil(n / 256)
    const long long* blk_begin;   // [n_tracks] first global step index of each track
    const float* blk_absmax;
    const uint32_t* absmax_bits;
    float* true_peak;             // [n_tracks], zero-initialised, updated with atomicMax on the float bits
    float ratio;                  // g0 / G
    int up;                       // oversampling factor (<= TP_MAX_UP)
    float coef[TP_MAX_UP * 21];
};

// UPT: compile-time oversampling factor (8, the reference's default: fully unrolled) or 0 = p.up at run time.
template <int UPT>
__global__ void __launch_bounds__(256) true_peak_kernel(const __grid_constant__ TpParams p) {
    const int up = UPT ? UPT : p.up;
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (long long)gridDim.x * 8;
    for (long long w = gw; w < p.total_blocks; w += nw) {
        int lo = 0, hi = p.n_tracks - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.blk_begin[mid] <= w) lo = mid; else hi = mid - 1;
        }
        const int trk = lo;
        const TrackDesc td = p.tracks[trk];
        const long long nblk = (td.n_samples + 255) / 256;
        const long long b = w - p.blk_begin[trk];
        const float* bm = p.blk_absmax + size_t(trk) * p.blk_pitch;
        float wmax = bm[b];
        if (b > 0) wmax = fmaxf(wmax, bm[b - 1]);
        if (b + 1 < nblk) wmax = fmaxf(wmax, bm[b + 1]);
        const float M = __uint_as_float(p.absmax_bits[trk]);
        if (!(wmax >= p.ratio * M)) continue;  // warp-uniform: this step cannot hold the maximum
        // 8 input samples per lane and the 10 + 10 neighbours, mono mix as utils.py:116 (float32 mean of L and R)
        const long long q0 = b * 256 + (long long)lane * 8;
        float x[28];
#pragma unroll
        for (int j = 0; j < 28; ++j) {
            const long long n = q0 - 10 + j;
            float v = 0.f;
            if (n >= 0 && n < td.n_samples) v = p.stereo ? 0.5f * (__ldg(td.ch0 + n) + __ldg(td.ch1 + n)) : __ldg(td.ch0 + n);
            x[j] = v;
        }
        float best = 0.f;
#pragma unroll
        for (int ph = 0; ph < (UPT ? UPT : up); ++ph) {
            float y[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) y[u] = 0.f;
#pragma unroll
            for (int i = -10; i <= 10; ++i) {
                const float c = p.coef[ph * 21 + (i + 10)];
#pragma unroll
                for (int u = 0; u < 8; ++u) y[u] = fmaf(c, x[10 + u - i], y[u]);  // x[q0 + u - i]
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (q0 + u < td.n_samples) best = fmaxf(best, fabsf(y[u]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(p.true_peak) + trk, __float_as_uint(best));
    }
}

// scipy.signal.resample_poly(x, up, 1): h = firwin(20 up + 1, 1 / up, window=("kaiser", 5.0)) cast to float32 (scipy matches
// the dtype of x), scaled by up; y[up q + ph] = sum_{i=-10..10} c[ph][i + 10] x[q - i].  gain = max_ph sum |c[ph][.]| (a bound on
// |y| / max |x|), floor = |c[0][0]| - sum_{i != 0} |c[0][i]| (|y| at the largest sample is at least floor * |x|).
void true_peak_design(int up, float* coef /* [up * 21] */, float& gain, float& floor_) {
    const double PI = 3.14159265358979323846;
    auto bessel_i0 = [](double x) {
        double s = 1.0, t = 1.0;
        for (int k = 1; k < 200; ++k) {
            t *= (x * 0.5) * (x * 0.5) / (double(k) * double(k));
            s += t;
            if (t < 1e-18 * s) break;
        }
        return s;
    };
    if (up == 1) {  // the reference takes max |x| itself (loudness.py:92-93)
        for (int i = 0; i < 21; ++i) coef[i] = (i == 10) ? 1.f : 0.f;
        gain = 1.0001f;
        floor_ = 0.9999f;
        return;
    }
    const int HL = 10 * up, NT = 2 * HL + 1;
    const double fc = 1.0 / up, beta = 5.0;
    std::vector<double> h(NT);
    double sum = 0.0;
    for (int n = 0; n < NT; ++n) {
        const double m = double(n - HL);
        const double sinc = (m == 0.0) ? 1.0 : std::sin(PI * fc * m) / (PI * fc * m);
        const double r = m / double(HL);
        h[n] = fc * sinc * bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / bessel_i0(beta);
        sum += h[n];
    }
    gain = 0.f;
    for (int ph = 0; ph < up; ++ph) {
        float g = 0.f;
        for (int i = -10; i <= 10; ++i) {
            const int idx = HL + ph + up * i;
            const float c = (idx >= 0 && idx < NT) ? float(h[idx] / sum) * float(up) : 0.f;
            coef[ph * 21 + (i + 10)] = c;
            g += std::fabs(c);
        }
        gain = std::max(gain, g);
    }
    float others = 0.f;
    for (int i = -10; i <= 10; ++i)
        if (i != 0) others += std::fabs(coef[i + 10]);
    gain *= 1.0001f;                                        // margins cover float32 accumulation
    floor_ = (std::fabs(coef[10]) - others) * 0.9999f;
}

int run_true_peak(const ta_plan* plan, const HostBatch& hb, const Workspace& ws, float* true_peak, int oversample,
                  cudaStream_t stream) {
    TA_REQUIRE(oversample >= 1 && oversample <= TP_MAX_UP, "true-peak oversample must lie in 1 .. 32");
    TpParams p{};
    p.tracks = ws.d_tracks;
    p.n_tracks = hb.n_tracks;
    p.stereo = hb.channels == 2;
    p.blk_pitch = ws.blk_pitch;
    p.blk_absmax = ws.d_blk_absmax;
    p.absmax_bits = ws.d_absmax_bits;
    p.true_peak = true_peak;
    p.up = oversample;
    if (oversample == 8) {
        p.ratio = plan->tp_floor / plan->tp_gain;
        for (int i = 0; i < 8 * 21; ++i) p.coef[i] = plan->tp_coef[i];
    } else {
        float gain, floor_;
        true_peak_design(oversample, p.coef, gain, floor_);
        p.ratio = floor_ > 0.f ? floor_ / gain : 0.f;   // 0: no step can be ruled out, every one is evaluated
    }
    // per-track first step index (global step numbering of the grid-stride loop)
    std::vector<long long> begin(hb.n_tracks);
    long long total = 0;
    for (int i = 0; i < hb.n_tracks; ++i) {
        begin[i] = total;
        total += (hb.tracks[i].n_samples + 255) / 256;
    }
    p.total_blocks = total;
    long long* d_begin = reinterpret_cast<long long*>(ws.d_tp_begin);
    TA_CUDA(cudaMemcpyAsync(d_begin, begin.data(), sizeof(long long) * hb.n_tracks, cudaMemcpyHostToDevice, stream));
    p.blk_begin = d_begin;
    if (total == 0) return TA_OK;
    const int grid = int(std::min<long long>((total + 7) / 8, (long long)plan->sm_count * 8));
    if (oversample == 8) true_peak_kernel<8><<<grid, 256, 0, stream>>>(p);
    else true_peak_kernel<0><<<grid, 256, 0, stream>>>(p);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
