// K9: harmonic/percussive separation curves.
//
// Replaces librosa.decompose.hpss(magnitude) at analysis/structure.py:52 for everything the reference does with
// its result: `np.sum(percussive, axis=0)` / `np.sum(harmonic, axis=0)` (structure.py:212-213) and the per-segment
// sums of structure.py:143-144, which are sums of those per-frame values.  hpss() is two 31-wide median filters
// over the (1 + n_fft/2, T) magnitude -- along time for the harmonic reference, along frequency for the
// percussive one, scipy.ndimage "reflect" borders -- followed by soft masks with power 2, margin 1:
//     mask_h = h^2 / (h^2 + p^2) (0.5 where both vanish),  harmonic = S * mask_h,  percussive = S * (1 - mask_h).
// The full-size harmonic/percussive matrices never need to exist: kernel 1 writes the time medians to a scratch
// matrix, kernel 2 walks each frame down the frequency axis with a sliding median, forms the masks and keeps the two
// column sums in registers.
// Both sliding medians keep the window SORTED IN REGISTERS and update it per step with a branch-free delete
// (shift left everything not below the leaving value) and insert (min/max ripple): 4 ALU ops per element.
#include "common.cuh"

namespace ta {

static constexpr int HW = 31;       // median window (librosa kernel_size)
static constexpr int HH = HW / 2;   // 15
static constexpr int HSEG = 128;    // outputs per warp segment in the time-direction kernel

// scipy.ndimage mode="reflect" (d c b a | a b c d | d c b a), valid for any n >= 1
__device__ __forceinline__ int reflect_index(int i, int n) {
    while (i < 0 || i >= n) i = (i < 0) ? -i - 1 : 2 * n - i - 1;
    return i;
}

// sorted window w[0] <= ... <= w[HW-1]: insert x, dropping the largest element
__device__ __forceinline__ void window_insert_drop(float (&w)[HW], float x) {
    float prev = w[0];
    w[0] = fminf(prev, x);
#pragma unroll
    for (int i = 1; i < HW; ++i) {
        const float cur = w[i];
        w[i] = fmaxf(prev, fminf(cur, x));
        prev = cur;
    }
}

// replace one instance of `old` (which is in the window) by x, keeping the window sorted
__device__ __forceinline__ void window_replace(float (&w)[HW], float old, float x) {
    // delete: elements below `old` stay, the rest shift left by one (the first element equal to `old` disappears)
#pragma unroll
    for (int i = 0; i < HW - 1; ++i) w[i] = (w[i] < old) ? w[i] : w[i + 1];
    // insert x into the sorted HW-1 prefix
    float prev = w[0];
    w[0] = fminf(prev, x);
#pragma unroll
    for (int i = 1; i < HW - 1; ++i) {
        const float cur = w[i];
        w[i] = fmaxf(prev, fminf(cur, x));
        prev = cur;
    }
    w[HW - 1] = fmaxf(prev, x);
}

// ---- kernel 1: median along time.  A warp owns 32 adjacent bins (lane = bin) and one segment of HSEG frames.
__global__ void __launch_bounds__(256) hpss_time_median_kernel(const TrackDesc* __restrict__ tracks, const float* __restrict__ mag,
                                                              float* __restrict__ harm, int n_bins) {
    extern __shared__ float smem[];
    constexpr int ROWS = HSEG + 2 * HH;  // 158 staged frames per segment
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* buf = smem + size_t(warp) * ROWS * 33;  // [ROWS][33]
    const TrackDesc td = tracks[blockIdx.z];
    const int T = td.n_frames;
    const int ts = (blockIdx.x * 8 + warp) * HSEG;
    if (ts >= T) return;  // warp-uniform
    const int nsteps = min(HSEG, T - ts);
    const int k0 = blockIdx.y * 32;
    const float* __restrict__ base = mag + size_t(td.pitch_off) * n_bins;
    // stage S[k0 + r][reflect(ts - HH + j)] at buf[j][r]: coalesced along time, conflict-free (pitch 33)
    for (int r = 0; r < 32; ++r) {
        const int k = k0 + r;
        for (int j = lane; j < nsteps + 2 * HH; j += 32) {
            float v = 0.f;
            if (k < n_bins) v = __ldg(base + size_t(k) * td.ld + reflect_index(ts - HH + j, T));
            buf[j * 33 + r] = v;
        }
    }
    __syncwarp();
    float w[HW];
#pragma unroll
    for (int i = 0; i < HW; ++i) w[i] = __int_as_float(0x7f800000);
    for (int j = 0; j < HW; ++j) window_insert_drop(w, buf[j * 33 + lane]);
    for (int s = 0; s < nsteps; ++s) {
        const float old = buf[s * 33 + lane];
        buf[s * 33 + lane] = w[HH];  // slot s is no longer needed as input: it leaves the window now
        if (s + 1 < nsteps) window_replace(w, old, buf[(s + HW) * 33 + lane]);
    }
    __syncwarp();
    float* __restrict__ obase = harm + size_t(td.pitch_off) * n_bins;
    for (int r = 0; r < 32; ++r) {
        const int k = k0 + r;
        if (k >= n_bins) break;
        for (int j = lane; j < nsteps; j += 32) obase[size_t(k) * td.ld + ts + j] = buf[j * 33 + r];
    }
}

// ---- kernel 2: median along frequency + soft masks + column sums.  One thread per frame.
__global__ void __launch_bounds__(128) hpss_freq_median_kernel(const TrackDesc* __restrict__ tracks, const float* __restrict__ mag,
                                                              const float* __restrict__ harm, float* __restrict__ harm_sum,
                                                              float* __restrict__ perc_sum, int n_bins) {
    const TrackDesc td = tracks[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= td.n_frames) return;
    const float* __restrict__ S = mag + size_t(td.pitch_off) * n_bins + t;
    const float* __restrict__ H = harm + size_t(td.pitch_off) * n_bins + t;
    float w[HW];
#pragma unroll
    for (int i = 0; i < HW; ++i) w[i] = __int_as_float(0x7f800000);
    for (int j = -HH; j <= HH; ++j) window_insert_drop(w, __ldg(S + size_t(reflect_index(j, n_bins)) * td.ld));
    double acc_h = 0.0, acc_p = 0.0;
    for (int k = 0; k < n_bins; ++k) {
        const float pr = w[HH];
        const float hr = __ldg(H + size_t(k) * td.ld);
        const float s = __ldg(S + size_t(k) * td.ld);
        // librosa.util.softmask(power=2, split_zeros=True)
        const float z = fmaxf(hr, pr);
        float mh = 0.5f, mp = 0.5f;
        if (!(z < 1.1754943508222875e-38f)) {
            const float a = (hr / z) * (hr / z), b = (pr / z) * (pr / z);
            mh = a / (a + b);
            mp = b / (b + a);
        }
        acc_h += double(s * mh);
        acc_p += double(s * mp);
        if (k + 1 < n_bins)
            window_replace(w, __ldg(S + size_t(reflect_index(k - HH, n_bins)) * td.ld),
                           __ldg(S + size_t(reflect_index(k + HH + 1, n_bins)) * td.ld));
    }
    harm_sum[td.pitch_off + t] = float(acc_h);
    perc_sum[td.pitch_off + t] = float(acc_p);
}

int run_hpss(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mag, float* scratch,
             float* harm_sum, float* perc_sum, cudaStream_t stream) {
    TA_REQUIRE(mag && scratch && harm_sum && perc_sum, "hpss needs magnitude, hpss_scratch and both sum outputs");
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    const int B = plan->n_bins;
    const size_t smem = size_t(8) * (HSEG + 2 * HH) * 33 * sizeof(float);
    TA_CUDA(cudaFuncSetAttribute(hpss_time_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 g1((hb.max_frames + 8 * HSEG - 1) / (8 * HSEG), (B + 31) / 32, hb.n_tracks);
    hpss_time_median_kernel<<<g1, 256, smem, stream>>>(d_tracks, mag, scratch, B);
    count_launch();
    TA_CUDA(cudaGetLastError());
    dim3 g2((hb.max_frames + 127) / 128, hb.n_tracks);
    hpss_freq_median_kernel<<<g2, 128, 0, stream>>>(d_tracks, mag, scratch, harm_sum, perc_sum, B);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
