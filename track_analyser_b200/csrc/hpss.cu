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
// (shift left everything not below the leaving value) and insert (min/max ripple): 4 ops per window slot; the value
// that leaves the window comes from a 32-deep ring in shared memory that each thread keeps for itself.
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

// replace one instance of `old` (which is in the window) by x, keeping the window sorted.
// `one` is 1.0f that the compiler cannot see (a kernel argument): the conditional shift of the delete step is written
// as a predicated multiply by it, which is exact and executes on the FMA pipe; the compare/select/min/max of the
// sorted-window update otherwise all queue on the half-rate ALU pipe (30 FSETP + 30 FSEL + 60 FMNMX per element).
__device__ __forceinline__ void window_replace(float (&w)[HW], float old, float x, float one) {
    // delete: elements below `old` stay, the rest shift left by one (the first element equal to `old` disappears)
#pragma unroll
    for (int i = 0; i < HW - 1; ++i)
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %0, %1;\n\t@!p mul.f32 %0, %2, %3;\n\t}" : "+f"(w[i]) : "f"(old), "f"(w[i + 1]), "f"(one));
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

// ---- kernel 1: median along time.  A warp owns ONE bin and 32 consecutive segments of HSEG frames (lane = segment), so
// the 32 lanes stream through one row of the magnitude matrix, 512 bytes apart: every load is a full 32-byte sector and
// all of a warp's sectors come from the same 16 KB of the row.  A lane reads its segment in aligned chunks of 8 frames
// (two 128-bit loads, the next chunk in flight while the current one is consumed), keeps the last 32 inputs in a
// shared-memory ring (4.2 KB per warp, its own column: no synchronisation at all) to know which value leaves the
// window, and writes 8 medians per two 128-bit stores.  (The first version staged 32 bins x 158 frames per warp in
// shared memory -- 20.8 KB per warp, 8 warps per SM, 11.5 % occupancy: 0.34 ms per track against 0.17 ms for kernel 2.)
static constexpr int HT_WARPS = 4;

__device__ __forceinline__ void load8(const float* __restrict__ row, int f0, int T, float (&v)[8]) {
    if (f0 >= 0 && f0 + 8 <= T) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(row + f0));
        const float4 b = __ldg(reinterpret_cast<const float4*>(row + f0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {  // track borders: scipy.ndimage "reflect"
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(row + reflect_index(f0 + i, T));
    }
}

__global__ void __launch_bounds__(32 * HT_WARPS) hpss_time_median_kernel(const TrackDesc* __restrict__ tracks,
                                                                         const float* __restrict__ mag, float* __restrict__ harm,
                                                                         int n_bins, float one) {
    __shared__ float ring_all[HT_WARPS][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ring = ring_all[warp] + lane;  // slot j at ring[j * 33]
    const TrackDesc td = tracks[blockIdx.z];
    const int T = td.n_frames;
    const int k = blockIdx.y * HT_WARPS + warp;
    const int ts = (blockIdx.x * 32 + lane) * HSEG;
    if (k >= n_bins || ts >= T) return;
    const float* __restrict__ row = mag + size_t(td.pitch_off) * n_bins + size_t(k) * td.ld;
    float* __restrict__ orow = harm + size_t(td.pitch_off) * n_bins + size_t(k) * td.ld;
    float w[HW];
#pragma unroll
    for (int i = 0; i < HW; ++i) w[i] = __int_as_float(0x7f800000);
    // prime with frames ts-15 .. ts+15: four chunks starting at ts-16 (frame ts-16 itself is not part of the window);
    // ring slot of frame f is (f - ts + 16) & 31
    float v[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        load8(row, ts - 16 + 8 * c, T, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ring[(8 * c + i) * 33] = v[i];
            if (8 * c + i >= 1) window_insert_drop(w, v[i]);
        }
    }
    float nxt[8];
    load8(row, ts + 16, T, nxt);
    const int nsteps = min(HSEG, T - ts);
    for (int c = 0; 8 * c < nsteps; ++c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = nxt[i];
        if (8 * (c + 1) < nsteps) load8(row, ts + 16 + 8 * (c + 1), T, nxt);  // in flight during the next 8 steps
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int s = 8 * c + i;  // output frame ts+s; frame ts+s-15 leaves, frame ts+s+16 = v[i] enters
            o[i] = w[HH];
            const float old = ring[((s + 1) & 31) * 33];
            ring[(s & 31) * 33] = v[i];
            window_replace(w, old, v[i], one);
        }
        const int f0 = ts + 8 * c;
        if (f0 + 8 <= T) {
            *reinterpret_cast<float4*>(orow + f0) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(orow + f0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (f0 + i < T) orow[f0 + i] = o[i];
        }
    }
}

// ---- kernel 2: median along frequency + soft masks + column sums.  One thread per frame walks the bins; adjacent
// threads read adjacent frames of a magnitude row, so every access is coalesced.  The last 32 magnitudes of the
// thread's column live in a shared-memory ring (its own bank: no conflicts, no synchronisation), which supplies both
// the value that leaves the window and the centre value, so a step costs two global loads (the entering magnitude and
// the time-direction median) instead of four; the loads of the next four steps are in flight during the current four.
// Soft mask: librosa.util.softmask(power=2, split_zeros=True) with correctly rounded reciprocals instead of four
// divisions (<= 2 ulp from the reference's float32 masks; the outputs are sums of ~1000 masked magnitudes).
static constexpr int HF_THREADS = 128;

// FULL: also store the two component matrices S * mask_h and S * mask_p ((bins, T) like the magnitude): what
// librosa.decompose.hpss returns, for callers that run the reference's own analyse_structure on it (compat/).
template <bool FULL>
__global__ void __launch_bounds__(HF_THREADS) hpss_freq_median_kernel(const TrackDesc* __restrict__ tracks,
                                                                      const float* __restrict__ mag, const float* __restrict__ harm,
                                                                      float* __restrict__ harm_sum, float* __restrict__ perc_sum,
                                                                      float* __restrict__ harm_full, float* __restrict__ perc_full,
                                                                      int n_bins, float one) {
    __shared__ float ring_all[32 * HF_THREADS];
    const TrackDesc td = tracks[blockIdx.y];
    const int t = blockIdx.x * HF_THREADS + threadIdx.x;
    if (t >= td.n_frames) return;
    float* ring = ring_all + threadIdx.x;  // slot j at ring[j * HF_THREADS]: magnitude of bin b lives in slot b & 31
    const size_t ld = td.ld;
    const float* __restrict__ S = mag + size_t(td.pitch_off) * n_bins + t;
    const float* __restrict__ H = harm + size_t(td.pitch_off) * n_bins + t;
    float w[HW];
#pragma unroll
    for (int i = 0; i < HW; ++i) w[i] = __int_as_float(0x7f800000);
    // window of bin 0: bins reflect(-15 .. 15) = 14 .. 0, 0 .. 15 (n_bins >= 32 is guaranteed by the plan's n_fft >= 1024)
    for (int j = 0; j <= HH; ++j) {
        const float v = __ldg(S + size_t(j) * ld);
        ring[(j & 31) * HF_THREADS] = v;
        window_insert_drop(w, v);
        if (j < HH) window_insert_drop(w, v);  // its mirror image below bin 0
    }
    auto entering = [&](int k) {  // bin that enters after step k, mirrored at the top edge
        const int e = k + HH + 1;
        return __ldg(S + size_t(e < n_bins ? e : 2 * n_bins - e - 1) * ld);
    };
    // np.sum(harmonic, axis=0) as numpy evaluates it on the float32 component matrices (structure.py:212-213): the products
    // S * mask rounded to float32, added row after row in float32 -- a sequential chain per frame, which is this thread's walk
    float acc_h = 0.f, acc_p = 0.f;
    float nv[4], hv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        nv[u] = (u < n_bins) ? entering(u) : 0.f;
        hv[u] = (u < n_bins) ? __ldg(H + size_t(u) * ld) : 0.f;
    }
    for (int k0 = 0; k0 < n_bins; k0 += 4) {
        float cv[4], ch[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { cv[u] = nv[u]; ch[u] = hv[u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + 4 + u;
            if (k < n_bins) {
                nv[u] = entering(k);
                hv[u] = __ldg(H + size_t(k) * ld);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + u;
            if (k < n_bins) {
                const float pr = w[HH], hr = ch[u];
                const float sv = ring[(k & 31) * HF_THREADS];
                // librosa.util.softmask(power=2, split_zeros=True)
                const float z = fmaxf(hr, pr);
                float mh = 0.5f, mp = 0.5f;
                if (!(z < 1.1754943508222875e-38f)) {
                    const float rz = __frcp_rn(z);
                    const float hn = hr * rz, pn = pr * rz;
                    const float a = hn * hn, b = pn * pn;
                    const float r = __frcp_rn(a + b);
                    mh = a * r;
                    mp = b * r;
                }
                const float hk = __fmul_rn(sv, mh), pk = __fmul_rn(sv, mp);
                acc_h = __fadd_rn(acc_h, hk);
                acc_p = __fadd_rn(acc_p, pk);
                if (FULL) {
                    const size_t o = size_t(td.pitch_off) * n_bins + size_t(k) * ld + t;
                    harm_full[o] = hk;
                    perc_full[o] = pk;
                }
                // bin k-15 (mirrored below bin 0) leaves, bin k+16 enters
                const int lo = (k >= HH) ? k - HH : HH - 1 - k;
                const float old = ring[(lo & 31) * HF_THREADS];
                const int e = k + HH + 1;
                if (e < n_bins) ring[(e & 31) * HF_THREADS] = cv[u];  // mirrored bins are already in the ring's past; not needed again
                window_replace(w, old, cv[u], one);
            }
        }
    }
    harm_sum[td.pitch_off + t] = acc_h;
    perc_sum[td.pitch_off + t] = acc_p;
}

int run_hpss(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mag, float* scratch,
             float* harm_sum, float* perc_sum, cudaStream_t stream, float* harm_full, float* perc_full) {
    TA_REQUIRE(mag && scratch && harm_sum && perc_sum, "hpss needs magnitude, hpss_scratch and both sum outputs");
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    const int B = plan->n_bins;
    TA_REQUIRE((reinterpret_cast<uintptr_t>(mag) & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0,
               "magnitude and hpss_scratch must be 16-byte aligned");
    dim3 g1((hb.max_frames + 32 * HSEG - 1) / (32 * HSEG), (B + HT_WARPS - 1) / HT_WARPS, hb.n_tracks);
    hpss_time_median_kernel<<<g1, 32 * HT_WARPS, 0, stream>>>(d_tracks, mag, scratch, B, 1.0f);
    count_launch();
    TA_CUDA(cudaGetLastError());
    TA_REQUIRE(B >= 32, "hpss needs at least 32 frequency bins");
    dim3 g2((hb.max_frames + HF_THREADS - 1) / HF_THREADS, hb.n_tracks);
    if (harm_full && perc_full)
        hpss_freq_median_kernel<true><<<g2, HF_THREADS, 0, stream>>>(d_tracks, mag, scratch, harm_sum, perc_sum, harm_full, perc_full, B, 1.0f);
    else
        hpss_freq_median_kernel<false><<<g2, HF_THREADS, 0, stream>>>(d_tracks, mag, scratch, harm_sum, perc_sum, nullptr, nullptr, B, 1.0f);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
