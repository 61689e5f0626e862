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
#include <algorithm>

#include "common.cuh"
#include "fft2_core.cuh"

namespace ta {

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
// (hidden: the kernel is HBM-bound), a checkpoint of the running sum every ROLL_BLOCK bins in shared memory, and once
// the total is known a second walk of at most ROLL_BLOCK bins from the last checkpoint below the threshold.
#ifndef CP_ROWS
#define CP_ROWS 8  // magnitude rows whose loads are in flight per thread
#endif
static constexpr int CP_THREADS = 256;
static constexpr int ROLL_BLOCK = 128;
static_assert(ROLL_BLOCK % CP_ROWS == 0, "checkpoints fall on row-group boundaries");

template <bool CHROMA, bool ROLL>
__global__ void __launch_bounds__(CP_THREADS, 3) chroma_project_kernel(const TrackDesc* __restrict__ tracks,
                                                                       const float* __restrict__ mag, const float* __restrict__ fb,
                                                                       float* __restrict__ out, int32_t* __restrict__ rolloff_bin,
                                                                       float roll_percent, int n_bins) {
    using namespace p2;
    extern __shared__ __align__(16) float wsm[];  // CHROMA: [n_bins * 12] filterbank; ROLL: then [n_ck][2 * CP_THREADS] checkpoints
    const TrackDesc td = tracks[blockIdx.y];
    if (blockIdx.x * CP_THREADS * 2 >= td.n_frames) return;
    const int t = (blockIdx.x * CP_THREADS + threadIdx.x) * 2;
    const bool ok = t < td.n_frames;  // rows are padded to a multiple of 32 frames, so t, t+1 stay inside the row
    const float* __restrict__ col = mag + size_t(td.pitch_off) * n_bins + (ok ? t : 0);
    const float* __restrict__ w = fb + size_t(blockIdx.y) * n_bins * 12;
    float2* ck = reinterpret_cast<float2*>(wsm + (CHROMA ? n_bins * 12 : 0)) + threadIdx.x;  // [block * CP_THREADS]
    float2 acc[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) acc[c] = make_float2(0.f, 0.f);
    float2 run = make_float2(0.f, 0.f);   // sequential float32 running sums of the two frames
    auto accumulate = [&](const float2 m, const float* wk) {
        if (ROLL) {
            run.x = __fadd_rn(run.x, m.x);
            run.y = __fadd_rn(run.y, m.y);
        }
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
        if (ROLL && kk % ROLL_BLOCK == 0) ck[(kk / ROLL_BLOCK) * CP_THREADS] = run;  // sum of bins [0, kk)
        float2 m[CP_ROWS];
#pragma unroll
        for (int u = 0; u < CP_ROWS; ++u) m[u] = __ldg(reinterpret_cast<const float2*>(col + size_t(kk + u) * td.ld));
#pragma unroll
        for (int u = 0; u < CP_ROWS; ++u) accumulate(m[u], wsm + (kk + u) * 12);
    }
    for (; kk < n_bins; ++kk) {
        if (ROLL && kk % ROLL_BLOCK == 0) ck[(kk / ROLL_BLOCK) * CP_THREADS] = run;
        accumulate(__ldg(reinterpret_cast<const float2*>(col + size_t(kk) * td.ld)), wsm + kk * 12);
    }
    if (!ok) return;
    if (ROLL) {
        const int n_ck = (n_bins + ROLL_BLOCK - 1) / ROLL_BLOCK;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (t + h >= td.n_frames) break;
            const float thr = __fmul_rn(roll_percent, h ? run.y : run.x);
            int b = 0;   // last checkpoint still below the threshold (running sums of non-negative values never decrease)
            for (int j = 1; j < n_ck; ++j) {
                const float2 c = ck[j * CP_THREADS];
                if ((h ? c.y : c.x) < thr) b = j;
            }
            const float2 c0 = ck[b * CP_THREADS];
            float r = h ? c0.y : c0.x;
            int first = n_bins - 1;
            const float* __restrict__ colh = col + h;
            // second walk, 16 rows in flight (reading a few rows past the answer is harmless; the adds stay sequential)
            for (int k0 = b * ROLL_BLOCK; k0 < n_bins && first == n_bins - 1; k0 += 16) {
                float v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = __ldg(colh + size_t(min(k0 + u, n_bins - 1)) * td.ld);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    r = __fadd_rn(r, v[u]);
                    if (first == n_bins - 1 && k0 + u < n_bins - 1 && !(r < thr)) first = k0 + u;
                }
            }
            rolloff_bin[td.pitch_off + t + h] = first;
        }
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

template <bool CHROMA, bool ROLL>
static int launch_project(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mag, const float* fb,
                          float* chroma, int32_t* rolloff_bin, cudaStream_t stream) {
    const size_t n_ck = (size_t(plan->n_bins) + ROLL_BLOCK - 1) / ROLL_BLOCK;
    const size_t smem = (CHROMA ? size_t(plan->n_bins) * 12 * sizeof(float) : 0) + (ROLL ? n_ck * CP_THREADS * sizeof(float2) : 0);
    auto kern = chroma_project_kernel<CHROMA, ROLL>;
    TA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((hb.max_frames + 2 * CP_THREADS - 1) / (2 * CP_THREADS), hb.n_tracks), CP_THREADS, smem, stream>>>(
        d_tracks, mag, fb, chroma, rolloff_bin, float(plan->desc.roll_percent), plan->n_bins);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

// chroma (may be NULL) and / or the roll-off bins (may be NULL) from an existing magnitude spectrogram
int run_chroma(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mag, const float* frame_max,
               float* chroma, double* tuning, int32_t* rolloff_bin, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    TA_REQUIRE(mag && (reinterpret_cast<uintptr_t>(mag) & 15) == 0, "the magnitude buffer must be present and 16-byte aligned");
    if (!chroma) {
        if (!rolloff_bin) return TA_OK;
        return launch_project<false, true>(plan, hb, d_tracks, mag, nullptr, nullptr, rolloff_bin, stream);
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
    return rolloff_bin ? launch_project<true, true>(plan, hb, d_tracks, mag, fb, chroma, rolloff_bin, stream)
                       : launch_project<true, false>(plan, hb, d_tracks, mag, fb, chroma, nullptr, stream);
}

}  // namespace ta
