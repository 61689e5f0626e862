// K3: onset-strength spectral flux.
//
// Replaces librosa.onset.onset_strength as called from tempo.py:19 (on the dB mel
// spectrogram, float32) and from analysis/structure.py:195 (on linear mel power,
// float64).  For frame t >= pad (pad = lag + 2048 // (2*hop) = 3 at hop 512):
//     env[t] = mean_m max(0, S[m, t-pad+1] - S[m, t-pad]),   env[t < pad] = 0
// with S = max(10*log10(max(1e-10, mel)), 10*log10(max(1e-10, max(mel))) - 80).
// A thread owns four adjacent mel columns (one aligned 128-bit load per band, four bands in flight) and walks
// the mel axis sequentially (the float32 mean accumulates in mel order like numpy); the column to the left of
// its first one comes from the left lane by warp shuffle, so each dB is computed once.  Lane 0 of a warp is
// the halo lane: strips of 128 columns overlap by four, a warp emits 124 new columns.
// HBM-bound: algorithmic bytes 4*M*T read + 12*T written per track.
#include "common.cuh"

namespace ta {

#ifndef TA_ONSET_FAST_LOG
#define TA_ONSET_FAST_LOG 1
#endif
__device__ __forceinline__ float db10(float x) {
#if TA_ONSET_FAST_LOG
    return 3.0102999566398120f * __log2f(fmaxf(1e-10f, x));
#else
    return 10.0f * log10f(fmaxf(1e-10f, x));
#endif
}

static constexpr int ONSET_ROWS = 4;     // mel bands whose loads are in flight per thread
static constexpr int ONSET_STRIP = 124;  // new columns per warp (32 lanes x 4 columns, minus the halo lane)

template <bool ENV, bool FLUX>
__global__ void __launch_bounds__(256) onset_flux_kernel(const TrackDesc* __restrict__ tracks, const float* __restrict__ mel,
                                                         const uint32_t* __restrict__ mel_max, float* __restrict__ env,
                                                         double* __restrict__ flux, int n_mels, int pad) {
    const TrackDesc td = tracks[blockIdx.y];
    const int T = td.n_frames;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);  // strip index within the track
    const int last = T - pad;                           // last mel column that produces an output (frame T-1)
    const size_t obase = size_t(td.pitch_off);
    if (w == 0) {  // the first `pad` frames are zero by definition
        for (int t = lane; t < min(pad, T); t += 32) {
            if (ENV) env[obase + t] = 0.f;
            if (FLUX) flux[obase + t] = 0.0;
        }
    }
    if (w * ONSET_STRIP + 1 > last) return;  // warp-uniform: no column of this strip has an output
    const int c0 = w * ONSET_STRIP + 4 * lane;
    const bool ok = c0 < td.ld;  // ld is a multiple of 32: the whole 128-bit load is inside the row or outside
    const float* __restrict__ base = mel + size_t(td.pitch_off) * n_mels + c0;
    const float floor_db = db10(__uint_as_float(mel_max[blockIdx.y])) - 80.0f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    double accl[4] = {0.0, 0.0, 0.0, 0.0};
    auto band = [&](const float4 v) {
        if (ENV) {
            const float d0 = fmaxf(db10(v.x), floor_db), d1 = fmaxf(db10(v.y), floor_db);
            const float d2 = fmaxf(db10(v.z), floor_db), d3 = fmaxf(db10(v.w), floor_db);
            const float dl = __shfl_up_sync(0xffffffffu, d3, 1);
            acc[0] += fmaxf(0.f, d0 - dl);
            acc[1] += fmaxf(0.f, d1 - d0);
            acc[2] += fmaxf(0.f, d2 - d1);
            acc[3] += fmaxf(0.f, d3 - d2);
        }
        if (FLUX) {
            // max(0, x1 - x0) in float64 of float32 inputs: the sign is decided by the float32 comparison (exact), so the
            // half-wave rectification is a predicated DADD instead of a float64 max (DSETP + two selects per element)
            const float vl = __shfl_up_sync(0xffffffffu, v.w, 1);
            const double x0 = double(v.x), x1 = double(v.y), x2 = double(v.z), x3 = double(v.w);
            if (v.x > vl) accl[0] += x0 - double(vl);
            if (v.y > v.x) accl[1] += x1 - x0;
            if (v.z > v.y) accl[2] += x2 - x1;
            if (v.w > v.z) accl[3] += x3 - x2;
        }
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    int m = 0;
    for (; m + ONSET_ROWS <= n_mels; m += ONSET_ROWS) {
        float4 v[ONSET_ROWS];
#pragma unroll
        for (int r = 0; r < ONSET_ROWS; ++r) v[r] = ok ? __ldg(reinterpret_cast<const float4*>(base + size_t(m + r) * td.ld)) : zero;
#pragma unroll
        for (int r = 0; r < ONSET_ROWS; ++r) band(v[r]);
    }
    for (; m < n_mels; ++m) band(ok ? __ldg(reinterpret_cast<const float4*>(base + size_t(m) * td.ld)) : zero);
    // lane 0 has no left neighbour: its first column belongs to the previous strip, and so do the other three unless
    // this is the first strip of the track
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + i;
        if (c < 1 || c > last || (lane == 0 && (i == 0 || w > 0))) continue;
        const size_t o = obase + c + pad - 1;
        if (ENV) env[o] = acc[i] / float(n_mels);
        if (FLUX) flux[o] = accl[i] / double(n_mels);
    }
}

int run_onset_flux(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mel,
                   const uint32_t* mel_max, float* onset_env, double* flux_linear, cudaStream_t stream) {
    if (!onset_env && !flux_linear) return TA_OK;
    TA_REQUIRE(plan->desc.n_mels > 0, "plan has no mel bands");
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    const int pad = 1 + 2048 / (2 * plan->desc.hop);  // onset_strength's own n_fft default (tempo.py:19, structure.py:195)
    const int strips = (int)((hb.max_frames + ONSET_STRIP - 1) / ONSET_STRIP);
    dim3 grid((strips + 7) / 8, hb.n_tracks);
    const int M = plan->desc.n_mels;
    if (onset_env && flux_linear)
        onset_flux_kernel<true, true><<<grid, 256, 0, stream>>>(d_tracks, mel, mel_max, onset_env, flux_linear, M, pad);
    else if (onset_env)
        onset_flux_kernel<true, false><<<grid, 256, 0, stream>>>(d_tracks, mel, mel_max, onset_env, flux_linear, M, pad);
    else
        onset_flux_kernel<false, true><<<grid, 256, 0, stream>>>(d_tracks, mel, mel_max, onset_env, flux_linear, M, pad);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
