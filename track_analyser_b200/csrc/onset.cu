// K3: onset-strength spectral flux.
//
// Replaces librosa.onset.onset_strength as called from tempo.py:19 (on the dB mel
// spectrogram, float32) and from analysis/structure.py:195 (on linear mel power,
// float64).  For frame t >= pad (pad = lag + 2048 // (2*hop) = 3 at hop 512):
//     env[t] = mean_m max(0, S[m, t-pad+1] - S[m, t-pad]),   env[t < pad] = 0
// with S = max(10*log10(max(1e-10, mel)), 10*log10(max(1e-10, max(mel))) - 80).
// One thread owns one column j = t - pad + 1 and walks the mel axis sequentially
// (the float32 mean accumulates in mel order like numpy); the value of column
// j-1 comes from the left lane by warp shuffle, so each dB is computed once.
// HBM-bound: algorithmic bytes 4*M*T read + 12*T written per track.
#include "common.cuh"

namespace ta {

__device__ __forceinline__ float db10(float x) { return 10.0f * log10f(fmaxf(1e-10f, x)); }

__global__ void __launch_bounds__(256) onset_flux_kernel(const TrackDesc* __restrict__ tracks, const float* __restrict__ mel,
                                                         const uint32_t* __restrict__ mel_max, float* __restrict__ env,
                                                         double* __restrict__ flux, int n_mels, int pad) {
    const TrackDesc td = tracks[blockIdx.y];
    const int T = td.n_frames;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;  // output frame
    if (blockIdx.x * blockDim.x >= T) return;
    const int lane = threadIdx.x & 31;
    const int j = t - pad + 1;  // "current" column; previous is j-1
    const bool valid = t < T && j >= 1;
    const bool cur_ok = t < T && j >= 0;  // lanes that must still supply a value to the right neighbour
    const float* __restrict__ base = mel + size_t(td.pitch_off) * n_mels;
    const float floor_db = db10(__uint_as_float(mel_max[blockIdx.y])) - 80.0f;
    float acc = 0.f;
    double accl = 0.0;
    for (int m = 0; m < n_mels; ++m) {
        const float* row = base + size_t(m) * td.ld;
        const float lin = cur_ok ? __ldg(row + j) : 0.f;
        const float cur = fmaxf(db10(lin), floor_db);
        float prev = __shfl_up_sync(0xffffffffu, cur, 1);
        float plin = __shfl_up_sync(0xffffffffu, lin, 1);
        if (lane == 0 && valid) {
            plin = __ldg(row + j - 1);
            prev = fmaxf(db10(plin), floor_db);
        }
        acc += fmaxf(0.f, cur - prev);
        accl += fmax(0.0, double(lin) - double(plin));
    }
    if (t < T) {
        const size_t o = size_t(td.pitch_off) + t;
        if (env) env[o] = valid ? acc / float(n_mels) : 0.f;
        if (flux) flux[o] = valid ? accl / double(n_mels) : 0.0;
    }
}

int run_onset_flux(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mel,
                   const uint32_t* mel_max, float* onset_env, double* flux_linear, cudaStream_t stream) {
    if (!onset_env && !flux_linear) return TA_OK;
    TA_REQUIRE(plan->desc.n_mels > 0, "plan has no mel bands");
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    const int pad = 1 + 2048 / (2 * plan->desc.hop);  // onset_strength's own n_fft default (tempo.py:19, structure.py:195)
    dim3 grid((hb.max_frames + 255) / 256, hb.n_tracks);
    onset_flux_kernel<<<grid, 256, 0, stream>>>(d_tracks, mel, mel_max, onset_env, flux_linear, plan->desc.n_mels, pad);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
