// K10: log-mel cepstrum for the structure stage ("MFCC/self-similarity novelty").
//
// Replaces, on the mel power spectrogram K1 leaves in HBM, the chain at analysis/structure.py:192,199:
//     log_mel = librosa.power_to_db(float64(mel) + 1e-9)      10*log10(max(1e-10, .)), clipped 80 dB below its maximum
//     mfcc    = librosa.feature.mfcc(S=log_mel, n_mfcc=13)    orthonormal DCT-II along the mel axis, first 13 rows
// in float64 like the reference.  The clip level follows from the track's largest mel value (log10 is monotone),
// which K1 already reduced into mel_max.  One thread owns one frame, walks the mel axis once and keeps the 13
// cepstral sums in registers; the 13 x M cosine table is staged in shared memory.  With this output the host no
// longer needs the (M, T) mel matrix: 13 x T float64 go back instead of M x T float32.
#include "common.cuh"

namespace ta {

__global__ void __launch_bounds__(128) mfcc_kernel(const TrackDesc* __restrict__ tracks, const float* __restrict__ mel,
                                                   const uint32_t* __restrict__ mel_max, const double* __restrict__ dct,
                                                   double* __restrict__ mfcc, int n_mels) {
    extern __shared__ double s_dct[];  // [TA_N_MFCC][n_mels]
    for (int i = threadIdx.x; i < TA_N_MFCC * n_mels; i += blockDim.x) s_dct[i] = dct[i];
    __syncthreads();
    const TrackDesc td = tracks[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= td.n_frames) return;
    const double top = 10.0 * log10(fmax(1e-10, double(__uint_as_float(mel_max[blockIdx.y])) + 1e-9));
    const double floor_db = top - 80.0;
    const float* __restrict__ col = mel + size_t(td.pitch_off) * n_mels + t;
    double acc[TA_N_MFCC];
#pragma unroll
    for (int k = 0; k < TA_N_MFCC; ++k) acc[k] = 0.0;
    for (int m = 0; m < n_mels; ++m) {
        const double v = fmax(10.0 * log10(fmax(1e-10, double(__ldg(col + size_t(m) * td.ld)) + 1e-9)), floor_db);
#pragma unroll
        for (int k = 0; k < TA_N_MFCC; ++k) acc[k] = fma(s_dct[k * n_mels + m], v, acc[k]);
    }
    double* __restrict__ o = mfcc + size_t(td.pitch_off) * TA_N_MFCC + t;
#pragma unroll
    for (int k = 0; k < TA_N_MFCC; ++k) o[size_t(k) * td.ld] = acc[k];
}

int run_mfcc(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mel, const uint32_t* mel_max,
             double* mfcc, cudaStream_t stream) {
    TA_REQUIRE(plan->desc.n_mels > 0 && plan->d_dct, "plan has no mel bands");
    TA_REQUIRE(mel && mel_max && mfcc, "mfcc needs the mel buffer");
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    const size_t smem = sizeof(double) * TA_N_MFCC * plan->desc.n_mels;
    TA_REQUIRE(smem <= 48 * 1024, "too many mel bands for the cepstrum table");
    dim3 grid((hb.max_frames + 127) / 128, hb.n_tracks);
    mfcc_kernel<<<grid, 128, smem, stream>>>(d_tracks, mel, mel_max, plan->d_dct, mfcc, plan->desc.n_mels);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
