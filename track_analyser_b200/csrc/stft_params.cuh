// Parameter block of the fused STFT kernel and the per-n_fft launchers (stft_n*.cu).
#pragma once
#include "common.cuh"

namespace ta {

struct StftParams {
    const TrackDesc* tracks;
    int n_tracks;
    int total_tiles;
    int hop;
    int n_mels;
    int mel_nnz;
    int mel_in_smem;
    float roll_percent;
    const float2* tw1;
    const float2* tw2;
    const float* window;
    const double* freqs;
    const int* mel_start;
    const int* mel_len;
    const int* mel_woff;
    const float* mel_w;
    // outputs (nullable)
    float* mag;
    float* mel;
    double* centroid;
    int32_t* rolloff_bin;
    float* frame_max;     // [P] max_f |X|
    float* frame_sum;     // [P] sum_f |X| (float32 chunk sums combined in double): the roll-off walk's first estimate of its total
    double* ltas;         // [n_tracks][B]
    double* band_energy;  // [n_tracks][2][B]
    uint32_t* mel_max;    // [n_tracks]
};

// sh = hop / (n_fft/16) when that is 1, 2 or 4 (frames of a slot then share loaded samples), else 0
int launch_stft_n1024(const ta_plan* plan, const StftParams& p, bool stereo, int sh, cudaStream_t stream);
int launch_stft_small(const ta_plan* plan, const StftParams& p, bool stereo, int d, cudaStream_t stream);
int launch_stft_n2048(const ta_plan* plan, const StftParams& p, bool stereo, int sh, cudaStream_t stream);
int launch_stft_n4096(const ta_plan* plan, const StftParams& p, bool stereo, int sh, cudaStream_t stream);

}  // namespace ta
