// Internal declarations shared by the translation units of libta_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/ta_b200.h"

namespace ta {

void set_error(const std::string& msg);
void count_launch(int n = 1);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define TA_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return ::ta::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define TA_REQUIRE(cond, msg)                    \
    do {                                         \
        if (!(cond)) {                           \
            ::ta::set_error(std::string(msg));   \
            return TA_ERR_INVALID;               \
        }                                        \
    } while (0)

// Device-side per-track descriptor (built from ta_batch into the workspace).
struct TrackDesc {
    const float* ch0;    // L (or mono)
    const float* ch1;    // R (nullptr for mono)
    int64_t n_samples;
    int64_t pitch_off;   // sum of frame pitches of earlier tracks
    int32_t n_frames;    // T
    int32_t ld;          // frame pitch of this track
    int32_t tile_begin;  // first global tile index (K1)
    int32_t chunk_begin; // first global chunk index (K5)
};

struct Biquad {
    double b0, b1, b2, a1, a2;
};

struct CqtTables;  // constant-Q tables of a plan (cqt.cu), built on first use

}  // namespace ta

struct ta_plan {
    ta_plan_desc desc;
    int n_bins;
    int sm_count;
    // device tables
    float2* d_tw1 = nullptr;    // [15 * n_fft/16]
    float2* d_tw2 = nullptr;    // [16 * n_fft/256]
    float* d_window = nullptr;  // [n_fft] periodic Hann (float32 of the float64 window)
    double* d_freqs = nullptr;  // [n_bins]
    int* d_mel_start = nullptr; // [n_mels]
    int* d_mel_len = nullptr;   // [n_mels]
    int* d_mel_woff = nullptr;  // [n_mels]
    float* d_mel_w = nullptr;   // [nnz]
    int mel_nnz = 0;
    double* d_dct = nullptr;    // [TA_N_MFCC * n_mels] orthonormal DCT-II rows (K10)
    float2* d_tg_tw1 = nullptr;   // tempogram transform (N = 1024) twiddles and window
    float2* d_tg_tw2 = nullptr;
    float* d_tg_window = nullptr;
    // true peak (K8): 8 polyphase branches x 21 taps of 8 * firwin(161, 1/8, kaiser 5.0), float32 like scipy
    float tp_coef[8 * 21];
    float tp_gain;   // max over branches of sum |c|: |y| <= tp_gain * max |x| over the 21-sample window
    float tp_floor;  // |y[8q]| >= tp_floor * |x[q]| at the sample of largest magnitude
    // second stream for the time-domain pass of the fused run (it only reads the PCM, so it forks at the
    // start of ta_frontend_run and joins at its end; see plan.cu frontend_impl)
    cudaStream_t aux_stream = nullptr;
    // host copies
    std::vector<float> h_window;
    std::vector<float> h_mel_dense;
    std::vector<double> h_freqs;
    ta::Biquad shelf, highpass;
    // loudness framing
    int kw_block;   // samples per gating block (0.4 s)
    int kw_step;    // gcd-granule of block bounds
    int rms_m_frame, rms_m_hop, rms_s_frame, rms_s_hop;
    // constant-Q transform tables (chroma_cqt), built lazily under the mutex by the first call that needs them
    mutable std::mutex cqt_mutex;
    mutable ta::CqtTables* cqt = nullptr;
};

namespace ta {
// stage launchers (defined in the .cu files)
struct HostBatch {
    int n_tracks = 0, channels = 0;
    std::vector<TrackDesc> tracks;
    int64_t total_pitch = 0;   // P
    int64_t total_samples = 0;
    int total_tiles = 0;
    int total_chunks = 0;
    int max_frames = 0;
};
int build_host_batch(const ta_plan* plan, const ta_batch* b, HostBatch& hb);

// workspace carve-up
struct Workspace {
    TrackDesc* d_tracks;       // [n_tracks]
    uint32_t* d_mel_max;       // [n_tracks]
    double* d_novelty;         // [2 * TA_N_MFCC * P] smoothed cepstrum and unit window means (self-similarity, K13)
    float* d_frame_sum;        // [P] sum_f |X| per frame from K1 (approximate total for the roll-off walk)
    void* d_tmaps;             // [n_tracks] CUtensorMap (128 bytes each) of the magnitude matrices: the projection's TMA loads
    double* d_granules;        // granule sums of the time-domain pass (three areas)
    size_t gran_doubles;
    double* d_fft;             // autocorrelation scratch (complex double) [..]
    size_t fft_elems;
    void* d_chroma;            // chroma_stft scratch (peak lists, filterbanks)
    size_t chroma_bytes;
    float* d_blk_absmax;       // [n_tracks][blk_pitch] max |mono| per 256-sample step (true-peak screening)
    uint32_t* d_absmax_bits;   // [n_tracks] float bits of max |mono|
    int blk_pitch;
    void* d_tp_begin;          // [n_tracks] int64 first step index per track (true-peak kernel)
    unsigned char* end;
};
int stft_tile_frames(int n_fft);
int stft_transform_length(int n_fft);  // 1024 for n_fft 256 / 512 (zero-padded), else n_fft
size_t carve_workspace(const ta_plan* plan, const HostBatch& hb, void* base, Workspace& ws);
}  // namespace ta
