// Host-side dispatch of the fused STFT kernel (stft_kernel.cuh); the kernels are instantiated per n_fft in
// stft_n1024.cu / stft_n2048.cu / stft_n4096.cu so that they compile in parallel.
#include "stft_params.cuh"

namespace ta {

int stft_tile_frames(int n_fft) { return n_fft == 4096 ? 8 : (n_fft == 2048 ? 16 : 32); }
int stft_transform_length(int n_fft) { return n_fft < 1024 ? 1024 : n_fft; }

int run_stft_features(const ta_plan* plan, const HostBatch& hb, const Workspace& ws,
                      const ta_frontend_out* out, cudaStream_t stream) {
    StftParams p{};
    p.tracks = ws.d_tracks;
    p.n_tracks = hb.n_tracks;
    p.total_tiles = hb.total_tiles;
    p.hop = plan->desc.hop;
    p.n_mels = plan->desc.n_mels;
    p.mel_nnz = plan->mel_nnz;
    p.roll_percent = float(plan->desc.roll_percent);
    p.tw1 = plan->d_tw1;
    p.tw2 = plan->d_tw2;
    p.window = plan->d_window;
    p.freqs = plan->d_freqs;
    p.mel_start = plan->d_mel_start;
    p.mel_len = plan->d_mel_len;
    p.mel_woff = plan->d_mel_woff;
    p.mel_w = plan->d_mel_w;
    p.mag = out->magnitude;

    p.mel = (plan->desc.n_mels > 0) ? out->mel : nullptr;
    p.centroid = out->centroid;
    p.rolloff_bin = out->rolloff_bin;
    p.frame_max = out->frame_max;
    p.frame_sum = out->rolloff_bin ? ws.d_frame_sum : nullptr;
    p.ltas = out->ltas;
    p.band_energy = out->band_energy;
    p.mel_max = p.mel ? ws.d_mel_max : nullptr;
    TA_REQUIRE((reinterpret_cast<uintptr_t>(p.mag) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.mel) & 15) == 0,
               "magnitude and mel outputs must be 16-byte aligned");
    const int B = plan->n_bins;
    if (p.ltas) TA_CUDA(cudaMemsetAsync(p.ltas, 0, sizeof(double) * hb.n_tracks * B, stream));
    if (p.band_energy) TA_CUDA(cudaMemsetAsync(p.band_energy, 0, sizeof(double) * hb.n_tracks * 2 * B, stream));
    if (p.mel_max) TA_CUDA(cudaMemsetAsync(p.mel_max, 0, sizeof(uint32_t) * hb.n_tracks, stream));
    const bool stereo = hb.channels == 2;
    const int m = plan->desc.n_fft / 16;
    const int sh = (p.hop % m == 0) ? p.hop / m : 0;
    switch (plan->desc.n_fft) {
        case 512: return launch_stft_small(plan, p, stereo, 2, stream);
        case 256: return launch_stft_small(plan, p, stereo, 4, stream);
        case 2048: return launch_stft_n2048(plan, p, stereo, sh, stream);
        case 1024: return launch_stft_n1024(plan, p, stereo, sh, stream);
        case 4096: return launch_stft_n4096(plan, p, stereo, sh, stream);
    }
    set_error("unsupported n_fft");
    return TA_ERR_UNSUPPORTED;
}

}  // namespace ta
