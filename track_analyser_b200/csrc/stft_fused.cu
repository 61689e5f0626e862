// Host-side dispatch of the fused STFT kernel (stft_kernel.cuh); the kernels are instantiated per n_fft in
// stft_n1024.cu / stft_n2048.cu / stft_n4096.cu so that they compile in parallel.
#include <mutex>

#include "stft_params.cuh"

namespace ta {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// One 2-d tensor map per track over its (bins, T) float32 magnitude matrix (row pitch ld): box = {4 frames, 256 bins},
// the shape of one store from a sub-tile of K1's shared-memory tile; extents are T and B, so partial tiles are clipped.
static int build_magnitude_maps(const ta_plan* plan, const HostBatch& hb, float* mag, void* d_maps, cudaStream_t stream) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return TA_ERR_CUDA;
    }
    std::vector<CUtensorMap> maps(hb.n_tracks);
    const int B = plan->n_bins;
    for (int i = 0; i < hb.n_tracks; ++i) {
        const TrackDesc& t = hb.tracks[i];
        const cuuint64_t dims[2] = {cuuint64_t(t.n_frames), cuuint64_t(B)};
        const cuuint64_t strides[1] = {cuuint64_t(t.ld) * sizeof(float)};
        const cuuint32_t box[2] = {4, 256};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, mag + size_t(t.pitch_off) * B, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
            return TA_ERR_CUDA;
        }
    }
    TA_CUDA(cudaMemcpyAsync(d_maps, maps.data(), sizeof(CUtensorMap) * hb.n_tracks, cudaMemcpyHostToDevice, stream));
    return TA_OK;
}

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
    p.tmaps = nullptr;
    if (out->magnitude) {
        int rc = build_magnitude_maps(plan, hb, out->magnitude, ws.d_tmaps, stream);
        if (rc != TA_OK) return rc;
        p.tmaps = reinterpret_cast<const CUtensorMap*>(ws.d_tmaps);
    }
    p.mel = (plan->desc.n_mels > 0) ? out->mel : nullptr;
    p.centroid = out->centroid;
    p.rolloff_bin = out->rolloff_bin;
    p.frame_max = out->frame_max;
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
