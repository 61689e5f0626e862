// Instantiations of the fused STFT kernel for n_fft = 4096 (tile of 8 frames).
#include "stft_kernel.cuh"

namespace ta {

int launch_stft_n4096(const ta_plan* plan, const StftParams& p, bool stereo, int sh, cudaStream_t stream) {
    if (sh == 1) return stereo ? launch_stft<4096, 8, true, 1>(plan, p, stream) : launch_stft<4096, 8, false, 1>(plan, p, stream);
    if (sh == 4) return stereo ? launch_stft<4096, 8, true, 4>(plan, p, stream) : launch_stft<4096, 8, false, 4>(plan, p, stream);
    return stereo ? launch_stft<4096, 8, true, 0>(plan, p, stream) : launch_stft<4096, 8, false, 0>(plan, p, stream);
}

}  // namespace ta
