// Instantiations of the fused STFT kernel for n_fft = 2048 (tile of 16 frames).
#include "stft_kernel.cuh"

namespace ta {

int launch_stft_n2048(const ta_plan* plan, const StftParams& p, bool stereo, int sh, cudaStream_t stream) {
    if (sh == 4) return stereo ? launch_stft<2048, 16, true, 4>(plan, p, stream) : launch_stft<2048, 16, false, 4>(plan, p, stream);
    return stereo ? launch_stft<2048, 16, true, 0>(plan, p, stream) : launch_stft<2048, 16, false, 0>(plan, p, stream);
}

}  // namespace ta
