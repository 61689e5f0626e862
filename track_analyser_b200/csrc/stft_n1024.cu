// Instantiations of the fused STFT kernel for n_fft = 1024 (tile of 32 frames).
#include "stft_kernel.cuh"

namespace ta {

int launch_stft_n1024(const ta_plan* plan, const StftParams& p, bool stereo, int sh, cudaStream_t stream) {
    if (sh == 4) return stereo ? launch_stft<1024, 32, true, 4>(plan, p, stream) : launch_stft<1024, 32, false, 4>(plan, p, stream);
    return stereo ? launch_stft<1024, 32, true, 0>(plan, p, stream) : launch_stft<1024, 32, false, 0>(plan, p, stream);
}

}  // namespace ta
