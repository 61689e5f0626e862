// Instantiations of the fused STFT kernel on the 1024-point transform (tile of 32 frames): n_fft = 1024, and n_fft = 512 / 256
// zero-padded into it.
#include "stft_kernel.cuh"

namespace ta {

int launch_stft_n1024(const ta_plan* plan, const StftParams& p, bool stereo, int sh, cudaStream_t stream) {
    if (sh == 4) return stereo ? launch_stft<1024, 32, true, 4>(plan, p, stream) : launch_stft<1024, 32, false, 4>(plan, p, stream);
    return stereo ? launch_stft<1024, 32, true, 0>(plan, p, stream) : launch_stft<1024, 32, false, 0>(plan, p, stream);
}

// n_fft 512 / 256: the frame zero-padded into the 1024-point transform, every 2nd / 4th bin kept
int launch_stft_small(const ta_plan* plan, const StftParams& p, bool stereo, int d, cudaStream_t stream) {
    if (d == 2) return stereo ? launch_stft<1024, 32, true, 0, 2>(plan, p, stream) : launch_stft<1024, 32, false, 0, 2>(plan, p, stream);
    return stereo ? launch_stft<1024, 32, true, 0, 4>(plan, p, stream) : launch_stft<1024, 32, false, 0, 4>(plan, p, stream);
}

}  // namespace ta
