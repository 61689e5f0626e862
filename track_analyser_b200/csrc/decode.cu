// K0: PCM decode -- interleaved samples as WAV files hold them -> planar float32 (channels, n_frames), the layout
// io.load_audio returns (io.py:72-79: soundfile.read(dtype="float32", always_2d=True).T).  libsndfile's conversions:
// int16 / 32768, 24-bit packed little endian / 8388608, int32 / 2147483648 (int -> float rounds to nearest, the scale
// is a power of two), float32 copied.  One thread per frame; reads and writes are coalesced; HBM-bound.
#include "common.cuh"

namespace ta {

template <int FMT>
__device__ __forceinline__ float pcm_sample(const unsigned char* __restrict__ src, size_t idx) {
    if (FMT == TA_PCM_S16) return float(reinterpret_cast<const int16_t*>(src)[idx]) * (1.0f / 32768.0f);
    if (FMT == TA_PCM_S24) {
        const unsigned char* b = src + idx * 3;
        int v = int(b[0]) | (int(b[1]) << 8) | (int(b[2]) << 16);
        v = (v << 8) >> 8;  // sign-extend 24 -> 32 bits
        return float(v) * (1.0f / 8388608.0f);
    }
    if (FMT == TA_PCM_S32) return float(reinterpret_cast<const int32_t*>(src)[idx]) * (1.0f / 2147483648.0f);
    return reinterpret_cast<const float*>(src)[idx];
}

template <int FMT>
__global__ void __launch_bounds__(256) decode_pcm_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst,
                                                        int channels, long long n_frames) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += stride)
        for (int c = 0; c < channels; ++c) dst[(size_t)c * n_frames + f] = pcm_sample<FMT>(src, size_t(f) * channels + c);
}

// Fingerprint of the float32 mono mix (L[i] + R[i]) * 0.5f of a planar pair: fp[0] = sum of the bit patterns, fp[1] = sum of
// bit pattern * ((i & 0xffff) + 1), both modulo 2^64.  The host forms the same two sums over AudioInput.samples; equal
// fingerprints mean the mono samples are the mean of the stereo pair (utils.py:116) without uploading them.
__global__ void __launch_bounds__(256) mono_mix_fingerprint_kernel(const float* __restrict__ left, const float* __restrict__ right,
                                                                  long long n, unsigned long long* __restrict__ fp) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long s1 = 0, s2 = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long b = __float_as_uint(__fmul_rn(__fadd_rn(left[i], right[i]), 0.5f));
        s1 += b;
        s2 += b * (unsigned long long)((i & 0xffff) + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(fp, s1);
        atomicAdd(fp + 1, s2);
    }
}

}  // namespace ta

extern "C" int ta_mono_mix_fingerprint(const float* planar_stereo, int64_t n_samples, uint64_t* fingerprint, void* stream) {
    using namespace ta;
    TA_REQUIRE(planar_stereo && fingerprint, "pointers must not be NULL");
    TA_REQUIRE(n_samples >= 0, "n_samples must be >= 0");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    TA_CUDA(cudaMemsetAsync(fingerprint, 0, 2 * sizeof(uint64_t), st));
    if (n_samples == 0) return TA_OK;
    const int grid = int(std::min<long long>((n_samples + 255) / 256, 148 * 8));
    mono_mix_fingerprint_kernel<<<grid, 256, 0, st>>>(planar_stereo, planar_stereo + n_samples, n_samples,
                                                      reinterpret_cast<unsigned long long*>(fingerprint));
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

extern "C" int ta_decode_pcm(const void* interleaved, int format, int channels, int64_t n_frames, float* planar_out,
                             void* stream) {
    using namespace ta;
    TA_REQUIRE(interleaved && planar_out, "interleaved / planar_out must not be NULL");
    TA_REQUIRE(channels >= 1 && channels <= 8, "channels must be 1..8");
    TA_REQUIRE(n_frames >= 0, "n_frames must be >= 0");
    if (n_frames == 0) return TA_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(interleaved);
    const int grid = int(std::min<long long>((n_frames + 255) / 256, 148 * 16));
    switch (format) {
        case TA_PCM_S16: decode_pcm_kernel<TA_PCM_S16><<<grid, 256, 0, st>>>(src, planar_out, channels, n_frames); break;
        case TA_PCM_S24: decode_pcm_kernel<TA_PCM_S24><<<grid, 256, 0, st>>>(src, planar_out, channels, n_frames); break;
        case TA_PCM_S32: decode_pcm_kernel<TA_PCM_S32><<<grid, 256, 0, st>>>(src, planar_out, channels, n_frames); break;
        case TA_PCM_F32: decode_pcm_kernel<TA_PCM_F32><<<grid, 256, 0, st>>>(src, planar_out, channels, n_frames); break;
        default: set_error("unknown PCM format"); return TA_ERR_INVALID;
    }
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}
