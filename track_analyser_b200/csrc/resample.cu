// K11: band-limited sinc resampling ("decode + resample").
//
// Replaces resampy.resample(x, sr_orig, sr_new) -- default "kaiser_best" filter -- as the reference calls it once per
// channel from utils._resample (utils.py:55-70; coerce_audio utils.py:91-96,112-116,141-143; load_audio io.py:126-128).
// The algorithm is J. O. Smith's table-driven band-limited interpolation as resampy 0.4 implements it
// (core.resample + interpn.resample_f), restated in oracle/resampy_np.py; every arithmetic step keeps that
// implementation's order and precision so the two agree bit for bit:
//   * t_out[t] = t * (1 / ratio) in float64, n = int(t_out), frac = scale * (t_out - n), scale = min(1, ratio);
//   * left wing over x[n], x[n-1], ..., then right wing over x[n+1], x[n+2], ...; tap weight = win[j] + eta * delta[j]
//     at table index j = offset + i * int(scale * num_table) (float64, separate multiply and add);
//   * the running sum lives in the float32 output element: each tap adds a float64 product to it in float64 and
//     rounds the result back to float32.
// The half window (caller-built, 64 * 512 + 1 float64 values for kaiser_best) is scaled by the ratio when
// down-sampling and stored next to its forward difference as one 16-byte entry, so a tap costs one table load.
// One thread owns one output sample.  The table (512 KB) lives in L2; at any moment the threads of a block read
// the ~index_step entries belonging to one tap index, which L1 holds.
#include <cstdlib>
#include <string>

#include "common.cuh"

struct ta_resampler {
    int device = 0;
    double2* d_tab = nullptr;  // [nwin] {window, forward difference}
    int nwin = 0, num_table = 0, index_step = 0;
    double ratio = 1.0, scale = 1.0, inc = 1.0;
    // rational structure sr_new / sr_orig = lo / li (a whole number of periods, scaled up so that li >= 128):
    // output t = j * lo + p reads around input j * li + const(p).  li == 0: ratio too irregular, generic kernel only.
    int li = 0, lo = 0, tz = 0, halo = 0;
};

namespace ta {

__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ src, long long n_in, long long src_pitch,
                                                       float* __restrict__ dst, long long n_out, long long dst_pitch,
                                                       const double2* __restrict__ tab, int nwin, int num_table, int index_step,
                                                       double scale, double inc) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    const float* __restrict__ x = src + (long long)blockIdx.y * src_pitch;
    const double time_register = __dmul_rn(double(t), inc);
    const long long n = (long long)time_register;
    float acc = 0.f;
    auto tap = [&](const double2 e, double eta, float xv) {
        const double weight = __dadd_rn(e.x, __dmul_rn(eta, e.y));
        acc = __double2float_rn(__dadd_rn(double(acc), __dmul_rn(weight, double(xv))));
    };
    {   // left wing, including the sample at n
        const double frac = __dmul_rn(scale, __dsub_rn(time_register, double(n)));
        const double index_frac = __dmul_rn(frac, double(num_table));
        const int offset = int(index_frac);
        const double eta = __dsub_rn(index_frac, double(offset));
        const long long i_max = min(n + 1, (long long)((nwin - offset) / index_step));
        const double2* e = tab + offset;
        const float* xp = x + n;
        for (long long i = 0; i < i_max; ++i, e += index_step, --xp) tap(__ldg(e), eta, __ldg(xp));
    }
    {   // right wing
        const double frac0 = __dmul_rn(scale, __dsub_rn(time_register, double(n)));
        const double frac = __dsub_rn(scale, frac0);
        const double index_frac = __dmul_rn(frac, double(num_table));
        const int offset = int(index_frac);
        const double eta = __dsub_rn(index_frac, double(offset));
        const long long k_max = min(n_in - n - 1, (long long)((nwin - offset) / index_step));
        const double2* e = tab + offset;
        const float* xp = x + n + 1;
        for (long long k = 0; k < k_max; ++k, e += index_step, ++xp) tap(__ldg(e), eta, __ldg(xp));
    }
    dst[(long long)blockIdx.y * dst_pitch + t] = acc;
}

// Period-major variant for rational ratios.  A CTA owns 32 consecutive periods (lane = period) and stages their input
// samples in shared memory once; a warp takes one output phase p at a time, so its 32 lanes evaluate the SAME filter
// phase for 32 different periods: the table entry is (up to a last-bit difference in the computed time) the same
// address for all lanes -- one broadcast load instead of up to 32 L1 wavefronts -- and the samples come from shared
// memory, skewed by one word per 2^tz (tz = trailing zeros of the period length) so that the lanes' reads, li words
// apart, fall into 32 different banks.  Every thread still derives n, offset and eta from its own output index
// exactly as the generic kernel does, so the results are bit-identical.
__global__ void __launch_bounds__(256) resample_periodic_kernel(const float* __restrict__ src, long long n_in, long long src_pitch,
                                                                float* __restrict__ dst, long long n_out, long long dst_pitch,
                                                                const double2* __restrict__ tab, int nwin, int num_table,
                                                                int index_step, double scale, double inc, int li, int lo, int tz,
                                                                int halo) {
    extern __shared__ float xs[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const long long j0 = (long long)blockIdx.x * 32;          // first period of this CTA
    const long long base = j0 * li - halo;                    // input index of xs[skew(0)]
    const int span = 32 * li + 2 * halo;
    const float* __restrict__ x = src + (long long)blockIdx.y * src_pitch;
    auto skew = [&](int m) { return tz ? m + (m >> tz) : m; };
    for (int m = threadIdx.x; m < span; m += blockDim.x) {
        const long long g = base + m;
        xs[skew(m)] = (g >= 0 && g < n_in) ? __ldg(x + g) : 0.f;
    }
    __syncthreads();
    const long long t_first = (j0 + lane) * lo;
    for (int p = warp; p < lo; p += nwarps) {
        const long long t = t_first + p;
        if (t >= n_out) continue;
        const double time_register = __dmul_rn(double(t), inc);
        const long long n = (long long)time_register;
        float acc = 0.f;
        auto tap = [&](const double2 e, double eta, float xv) {
            const double weight = __dadd_rn(e.x, __dmul_rn(eta, e.y));
            acc = __double2float_rn(__dadd_rn(double(acc), __dmul_rn(weight, double(xv))));
        };
        const double frac0 = __dmul_rn(scale, __dsub_rn(time_register, double(n)));
        {   // left wing, including the sample at n
            const double index_frac = __dmul_rn(frac0, double(num_table));
            const int offset = int(index_frac);
            const double eta = __dsub_rn(index_frac, double(offset));
            const int i_max = int(min(n + 1, (long long)((nwin - offset) / index_step)));
            const double2* e = tab + offset;
            int m = int(n - base);
            for (int i = 0; i < i_max; ++i, e += index_step, --m) tap(__ldg(e), eta, xs[skew(m)]);
        }
        {   // right wing
            const double frac = __dsub_rn(scale, frac0);
            const double index_frac = __dmul_rn(frac, double(num_table));
            const int offset = int(index_frac);
            const double eta = __dsub_rn(index_frac, double(offset));
            const int k_max = int(min(n_in - n - 1, (long long)((nwin - offset) / index_step)));
            const double2* e = tab + offset;
            int m = int(n + 1 - base);
            for (int k = 0; k < k_max; ++k, e += index_step, ++m) tap(__ldg(e), eta, xs[skew(m)]);
        }
        dst[(long long)blockIdx.y * dst_pitch + t] = acc;
    }
}

}  // namespace ta

extern "C" int ta_resampler_create(int device, int sr_orig, int sr_new, const double* half_window, int n_window, int num_table,
                                   ta_resampler** out) {
    using namespace ta;
    TA_REQUIRE(out, "out must not be NULL");
    *out = nullptr;
    TA_REQUIRE(sr_orig > 0 && sr_new > 0, "Invalid sample rate");  // resampy's message
    TA_REQUIRE(half_window && n_window >= 2 && num_table >= 1, "half_window / n_window / num_table are invalid");
    TA_CUDA(cudaSetDevice(device));
    auto* r = new ta_resampler();
    r->device = device;
    r->ratio = double(sr_new) / double(sr_orig);
    r->scale = std::min(1.0, r->ratio);
    r->inc = 1.0 / r->ratio;
    r->nwin = n_window;
    r->num_table = num_table;
    r->index_step = int(r->scale * num_table);
    if (r->index_step < 1) {
        delete r;
        set_error("sample-rate ratio too small for this interpolation table");
        return TA_ERR_INVALID;
    }
    {
        long long a = sr_orig, b = sr_new;
        while (b) { const long long r = a % b; a = b; b = r; }
        long long li = sr_orig / a, lo = sr_new / a;
        while (li < 128) { li *= 2; lo *= 2; }
        // taps reach nwin / index_step samples to either side of n; n itself may sit one below j * li + const
        const int halo = n_window / r->index_step + 4;
        const int tz = __builtin_ctzll((unsigned long long)li);
        const size_t words = size_t(32) * li + 2 * halo;
        if (li <= 2048 && lo <= 8192 && (words + (tz ? (words >> tz) : 0) + 1) * sizeof(float) <= 200 * 1024) {
            r->li = int(li);
            r->lo = int(lo);
            r->tz = tz;
            r->halo = halo;
        }
    }
    std::vector<double2> tab(n_window);
    for (int i = 0; i < n_window; ++i) tab[i].x = r->ratio < 1.0 ? half_window[i] * r->ratio : half_window[i];
    for (int i = 0; i < n_window; ++i) tab[i].y = i + 1 < n_window ? tab[i + 1].x - tab[i].x : 0.0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&r->d_tab), sizeof(double2) * n_window);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_tab, tab.data(), sizeof(double2) * n_window, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(r->d_tab);
        delete r;
        set_error(cudaGetErrorString(e));
        return TA_ERR_CUDA;
    }
    *out = r;
    return TA_OK;
}

extern "C" void ta_resampler_destroy(ta_resampler* r) {
    if (!r) return;
    cudaFree(r->d_tab);
    delete r;
}

extern "C" int64_t ta_resampler_out_len(const ta_resampler* r, int64_t n_in) {
    if (!r || n_in < 0) return -1;
    return int64_t(double(n_in) * r->ratio);  // int(shape * sample_ratio)
}

extern "C" int ta_resample(const ta_resampler* r, const float* src, int64_t n_in, int64_t src_pitch, int n_rows, float* dst,
                           int64_t dst_pitch, void* stream) {
    using namespace ta;
    TA_REQUIRE(r && src && dst, "resampler / src / dst must not be NULL");
    TA_REQUIRE(n_rows >= 1 && n_rows <= 65535, "n_rows must be 1..65535");
    const int64_t n_out = ta_resampler_out_len(r, n_in);
    TA_REQUIRE(n_out >= 1, "Input signal is too small to resample");  // resampy raises ValueError here
    TA_REQUIRE(src_pitch >= n_in && dst_pitch >= n_out, "row pitches are smaller than the rows");
    static const bool generic_only = [] { const char* e = getenv("TA_RESAMPLE"); return e && std::string(e) == "generic"; }();
    if (r->li > 0 && !generic_only) {
        const size_t words = size_t(32) * r->li + 2 * r->halo;
        const size_t smem = (words + (r->tz ? (words >> r->tz) : 0) + 1) * sizeof(float);
        TA_CUDA(cudaFuncSetAttribute(resample_periodic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long periods = (n_out + r->lo - 1) / r->lo;
        dim3 grid((unsigned)((periods + 31) / 32), (unsigned)n_rows);
        resample_periodic_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
            src, n_in, src_pitch, dst, n_out, dst_pitch, r->d_tab, r->nwin, r->num_table, r->index_step, r->scale, r->inc, r->li,
            r->lo, r->tz, r->halo);
    } else {
        dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)n_rows);
        resample_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, n_in, src_pitch, dst, n_out, dst_pitch,
                                                                                   r->d_tab, r->nwin, r->num_table, r->index_step,
                                                                                   r->scale, r->inc);
    }
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}
