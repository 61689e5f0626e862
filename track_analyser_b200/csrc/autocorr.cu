// K4: full-length autocorrelation of the onset envelope in float64.
//
// Replaces librosa.autocorrelate (tempo.py:38): irfft(|rfft(x, n_pad)|^2)[:T].
// numpy 1.26 transforms in double, so the reference result is float64; B200 has
// a full-rate FP64 pipe and the work is tiny (one 32768-point transform pair per
// 3-minute track), so this stage computes in double too.  Linear autocorrelation
// does not depend on the pad length, so n_pad is the next power of two >= 2T-1
// instead of scipy's next_fast_len.
//
// n = n1 * n2 (n2 <= 4096) four-step transform, three launches:
//   A  columns : length-n1 DIF FFT over a of x[n2*a + b], times W_n^{b*c}
//   B  rows    : length-n2 DIF FFT, |.|^2, length-n2 inverse DIT FFT, times conj twiddle
//   C  columns : length-n1 inverse DIT FFT, real part / n -> autocorr
// DIF leaves bit-reversed order and the inverse DIT consumes it, so no reordering
// pass exists anywhere.  Twiddles come from sincospi() in double.
#include "common.cuh"

namespace ta {

static constexpr int AC_SMEM_ELEMS = 4096;  // complex doubles per CTA (64 KB)
static constexpr int AC_THREADS = 512;

__host__ __device__ inline int ilog2(unsigned v) {
    int l = 0;
    while ((1u << l) < v) ++l;
    return l;
}

struct AcGeom {
    int n, n1, n2;
};
__host__ __device__ inline AcGeom ac_geom(int T) {
    AcGeom g;
    unsigned need = (T > 0) ? unsigned(2 * T - 1) : 1u;
    g.n = 1 << ilog2(need);
    if (g.n < 2) g.n = 2;
    g.n2 = g.n < AC_SMEM_ELEMS ? g.n : AC_SMEM_ELEMS;
    g.n1 = g.n / g.n2;
    return g;
}

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return bits ? (__brev(v) >> (32 - bits)) : 0u; }

// `batch` independent length-L transforms living in shared memory at s[q*L + i] (stride 1).
// Forward: DIF, natural in -> bit-reversed out.  Inverse: DIT, bit-reversed in -> natural out (unscaled).
__device__ void smem_fft(double2* s, int L, int batch, bool inverse) {
    const int half = L >> 1;
    const int total = batch * half;
    if (!inverse) {
        for (int span = half; span >= 1; span >>= 1) {
            for (int i = threadIdx.x; i < total; i += blockDim.x) {
                const int q = i / half, bi = i - q * half;
                const int pos = bi & (span - 1);
                const int j = q * L + ((bi - pos) << 1) + pos;
                const double2 u = s[j], v = s[j + span];
                double sn, cs;
                sincospi(-double(pos) / double(span), &sn, &cs);
                const double dx = u.x - v.x, dy = u.y - v.y;
                s[j] = make_double2(u.x + v.x, u.y + v.y);
                s[j + span] = make_double2(dx * cs - dy * sn, dx * sn + dy * cs);
            }
            __syncthreads();
        }
    } else {
        for (int span = 1; span <= half; span <<= 1) {
            for (int i = threadIdx.x; i < total; i += blockDim.x) {
                const int q = i / half, bi = i - q * half;
                const int pos = bi & (span - 1);
                const int j = q * L + ((bi - pos) << 1) + pos;
                const double2 u = s[j], w = s[j + span];
                double sn, cs;
                sincospi(double(pos) / double(span), &sn, &cs);
                const double vx = w.x * cs - w.y * sn, vy = w.x * sn + w.y * cs;
                s[j] = make_double2(u.x + vx, u.y + vy);
                s[j + span] = make_double2(u.x - vx, u.y - vy);
            }
            __syncthreads();
        }
    }
}

// Kernel A / C: column transforms.  grid = (ceil(n2 / cols_per_cta), n_tracks).
template <bool INVERSE>
__global__ void __launch_bounds__(AC_THREADS) ac_columns_kernel(const TrackDesc* __restrict__ tracks,
                                                                const float* __restrict__ env, double2* __restrict__ work,
                                                                const size_t* __restrict__ work_off, double* __restrict__ out) {
    extern __shared__ double2 sm[];
    const TrackDesc td = tracks[blockIdx.y];
    const AcGeom g = ac_geom(td.n_frames);
    if (g.n1 == 1) return;
    const int cpb = AC_SMEM_ELEMS / g.n1;  // columns per CTA
    const int b0 = blockIdx.x * cpb;
    if (b0 >= g.n2) return;
    const int bits1 = ilog2(g.n1);
    double2* w = work + work_off[blockIdx.y];
    const int T = td.n_frames;
    if (!INVERSE) {
        const float* x = env + td.pitch_off;
        // smem layout [col][a] so each column is a stride-1 transform
        for (int i = threadIdx.x; i < cpb * g.n1; i += blockDim.x) {
            const int a = i / cpb, c = i - a * cpb;  // consecutive threads -> consecutive b (coalesced)
            const long long idx = (long long)g.n2 * a + b0 + c;
            sm[c * g.n1 + a] = make_double2(idx < T ? double(x[idx]) : 0.0, 0.0);
        }
        __syncthreads();
        smem_fft(sm, g.n1, cpb, false);
        for (int i = threadIdx.x; i < cpb * g.n1; i += blockDim.x) {
            const int pc = i / cpb, c = i - pc * cpb;
            const int b = b0 + c;
            const unsigned cfreq = bitrev(unsigned(pc), bits1);  // true frequency index of this row
            double sn, cs;
            // W_n^{b*c} = exp(-2 pi i b c / n); reduce b*c mod n exactly in integers first
            const long long bc = ((long long)b * cfreq) % g.n;
            sincospi(-2.0 * double(bc) / double(g.n), &sn, &cs);
            const double2 v = sm[c * g.n1 + pc];
            w[(size_t)pc * g.n2 + b] = make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
        }
    } else {
        for (int i = threadIdx.x; i < cpb * g.n1; i += blockDim.x) {
            const int pc = i / cpb, c = i - pc * cpb;
            sm[c * g.n1 + pc] = w[(size_t)pc * g.n2 + b0 + c];
        }
        __syncthreads();
        smem_fft(sm, g.n1, cpb, true);
        double* y = out + td.pitch_off;
        const double scale = 1.0 / double(g.n);
        for (int i = threadIdx.x; i < cpb * g.n1; i += blockDim.x) {
            const int a = i / cpb, c = i - a * cpb;
            const long long idx = (long long)g.n2 * a + b0 + c;
            if (idx < T) y[idx] = sm[c * g.n1 + a].x * scale;
        }
    }
}

// Kernel B: row transforms + power spectrum.  grid = (max n1, n_tracks).
__global__ void __launch_bounds__(AC_THREADS) ac_rows_kernel(const TrackDesc* __restrict__ tracks, const float* __restrict__ env,
                                                             double2* __restrict__ work, const size_t* __restrict__ work_off,
                                                             double* __restrict__ out) {
    extern __shared__ double2 sm[];
    const TrackDesc td = tracks[blockIdx.y];
    const AcGeom g = ac_geom(td.n_frames);
    const int pc = blockIdx.x;
    if (pc >= g.n1) return;
    const int T = td.n_frames;
    double2* w = work + work_off[blockIdx.y] + (size_t)pc * g.n2;
    if (g.n1 == 1) {
        const float* x = env + td.pitch_off;
        for (int i = threadIdx.x; i < g.n2; i += blockDim.x) sm[i] = make_double2(i < T ? double(x[i]) : 0.0, 0.0);
    } else {
        for (int i = threadIdx.x; i < g.n2; i += blockDim.x) sm[i] = w[i];
    }
    __syncthreads();
    smem_fft(sm, g.n2, 1, false);
    for (int i = threadIdx.x; i < g.n2; i += blockDim.x) {
        const double2 v = sm[i];
        sm[i] = make_double2(v.x * v.x + v.y * v.y, 0.0);
    }
    __syncthreads();
    smem_fft(sm, g.n2, 1, true);
    if (g.n1 == 1) {
        double* y = out + td.pitch_off;
        const double scale = 1.0 / double(g.n);
        for (int i = threadIdx.x; i < g.n2; i += blockDim.x)
            if (i < T) y[i] = sm[i].x * scale;
    } else {
        const unsigned cfreq = bitrev(unsigned(pc), ilog2(g.n1));
        for (int b = threadIdx.x; b < g.n2; b += blockDim.x) {
            double sn, cs;
            const long long bc = ((long long)b * cfreq) % g.n;
            sincospi(2.0 * double(bc) / double(g.n), &sn, &cs);
            const double2 v = sm[b];
            w[b] = make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
        }
    }
}

size_t autocorr_scratch_elems(const HostBatch& hb) {
    // per-track complex work area (2 doubles per point) + offset table (size_t per track, kept in doubles)
    size_t elems = 0;
    for (auto& t : hb.tracks) {
        const AcGeom g = ac_geom(t.n_frames);
        if (g.n1 > 1) elems += 2 * size_t(g.n);
    }
    return elems + hb.n_tracks + 2;
}

int run_autocorrelate(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* env, double* out,
                      double* scratch, size_t scratch_elems, cudaStream_t stream) {
    (void)plan;
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    TA_REQUIRE(scratch_elems >= autocorr_scratch_elems(hb), "autocorrelation scratch too small");
    // offsets (in complex elements) of each track's work area, stored after the work areas
    std::vector<size_t> off(hb.n_tracks);
    size_t cur = 0;
    int max_n1 = 1, max_cols_grid = 1;
    for (int i = 0; i < hb.n_tracks; ++i) {
        const AcGeom g = ac_geom(hb.tracks[i].n_frames);
        off[i] = cur;
        if (g.n1 > 1) {
            cur += size_t(g.n);
            max_n1 = std::max(max_n1, g.n1);
            max_cols_grid = std::max(max_cols_grid, g.n2 / (AC_SMEM_ELEMS / g.n1));
        }
    }
    size_t* d_off = reinterpret_cast<size_t*>(scratch + 2 * cur);
    double2* d_work = reinterpret_cast<double2*>(scratch);
    TA_CUDA(cudaMemcpyAsync(d_off, off.data(), sizeof(size_t) * hb.n_tracks, cudaMemcpyHostToDevice, stream));
    const size_t smem = sizeof(double2) * AC_SMEM_ELEMS;
    // the attribute is per device and plans may live on several devices of one process: set it on every call
    TA_CUDA(cudaFuncSetAttribute(ac_columns_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TA_CUDA(cudaFuncSetAttribute(ac_columns_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TA_CUDA(cudaFuncSetAttribute(ac_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (max_n1 > 1) {
        ac_columns_kernel<false><<<dim3(max_cols_grid, hb.n_tracks), AC_THREADS, smem, stream>>>(d_tracks, env, d_work, d_off, out);
        count_launch();
        TA_CUDA(cudaGetLastError());
    }
    ac_rows_kernel<<<dim3(max_n1, hb.n_tracks), AC_THREADS, smem, stream>>>(d_tracks, env, d_work, d_off, out);
        count_launch();
    TA_CUDA(cudaGetLastError());
    if (max_n1 > 1) {
        ac_columns_kernel<true><<<dim3(max_cols_grid, hb.n_tracks), AC_THREADS, smem, stream>>>(d_tracks, env, d_work, d_off, out);
        count_launch();
        TA_CUDA(cudaGetLastError());
    }
    return TA_OK;
}

}  // namespace ta
