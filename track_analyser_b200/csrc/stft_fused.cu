// K1 + K2(mel) + K7: fused frame + Hann + FFT + |X| kernel with in-CTA epilogue.
//
// Replaces, per track, every mono/mid and side STFT of the reference
// (features.py:79,97,116; stereo.py:95-96; structure.py:48,53; tempo.py:19 via
// melspectrogram) with ONE complex FFT per frame: z = mid + i*side (stereo) or
// z = frame(t) + i*frame(t+1) (mono), split afterwards by Hermitian symmetry.
//
// A persistent CTA (one per SM, 512 threads = NG groups of N/16 threads) walks a
// contiguous range of (track, tile) work items; a tile is TF consecutive frames.
// Per tile:   FFT phase   each group transforms frames g, g+NG, ... and writes
//                         |X_mid| into a shared [bin][frame] tile (pitch TF+1);
//                         per-bin time sums (LTAS, |mid|^2, |side|^2) stay in
//                         registers across tiles of the same track.
//             epilogue    (a) magnitude tile -> global, rows of TF contiguous floats
//                         (b) sparse Slaney mel projection of tile^2 -> global
//                         (c) per-frame centroid / roll-off from the tile
// Algorithmic HBM bytes per tile: read C*TF*hop*4 (PCM), write (B+M)*TF*4.
#include <algorithm>

#include "common.cuh"
#include "fft_core.cuh"

namespace ta {

struct StftParams {
    const TrackDesc* tracks;
    int n_tracks;
    int total_tiles;
    int hop;
    int n_mels;
    int mel_nnz;
    float roll_percent;
    const float2* tw1;
    const float2* tw2;
    const float* window;
    const double* freqs;
    const int* mel_start;
    const int* mel_len;
    const int* mel_woff;
    const float* mel_w;
    // outputs (nullable)
    float* mag;
    float* mel;
    double* centroid;
    int32_t* rolloff_bin;
    float* frame_max;     // [P] max_f |X|
    double* ltas;         // [n_tracks][B]
    double* band_energy;  // [n_tracks][2][B]
    uint32_t* mel_max;    // [n_tracks]
};

// |X| = sqrt(re^2+im^2) through MUFU.SQRT (max rel. error 2^-22): two orders of magnitude below the
// fp32 FFT's own rounding noise and 1e3 below the parity tolerance, at a quarter of __fsqrt_rn's cost.
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void group_barrier(int g, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nthreads) : "memory");
}

template <int N, int TF, bool STEREO>
struct StftSmem {
    using C = FftCfg<N>;
    static constexpr int THREADS = 512;
    static constexpr int NG = THREADS / C::M;
    static constexpr int B = N / 2 + 1;
    static constexpr int TFP = TF + 1;
    static constexpr size_t tile_bytes = ((size_t(B) * TFP * 4 + 15) / 16) * 16;
    static constexpr size_t ex_bytes = size_t(NG) * C::EX * 8;
    static constexpr bool TW1_SMEM = (N != 4096);  // 4096: tile + exchange leave no room, read tw1 through L1/L2
    static constexpr size_t tw1_bytes = TW1_SMEM ? size_t(15) * C::M * 8 : 0;
    static constexpr size_t tw2_bytes = size_t(16) * C::Q * 8;
    static constexpr size_t fixed = tile_bytes + ex_bytes + tw1_bytes + tw2_bytes;
    // + mel tables (runtime size): 3 ints per band, then the packed weights
    static size_t total(int n_mels, int mel_nnz) { return fixed + ((size_t(3) * n_mels * 4 + size_t(mel_nnz) * 4 + 15) / 16) * 16; }
};

__device__ __forceinline__ size_t col_out(const TrackDesc& td, int t0, int f) { return size_t(td.pitch_off) + t0 + f; }

template <int N, int TF, bool STEREO>
__global__ void __launch_bounds__(512, 1) stft_fused_kernel(const StftParams p) {
    using C = FftCfg<N>;
    using S = StftSmem<N, TF, STEREO>;
    constexpr int M = C::M, NG = S::NG, B = S::B, TFP = S::TFP;
    constexpr int NBIN = 8;  // bins k = r + M*i, i < 8 (plus bin N/2 on thread r == 0)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    float2* ex_all = reinterpret_cast<float2*>(smem_raw + S::tile_bytes);
    float2* tw1s = reinterpret_cast<float2*>(smem_raw + S::tile_bytes + S::ex_bytes);
    float2* tw2s = reinterpret_cast<float2*>(smem_raw + S::tile_bytes + S::ex_bytes + S::tw1_bytes);
    int* mel_tab = reinterpret_cast<int*>(smem_raw + S::fixed);  // [3][n_mels]: start, len, woff
    float* mel_ws = reinterpret_cast<float*>(mel_tab + 3 * p.n_mels);

    const int tid = threadIdx.x;
    const int g = tid / M, r = tid % M;
    const int warp = tid >> 5, lane = tid & 31;
    float2* ex = ex_all + size_t(g) * C::EX;

    if (S::TW1_SMEM)
        for (int i = tid; i < 15 * M; i += S::THREADS) tw1s[i] = p.tw1[i];
    const float2* tw1 = S::TW1_SMEM ? tw1s : p.tw1;
    for (int i = tid; i < 16 * C::Q; i += S::THREADS) tw2s[i] = p.tw2[i];
    if (p.mel) {
        for (int i = tid; i < p.n_mels; i += S::THREADS) {
            mel_tab[i] = p.mel_start[i];
            mel_tab[p.n_mels + i] = p.mel_len[i];
            mel_tab[2 * p.n_mels + i] = p.mel_woff[i];
        }
        for (int i = tid; i < p.mel_nnz; i += S::THREADS) mel_ws[i] = p.mel_w[i];
    }
    // window, pre-scaled: 1/2 for the Hermitian split, another 1/2 for (L+-R)/2
    float wreg[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) wreg[n1] = p.window[n1 * M + r] * (STEREO ? 0.25f : 0.5f);
    __syncthreads();

    const int w0 = int((long long)blockIdx.x * p.total_tiles / gridDim.x);
    const int w1 = int((long long)(blockIdx.x + 1) * p.total_tiles / gridDim.x);
    if (w0 >= w1) return;

    // locate the track of the first tile
    int trk = 0;
    {
        int lo = 0, hi = p.n_tracks - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (p.tracks[mid].tile_begin <= w0) lo = mid; else hi = mid - 1;
        }
        trk = lo;
    }

    float acc_l[NBIN + 1], acc_m[NBIN + 1], acc_s[NBIN + 1];
#pragma unroll
    for (int i = 0; i <= NBIN; ++i) acc_l[i] = acc_m[i] = acc_s[i] = 0.f;

    auto flush = [&](int t) {
#pragma unroll
        for (int i = 0; i <= NBIN; ++i) {
            const int k = (i < NBIN) ? r + M * i : N / 2;
            if (i == NBIN && r != 0) continue;
            if (p.ltas) atomicAdd(&p.ltas[size_t(t) * B + k], double(acc_l[i]));
            if (p.band_energy) {
                atomicAdd(&p.band_energy[(size_t(t) * 2 + 0) * B + k], double(acc_m[i]));
                if (STEREO) atomicAdd(&p.band_energy[(size_t(t) * 2 + 1) * B + k], double(acc_s[i]));
            }
            acc_l[i] = acc_m[i] = acc_s[i] = 0.f;
        }
    };

    for (int w = w0; w < w1; ++w) {
        while (trk + 1 < p.n_tracks && w >= p.tracks[trk + 1].tile_begin) {
            flush(trk);
            ++trk;
        }
        const TrackDesc td = p.tracks[trk];
        const int t0 = (w - td.tile_begin) * TF;
        const int nf = min(TF, td.n_frames - t0);
        const int slots = STEREO ? nf : (nf + 1) / 2;

        // ------------------------------ FFT phase ------------------------------
        for (int s = g; s < slots; s += NG) {
            const int f = STEREO ? s : 2 * s;        // frame slot in tile
            const int t = t0 + f;                    // absolute frame
            const long long base = (long long)t * p.hop - N / 2;
            {   // L2 prefetch of the samples this group's next frame adds (its last NG*hop samples)
                const long long nb = base + N + (long long)r * 32;  // one 128-byte line per thread
                if (r * 32 < NG * p.hop * (STEREO ? 1 : 2) && nb + 32 <= td.n_samples) {
                    prefetch_l2(td.ch0 + nb);
                    if (STEREO) prefetch_l2(td.ch1 + nb);
                }
            }
            float2 v[16];
            if (STEREO) {
                const float* __restrict__ L = td.ch0;
                const float* __restrict__ R = td.ch1;
                if (base >= 0 && base + N <= td.n_samples) {
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1) {
                        const float l = __ldg(L + base + n1 * M + r), rr = __ldg(R + base + n1 * M + r);
                        v[n1] = make_float2((l + rr) * wreg[n1], (l - rr) * wreg[n1]);
                    }
                } else {
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1) {
                        const long long n = base + n1 * M + r;
                        const bool ok = n >= 0 && n < td.n_samples;
                        const float l = ok ? __ldg(L + n) : 0.f, rr = ok ? __ldg(R + n) : 0.f;
                        v[n1] = make_float2((l + rr) * wreg[n1], (l - rr) * wreg[n1]);
                    }
                }
            } else {
                const float* __restrict__ X = td.ch0;
                const long long base2 = base + p.hop;
                if (base >= 0 && base2 + N <= td.n_samples) {
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1)
                        v[n1] = make_float2(__ldg(X + base + n1 * M + r) * wreg[n1],
                                            __ldg(X + base2 + n1 * M + r) * wreg[n1]);
                } else {
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1) {
                        const long long na = base + n1 * M + r, nb = na + p.hop;
                        const float a = (na >= 0 && na < td.n_samples) ? __ldg(X + na) : 0.f;
                        const float b = (nb >= 0 && nb < td.n_samples) ? __ldg(X + nb) : 0.f;
                        v[n1] = make_float2(a * wreg[n1], b * wreg[n1]);
                    }
                }
            }
            pass1<N>(v, r, tw1, ex);
            group_barrier(g, M);
            pass2_load<N>(v, r, ex);
            group_barrier(g, M);
            pass2_store<N>(v, r, tw2s, ex);
            group_barrier(g, M);
            pass3_load<N>(v, r, ex);
            group_barrier(g, M);
            pass3_store<N>(v, r, ex);
            group_barrier(g, M);
            const bool second_ok = !STEREO && (f + 1 < nf);
#pragma unroll
            for (int i = 0; i <= NBIN; ++i) {
                if (i == NBIN && r != 0) continue;
                const int k = (i < NBIN) ? r + M * i : N / 2;
                float2 xa, xb;
                split_pair(ex[k], ex[(N - k) & (N - 1)], xa, xb);
                const float pa = fmaf(xa.x, xa.x, xa.y * xa.y);
                const float pb = fmaf(xb.x, xb.x, xb.y * xb.y);
                const float ma = fast_sqrt(pa);
                tile[k * TFP + f] = ma;
                if (STEREO) {
                    acc_l[i] += ma;
                    acc_m[i] += ma * ma;
                    acc_s[i] += pb;
                } else {
                    const float mb = fast_sqrt(pb);
                    if (second_ok) {
                        tile[k * TFP + f + 1] = mb;
                        acc_l[i] += ma + mb;
                        acc_m[i] += ma * ma + mb * mb;
                    } else {
                        acc_l[i] += ma;
                        acc_m[i] += ma * ma;
                    }
                }
            }
            group_barrier(g, M);
        }
        __syncthreads();

        // ------------------------------ epilogue ------------------------------
        constexpr int RPW = 32 / TF;           // tile rows covered by one warp instruction
        const int f = lane % TF, sub = lane / TF;
        const bool fok = f < nf;
        // (a) magnitude rows -> global, TF contiguous floats per row
        if (p.mag) {
            float* dst = p.mag + size_t(td.pitch_off) * B + t0 + f;
            for (int k = warp * RPW + sub; k < B; k += 16 * RPW)
                if (fok) dst[size_t(k) * td.ld] = tile[k * TFP + f];
        }
        // (b) mel projection of tile^2 (power = magnitude**2 as librosa computes it)
        if (p.mel) {
            float* dst = p.mel + size_t(td.pitch_off) * p.n_mels + t0 + f;
            float vmax = 0.f;
            for (int m = warp * RPW + sub; m < p.n_mels; m += 16 * RPW) {
                const int ks = mel_tab[m], len = mel_tab[p.n_mels + m];
                const float* wt = mel_ws + mel_tab[2 * p.n_mels + m];
                const float* col = tile + ks * TFP + f;
                float a0 = 0.f, a1 = 0.f;
                int j = 0;
                for (; j + 4 <= len; j += 4) {
                    const float x0 = col[(j + 0) * TFP], x1 = col[(j + 1) * TFP], x2 = col[(j + 2) * TFP], x3 = col[(j + 3) * TFP];
                    a0 = fmaf(wt[j + 0], x0 * x0, a0);
                    a1 = fmaf(wt[j + 1], x1 * x1, a1);
                    a0 = fmaf(wt[j + 2], x2 * x2, a0);
                    a1 = fmaf(wt[j + 3], x3 * x3, a1);
                }
                for (; j < len; ++j) {
                    const float x0 = col[j * TFP];
                    a0 = fmaf(wt[j], x0 * x0, a0);
                }
                const float acc = a0 + a1;
                if (fok) {
                    dst[size_t(m) * td.ld] = acc;
                    vmax = fmaxf(vmax, acc);
                }
            }
            if (p.mel_max) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
                if (lane == 0) atomicMax(&p.mel_max[trk], __float_as_uint(vmax));
            }
        }
        // (c) per-frame centroid / roll-off; bins split into 16*RPW chunks.  Per chunk: fp32 partial sums
        // s1 = sum |X| and s2 = sum (k-kb)|X| (<= 65 terms each), combined across chunks in double:
        // centroid = df * sum_c (s2_c + kb_c*s1_c) / sum_c s1_c.  (librosa rounds |X|/sum to float32 before
        // the float64 dot product; that changes the result by ~2e-9 relative, far inside rtol 1e-4.)
        if (p.centroid || p.rolloff_bin || p.frame_max) {
            constexpr int NCH = 16 * RPW;
            constexpr int CH = (B + NCH - 1) / NCH;
            float* part_1 = reinterpret_cast<float*>(ex_all);   // [NCH][TF]  (aliases the idle exchange buffers)
            float* part_2 = part_1 + NCH * TF;                  // [NCH][TF]
            float* part_3 = part_2 + NCH * TF;                  // [NCH][TF] chunk maxima
            int* roll_s = reinterpret_cast<int*>(part_3 + NCH * TF);  // [TF]
            const int c = warp * RPW + sub;
            const int kb = c * CH, ke = min(kb + CH, B);
            const float* col = tile + f;
            float s1 = 0.f, s2 = 0.f, s3 = 0.f;
            for (int k = kb; k < ke; ++k) {
                const float a = col[k * TFP];
                s1 += a;
                s2 = fmaf(float(k - kb), a, s2);
                s3 = fmaxf(s3, a);
            }
            part_1[c * TF + f] = s1;
            part_2[c * TF + f] = s2;
            part_3[c * TF + f] = s3;
            if (tid < TF) roll_s[tid] = B;
            __syncthreads();
            float prefix = 0.f, total_f = 0.f;
            for (int cc = 0; cc < NCH; ++cc) {
                const float pf = part_1[cc * TF + f];
                if (cc < c) prefix += pf;
                total_f += pf;
            }
            const float thr = p.roll_percent * total_f;
            int first = B;
            if (c > 0 && !(prefix < thr)) {
                first = kb;  // the crossing happened in an earlier chunk
            } else {
                float run = prefix;
                for (int k = kb; k < ke; ++k) {
                    run += col[k * TFP];
                    if (!(run < thr)) { first = k; break; }
                }
            }
            if (first < B) atomicMin(&roll_s[f], first);
            if (c == 0 && fok && p.centroid) {
                double num = 0.0, den = 0.0;
                for (int cc = 0; cc < NCH; ++cc) {
                    const double a1 = double(part_1[cc * TF + f]);
                    den += a1;
                    num += double(part_2[cc * TF + f]) + double(cc * CH) * a1;
                }
                const double df = p.freqs[1];
                p.centroid[col_out(td, t0, f)] = (den < 1.1754943508222875e-38) ? df * num : df * num / den;
            }
            if (c == 0 && fok && p.frame_max) {
                float mx = 0.f;
                for (int cc = 0; cc < NCH; ++cc) mx = fmaxf(mx, part_3[cc * TF + f]);
                p.frame_max[col_out(td, t0, f)] = mx;
            }
            __syncthreads();
            if (c == 0 && fok && p.rolloff_bin) p.rolloff_bin[col_out(td, t0, f)] = roll_s[f];
        }
        __syncthreads();
    }
    flush(trk);
}

template <int N, int TF, bool STEREO>
static int launch_stft(const ta_plan* plan, const StftParams& p, cudaStream_t stream) {
    using S = StftSmem<N, TF, STEREO>;
    auto kern = stft_fused_kernel<N, TF, STEREO>;
    const size_t smem = S::total(p.n_mels, p.mel_nnz);
    if (smem > 232448) {
        set_error("mel tables do not fit in shared memory next to the STFT tile");
        return TA_ERR_UNSUPPORTED;
    }
    TA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::max(1, std::min(plan->sm_count, p.total_tiles));
    kern<<<grid, 512, smem, stream>>>(p);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

int stft_tile_frames(int n_fft) { return n_fft == 4096 ? 16 : 32; }

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
    p.ltas = out->ltas;
    p.band_energy = out->band_energy;
    p.mel_max = p.mel ? ws.d_mel_max : nullptr;
    const int B = plan->n_bins;
    if (p.ltas) TA_CUDA(cudaMemsetAsync(p.ltas, 0, sizeof(double) * hb.n_tracks * B, stream));
    if (p.band_energy) TA_CUDA(cudaMemsetAsync(p.band_energy, 0, sizeof(double) * hb.n_tracks * 2 * B, stream));
    if (p.mel_max) TA_CUDA(cudaMemsetAsync(p.mel_max, 0, sizeof(uint32_t) * hb.n_tracks, stream));
    const bool stereo = hb.channels == 2;
    switch (plan->desc.n_fft) {
        case 2048:
            return stereo ? launch_stft<2048, 32, true>(plan, p, stream) : launch_stft<2048, 32, false>(plan, p, stream);
        case 1024:
            return stereo ? launch_stft<1024, 32, true>(plan, p, stream) : launch_stft<1024, 32, false>(plan, p, stream);
        case 4096:
            return stereo ? launch_stft<4096, 16, true>(plan, p, stream) : launch_stft<4096, 16, false>(plan, p, stream);
    }
    set_error("unsupported n_fft");
    return TA_ERR_UNSUPPORTED;
}

}  // namespace ta
