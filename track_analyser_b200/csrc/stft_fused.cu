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
    double* ltas;         // [n_tracks][B]
    double* band_energy;  // [n_tracks][2][B]
    uint32_t* mel_max;    // [n_tracks]
};

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
    static constexpr size_t total = tile_bytes + ex_bytes + tw1_bytes + tw2_bytes;
};

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

    const int tid = threadIdx.x;
    const int g = tid / M, r = tid % M;
    const int warp = tid >> 5, lane = tid & 31;
    float2* ex = ex_all + size_t(g) * C::EX;

    if (S::TW1_SMEM)
        for (int i = tid; i < 15 * M; i += S::THREADS) tw1s[i] = p.tw1[i];
    const float2* tw1 = S::TW1_SMEM ? tw1s : p.tw1;
    for (int i = tid; i < 16 * C::Q; i += S::THREADS) tw2s[i] = p.tw2[i];
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
                const float ma = __fsqrt_rn(pa);
                tile[k * TFP + f] = ma;
                if (STEREO) {
                    acc_l[i] += ma;
                    acc_m[i] += ma * ma;
                    acc_s[i] += pb;
                } else {
                    const float mb = __fsqrt_rn(pb);
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
        const size_t col = size_t(td.pitch_off) + t0 + f;  // column inside a packed per-frame series
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
                const int ks = p.mel_start[m], len = p.mel_len[m];
                const float* __restrict__ wt = p.mel_w + p.mel_woff[m];
                float acc = 0.f;
                for (int j = 0; j < len; ++j) {
                    const float a = tile[(ks + j) * TFP + f];
                    acc = fmaf(__ldg(wt + j), a * a, acc);
                }
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
        // (c) per-frame centroid / roll-off; bins split into 16*RPW chunks
        if (p.centroid || p.rolloff_bin) {
            constexpr int NCH = 16 * RPW;
            constexpr int CH = (B + NCH - 1) / NCH;
            // scratch aliases the (idle) exchange buffers
            float* part_f = reinterpret_cast<float*>(ex_all);            // [NCH][TF]
            double* part_d = reinterpret_cast<double*>(part_f + NCH * TF);  // [NCH][TF]
            int* roll_s = reinterpret_cast<int*>(part_d + NCH * TF);        // [TF]
            const int c = warp * RPW + sub;
            const int kb = c * CH, ke = min(kb + CH, B);
            float sf = 0.f;
            double sd = 0.0;
            for (int k = kb; k < ke; ++k) {
                const float a = tile[k * TFP + f];
                sf += a;
                sd += double(a);
            }
            part_f[c * TF + f] = sf;
            part_d[c * TF + f] = sd;
            if (tid < TF) roll_s[tid] = B;
            __syncthreads();
            float prefix = 0.f, total_f = 0.f;
            double total_d = 0.0;
            for (int cc = 0; cc < NCH; ++cc) {
                const float pf = part_f[cc * TF + f];
                if (cc < c) prefix += pf;
                total_f += pf;
                total_d += part_d[cc * TF + f];
            }
            const float thr = p.roll_percent * total_f;
            const double len = (total_d < 1.1754943508222875e-38) ? 1.0 : total_d;
            double cen = 0.0;
            float run = prefix;
            int first = B;
            for (int k = kb; k < ke; ++k) {
                const float a = tile[k * TFP + f];
                run += a;
                if (first == B && !(run < thr)) first = k;
                const float an = float(double(a) / len);
                cen = fma(p.freqs[k], double(an), cen);
            }
            if (first < B) atomicMin(&roll_s[f], first);
            __syncthreads();
            part_d[c * TF + f] = cen;
            __syncthreads();
            if (c == 0 && fok) {
                double tot = 0.0;
                for (int cc = 0; cc < NCH; ++cc) tot += part_d[cc * TF + f];
                if (p.centroid) p.centroid[col] = tot;
                if (p.rolloff_bin) p.rolloff_bin[col] = roll_s[f];
            }
        }
        __syncthreads();
    }
    flush(trk);
}

template <int N, int TF, bool STEREO>
static int launch_stft(const ta_plan* plan, const StftParams& p, cudaStream_t stream) {
    using S = StftSmem<N, TF, STEREO>;
    auto kern = stft_fused_kernel<N, TF, STEREO>;
    static bool configured[16] = {false};
    int dev = plan->desc.device;
    if (dev < 16 && !configured[dev]) {
        TA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total));
        configured[dev] = true;
    }
    const int grid = std::max(1, std::min(plan->sm_count, p.total_tiles));
    kern<<<grid, 512, S::total, stream>>>(p);
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
