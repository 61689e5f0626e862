// K1 + K2(mel) + K7: fused frame + Hann + FFT + |X| kernel with in-CTA epilogue.
//
// Replaces, per track, every mono/mid and side STFT of the reference
// (features.py:79,97,116; stereo.py:95-96; structure.py:48,53; tempo.py:19 via
// melspectrogram; harmony.py:254) with ONE complex FFT per frame (stereo: z = mid + i*side) or
// per two frames (mono: z = frame_a + i*frame_b), split afterwards by Hermitian symmetry -- and
// every thread carries TWO such transforms in packed f32x2 registers (fft2_core.cuh), so one
// "slot" of work is 2 stereo frames or 4 mono frames.
//
// A persistent CTA (one per SM, 512 threads = NG groups of N/16 threads) walks a contiguous range
// of (track, tile) work items; a tile is TF consecutive frames.
// Per tile:   FFT phase   each group transforms slots g, g+NG, ... and writes |X_mid| pairs into a
//                         shared [bin][frame] tile (pitch TF+2 floats; (TF+2)/2 odd keeps the 64-bit
//                         column-pair stores and the row reads conflict free); per-bin sums of
//                         |side|^2 stay in registers across the tiles of a track.
//             epilogue    (a) magnitude tile -> global, row segments of TF contiguous floats;
//                             per-bin time sums (LTAS, |mid|^2), one thread per bin
//                         (b) sparse Slaney mel projection of tile^2 -> global (two frames per thread)
//                         (c) per-frame centroid / max / sum from partial sums taken during (a)'s walk
//                             (the roll-off bin is an exact sequential float32 chain: chroma.cu walks it)
// Algorithmic HBM bytes per tile: read C*TF*hop*4 (PCM), write (B+M)*TF*4 + 16*TF.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "fft2_core.cuh"
#include "stft_params.cuh"

namespace ta {


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

// D > 1: the plan's n_fft is N / D (256 or 512 on the 1024-point transform): the frame fills the first n_fft rows of the
// transform (the window table is zero beyond them) and its spectrum is every D-th bin of the zero-padded transform.
template <int N, int TF, bool STEREO, int D = 1>
struct StftCfg {
    using C = FftCfg<N>;
    static constexpr int THREADS = 512;
    static constexpr int NG = THREADS / C::M;     // transform groups per CTA
    static constexpr int B = N / (2 * D) + 1;     // bins of the plan's n_fft
    static constexpr int FPS = STEREO ? 2 : 4;    // frames per slot (two packed transforms)
    static constexpr int SLOTS = TF / FPS;
    static constexpr int TFP = TF + 2;            // tile row pitch (floats)
    static constexpr int HP = TF / 2;             // frame pairs per tile
    static constexpr int NACC = (B + THREADS - 1) / THREADS;
    static constexpr int CH = (B + 31) / 32;      // bins per lane in the per-frame feature pass
    static_assert(SLOTS % NG == 0, "tile must hold a whole number of rounds");
    static_assert((TFP / 2) % 2 == 1, "tile pitch / 2 must be odd");
    static constexpr size_t tile_bytes = ((size_t(B) * TFP * 4 + 15) / 16) * 16;
    static constexpr size_t ex_bytes = size_t(NG) * p2::Ex<N>::SLOTS * 16;
    static constexpr bool TW1_SMEM = (N != 4096);  // 4096: tile + exchange leave no room, read tw1 through L2
    static constexpr size_t tw1_bytes = TW1_SMEM ? size_t(15) * C::M * 8 : 0;
    static constexpr size_t tw2_bytes = size_t(16) * C::Q * 8;
    static constexpr size_t fixed = tile_bytes + ex_bytes + tw1_bytes + tw2_bytes;
    __host__ __device__ static size_t mel_tab_bytes(int n_mels) { return ((size_t(3) * n_mels * 4 + 15) / 16) * 16; }
    __host__ __device__ static size_t mel_w_bytes(int mel_nnz) { return ((size_t(mel_nnz) * 4 + 15) / 16) * 16; }
};

static constexpr size_t SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA on sm_100

__device__ __forceinline__ size_t col_out(const TrackDesc& td, int t0, int f) { return size_t(td.pitch_off) + t0 + f; }

// mel band m, frame pair fp: sum_j w[j] * tile[ks+j][2fp..2fp+1]^2
template <int TFP>
__device__ __forceinline__ float2 mel_taps(const float* __restrict__ col, const float* __restrict__ wt, int len) {
    using namespace p2;
    float2 a0 = make_float2(0.f, 0.f), a1 = a0;
    int j = 0;
    for (; j + 2 <= len; j += 2) {
        const float2 x0 = *reinterpret_cast<const float2*>(col + (j + 0) * TFP);
        const float2 x1 = *reinterpret_cast<const float2*>(col + (j + 1) * TFP);
        a0 = pfmas(pmul(x0, x0), wt[j + 0], a0);
        a1 = pfmas(pmul(x1, x1), wt[j + 1], a1);
    }
    if (j < len) {
        const float2 x0 = *reinterpret_cast<const float2*>(col + j * TFP);
        a0 = pfmas(pmul(x0, x0), wt[j], a0);
    }
    return padd(a0, a1);
}

// SH > 0: hop == SH * (N/16), so frame q of a slot reads the samples of frame 0 shifted by q*SH rows of the
// (16, N/16) sample matrix and each thread loads 16 + (frames-1)*SH values per channel instead of 16 per frame.
template <int N, int TF, bool STEREO, int SH, int D = 1>
__global__ void __launch_bounds__(512, 1) stft_fused_kernel(const StftParams p) {
    using namespace p2;
    using C = FftCfg<N>;
    using S = StftCfg<N, TF, STEREO, D>;
    static_assert(D == 1 || (SH == 0 && C::NB >= 2), "zero-padded transforms use the generic load path of the paired core");
    using E = Ex<N>;
    constexpr int M = C::M, NG = S::NG, B = S::B, TFP = S::TFP, HP = S::HP, NACC = S::NACC, CH = S::CH;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    float4* ex_all = reinterpret_cast<float4*>(smem_raw + S::tile_bytes);
    float2* tw1s = reinterpret_cast<float2*>(smem_raw + S::tile_bytes + S::ex_bytes);
    float2* tw2s = reinterpret_cast<float2*>(smem_raw + S::tile_bytes + S::ex_bytes + S::tw1_bytes);
    int* mel_tab = reinterpret_cast<int*>(smem_raw + S::fixed);  // [3][n_mels]: start, len, woff
    float* mel_ws = reinterpret_cast<float*>(smem_raw + S::fixed + S::mel_tab_bytes(p.n_mels));

    const int tid = threadIdx.x;
    const int g = tid / M, r = tid % M;
    const int warp = tid >> 5, lane = tid & 31;
    float4* ex = ex_all + size_t(g) * E::SLOTS;

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
        if (p.mel_in_smem)
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

    int trk = 0;
    {
        int lo = 0, hi = p.n_tracks - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (p.tracks[mid].tile_begin <= w0) lo = mid; else hi = mid - 1;
        }
        trk = lo;
    }

    float acc_s[9];            // sum_t |side[k]|^2 for the 8 kept bins (+ bin N/2 on r == 0)
    float acc_l[NACC], acc_m[NACC];  // sum_t |mid[k]|, |mid[k]|^2 for bins tid + 512*j
#pragma unroll
    for (int i = 0; i < 9; ++i) acc_s[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc_l[j] = acc_m[j] = 0.f;

    auto flush = [&](int t) {
        if (STEREO && p.band_energy) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int k;
                if constexpr (C::NB >= 2) k = Pair3<N>::owned_bin(r, i);
                else k = kept_bin<N>(r, i);
                if (D == 1 || k % D == 0) atomicAdd(&p.band_energy[(size_t(t) * 2 + 1) * B + k / D], double(acc_s[i]));
            }
            if (r == 0) atomicAdd(&p.band_energy[(size_t(t) * 2 + 1) * B + N / (2 * D)], double(acc_s[8]));
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) acc_s[i] = 0.f;
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            const int k = tid + S::THREADS * j;
            if (k < B) {
                if (p.ltas) atomicAdd(&p.ltas[size_t(t) * B + k], double(acc_l[j]));
                if (p.band_energy) atomicAdd(&p.band_energy[(size_t(t) * 2 + 0) * B + k], double(acc_m[j]));
            }
            acc_l[j] = acc_m[j] = 0.f;
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
        const bool want_feat = p.centroid || p.frame_max || p.frame_sum;

        // ------------------------------ FFT phase ------------------------------
        // Every slot of the tile is transformed, frames past the end of the track as zeros, so the
        // whole tile is defined for the epilogue.
        for (int s = g; s < S::SLOTS; s += NG) {
            const int f = s * S::FPS;                // first frame of the slot inside the tile
            const int t = t0 + f;                    // absolute frame
            const long long base = (long long)t * p.hop - N / (2 * D);
            C2 v[16];
            constexpr int NFR = S::FPS;  // frames per slot
            const bool interior = t + NFR - 1 < td.n_frames && base >= 0 && base + (long long)(NFR - 1) * p.hop + N <= td.n_samples;
            if (SH > 0 && interior) {
                // L2 prefetch of the rows this group's next slot adds (NG slots further on in the same track)
                {
                    const long long nb = base + (long long)NG * NFR * p.hop + (long long)r * 32;
                    if (r * 32 < (16 + (NFR - 1) * SH) * M && nb + 32 <= td.n_samples) {
                        prefetch_l2(td.ch0 + nb);
                        if (STEREO) prefetch_l2(td.ch1 + nb);
                    }
                }
                if (STEREO) {
                    const float* __restrict__ La = td.ch0 + base + r;
                    const float* __restrict__ Ra = td.ch1 + base + r;
                    float m[16 + SH], sd[16 + SH];
#pragma unroll
                    for (int j = 0; j < 16 + SH; ++j) {
                        const float l = __ldg(La + j * M), rr = __ldg(Ra + j * M);
                        m[j] = l + rr;
                        sd[j] = l - rr;
                    }
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1) {
                        v[n1].re = make_float2(m[n1] * wreg[n1], m[n1 + SH] * wreg[n1]);
                        v[n1].im = make_float2(sd[n1] * wreg[n1], sd[n1 + SH] * wreg[n1]);
                    }
                } else {
                    const float* __restrict__ Xa = td.ch0 + base + r;
                    float x[16 + 3 * SH];
#pragma unroll
                    for (int j = 0; j < 16 + 3 * SH; ++j) x[j] = __ldg(Xa + j * M);
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1) {
                        v[n1].re = make_float2(x[n1] * wreg[n1], x[n1 + SH] * wreg[n1]);
                        v[n1].im = make_float2(x[n1 + 2 * SH] * wreg[n1], x[n1 + 3 * SH] * wreg[n1]);
                    }
                }
            } else if (STEREO) {
                const float* __restrict__ L = td.ch0;
                const float* __restrict__ R = td.ch1;
                const bool va = t < td.n_frames, vb = t + 1 < td.n_frames;
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const long long na = base + n1 * M + r, nb = na + p.hop;
                    const bool oka = va && na >= 0 && na < td.n_samples, okb = vb && nb >= 0 && nb < td.n_samples;
                    const float2 l = make_float2(oka ? __ldg(L + na) : 0.f, okb ? __ldg(L + nb) : 0.f);
                    const float2 rr = make_float2(oka ? __ldg(R + na) : 0.f, okb ? __ldg(R + nb) : 0.f);
                    v[n1].re = pmuls(padd(l, rr), wreg[n1]);
                    v[n1].im = pmuls(psub(l, rr), wreg[n1]);
                }
            } else {
                // transform A = frame t + i*frame t+2, transform B = frame t+1 + i*frame t+3: the two real
                // spectra of a packed pair are then adjacent frames (t, t+1) and (t+2, t+3).
                const float* __restrict__ X = td.ch0;
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    float x[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const long long n = base + (long long)q * p.hop + n1 * M + r;
                        x[q] = (t + q < td.n_frames && n >= 0 && n < td.n_samples) ? __ldg(X + n) : 0.f;
                    }
                    v[n1].re = pmuls(make_float2(x[0], x[1]), wreg[n1]);
                    v[n1].im = pmuls(make_float2(x[2], x[3]), wreg[n1]);
                }
            }
            pass1<N>(v, r, tw1, ex);
            group_barrier(g, M);
            pass2<N>(v, r, tw2s, ex);
            group_barrier(g, M);
            // one lower-half bin (row k of the tile) from its spectrum value and its mirror value
            auto emit = [&](int kt, const C2& zk, const C2& zn, float& side_acc) {
                if (D > 1 && kt % D) return;   // not a bin of the plan's n_fft
                const int k = kt / D;
                C2 xa, xb;
                split_pair(zk, zn, xa, xb);
                const float2 pa = pfma(xa.re, xa.re, pmul(xa.im, xa.im));
                const float2 pb = pfma(xb.re, xb.re, pmul(xb.im, xb.im));
                *reinterpret_cast<float2*>(tile + k * TFP + f) = make_float2(fast_sqrt(pa.x), fast_sqrt(pa.y));
                if (STEREO) side_acc += pb.x + pb.y;
                else *reinterpret_cast<float2*>(tile + k * TFP + f + 2) = make_float2(fast_sqrt(pb.x), fast_sqrt(pb.y));
            };
            if constexpr (C::NB >= 2) {
                // paired pass 3: every mirror bin is in this thread's registers, no exchange for the split
                using P = Pair3<N>;
                constexpr int Q = C::Q, HB = P::HB;
                pass3_paired<N>(v, r, ex);
                float dump = 0.f;
#pragma unroll
                for (int b = 0; b < HB; ++b) {
                    const bool special = (b == 0 && r == 0);  // (0,0)/(0,8) pair themselves: redone below
#pragma unroll
                    for (int k3 = 0; k3 < Q / 2; ++k3) {
                        emit(P::bin(r, b, 0, k3), v[b * Q + k3], v[(HB + b) * Q + Q - 1 - k3],
                             special ? dump : acc_s[b * Q + k3]);
                        emit(P::bin(r, b, 1, k3), v[(HB + b) * Q + k3], v[b * Q + Q - 1 - k3],
                             special ? dump : acc_s[b * Q + Q / 2 + k3]);
                    }
                }
                if (r == 0) {  // thread 0, pair 0: bins 256*k3 (k3 <= Q/2) mirror inside butterfly (0,0), 128 + 256*k3 inside (0,8)
#pragma unroll
                    for (int k3 = 0; k3 < Q / 2; ++k3) {
                        emit(256 * k3, v[k3], v[(Q - k3) % Q], acc_s[k3]);
                        emit(128 + 256 * k3, v[HB * Q + k3], v[HB * Q + Q - 1 - k3], acc_s[Q / 2 + k3]);
                    }
                    emit(N / 2, v[Q / 2], v[Q / 2], acc_s[8]);
                }
            } else {
                pass3<N>(v, r, ex);
                group_barrier(g, M);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = kept_bin<N>(r, i);
                    C2 zn = unpack(ex[E::slot_of((N - k) & (N - 1))]);
                    if (i == 0 && r == 0) zn = v[kept_reg<N>(i)];  // bin 0 mirrors itself
                    emit(k, v[kept_reg<N>(i)], zn, acc_s[i]);
                }
                if (r == 0) emit(N / 2, v[C::Q / 2], v[C::Q / 2], acc_s[8]);  // bin N/2 (thread 0, k3 = Q/2) mirrors itself
            }
            group_barrier(g, M);  // pass-3 / mirror reads done before the next slot's pass 1 overwrites the slots
        }
        __syncthreads();

        // ------------------------------ epilogue ------------------------------
        {   // (a) one walk over the tile: magnitude rows -> global (64-bit stores, TF contiguous floats per row) and, on
            // the way, this thread's share of the per-frame sums of (c).  A half-warp reads RPH rows that are DR rows
            // apart so that its 16 64-bit words fall into 16 distinct bank pairs; a thread keeps its frame pair and
            // steps KS rows at a time.
            constexpr int RPH = 16 / HP, DR = (TF == 16) ? 8 : (TF == 8) ? 4 : 1, RB = DR * RPH, KS = (S::THREADS / 16 / DR) * RB;
            const int hw = tid >> 4, l16 = tid & 15, fp = l16 % HP, ri = l16 / HP;
            const int k0 = (hw / DR) * RB + (hw % DR) + ri * DR;
            if (p.mag || want_feat) {
                const float* src = tile + k0 * TFP + 2 * fp;
                float2 s1 = make_float2(0.f, 0.f), s2 = s1, mx = s1;
                float kf = float(k0);
                auto take = [&](const float2 a) {
                    s1 = padd(s1, a);
                    s2 = pfmas(a, kf, s2);
                    mx.x = fmaxf(mx.x, a.x);
                    mx.y = fmaxf(mx.y, a.y);
                    kf += float(KS);
                };
                if (p.mag) {
                    float* dst = p.mag + size_t(td.pitch_off) * B + size_t(k0) * td.ld + t0 + 2 * fp;
                    const size_t dstep = size_t(KS) * td.ld;
                    if (nf == TF) {
#pragma unroll 4
                        for (int k = k0; k < B; k += KS, src += KS * TFP, dst += dstep) {
                            const float2 a = *reinterpret_cast<const float2*>(src);
                            *reinterpret_cast<float2*>(dst) = a;
                            take(a);
                        }
                    } else {
                        const bool ok0 = 2 * fp < nf, ok1 = 2 * fp + 1 < nf;
                        for (int k = k0; k < B; k += KS, src += KS * TFP, dst += dstep) {
                            const float2 a = *reinterpret_cast<const float2*>(src);
                            if (ok1) *reinterpret_cast<float2*>(dst) = a;
                            else if (ok0) dst[0] = a.x;
                            take(a);
                        }
                    }
                } else {
#pragma unroll 4
                    for (int k = k0; k < B; k += KS, src += KS * TFP) take(*reinterpret_cast<const float2*>(src));
                }
                if (want_feat) {
                    // the 32 / HP lanes of a warp that share a frame pair, then one slot per (warp, frame pair) in the
                    // exchange area of the last transform group (idle until the next tile's pass 1)
#pragma unroll
                    for (int o = HP; o < 32; o <<= 1) {
                        s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o);
                        s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
                        s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o);
                        s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
                        mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, o));
                        mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, o));
                    }
                    if (lane < HP) {
                        float2* pp = reinterpret_cast<float2*>(ex_all + size_t(NG - 1) * E::SLOTS) + (warp * HP + lane) * 3;
                        pp[0] = s1;
                        pp[1] = s2;
                        pp[2] = mx;
                    }
                }
            }
            // ... and per-bin time sums, one thread per bin (frames past the track end are exact zeros)
            if (p.ltas || p.band_energy) {
#pragma unroll
                for (int j = 0; j < NACC; ++j) {
                    const int k = tid + S::THREADS * j;
                    if (k < B) {
                        const float* row = tile + k * TFP;
                        float2 s = make_float2(0.f, 0.f), q = s;
#pragma unroll
                        for (int h = 0; h < HP; ++h) {
                            const float2 a = *reinterpret_cast<const float2*>(row + 2 * h);
                            s = padd(s, a);
                            q = pfma(a, a, q);
                        }
                        acc_l[j] += s.x + s.y;
                        acc_m[j] += q.x + q.y;
                    }
                }
            }
        }
        // (b) mel projection of tile^2 (power = magnitude**2 as librosa computes it), two frames per thread
        if (p.mel) {
            float vmax = 0.f;
            for (int it = tid; it < p.n_mels * HP; it += S::THREADS) {
                const int fp = it % HP, m = it / HP;
                const int ks = mel_tab[m], len = mel_tab[p.n_mels + m], wo = mel_tab[2 * p.n_mels + m];
                const float* col = tile + ks * TFP + 2 * fp;
                const float2 acc = p.mel_in_smem ? mel_taps<TFP>(col, mel_ws + wo, len) : mel_taps<TFP>(col, p.mel_w + wo, len);
                float* dst = p.mel + size_t(td.pitch_off) * p.n_mels + size_t(m) * td.ld + t0 + 2 * fp;
                if (2 * fp + 1 < nf) {
                    *reinterpret_cast<float2*>(dst) = acc;
                    vmax = fmaxf(vmax, fmaxf(acc.x, acc.y));
                } else if (2 * fp < nf) {
                    dst[0] = acc.x;
                    vmax = fmaxf(vmax, acc.x);
                }
            }
            if (p.mel_max) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
                if (lane == 0) atomicMax(&p.mel_max[trk], __float_as_uint(vmax));
            }
        }
        __syncthreads();
        // (c) per-frame centroid / max / sum: the 16 warps' partial sums s1 = sum |X|, s2 = sum k |X| (float32 over the <= 17
        // rows a thread walked and its 32 / HP lanes) combined in double, in warp order, by one thread per frame:
        // centroid = df * s2 / s1.  (librosa rounds |X| / sum to float32 before the float64 dot product; that changes
        // the result by ~2e-9 relative.)
        if (want_feat && g == NG - 1) {
            if (r < nf) {
                const float* part = reinterpret_cast<const float*>(ex_all + size_t(NG - 1) * E::SLOTS);
                const int fp = r >> 1, h = r & 1;
                double den = 0.0, num = 0.0;
                float m = 0.f;
#pragma unroll 4
                for (int w = 0; w < S::THREADS / 32; ++w) {
                    const float* pp = part + (w * HP + fp) * 6;
                    den += double(pp[h]);
                    num += double(pp[2 + h]);
                    m = fmaxf(m, pp[4 + h]);
                }
                const double df = p.freqs[1];
                if (p.centroid) p.centroid[col_out(td, t0, r)] = (den < 1.1754943508222875e-38) ? df * num : df * num / den;
                if (p.frame_sum) p.frame_sum[col_out(td, t0, r)] = float(den);
                if (p.frame_max) p.frame_max[col_out(td, t0, r)] = m;
            }
            group_barrier(g, M);   // the partial sums are read before this group's next pass 1 overwrites them
        }
    }
    flush(trk);
}

template <int N, int TF, bool STEREO, int SH, int D = 1>
static int launch_stft(const ta_plan* plan, StftParams p, cudaStream_t stream) {
    using S = StftCfg<N, TF, STEREO, D>;
    auto kern = stft_fused_kernel<N, TF, STEREO, SH, D>;
    size_t smem = S::fixed + S::mel_tab_bytes(p.n_mels);
    p.mel_in_smem = (smem + S::mel_w_bytes(p.mel_nnz) <= SMEM_LIMIT) ? 1 : 0;
    if (p.mel_in_smem) smem += S::mel_w_bytes(p.mel_nnz);
    if (smem > SMEM_LIMIT) {
        set_error("mel band table does not fit in shared memory next to the STFT tile");
        return TA_ERR_UNSUPPORTED;
    }
    TA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::max(1, std::min(plan->sm_count, p.total_tiles));
    kern<<<grid, 512, smem, stream>>>(p);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
