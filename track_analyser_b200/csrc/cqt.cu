// K12: chroma_cqt = tuning estimate + 2:1 decimation chain + constant-Q transform + chroma fold + per-frame inf-norm.
//
// Replaces librosa.feature.chroma_cqt(y=y, sr=sr) (every default: hop 512, fmin C1, 7 octaves x 36 bins, norm=inf,
// tuning=None) as called from harmony.py:107 (key_estimate) and harmony.py:148 (analyse_harmony).  librosa.cqt is the
// recursive ("pseudo") constant-Q transform: per octave, from the top one down, a centred rectangular-window STFT of
// the (progressively 2:1 decimated) signal is multiplied by the 1 %-sparsified spectra of 36 Hann-windowed complex
// exponentials; hop and rate halve from octave to octave, so every octave has the same transform length, the same
// number of frames and -- up to float rounding -- the same basis.  Stages here, all on the device:
//   tuning     librosa.estimate_tuning(y=y, bins_per_octave=36): piptrack on the MAGNITUDE spectrogram K1 already wrote
//              (chroma.cu: run_tuning), histogram arg-max bin 0..99 -> tuning = -0.5 + 0.01 * bin.
//   basis      at first use, for all 100 possible tunings x 7 octaves x 36 filters: the time-domain wavelet in float64
//              (phasor x periodic Hann, L1-normalised, rounded to complex64, scaled by length / n_fft), its float64 DFT,
//              librosa.util.sparsify_rows(quantile = 0.01), times sqrt(2^octave), complex64 -- cqt_basis_kernel.
//   decimate   y -> mono -> x_1 -> x_2 ...: x_{s+1}[m] = sqrt(2) * sum_j h[j] x_s[2m + j], ceil(n/2) samples.  librosa uses
//              libsoxr ("soxr_hq"), which cannot be restated here: h is the zero-phase Kaiser-windowed sinc with soxr HQ's
//              published band edges (0.913 / 1.0 of the new Nyquist, 126.4 dB, 381 taps), the same filter as
//              oracle/cqt_np.py:decimator_taps -- a stated, swappable stage (PARITY UNPINNED).
//   transform  cqt_chroma_kernel: one CTA walks tiles of 32 frames; per octave, eight thread groups each transform four
//              real frames with one packed complex 1024-point FFT pair (fft2_core.cuh), the Hermitian split puts the
//              complex spectrum of the ~115 bins any filter touches into shared memory, 36 x 32 (filter, frame)
//              projections, |.| / sqrt(length), and the chroma fold (three adjacent bins per pitch class) accumulated
//              in registers over the seven octaves; finally each frame is divided by its maximum.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "fft2_core.cuh"

namespace ta {

static constexpr int CQ_BPO = 36, CQ_OCT = 7, CQ_BINS = CQ_BPO * CQ_OCT, CQ_TUNINGS = 100, CQ_SPAN = 32;
static constexpr int CQ_N = 1024, CQ_TF = 32, CQ_SPITCH = CQ_TF + 1;  // transform length, frames per tile, spectrum row pitch
static constexpr int CQ_HOP = 512;                                    // chroma_cqt's default hop_length
static constexpr int DEC_HALF = 190;                                  // 381 taps
static constexpr int DEC_NE = 95;                                     // even taps h[2i], i = -95..95; odd taps h[2i+1], i = -95..94
static constexpr int DEC_THREADS = 256, DEC_R = 4, DEC_TILE = DEC_THREADS * DEC_R;  // outputs per CTA
static constexpr int DEC_HALO = 96;                                   // halo of the even/odd streams (multiple of 4, >= 95)
static constexpr int DEC_SM = DEC_TILE + 2 * DEC_HALO;                // floats per stream in shared memory
static constexpr int CQ_MAX_STAGES = 10;

__constant__ float c_dec_even[2 * DEC_NE + 1];  // sqrt(2) * h[2i],     i = -95..95
__constant__ float c_dec_odd[2 * DEC_NE];       // sqrt(2) * h[2i + 1], i = -95..94

struct CqtTables {
    int early = 0;            // librosa's early down-sampling count
    int nfft = 0;             // per-octave transform length (256, 512 or 1024)
    int bin_step = 1;         // CQ_N / nfft: basis bin j is bin j * bin_step of the zero-padded 1024-point transform
    int first_stage = 0;      // first decimation stage kept in scratch (0: the mono mix itself is octave 0's signal)
    int last_stage = 0;       // early + 6
    int j_lo = 0, j_hi = 0;   // range of basis bins (in units of bin_step) any filter touches
    float2* d_basis = nullptr;        // [100][7][36][CQ_SPAN]
    short2* d_span = nullptr;         // [100][7][36] (first basis bin, count)
    float* d_inv_sqrt_len = nullptr;  // [100][252]
    float2* d_tw1 = nullptr;          // 1024-point twiddles of the packed FFT core
    float2* d_tw2 = nullptr;
    std::string error;        // non-empty: this configuration is not supported
};

struct CqtTrack {
    const float* sig[CQ_OCT];
    long long len[CQ_OCT];
    long long pitch_off;
    int n_frames;
    int ld;
};

struct DecTrack {
    const float* a;
    const float* b;   // second channel (mono mix of a stereo batch), else nullptr
    float* dst;
    long long n_in, n_out;
};

// ------------------------------------------------------------------------------------------------------------------
// basis: librosa.filters.wavelet -> __vqt_filter_fft -> util.sparsify_rows, one CTA per (octave * 36 + filter, tuning)
// ------------------------------------------------------------------------------------------------------------------
struct BasisParams {
    const double* freqs;   // [100][252]
    double Q;
    double sr0;            // rate of the top octave
    int nfft;
    float2* basis;
    short2* span;
    int* range;            // [0] min first bin, [1] max last bin + 1, [2] error flag (a span wider than CQ_SPAN)
};

__global__ void __launch_bounds__(256) cqt_basis_kernel(const BasisParams p) {
    __shared__ float2 sig[CQ_N];
    __shared__ double2 tw[CQ_N];
    __shared__ double hre[CQ_N / 2 + 1], him[CQ_N / 2 + 1];
    __shared__ double srt[CQ_N];
    __shared__ double red[256];
    __shared__ double s_thr;
    __shared__ int s_first, s_last;
    const int tid = threadIdx.x;
    const int oct = blockIdx.x / CQ_BPO, filt = blockIdx.x % CQ_BPO, ti = blockIdx.y;
    const int k = CQ_BINS - CQ_BPO * (oct + 1) + filt;     // global bin
    const double freq = p.freqs[ti * CQ_BINS + k];
    const double sr = p.sr0 / double(1 << oct);
    const int nfft = p.nfft, nb = nfft / 2 + 1;
    const double ilen = p.Q * sr / freq;                    // wavelet_lengths, gamma = 0
    const double t0 = floor(-ilen / 2.0), t1 = floor(ilen / 2.0);
    const int n = int(t1 - t0);
    const int lpad = (nfft - n) / 2;                        // util.pad_center
    const double PI = 3.141592653589793;
    for (int i = tid; i < nfft; i += 256) sig[i] = make_float2(0.f, 0.f);
    if (tid == 0) {
        s_first = nb;
        s_last = -1;
    }
    // phasor * periodic Hann in float64 (real parts staged in srt, imaginary parts in the not yet filled twiddle table)
    double* stage_im = reinterpret_cast<double*>(tw);
    double part = 0.0;
    for (int i = tid; i < n; i += 256) {
        const double t = t0 + double(i);
        const double ang = t * 2.0 * PI * freq / sr;        // evaluated left to right like numpy
        const double fac = double(i) * (2.0 * PI / double(n)) + (-PI);
        const double w = 0.5 + 0.5 * cos(fac);              // scipy general_cosine(M, [0.5, 0.5], sym=False)
        const double re = cos(ang) * w, im = sin(ang) * w;
        part += hypot(re, im);
        srt[i] = re;
        stage_im[i] = im;
    }
    red[tid] = part;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    const double norm = red[0];                             // util.normalize(norm=1)
    const double scale = ilen / double(nfft);
    for (int i = tid; i < n; i += 256) {
        const float re32 = float(srt[i] / norm), im32 = float(stage_im[i] / norm);               // complex64
        sig[lpad + i] = make_float2(float(double(re32) * scale), float(double(im32) * scale));    // basis *= lengths / n_fft
    }
    __syncthreads();
    for (int i = tid; i < nfft; i += 256) {
        double sn, cs;
        sincospi(2.0 * double(i) / double(nfft), &sn, &cs);
        tw[i] = make_double2(cs, -sn);
    }
    __syncthreads();
    // float64 DFT of the padded filter, bins 0 .. nfft/2
    for (int j = tid; j < nb; j += 256) {
        double ar = 0.0, ai = 0.0;
        int idx = int((long long)j * lpad % nfft);
        for (int i = 0; i < n; ++i) {
            const float2 b = sig[lpad + i];
            const double2 w = tw[idx];
            ar += double(b.x) * w.x - double(b.y) * w.y;
            ai += double(b.x) * w.y + double(b.y) * w.x;
            idx += j;
            if (idx >= nfft) idx -= nfft;
        }
        hre[j] = ar;
        him[j] = ai;
    }
    __syncthreads();
    // sparsify_rows: mags, their sum, ascending sort, sequential cumsum of mag_sort / norm, threshold = first >= 0.01
    part = 0.0;
    for (int j = tid; j < CQ_N; j += 256) {
        const double m = (j < nb) ? hypot(hre[j], him[j]) : INFINITY;
        srt[j] = m;
        if (j < nb) part += m;
    }
    red[tid] = part;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    const double norms = red[0];
    for (int size = 2; size <= CQ_N; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = tid; i < CQ_N / 2; i += 256) {
                const int lo = 2 * i - (i & (stride - 1));   // index with bit `stride` clear
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const double a = srt[lo], b = srt[hi];
                if ((a > b) == up) {
                    srt[lo] = b;
                    srt[hi] = a;
                }
            }
        }
    __syncthreads();
    if (tid == 0) {
        double c = 0.0;
        int idx = 0;
        for (int j = 0; j < nb; ++j) {
            c += srt[j] / norms;
            if (!(c < 0.01)) {
                idx = j;
                break;
            }
        }
        s_thr = srt[idx];
    }
    __syncthreads();
    const double thr = s_thr;
    for (int j = tid; j < nb; j += 256)
        if (hypot(hre[j], him[j]) >= thr) {
            atomicMin(&s_first, j);
            atomicMax(&s_last, j);
        }
    __syncthreads();
    const int first = s_first, last = s_last, cnt = last - first + 1;
    const size_t row = (size_t(ti) * CQ_OCT + oct) * CQ_BPO + filt;
    if (cnt > CQ_SPAN || cnt <= 0) {
        if (tid == 0) {
            atomicExch(&p.range[2], 1);
            p.span[row] = make_short2(0, 0);
        }
        return;
    }
    const float oscale = float(sqrt(double(1 << oct)));     // fft_basis[:] *= sqrt(sr / my_sr)
    if (tid < CQ_SPAN) {
        const int j = first + tid;
        float2 v = make_float2(0.f, 0.f);
        if (j <= last && hypot(hre[j], him[j]) >= thr) v = make_float2(float(hre[j]) * oscale, float(him[j]) * oscale);
        p.basis[row * CQ_SPAN + tid] = v;
    }
    if (tid == 0) {
        p.span[row] = make_short2(short(first), short(cnt));
        atomicMin(&p.range[0], first);
        atomicMax(&p.range[1], last + 1);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// decimation: x_{s+1}[m] = sum_j taps[j] x_s[2m + j]  (taps already carry the sqrt(2) of librosa's scale=True)
// Even and odd input samples are staged as two streams so that the taps of one parity walk one stream with unit
// stride; a thread owns DEC_R = 4 consecutive outputs, reads its window with 128-bit loads (conflict-free, every value
// feeds up to 4 outputs) and takes the taps as immediate constant-bank operands (the tap loop is fully unrolled).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DEC_THREADS) cqt_decimate_kernel(const DecTrack* __restrict__ tracks) {
    __shared__ __align__(16) float se[DEC_SM];
    __shared__ __align__(16) float so[DEC_SM];
    const DecTrack td = tracks[blockIdx.y];
    const long long m_tile = (long long)blockIdx.x * DEC_TILE;
    if (m_tile >= td.n_out) return;
    const int tid = threadIdx.x;
    // stream index q <-> input samples 2 * (m_tile - HALO + q) (even stream) and the one after it (odd stream)
    const long long n_base = 2 * (m_tile - DEC_HALO);
    for (int j = tid; j < 2 * DEC_SM; j += DEC_THREADS) {
        const long long n = n_base + j;
        float v = 0.f;
        if (n >= 0 && n < td.n_in) v = td.b ? (__ldg(td.a + n) + __ldg(td.b + n)) * 0.5f : __ldg(td.a + n);
        ((j & 1) ? so : se)[j >> 1] = v;
    }
    __syncthreads();
    // outputs m_tile + 4 tid + r;  out[r] = sum_i he[i] E[m + i] + sum_i ho[i] O[m + i], window chunk c holds d = 4c + e - 96
    float acc[DEC_R] = {0.f, 0.f, 0.f, 0.f};
    const float4* pe = reinterpret_cast<const float4*>(se + DEC_R * tid);
    const float4* po = reinterpret_cast<const float4*>(so + DEC_R * tid);
    constexpr int NCH = (2 * DEC_HALO + DEC_R) / 4;  // 49 chunks cover d = -96 .. 99
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const float4 ve = pe[c], vo = po[c];
        const float xe[4] = {ve.x, ve.y, ve.z, ve.w}, xo[4] = {vo.x, vo.y, vo.z, vo.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int d = 4 * c + e - DEC_HALO;
#pragma unroll
            for (int r = 0; r < DEC_R; ++r) {
                const int i = d - r;
                if (i >= -DEC_NE && i <= DEC_NE) acc[r] = fmaf(c_dec_even[i + DEC_NE], xe[e], acc[r]);
                if (i >= -DEC_NE && i < DEC_NE) acc[r] = fmaf(c_dec_odd[i + DEC_NE], xo[e], acc[r]);
            }
        }
    }
    const long long m0 = m_tile + DEC_R * tid;
    if (m0 + DEC_R <= td.n_out) {
        *reinterpret_cast<float4*>(td.dst + m0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
        for (int r = 0; r < DEC_R; ++r)
            if (m0 + r < td.n_out) td.dst[m0 + r] = acc[r];
    }
}

// mono mix of a stereo batch / copy of a mono one (only when the top octave reads the undecimated signal: early == 0)
__global__ void __launch_bounds__(256) cqt_mono_kernel(const DecTrack* __restrict__ tracks) {
    const DecTrack td = tracks[blockIdx.y];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < td.n_out; i += (long long)gridDim.x * blockDim.x)
        td.dst[i] = td.b ? (td.a[i] + td.b[i]) * 0.5f : td.a[i];
}

// ------------------------------------------------------------------------------------------------------------------
// transform + projection + chroma
// ------------------------------------------------------------------------------------------------------------------
struct CqtParams {
    const CqtTrack* tracks;
    const int* tuning_idx;       // [n_tracks]
    const float2* basis;
    const short2* span;
    const float* inv_sqrt_len;
    const float2* tw1;
    const float2* tw2;
    int nfft, bin_step, hop0;    // hop of the top octave
    int j_lo, n_rows;            // first basis bin kept in shared memory, number of rows
    float* chroma;               // [12 * Pc]
    float* cqt_mag;              // [252 * Pc] or nullptr
};

__device__ __forceinline__ void cq_barrier(int g, int n) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(n) : "memory"); }

// BS: bin step (1024 / n_fft of the octave transforms), a compile-time 1, 2 or 4
template <int BS>
__global__ void __launch_bounds__(512, 1) cqt_chroma_kernel(const CqtParams p) {
    using namespace p2;
    using C = FftCfg<CQ_N>;
    using E = Ex<CQ_N>;
    using P = Pair3<CQ_N>;
    constexpr int M = C::M, NG = 512 / M, N = CQ_N, Q = C::Q, HB = P::HB;
    static_assert(NG * 4 == CQ_TF, "one round of slots (4 frames each) fills a tile");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* ex_all = reinterpret_cast<float4*>(smem_raw);
    float2* tw1s = reinterpret_cast<float2*>(ex_all + size_t(NG) * E::SLOTS);
    float2* tw2s = tw1s + 15 * M;
    float2* spec = tw2s + 16 * Q;                                   // [n_rows][CQ_SPITCH]
    float* mags = reinterpret_cast<float*>(spec + size_t(p.n_rows) * CQ_SPITCH);  // [36][32]
    float* chr = mags + CQ_BPO * CQ_TF;                             // [12][32]

    const int tid = threadIdx.x, g = tid / M, r = tid % M;
    float4* ex = ex_all + size_t(g) * E::SLOTS;
    for (int i = tid; i < 15 * M; i += 512) tw1s[i] = p.tw1[i];
    for (int i = tid; i < 16 * Q; i += 512) tw2s[i] = p.tw2[i];
    __syncthreads();

    const int trk = blockIdx.y;
    const CqtTrack* __restrict__ tdp = p.tracks + trk;
    struct { long long pitch_off; int n_frames, ld; } td = {tdp->pitch_off, tdp->n_frames, tdp->ld};
    const int ti = p.tuning_idx[trk];
    const int n_tiles = (td.n_frames + CQ_TF - 1) / CQ_TF;
    const int nfft = CQ_N / BS;
    const int cfr = tid & 31, cc = tid >> 5;   // chroma owner: frame, pitch class (tid < 384)

    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x) {
        const int t0 = w * CQ_TF, nf = min(CQ_TF, td.n_frames - t0);
        float cacc = 0.f;
        for (int oct = 0; oct < CQ_OCT; ++oct) {
            const float* __restrict__ x = tdp->sig[oct];
            const long long len = tdp->len[oct];
            const int hop = p.hop0 >> oct;
            {   // four frames t .. t+3 of this group: A = frame t + i*frame t+2, B = frame t+1 + i*frame t+3
                const int f = 4 * g, t = t0 + f;
                C2 v[16];
                const long long s0 = (long long)t * hop - nfft / 2 + r;   // sample of (frame t, row 0) for this thread
                // the whole slot inside the signal and the track: unchecked loads (rows n >= nfft of a zero-padded transform
                // are compile-time zeros)
                if (t + 3 < td.n_frames && s0 - r >= 0 && s0 - r + 3LL * hop + nfft <= len) {
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1) {
                        if (n1 * M >= nfft) {
                            v[n1].re = v[n1].im = make_float2(0.f, 0.f);
                        } else {
                            const float* __restrict__ q0 = x + s0 + n1 * M;
                            v[n1].re = pmuls(make_float2(__ldg(q0), __ldg(q0 + hop)), 0.5f);
                            v[n1].im = pmuls(make_float2(__ldg(q0 + 2 * hop), __ldg(q0 + 3 * hop)), 0.5f);
                        }
                    }
                } else {
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1) {
                        const int n = n1 * M + r;
                        float xs[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const long long s = s0 + (long long)q * hop + n1 * M;
                            xs[q] = (n < nfft && t + q < td.n_frames && s >= 0 && s < len) ? __ldg(x + s) : 0.f;
                        }
                        v[n1].re = pmuls(make_float2(xs[0], xs[1]), 0.5f);
                        v[n1].im = pmuls(make_float2(xs[2], xs[3]), 0.5f);
                    }
                }
                pass1<N>(v, r, tw1s, ex);
                cq_barrier(g, M);
                pass2<N>(v, r, tw2s, ex);
                cq_barrier(g, M);
                pass3_paired<N>(v, r, ex);
                auto emit = [&](int k, const C2& zk, const C2& zn) {
                    if (BS > 1 && k % BS) return;
                    const int row = k / BS - p.j_lo;
                    if (row < 0 || row >= p.n_rows) return;
                    C2 xa, xb;
                    split_pair(zk, zn, xa, xb);
                    float2* dst = spec + row * CQ_SPITCH + f;
                    dst[0] = make_float2(xa.re.x, xa.im.x);
                    dst[1] = make_float2(xa.re.y, xa.im.y);
                    dst[2] = make_float2(xb.re.x, xb.im.x);
                    dst[3] = make_float2(xb.re.y, xb.im.y);
                };
#pragma unroll
                for (int b = 0; b < HB; ++b) {
                    if (b == 0 && r == 0) continue;  // (0,0)/(0,8) pair themselves: handled below
#pragma unroll
                    for (int k3 = 0; k3 < Q / 2; ++k3) {
                        emit(P::bin(r, b, 0, k3), v[b * Q + k3], v[(HB + b) * Q + Q - 1 - k3]);
                        emit(P::bin(r, b, 1, k3), v[(HB + b) * Q + k3], v[b * Q + Q - 1 - k3]);
                    }
                }
                if (r == 0) {
#pragma unroll
                    for (int k3 = 0; k3 < Q / 2; ++k3) {
                        emit(256 * k3, v[k3], v[(Q - k3) % Q]);
                        emit(128 + 256 * k3, v[HB * Q + k3], v[HB * Q + Q - 1 - k3]);
                    }
                    emit(N / 2, v[Q / 2], v[Q / 2]);
                }
            }
            __syncthreads();
            // projection: (filter, frame) pairs, lane = frame
            const size_t brow = (size_t(ti) * CQ_OCT + oct) * CQ_BPO;
            for (int idx = tid; idx < CQ_BPO * CQ_TF; idx += 512) {
                const int fr = idx & 31, filt = idx >> 5;
                const short2 sp = p.span[brow + filt];
                const float2* __restrict__ bw = p.basis + (brow + filt) * CQ_SPAN;
                const float2* col = spec + (sp.x - p.j_lo) * CQ_SPITCH + fr;
                float ar = 0.f, ai = 0.f;
                for (int i = 0; i < sp.y; ++i) {
                    const float2 b = __ldg(bw + i), s = col[i * CQ_SPITCH];
                    ar = fmaf(b.x, s.x, fmaf(-b.y, s.y, ar));
                    ai = fmaf(b.x, s.y, fmaf(b.y, s.x, ai));
                }
                const int kg = CQ_BINS - CQ_BPO * (oct + 1) + filt;
                const float sc = p.inv_sqrt_len[ti * CQ_BINS + kg];
                ar *= sc;
                ai *= sc;
                const float m = sqrtf(ar * ar + ai * ai);
                mags[filt * CQ_TF + fr] = m;
                if (p.cqt_mag && fr < nf) p.cqt_mag[size_t(td.pitch_off) * CQ_BINS + size_t(kg) * td.ld + t0 + fr] = m;
            }
            __syncthreads();
            // chroma fold: pitch class c collects filters 3c-1, 3c, 3c+1 (mod 36) of every octave (filters.cq_to_chroma)
            if (tid < 12 * CQ_TF) {
                const int f0 = (3 * cc + CQ_BPO - 1) % CQ_BPO;
                cacc += mags[f0 * CQ_TF + cfr] + mags[(3 * cc) * CQ_TF + cfr] + mags[(3 * cc + 1) * CQ_TF + cfr];
            }
        }
        // util.normalize(norm=inf) over the 12 pitch classes of each frame
        if (tid < 12 * CQ_TF) chr[cc * CQ_TF + cfr] = cacc;
        __syncthreads();
        if (tid < 12 * CQ_TF && cfr < nf) {
            float mx = 0.f;
#pragma unroll
            for (int c = 0; c < 12; ++c) mx = fmaxf(mx, chr[c * CQ_TF + cfr]);
            const float l = (mx < 1.1754943508222875e-38f) ? 1.0f : mx;
            p.chroma[size_t(td.pitch_off) * 12 + size_t(cc) * td.ld + t0 + cfr] = cacc / l;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static double bessel_i0(double x) {
    double s = 1.0, t = 1.0;
    for (int k = 1; k < 500; ++k) {
        t *= (x * 0.5) * (x * 0.5) / (double(k) * double(k));
        s += t;
        if (t < 1e-20 * s) break;
    }
    return s;
}

// the decimator of oracle/cqt_np.py:decimator_taps (Kaiser window method on soxr HQ's band edges), float64
static void decimator_taps(std::vector<double>& h) {
    const double PI = 3.14159265358979323846;
    const double att = 21.0 * 20.0 * std::log10(2.0);
    const double beta = 0.1102 * (att - 8.7);
    const double width = PI * (1.0 - 0.913) / 2.0;
    int n = int(std::ceil((att - 7.95) / (2.285 * width))) + 1;
    if (n % 2 == 0) n += 1;
    const int half = n / 2;
    const double fc = 0.25 * (0.913 + 1.0) / 2.0;
    h.assign(n, 0.0);
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const double m = double(i - half);
        const double x = 2.0 * fc * m;
        const double sinc = (m == 0.0) ? 1.0 : std::sin(PI * x) / (PI * x);
        const double rr = 2.0 * double(i) / double(n - 1) - 1.0;            // np.kaiser: I0(beta sqrt(1 - ((i - a)/a)^2)) / I0(beta)
        const double wk = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - rr * rr))) / bessel_i0(beta);
        h[i] = 2.0 * fc * sinc * wk;
        sum += h[i];
    }
    for (auto& v : h) v /= sum;
}

static const double CQ_WINDOW_BW = 1.50018310546875;  // librosa.filters.WINDOW_BANDWIDTHS["hann"]

static void cqt_freqs(int ti, double* freqs /* [252] */) {
    const double C1 = 440.0 * std::pow(2.0, (24 - 69.0) / 12.0);      // note_to_hz("C1")
    const double tuning = double(ti) * 0.01 + (-0.5);                   // np.linspace(-0.5, 0.5, 101)[ti]
    const double fmin = C1 * std::pow(2.0, tuning / CQ_BPO);
    for (int k = 0; k < CQ_BINS; ++k) freqs[k] = fmin * std::pow(2.0, double(k) / CQ_BPO);
}

static CqtTables* cqt_tables_build(const ta_plan* plan) {
    CqtTables* t = new CqtTables();
    const double sr = double(plan->desc.sample_rate);
    const double rr = std::pow(2.0, 1.0 / CQ_BPO);
    const double alpha = (rr * rr - 1.0) / (rr * rr + 1.0);
    const double Q = 1.0 / alpha;
    std::vector<double> freqs(size_t(CQ_TUNINGS) * CQ_BINS);
    int early = -1, nfft = -1;
    for (int ti = 0; ti < CQ_TUNINGS; ++ti) {
        double* f = freqs.data() + size_t(ti) * CQ_BINS;
        cqt_freqs(ti, f);
        double cutoff = 0.0;
        for (int k = 0; k < CQ_BINS; ++k) cutoff = std::max(cutoff, f[k] * (1.0 + 0.5 * CQ_WINDOW_BW / Q));
        const double nyq = sr / 2.0;
        if (cutoff > nyq) {
            t->error = "chroma_cqt: the constant-Q basis exceeds the Nyquist frequency at this sample rate (librosa raises ParameterError)";
            return t;
        }
        const int c1 = std::max(0, int(std::ceil(std::log2(nyq / cutoff)) - 1) - 1);
        const int c2 = std::max(0, 9 - CQ_OCT + 1);  // hop 512 = 2^9
        const int e = std::min(c1, c2);
        // transform length of every octave: lengths of the lowest filter of the octave at the octave's rate
        const double sr0 = sr / double(1 << e);
        int nf = -1;
        for (int o = 0; o < CQ_OCT; ++o) {
            const double sro = sr0 / double(1 << o);
            const double lmax = Q * sro / f[CQ_BINS - CQ_BPO * (o + 1)];
            const int v = int(std::pow(2.0, std::ceil(std::log2(lmax))));
            if (nf < 0) nf = v;
            if (v != nf) nf = 0;
        }
        if (early < 0) {
            early = e;
            nfft = nf;
        }
        if (e != early || nf != nfft || nf <= 0) {
            t->error = "chroma_cqt: this sample rate sits on a boundary of librosa's early-downsampling / filter-length rules "
                       "(the configuration would depend on the estimated tuning); not supported";
            return t;
        }
    }
    if (nfft != 256 && nfft != 512 && nfft != 1024) {
        t->error = "chroma_cqt: unsupported constant-Q transform length " + std::to_string(nfft);
        return t;
    }
    t->early = early;
    t->nfft = nfft;
    t->bin_step = CQ_N / nfft;
    t->first_stage = early == 0 ? 0 : 1;
    t->last_stage = early + CQ_OCT - 1;
    const double sr0 = sr / double(1 << early);

    auto fail = [&](const char* what) {
        t->error = std::string("chroma_cqt table construction failed: ") + what + ": " + cudaGetErrorString(cudaGetLastError());
        return t;
    };
    double* d_freqs = nullptr;
    int* d_range = nullptr;
    if (cudaMalloc(&d_freqs, freqs.size() * sizeof(double)) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&d_range, 3 * sizeof(int)) != cudaSuccess) return fail("cudaMalloc");
    const size_t rows = size_t(CQ_TUNINGS) * CQ_OCT * CQ_BPO;
    if (cudaMalloc(&t->d_basis, rows * CQ_SPAN * sizeof(float2)) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&t->d_span, rows * sizeof(short2)) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&t->d_inv_sqrt_len, size_t(CQ_TUNINGS) * CQ_BINS * sizeof(float)) != cudaSuccess) return fail("cudaMalloc");
    const int range0[3] = {1 << 20, 0, 0};
    cudaMemcpy(d_freqs, freqs.data(), freqs.size() * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_range, range0, sizeof(range0), cudaMemcpyHostToDevice);
    BasisParams bp{d_freqs, Q, sr0, nfft, t->d_basis, t->d_span, d_range};
    cqt_basis_kernel<<<dim3(CQ_OCT * CQ_BPO, CQ_TUNINGS), 256>>>(bp);
    count_launch();
    int range[3] = {0, 0, 0};
    if (cudaMemcpy(range, d_range, sizeof(range), cudaMemcpyDeviceToHost) != cudaSuccess) return fail("cqt_basis_kernel");
    cudaFree(d_freqs);
    cudaFree(d_range);
    if (range[2] || range[1] <= range[0]) {
        t->error = "chroma_cqt: a sparsified basis row spans more than " + std::to_string(CQ_SPAN) + " bins";
        return t;
    }
    t->j_lo = range[0];
    t->j_hi = range[1];
    // V /= sqrt(lengths), lengths at the rate of the top octave (vqt recomputes them after the early down-sampling)
    std::vector<float> isl(size_t(CQ_TUNINGS) * CQ_BINS);
    for (int ti = 0; ti < CQ_TUNINGS; ++ti)
        for (int k = 0; k < CQ_BINS; ++k) isl[size_t(ti) * CQ_BINS + k] = float(1.0 / std::sqrt(Q * sr0 / freqs[size_t(ti) * CQ_BINS + k]));
    cudaMemcpy(t->d_inv_sqrt_len, isl.data(), isl.size() * sizeof(float), cudaMemcpyHostToDevice);
    {   // twiddles of the 1024-point packed core (same tables as the tempogram's)
        const double PI = 3.14159265358979323846;
        const int TM = CQ_N / 16, TQ = CQ_N / 256;
        std::vector<float2> t1(size_t(15) * TM), t2(size_t(16) * TQ);
        for (int k1 = 1; k1 < 16; ++k1)
            for (int r = 0; r < TM; ++r) {
                const double a = -2.0 * PI * double((r * k1) % CQ_N) / CQ_N;
                t1[size_t(k1 - 1) * TM + r] = make_float2(float(std::cos(a)), float(std::sin(a)));
            }
        for (int k2 = 0; k2 < 16; ++k2)
            for (int n3 = 0; n3 < TQ; ++n3) {
                const double a = -2.0 * PI * double(n3 * k2) / TM;
                t2[size_t(k2) * TQ + n3] = make_float2(float(std::cos(a)), float(std::sin(a)));
            }
        if (cudaMalloc(&t->d_tw1, t1.size() * sizeof(float2)) != cudaSuccess) return fail("cudaMalloc");
        if (cudaMalloc(&t->d_tw2, t2.size() * sizeof(float2)) != cudaSuccess) return fail("cudaMalloc");
        cudaMemcpy(t->d_tw1, t1.data(), t1.size() * sizeof(float2), cudaMemcpyHostToDevice);
        cudaMemcpy(t->d_tw2, t2.data(), t2.size() * sizeof(float2), cudaMemcpyHostToDevice);
    }
    {   // decimator taps, sqrt(2) folded in, split by parity
        std::vector<double> h;
        decimator_taps(h);
        if (int(h.size()) != 2 * DEC_HALF + 1) {
            t->error = "chroma_cqt: decimator design produced " + std::to_string(h.size()) + " taps, expected 381";
            return t;
        }
        float he[2 * DEC_NE + 1], ho[2 * DEC_NE];
        const double s2 = std::sqrt(2.0);
        for (int i = -DEC_NE; i <= DEC_NE; ++i) he[i + DEC_NE] = float(h[2 * i + DEC_HALF] * s2);
        for (int i = -DEC_NE; i < DEC_NE; ++i) ho[i + DEC_NE] = float(h[2 * i + 1 + DEC_HALF] * s2);
        if (cudaMemcpyToSymbol(c_dec_even, he, sizeof(he)) != cudaSuccess) return fail("cudaMemcpyToSymbol");
        if (cudaMemcpyToSymbol(c_dec_odd, ho, sizeof(ho)) != cudaSuccess) return fail("cudaMemcpyToSymbol");
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return fail("cudaDeviceSynchronize");
    return t;
}

void cqt_tables_free(CqtTables* t) {
    if (!t) return;
    cudaFree(t->d_basis);
    cudaFree(t->d_span);
    cudaFree(t->d_inv_sqrt_len);
    cudaFree(t->d_tw1);
    cudaFree(t->d_tw2);
    delete t;
}

// the tables of a plan, built by the first caller; nullptr + last error when the configuration is unsupported
static const CqtTables* cqt_tables(const ta_plan* plan) {
    std::lock_guard<std::mutex> lock(plan->cqt_mutex);
    if (!plan->cqt) {
        cudaSetDevice(plan->desc.device);
        plan->cqt = cqt_tables_build(plan);
    }
    if (!plan->cqt->error.empty()) {
        set_error(plan->cqt->error);
        return nullptr;
    }
    return plan->cqt;
}

// stage lengths x_0 = n, x_{s+1} = ceil(x_s / 2), and the frame count of the stacked transform (librosa __trim_stack)
static void cqt_lengths(const CqtTables* t, int64_t n, int64_t* len /* [last_stage + 1] */, int64_t& frames) {
    len[0] = n;
    for (int s = 1; s <= t->last_stage; ++s) len[s] = (len[s - 1] + 1) / 2;
    frames = INT64_MAX;
    for (int o = 0; o < CQ_OCT; ++o) {
        const int s = t->early + o;
        frames = std::min<int64_t>(frames, 1 + len[s] / (CQ_HOP >> s));
    }
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

size_t tuning_scratch_bytes(const ta_plan* plan, const HostBatch& hb);
int run_tuning(const ta_plan*, const HostBatch&, const TrackDesc*, const float* mag, const float* frame_max, bool power, int bpo,
               void* scratch, double* tuning, int* tuning_idx, cudaStream_t);

int cqt_supported(const ta_plan* plan) {
    if (plan->desc.n_fft != 2048 || plan->desc.hop != CQ_HOP) {
        set_error("chroma_cqt needs the default plan (n_fft 2048, hop 512): librosa.estimate_tuning(y=...) and chroma_cqt use those");
        return TA_ERR_UNSUPPORTED;
    }
    return cqt_tables(plan) ? TA_OK : TA_ERR_UNSUPPORTED;
}

int64_t cqt_frame_count(const ta_plan* plan, int64_t n_samples) {
    const CqtTables* t = cqt_tables(plan);
    if (!t) return -1;
    int64_t len[CQ_MAX_STAGES], frames;
    cqt_lengths(t, n_samples, len, frames);
    return frames;
}

size_t cqt_scratch_bytes(const ta_plan* plan, const HostBatch& hb) {
    const CqtTables* t = cqt_tables(plan);
    if (!t) return 0;
    size_t bytes = 0;
    bytes += align256(sizeof(CqtTrack) * hb.n_tracks);
    bytes += align256(sizeof(DecTrack) * hb.n_tracks) * (t->last_stage + 1);
    bytes += align256(sizeof(int) * hb.n_tracks);
    bytes += align256(tuning_scratch_bytes(plan, hb));
    for (auto& tr : hb.tracks) {
        int64_t len[CQ_MAX_STAGES], frames;
        cqt_lengths(t, tr.n_samples, len, frames);
        for (int s = t->first_stage; s <= t->last_stage; ++s) bytes += align256(size_t(len[s]) * sizeof(float) + 16);
    }
    return bytes;
}

int run_chroma_cqt(const ta_plan* plan, const HostBatch& hb, const TrackDesc* d_tracks, const float* mag, const float* frame_max,
                   float* chroma, float* cqt_mag, double* tuning, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    int rc = cqt_supported(plan);
    if (rc != TA_OK) return rc;
    const CqtTables* t = cqt_tables(plan);
    TA_REQUIRE(hb.n_tracks <= 65535, "at most 65535 tracks per call");
    TA_REQUIRE(mag && frame_max && chroma && tuning, "chroma_cqt needs the magnitude, frame_max, chroma_cqt and cqt_tuning buffers");
    TA_REQUIRE(scratch && scratch_bytes >= cqt_scratch_bytes(plan, hb), "cqt scratch too small (ta_cqt_scratch_bytes)");
    TA_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "cqt scratch must be 256-byte aligned");
    unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
    size_t cur = 0;
    auto take = [&](size_t b) {
        unsigned char* r = base + cur;
        cur += align256(b);
        return r;
    };
    const int nt = hb.n_tracks, ns = t->last_stage + 1;
    CqtTrack* d_ct = reinterpret_cast<CqtTrack*>(take(sizeof(CqtTrack) * nt));
    DecTrack* d_dec = reinterpret_cast<DecTrack*>(take(align256(sizeof(DecTrack) * nt) * ns));
    int* d_tidx = reinterpret_cast<int*>(take(sizeof(int) * nt));
    void* d_tune = take(tuning_scratch_bytes(plan, hb));

    std::vector<CqtTrack> ct(nt);
    std::vector<std::vector<DecTrack>> dec(ns, std::vector<DecTrack>(nt));
    long long pitch = 0;
    int64_t max_out[CQ_MAX_STAGES] = {0};
    int max_frames = 0;
    for (int i = 0; i < nt; ++i) {
        const TrackDesc& tr = hb.tracks[i];
        int64_t len[CQ_MAX_STAGES], frames;
        cqt_lengths(t, tr.n_samples, len, frames);
        TA_REQUIRE(frames < (int64_t(1) << 30), "track too long");
        float* sig[CQ_MAX_STAGES] = {nullptr};
        for (int s = t->first_stage; s <= t->last_stage; ++s) sig[s] = reinterpret_cast<float*>(take(size_t(len[s]) * sizeof(float) + 16));
        // stage s is produced from stage s-1 (stage 0 = the mono mix of the PCM itself)
        for (int s = t->first_stage; s <= t->last_stage; ++s) {
            DecTrack& d = dec[s][i];
            if (s <= 1) {
                d.a = tr.ch0;
                d.b = tr.ch1;
                d.n_in = tr.n_samples;
            } else {
                d.a = sig[s - 1];
                d.b = nullptr;
                d.n_in = len[s - 1];
            }
            d.dst = sig[s];
            d.n_out = len[s];
            max_out[s] = std::max(max_out[s], len[s]);
        }
        CqtTrack& c = ct[i];
        for (int o = 0; o < CQ_OCT; ++o) {
            c.sig[o] = sig[t->early + o];
            c.len[o] = len[t->early + o];
        }
        c.n_frames = int(frames);
        c.ld = int(ta_frame_pitch(frames));
        c.pitch_off = pitch;
        pitch += c.ld;
        max_frames = std::max(max_frames, c.n_frames);
    }
    TA_REQUIRE(cur <= scratch_bytes, "cqt scratch too small (ta_cqt_scratch_bytes)");
    TA_CUDA(cudaMemcpyAsync(d_ct, ct.data(), sizeof(CqtTrack) * nt, cudaMemcpyHostToDevice, stream));
    const size_t dec_stride = align256(sizeof(DecTrack) * nt);
    for (int s = t->first_stage; s <= t->last_stage; ++s)
        TA_CUDA(cudaMemcpyAsync(reinterpret_cast<unsigned char*>(d_dec) + dec_stride * s, dec[s].data(), sizeof(DecTrack) * nt,
                                cudaMemcpyHostToDevice, stream));
    // tuning from the magnitude spectrogram (bins per octave 36, magnitude not power)
    rc = run_tuning(plan, hb, d_tracks, mag, frame_max, false, CQ_BPO, d_tune, tuning, d_tidx, stream);
    if (rc != TA_OK) return rc;
    // decimation chain
    for (int s = t->first_stage; s <= t->last_stage; ++s) {
        const DecTrack* d = reinterpret_cast<const DecTrack*>(reinterpret_cast<unsigned char*>(d_dec) + dec_stride * s);
        if (max_out[s] <= 0) continue;
        if (s == 0) {
            cqt_mono_kernel<<<dim3(unsigned(std::min<int64_t>((max_out[s] + 255) / 256, 1024)), nt), 256, 0, stream>>>(d);
        } else {
            TA_REQUIRE((max_out[s] + DEC_TILE - 1) / DEC_TILE < (int64_t(1) << 31), "track too long");
            cqt_decimate_kernel<<<dim3(unsigned((max_out[s] + DEC_TILE - 1) / DEC_TILE), nt), DEC_THREADS, 0, stream>>>(d);
        }
        count_launch();
        TA_CUDA(cudaGetLastError());
    }
    // transform + chroma
    CqtParams p{};
    p.tracks = d_ct;
    p.tuning_idx = d_tidx;
    p.basis = t->d_basis;
    p.span = t->d_span;
    p.inv_sqrt_len = t->d_inv_sqrt_len;
    p.tw1 = t->d_tw1;
    p.tw2 = t->d_tw2;
    p.nfft = t->nfft;
    p.bin_step = t->bin_step;
    p.hop0 = CQ_HOP >> t->early;
    p.j_lo = t->j_lo;
    p.n_rows = t->j_hi - t->j_lo;
    p.chroma = chroma;
    p.cqt_mag = cqt_mag;
    using E = p2::Ex<CQ_N>;
    const size_t smem = size_t(512 / (CQ_N / 16)) * E::SLOTS * 16 + size_t(15) * (CQ_N / 16) * 8 + size_t(16) * (CQ_N / 256) * 8 +
                        size_t(p.n_rows) * CQ_SPITCH * 8 + size_t(CQ_BPO) * CQ_TF * 4 + size_t(12) * CQ_TF * 4;
    TA_REQUIRE(smem <= 232448, "constant-Q spectrum tile does not fit in shared memory");
    const int tiles = (max_frames + CQ_TF - 1) / CQ_TF;
    const int gx = std::max(1, std::min(tiles, std::max(1, (4 * plan->sm_count + nt - 1) / nt)));
    auto kern = t->bin_step == 1 ? cqt_chroma_kernel<1> : (t->bin_step == 2 ? cqt_chroma_kernel<2> : cqt_chroma_kernel<4>);
    TA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(gx, nt), 512, smem, stream>>>(p);
    count_launch();
    TA_CUDA(cudaGetLastError());
    return TA_OK;
}

}  // namespace ta
