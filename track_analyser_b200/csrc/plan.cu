// Plan construction (host-side tables computed in double), error plumbing, batch
// descriptors, workspace carve-up and the fused schedule of the C ABI.
#include <atomic>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <numeric>

#include "common.cuh"

namespace ta {

static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(uint64_t(n), std::memory_order_relaxed); }

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " at " + file + ":" +
                   std::to_string(line);
    return TA_ERR_CUDA;
}

// stage launchers defined in the other translation units
int run_stft_features(const ta_plan*, const HostBatch&, const Workspace&, const ta_frontend_out*, cudaStream_t);
int run_onset_flux(const ta_plan*, const HostBatch&, const TrackDesc*, const float* mel, const uint32_t* mel_max,
                   float* onset_env, double* flux_linear, cudaStream_t);
int run_autocorrelate(const ta_plan*, const HostBatch&, const TrackDesc*, const float* env, double* out,
                      double* scratch, size_t scratch_elems, cudaStream_t);
size_t autocorr_scratch_elems(const HostBatch&);
int run_time_domain(const ta_plan*, const HostBatch&, const Workspace&, const ta_frontend_out*, cudaStream_t);
int time_chunk_samples(const ta_plan*, int64_t total_samples);
void true_peak_design(int up, float* coef, float& gain, float& floor_);

// ---------------------------------------------------------------------------
// Slaney mel scale exactly as librosa.filters.mel(htk=False, norm="slaney")
// ---------------------------------------------------------------------------
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static void build_mel(const ta_plan_desc& d, int n_bins, const std::vector<double>& fftfreqs,
                      std::vector<float>& dense) {
    const int M = d.n_mels;
    const double fmax = d.fmax > 0 ? double(d.fmax) : double(d.sample_rate) / 2;
    const double lo = hz_to_mel(double(d.fmin)), hi = hz_to_mel(fmax);
    std::vector<double> mel_f(M + 2);
    for (int i = 0; i < M + 2; ++i) {
        // numpy.linspace: start + i*step with step = (stop-start)/(num-1); last point is stop exactly
        const double step = (hi - lo) / double(M + 1);
        const double m = (i == M + 1) ? hi : lo + i * step;
        mel_f[i] = mel_to_hz(m);
    }
    dense.assign(size_t(M) * n_bins, 0.f);
    for (int i = 0; i < M; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < n_bins; ++k) {
            const double lower = -(mel_f[i] - fftfreqs[k]) / fd0;
            const double upper = (mel_f[i + 2] - fftfreqs[k]) / fd1;
            const double w = std::max(0.0, std::min(lower, upper));
            // librosa stores float32 weights, then scales in place by the float64 enorm
            dense[size_t(i) * n_bins + k] = float(double(float(w)) * enorm);
        }
    }
}

static void biquad_kweight(int rate, Biquad& shelf, Biquad& hp, double coefs[12]) {
    const double PI = 3.14159265358979323846;
    {
        const double G = 4.0, Q = 1.0 / std::sqrt(2.0), fc = 1500.0;
        const double A = std::pow(10.0, G / 40.0);
        const double w0 = 2.0 * PI * (fc / rate);
        const double alpha = std::sin(w0) / (2.0 * Q);
        const double c = std::cos(w0), sA = std::sqrt(A);
        const double b0 = A * ((A + 1) + (A - 1) * c + 2 * sA * alpha);
        const double b1 = -2 * A * ((A - 1) + (A + 1) * c);
        const double b2 = A * ((A + 1) + (A - 1) * c - 2 * sA * alpha);
        const double a0 = (A + 1) - (A - 1) * c + 2 * sA * alpha;
        const double a1 = 2 * ((A - 1) - (A + 1) * c);
        const double a2 = (A + 1) - (A - 1) * c - 2 * sA * alpha;
        shelf = {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
        const double t[6] = {b0 / a0, b1 / a0, b2 / a0, a0 / a0, a1 / a0, a2 / a0};
        std::memcpy(coefs, t, sizeof(t));
    }
    {
        const double Q = 0.5, fc = 38.0;
        const double w0 = 2.0 * PI * (fc / rate);
        const double alpha = std::sin(w0) / (2.0 * Q);
        const double c = std::cos(w0);
        const double b0 = (1 + c) / 2, b1 = -(1 + c), b2 = (1 + c) / 2;
        const double a0 = 1 + alpha, a1 = -2 * c, a2 = 1 - alpha;
        hp = {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
        const double t[6] = {b0 / a0, b1 / a0, b2 / a0, a0 / a0, a1 / a0, a2 / a0};
        std::memcpy(coefs + 6, t, sizeof(t));
    }
}

template <typename T>
static int upload(T** dptr, const std::vector<T>& h) {
    TA_CUDA(cudaMalloc(reinterpret_cast<void**>(dptr), std::max<size_t>(1, h.size()) * sizeof(T)));
    if (!h.empty()) TA_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return TA_OK;
}

static int plan_build(ta_plan* p, const double* user_window) {
    const ta_plan_desc& d = p->desc;
    const int NF = d.n_fft;                      // the plan's n_fft: window length, bins
    const int N = stft_transform_length(NF);    // length of the transform that evaluates it (NF zero-padded when NF < 1024)
    const int M = N / 16, Q = N / 256;
    const double PI = 3.14159265358979323846;
    TA_CUDA(cudaSetDevice(d.device));
    cudaDeviceProp prop;
    TA_CUDA(cudaGetDeviceProperties(&prop, d.device));
    p->sm_count = prop.multiProcessorCount;
    p->n_bins = NF / 2 + 1;

    // twiddles, exact to float rounding (angle reduced with integer arithmetic first)
    std::vector<float2> tw1(size_t(15) * M), tw2(size_t(16) * Q);
    for (int k1 = 1; k1 < 16; ++k1)
        for (int r = 0; r < M; ++r) {
            const double a = -2.0 * PI * double((r * k1) % N) / N;
            tw1[size_t(k1 - 1) * M + r] = make_float2(float(std::cos(a)), float(std::sin(a)));
        }
    for (int k2 = 0; k2 < 16; ++k2)
        for (int n3 = 0; n3 < Q; ++n3) {
            const double a = -2.0 * PI * double(n3 * k2) / M;
            tw2[size_t(k2) * Q + n3] = make_float2(float(std::cos(a)), float(std::sin(a)));
        }
    // periodic Hann, scipy.signal.get_window("hann", N, fftbins=True): 0.5 - 0.5 cos(2 pi n / N)
    p->h_window.assign(N, 0.f);                  // zero beyond the frame when the transform is longer than it
    for (int n = 0; n < NF; ++n)
        p->h_window[n] = user_window ? float(user_window[n]) : float(0.5 - 0.5 * std::cos(2.0 * PI * n / NF));
    // numpy.fft.rfftfreq(n, d=1/sr): arange(n//2+1) * (1.0 / (n * d))
    p->h_freqs.resize(p->n_bins);
    {
        const double dd = 1.0 / double(d.sample_rate);
        const double val = 1.0 / (double(NF) * dd);
        for (int k = 0; k < p->n_bins; ++k) p->h_freqs[k] = double(k) * val;
    }
    int rc;
    if ((rc = upload(&p->d_tw1, tw1))) return rc;
    if ((rc = upload(&p->d_tw2, tw2))) return rc;
    if ((rc = upload(&p->d_window, p->h_window))) return rc;
    if ((rc = upload(&p->d_freqs, p->h_freqs))) return rc;

    if (d.n_mels > 0) {
        build_mel(d, p->n_bins, p->h_freqs, p->h_mel_dense);
        std::vector<int> start(d.n_mels), len(d.n_mels), woff(d.n_mels);
        std::vector<float> w;
        for (int m = 0; m < d.n_mels; ++m) {
            int lo = p->n_bins, hi = -1;
            for (int k = 0; k < p->n_bins; ++k)
                if (p->h_mel_dense[size_t(m) * p->n_bins + k] != 0.f) {
                    lo = std::min(lo, k);
                    hi = std::max(hi, k);
                }
            start[m] = (hi < 0) ? 0 : lo;
            len[m] = (hi < 0) ? 0 : hi - lo + 1;
            woff[m] = int(w.size());
            for (int k = 0; k < len[m]; ++k) w.push_back(p->h_mel_dense[size_t(m) * p->n_bins + start[m] + k]);
        }
        if ((rc = upload(&p->d_mel_start, start))) return rc;
        if ((rc = upload(&p->d_mel_len, len))) return rc;
        if ((rc = upload(&p->d_mel_woff, woff))) return rc;
        if ((rc = upload(&p->d_mel_w, w))) return rc;
        // scipy.fft.dct(type=2, norm="ortho") rows 0..12 over the mel axis (librosa.feature.mfcc)
        const int M = p->desc.n_mels;
        std::vector<double> dct(size_t(TA_N_MFCC) * M);
        const double pi = 3.14159265358979323846;
        for (int k = 0; k < TA_N_MFCC; ++k)
            for (int m = 0; m < M; ++m)
                dct[size_t(k) * M + m] = k == 0 ? 1.0 / std::sqrt(double(M))
                                                : std::sqrt(2.0 / M) * std::cos(pi * k * (2 * m + 1) / (2.0 * M));
        if ((rc = upload(&p->d_dct, dct))) return rc;
        p->mel_nnz = int(w.size());
    }

    {   // tempogram: 1024-point transform twiddles and periodic Hann(win)
        const int TN = 1024, TM = TN / 16, TQ = TN / 256;
        std::vector<float2> t1(size_t(15) * TM), t2(size_t(16) * TQ);
        for (int k1 = 1; k1 < 16; ++k1)
            for (int r = 0; r < TM; ++r) {
                const double a = -2.0 * PI * double((r * k1) % TN) / TN;
                t1[size_t(k1 - 1) * TM + r] = make_float2(float(std::cos(a)), float(std::sin(a)));
            }
        for (int k2 = 0; k2 < 16; ++k2)
            for (int n3 = 0; n3 < TQ; ++n3) {
                const double a = -2.0 * PI * double(n3 * k2) / TM;
                t2[size_t(k2) * TQ + n3] = make_float2(float(std::cos(a)), float(std::sin(a)));
            }
        const int win = std::max(2, d.tempogram_win);
        std::vector<float> tw(win);
        for (int n = 0; n < win; ++n) tw[n] = float(0.5 - 0.5 * std::cos(2.0 * PI * n / win));
        if ((rc = upload(&p->d_tg_tw1, t1))) return rc;
        if ((rc = upload(&p->d_tg_tw2, t2))) return rc;
        if ((rc = upload(&p->d_tg_window, tw))) return rc;
    }

    double coefs[12];
    biquad_kweight(d.sample_rate, p->shelf, p->highpass, coefs);

    true_peak_design(8, p->tp_coef, p->tp_gain, p->tp_floor);  // the reference's default oversample (loudness.py:81)

    // loudness framing (analysis/loudness.py:35-38 and pyloudnorm block bounds)
    auto rms_frame = [&](double seconds) {
        int fl = std::max(1024, int(std::nearbyint(double(d.sample_rate) * seconds)));
        if (fl % 2) fl += 1;
        return fl;
    };
    p->rms_m_frame = rms_frame(double(d.meter_block));
    p->rms_m_hop = std::max(1, p->rms_m_frame / 2);
    p->rms_s_frame = rms_frame(3.0);
    p->rms_s_hop = std::max(1, p->rms_s_frame / 2);
    return TA_OK;
}

// pyloudnorm block bounds, evaluated with the identical double expressions
void kw_block_bounds(const ta_plan* plan, int64_t n_samples, std::vector<int64_t>& lo, std::vector<int64_t>& hi) {
    const double T_g = double(plan->desc.meter_block), step = 0.25;
    const double rate = double(plan->desc.sample_rate);
    const double T = double(n_samples) / rate;
    lo.clear();
    hi.clear();
    if (double(n_samples) < T_g * rate) return;
    const long long nb = (long long)(std::nearbyint((T - T_g) / (T_g * step)) + 1);
    for (long long j = 0; j < nb; ++j) {
        lo.push_back((int64_t)(T_g * (double(j) * step) * rate));
        hi.push_back((int64_t)(T_g * (double(j) * step + 1) * rate));
    }
}

int build_host_batch(const ta_plan* plan, const ta_batch* b, HostBatch& hb) {
    TA_REQUIRE(b && b->n_tracks > 0, "batch must hold at least one track");
    TA_REQUIRE(b->channels == 1 || b->channels == 2, "channels must be 1 or 2");
    TA_REQUIRE(b->pcm && b->pcm_offset && b->n_samples, "batch pointers must not be NULL");
    const int hop = plan->desc.hop, TF = stft_tile_frames(plan->desc.n_fft);
    hb.n_tracks = b->n_tracks;
    hb.channels = b->channels;
    hb.tracks.resize(b->n_tracks);
    int64_t pitch = 0, samples = 0;
    long long tiles = 0;
    for (int i = 0; i < b->n_tracks; ++i) {
        const int64_t ns = b->n_samples[i];
        TA_REQUIRE(ns >= 0, "n_samples must be >= 0");
        TA_REQUIRE(b->pcm_offset[i] % 4 == 0, "pcm_offset must be a multiple of 4 elements");
        TrackDesc& t = hb.tracks[i];
        t.ch0 = b->pcm + b->pcm_offset[i];
        t.ch1 = (b->channels == 2) ? t.ch0 + ns : nullptr;
        t.n_samples = ns;
        const int64_t T = ta_frame_count(ns, hop);
        TA_REQUIRE(T < (int64_t(1) << 30), "track too long");
        t.n_frames = int(T);
        t.ld = int(ta_frame_pitch(T));
        t.pitch_off = pitch;
        t.tile_begin = int(tiles);
        t.chunk_begin = 0;  // filled by the time-domain stage
        pitch += t.ld;
        samples += ns;
        tiles += (T + TF - 1) / TF;
        hb.max_frames = std::max(hb.max_frames, t.n_frames);
        TA_REQUIRE(tiles < (1ll << 31), "batch too large");
    }
    hb.total_pitch = pitch;
    hb.total_samples = samples;
    hb.total_tiles = int(tiles);
    // time-domain chunks
    const int cs = time_chunk_samples(plan, samples);
    long long chunks = 0;
    for (int i = 0; i < b->n_tracks; ++i) {
        hb.tracks[i].chunk_begin = int(chunks);
        chunks += (hb.tracks[i].n_samples + cs - 1) / cs;
    }
    TA_REQUIRE(chunks < (1ll << 31), "batch too large");
    hb.total_chunks = int(chunks);
    return TA_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t td_granule_doubles(const ta_plan* plan, const HostBatch& hb);
size_t chroma_scratch_bytes(const ta_plan* plan, const HostBatch& hb);
int run_chroma(const ta_plan*, const HostBatch&, const TrackDesc*, const float* mag, const float* frame_max, float* chroma,
               double* tuning, int32_t* rolloff_bin, const float* frame_sum, void* scratch, size_t scratch_bytes, void* d_maps, cudaStream_t);
int run_tempogram(const ta_plan*, const HostBatch&, const TrackDesc*, const float* env, float* out, cudaStream_t);
int run_mfcc(const ta_plan*, const HostBatch&, const TrackDesc*, const float* mel, const uint32_t* mel_max, double* mfcc,
             cudaStream_t);
int run_hpss(const ta_plan*, const HostBatch&, const TrackDesc*, const float* mag, float* scratch, float* harm_sum, float* perc_sum,
             cudaStream_t, float* harm_full = nullptr, float* perc_full = nullptr);
int run_self_similarity(const ta_plan*, const HostBatch&, const TrackDesc*, const double* mfcc, double* scratch, double* out,
                        cudaStream_t);

int cqt_supported(const ta_plan* plan);
int64_t cqt_frame_count(const ta_plan* plan, int64_t n_samples);
size_t cqt_scratch_bytes(const ta_plan* plan, const HostBatch& hb);
void cqt_tables_free(CqtTables* t);
int run_chroma_cqt(const ta_plan*, const HostBatch&, const TrackDesc*, const float* mag, const float* frame_max, float* chroma,
                   float* cqt_mag, double* tuning, void* scratch, size_t scratch_bytes, cudaStream_t);

size_t carve_workspace(const ta_plan* plan, const HostBatch& hb, void* base, Workspace& ws) {
    unsigned char* p = reinterpret_cast<unsigned char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char* r = p ? p + off : nullptr;
        off = align_up(off + bytes, 256);
        return r;
    };
    ws.d_tracks = reinterpret_cast<TrackDesc*>(take(sizeof(TrackDesc) * hb.n_tracks));
    ws.d_mel_max = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * hb.n_tracks));
    ws.d_tmaps = take(size_t(128) * hb.n_tracks);
    ws.d_frame_sum = reinterpret_cast<float*>(take(sizeof(float) * size_t(hb.total_pitch)));
    ws.d_novelty = reinterpret_cast<double*>(take(plan->desc.n_mels > 0 ? sizeof(double) * 2 * TA_N_MFCC * size_t(hb.total_pitch) : 0));
    // granule sums: K-weighted, momentary hop, short-term hop
    ws.gran_doubles = td_granule_doubles(plan, hb);
    ws.d_granules = reinterpret_cast<double*>(take(sizeof(double) * ws.gran_doubles));
    ws.fft_elems = autocorr_scratch_elems(hb);
    ws.d_fft = reinterpret_cast<double*>(take(sizeof(double) * ws.fft_elems));
    ws.chroma_bytes = chroma_scratch_bytes(plan, hb);
    ws.d_chroma = take(ws.chroma_bytes);
    {
        int64_t max_samples = 0;
        for (auto& t : hb.tracks) max_samples = std::max(max_samples, t.n_samples);
        ws.blk_pitch = int(max_samples / 256 + 2);
        ws.d_blk_absmax = reinterpret_cast<float*>(take(sizeof(float) * size_t(hb.n_tracks) * ws.blk_pitch));
        ws.d_absmax_bits = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * hb.n_tracks));
        ws.d_tp_begin = take(sizeof(long long) * hb.n_tracks);
    }
    ws.end = p ? p + off : nullptr;
    return off;
}

}  // namespace ta

using namespace ta;

extern "C" {

int ta_abi_version(void) { return TA_ABI_VERSION; }
const char* ta_last_error(void) { return g_last_error.c_str(); }

int ta_plan_create(const ta_plan_desc* desc, ta_plan** out) { return ta_plan_create_window(desc, nullptr, out); }

int ta_plan_create_window(const ta_plan_desc* desc, const double* window, ta_plan** out) {
    TA_REQUIRE(desc && out, "desc/out must not be NULL");
    *out = nullptr;
    TA_REQUIRE(desc->n_fft == 256 || desc->n_fft == 512 || desc->n_fft == 1024 || desc->n_fft == 2048 || desc->n_fft == 4096,
               "n_fft must be a power of two between 256 and 4096");
    TA_REQUIRE(desc->hop > 0, "hop must be positive");
    TA_REQUIRE(desc->sample_rate > 0, "sample_rate must be positive");
    TA_REQUIRE(desc->n_mels >= 0 && desc->n_mels <= 1024, "n_mels out of range");
    TA_REQUIRE(desc->roll_percent > 0.0 && desc->roll_percent < 1.0, "roll_percent must be in (0,1)");
    TA_REQUIRE(desc->meter_block > 0.0, "meter_block must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error(std::string("no CUDA device available (no CPU fallback): ") + cudaGetErrorString(e));
        return TA_ERR_CUDA;
    }
    TA_REQUIRE(desc->device >= 0 && desc->device < ndev, "device ordinal out of range");
    ta_plan* p = new (std::nothrow) ta_plan();
    TA_REQUIRE(p, "out of host memory");
    p->desc = *desc;
    int rc = plan_build(p, window);
    if (rc == TA_OK && cudaStreamCreateWithFlags(&p->aux_stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreateWithFlags failed");
        rc = TA_ERR_CUDA;
    }
    if (rc != TA_OK) {
        ta_plan_destroy(p);
        return rc;
    }
    *out = p;
    return TA_OK;
}

void ta_plan_destroy(ta_plan* p) {
    if (!p) return;
    cudaFree(p->d_tw1);
    cudaFree(p->d_tw2);
    cudaFree(p->d_window);
    cudaFree(p->d_freqs);
    cudaFree(p->d_mel_start);
    cudaFree(p->d_mel_len);
    cudaFree(p->d_mel_woff);
    if (p->aux_stream) cudaStreamDestroy(p->aux_stream);
    cudaFree(p->d_mel_w);
    cudaFree(p->d_dct);
    cudaFree(p->d_tg_tw1);
    cudaFree(p->d_tg_tw2);
    cudaFree(p->d_tg_window);
    cqt_tables_free(p->cqt);
    delete p;
}

int ta_plan_n_bins(const ta_plan* plan) { return plan ? plan->n_bins : TA_ERR_INVALID; }

int ta_plan_table(const ta_plan* plan, int which, void* host_out, size_t bytes) {
    TA_REQUIRE(plan && host_out, "plan/host_out must not be NULL");
    const void* src = nullptr;
    size_t need = 0;
    double coefs[12];
    switch (which) {
        case 0: src = plan->h_window.data(); need = size_t(plan->desc.n_fft) * sizeof(float); break;
        case 1: src = plan->h_mel_dense.data(); need = plan->h_mel_dense.size() * sizeof(float); break;
        case 2: src = plan->h_freqs.data(); need = plan->h_freqs.size() * sizeof(double); break;
        case 3: {
            Biquad s, h;
            biquad_kweight(plan->desc.sample_rate, s, h, coefs);
            src = coefs;
            need = sizeof(coefs);
            break;
        }
        default: set_error("unknown table id"); return TA_ERR_INVALID;
    }
    TA_REQUIRE(bytes >= need, "host_out too small");
    std::memcpy(host_out, src, need);
    return TA_OK;
}

size_t ta_workspace_bytes(const ta_plan* plan, const ta_batch* batch) {
    if (!plan || !batch) return 0;
    HostBatch hb;
    if (build_host_batch(plan, batch, hb) != TA_OK) return 0;
    Workspace ws;
    return carve_workspace(plan, hb, nullptr, ws);
}

static int prepare(const ta_plan* plan, const ta_batch* batch, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream, HostBatch& hb, Workspace& ws) {
    TA_REQUIRE(plan, "plan must not be NULL");
    int rc = build_host_batch(plan, batch, hb);
    if (rc != TA_OK) return rc;
    const size_t need = carve_workspace(plan, hb, workspace, ws);
    if (!workspace || workspace_bytes < need) {
        set_error("workspace too small: need " + std::to_string(need) + " bytes");
        return TA_ERR_WORKSPACE;
    }
    TA_CUDA(cudaSetDevice(plan->desc.device));
    // descriptor upload: stream-ordered copy from a pageable host vector is staged by the runtime
    TA_CUDA(cudaMemcpyAsync(ws.d_tracks, hb.tracks.data(), sizeof(TrackDesc) * hb.n_tracks, cudaMemcpyHostToDevice, stream));
    return TA_OK;
}

int ta_stft_features(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out, void* workspace,
                     size_t workspace_bytes, void* stream) {
    TA_REQUIRE(out, "out must not be NULL");
    HostBatch hb;
    Workspace ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = prepare(plan, batch, workspace, workspace_bytes, st, hb, ws);
    if (rc != TA_OK) return rc;
    TA_REQUIRE(!out->rolloff_bin || out->magnitude, "rolloff_bin output needs the magnitude buffer");
    if ((rc = run_stft_features(plan, hb, ws, out, st)) != TA_OK) return rc;
    // the roll-off bins come from a sequential walk down the magnitude columns (chroma.cu)
    if (out->rolloff_bin)
        return run_chroma(plan, hb, ws.d_tracks, out->magnitude, nullptr, nullptr, nullptr, out->rolloff_bin, ws.d_frame_sum, nullptr, 0, ws.d_tmaps, st);
    return TA_OK;
}

int ta_onset_flux(const ta_plan* plan, const ta_batch* batch, const float* mel, const uint32_t* mel_max_bits,
                  float* onset_env, double* flux_linear, void* stream) {
    TA_REQUIRE(plan && mel && mel_max_bits, "plan/mel/mel_max_bits must not be NULL");
    HostBatch hb;
    int rc = build_host_batch(plan, batch, hb);
    if (rc != TA_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    TA_CUDA(cudaSetDevice(plan->desc.device));
    // this stand-alone entry point has no workspace: descriptors travel through a stream-ordered allocation
    TrackDesc* d_tracks = nullptr;
    TA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_tracks), sizeof(TrackDesc) * hb.n_tracks, st));
    TA_CUDA(cudaMemcpyAsync(d_tracks, hb.tracks.data(), sizeof(TrackDesc) * hb.n_tracks, cudaMemcpyHostToDevice, st));
    rc = run_onset_flux(plan, hb, d_tracks, mel, mel_max_bits, onset_env, flux_linear, st);
    cudaFreeAsync(d_tracks, st);
    return rc;
}

int ta_autocorrelate(const ta_plan* plan, const ta_batch* batch, const float* onset_env, double* autocorr,
                     void* workspace, size_t workspace_bytes, void* stream) {
    TA_REQUIRE(onset_env && autocorr, "onset_env/autocorr must not be NULL");
    HostBatch hb;
    Workspace ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = prepare(plan, batch, workspace, workspace_bytes, st, hb, ws);
    if (rc != TA_OK) return rc;
    return run_autocorrelate(plan, hb, ws.d_tracks, onset_env, autocorr, ws.d_fft, ws.fft_elems, st);
}

int ta_time_domain(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out, void* workspace,
                   size_t workspace_bytes, void* stream) {
    TA_REQUIRE(out, "out must not be NULL");
    HostBatch hb;
    Workspace ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = prepare(plan, batch, workspace, workspace_bytes, st, hb, ws);
    if (rc != TA_OK) return rc;
    return run_time_domain(plan, hb, ws, out, st);
}

int ta_chroma_stft(const ta_plan* plan, const ta_batch* batch, const float* magnitude, const float* frame_max, float* chroma,
                   double* tuning, void* workspace, size_t workspace_bytes, void* stream) {
    TA_REQUIRE(magnitude && frame_max && chroma && tuning, "magnitude/frame_max/chroma/tuning must not be NULL");
    HostBatch hb;
    Workspace ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = prepare(plan, batch, workspace, workspace_bytes, st, hb, ws);
    if (rc != TA_OK) return rc;
    return run_chroma(plan, hb, ws.d_tracks, magnitude, frame_max, chroma, tuning, nullptr, nullptr, ws.d_chroma, ws.chroma_bytes, ws.d_tmaps, st);
}

int ta_tempogram(const ta_plan* plan, const ta_batch* batch, const float* onset_env, float* tempogram, void* workspace,
                 size_t workspace_bytes, void* stream) {
    TA_REQUIRE(onset_env && tempogram, "onset_env/tempogram must not be NULL");
    HostBatch hb;
    Workspace ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = prepare(plan, batch, workspace, workspace_bytes, st, hb, ws);
    if (rc != TA_OK) return rc;
    return run_tempogram(plan, hb, ws.d_tracks, onset_env, tempogram, st);
}

int ta_hpss_curves(const ta_plan* plan, const ta_batch* batch, const float* magnitude, float* scratch, float* harmonic_sum,
                   float* percussive_sum, void* workspace, size_t workspace_bytes, void* stream) {
    HostBatch hb;
    Workspace ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = prepare(plan, batch, workspace, workspace_bytes, st, hb, ws);
    if (rc != TA_OK) return rc;
    return run_hpss(plan, hb, ws.d_tracks, magnitude, scratch, harmonic_sum, percussive_sum, st);
}

int64_t ta_cqt_frame_count(const ta_plan* plan, int64_t n_samples) {
    if (!plan || n_samples < 0) {
        set_error("plan must not be NULL and n_samples must be >= 0");
        return TA_ERR_INVALID;
    }
    if (cqt_supported(plan) != TA_OK) return TA_ERR_UNSUPPORTED;
    return cqt_frame_count(plan, n_samples);
}

size_t ta_cqt_scratch_bytes(const ta_plan* plan, const ta_batch* batch) {
    if (!plan || !batch) return 0;
    HostBatch hb;
    if (build_host_batch(plan, batch, hb) != TA_OK) return 0;
    if (cqt_supported(plan) != TA_OK) return 0;
    return cqt_scratch_bytes(plan, hb);
}

int ta_chroma_cqt(const ta_plan* plan, const ta_batch* batch, const float* magnitude, const float* frame_max, float* chroma_cqt,
                  float* cqt_mag, double* cqt_tuning, void* cqt_scratch, size_t cqt_scratch_bytes_, void* workspace,
                  size_t workspace_bytes, void* stream) {
    HostBatch hb;
    Workspace ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = prepare(plan, batch, workspace, workspace_bytes, st, hb, ws);
    if (rc != TA_OK) return rc;
    return run_chroma_cqt(plan, hb, ws.d_tracks, magnitude, frame_max, chroma_cqt, cqt_mag, cqt_tuning, cqt_scratch,
                          cqt_scratch_bytes_, st);
}

int ta_hpss_components(const ta_plan* plan, const ta_batch* batch, const float* magnitude, float* scratch, float* harmonic,
                       float* percussive, float* harmonic_sum, float* percussive_sum, void* workspace, size_t workspace_bytes,
                       void* stream) {
    TA_REQUIRE(harmonic && percussive, "harmonic / percussive must not be NULL");
    HostBatch hb;
    Workspace ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = prepare(plan, batch, workspace, workspace_bytes, st, hb, ws);
    if (rc != TA_OK) return rc;
    return run_hpss(plan, hb, ws.d_tracks, magnitude, scratch, harmonic_sum, percussive_sum, st, harmonic, percussive);
}

uint64_t ta_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static int frontend_impl(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out, void* workspace,
                         size_t workspace_bytes, void* stream, cudaEvent_t* ev) {
    TA_REQUIRE(out, "out must not be NULL");
    HostBatch hb;
    Workspace ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = prepare(plan, batch, workspace, workspace_bytes, st, hb, ws);
    if (rc != TA_OK) return rc;
    auto mark = [&](int i) { if (ev) cudaEventRecord(ev[i], st); };
    mark(0);
    // Two chains are independent of the main one (K1 -> onset -> autocorrelation -> tempogram): the time-domain pass
    // (K5/K6/K8) depends on nothing but the PCM, and the chroma / HPSS kernels only on K1's magnitude.  Outside the
    // profiled run both go to the plan's second stream -- the time-domain pass right here, chroma/HPSS after an event
    // that marks K1's completion -- so that the HBM-bound projection overlaps the FP64-bound tempogram, and are joined
    // before returning (the caller's stream sees all results in order).  TA_OVERLAP=0: everything on the caller's stream;
    // TA_OVERLAP=1: only the time-domain pass is forked.
    const bool need_td = out->moments || out->kw_blocks || out->lufs || out->rms_momentary || out->rms_short || out->true_peak;
    static const int overlap_mode = [] { const char* e = std::getenv("TA_OVERLAP"); return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2; }();
    const bool fork = !ev && overlap_mode > 0 && plan->aux_stream;
    const bool fork_td = need_td && fork;
    const bool fork_mag = fork && overlap_mode > 1;  // chroma / HPSS on the second stream
    cudaStream_t aux = plan->aux_stream;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_k1 = nullptr;
    if (fork) {
        TA_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        TA_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        TA_CUDA(cudaEventCreateWithFlags(&ev_k1, cudaEventDisableTiming));
        TA_CUDA(cudaEventRecord(ev_fork, st));
        TA_CUDA(cudaStreamWaitEvent(aux, ev_fork, 0));
    }
    auto join = [&]() {
        if (fork) {
            cudaEventRecord(ev_join, aux);
            cudaStreamWaitEvent(st, ev_join, 0);
            cudaEventDestroy(ev_fork);  // destruction is deferred by the runtime until the events have completed
            cudaEventDestroy(ev_join);
            cudaEventDestroy(ev_k1);
        }
    };
    if (fork_td && (rc = run_time_domain(plan, hb, ws, out, aux)) != TA_OK) { join(); return rc; }
    const bool need_flux = out->onset_env || out->flux_linear || out->autocorr || out->tempogram;
    if (need_flux && !out->mel) { join(); set_error("onset/autocorr outputs need the mel output buffer"); return TA_ERR_INVALID; }
    if (out->autocorr && !out->onset_env) { join(); set_error("autocorr output needs the onset_env output buffer"); return TA_ERR_INVALID; }
    if (out->chroma && !(out->magnitude && out->frame_max && out->tuning)) {
        join();
        set_error("chroma output needs the magnitude, frame_max and tuning buffers");
        return TA_ERR_INVALID;
    }
    if (out->tempogram && !out->onset_env) { join(); set_error("tempogram output needs the onset_env buffer"); return TA_ERR_INVALID; }
    if (out->rolloff_bin && !out->magnitude) { join(); set_error("rolloff_bin output needs the magnitude buffer"); return TA_ERR_INVALID; }
    if (out->chroma_cqt || out->cqt_mag) {
        if (!(out->chroma_cqt && out->magnitude && out->frame_max && out->cqt_tuning && out->cqt_scratch)) {
            join();
            set_error("chroma_cqt output needs the magnitude, frame_max, cqt_tuning and cqt_scratch buffers");
            return TA_ERR_INVALID;
        }
        if ((rc = cqt_supported(plan)) != TA_OK) { join(); return rc; }
    }
    const bool need_stft = out->magnitude || out->mel || out->ltas || out->centroid || out->rolloff_bin || out->band_energy ||
                           out->frame_max;
    if (need_stft && (rc = run_stft_features(plan, hb, ws, out, st)) != TA_OK) { join(); return rc; }
    mark(1);
    if ((out->hpss_harmonic || out->hpss_percussive) &&
        !(out->magnitude && out->hpss_scratch && out->hpss_harmonic && out->hpss_percussive)) {
        join();
        set_error("hpss outputs need the magnitude buffer, hpss_scratch and both sum buffers");
        return TA_ERR_INVALID;
    }
    bool chroma_done = false;
    const bool need_proj = out->chroma || out->rolloff_bin;  // the walk down the magnitude columns: chroma and / or roll-off
    if (fork_mag && (need_proj || out->hpss_harmonic || out->chroma_cqt)) {  // magnitude consumers on the second stream, behind K1
        cudaEventRecord(ev_k1, st);
        cudaStreamWaitEvent(aux, ev_k1, 0);
        if (need_proj && (rc = run_chroma(plan, hb, ws.d_tracks, out->magnitude, out->frame_max, out->chroma, out->tuning,
                                          out->rolloff_bin, ws.d_frame_sum, ws.d_chroma, ws.chroma_bytes, ws.d_tmaps, aux)) != TA_OK) {
            join();
            return rc;
        }
        if (out->hpss_harmonic && (rc = run_hpss(plan, hb, ws.d_tracks, out->magnitude, out->hpss_scratch, out->hpss_harmonic,
                                                 out->hpss_percussive, aux)) != TA_OK) {
            join();
            return rc;
        }
        if (out->chroma_cqt && (rc = run_chroma_cqt(plan, hb, ws.d_tracks, out->magnitude, out->frame_max, out->chroma_cqt,
                                                    out->cqt_mag, out->cqt_tuning, out->cqt_scratch,
                                                    size_t(out->cqt_scratch_bytes), aux)) != TA_OK) {
            join();
            return rc;
        }
        chroma_done = true;
    }
    if (need_flux && (rc = run_onset_flux(plan, hb, ws.d_tracks, out->mel, ws.d_mel_max, out->onset_env,
                                          out->flux_linear, st)) != TA_OK) {
        join();
        return rc;
    }
    if (out->mfcc) {
        if (!out->mel) { join(); set_error("mfcc output needs the mel output buffer"); return TA_ERR_INVALID; }
        if ((rc = run_mfcc(plan, hb, ws.d_tracks, out->mel, ws.d_mel_max, out->mfcc, st)) != TA_OK) { join(); return rc; }
    }
    if (out->self_similarity) {
        if (!out->mfcc) { join(); set_error("self_similarity output needs the mfcc output buffer"); return TA_ERR_INVALID; }
        if ((rc = run_self_similarity(plan, hb, ws.d_tracks, out->mfcc, ws.d_novelty, out->self_similarity, st)) != TA_OK) {
            join();
            return rc;
        }
    }
    mark(2);
    if (out->autocorr &&
        (rc = run_autocorrelate(plan, hb, ws.d_tracks, out->onset_env, out->autocorr, ws.d_fft, ws.fft_elems, st)) != TA_OK) {
        join();
        return rc;
    }
    mark(3);
    if (out->tempogram && (rc = run_tempogram(plan, hb, ws.d_tracks, out->onset_env, out->tempogram, st)) != TA_OK) { join(); return rc; }
    mark(4);
    if (need_proj && !chroma_done &&
        (rc = run_chroma(plan, hb, ws.d_tracks, out->magnitude, out->frame_max, out->chroma, out->tuning, out->rolloff_bin,
                         ws.d_frame_sum, ws.d_chroma, ws.chroma_bytes, ws.d_tmaps, st)) != TA_OK) {
        join();
        return rc;
    }
    if ((out->hpss_harmonic || out->hpss_percussive) && !chroma_done &&
        (rc = run_hpss(plan, hb, ws.d_tracks, out->magnitude, out->hpss_scratch, out->hpss_harmonic, out->hpss_percussive,
                       st)) != TA_OK) {
        join();
        return rc;
    }
    if (out->chroma_cqt && !chroma_done &&
        (rc = run_chroma_cqt(plan, hb, ws.d_tracks, out->magnitude, out->frame_max, out->chroma_cqt, out->cqt_mag,
                             out->cqt_tuning, out->cqt_scratch, size_t(out->cqt_scratch_bytes), st)) != TA_OK) {
        join();
        return rc;
    }
    mark(5);
    if (need_td && !fork_td && (rc = run_time_domain(plan, hb, ws, out, st)) != TA_OK) return rc;
    mark(6);
    join();
    return TA_OK;
}

int ta_frontend_run(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out, void* workspace,
                    size_t workspace_bytes, void* stream) {
    return frontend_impl(plan, batch, out, workspace, workspace_bytes, stream, nullptr);
}

// ---- host-buffer entry point ------------------------------------------------------------------------------------
namespace {
struct DevAllocs {   // stream-ordered allocations of one ta_frontend_run_host call, released on every exit path
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit DevAllocs(cudaStream_t s) : st(s) {}
    void* get(size_t bytes) {
        void* p = nullptr;
        if (cudaMallocAsync(&p, std::max<size_t>(bytes, 16), st) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return p;
    }
    ~DevAllocs() {
        for (void* p : ptrs) cudaFreeAsync(p, st);
    }
};
}  // namespace

int ta_frontend_run_host(const ta_plan* plan, const ta_batch* hbatch, const ta_frontend_out* hout, void* stream) {
    TA_REQUIRE(plan && hbatch && hout, "plan/batch/out must not be NULL");
    TA_REQUIRE(hbatch->n_tracks > 0 && hbatch->pcm && hbatch->pcm_offset && hbatch->n_samples, "batch pointers must not be NULL");
    TA_REQUIRE(hbatch->channels == 1 || hbatch->channels == 2, "channels must be 1 or 2");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    TA_CUDA(cudaSetDevice(plan->desc.device));
    DevAllocs mem(st);
    // PCM span on the host -> one device buffer with the same element offsets
    int64_t span = 0;
    for (int i = 0; i < hbatch->n_tracks; ++i) {
        TA_REQUIRE(hbatch->n_samples[i] >= 0 && hbatch->pcm_offset[i] >= 0, "n_samples / pcm_offset must be >= 0");
        span = std::max<int64_t>(span, hbatch->pcm_offset[i] + int64_t(hbatch->channels) * hbatch->n_samples[i]);
    }
    float* d_pcm = reinterpret_cast<float*>(mem.get(size_t(span) * sizeof(float)));
    if (!d_pcm) return cuda_fail(cudaGetLastError(), "cudaMallocAsync(pcm)", __FILE__, __LINE__);
    if (span) TA_CUDA(cudaMemcpyAsync(d_pcm, hbatch->pcm, size_t(span) * sizeof(float), cudaMemcpyHostToDevice, st));
    ta_batch dbatch = *hbatch;
    dbatch.pcm = d_pcm;
    HostBatch hb;
    int rc = build_host_batch(plan, &dbatch, hb);
    if (rc != TA_OK) return rc;
    const size_t P = size_t(hb.total_pitch), nt = size_t(hb.n_tracks), B = size_t(plan->n_bins), M = size_t(plan->desc.n_mels);
    const size_t W = size_t(std::max(2, plan->desc.tempogram_win));
    const bool want_cqt = hout->chroma_cqt || hout->cqt_mag || hout->cqt_tuning;
    size_t Pc = 0;
    if (want_cqt) {
        if ((rc = cqt_supported(plan)) != TA_OK) return rc;
        for (auto& t : hb.tracks) Pc += size_t(ta_frame_pitch(cqt_frame_count(plan, t.n_samples)));
    }
    // dependency closure: buffers the schedule needs on the device although the caller did not ask for them
    const bool w_tempo = hout->tempogram, w_ac = hout->autocorr, w_env = hout->onset_env || w_ac || w_tempo;
    const bool w_mel = hout->mel || w_env || hout->flux_linear || hout->mfcc || hout->self_similarity;
    const bool w_chroma = hout->chroma || hout->tuning;
    const bool w_hpss = hout->hpss_harmonic || hout->hpss_percussive;
    const bool w_mag = hout->magnitude || w_chroma || w_hpss || want_cqt || hout->rolloff_bin;
    const bool w_fmax = hout->frame_max || w_chroma || want_cqt;
    ta_frontend_out d{};
    d.kw_pitch = hout->kw_pitch;
    d.rms_pitch = hout->rms_pitch;
    d.true_peak_oversample = hout->true_peak_oversample;
    struct Copy { void* host; void* dev; size_t bytes; };
    std::vector<Copy> copies;
    bool oom = false;
    auto want = [&](bool need, void* host, size_t bytes) -> void* {
        if (!need && !host) return nullptr;
        void* p = mem.get(bytes);
        if (!p) { oom = true; return nullptr; }
        if (host) copies.push_back({host, p, bytes});
        return p;
    };
    d.magnitude = (float*)want(w_mag, hout->magnitude, B * P * 4);
    d.mel = (float*)want(w_mel, hout->mel, M * P * 4);
    d.onset_env = (float*)want(w_env, hout->onset_env, P * 4);
    d.autocorr = (double*)want(false, hout->autocorr, P * 8);
    d.flux_linear = (double*)want(false, hout->flux_linear, P * 8);
    d.ltas = (double*)want(false, hout->ltas, nt * B * 8);
    d.centroid = (double*)want(false, hout->centroid, P * 8);
    d.rolloff_bin = (int32_t*)want(false, hout->rolloff_bin, P * 4);
    d.band_energy = (double*)want(false, hout->band_energy, nt * 2 * B * 8);
    d.moments = (double*)want(false, hout->moments, nt * TA_N_MOMENTS * 8);
    d.kw_blocks = (double*)want(false, hout->kw_blocks, nt * size_t(std::max(1, hout->kw_pitch)) * 8);
    d.lufs = (double*)want(false, hout->lufs, nt * 8);
    d.rms_momentary = (double*)want(false, hout->rms_momentary, nt * size_t(std::max(1, hout->rms_pitch)) * 8);
    d.rms_short = (double*)want(false, hout->rms_short, nt * size_t(std::max(1, hout->rms_pitch)) * 8);
    d.frame_max = (float*)want(w_fmax, hout->frame_max, P * 4);
    d.chroma = (float*)want(w_chroma, hout->chroma, 12 * P * 4);
    d.tuning = (double*)want(w_chroma, hout->tuning, nt * 8);
    d.tempogram = (float*)want(false, hout->tempogram, W * P * 4);
    d.true_peak = (float*)want(false, hout->true_peak, nt * 4);
    d.hpss_harmonic = (float*)want(w_hpss, hout->hpss_harmonic, P * 4);
    d.hpss_percussive = (float*)want(w_hpss, hout->hpss_percussive, P * 4);
    d.hpss_scratch = (float*)want(w_hpss, nullptr, B * P * 4);
    d.mfcc = (double*)want(hout->self_similarity != nullptr, hout->mfcc, size_t(TA_N_MFCC) * P * 8);
    d.self_similarity = (double*)want(false, hout->self_similarity, P * 8);
    d.chroma_cqt = (float*)want(want_cqt, hout->chroma_cqt, 12 * Pc * 4);
    d.cqt_tuning = (double*)want(want_cqt, hout->cqt_tuning, nt * 8);
    d.cqt_mag = (float*)want(false, hout->cqt_mag, 252 * Pc * 4);
    if (want_cqt) {
        d.cqt_scratch_bytes = cqt_scratch_bytes(plan, hb);
        d.cqt_scratch = want(true, nullptr, size_t(d.cqt_scratch_bytes));
    }
    Workspace ws;
    const size_t ws_bytes = carve_workspace(plan, hb, nullptr, ws);
    void* d_ws = mem.get(ws_bytes);
    if (oom || !d_ws) return cuda_fail(cudaGetLastError(), "cudaMallocAsync(outputs)", __FILE__, __LINE__);
    rc = frontend_impl(plan, &dbatch, &d, d_ws, ws_bytes, stream, nullptr);
    if (rc != TA_OK) return rc;
    for (const Copy& c : copies) TA_CUDA(cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, st));
    TA_CUDA(cudaStreamSynchronize(st));
    return TA_OK;
}

int ta_frontend_run_profiled(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out, void* workspace,
                             size_t workspace_bytes, void* stream, float stage_ms[TA_N_STAGES]) {
    TA_REQUIRE(stage_ms, "stage_ms must not be NULL");
    cudaEvent_t ev[TA_N_STAGES + 1];
    for (auto& e : ev) TA_CUDA(cudaEventCreate(&e));
    int rc = frontend_impl(plan, batch, out, workspace, workspace_bytes, stream, ev);
    if (rc == TA_OK) {
        cudaError_t e = cudaEventSynchronize(ev[TA_N_STAGES]);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventSynchronize", __FILE__, __LINE__);
        for (int i = 0; i < TA_N_STAGES && rc == TA_OK; ++i) cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]);
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return rc;
}

}  // extern "C"
