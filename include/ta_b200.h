/* ta_b200.h -- C ABI of the B200-native spectral frontend for track-analyser.
 *
 * The reference (cillianjoy/track-analyser) has no FFI of its own: its hot path is
 * reached through module-level Python functions that delegate to librosa /
 * pyloudnorm.  This header is the boundary a maintainer
 * binds instead (ctypes stub: INTEGRATION.md).  Each entry point names the
 * reference call sites (paths under /root/reference/src/track_analyser) whose
 * arithmetic it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative TA_ERR_* code; the text of
 *     the last error on the calling thread is ta_last_error().  No C++ exception
 *     and no abort() crosses this boundary.
 *   - all `float*`/`double*`/`int32_t*` data pointers are DEVICE pointers owned by
 *     the caller (torch tensors in the shipped Python shim).  The library never
 *     allocates or frees caller-visible memory; scratch is the caller-provided
 *     workspace (ta_workspace_bytes).  Plan tables are owned by the plan.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work
 *     on it and returns.  The one exception is ta_frontend_run_host, which takes
 *     HOST buffers, does the copies itself and synchronises the stream before
 *     returning (for callers that own no device memory, e.g. a numpy-only binding).
 *   - ragged batches: track i has n_samples[i] samples per channel and
 *     T_i = 1 + n_samples[i] / hop frames.  All (rows, T_i) matrices of track i
 *     are row-major with row pitch ld_i = ta_frame_pitch(T_i) (T_i rounded up to
 *     a multiple of 32 floats so that every row starts on a 128-byte line) and
 *     are packed track after track: matrix of track i starts at element
 *     rows * frame_pitch_prefix[i].  Per-frame series use the same pitch.
 *   - there is NO CPU fallback: with no CUDA device every entry point that needs
 *     one fails with TA_ERR_CUDA.
 */
#ifndef TA_B200_H
#define TA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TA_ABI_VERSION 4

#if defined(__GNUC__)
#define TA_API __attribute__((visibility("default")))
#else
#define TA_API
#endif

#define TA_OK 0
#define TA_ERR_INVALID (-1)   /* bad argument */
#define TA_ERR_CUDA (-2)      /* CUDA runtime/driver error (text in ta_last_error) */
#define TA_ERR_UNSUPPORTED (-3)
#define TA_ERR_WORKSPACE (-4) /* workspace too small */

#define TA_N_MOMENTS 10      /* doubles per track in ta_frontend_out.moments */
#define TA_N_MFCC 13         /* cepstral coefficients per frame in ta_frontend_out.mfcc (structure.py:199) */

typedef struct ta_plan ta_plan;

TA_API int ta_abi_version(void);
TA_API const char* ta_last_error(void);

/* Immutable per-configuration state: FFT twiddles, periodic Hann window (computed
 * in double), sparse Slaney mel filterbank, bin frequencies, K-weighting biquads.
 * Replaces the table construction inside librosa.stft / librosa.filters.mel /
 * pyloudnorm.Meter.__init__ reached from features.py:79, tempo.py:19,
 * analysis/structure.py:48-59, analysis/loudness.py:60. */
typedef struct ta_plan_desc {
    int32_t device;        /* CUDA device ordinal */
    int32_t sample_rate;   /* Hz */
    int32_t n_fft;         /* 256, 512, 1024, 2048 or 4096 (256 / 512 run zero-padded on the 1024-point transform) */
    int32_t hop;           /* hop length in samples, any positive value (hop = n_fft/4 takes the shared-sample fast path) */
    int32_t n_mels;        /* mel bands (0: no mel tables) */
    int32_t n_chroma;      /* chroma bins (12) */
    int32_t tempogram_win; /* tempogram window in frames (384) */
    int32_t reserved;      /* must be 0 */
    double fmin;           /* mel lower edge, Hz */
    double fmax;           /* mel upper edge, Hz; <= 0 means sample_rate / 2 */
    double roll_percent;   /* spectral roll-off fraction (0.85) */
    double meter_block;    /* loudness gating block in seconds (0.4); a double because pyloudnorm's
                              block bounds int(T_g*(j*step)*rate) are evaluated in Python floats */
} ta_plan_desc;

TA_API int ta_plan_create(const ta_plan_desc* desc, ta_plan** out);
/* Same with an analysis window other than Hann: `window` = n_fft float64 values on the HOST, what
 * scipy.signal.get_window(name, n_fft, fftbins=True) returns for librosa.stft(window=name) (features.py:66-79 passes
 * its `window` argument through); NULL = periodic Hann. */
TA_API int ta_plan_create_window(const ta_plan_desc* desc, const double* window, ta_plan** out);
TA_API void ta_plan_destroy(ta_plan* plan);
TA_API int ta_plan_n_bins(const ta_plan* plan);

/* Copies plan tables to host for inspection/tests: which = 0 window (n_fft f32),
 * 1 dense mel basis (n_mels * n_bins f32), 2 bin frequencies (n_bins f64),
 * 3 biquad coefficients (12 f64: b0 b1 b2 a0 a1 a2 shelf, then high-pass). */
TA_API int ta_plan_table(const ta_plan* plan, int which, void* host_out, size_t bytes);

static inline int64_t ta_frame_count(int64_t n_samples, int32_t hop) { return 1 + n_samples / hop; }
static inline int64_t ta_frame_pitch(int64_t n_frames) { return (n_frames + 31) & ~(int64_t)31; }

/* Track batch (device PCM, host metadata). */
typedef struct ta_batch {
    int32_t n_tracks;
    int32_t channels;          /* 1: mono rows; 2: planar L then R per track */
    const float* pcm;          /* device base; channel c of track i at pcm + pcm_offset[i] + c*n_samples[i] */
    const int64_t* pcm_offset; /* host [n_tracks], element offsets, multiples of 4 */
    const int64_t* n_samples;  /* host [n_tracks] */
} ta_batch;

/* Outputs of the fused frontend.  Any pointer may be NULL to skip that output
 * (the dependent stages still run if a later output needs them).  Sizes use
 * P = sum_i ta_frame_pitch(T_i), B = n_fft/2+1, M = n_mels. */
typedef struct ta_frontend_out {
    float* magnitude;      /* [B * P]  |STFT| of the mono (mid) signal: structure.py:48-51, features.py:79-80 */
    float* mel;            /* [M * P]  mel power spectrogram: structure.py:53-59, inside tempo.py:19 */
    float* onset_env;      /* [P]      onset-strength envelope: tempo.py:16-24 */
    double* autocorr;      /* [P]      librosa.autocorrelate(onset_env): tempo.py:38 (float64 like numpy 1.26) */
    double* flux_linear;   /* [P]      onset_strength(S=mel) on linear power: structure.py:194-196 */
    double* ltas;          /* [n_tracks * B] time-mean of magnitude: features.py:80 */
    double* centroid;      /* [P]      spectral centroid, Hz: features.py:97-100 */
    int32_t* rolloff_bin;  /* [P]      roll-off bin index k (frequency = k*sr/n_fft): features.py:116-123; numpy's sequential
                              float32 cumsum is reproduced on the magnitude matrix, so this output needs `magnitude` */
    double* band_energy;   /* [n_tracks * 2 * B] per-bin time sums of |mid|^2 then |side|^2: stereo.py:95-122 */
    double* moments;       /* [n_tracks * TA_N_MOMENTS] sum L, R, L^2, R^2, LR, mid^2, side^2, n, sum |L|, sum |R|:
                              stereo.py:62-83, loudness.py:118, harmony.py:270-282 (mono batches: L = the signal, R = 0) */
    double* kw_blocks;     /* [n_tracks * kw_pitch] K-weighted gating-block mean squares z_j: loudness.py:60-61 */
    double* lufs;          /* [n_tracks] gated integrated loudness: loudness.py:61 */
    double* rms_momentary; /* [n_tracks * rms_pitch] mean-square of centred frames, 0.4 s window: loudness.py:57 */
    double* rms_short;     /* [n_tracks * rms_pitch] same, 3.0 s window: loudness.py:56 */
    float* frame_max;      /* [P]      max_f |X| per frame (input of the tuning estimate) */
    float* chroma;         /* [12 * P] chroma_stft, inf-normalised per frame: harmony.py:108,149 (needs magnitude) */
    double* tuning;        /* [n_tracks] estimated tuning in fractions of a bin (librosa.estimate_tuning) */
    float* tempogram;      /* [win * P] autocorrelation tempogram of onset_env: report.py:260 */
    float* true_peak;      /* [n_tracks] max |y| of the 8x polyphase-oversampled mono signal (linear; dBTP =
                              20 log10(. + 1e-12)): true_peak_dbtp, analysis/loudness.py:81-97 */
    float* hpss_harmonic;  /* [P]      sum over bins of librosa.decompose.hpss(magnitude)[0]: structure.py:52,143,213 */
    float* hpss_percussive;/* [P]      same for the percussive component (structure.py:144,212) */
    float* hpss_scratch;   /* [B * P]  caller-provided scratch (time-direction medians); required with the two above */
    double* mfcc;          /* [TA_N_MFCC * P] librosa.feature.mfcc(S=power_to_db(mel + 1e-9), n_mfcc=13), float64:
                              analysis/structure.py:192,199 (needs mel) */
    double* self_similarity; /* [P]    MFCC self-similarity novelty (Gaussian-smoothed cepstrum, 2 s context windows, 1 - cosine):
                              analysis/structure.py:199-210 (needs mfcc) */
    float* chroma_cqt;     /* [12 * Pc] librosa.feature.chroma_cqt(y, sr), inf-normalised per frame: harmony.py:107,148.
                              Pc = sum_i ta_frame_pitch(ta_cqt_frame_count(plan, n_samples[i])); needs magnitude, frame_max,
                              cqt_tuning and cqt_scratch; plan must be n_fft 2048 / hop 512 (librosa's defaults there) */
    double* cqt_tuning;    /* [n_tracks] librosa.estimate_tuning(y=y, bins_per_octave=36), the tuning chroma_cqt applies */
    float* cqt_mag;        /* [252 * Pc] optional: |librosa.cqt(y, sr, n_bins=252, bins_per_octave=36, tuning=cqt_tuning)| */
    void* cqt_scratch;     /* caller-provided scratch of ta_cqt_scratch_bytes() bytes, 256-byte aligned (decimated signals) */
    int32_t kw_pitch;      /* capacity per track of kw_blocks */
    int32_t rms_pitch;     /* capacity per track of rms_momentary / rms_short */
    uint64_t cqt_scratch_bytes; /* size of cqt_scratch */
    int32_t true_peak_oversample; /* oversampling factor of true_peak, 1 .. 32; 0 = the reference's default 8 (loudness.py:81) */
    int32_t reserved;      /* must be 0 */
} ta_frontend_out;

TA_API size_t ta_workspace_bytes(const ta_plan* plan, const ta_batch* batch);

/* Fused schedule K1..K11 for a batch of tracks; see DESIGN.md for the kernels.  Work is enqueued on `stream`; the
 * chains that do not depend on the main one (time-domain pass; chroma / HPSS behind K1) run on a second stream owned by
 * the plan and are joined back into `stream` before the call returns, so the caller sees plain stream ordering
 * (environment TA_OVERLAP=0 keeps everything on `stream`). */
TA_API int ta_frontend_run(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ta_frontend_run for callers without device memory: every data pointer of `batch` (pcm) and `out` is a HOST pointer
 * (pageable or pinned) with the element counts documented on ta_frontend_out; hpss_scratch / cqt_scratch are ignored.
 * The library takes device memory for the PCM, the requested outputs (and the intermediates they depend on), scratch and
 * workspace from the stream-ordered pool (cudaMallocAsync), copies in, runs the fused schedule, copies the outputs back,
 * frees, and synchronises `stream` before returning.  kw_pitch / rms_pitch as for ta_frontend_run. */
TA_API int ta_frontend_run_host(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out, void* stream);

/* Same schedule as ta_frontend_run, but brackets each stage with CUDA events on
 * `stream`, synchronises, and returns the device time of each stage in
 * milliseconds: stage_ms[0] STFT+mel+features (K1/K2/K7), [1] onset flux (K3),
 * [2] autocorrelation (K4), [3] tempogram (K4b), [4] chroma_stft (K2b),
 * [5] time-domain pass + gating (K5/K6).  For bench.py's roofline figure; not for
 * production use (it blocks the host). */
#define TA_N_STAGES 6
TA_API int ta_frontend_run_profiled(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out,
                                    void* workspace, size_t workspace_bytes, void* stream, float stage_ms[TA_N_STAGES]);

/* Number of kernels this library has launched in the calling process so far. */
TA_API uint64_t ta_launch_count(void);

/* --- single-stage entry points (each is also a step of ta_frontend_run) ------ */

/* K1+K2+K7: fused frame + Hann + FFT + |X| (+ mel, LTAS, centroid, roll-off, band sums).
 * Replaces librosa.stft / melspectrogram / spectral_centroid / spectral_rolloff at
 * features.py:79,97,116; stereo.py:95-96; structure.py:48,53; harmony.py:254. */
TA_API int ta_stft_features(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* K3: power_to_db (top_db 80 against the per-track max) + lag-1 diff + half-wave
 * rectification + mean over mel + 3-frame left pad: librosa.onset.onset_strength
 * at tempo.py:19.  mel_max_bits: per-track float bit pattern of max(mel). */
TA_API int ta_onset_flux(const ta_plan* plan, const ta_batch* batch, const float* mel,
                  const uint32_t* mel_max_bits, float* onset_env, double* flux_linear, void* stream);

/* K4: full-length autocorrelation irfft(|rfft(x, n_pad)|^2)[:T] in float64:
 * librosa.autocorrelate at tempo.py:38. */
TA_API int ta_autocorrelate(const ta_plan* plan, const ta_batch* batch, const float* onset_env,
                     double* autocorr, void* workspace, size_t workspace_bytes, void* stream);

/* K5+K6: one pass over the PCM: K-weighting biquad cascade as a chunked linear
 * recurrence scan (float64 state, float32 round trip between stages like
 * scipy.signal.lfilter inside pyloudnorm), gating-block energies, BS.1770 gating,
 * mid/side/LR moments and the centred RMS frames: loudness.py:30-61,118; stereo.py:62-83; and
 * (K8) the true peak: scipy.signal.resample_poly(x, 8, 1) + max |.| of loudness.py:81-97, evaluated
 * exactly but only where the 161-tap interpolator can reach the maximum. */
TA_API int ta_time_domain(const ta_plan* plan, const ta_batch* batch, const ta_frontend_out* out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* K2b: chroma_stft on an existing magnitude spectrogram: tuning estimate (piptrack on the power
 * spectrogram, median gate, 0.01-bin histogram), chroma filterbank for that tuning, projection and
 * per-frame inf-norm: librosa.feature.chroma_stft at harmony.py:108,149. */
TA_API int ta_chroma_stft(const ta_plan* plan, const ta_batch* batch, const float* magnitude, const float* frame_max,
                          float* chroma, double* tuning, void* workspace, size_t workspace_bytes, void* stream);

/* K0: PCM decode on the device.  `interleaved` holds n_frames * channels samples as a WAV data chunk stores them
 * (device pointer); planar_out receives (channels, n_frames) float32, the layout of io.load_audio (io.py:72-79) and of
 * ta_batch.pcm.  Conversions are libsndfile's: int16 / 2^15, packed 24-bit / 2^23, int32 / 2^31, float32 copied. */
#define TA_PCM_S16 1
#define TA_PCM_S24 2
#define TA_PCM_S32 3
#define TA_PCM_F32 4
TA_API int ta_decode_pcm(const void* interleaved, int format, int channels, int64_t n_frames, float* planar_out,
                         void* stream);

/* Fingerprint of the float32 mono mix (L[i] + R[i]) * 0.5f of the planar pair `planar_stereo` (L then R, n_samples each):
 * fingerprint[0] = sum of the IEEE bit patterns, fingerprint[1] = sum of bit pattern * ((i & 0xffff) + 1), modulo 2^64
 * (device uint64[2], written by the call).  The batch driver forms the same sums over AudioInput.samples on the host: equal
 * fingerprints mean the mono samples are the mean of the stereo pair, the way utils.coerce_audio builds them (utils.py:116),
 * and one stereo run may then serve the mono stages of the track too (mono == mid exactly). */
TA_API int ta_mono_mix_fingerprint(const float* planar_stereo, int64_t n_samples, uint64_t* fingerprint, void* stream);

/* K9: per-frame sums of the harmonic and percussive components of librosa.decompose.hpss (31-wide median
 * filters along time and frequency, soft masks with power 2) on an existing magnitude spectrogram:
 * analysis/structure.py:52 as consumed at :143-144 and :212-213.  scratch: B * P floats. */
TA_API int ta_hpss_curves(const ta_plan* plan, const ta_batch* batch, const float* magnitude, float* scratch,
                          float* harmonic_sum, float* percussive_sum, void* workspace, size_t workspace_bytes, void* stream);

/* K12: librosa.feature.chroma_cqt(y=y, sr=sr) with every default (hop 512, 7 octaves x 36 bins from C1, tuning estimated
 * from y, norm=inf) on an existing magnitude spectrogram (for the tuning estimate) and the batch's PCM: harmony.py:107,148.
 * The octave decimator is a stated stage (Kaiser-windowed sinc on soxr HQ's published band edges; librosa calls libsoxr):
 * DESIGN.md.  Frame count per track: ta_cqt_frame_count (librosa's per-octave STFTs can hold one frame more or less than
 * 1 + n/hop; the stack is trimmed to the shortest).  ta_cqt_frame_count < 0 / ta_cqt_scratch_bytes == 0: unsupported plan
 * (text in ta_last_error). */
TA_API int64_t ta_cqt_frame_count(const ta_plan* plan, int64_t n_samples);
TA_API size_t ta_cqt_scratch_bytes(const ta_plan* plan, const ta_batch* batch);
TA_API int ta_chroma_cqt(const ta_plan* plan, const ta_batch* batch, const float* magnitude, const float* frame_max,
                         float* chroma_cqt, float* cqt_mag, double* cqt_tuning, void* cqt_scratch, size_t cqt_scratch_bytes,
                         void* workspace, size_t workspace_bytes, void* stream);

/* K9 with the component matrices themselves: harmonic, percussive = librosa.decompose.hpss(magnitude) ([B * P] floats each,
 * laid out like the magnitude) next to their per-frame sums.  For callers that run the reference's own analyse_structure
 * (analysis/structure.py:52, 135-144, 212-213) on the result; the fused schedule only ever needs the sums. */
TA_API int ta_hpss_components(const ta_plan* plan, const ta_batch* batch, const float* magnitude, float* scratch, float* harmonic,
                              float* percussive, float* harmonic_sum, float* percussive_sum, void* workspace,
                              size_t workspace_bytes, void* stream);

/* K4b: windowed (tempogram_win frames, Hann, centred, inf-normalised) autocorrelation of the onset
 * envelope: librosa.feature.tempogram at report.py:260.  Output rows = lags, (win, T_i) per track. */
TA_API int ta_tempogram(const ta_plan* plan, const ta_batch* batch, const float* onset_env, float* tempogram,
                        void* workspace, size_t workspace_bytes, void* stream);

/* K11: resampy.resample(x, sr_orig, sr_new) (band-limited sinc interpolation, "kaiser_best") as the reference calls
 * it per channel from utils._resample (utils.py:55-70) and load_audio (io.py:126-128).  The caller supplies the
 * right half of the interpolation window (float64, n_window = num_zeros * num_table + 1 samples, num_table per
 * zero crossing; resampy.filters.sinc_window) as a HOST array; the resampler keeps its scaled copy on the device.
 * ta_resample converts n_rows planar rows of n_in samples (device, row pitch src_pitch) into rows of
 * ta_resampler_out_len(n_in) = int(n_in * sr_new / sr_orig) samples. */
typedef struct ta_resampler ta_resampler;
TA_API int ta_resampler_create(int device, int sr_orig, int sr_new, const double* half_window, int n_window, int num_table,
                               ta_resampler** out);
TA_API void ta_resampler_destroy(ta_resampler* resampler);
TA_API int64_t ta_resampler_out_len(const ta_resampler* resampler, int64_t n_in);
TA_API int ta_resample(const ta_resampler* resampler, const float* src, int64_t n_in, int64_t src_pitch, int n_rows, float* dst,
                       int64_t dst_pitch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TA_B200_H */
