/* sfx.h -- C ABI of the sm_100a batched speech feature extractor (libsfx_b200.so).
 *
 * The reference has no FFI/plugin layer for this path: the boundary is the Python module
 * preprocessing/audio_preprocessing.py (reference :12-46), whose per-clip functions call librosa.
 * These entry points are what a maintainer binds (ctypes) to replace, in one
 * batched device pass, the four calls the reference makes per clip:
 *
 *   sfx_extract / sfx_extract_host  <->  extract_mfcc             (audio_preprocessing.py:22-24)
 *   / sfx_extract_host_pcm16
 *                                        extract_chroma           (audio_preprocessing.py:27-29)
 *                                        extract_spectral_features(audio_preprocessing.py:32-37)
 *                                        np.concatenate -> f32[56] (audio_preprocessing.py:45-46)
 *   lengths[] / n_default            <->  the pad/trim of load_audio (audio_preprocessing.py:14-18):
 *                                        a clip shorter than the row is read up to lengths[i] only.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns 0 on success
 * or a negative sfx_status; sfx_last_error() gives the text.  Nothing here falls back to the CPU:
 * without a CUDA device every compute entry point fails with SFX_ERR_CUDA.
 * Device entry points are stream-ordered (no host sync inside) and re-entrant across streams and host
 * threads given distinct workspaces (table sets and the pipeline mode are read under a lock / atomically
 * once per call).  The caller owns every buffer.
 */
#ifndef SFX_B200_H
#define SFX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFX_ABI_VERSION 3
#define SFX_N_CHROMA    12
#define SFX_N_SPECTRAL  4      /* [zcr, spectral_centroid, spectral_rolloff, rms] (reference :33-37) */
#define SFX_N_FFT       2048
#define SFX_HOP         512
#define SFX_N_BINS      1025
#define SFX_N_MELS      128
#define SFX_P_STRIDE    1056   /* padded bin row */
#define SFX_N_TUNINGS   100

typedef enum {
    SFX_OK              =  0,
    SFX_ERR_ARG         = -1,  /* bad argument (null pointer, negative size, n_mfcc out of range ...) */
    SFX_ERR_CUDA        = -2,  /* CUDA runtime error (including "no device") */
    SFX_ERR_NOT_INIT    = -3,  /* sfx_init_tables has not been called for this device */
    SFX_ERR_WORKSPACE   = -4,  /* workspace too small */
    SFX_ERR_BAD_CLIP    = -5   /* host entry points only: a clip has length <= 0 (rejected before any work) or a
                                * non-finite sample (found on the device: the clip's feature row is NaN; every row of
                                * the batch has been delivered when the call returns this code) */
} sfx_status;

/* Host-side constant tables (generated in float64 by sfx_b200/tables.py; see that file for layouts). */
typedef struct {
    int32_t       sr;            /* sample rate the banks were built for (reference config.py:57) */
    int32_t       pip_kmin;      /* first/last rFFT bin with 150 Hz <= f < 4000 Hz (librosa.piptrack) */
    int32_t       pip_kmax;
    int32_t       mel_ps;        /* row stride of the per-lane mel partial-sum slots (odd, <= 31) */
    int32_t       mel_flush32;   /* 1 if the Slaney interval index advances at the Nyquist bin */
    const float  *hann;          /* [2048] periodic Hann window; uploaded and validated, but since round 2 the kernels build the
                                  * window from cos/sin of the sample phase (within 1 ulp(0.5) of these values) instead of reading it */
    const float  *tw1;           /* [32][32][2] */
    const float  *tw2;           /* [32][32][2] */
    const float  *mel_ab;        /* [33][32][2] (falling, rising) weights of bin 32*lane + j at [j][lane] */
    const uint32_t*mel_mask;     /* [32] bit j: interval index advances at bin 32*lane + j */
    const int32_t*mel_src;       /* [128][3] partial-sum slots of each filter (32*mel_ps = zero slot) */
    const uint16_t*chroma16;     /* [100][2][12][1056] IEEE half: bank = hi + 2^-11 * lo (tensor-core operands) */
    const float  *chroma_ny;     /* [100][12] float32 weights of the Nyquist bin */
    const double *dct;           /* [128][128] */
    const double *edges;         /* [101] */
    const uint32_t*chroma_frag;  /* [100][32][2][2][32][4] the bank as mma.m16n8k16 A fragments: (tuning, 32-bin step,
                                  * half step, hi/lo, lane) -> one 16-byte quad (stream pipeline; see tables.py) */
    const uint8_t *chroma_umma;  /* [100][16][4096] the bank as tcgen05 B-operand shared-memory images (K-major, 128-byte
                                  * swizzle), one per (tuning, 64-bin block); used by builds with SFX_CHROMA_UMMA */
} sfx_tables_host;

/* Optional per-clip / per-frame intermediates for parity triage (device pointers, any may be NULL). */
typedef struct {
    float   *P;          /* [B][T_dbg][1056] power spectrum |X|^2 */
    float   *logmel;     /* [B][T_dbg][128]  10*log10(max(1e-10, mel)) before the top_db clamp */
    float   *frame_feat; /* [B][T_dbg][4]    per-frame centroid(Hz), rolloff(Hz), rms, zc-count of the frame's hop */
    float   *clip_info;  /* [B][8]           tuning, gmax, n_peaks, median threshold, n_selected, T,
                          *                   peaks whose fast-path result differs from the reference form (must be 0), 0 */
    int32_t  T_dbg;      /* frames allocated per clip in the arrays above */
} sfx_debug_out;

int         sfx_abi_version(void);
const char *sfx_last_error(void);

/* Number of CUDA devices visible (<0 on error; 0 = none: compute calls will fail). */
int         sfx_device_count(void);

/* Upload the constant tables of sample rate tables->sr to `device`.  Idempotent per (device, sr); table sets
 * of several sample rates coexist and are selected by the `sr` argument of the extract calls. */
int         sfx_init_tables(int device, const sfx_tables_host *tables);

/* Bytes of device workspace sfx_extract needs for clips of at most max_samples samples, whatever the batch size
 * (persistent CTAs each own a fixed number of scratch slices).  0 on error. */
size_t      sfx_workspace_bytes(int device, int64_t max_samples);

/* The same for batches of at most B clips: a single-clip request needs one slice, not one per resident CTA
 * (B <= 0: any batch size, i.e. sfx_workspace_bytes).  The host entry points size their cached workspaces with it. */
size_t      sfx_workspace_bytes_batch(int device, int64_t max_samples, int64_t B);

/* Kernels launched by the most recent sfx_extract / sfx_extract_debug / sfx_extract_host call of this thread (1 before
 * any call; stream and fused pipelines: 1 per call or host chunk, +1 for a ragged batch (lengths != NULL, more clips than
 * CTAs), whose clips are processed longest first; split pipeline: 3 per chunk of <= 1024 clips). */
int         sfx_launches_per_extract(void);

/* Pipeline selection: 0 = auto (default), 1 = fused, 2 = split, 3 = stream, 4 = fused_umma.  One arithmetic:
 *   stream  one persistent 16-warp CTA per SM; every warp pulls STFT frames or whole per-clip tails (tuning estimate,
 *           MFCC, chroma, pooled row) from a CTA-local scheduler, so a tail occupies one warp while 15 keep transforming
 *           frames.  Built and measured (1.42 M clips/s against the fused kernel's 2.03 M on the bench mix); kept as an A/B.
 *   fused   persistent 8-warp CTAs, two per SM, one clip per CTA at a time (frames, barrier, tail by all 8 warps).
 *           Highest throughput on large batches: the kernel bench.py times.
 *   split   frame-parallel two-kernel pipeline per chunk of <= 1024 clips: lowest latency for small batches.
 *   fused_umma  the fused kernel with its chroma projection on tcgen05 (UMMA, accumulator in tensor memory) instead of
 *           mma.sync: same throughput (measured), kept as the A/B of the two tensor paths.
 * auto = split for batches that fit one chunk of at most 1024 clips (256 when ragged), fused above that.  Also settable
 * through the environment variable SFX_PIPELINE=auto|fused|split|stream|fused_umma before the first call.  The mode is read once per
 * call (atomically); call sfx_workspace_bytes again after changing it. */
int         sfx_set_pipeline(int mode);

/* Batched extraction, device buffers.
 *   sr          sample rate of a table set uploaded with sfx_init_tables
 *   wave        [B] rows of float32 samples, row i at wave + i*row_stride (device)
 *   lengths     [B] int32 sample counts (device) or NULL = every clip has n_default samples
 *   out         [B] rows of (n_mfcc + 12 + 4) float32 at out + i*out_stride (device):
 *               [mfcc_0..n_mfcc-1 | chroma C..B | zcr, centroid_Hz, rolloff_Hz, rms]
 *   workspace   >= sfx_workspace_bytes(device, max length in the batch)
 *   stream      cudaStream_t (NULL = legacy default stream)
 * A clip with length <= 0 or > max_samples yields a row of NaN (the host wrapper raises, as librosa would); lengths are
 * read on the device, so they are not validated by the call itself.
 */
int         sfx_extract(int device, int32_t sr, const float *wave, int64_t row_stride, const int32_t *lengths,
                        int64_t n_default, int64_t max_samples, int32_t B, int32_t n_mfcc,
                        float *out, int64_t out_stride, void *workspace, size_t workspace_bytes,
                        void *stream);

/* Same, additionally filling `dbg`. */
int         sfx_extract_debug(int device, int32_t sr, const float *wave, int64_t row_stride, const int32_t *lengths,
                              int64_t n_default, int64_t max_samples, int32_t B, int32_t n_mfcc,
                              float *out, int64_t out_stride, void *workspace, size_t workspace_bytes,
                              void *stream, const sfx_debug_out *dbg);

/* Batched extraction, HOST buffers (the reference-facing plugin path): pinned staging, chunked
 * H2D copy overlapped with the kernel on three streams, D2H of the feature rows, then a stream sync.
 * host_lengths may be NULL; a length outside [1, row_stride] is rejected with SFX_ERR_BAD_CLIP before any work.
 * Samples are not scanned on the host (that would cost more than the PCIe copy): a clip with a NaN / Inf sample gets a
 * NaN feature row, and the call returns SFX_ERR_BAD_CLIP naming the first such clip after delivering every row (the
 * reference raises librosa.ParameterError per clip; the Python wrapper raises for the NaN rows).  Allocates and caches its
 * own device buffers per device; calls on one device are serialised.  chunk_clips <= 0 selects the default chunk. */
int         sfx_extract_host(int device, int32_t sr, const float *host_wave, int64_t row_stride,
                             const int32_t *host_lengths, int64_t n_default, int32_t B, int32_t n_mfcc,
                             float *host_out, int64_t out_stride, int32_t chunk_clips);

/* Same for 16-bit PCM rows as they sit in a WAV file: converted on the device exactly as libsndfile does for librosa.load
 * (x / 32768, reference :13), then extracted.  Half the host-to-device bytes of sfx_extract_host; valid for audio that is
 * already at the sample rate `sr` (librosa.load then does no resampling) and mono. */
int         sfx_extract_host_pcm16(int device, int32_t sr, const int16_t *host_pcm, int64_t row_stride,
                                   const int32_t *host_lengths, int64_t n_default, int32_t B, int32_t n_mfcc,
                                   float *host_out, int64_t out_stride, int32_t chunk_clips);

/* ---- scope row f3: load_audio (reference :12-19) on the device for 16-bit PCM ------------------------------------
 * Polyphase resampler description = scipy.signal.resample_poly(x, up, down) as this package's load_audio uses it:
 * taps = the zero-padded filter (n_pre_pad zeros, then firwin(2*half_len+1, 1/max(up,down), ('kaiser', 5.0)) * up with
 * half_len = 10*max(up,down)), n_pre_remove = (half_len + n_pre_pad) / down.  sfx_b200/resample.py builds it. */
typedef struct {
    int32_t       up, down;       /* target_sr / g, native_sr / g, g = gcd */
    int32_t       n_taps;
    int32_t       n_pre_remove;
    const double *taps;           /* [n_taps] host */
} sfx_resampler_host;

/* file -> features for B clips of raw 16-bit PCM frames (mono or interleaved stereo) at their native rate:
 * x/32768 (soundfile), channel mean (float32), resample_poly in float64 (bit-identical to scipy's result; rs == NULL or
 * up == down: native rate, no filter), float32, zero pad / trim to n_target = sr * duration samples, then the extractor.
 *   host_frames  [B] frames (samples per channel) to take from each row, already limited to round(native_sr * duration)
 *                as load_audio does (:13), or NULL = frames_default
 *   row_stride   int16 elements between rows of host_pcm
 * Same chunk-pipelined H2D || kernels || D2H structure as sfx_extract_host.  librosa's own resampler (soxr_hq) is not
 * restated: see DESIGN.md, row f3. */
int         sfx_preprocess_host_pcm16(int device, int32_t sr, const sfx_resampler_host *rs, const int16_t *host_pcm,
                                      int64_t row_stride, int32_t channels, const int32_t *host_frames,
                                      int64_t frames_default, int64_t n_target, int32_t B, int32_t n_mfcc,
                                      float *host_out, int64_t out_stride, int32_t chunk_clips);
const char *sfx_frontend_last_error(void);
int         sfx_frontend_release(int device);

/* Release cached device buffers/tables of `device` (tests; process exit does it implicitly). */
int         sfx_release(int device);

/* ---- scope row f1: device-resident StandardScaler + speech DNN forward ------------------------------------------
 * Consumer of the feature rows in the reference's inference/speech_inference.py:66-76 (scaler.transform + model.predict)
 * and :85-105 (layers[-3] tap); architecture of model_training/train_speech_model.py:55-90.  FP32 kernels. */
typedef struct {
    int32_t             n_layers;      /* dense layers (6 for the reference model) */
    const int32_t      *dims;          /* [n_layers + 1] widths, e.g. 56,512,512,256,128,64,7 */
    const float *const *kernel;        /* [n_layers] Keras layout [in][out] */
    const float *const *bias;          /* [n_layers] */
    const float *const *bn_gamma;      /* [n_layers - 1] BatchNormalization after every hidden Dense (may be NULL) */
    const float *const *bn_beta;
    const float *const *bn_mean;
    const float *const *bn_var;
    float               bn_eps;        /* Keras default 1e-3 */
    const double       *scaler_mean;   /* [dims[0]] sklearn StandardScaler.mean_  (NULL = no scaler) */
    const double       *scaler_scale;  /* [dims[0]] sklearn StandardScaler.scale_ */
} sfx_dnn_host;

int         sfx_dnn_create(int device, const sfx_dnn_host *model, void **handle);
int         sfx_dnn_destroy(void *handle);
size_t      sfx_dnn_workspace_bytes(void *handle, int32_t B);
int         sfx_dnn_launches_per_forward(void *handle);
const char *sfx_dnn_last_error(void);
/* feats [B] rows of dims[0] float32 (device, e.g. the output of sfx_extract); probs [B][dims[n]] softmax outputs;
 * tap [B][dims[n-1]] last hidden activation or NULL.  Stream-ordered. */
int         sfx_dnn_forward(void *handle, const float *feats, int64_t feat_stride, int32_t B, float *probs,
                            int64_t probs_stride, float *tap, int64_t tap_stride, void *workspace,
                            size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SFX_B200_H */
