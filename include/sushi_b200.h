/*
 * sushi_b200.h -- C ABI of the B200-native audio template-matching library.
 *
 * This is the drop-in boundary for ONE path of tp7/Sushi: the per-event audio
 * template match that the reference performs in Python through OpenCV,
 *
 *     WavStream.get_substream / find_substream      reference wav.py:168-188
 *     cv2.matchTemplate(..., TM_SQDIFF_NORMED)      reference wav.py:185
 *     result.argmin(axis=1)[0]                      reference wav.py:186
 *     load-time downsample / pad / normalise        reference wav.py:64-91,108-156
 *
 * The reference has no FFI of its own (it is pure Python calling cv2), so these
 * entry points are what a ctypes binding inside the reference's wav.py would
 * call (see INTEGRATION.md for that stub).  Conventions:
 *
 *   - plain C: pointers, sizes, integer sample offsets.  All time -> sample
 *     conversion (wav.py:173-175) stays on the Python side so that it is
 *     reproduced bit-for-bit; the C side never sees seconds.
 *   - every call returns SB_OK (0) or a negative SB_E* code; the text of the
 *     last failure on the calling thread is available from sb_last_error().
 *   - the caller owns every host buffer; the library owns device memory behind
 *     opaque handles.  One host thread drives one GPU (one process per GPU).
 *   - blocking calls return after their results are in the caller's host
 *     buffers.  *_device variants take/return device pointers and only enqueue
 *     work on the library stream (sb_sync() waits for it).
 *
 * Result convention (mirrors wav.py:185-188): for a query with template length
 * n and nlags candidate positions, curve[j] is OpenCV's TM_SQDIFF_NORMED value
 * of the template against image[lag0+j .. lag0+j+n), j in [0, nlags), rounded
 * to float32; the call returns diff = min_j curve[j] and idx = the FIRST j that
 * attains it (numpy argmin rule).
 */
#ifndef SUSHI_B200_H
#define SUSHI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_ABI_VERSION 1

/* status codes */
#define SB_OK            0
#define SB_EINVAL       -1   /* bad argument (range, NULL, dtype)            */
#define SB_ECUDA        -2   /* CUDA runtime / cuFFT failure                 */
#define SB_ENOMEM       -3   /* host or device allocation failed             */
#define SB_ESTATE       -4   /* library not initialised / already shut down  */

/* sample types of a resident stream (reference wav.py:108-110,153-156) */
#define SB_U8  0             /* 'uint8'   : reference default (sushi.py:769) */
#define SB_F32 1             /* 'float32'                                    */

typedef struct sb_stream sb_stream;      /* opaque: one normalised stream in HBM */

/* ---- life cycle ------------------------------------------------------- */

/* Bind the calling process to CUDA device `device` and create the library
 * stream, plans and scratch.  Idempotent for the same device. */
int sb_init(int device);
int sb_shutdown(void);
/* ABI version of the loaded library (compare with SB_ABI_VERSION). */
int sb_abi_version(void);
/* Text of the last error on this thread ("" if none). Never NULL. */
const char* sb_last_error(void);
/* Wait until everything enqueued on the library stream has finished. */
int sb_sync(void);
/* Tuning knobs (power-of-two lag-block size in samples, 1024..65536; items per
 * chunk).  Changing the block size drops cached block spectra. */
int sb_set_block_size(int block);
int sb_get_block_size(void);
/* Engine behind sb_find*:
 *   2 (default) = the packed fused lag-block kernels (sb_fused2.cu): spectral multiply, inverse FFT in
 *       shared memory, normalisation and argmin in one launch, written around the two-wide fp32
 *       instructions of sm_100 (FFMA2/FADD2) on spectra stored in a paired layout; lag blocks of 16384
 *       at hop B -- any other geometry silently runs engine 1, and templates of 12+ partitions keep
 *       engine 1's blocked multiply.  One CTA handles one lag block, or (batches averaging >= 1.5 template
 *       partitions) a pair of consecutive lag blocks that share their template rows, the second product
 *       spectrum waiting in tensor memory; both give bit-identical results;
 *   4 / 5 = engine 2 with pairs always / never;
 *   1 = the first fused lag-block kernel (sb_fused.cu; lag blocks of 8192 or 16384, hop B or B/2);
 *   0 = the cuFFT-planned pipeline (any block size; kept as the cross-check and for odd block sizes).
 * Engines agree to float32 FFT rounding (~2e-7 of the curve), not bit for bit. */
int sb_set_engine(int engine);
int sb_get_engine(void);
/* Spectral multiply of engine 1: 0 (default) = per lag block inside the fused kernel, except for
 * queries whose template spans 12 or more partitions (>= 16 s at the default block size), which go
 * through the register-blocked multiply kernel (8 lag blocks share each template row); 1 = never
 * blocked, 2 = always blocked.  The route depends only on the query, not on the rest of the batch. */
int sb_set_premac_mode(int mode);
/* Overlap-save geometry of the fused engines: 1 (default) = hop B (half of each inverse FFT is valid
 * lags), 2 = hop B/2 (three quarters valid, but twice as many template partitions to multiply:
 * pays off only for templates shorter than B/2), 0 = chosen per batch by a cost rule.  Geometries
 * agree to float32 FFT rounding (~1e-7), not bit for bit, so a run should stick to one. */
int sb_set_hop_mode(int mode);
/* Body variant of the packed kernels (engines 2, 4, 5) on uint8 streams: 3 (default) = per run of 8 lags a lower
 * and an upper bound of the screening values from the run's largest correlation value and its exact head sums; only
 * the runs whose lower bound does not exceed the lag block's smallest upper bound can hold the minimum -- they leave
 * the match kernel as records (query, first lag, 8 correlation values) and a second kernel evaluates their lags in
 * fp64; window sums slide on the staged sample windows, per-query constants travel through shared memory, the
 * self-mirrored quad is prefetched; 1 = the first version (fp32 screening of every lag and fp64 evaluation of the
 * candidates inside the match kernel, running sums read from HBM).  Either way the result is the fp64 evaluation of
 * every lag that can be the minimum, so the two agree bit for bit (checked on the GPU by the test-suite); float32
 * streams and the other engines ignore the setting.  (2, the per-lag loop of 3 run over all lags, was measured and
 * dropped in round 2.) */
int sb_set_epilogue(int variant);
int sb_get_epilogue(void);
/* Lag blocks processed per multiply / inverse-FFT / normalise launch (>= 1). */
int sb_set_chunk_items(int items);
/* Template partition spectra kept resident per pass over a batch (>= 1); batches needing more are
 * processed in several passes of whole queries. */
int sb_set_max_parts(int64_t parts);

/* The library's CUDA stream (a cudaStream_t) so that a host framework can order its own
 * work (e.g. an NCCL broadcast issued through torch.distributed) on the same stream. */
void* sb_get_stream(void);
/* Page-locked host buffers for callers that want true asynchronous H2D/D2H copies. */
int sb_pinned_alloc(int64_t bytes, void** out);
int sb_pinned_free(void* p);

/* Plain device buffers for callers that keep results on the GPU (sb_find_batch_device). */
int sb_device_alloc(int64_t bytes, void** out);
int sb_device_free(void* p);
int sb_copy_to_host(void* host_dst, const void* dev_src, int64_t bytes);   /* ordered on the library stream, blocking */
int sb_copy_to_device(void* dev_dst, const void* host_src, int64_t bytes); /* ordered on the library stream, blocking */
int sb_copy_on_device(void* dev_dst, const void* dev_src, int64_t bytes);  /* enqueued on the library stream */

/* ---- resident streams:  WavStream.data  (wav.py:119,140-156) ----------- */

/* Upload a normalised stream of n samples (dtype SB_U8 or SB_F32) from host
 * memory, build its running sums (the integral image cv2 builds per call,
 * wav.py:185) and keep it resident.  */
int sb_stream_create(const void* host_samples, int64_t n, int dtype, sb_stream** out);
/* Same, but `dev_samples` is already a device pointer on the bound GPU (e.g.
 * the receive buffer of an NCCL broadcast); it is copied device-to-device. */
int sb_stream_create_device(const void* dev_samples, int64_t n, int dtype, sb_stream** out);
int sb_stream_destroy(sb_stream* s);
/* Device address of the raw samples of a resident stream (read-only for the caller). */
const void* sb_stream_device_ptr(const sb_stream* s);
int64_t sb_stream_length(const sb_stream* s);
int sb_stream_dtype(const sb_stream* s);
/* Copy samples [off, off+n) back to host (tests, WavStream.data mirror). */
int sb_stream_read(const sb_stream* s, int64_t off, int64_t n, void* host_out);

/* ---- the matcher:  WavStream.find_substream  (wav.py:177-188) ---------- */

/* One query whose template is a raw host array (pattern is "any (1,n)
 * ndarray", e.g. np.split halves, sushi.py:445).  dtype must equal the image
 * stream's.  Template bytes are uploaded inside the call. */
int sb_find(const sb_stream* image, const void* tmpl_host, int64_t tmpl_len,
            int64_t lag0, int64_t nlags, float* diff_out, int64_t* idx_out);

/* `count` independent queries; template q = tmpl[tmpl_off[q] .. +tmpl_len[q])
 * (get_substream as an (offset,length) descriptor: no bytes move,
 * wav.py:168-171), searched over image positions [lag0[q], lag0[q]+nlags[q]).
 * Requirements per query: tmpl_len>=1, nlags>=1,
 * tmpl_off+tmpl_len <= len(tmpl), lag0>=0, lag0+nlags-1+tmpl_len <= len(image).
 * image and tmpl may be the same stream.  Host arrays in, host arrays out. */
int sb_find_batch(const sb_stream* image, const sb_stream* tmpl, int64_t count,
                  const int64_t* tmpl_off, const int64_t* tmpl_len,
                  const int64_t* lag0, const int64_t* nlags,
                  float* diff_out, int64_t* idx_out);

/* Same queries (descriptor arrays still in HOST memory: they are planned on the
 * host), but the two result arrays are DEVICE pointers and the call only
 * enqueues work on the library stream (no host synchronisation; sb_sync()
 * waits).  Used by the multi-GPU path, where the per-rank results feed an NCCL
 * all-gather without a round trip through the host. */
int sb_find_batch_device(const sb_stream* image, const sb_stream* tmpl, int64_t count,
                         const int64_t* tmpl_off, const int64_t* tmpl_len,
                         const int64_t* lag0, const int64_t* nlags,
                         float* d_diff_out, int64_t* d_idx_out);

/* Whole curves of `count` queries, concatenated in query order into curves_out[sum(nlags)] (host).
 * Every lag is evaluated with the exact fp64 rule.  Because a value depends only on (template,
 * absolute position), a curve over a wider range answers any sub-range query exactly: the shift
 * solver uses this to precompute the next groups' searches in one launch (sushi_b200/shifts.py). */
int sb_match_curves(const sb_stream* image, const sb_stream* tmpl, int64_t count,
                    const int64_t* tmpl_off, const int64_t* tmpl_len,
                    const int64_t* lag0, const int64_t* nlags, float* curves_out);
/* The whole curve of one query (debug / parity tests): curve_out[nlags]. */
int sb_match_curve(const sb_stream* image, const sb_stream* tmpl,
                   int64_t tmpl_off, int64_t tmpl_len, int64_t lag0, int64_t nlags,
                   float* curve_out);

/* ---- the loader:  WavStream.__init__  (wav.py:64-91,108-156) ----------- */

/* Decode interleaved PCM (sample_width 2 = int16 LE, 3 = int24 LE top 16 bits,
 * wav.py:68-74), average `channels` channels (wav.py:80-90) and resample each
 * READ_CHUNK (= `framerate` frames) with OpenCV's INTER_NEAREST index map
 * (wav.py:125-137) into a padded float32 buffer of total_len samples whose
 * first `padding` samples and last `padding` samples repeat the edge values
 * (wav.py:140-141).  Output stays on the device behind `*out_f32` (a SB_F32
 * stream holding the UN-normalised samples). */
int sb_load_pcm(const void* pcm_host, int64_t frames, int channels, int sample_width,
                int framerate, int sample_rate, int64_t padding, int64_t total_len,
                sb_stream** out_f32);
/* Median-clip normalise (wav.py:145-151) and optionally quantise to uint8
 * (wav.py:153-156) a stream produced by sb_load_pcm; returns a new resident
 * stream of dtype `dtype` ready for sb_find*.  min3/max3 (3 x the medians) are
 * returned for the host mirror. */
int sb_normalise(const sb_stream* raw_f32, int dtype, sb_stream** out,
                 float* min3_out, float* max3_out);

/* ---- multi-GPU: events shard across ranks ---------------- */

/* One process per GPU.  The library owns an NCCL communicator (libnccl is opened with dlopen on the first
 * of these calls; single-GPU users never need it).  The path has exactly two collectives, both outside the
 * kernels: a broadcast of the normalised streams (the reference's WavStream.data, wav.py:119-156, which every
 * rank needs) and an all-gather of the per-event (diff, idx) pairs find_substream returns (wav.py:188).
 * Rendezvous is the caller's business: rank 0 obtains SB_COMM_ID_BYTES opaque bytes from sb_comm_unique_id,
 * hands them to the other ranks by any means, then every rank calls sb_comm_init (collective). */
#define SB_COMM_ID_BYTES 128
int sb_comm_unique_id(void* id_out);
int sb_comm_init(const void* id, int world_size, int rank);
int sb_comm_destroy(void);
int sb_comm_world_size(void);                 /* 1 without a communicator */
int sb_comm_rank(void);
int sb_comm_nccl_version(void);               /* e.g. 22703; -1 if libnccl cannot be opened */
/* Broadcast `bytes` bytes of device memory in place from `root`.  Runs on the library's communication
 * stream, ordered behind everything the library stream has been given so far; `slot` (0..3) names the
 * completion event.  sb_comm_wait(slot) makes the library stream wait for that broadcast only, so the
 * broadcast of one stream overlaps the running sums and spectra of another. */
int sb_comm_broadcast(void* dev_buf, int64_t bytes, int root, int slot);
int sb_comm_wait(int slot);
/* All-gather on the library stream: every rank contributes bytes_per_rank bytes of device memory,
 * dev_recv receives world_size * bytes_per_rank bytes in rank order. */
int sb_comm_all_gather(const void* dev_send, void* dev_recv, int64_t bytes_per_rank);
/* Blocking helpers for measurement: element-wise maximum over ranks of up to 32 host floats (device
 * times are reported as the maximum over ranks), and a barrier that returns once every rank's
 * library and communication streams have drained. */
int sb_comm_max_f32(float* host_inout, int count);
int sb_comm_barrier(void);

/* ---- measurement ------------------------------------------------------- */

/* Device timer on the library stream (CUDA events). */
int sb_timer_start(void);
int sb_timer_stop(float* ms_out);
/* Per-kernel accounting: when enabled every kernel class is bracketed by CUDA
 * events on the library stream; sb_profile_get returns accumulated device ms
 * and launch count for a kernel class name, sb_profile_names
 * a comma-separated list. Costs a little throughput while on. */
int sb_profile_enable(int on);
int sb_profile_reset(void);
int sb_profile_get(const char* name, double* ms_out, int64_t* launches_out);
const char* sb_profile_names(void);
/* Total kernels launched by this library since sb_init (or last reset). */
int64_t sb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SUSHI_B200_H */
