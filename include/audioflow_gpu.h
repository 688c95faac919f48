/*
 * audioflow_gpu.h -- C ABI of libaudioflow_gpu.so, the B200 (sm_100a) implementation of the
 * audio hot path of forfd8960/audio-flow-rs:
 *
 *     interleaved f32/i16 PCM -> downmix -> resample to 16 kHz -> 25 ms/10 ms frames
 *     -> Hann -> 512-pt STFT -> log-mel,  and energy VAD frame decisions.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference repo root).  The reference has no FFI of its own (plain Rust structs re-exported
 * at src-tauri/src/modules/audio/mod.rs:9-11), so these are exactly the functions a Rust
 * `extern "C"` block for that module would bind; INTEGRATION.md shows that binding.
 *
 * Conventions
 *   - plain C types only; every function returns an af_status (0 = ok) unless noted.
 *   - on failure a thread-local message is available from af_last_error(); the Rust shim
 *     turns it into AudioError::ResamplingFailed(msg) (src-tauri/src/error.rs:109-110).
 *   - handles are NOT thread-safe individually (they mirror `&mut self`); distinct handles
 *     may be used from distinct threads.
 *   - there is no CPU fallback: without a CUDA device every call fails with AF_ERR_NO_DEVICE.
 */
#ifndef AUDIOFLOW_GPU_H
#define AUDIOFLOW_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AF_API __attribute__((visibility("default")))
#else
#define AF_API
#endif

typedef enum af_status {
    AF_OK = 0,
    AF_ERR_INVALID = 1,           /* bad argument */
    AF_ERR_RESAMPLING_FAILED = 2, /* AudioError::ResamplingFailed (error.rs:109-110) */
    AF_ERR_CUDA = 3,              /* CUDA runtime failure (also surfaced as ResamplingFailed by the shim) */
    AF_ERR_NO_DEVICE = 4,         /* no CUDA device / not initialised */
    AF_ERR_CAPACITY = 5           /* caller-provided output buffer too small */
} af_status;

/* ---- library ------------------------------------------------------------------------- */
#define AF_MAX_GPUS 16
/* Selects GPU `device` for the calling THREAD, initialising the library's context on it at first use (device < 0:
 * keep the thread's selection / the process default).  A process may use several GPUs: objects remember the GPU they
 * were created on and every call that takes a handle runs there, whatever the calling thread has selected. */
AF_API int af_init(int device);
AF_API int af_current_device(void);             /* the calling thread's GPU, -1 before any af_init */
/* Makes `cuda_stream` (a cudaStream_t) the stream the library enqueues on for the calling thread's GPU wherever an
 * entry point has no stream argument of its own (sessions, sharded batches, the NULL-stream form of af_batch_run);
 * NULL restores the library's own non-blocking stream.  The caller keeps the stream alive. */
AF_API int af_set_stream(void *cuda_stream);
AF_API int af_shutdown(void);
AF_API size_t af_last_error(char *buf, size_t cap);   /* returns strlen of the full message */
AF_API int af_device_count(int *count);
AF_API const char *af_version(void);
/* Number of kernels this library has launched since af_init (bench.py's gpu_launches). */
AF_API uint64_t af_kernel_launch_count(void);
/* Debugging aid: cycles the fused kernel's warp roles spent waiting on each pipeline barrier since the last
 * call ([role][total, t0 .. t6], roles FFT / mel / VAD / resample).  All zero unless the library
 * was built with `make STATS=1`. */
AF_API int af_debug_pipe_stats(uint64_t out[32]);

/* pinned host memory for the host-buffer entry points (optional but much faster) */
AF_API int af_host_alloc(void **ptr, size_t bytes);
AF_API int af_host_free(void *ptr);

/* ---- AudioFrame::to_mono  (src-tauri/src/modules/audio/capture.rs:30-42) ---------------- */
/* host buffers; out needs ceil(n_samples / channels) floats.  channels == 1 -> copy. */
AF_API int af_to_mono(const float *samples, size_t n_samples, uint16_t channels, float *out, size_t out_cap,
                      size_t *n_out);

/* ---- AudioResampler  (src-tauri/src/modules/audio/resampler.rs:12-112, 169-178) -------- */
typedef struct af_resampler af_resampler;
/* AudioResampler::new (resampler.rs:32-57); create_48k_to_16k (:60-62) == create(48000,16000) */
AF_API int af_resampler_create(uint32_t input_rate, uint32_t output_rate, af_resampler **out);
AF_API void af_resampler_destroy(af_resampler *r);
/* AudioResampler::process (resampler.rs:71-93): equal rates -> copy; otherwise ONE rubato
 * FastFixedIn step: consumes exactly 128 frames (extra ignored), fewer than 128 ->
 * AF_ERR_RESAMPLING_FAILED.  Host buffers. */
AF_API int af_resampler_process(af_resampler *r, const float *input, size_t n, float *out, size_t out_cap,
                                size_t *n_out);
AF_API uint32_t af_resampler_input_rate(const af_resampler *r);   /* resampler.rs:96-98  */
AF_API uint32_t af_resampler_output_rate(const af_resampler *r);  /* resampler.rs:101-103 */
AF_API int af_resampler_needs_resampling(const af_resampler *r);  /* resampler.rs:106-108 */
AF_API size_t af_resampler_chunk_size(const af_resampler *r);     /* 128, or 0 for passthrough */
/* upper bound of the number of output frames produced for n_in input frames */
AF_API size_t af_resample_max_output(uint32_t input_rate, uint32_t output_rate, size_t n_in);
/* exact number of output frames of BatchResampler::process(all n_in) + flush() */
AF_API int af_resample_output_len(uint32_t input_rate, uint32_t output_rate, size_t n_in, size_t *n_out);

/* ---- BatchResampler  (src-tauri/src/modules/audio/resampler.rs:115-166) ---------------- */
typedef struct af_batch_resampler af_batch_resampler;
AF_API int af_batch_resampler_create(uint32_t input_rate, uint32_t output_rate, af_batch_resampler **out);
AF_API void af_batch_resampler_destroy(af_batch_resampler *b);
/* BatchResampler::process (:132-147): buffers input, emits the output of every complete
 * 128-frame chunk.  Equal rates: passthrough (the reference would loop forever there). */
AF_API int af_batch_resampler_process(af_batch_resampler *b, const float *input, size_t n, float *out,
                                      size_t out_cap, size_t *n_out);
/* BatchResampler::flush (:150-166): zero-pads the residual to one chunk and processes it. */
AF_API int af_batch_resampler_flush(af_batch_resampler *b, float *out, size_t out_cap, size_t *n_out);

/* ---- VoiceActivityDetector  (src-tauri/src/modules/audio/vad.rs) ---------------------- */
typedef struct af_vad_config {      /* VadConfig, vad.rs:21-32 */
    float threshold_db;
    float smoothing_factor;
    uint64_t silence_timeout_frames;
    uint64_t min_speech_frames;
} af_vad_config;

enum { AF_VAD_SILENCE = 0, AF_VAD_SPEECH = 1, AF_VAD_ENDING = 2 };   /* VadState, vad.rs:47-54 */

typedef struct af_vad af_vad;
AF_API void af_vad_config_default(af_vad_config *cfg);                 /* vad.rs:34-43 */
AF_API int af_vad_create(const af_vad_config *cfg, af_vad **out);      /* vad.rs:80-88 */
AF_API void af_vad_destroy(af_vad *v);
/* detect (vad.rs:97-154): one frame of any length -> post-transition state */
AF_API int af_vad_detect(af_vad *v, const float *frame, size_t n, uint8_t *state);
/* the same detector driven over many frames [f*hop, f*hop+frame_len) of one host signal in a
 * single launch; writes one state per frame */
AF_API int af_vad_detect_frames(af_vad *v, const float *samples, size_t n, uint32_t frame_len, uint32_t hop,
                                uint8_t *states, size_t states_cap, size_t *n_frames);
AF_API int af_vad_reset(af_vad *v);                                    /* vad.rs:179-184 */
AF_API int af_vad_state(const af_vad *v);                              /* vad.rs:187-189 */
AF_API float af_vad_energy_db(const af_vad *v);                        /* vad.rs:192-194 */
AF_API int af_vad_is_speaking(const af_vad *v);                        /* vad.rs:197-199 */
AF_API uint64_t af_vad_speech_frame_count(const af_vad *v);            /* vad.rs:202-204 */
AF_API float af_vad_smoothed_energy(const af_vad *v);                  /* the field behind energy_db */
/* mean-square frame energy exactly as calculate_energy (vad.rs:157-168), on the GPU */
AF_API int af_vad_frame_energy(const float *frame, size_t n, float *energy);

/* ---- PCM16 wire encode  (src-tauri/src/modules/network/websocket.rs:246-251) ---------- */
AF_API int af_pcm16_encode(const float *samples, size_t n, int16_t *out);
/* The wire payload itself: base64 (standard alphabet, '=' padding) of the little-endian PCM16 bytes -- the
 * "audio_base_64" field of WebSocketClient::send_audio / MessageBuilder::audio_message (websocket.rs:244-254, :338-348).
 * Writes af_pcm16_base64_len(n) = 4 * ceil(2 n / 3) characters (no terminator) into out[out_cap]. */
AF_API size_t af_pcm16_base64_len(size_t n_samples);
AF_API int af_pcm16_base64(const float *samples, size_t n, char *out, size_t out_cap, size_t *n_out);

/* ======================================================================================= */
/* Batched fast path: many independent streams through                                     */
/*   to_mono -> BatchResampler(all)+flush -> STFT/log-mel -> framed VAD                     */
/* in one launch sequence.  Results are identical to driving the per-object API above      */
/* stream by stream.                                                                       */
/* ======================================================================================= */
enum { AF_FMT_F32 = 0, AF_FMT_I16 = 1 };            /* sample format; i16 decodes as s / 32768 */
enum { AF_MEM_DEVICE = 0, AF_MEM_HOST = 1 };        /* where stream data and outputs live */

typedef struct af_pipeline_config {
    uint32_t n_mels;          /* 0 = no STFT/log-mel; else 1..128 (80 and 128 are the BASELINE configs) */
    float f_min, f_max;       /* mel band edges in Hz (0, 8000) */
    float log_floor;          /* log(max(mel, floor)), 1e-10 */
    uint32_t log10;           /* 0 natural log, 1 log10 */
    uint32_t vad_enable;      /* run the VAD */
    af_vad_config vad;
    uint32_t vad_frame_len;   /* 0 -> the STFT frames (400 / 160); else e.g. 320 / 320 (20 ms) */
    uint32_t vad_hop;
    uint32_t write_pcm;       /* write the resampled 16 kHz PCM */
    uint32_t pcm16;           /* batch runs: deliver the PCM as i16 wire samples, (x.clamp(-1,1) * 32767) as i16
                               * (websocket.rs:246-251): af_outputs.pcm then points at int16_t rows and pcm_stride counts
                               * int16_t elements (still a multiple of 4).  Halves the PCM bytes that leave the GPU. */
} af_pipeline_config;

AF_API void af_pipeline_config_default(af_pipeline_config *cfg);

typedef struct af_stream_desc {
    const void *data;         /* interleaved samples [frame][channel] */
    uint64_t n_samples;       /* TOTAL interleaved sample count (AudioFrame::samples.len()) */
    uint32_t sample_rate;     /* AudioFrame::sample_rate; output is always 16 kHz */
    uint16_t channels;        /* AudioFrame::channels */
    uint16_t format;          /* AF_FMT_* */
} af_stream_desc;

typedef struct af_vad_final {   /* detector state after the last frame of a stream */
    float smoothed_energy;
    int32_t state;
    uint64_t silence_frames;
    uint64_t speech_frames;
} af_vad_final;

typedef struct af_outputs {
    float *pcm;               /* [n_streams][pcm_stride]    resampled mono 16 kHz (or NULL); device pointers must be
                               * 16-byte aligned (vector stores), else AF_ERR_INVALID */
    uint64_t pcm_stride;      /* floats per row, multiple of 4 */
    float *logmel;            /* [n_streams][logmel_stride] rows of [frame][mel] (or NULL) */
    uint64_t logmel_stride;   /* floats per row, multiple of 4 */
    uint8_t *vad;             /* [n_streams][vad_stride]    VadState per frame (or NULL) */
    uint64_t vad_stride;
    float *energy;            /* [n_streams][energy_stride] mean-square energy per VAD frame (or NULL) */
    uint64_t energy_stride;
    af_vad_final *vad_final;  /* [n_streams] (or NULL) */
} af_outputs;

typedef struct af_pipeline af_pipeline;
typedef struct af_batch af_batch;

AF_API int af_pipeline_create(const af_pipeline_config *cfg, af_pipeline **out);
AF_API void af_pipeline_destroy(af_pipeline *p);

/* Plans a batch: validates the descriptors, computes per-stream output lengths with the exact
 * chunk recurrence of the reference resampler, builds the device-side stream/tile tables.
 * mem = AF_MEM_DEVICE: desc.data are device pointers (16-byte aligned), the batch can be run
 * many times.  mem = AF_MEM_HOST: desc.data are host pointers, used by af_batch_run_host. */
AF_API int af_batch_create(af_pipeline *p, const af_stream_desc *streams, size_t n_streams, int mem,
                           af_batch **out);
AF_API void af_batch_destroy(af_batch *b);
AF_API size_t af_batch_n_streams(const af_batch *b);
/* per-stream counts (host arrays of n_streams): resampled length, STFT frames, VAD frames */
AF_API int af_batch_counts(const af_batch *b, uint32_t *n_out, uint32_t *n_feat_frames, uint32_t *n_vad_frames);
/* minimal strides (already rounded to the required multiples) */
AF_API int af_batch_strides(const af_batch *b, uint64_t *pcm_stride, uint64_t *logmel_stride,
                            uint64_t *vad_stride);

/* Device-resident run: outputs are device pointers.  Enqueues on `cuda_stream`
 * (a cudaStream_t passed as void*, NULL = the library's own stream) and returns without
 * synchronising when cuda_stream != NULL.  The library's own stream is non-blocking: with cuda_stream == NULL the
 * caller must have finished (synchronised) whatever produced the inputs or still touches the output buffers on
 * other streams, the legacy default stream included. */
AF_API int af_batch_run(af_batch *b, const af_outputs *out, void *cuda_stream);
/* Host-buffer run (the reference-facing call): copies the streams H2D (chunked, overlapped
 * with compute), runs the pipeline and copies the requested outputs back; blocking. */
AF_API int af_batch_run_host(af_batch *b, const af_outputs *out);
/* One-shot convenience: create + run_host + destroy. */
AF_API int af_pipeline_run(af_pipeline *p, const af_stream_desc *streams, size_t n_streams,
                           const af_outputs *out);

/* which variant of the fused kernel to launch (default "auto"); for tests and ncu comparisons.
 * names: "auto", "sync" (plain loads), "tma" (bulk-copy staged input).  Returns AF_ERR_INVALID
 * for unknown names. */
AF_API int af_set_kernel_variant(const char *name);

/* ---- VAD segmentation (consumer of VadState) --------------------------- */
/* device buffers: states [n_streams][vad_stride] -> seg [n_streams][seg_cap][2] = [start,end)
 * frames, n_seg [n_streams].  A segment runs from the first Speech frame to the frame that
 * reports Ending (inclusive) or that falls back to Silence (exclusive). */
AF_API int af_vad_segments(const uint8_t *states, uint64_t vad_stride, const uint32_t *n_frames,
                           size_t n_streams, uint32_t *seg, uint32_t seg_cap, uint32_t *n_seg,
                           void *cuda_stream);

/* ---- VAD-gated output (second half) ------------------------------------- */
/* What the reference's spec wants in front of the wire (specs/0001-spec.md:466: drop the silent stretches before
 * ScribeClient::send_audio, src-tauri/src/modules/network/scribe_client.rs:194-196): only the audio and the feature
 * rows of the speech segments, packed.  Frame f of a segment [start, end) contributes its hop -- samples
 * [f*hop, (f+1)*hop) of the 16 kHz row -- and log-mel row f; the kept frames of a stream are stored back to back in
 * order.  seg_offset[s][k] is the compacted frame index where segment k starts (PCM offset = that times hop),
 * seg_offset[s][n_seg] == n_frames[s] is the number of kept frames.  All pointers are device pointers;
 * seg / n_seg are af_vad_segments' outputs (segments beyond seg_cap are ignored). */
typedef struct af_gate_outputs {
    float *pcm;               /* [n_streams][pcm_stride]    gated 16 kHz samples (or NULL) */
    uint64_t pcm_stride;
    float *logmel;            /* [n_streams][logmel_stride] gated rows of [frame][mel] (or NULL) */
    uint64_t logmel_stride;
    uint32_t *seg_offset;     /* [n_streams][seg_cap + 1] */
    uint32_t *n_frames;       /* [n_streams] kept frames */
} af_gate_outputs;
AF_API int af_vad_gate(const float *pcm, uint64_t pcm_stride, const float *logmel, uint64_t logmel_stride, uint32_t n_mels,
                       const uint32_t *n_out, uint32_t hop, const uint32_t *seg, uint32_t seg_cap, const uint32_t *n_seg,
                       size_t n_streams, const af_gate_outputs *out, void *cuda_stream);

/* ======================================================================================= */
/* Multi-GPU: independent streams sharded over the GPUs of one box, results gathered with NCCL  */
/* (SURVEY 8(e)).  Streams never talk to each other, so the data path has NO collective; the only */
/* exchange is the gather of the per-stream VAD states to every GPU, issued on a side stream so    */
/* that the next batch never waits for it.  Two ways to form the communicator:                     */
/*   - af_init_multi(n): ONE process (e.g. the Rust host) drives GPUs 0..n-1 (ncclCommInitAll);    */
/*   - one process per GPU: af_init(local_gpu), rank 0 calls af_comm_unique_id, the host ships the  */
/*     128 bytes to the other ranks by any channel, every rank calls af_comm_init_rank.             */
/* NCCL is loaded at run time (libnccl.so.2); single-GPU use never touches it.                      */
/* ======================================================================================= */
AF_API int af_init_multi(int n_gpus);
AF_API int af_comm_unique_id(uint8_t id[128]);
AF_API int af_comm_init_rank(int n_ranks, int rank, const uint8_t id[128]);
AF_API int af_comm_size(void);                  /* ranks of the communicator, 0 when there is none */
AF_API int af_comm_shutdown(void);
/* Contiguous blocks of stream indices, balanced by input BYTES (44.1 kHz and 48 kHz streams differ by 8 %):
 * shard r owns streams [first[r], first[r + 1]); first has n_shards + 1 entries.  Pure host code. */
AF_API int af_shard_partition(const af_stream_desc *streams, size_t n_streams, int n_shards, size_t *first);

typedef struct af_sharded_batch af_sharded_batch;
typedef struct af_sharded_outputs {             /* per rank; only the entries of this process's ranks are read */
    af_outputs shard[AF_MAX_GPUS];              /* device pointers on the rank's GPU, rows = the rank's streams */
} af_sharded_outputs;
/* Plans n_streams GLOBAL streams over the communicator's ranks (af_shard_partition) and builds the batches of the
 * ranks this process owns (all of them after af_init_multi, one after af_comm_init_rank).  Every process passes the
 * same descriptor list; the `data` of streams owned elsewhere is not used (may be NULL).  mem as in af_batch_create;
 * AF_MEM_DEVICE data must live on the owner's GPU. */
AF_API int af_sharded_batch_create(af_pipeline *p, const af_stream_desc *streams, size_t n_streams, int mem,
                                   af_sharded_batch **out);
AF_API void af_sharded_batch_destroy(af_sharded_batch *b);
/* rank r owns streams [*first, *first + *count); *device = its GPU when this process owns the rank, else -1 */
AF_API int af_sharded_batch_shard(const af_sharded_batch *b, int rank, size_t *first, size_t *count, int *device);
AF_API af_batch *af_sharded_batch_local(af_sharded_batch *b, int rank);   /* the rank's ordinary batch (counts, strides); NULL if remote */
/* Runs every local shard on its GPU (all GPUs concurrently, nothing blocks in between).  gather != 0: the VAD kernel
 * of each shard writes its states straight into the rank's slot of a double-buffered gather buffer and ONE in-place
 * ncclAllGather per GPU follows on the side stream; shard[r].vad is then ignored.  async == 0: returns when every GPU
 * (and the gather) is done; else af_sharded_batch_wait does that. */
AF_API int af_sharded_batch_run(af_sharded_batch *b, const af_sharded_outputs *out, int gather, int async);
AF_API int af_sharded_batch_wait(af_sharded_batch *b);
/* Orders the compute stream of every local GPU after the last gather WITHOUT blocking the host: work (or an event)
 * enqueued afterwards sees the gathered states. */
AF_API int af_sharded_batch_join(af_sharded_batch *b);
/* The gathered states of the LAST run on a local rank's GPU: row (r * rows_per_rank + i) holds stream i of rank r,
 * n_vad_frames[global stream] of them valid (host array of n_streams, may be NULL).  The buffer is reused by the run
 * after the next one. */
AF_API int af_sharded_batch_gathered(af_sharded_batch *b, int rank, const uint8_t **states, uint64_t *row_stride,
                                     uint64_t *rows_per_rank, uint32_t *n_vad_frames);
/* The same result on the host, in GLOBAL stream order: states[n_streams][stride] (stride >= the longest stream's
 * frame count).  Waits for the gather. */
AF_API int af_sharded_batch_gathered_host(af_sharded_batch *b, int rank, uint8_t *states, uint64_t stride);
/* Device time of the last gather on a local rank (side stream, event to event), milliseconds. */
AF_API int af_sharded_batch_gather_ms(af_sharded_batch *b, int rank, float *ms);
/* Host-buffer form (mem == AF_MEM_HOST): `out` holds rows for ALL n_streams streams in host memory; every local shard
 * runs af_batch_run_host on its GPU from its own host thread, writing its rows.  Blocking. */
AF_API int af_sharded_batch_run_host(af_sharded_batch *b, const af_outputs *out);

/* ---- host-side planning diagnostics (no GPU needed; used by the CPU test-suite) ---------- */
/* smallest f32 energy e with 20*log10f(e) > threshold_db under the host libm (NaN if none):
 * the device compares energies against this instead of taking a log (VAD). */
AF_API float af_debug_vad_energy_threshold(float threshold_db);
/* the f32 fractional offsets the host plan derives for the first n_chunks 128-frame chunks of
 * an (input_rate -> output_rate) stream; returns the number of outputs (may exceed cap). */
AF_API size_t af_debug_resample_plan(uint32_t input_rate, uint32_t output_rate, size_t n_chunks, float *frac,
                                     size_t cap, int *mode);

/* ---- streaming sessions (persistent per-stream state; BASELINE config 5) ---------------- */
typedef struct af_session af_session;
/* n_streams concurrent streams of one (rate, channels, format); state (resampler position,
 * residual input, STFT overlap, VAD) stays on the device between ticks. */
AF_API int af_session_create(af_pipeline *p, size_t n_streams, uint32_t sample_rate, uint16_t channels,
                             uint16_t format, uint32_t max_tick_samples, af_session **out);
AF_API void af_session_destroy(af_session *s);
/* One tick: every stream pushes `n_samples` interleaved samples (device or host per `mem`,
 * rows `in_stride` samples apart).  Outputs produced by this tick are appended at the row
 * starts of `out`; counts (host arrays of n_streams, may be NULL) receive how many
 * PCM samples / feature frames / VAD frames each stream emitted. */
AF_API int af_session_push(af_session *s, const void *data, uint64_t in_stride, uint32_t n_samples, int mem,
                           const af_outputs *out, uint32_t *n_pcm, uint32_t *n_feat, uint32_t *n_vad);
AF_API int af_session_reset(af_session *s);

/* ---- level metering for the UI events ---------------------------------------------------- */
/* AudioLevel{level, peak} (src-tauri/src/events/mod.rs:41,73-75) and VolumeLevel{level, is_speech}
 * (src-tauri/src/modules/events/mod.rs:22,182-185) as by-products of a session tick.  The reference only declares
 * the payloads; the values are defined here: level_db = VoiceActivityDetector::energy_db() (vad.rs:192-194:
 * 20*log10 of the smoothed energy after the tick's last frame, -inf when it is <= 0 or the VAD is off),
 * peak = max |y| over the 16 kHz samples the tick produced (0 if none), is_speech = is_speaking() (vad.rs:197-199).
 * Enable once, then read after every push; host arrays of n_streams, any may be NULL. */
AF_API int af_session_enable_levels(af_session *s, int enable);
AF_API int af_session_levels(const af_session *s, float *level_db, float *peak, uint8_t *is_speech);

/* ---- RingBuffer  (src-tauri/src/modules/audio/capture.rs:84-161): capture hand-off ------ */
/* The cpal callback writes captured f32 samples, the processing task reads them.  Same semantics as the
 * reference: one slot stays free (capacity - 1 usable), a write drops what does not fit and returns the
 * count written, a read returns min(size, available).  The storage is pinned host memory when a device is
 * bound, so a session tick copies straight out of it.  Calls lock like the reference's Mutex<Vec<f32>>. */
typedef struct af_ring af_ring;
#define AF_RING_EMPTY (-1) /* af_ring_read: nothing available -- the reference's `None`, not an error */
AF_API int af_ring_create(size_t capacity_samples, af_ring **out);       /* RingBuffer::new  capture.rs:91-99 */
AF_API void af_ring_destroy(af_ring *r);
AF_API size_t af_ring_capacity(const af_ring *r);
AF_API size_t af_ring_write(af_ring *r, const float *data, size_t n);    /* RingBuffer::write capture.rs:101-121 */
AF_API int af_ring_read(af_ring *r, float *out, size_t size, size_t *n_read);   /* RingBuffer::read capture.rs:123-147 */
AF_API size_t af_ring_available(const af_ring *r);                       /* capture.rs:149-154 */
AF_API void af_ring_clear(af_ring *r);                                   /* capture.rs:156-161 */
/* AudioCapturer::read_frame (capture.rs:310-319) for every stream of a session at once: takes exactly
 * n_samples from each of the n_streams rings and pushes them as one tick (af_session_push).  Nothing is
 * consumed unless every ring holds n_samples. */
AF_API int af_session_push_rings(af_session *s, af_ring *const *rings, uint32_t n_samples, const af_outputs *out,
                                 uint32_t *n_pcm, uint32_t *n_feat, uint32_t *n_vad);

#ifdef __cplusplus
}
#endif
#endif /* AUDIOFLOW_GPU_H */
