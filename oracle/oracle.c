/*
 * oracle.c -- CPU restatement of the audio hot path of forfd8960/audio-flow-rs.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under audio-flow-rs_b200/ (the product) may
 * include, link, import or call this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker or
 * as the timed CPU baseline -- never as the thing shipped.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   (A) reference-pinned   : orc_to_mono, orc_vad_*   -- pinned by the reference's own
 *                            unit tests (capture.rs:372-400, vad.rs:212-298), carried
 *                            verbatim in tests/golden/reference_kat.json.
 *   (B) PARITY UNPINNED    : orc_resampler_* / orc_batch_*  -- arithmetic lives in the
 *                            third-party crate rubato 0.16.2 (Cargo.lock:4198-4201),
 *                            which is NOT under /root/reference and cannot be built here
 *                            (no Rust toolchain).  Restated from rubato's published
 *                            FastFixedIn/interp_cubic algorithm; anchored on the
 *                            reference call sites resampler.rs:43-49, :84, :132-166 and
 *                            the one reference test that touches it (passthrough,
 *                            resampler.rs:185-190).
 *   (C) spec-defined       : orc_logmel, orc_frame_energy -- the reference has no
 *                            STFT/mel code at all; the spec is written in DESIGN.md and
 *                            this file is its definition.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off matters: Rust never fuses a*b+c, so neither may this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Downmix: AudioFrame::to_mono, capture.rs:30-42                             */
/*   chunks(C).map(|c| c.iter().sum::<f32>() / C as f32)                      */
/*   - sequential left-to-right sum starting from 0.0                         */
/*   - a trailing partial chunk is summed and still divided by C              */
/*   - C == 1 -> clone                                                        */
/* ------------------------------------------------------------------------- */
ORC_API size_t orc_to_mono(const float *in, size_t n_samples, unsigned channels, float *out)
{
    if (channels <= 1) {
        memcpy(out, in, n_samples * sizeof(float));
        return n_samples;
    }
    size_t n_out = 0;
    for (size_t i = 0; i < n_samples; i += channels) {
        size_t m = n_samples - i < channels ? n_samples - i : channels;
        float sum = 0.0f;
        for (size_t c = 0; c < m; ++c)
            sum = sum + in[i + c];
        out[n_out++] = sum / (float)channels;
    }
    return n_out;
}

/* i16 capture decode (north-star input format; the reference captures f32 only,
 * capture.rs:266-268).  Spec: x = s / 32768 (exact in f32). */
ORC_API void orc_i16_to_f32(const int16_t *in, size_t n, float *out)
{
    for (size_t i = 0; i < n; ++i)
        out[i] = (float)in[i] / 32768.0f;
}

/* ------------------------------------------------------------------------- */
/* rubato 0.16.2 FastFixedIn<f32>, PolynomialDegree::Cubic  (asynchro_fast.rs) */
/* as constructed at resampler.rs:43-49 with the INTENDED arguments            */
/* (ratio = out/in, 1 channel, chunk 128); the literal argument order there    */
/* makes every call fail and is not replicated.                 */
/* ------------------------------------------------------------------------- */
#define POLY_LEN 8           /* rubato POLYNOMIAL_LEN_U */
#define CHUNK 128            /* resampler.rs:47,55 */

typedef struct {
    int passthrough;         /* resampler.rs:33-40: equal rates -> resampler None */
    uint32_t in_rate, out_rate;
    double resample_ratio;   /* out/in */
    double last_index;       /* starts at -(POLY_LEN/2) = -4.0 */
    float buffer[CHUNK + 2 * POLY_LEN];
} orc_resampler;

/* rubato interp_cubic: points at x = -1, 0, 1, 2 */
static inline float interp_cubic(float x, const float *y)
{
    float a0 = y[1];
    float a1 = -(1.0f / 3.0f) * y[0] - 0.5f * y[1] + y[2] - (1.0f / 6.0f) * y[3];
    float a2 = 0.5f * (y[0] + y[2]) - y[1];
    float a3 = 0.5f * (y[1] - y[2]) + (1.0f / 6.0f) * (y[3] - y[0]);
    float x2 = x * x;
    float x3 = x2 * x;
    return a0 + a1 * x + a2 * x2 + a3 * x3;
}

ORC_API orc_resampler *orc_resampler_new(uint32_t in_rate, uint32_t out_rate)
{
    orc_resampler *r = (orc_resampler *)calloc(1, sizeof(*r));
    r->in_rate = in_rate;
    r->out_rate = out_rate;
    r->passthrough = (in_rate == out_rate);
    r->resample_ratio = (double)out_rate / (double)in_rate;
    r->last_index = -(double)(POLY_LEN / 2);
    return r;
}
ORC_API void orc_resampler_free(orc_resampler *r) { free(r); }
ORC_API size_t orc_resampler_chunk_size(const orc_resampler *r) { return r->passthrough ? 0 : CHUNK; }

/* One FastFixedIn::process_into_buffer step on exactly CHUNK frames.
 * If frac_out != NULL, the f32 fractional offsets are recorded too (used by tests
 * to check the product's host-side resample plan). */
static size_t fastfixedin_step(orc_resampler *r, const float *in, float *out, float *frac_out)
{
    /* shift the last 2*POLY_LEN samples to the front, append the new chunk */
    memmove(r->buffer, r->buffer + CHUNK, 2 * POLY_LEN * sizeof(float));
    memcpy(r->buffer + 2 * POLY_LEN, in, CHUNK * sizeof(float));

    double t_ratio = 1.0 / r->resample_ratio;
    double t_ratio_end = 1.0 / r->resample_ratio;           /* target_ratio == resample_ratio */
    double approx_frames = (double)CHUNK * (0.5 * r->resample_ratio + 0.5 * r->resample_ratio);
    double t_ratio_increment = (t_ratio_end - t_ratio) / approx_frames;   /* == 0.0 */
    long end_idx = (long)CHUNK - (POLY_LEN + 1) - (long)ceil(t_ratio_end);

    double idx = r->last_index;
    size_t n = 0;
    while (idx < (double)end_idx) {
        t_ratio += t_ratio_increment;
        idx += t_ratio;
        double idx_floor = floor(idx);
        long start_idx = (long)idx_floor - 1;
        float frac = (float)(idx - idx_floor);
        const float *p = r->buffer + (start_idx + 2 * POLY_LEN);
        out[n] = interp_cubic(frac, p);
        if (frac_out) frac_out[n] = frac;
        n++;
    }
    r->last_index = idx - (double)CHUNK;
    return n;
}

/* rubato sizes its output buffer as chunk * ratio + 10 frames */
static size_t max_out_per_chunk(const orc_resampler *r)
{
    return (size_t)((double)CHUNK * r->resample_ratio + 10.0) + 8;
}

/* AudioResampler::process, resampler.rs:71-93.
 * returns 0 ok, 1 = ResamplingFailed (fewer than 128 input frames: rubato's
 * InsufficientInputBufferSize), 2 = output capacity too small (caller bug). */
ORC_API int orc_resampler_process(orc_resampler *r, const float *in, size_t n, float *out, size_t cap,
                                  size_t *n_out)
{
    if (r->passthrough) {
        if (cap < n) return 2;
        memcpy(out, in, n * sizeof(float));
        *n_out = n;
        return 0;
    }
    if (n < CHUNK) return 1;
    float *tmp = (float *)malloc(max_out_per_chunk(r) * sizeof(float));
    size_t m = fastfixedin_step(r, in, tmp, NULL);   /* extra input beyond 128 is ignored */
    if (m > cap) { free(tmp); return 2; }
    memcpy(out, tmp, m * sizeof(float));
    free(tmp);
    *n_out = m;
    return 0;
}

/* BatchResampler, resampler.rs:115-166 */
typedef struct {
    orc_resampler *rs;
    float *buf;
    size_t len, cap;
} orc_batch;

ORC_API orc_batch *orc_batch_new(uint32_t in_rate, uint32_t out_rate)
{
    orc_batch *b = (orc_batch *)calloc(1, sizeof(*b));
    b->rs = orc_resampler_new(in_rate, out_rate);
    return b;
}
ORC_API void orc_batch_free(orc_batch *b)
{
    if (!b) return;
    orc_resampler_free(b->rs);
    free(b->buf);
    free(b);
}

/* BatchResampler::process, resampler.rs:132-147.  Equal rates: the reference loops forever
 * (chunk_size == 0 => `while len >= 0`); documented deviation: passthrough. */
ORC_API int orc_batch_process(orc_batch *b, const float *in, size_t n, float *out, size_t cap, size_t *n_out)
{
    *n_out = 0;
    if (b->rs->passthrough) {
        if (cap < n) return 2;
        memcpy(out, in, n * sizeof(float));
        *n_out = n;
        return 0;
    }
    if (b->len + n > b->cap) {
        b->cap = (b->len + n) * 2 + CHUNK;
        b->buf = (float *)realloc(b->buf, b->cap * sizeof(float));
    }
    memcpy(b->buf + b->len, in, n * sizeof(float));
    b->len += n;
    size_t pos = 0, w = 0;
    float *tmp = (float *)malloc(max_out_per_chunk(b->rs) * sizeof(float));
    while (b->len - pos >= CHUNK) {
        size_t m = fastfixedin_step(b->rs, b->buf + pos, tmp, NULL);
        if (w + m > cap) { free(tmp); return 2; }
        memcpy(out + w, tmp, m * sizeof(float));
        w += m;
        pos += CHUNK;
    }
    free(tmp);
    memmove(b->buf, b->buf + pos, (b->len - pos) * sizeof(float));   /* drain(..128) */
    b->len -= pos;
    *n_out = w;
    return 0;
}

/* BatchResampler::flush, resampler.rs:150-166: zero-pad the residual to one chunk. */
ORC_API int orc_batch_flush(orc_batch *b, float *out, size_t cap, size_t *n_out)
{
    *n_out = 0;
    if (b->len == 0) return 0;
    if (b->rs->passthrough) { b->len = 0; return 0; }
    float chunk[CHUNK];
    memset(chunk, 0, sizeof(chunk));
    memcpy(chunk, b->buf, b->len * sizeof(float));
    float *tmp = (float *)malloc(max_out_per_chunk(b->rs) * sizeof(float));
    size_t m = fastfixedin_step(b->rs, chunk, tmp, NULL);
    if (m > cap) { free(tmp); return 2; }
    memcpy(out, tmp, m * sizeof(float));
    free(tmp);
    *n_out = m;
    b->len = 0;
    return 0;
}

/* Whole-stream convenience: BatchResampler::process(all) followed by flush().
 * frac_out (optional, same capacity as out) receives the f32 fractional offsets. */
ORC_API size_t orc_resample_stream(uint32_t in_rate, uint32_t out_rate, const float *in, size_t n,
                                   float *out, size_t cap, float *frac_out)
{
    if (in_rate == out_rate) {
        size_t m = n < cap ? n : cap;
        memcpy(out, in, m * sizeof(float));
        return m;
    }
    orc_resampler *r = orc_resampler_new(in_rate, out_rate);
    size_t w = 0;
    float chunk[CHUNK];
    float *tmp = (float *)malloc(2 * max_out_per_chunk(r) * sizeof(float));
    float *ftmp = tmp + max_out_per_chunk(r);
    for (size_t pos = 0; pos < n; pos += CHUNK) {
        const float *src = in + pos;
        if (n - pos < CHUNK) {
            memset(chunk, 0, sizeof(chunk));
            memcpy(chunk, in + pos, (n - pos) * sizeof(float));
            src = chunk;
        }
        size_t m = fastfixedin_step(r, src, tmp, ftmp);
        if (w + m > cap) m = cap - w;
        memcpy(out + w, tmp, m * sizeof(float));
        if (frac_out) memcpy(frac_out + w, ftmp, m * sizeof(float));
        w += m;
    }
    free(tmp);
    orc_resampler_free(r);
    return w;
}

/* Upper bound on the output count for n input frames. */
ORC_API size_t orc_resample_max_output(uint32_t in_rate, uint32_t out_rate, size_t n)
{
    if (in_rate == out_rate) return n;
    size_t chunks = (n + CHUNK - 1) / CHUNK;
    double per = (double)CHUNK * (double)out_rate / (double)in_rate;
    return (size_t)((double)chunks * per) + chunks + 64;
}

/* ------------------------------------------------------------------------- */
/* VAD: vad.rs:21-204                                                         */
/* ------------------------------------------------------------------------- */
typedef struct {
    float threshold_db;             /* vad.rs:24 */
    float smoothing_factor;         /* vad.rs:27 */
    uint64_t silence_timeout_frames;/* vad.rs:29 */
    uint64_t min_speech_frames;     /* vad.rs:31 */
} orc_vad_config;

enum { ORC_SILENCE = 0, ORC_SPEECH = 1, ORC_ENDING = 2 };   /* vad.rs:47-54 */

typedef struct {
    orc_vad_config cfg;
    float smoothed_energy;
    uint64_t silence_frames;
    uint64_t speech_frames;
    int state;
} orc_vad;

ORC_API void orc_vad_default_config(orc_vad_config *c)   /* vad.rs:34-43 */
{
    c->threshold_db = -50.0f;
    c->smoothing_factor = 0.3f;
    c->silence_timeout_frames = 15;
    c->min_speech_frames = 3;
}

ORC_API orc_vad *orc_vad_new(const orc_vad_config *c)     /* vad.rs:80-88 */
{
    orc_vad *v = (orc_vad *)calloc(1, sizeof(*v));
    v->cfg = *c;
    v->state = ORC_SILENCE;
    return v;
}
ORC_API void orc_vad_free(orc_vad *v) { free(v); }

/* calculate_energy, vad.rs:157-168: MEAN SQUARE (the comments say RMS). */
ORC_API float orc_frame_energy(const float *frame, size_t n)
{
    if (n == 0) return 0.0f;
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i)
        sum = sum + frame[i] * frame[i];
    return sum / (float)n;
}

/* energy_to_dbfs, vad.rs:171-176 */
ORC_API float orc_energy_to_dbfs(float e)
{
    if (e <= 0.0f) return -INFINITY;
    return 20.0f * log10f(e);
}

/* detect() on a precomputed frame energy: vad.rs:101-153 */
ORC_API int orc_vad_detect_energy(orc_vad *v, float energy)
{
    float old = v->smoothed_energy;
    v->smoothed_energy = v->cfg.smoothing_factor * energy + (1.0f - v->cfg.smoothing_factor) * old;
    float det = v->cfg.smoothing_factor > 0.0f ? v->smoothed_energy : energy;
    float dbfs = orc_energy_to_dbfs(det);
    int is_speech = dbfs > v->cfg.threshold_db;

    switch (v->state) {
    case ORC_SILENCE:
        if (is_speech) {
            v->speech_frames = 1;
            v->silence_frames = 0;
            v->state = ORC_SPEECH;
        }
        break;
    case ORC_SPEECH:
        if (is_speech) {
            v->speech_frames += 1;
            v->silence_frames = 0;
        } else {
            v->silence_frames += 1;
            if (v->silence_frames >= v->cfg.silence_timeout_frames) {
                if (v->speech_frames >= v->cfg.min_speech_frames)
                    v->state = ORC_ENDING;
                else
                    v->state = ORC_SILENCE;
                v->speech_frames = 0;
            }
        }
        break;
    case ORC_ENDING:
        v->state = ORC_SILENCE;
        v->silence_frames = 0;
        break;
    }
    return v->state;
}

ORC_API int orc_vad_detect(orc_vad *v, const float *frame, size_t n)   /* vad.rs:97-154 */
{
    return orc_vad_detect_energy(v, orc_frame_energy(frame, n));
}

ORC_API void orc_vad_reset(orc_vad *v)     /* vad.rs:179-184 */
{
    v->smoothed_energy = 0.0f;
    v->silence_frames = 0;
    v->speech_frames = 0;
    v->state = ORC_SILENCE;
}
ORC_API int orc_vad_state(const orc_vad *v) { return v->state; }                           /* :187-189 */
ORC_API float orc_vad_energy_db(const orc_vad *v) { return orc_energy_to_dbfs(v->smoothed_energy); } /* :192-194 */
ORC_API int orc_vad_is_speaking(const orc_vad *v) { return v->state == ORC_SPEECH; }       /* :197-199 */
ORC_API uint64_t orc_vad_speech_frame_count(const orc_vad *v) { return v->speech_frames; } /* :202-204 */
/* the private field behind the timeout (vad.rs:72): no accessor in the reference; exposed for the state-exactness tests */
ORC_API uint64_t orc_vad_silence_frames(const orc_vad *v) { return v->silence_frames; }
ORC_API float orc_vad_smoothed_energy(const orc_vad *v) { return v->smoothed_energy; }

/* Framed VAD over a whole 16 kHz stream: frames [f*hop, f*hop+len), "valid" framing.
 * states[T] receives the post-transition state per frame, energies[T] (optional) the
 * mean-square energies.  Returns T. */
ORC_API size_t orc_vad_stream(orc_vad *v, const float *y, size_t n, size_t frame_len, size_t hop,
                              uint8_t *states, float *energies)
{
    if (n < frame_len || frame_len == 0 || hop == 0) return 0;
    size_t T = 1 + (n - frame_len) / hop;
    for (size_t f = 0; f < T; ++f) {
        float e = orc_frame_energy(y + f * hop, frame_len);
        if (energies) energies[f] = e;
        states[f] = (uint8_t)orc_vad_detect_energy(v, e);
    }
    return T;
}

/* ------------------------------------------------------------------------- */
/* Spec-defined features (no reference code exists): DESIGN.md "Feature spec" */
/*   window : periodic Hann, w[n] = 0.5 - 0.5 cos(2 pi n / win), f64 -> f32    */
/*   frame  : xw[n] = fl32(y[f*hop + n] * w[n]), zero-padded to n_fft          */
/*   STFT   : X[k] = sum_n xw[n] e^{-2 pi i k n / n_fft}, k = 0..n_fft/2       */
/*   power  : P[k] = Re^2 + Im^2                                               */
/*   mel    : HTK scale mel(f) = 2595 log10(1 + f/700); n_mels + 2 points      */
/*            equally spaced in mel on [fmin, fmax]; triangular weights        */
/*            evaluated at the bin centre frequencies, no area normalisation;  */
/*            weights rounded to f32                                           */
/*   log    : ln(max(mel, floor))   (log10 when log10_flag)                    */
/* The oracle evaluates DFT, power and mel sums in f64 from the f32 inputs     */
/* (xw and the f32 weights) and rounds the final log to f32.                   */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t sample_rate;   /* 16000 */
    uint32_t n_fft;         /* 512   */
    uint32_t win_length;    /* 400   */
    uint32_t hop_length;    /* 160   */
    uint32_t n_mels;        /* 80 / 128 */
    float f_min;            /* 0     */
    float f_max;            /* 8000  */
    float log_floor;        /* 1e-10 */
    uint32_t log10_flag;    /* 0: natural log */
} orc_feat_config;

ORC_API void orc_feat_default_config(orc_feat_config *c)
{
    c->sample_rate = 16000;
    c->n_fft = 512;
    c->win_length = 400;
    c->hop_length = 160;
    c->n_mels = 80;
    c->f_min = 0.0f;
    c->f_max = 8000.0f;
    c->log_floor = 1e-10f;
    c->log10_flag = 0;
}

ORC_API void orc_hann_window(uint32_t win, float *w)
{
    for (uint32_t n = 0; n < win; ++n)
        w[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)win));
}

static double hz_to_mel(double f) { return 2595.0 * log10(1.0 + f / 700.0); }
static double mel_to_hz(double m) { return 700.0 * (pow(10.0, m / 2595.0) - 1.0); }

/* fb[k * n_mels + m], k = 0..n_fft/2 */
ORC_API void orc_mel_filterbank(const orc_feat_config *c, float *fb)
{
    uint32_t n_bins = c->n_fft / 2 + 1, M = c->n_mels;
    double m_lo = hz_to_mel(c->f_min), m_hi = hz_to_mel(c->f_max);
    double *pts = (double *)malloc((M + 2) * sizeof(double));
    for (uint32_t i = 0; i < M + 2; ++i)
        pts[i] = mel_to_hz(m_lo + (m_hi - m_lo) * (double)i / (double)(M + 1));
    for (uint32_t k = 0; k < n_bins; ++k) {
        double f = (double)k * (double)c->sample_rate / (double)c->n_fft;
        for (uint32_t m = 0; m < M; ++m) {
            double up = (f - pts[m]) / (pts[m + 1] - pts[m]);
            double dn = (pts[m + 2] - f) / (pts[m + 2] - pts[m + 1]);
            double w = up < dn ? up : dn;
            if (w < 0.0) w = 0.0;
            fb[(size_t)k * M + m] = (float)w;
        }
    }
    free(pts);
}

/* radix-2 f64 FFT, in place, n power of two */
static void fft_f64(double *re, double *im, uint32_t n)
{
    for (uint32_t i = 1, j = 0; i < n; ++i) {
        uint32_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (uint32_t len = 2; len <= n; len <<= 1) {
        uint32_t half = len >> 1;
        for (uint32_t i = 0; i < n; i += len) {
            for (uint32_t j = 0; j < half; ++j) {
                double ang = -2.0 * M_PI * (double)j / (double)len;
                double wr = cos(ang), wi = sin(ang);
                double ur = re[i + j], ui = im[i + j];
                double vr = re[i + j + half] * wr - im[i + j + half] * wi;
                double vi = re[i + j + half] * wi + im[i + j + half] * wr;
                re[i + j] = ur + vr; im[i + j] = ui + vi;
                re[i + j + half] = ur - vr; im[i + j + half] = ui - vi;
            }
        }
    }
}

ORC_API size_t orc_num_frames(size_t n, uint32_t win, uint32_t hop)
{
    if (n < win) return 0;
    return 1 + (n - win) / hop;
}

/* y[n] 16 kHz mono -> out[T * n_mels] (row-major [frame][mel]); optional power[T * (n_fft/2+1)].
 * Returns T. */
ORC_API size_t orc_logmel(const float *y, size_t n, const orc_feat_config *c, float *out, float *power_out)
{
    size_t T = orc_num_frames(n, c->win_length, c->hop_length);
    if (T == 0) return 0;
    uint32_t N = c->n_fft, nb = N / 2 + 1, M = c->n_mels;
    float *w = (float *)malloc(c->win_length * sizeof(float));
    float *fb = (float *)malloc((size_t)nb * M * sizeof(float));
    double *re = (double *)malloc(N * sizeof(double));
    double *im = (double *)malloc(N * sizeof(double));
    double *P = (double *)malloc(nb * sizeof(double));
    /* per-filter nonzero range to keep the oracle fast */
    uint32_t *lo = (uint32_t *)malloc(M * sizeof(uint32_t));
    uint32_t *hi = (uint32_t *)malloc(M * sizeof(uint32_t));
    orc_hann_window(c->win_length, w);
    orc_mel_filterbank(c, fb);
    for (uint32_t m = 0; m < M; ++m) {
        lo[m] = nb; hi[m] = 0;
        for (uint32_t k = 0; k < nb; ++k)
            if (fb[(size_t)k * M + m] != 0.0f) {
                if (k < lo[m]) lo[m] = k;
                if (k + 1 > hi[m]) hi[m] = k + 1;
            }
    }
    for (size_t f = 0; f < T; ++f) {
        const float *x = y + f * c->hop_length;
        for (uint32_t i = 0; i < N; ++i) {
            float xw = i < c->win_length ? x[i] * w[i] : 0.0f;
            re[i] = (double)xw;
            im[i] = 0.0;
        }
        fft_f64(re, im, N);
        for (uint32_t k = 0; k < nb; ++k) {
            P[k] = re[k] * re[k] + im[k] * im[k];
            if (power_out) power_out[f * nb + k] = (float)P[k];
        }
        for (uint32_t m = 0; m < M; ++m) {
            double acc = 0.0;
            for (uint32_t k = lo[m]; k < hi[m]; ++k)
                acc += (double)fb[(size_t)k * M + m] * P[k];
            double fl = (double)c->log_floor;
            if (acc < fl) acc = fl;
            out[f * M + m] = (float)(c->log10_flag ? log10(acc) : log(acc));
        }
    }
    free(w); free(fb); free(re); free(im); free(P); free(lo); free(hi);
    return T;
}

/* ------------------------------------------------------------------------- */
/* f32 variant of the feature path, used ONLY as the timed CPU baseline        */
/* (a fair "what a CPU implementation costs" number: f32 radix-2 FFT with      */
/* precomputed twiddles and the sparse mel).  Not a parity oracle.             */
/* ------------------------------------------------------------------------- */
typedef struct {
    orc_feat_config c;
    float *w, *fbw;            /* window; compact filter weights */
    uint32_t *lo, *hi, *off;   /* per-filter [lo,hi) and offset into fbw */
    float *twr, *twi;          /* n_fft/2 twiddles */
    uint32_t *rev;
} orc_feat_plan;

ORC_API orc_feat_plan *orc_feat_plan_new(const orc_feat_config *c)
{
    orc_feat_plan *p = (orc_feat_plan *)calloc(1, sizeof(*p));
    p->c = *c;
    uint32_t N = c->n_fft, nb = N / 2 + 1, M = c->n_mels;
    p->w = (float *)malloc(c->win_length * sizeof(float));
    orc_hann_window(c->win_length, p->w);
    float *fb = (float *)malloc((size_t)nb * M * sizeof(float));
    orc_mel_filterbank(c, fb);
    p->lo = (uint32_t *)malloc(M * 4); p->hi = (uint32_t *)malloc(M * 4); p->off = (uint32_t *)malloc(M * 4);
    p->fbw = (float *)malloc((size_t)nb * 2 * sizeof(float) + M * 8);
    uint32_t o = 0;
    for (uint32_t m = 0; m < M; ++m) {
        uint32_t l = nb, h = 0;
        for (uint32_t k = 0; k < nb; ++k)
            if (fb[(size_t)k * M + m] != 0.0f) { if (k < l) l = k; if (k + 1 > h) h = k + 1; }
        if (h < l) { l = 0; h = 0; }
        p->lo[m] = l; p->hi[m] = h; p->off[m] = o;
        for (uint32_t k = l; k < h; ++k) p->fbw[o++] = fb[(size_t)k * M + m];
    }
    free(fb);
    p->twr = (float *)malloc(N / 2 * 4); p->twi = (float *)malloc(N / 2 * 4);
    for (uint32_t j = 0; j < N / 2; ++j) {
        p->twr[j] = (float)cos(-2.0 * M_PI * j / N);
        p->twi[j] = (float)sin(-2.0 * M_PI * j / N);
    }
    p->rev = (uint32_t *)malloc(N * 4);
    uint32_t bits = 0; while ((1u << bits) < N) bits++;
    for (uint32_t i = 0; i < N; ++i) {
        uint32_t r = 0;
        for (uint32_t b = 0; b < bits; ++b) if (i & (1u << b)) r |= 1u << (bits - 1 - b);
        p->rev[i] = r;
    }
    return p;
}
ORC_API void orc_feat_plan_free(orc_feat_plan *p)
{
    if (!p) return;
    free(p->w); free(p->fbw); free(p->lo); free(p->hi); free(p->off); free(p->twr); free(p->twi); free(p->rev);
    free(p);
}

ORC_API size_t orc_logmel_f32(const orc_feat_plan *p, const float *y, size_t n, float *out)
{
    const orc_feat_config *c = &p->c;
    size_t T = orc_num_frames(n, c->win_length, c->hop_length);
    uint32_t N = c->n_fft, nb = N / 2 + 1, M = c->n_mels;
    float re[4096], im[4096], P[2049];
    if (N > 4096) return 0;
    for (size_t f = 0; f < T; ++f) {
        const float *x = y + f * c->hop_length;
        for (uint32_t i = 0; i < N; ++i) {
            uint32_t r = p->rev[i];
            re[r] = i < c->win_length ? x[i] * p->w[i] : 0.0f;
            im[r] = 0.0f;
        }
        for (uint32_t len = 2; len <= N; len <<= 1) {
            uint32_t half = len >> 1, step = N / len;
            for (uint32_t i = 0; i < N; i += len)
                for (uint32_t j = 0; j < half; ++j) {
                    float wr = p->twr[j * step], wi = p->twi[j * step];
                    float vr = re[i + j + half] * wr - im[i + j + half] * wi;
                    float vi = re[i + j + half] * wi + im[i + j + half] * wr;
                    float ur = re[i + j], ui = im[i + j];
                    re[i + j] = ur + vr; im[i + j] = ui + vi;
                    re[i + j + half] = ur - vr; im[i + j + half] = ui - vi;
                }
        }
        for (uint32_t k = 0; k < nb; ++k) P[k] = re[k] * re[k] + im[k] * im[k];
        for (uint32_t m = 0; m < M; ++m) {
            float acc = 0.0f;
            const float *wq = p->fbw + p->off[m];
            for (uint32_t k = p->lo[m]; k < p->hi[m]; ++k) acc += wq[k - p->lo[m]] * P[k];
            if (acc < c->log_floor) acc = c->log_floor;
            out[f * M + m] = c->log10_flag ? log10f(acc) : logf(acc);
        }
    }
    return T;
}

/* ------------------------------------------------------------------------- */
/* PCM16 wire encode: websocket.rs:246-251, :340-347                          */
/*   (x.clamp(-1,1) * 32767.0) as i16  -- Rust `as` truncates toward zero,    */
/*   saturates, NaN -> 0 (clamp keeps NaN as NaN).                            */
/* ------------------------------------------------------------------------- */
ORC_API void orc_pcm16_encode(const float *x, size_t n, int16_t *out)
{
    for (size_t i = 0; i < n; ++i) {
        float v = x[i];
        if (v != v) { out[i] = 0; continue; }
        if (v < -1.0f) v = -1.0f;
        if (v > 1.0f) v = 1.0f;
        float s = v * 32767.0f;
        out[i] = (int16_t)s;   /* |s| <= 32767: C truncation == Rust `as` */
    }
}

/* ------------------------------------------------------------------------- */
/* VAD segmentation: maximal runs from the first Speech frame */
/* to the frame that reports Ending (inclusive) or falls back to Silence       */
/* (exclusive).  seg[2*i], seg[2*i+1] = [start_frame, end_frame).              */
/* ------------------------------------------------------------------------- */
ORC_API size_t orc_vad_segments(const uint8_t *states, size_t T, uint32_t *seg, size_t cap)
{
    size_t ns = 0;
    int in_seg = 0;
    size_t start = 0;
    for (size_t f = 0; f < T; ++f) {
        if (!in_seg) {
            if (states[f] == ORC_SPEECH) { in_seg = 1; start = f; }
        } else {
            if (states[f] == ORC_ENDING) {
                if (ns < cap) { seg[2 * ns] = (uint32_t)start; seg[2 * ns + 1] = (uint32_t)(f + 1); }
                ns++; in_seg = 0;
            } else if (states[f] == ORC_SILENCE) {
                if (ns < cap) { seg[2 * ns] = (uint32_t)start; seg[2 * ns + 1] = (uint32_t)f; }
                ns++; in_seg = 0;
            }
        }
    }
    if (in_seg) {
        if (ns < cap) { seg[2 * ns] = (uint32_t)start; seg[2 * ns + 1] = (uint32_t)T; }
        ns++;
    }
    return ns;
}

/* ------------------------------------------------------------------------- */
/* VAD-gated output (intent specs/0001-spec.md:466): keep only  */
/* the audio and the feature rows of the speech segments, packed in order.      */
/* Frame f of a segment contributes samples [f*hop, (f+1)*hop) and log-mel row f */
/* (samples beyond n read as 0).  off[k] = compacted frame where segment k       */
/* starts, off[n_seg] = kept frames (returned).                                  */
/* ------------------------------------------------------------------------- */
ORC_API size_t orc_vad_gate(const float *pcm, size_t n, const float *logmel, uint32_t n_mels, uint32_t hop,
                            const uint32_t *seg, size_t n_seg, float *out_pcm, float *out_logmel, uint32_t *off)
{
    size_t kept = 0;
    for (size_t k = 0; k < n_seg; ++k) {
        if (off) off[k] = (uint32_t)kept;
        for (uint32_t f = seg[2 * k]; f < seg[2 * k + 1]; ++f, ++kept) {
            if (pcm && out_pcm)
                for (uint32_t i = 0; i < hop; ++i) {
                    const size_t src = (size_t)f * hop + i;
                    out_pcm[kept * hop + i] = src < n ? pcm[src] : 0.0f;
                }
            if (logmel && out_logmel)
                memcpy(out_logmel + kept * n_mels, logmel + (size_t)f * n_mels, sizeof(float) * n_mels);
        }
    }
    if (off) off[n_seg] = (uint32_t)kept;
    return kept;
}

/* ------------------------------------------------------------------------- */
/* Whole-path CPU baseline for one stream (what bench.py times):               */
/* downmix -> BatchResampler(all)+flush -> [logmel f32] -> framed VAD.         */
/* scratch buffers are caller-provided so the timing excludes malloc.          */
/* ------------------------------------------------------------------------- */
ORC_API size_t orc_pipeline_stream(const float *in, size_t n_samples, unsigned channels, uint32_t in_rate,
                                   const orc_feat_plan *plan, const orc_vad_config *vc, uint32_t vad_len,
                                   uint32_t vad_hop, float *mono_scratch, float *pcm_out, size_t pcm_cap,
                                   float *logmel_out, uint8_t *vad_out, size_t *n_frames_out)
{
    size_t nm = orc_to_mono(in, n_samples, channels, mono_scratch);
    size_t ny = orc_resample_stream(in_rate, 16000, mono_scratch, nm, pcm_out, pcm_cap, NULL);
    size_t T = 0;
    if (plan && logmel_out) T = orc_logmel_f32(plan, pcm_out, ny, logmel_out);
    if (vc && vad_out) {
        orc_vad v;
        memset(&v, 0, sizeof(v));
        v.cfg = *vc;
        T = orc_vad_stream(&v, pcm_out, ny, vad_len, vad_hop, vad_out, NULL);
    }
    if (n_frames_out) *n_frames_out = T;
    return ny;
}

/* ------------------------------------------------------------------------- */
/* The same stages driven the way the reference's processing loop would drive   */
/* them (specs/0002-design.md:1113-1118: every ~100 ms read what the ring holds, */
/* to_mono, BatchResampler::process, detect per 20 ms frame), INCLUDING the      */
/* allocations and copies the Rust code performs: `samples.to_vec()` per read    */
/* (capture.rs:134-141), a fresh mono Vec (capture.rs:34-40), `buffer.extend`,   */
/* per 128-frame chunk `buffer[..128].to_vec()`, rubato's Vec<Vec<f32>> result,   */
/* `output.extend` and `buffer.drain(..128)` -- a memmove of the whole residual   */
/* (resampler.rs:132-147).  Timing row "reference-shaped" of bench.py; the        */
/* numbers it produces are identical to orc_pipeline_stream's.                    */
/* ------------------------------------------------------------------------- */
ORC_API size_t orc_pipeline_stream_shaped(const float *in, size_t n_samples, unsigned channels, uint32_t in_rate,
                                          size_t read_frames, const orc_vad_config *vc, uint32_t vad_len,
                                          float *pcm_out, size_t pcm_cap, uint8_t *vad_out, size_t *n_frames_out)
{
    orc_resampler *rs = orc_resampler_new(in_rate, 16000);
    orc_vad v;
    memset(&v, 0, sizeof(v));
    if (vc) v.cfg = *vc;
    float *buffer = NULL; size_t blen = 0, bcap = 0;          /* BatchResampler::buffer */
    float *pending = NULL; size_t plen = 0, pcap = 0;         /* 16 kHz samples waiting for a full VAD frame */
    size_t ny = 0, T = 0;
    const size_t per_read = read_frames * channels;
    for (size_t pos = 0; pos < n_samples; pos += per_read) {
        const size_t n = n_samples - pos < per_read ? n_samples - pos : per_read;
        float *frame = (float *)malloc((n ? n : 1) * sizeof(float));          /* RingBuffer::read -> Vec */
        memcpy(frame, in + pos, n * sizeof(float));
        float *mono = (float *)malloc((n / channels + 1) * sizeof(float));    /* to_mono -> new Vec */
        const size_t nm = orc_to_mono(frame, n, channels, mono);
        if (blen + nm > bcap) { bcap = (blen + nm) * 2 + CHUNK; buffer = (float *)realloc(buffer, bcap * sizeof(float)); }
        memcpy(buffer + blen, mono, nm * sizeof(float));                      /* buffer.extend_from_slice */
        blen += nm;
        float *output = NULL; size_t olen = 0, ocap = 0;                      /* let mut output = Vec::new() */
        while (!rs->passthrough && blen >= CHUNK) {
            float *chunk = (float *)malloc(CHUNK * sizeof(float));            /* buffer[..128].to_vec() */
            memcpy(chunk, buffer, CHUNK * sizeof(float));
            float *res = (float *)malloc(max_out_per_chunk(rs) * sizeof(float));   /* rubato's Vec<Vec<f32>> */
            const size_t m = fastfixedin_step(rs, chunk, res, NULL);
            if (olen + m > ocap) { ocap = (olen + m) * 2 + 64; output = (float *)realloc(output, ocap * sizeof(float)); }
            memcpy(output + olen, res, m * sizeof(float));                    /* output.extend(processed) */
            olen += m;
            free(res); free(chunk);
            memmove(buffer, buffer + CHUNK, (blen - CHUNK) * sizeof(float));  /* buffer.drain(..128) */
            blen -= CHUNK;
        }
        if (rs->passthrough) { output = (float *)malloc((blen ? blen : 1) * sizeof(float)); memcpy(output, buffer, blen * sizeof(float)); olen = blen; blen = 0; }
        if (ny + olen <= pcm_cap) memcpy(pcm_out + ny, output, olen * sizeof(float));
        ny += olen;
        if (vc && vad_len) {
            if (plen + olen > pcap) { pcap = (plen + olen) * 2 + vad_len; pending = (float *)realloc(pending, pcap * sizeof(float)); }
            memcpy(pending + plen, output, olen * sizeof(float));
            plen += olen;
            size_t used = 0;
            while (plen - used >= vad_len) {
                const int st = orc_vad_detect(&v, pending + used, vad_len);
                if (vad_out) vad_out[T] = (uint8_t)st;
                ++T;
                used += vad_len;
            }
            memmove(pending, pending + used, (plen - used) * sizeof(float));
            plen -= used;
        }
        free(output); free(mono); free(frame);
    }
    free(buffer); free(pending);
    orc_resampler_free(rs);
    if (n_frames_out) *n_frames_out = T;
    return ny;
}
