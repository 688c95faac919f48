// af_kernels.cu -- the small kernels around the fused path: compat objects (to_mono, chunked
// resampling, single-detector VAD), generic frame energies, the sequential VAD scan, PCM16
// encode and VAD segmentation.
#include "af_device.cuh"
#include "af_launch.h"

namespace af {

// ---- AudioFrame::to_mono (capture.rs:30-42) on a device buffer ----
__global__ void af_to_mono_kernel(const float *__restrict__ in, uint64_t n_samples, uint32_t channels,
                                  float *__restrict__ out, uint64_t n_frames)
{
    for (uint64_t f = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; f < n_frames;
         f += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t base = f * channels;
        uint32_t m = channels;
        if (base + m > n_samples) m = (uint32_t)(n_samples - base);
        float sum = 0.0f;
        for (uint32_t c = 0; c < m; ++c) sum = __fadd_rn(sum, in[base + c]);
        out[f] = channels == 1 ? in[f] : __fdiv_rn(sum, (float)channels);
    }
}

cudaError_t launch_to_mono(const float *in, uint64_t n_samples, uint32_t channels, float *out, uint64_t n_frames,
                           cudaStream_t st)
{
    if (n_frames == 0) return cudaSuccess;
    const int blocks = (int)std::min<uint64_t>((n_frames + 255) / 256, 148 * 8);
    af_to_mono_kernel<<<blocks, 256, 0, st>>>(in, n_samples, channels, out, n_frames);
    return cudaGetLastError();
}

// ---- generic chunk-exact resampling of mono device signals (AudioResampler / BatchResampler /
//      sessions): one thread per output sample, blockIdx.y = job ----
__global__ void af_resample_jobs_kernel(const ResampleJob *__restrict__ jobs)
{
    const ResampleJob J = jobs[blockIdx.y];
    const unsigned long long count = J.n_end - J.n_begin;
    for (unsigned long long o = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; o < count;
         o += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long n = J.n_begin + o;
        if (J.mode == RS_PASSTHROUGH) {
            const long long idx = (long long)n;
            J.out[o] = (idx < J.data_base || idx >= J.n_valid_end) ? 0.0f : J.data[idx - J.data_base];
            continue;
        }
        long long k; uint32_t rem;
        resample_pos(n, J.p, J.q, &k, &rem);
        float frac;
        if (J.mode == RS_TABLE) {
            frac = J.frac[n];
            if (rem == 0 && frac >= 0.5f) k -= 1;
        } else {
            frac = (float)rem * (1.0f / (float)J.q);
        }
        float y[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const long long idx = k - 1 + t;
            y[t] = (idx < J.data_base || idx >= J.n_valid_end) ? 0.0f : J.data[idx - J.data_base];
        }
        J.out[o] = interp_cubic(frac, y[0], y[1], y[2], y[3]);
    }
}

cudaError_t launch_resample_jobs(const ResampleJob *jobs_dev, uint32_t n_jobs, uint32_t max_outputs, cudaStream_t st)
{
    if (n_jobs == 0 || max_outputs == 0) return cudaSuccess;
    dim3 grid(std::min<uint32_t>((max_outputs + 255) / 256, 1024), n_jobs);
    af_resample_jobs_kernel<<<grid, 256, 0, st>>>(jobs_dev);
    return cudaGetLastError();
}

// ---- calculate_energy (vad.rs:157-168) for arbitrary (frame_len, hop): one thread per frame ----
__global__ void af_frame_energy_kernel(const EnergyJob J)
{
    const uint32_t s = blockIdx.y;
    const uint32_t T = J.n_frames ? J.n_frames[s] : J.n_frames_all;
    const float *y = J.y + (uint64_t)s * J.y_stride;
    float *e = J.energy + (uint64_t)s * J.energy_stride;
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < T; f += gridDim.x * blockDim.x) {
        const float *x = y + (uint64_t)f * J.hop;
        float sum = 0.0f;
        for (uint32_t i = 0; i < J.frame_len; ++i) {
            const float v = x[i];
            sum = __fadd_rn(sum, __fmul_rn(v, v));
        }
        e[f] = J.frame_len ? __fdiv_rn(sum, (float)J.frame_len) : 0.0f;
    }
}

cudaError_t launch_frame_energy(const EnergyJob &job, cudaStream_t st)
{
    if (job.n_streams == 0) return cudaSuccess;
    uint32_t maxT = job.n_frames_all ? job.n_frames_all : 1;
    dim3 grid(std::min<uint32_t>((maxT + 127) / 128, 2048), job.n_streams);
    af_frame_energy_kernel<<<grid, 128, 0, st>>>(job);
    return cudaGetLastError();
}

// ---- detect() EMA + threshold + state machine (vad.rs:101-153): strictly sequential per stream ----
// One thread per stream; latency bound.  Per 32 frames: (1) the EMA chain (fmul, fadd per frame, the only
// floating-point recurrence) produces a word of is_speech bits, (2) the state machine consumes the word
// RUN BY RUN (ffs on the bit word) instead of frame by frame: a run of equal decisions changes the state at
// most once (Silence->Speech on the first 1, Speech->Ending/Silence on the zero that reaches the timeout).
struct VadMachine {        // 32-bit working copy of (state, silence_frames, speech_frames); see the overflow guard below
    uint32_t st, sil, spk;
};

__device__ __forceinline__ void vad_emit(uint8_t *out, uint32_t f, uint32_t count, uint32_t v)
{
    if (!out) return;
    for (uint32_t i = 0; i < count; ++i) out[f + i] = (uint8_t)v;     // independent byte stores
}

// consumes n (<= 32) decisions, bit j of `bits` = frame f0 + j; exact vad.rs:121-153 semantics
__device__ __forceinline__ void vad_machine_word(VadMachine &m, uint32_t bits, uint32_t n, uint32_t f0, uint8_t *out,
                                                 uint32_t timeout, uint32_t min_speech)
{
    uint32_t j = 0;
    while (j < n) {
        const uint32_t rem = bits >> j;                   // bit 0 = decision of frame j (bits above n are zero)
        if (m.st == 0u) {                                 // Silence: zeros keep it; the first 1 enters Speech
            if (rem == 0u) { vad_emit(out, f0 + j, n - j, 0u); j = n; break; }
            const uint32_t z = (uint32_t)__ffs((int)rem) - 1u;
            vad_emit(out, f0 + j, z, 0u);
            m.st = 1u; m.spk = 1u; m.sil = 0u;
            vad_emit(out, f0 + j + z, 1u, 1u);
            j += z + 1u;
        } else if (m.st == 1u) {
            if (rem & 1u) {                               // run of speech frames
                uint32_t r = (uint32_t)__ffs((int)~rem) - 1u;        // ~rem != 0 or ffs gives 0 -> r = 0xffffffff
                if (~rem == 0u) r = 32u;
                r = min(r, n - j);
                m.spk += r; m.sil = 0u;
                vad_emit(out, f0 + j, r, 1u);
                j += r;
            } else {                                      // run of non-speech frames
                uint32_t r = rem ? (uint32_t)__ffs((int)rem) - 1u : 32u;
                r = min(r, n - j);
                const uint32_t need = timeout > m.sil ? timeout - m.sil : 1u;   // zeros until silence_frames >= timeout
                if (r < need) {
                    m.sil += r;
                    vad_emit(out, f0 + j, r, 1u);
                    j += r;
                } else {
                    vad_emit(out, f0 + j, need - 1u, 1u);
                    m.sil += need;
                    m.st = m.spk >= min_speech ? 2u : 0u;
                    m.spk = 0u;
                    vad_emit(out, f0 + j + need - 1u, 1u, m.st);
                    j += need;
                }
            }
        } else {                                          // Ending: one frame, whatever it is
            m.st = 0u; m.sil = 0u;
            vad_emit(out, f0 + j, 1u, 0u);
            j += 1u;
        }
    }
}

__global__ void af_vad_scan_kernel(const ScanJob J)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= J.n_streams) return;
    const uint32_t T = J.n_frames ? J.n_frames[s] : J.n_frames_all;
    const float *e = J.energy + (uint64_t)s * J.energy_stride;
    uint8_t *out = J.states ? J.states + (uint64_t)s * J.states_stride : nullptr;
    VadState v;
    if (J.state_io) v = J.state_io[s];
    else { v.smoothed = 0.0f; v.state = 0; v.silence_frames = 0; v.speech_frames = 0; }
    const VadParams prm = J.prm;
    // the fast path keeps the counters in 32 bits; fall back to the plain 64-bit step when they could overflow
    const bool small = v.silence_frames < 0x40000000ull && v.speech_frames < 0x40000000ull && T < 0x40000000u &&
                       prm.silence_timeout < 0x40000000ull && prm.min_speech < 0x80000000ull;
    if (!small) {
        for (uint32_t f = 0; f < T; ++f) {
            const int stv = vad_step(v, prm, e[f]);
            if (out) out[f] = (uint8_t)stv;
        }
        if (J.state_io) J.state_io[s] = v;
        if (J.final_out) J.final_out[s] = v;
        return;
    }
    const float alpha = prm.alpha, beta = __fsub_rn(1.0f, prm.alpha), e_min = prm.e_min;
    const bool use_smoothed = alpha > 0.0f;
    float sm = v.smoothed;
    VadMachine m{(uint32_t)v.state, (uint32_t)v.silence_frames, (uint32_t)v.speech_frames};
    const uint32_t timeout = (uint32_t)prm.silence_timeout, minsp = (uint32_t)prm.min_speech;
    uint32_t f = 0;
    const bool vec_in = ((reinterpret_cast<uintptr_t>(e) & 15) == 0);
    if (vec_in && T >= 32) {
        const float4 *e4 = reinterpret_cast<const float4 *>(e);
        const uint32_t n_words = T / 32;
        float4 cur[8], nxt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = e4[j];
        for (uint32_t w = 0; w < n_words; ++w) {
            if (w + 1 < n_words) {
#pragma unroll
                for (int j = 0; j < 8; ++j) nxt[j] = e4[(w + 1) * 8 + j];       // prefetch the next 32 energies
            }
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {                 // EMA chain (vad.rs:101-118)
                const float ev[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    sm = __fadd_rn(__fmul_rn(alpha, ev[c]), __fmul_rn(beta, sm));
                    bits |= ((use_smoothed ? sm : ev[c]) >= e_min ? 1u : 0u) << (4 * j + c);
                }
            }
            vad_machine_word(m, bits, 32u, w * 32u, out, timeout, minsp);
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
        }
        f = n_words * 32;
    }
    while (f < T) {                                       // tail (and unaligned rows): up to 32 frames at a time
        const uint32_t n = min(32u, T - f);
        uint32_t bits = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const float ev = e[f + j];
            sm = __fadd_rn(__fmul_rn(alpha, ev), __fmul_rn(beta, sm));
            bits |= ((use_smoothed ? sm : ev) >= e_min ? 1u : 0u) << j;
        }
        vad_machine_word(m, bits, n, f, out, timeout, minsp);
        f += n;
    }
    v.smoothed = sm; v.state = (int)m.st; v.silence_frames = m.sil; v.speech_frames = m.spk;
    if (J.state_io) J.state_io[s] = v;
    if (J.final_out) J.final_out[s] = v;
}

cudaError_t launch_vad_scan(const ScanJob &job, cudaStream_t st)
{
    if (job.n_streams == 0) return cudaSuccess;
    const int threads = 32;                       // one warp per CTA: spread the chains over the SMs
    af_vad_scan_kernel<<<(job.n_streams + threads - 1) / threads, threads, 0, st>>>(job);
    return cudaGetLastError();
}

// ---- streaming sessions (all streams advance in lockstep) ----
// ingest: carry the retained input frames to the front of the new buffer and append the downmixed new frames
__global__ void af_session_ingest_kernel(const SessionIngest J)
{
    const uint32_t s = blockIdx.y;
    const float *old = J.old_buf + (uint64_t)s * J.buf_stride;
    float *neu = J.new_buf + (uint64_t)s * J.buf_stride;
    const char *in = reinterpret_cast<const char *>(J.input) + (uint64_t)s * J.in_stride_bytes;
    const uint32_t total = J.keep + J.n_new_frames;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float v;
        if (i < J.keep) v = old[J.drop + i];
        else v = load_mono(in, J.n_samples, J.n_new_frames, J.channels, J.format, (int)(i - J.keep));
        neu[i] = v;
    }
}

cudaError_t launch_session_ingest(const SessionIngest &J, uint32_t n_streams, cudaStream_t st)
{
    const uint32_t total = J.keep + J.n_new_frames;
    if (n_streams == 0 || total == 0) return cudaSuccess;
    dim3 grid((total + 255) / 256, n_streams);
    af_session_ingest_kernel<<<grid, 256, 0, st>>>(J);
    return cudaGetLastError();
}

// resample the complete chunks of this tick for every stream; carries the unconsumed 16 kHz tail forward
__global__ void af_session_resample_kernel(const SessionResample J)
{
    const uint32_t s = blockIdx.y;
    const float *in = J.in_buf + (uint64_t)s * J.in_stride;
    const float *yold = J.y_old + (uint64_t)s * J.y_stride;
    float *ynew = J.y_new + (uint64_t)s * J.y_stride;
    const uint32_t n_new = (uint32_t)(J.n_end - J.n_begin);
    const uint32_t total = J.y_keep + n_new;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float v;
        if (i < J.y_keep) {
            v = yold[J.y_drop + i];
        } else {
            const unsigned long long n = J.n_begin + (i - J.y_keep);
            if (J.mode == RS_PASSTHROUGH) {
                const long long idx = (long long)n;
                v = (idx < J.data_base || idx >= J.n_valid_end) ? 0.0f : in[idx - J.data_base];
            } else {
                long long k; uint32_t rem;
                resample_pos(n, J.p, J.q, &k, &rem);
                float frac;
                if (J.mode == RS_TABLE) {
                    frac = J.frac[i - J.y_keep];
                    k += __float2int_rn((float)rem * (1.0f / (float)J.q) - frac);
                } else {
                    frac = (float)rem * (1.0f / (float)J.q);
                }
                float y[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const long long idx = k - 1 + t;
                    y[t] = (idx < J.data_base || idx >= J.n_valid_end) ? 0.0f : in[idx - J.data_base];
                }
                v = interp_cubic(frac, y[0], y[1], y[2], y[3]);
            }
        }
        ynew[i] = v;
    }
}

cudaError_t launch_session_resample(const SessionResample &J, uint32_t n_streams, cudaStream_t st)
{
    const uint32_t total = J.y_keep + (uint32_t)(J.n_end - J.n_begin);
    if (n_streams == 0 || total == 0) return cudaSuccess;
    dim3 grid((total + 255) / 256, n_streams);
    af_session_resample_kernel<<<grid, 256, 0, st>>>(J);
    return cudaGetLastError();
}

// every stream of a session has the same lengths: refresh the fused kernel's stream table in place
__global__ void af_session_setup_kernel(StreamDev *tab, uint32_t n_streams, uint32_t n, uint32_t n_frames, uint32_t n_vad)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    tab[s].n_samples = n; tab[s].n_in = n; tab[s].n_out = n; tab[s].n_frames = n_frames; tab[s].n_vad_frames = n_vad;
}

cudaError_t launch_session_setup(StreamDev *tab, uint32_t n_streams, uint32_t n, uint32_t n_frames, uint32_t n_vad,
                                 cudaStream_t st)
{
    if (n_streams == 0) return cudaSuccess;
    af_session_setup_kernel<<<(n_streams + 127) / 128, 128, 0, st>>>(tab, n_streams, n, n_frames, n_vad);
    return cudaGetLastError();
}

// ---- f32 -> PCM16 little endian (websocket.rs:246-251): (x.clamp(-1,1) * 32767.0) as i16 ----
__global__ void af_pcm16_kernel(const float *__restrict__ in, uint64_t n, int16_t *__restrict__ out)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        float v = in[i];
        int r = 0;
        if (v == v) {                             // NaN stays NaN through clamp and casts to 0
            v = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
            r = __float2int_rz(__fmul_rn(v, 32767.0f));
        }
        out[i] = (int16_t)r;
    }
}

cudaError_t launch_pcm16(const float *in, uint64_t n, int16_t *out, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const int blocks = (int)std::min<uint64_t>((n + 255) / 256, 148 * 8);
    af_pcm16_kernel<<<blocks, 256, 0, st>>>(in, n, out);
    return cudaGetLastError();
}

// ---- VAD segmentation: runs from the first Speech frame to Ending (inclusive) / Silence (exclusive) ----
// One CTA per stream.  A segment starts at every frame with state Speech whose predecessor is not Speech, and
// ends at the first later frame that is not Speech; starts and ends alternate, so the k-th start pairs with
// the k-th end.  Each thread owns a contiguous span of frames: count, block-wide exclusive scan, write.
__global__ void af_vad_segments_kernel(const uint8_t *__restrict__ states, uint64_t stride,
                                       const uint32_t *__restrict__ n_frames, uint32_t n_streams,
                                       uint32_t *__restrict__ seg, uint32_t seg_cap, uint32_t *__restrict__ n_seg)
{
    __shared__ uint32_t s_starts[1024], s_ends[1024];
    const uint32_t s = blockIdx.x;
    if (s >= n_streams) return;
    const uint8_t *st = states + (uint64_t)s * stride;
    uint32_t *sg = seg + (uint64_t)s * seg_cap * 2;
    const uint32_t T = n_frames[s];
    const uint32_t nt = blockDim.x, t = threadIdx.x;
    const uint32_t span = (T + nt - 1) / nt;
    const uint32_t lo = min(t * span, T), hi = min(lo + span, T);
    uint32_t ns = 0, ne = 0;
    uint8_t prev = lo > 0 ? st[lo - 1] : (uint8_t)0;
    for (uint32_t f = lo; f < hi; ++f) {
        const uint8_t v = st[f];
        ns += (v == 1 && prev != 1);
        ne += (v != 1 && prev == 1);
        prev = v;
    }
    s_starts[t] = ns; s_ends[t] = ne;
    __syncthreads();
    // exclusive scan (Hillis-Steele on both arrays)
    for (uint32_t d = 1; d < nt; d <<= 1) {
        uint32_t a = 0, b = 0;
        if (t >= d) { a = s_starts[t - d]; b = s_ends[t - d]; }
        __syncthreads();
        s_starts[t] += a; s_ends[t] += b;
        __syncthreads();
    }
    uint32_t ks = s_starts[t] - ns, ke = s_ends[t] - ne;      // index of this span's first start / end
    const uint32_t total = s_starts[nt - 1], total_ends = s_ends[nt - 1];
    prev = lo > 0 ? st[lo - 1] : (uint8_t)0;
    for (uint32_t f = lo; f < hi; ++f) {
        const uint8_t v = st[f];
        if (v == 1 && prev != 1) { if (ks < seg_cap) sg[2 * ks] = f; ks++; }
        if (v != 1 && prev == 1) { if (ke < seg_cap) sg[2 * ke + 1] = v == 2 ? f + 1 : f; ke++; }
        prev = v;
    }
    if (t == 0) {
        if (total > total_ends && total - 1 < seg_cap) sg[2 * (total - 1) + 1] = T;    // still speaking at the end
        n_seg[s] = total;
    }
}

cudaError_t launch_vad_segments(const uint8_t *states, uint64_t stride, const uint32_t *n_frames, uint32_t n_streams,
                                uint32_t *seg, uint32_t seg_cap, uint32_t *n_seg, cudaStream_t st)
{
    if (n_streams == 0) return cudaSuccess;
    af_vad_segments_kernel<<<n_streams, 256, 0, st>>>(states, stride, n_frames, n_streams, seg, seg_cap, n_seg);
    return cudaGetLastError();
}

}  // namespace af
