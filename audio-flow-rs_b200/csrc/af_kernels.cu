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
// One thread per stream.  The dependent chain is ~3 FP32 ops + a few integer ops per frame; the kernel is
// latency bound, so energies are fetched 16 frames ahead as 4 x float4 (software pipelined) and the states
// leave as one 16-byte store per 16 frames.
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
    uint32_t f = 0;
    const bool vec_in = ((reinterpret_cast<uintptr_t>(e) & 15) == 0);
    const bool vec_out = out && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec_in && T >= 16) {
        const float4 *e4 = reinterpret_cast<const float4 *>(e);
        float4 nb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) nb[j] = e4[j];
        for (; f + 16 <= T; f += 16) {
            float4 cb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) cb[j] = nb[j];
            if (f + 32 <= T) {
#pragma unroll
                for (int j = 0; j < 4; ++j) nb[j] = e4[(f + 16) / 4 + j];       // prefetch the next 16 frames
            }
            uint32_t packed[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t s0 = (uint32_t)vad_step(v, prm, cb[j].x);
                const uint32_t s1 = (uint32_t)vad_step(v, prm, cb[j].y);
                const uint32_t s2 = (uint32_t)vad_step(v, prm, cb[j].z);
                const uint32_t s3 = (uint32_t)vad_step(v, prm, cb[j].w);
                packed[j] = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
            }
            if (vec_out) {
                *reinterpret_cast<uint4 *>(out + f) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            } else if (out) {
#pragma unroll
                for (int j = 0; j < 16; ++j) out[f + j] = (uint8_t)(packed[j >> 2] >> (8 * (j & 3)));
            }
        }
    }
    for (; f < T; ++f) {
        const int st = vad_step(v, prm, e[f]);
        if (out) out[f] = (uint8_t)st;
    }
    if (J.state_io) J.state_io[s] = v;
    if (J.final_out) J.final_out[s] = v;
}

cudaError_t launch_vad_scan(const ScanJob &job, cudaStream_t st)
{
    if (job.n_streams == 0) return cudaSuccess;
    const int threads = 32;                       // few streams per CTA: spread the chains over the SMs
    af_vad_scan_kernel<<<(job.n_streams + threads - 1) / threads, threads, 0, st>>>(job);
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
__global__ void af_vad_segments_kernel(const uint8_t *__restrict__ states, uint64_t stride,
                                       const uint32_t *__restrict__ n_frames, uint32_t n_streams,
                                       uint32_t *__restrict__ seg, uint32_t seg_cap, uint32_t *__restrict__ n_seg)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const uint8_t *st = states + (uint64_t)s * stride;
    uint32_t *sg = seg + (uint64_t)s * seg_cap * 2;
    const uint32_t T = n_frames[s];
    uint32_t ns = 0, start = 0;
    bool in_seg = false;
    for (uint32_t f = 0; f < T; ++f) {
        const uint8_t v = st[f];
        if (!in_seg) {
            if (v == 1) { in_seg = true; start = f; }
        } else if (v != 1) {
            if (ns < seg_cap) { sg[2 * ns] = start; sg[2 * ns + 1] = v == 2 ? f + 1 : f; }
            ns++; in_seg = false;
        }
    }
    if (in_seg) {
        if (ns < seg_cap) { sg[2 * ns] = start; sg[2 * ns + 1] = T; }
        ns++;
    }
    n_seg[s] = ns;
}

cudaError_t launch_vad_segments(const uint8_t *states, uint64_t stride, const uint32_t *n_frames, uint32_t n_streams,
                                uint32_t *seg, uint32_t seg_cap, uint32_t *n_seg, cudaStream_t st)
{
    if (n_streams == 0) return cudaSuccess;
    af_vad_segments_kernel<<<(n_streams + 31) / 32, 32, 0, st>>>(states, stride, n_frames, n_streams, seg, seg_cap, n_seg);
    return cudaGetLastError();
}

}  // namespace af
