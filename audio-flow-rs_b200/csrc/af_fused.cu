// af_fused.cu -- the fused hot-path kernel for sm_100a:
//   K1 downmix + cubic (rubato FastFixedIn) resample  -> 16 kHz step buffer in shared memory; the raw
//      input of the NEXT step is staged into shared memory by a TMA bulk copy (cp.async.bulk +
//      mbarrier) while the current step is in its FFT phase
//   K2 Hann window + 512-point real FFT (packed 256-point complex, 16x16 in registers,
//      one half-warp per frame, transposed through shared memory, Hermitian split by shuffles)
//   K3 sparse banded mel projection + log
//   K4 (energy part) bit-exact sequential mean-square per frame on a dedicated VAD warp
// Replaces capture.rs:30-42, resampler.rs:71-93/132-166 (+ rubato) and the O(len) part of
// vad.rs:157-168; the STFT/mel stages are spec-defined.  The sequential EMA/state
// machine of vad.rs:101-153 runs in af_vad_scan_kernel (af_kernels.cu).
#include "af_device.cuh"
#include "af_launch.h"

namespace af {

struct __align__(128) FusedSmem {
    unsigned char stage[STAGE_BYTES];                // raw interleaved input of one step (bulk-copy target)
    float ybuf[YBUF_FLOATS];                         // padded 16 kHz samples of the current step
    float scr[16 * SCR_FLOATS_PER_FRAME];            // per half-warp transpose scratch / log-mel stage
    float pbuf[PBUF_FLOATS];                         // 4*|X[k]|^2, [bin][frame]
    FftTables fft;
    MelTables mel;
    StreamDev stream;                                // descriptor of the tile's stream
    unsigned long long mbar;                         // completion barrier of the in-flight stage fill
    unsigned long long st_lo, st_hi;                 // interleaved element range [lo, hi) held by the stage
    uint32_t st_interior;                            // 1: every tap of the step is inside the stage and the stream
    int tile_k;                                      // floor(position) of the tile's first output
    uint32_t tile_rem;                               // and its remainder (numerator units)
    uint32_t inc_k, inc_rem;                         // position increment for FUSED_THREADS outputs
};

// ---- mbarrier / bulk-copy wrappers (PTX; SASS: SYNCS.*, UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared-memory load the compiler cannot rematerialise at the use site: keeps per-lane constants in registers
__device__ __forceinline__ float2 lds_f2_pinned(const void *p)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- stage fill: issued by ONE thread for the step (tile, g); always completes one mbarrier phase ----
__device__ __forceinline__ void issue_fill(FusedSmem &sm, const FusedParams &P, uint32_t tile, uint32_t g)
{
    const TileDev td = P.tiles[tile];
    const StreamDev *sp = P.streams + td.stream;
    const uint32_t n_out = sp->n_out, n_in = sp->n_in, mode = sp->mode, p = sp->p, q = sp->q;
    const uint32_t ch = sp->channels, bps = sp->format == FMT_I16 ? 2u : 4u;
    const unsigned long long n_samples = sp->n_samples;
    unsigned long long lo = 0, hi = 0;
    uint32_t bytes = 0, interior = 0;
    const char *src = reinterpret_cast<const char *>(sp->data);
    {
        const unsigned long long step0 = (unsigned long long)td.tile * TILE_SAMPLES + (unsigned long long)g * STEP_SAMPLES;
        const unsigned long long n_first = step0 + (g == 0 ? 0 : CARRY);
        unsigned long long n_last = step0 + YLEN;
        if (n_last > n_out) n_last = n_out;
        if (n_first < n_last) {
            long long k0, k1;
            if (mode == RS_PASSTHROUGH) { k0 = (long long)n_first; k1 = (long long)n_last - 1; }
            else {
                uint32_t r;
                resample_pos(n_first, p, q, &k0, &r);
                resample_pos(n_last - 1, p, q, &k1, &r);
            }
            // interior step: no tap (k-2 .. k+2) leaves the stream, no output beyond n_out, whole channel frames
            const bool geom = (k0 - 2 >= 0) && (k1 + 3 <= (long long)n_in) && ((unsigned long long)(k1 + 3) * ch <= n_samples) &&
                              (n_last == step0 + YLEN) && ch <= 2;
            if (geom) interior = 2;                              // unchecked taps straight from global memory
            long long i_lo = k0 - 2, i_hi = k1 + 3;
            if (i_lo < 0) i_lo = 0;
            if (i_hi > (long long)n_in) i_hi = (long long)n_in;
            if (P.use_stage && sp->staged && i_lo < i_hi) {
                unsigned long long b_lo = ((unsigned long long)i_lo * ch * bps) & ~15ull;
                unsigned long long e_hi = (unsigned long long)i_hi * ch;
                if (e_hi > n_samples) e_hi = n_samples;
                unsigned long long b_hi = (e_hi * bps + 15ull) & ~15ull;
                const unsigned long long b_end = (n_samples * bps) & ~15ull;   // never read past the stream's last full 16 bytes
                if (b_hi > b_end) b_hi = b_end;
                if (b_hi > b_lo && b_hi - b_lo <= (unsigned long long)STAGE_BYTES) {
                    bytes = (uint32_t)(b_hi - b_lo);
                    lo = b_lo / bps; hi = b_hi / bps;
                    src += b_lo;
                    if (geom && (unsigned long long)(k1 + 3) * ch <= hi) interior = 1;   // ... and from the stage
                }
            }
        }
    }
    sm.st_lo = lo; sm.st_hi = hi; sm.st_interior = interior;
    if (bytes) {
        mbar_arrive_expect_tx(&sm.mbar, bytes);
        bulk_g2s(sm.stage, src, bytes, &sm.mbar);
    } else {
        mbar_arrive(&sm.mbar);
    }
}

// ---- one mono input frame (downmixed) for the resampler, from the stage when it holds it ----
enum { K_F32_1 = 0, K_I16_1 = 1, K_F32_2 = 2, K_I16_2 = 3, K_GENERIC = 4 };

template <int KIND>
__device__ __forceinline__ float tap(const FusedSmem &sm, const StreamDev &s, uint32_t st_lo, uint32_t st_hi, int idx)
{
    if ((uint32_t)idx >= s.n_in) return 0.0f;                        // also covers idx < 0
    if (KIND == K_F32_1) {
        const uint32_t e = (uint32_t)idx;
        if (e >= st_lo && e < st_hi) return reinterpret_cast<const float *>(sm.stage)[e - st_lo];
        return __ldg(reinterpret_cast<const float *>(s.data) + e);
    } else if (KIND == K_I16_1) {
        const uint32_t e = (uint32_t)idx;
        short v;
        if (e >= st_lo && e < st_hi) v = reinterpret_cast<const short *>(sm.stage)[e - st_lo];
        else v = __ldg(reinterpret_cast<const short *>(s.data) + e);
        return (float)v * (1.0f / 32768.0f);
    } else if (KIND == K_F32_2) {
        const uint32_t e = 2u * (uint32_t)idx;
        if (e >= st_lo && e + 2 <= st_hi) {
            const float2 v = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(sm.stage) + (e - st_lo));
            return __fmul_rn(__fadd_rn(__fadd_rn(0.0f, v.x), v.y), 0.5f);
        }
        return load_mono(s.data, s.n_samples, s.n_in, 2, FMT_F32, idx);
    } else if (KIND == K_I16_2) {
        const uint32_t e = 2u * (uint32_t)idx;
        if (e >= st_lo && e + 2 <= st_hi) {
            const short2 v = *reinterpret_cast<const short2 *>(reinterpret_cast<const short *>(sm.stage) + (e - st_lo));
            const float l = (float)v.x * (1.0f / 32768.0f), r = (float)v.y * (1.0f / 32768.0f);
            return __fmul_rn(__fadd_rn(__fadd_rn(0.0f, l), r), 0.5f);
        }
        return load_mono(s.data, s.n_samples, s.n_in, 2, FMT_I16, idx);
    } else {
        return load_mono(s.data, s.n_samples, s.n_in, s.channels, s.format, idx);
    }
}

// unchecked tap for interior steps: `off` = mono frame index relative to the first staged frame
template <int KIND>
__device__ __forceinline__ float tap_fast(const unsigned char *stage, int off)
{
    if (KIND == K_F32_1) return reinterpret_cast<const float *>(stage)[off];
    if (KIND == K_I16_1) return (float)reinterpret_cast<const short *>(stage)[off] * (1.0f / 32768.0f);
    if (KIND == K_F32_2) {
        const float2 v = reinterpret_cast<const float2 *>(stage)[off];
        return __fmul_rn(__fadd_rn(__fadd_rn(0.0f, v.x), v.y), 0.5f);
    }
    const short2 v = reinterpret_cast<const short2 *>(stage)[off];
    const float l = (float)v.x * (1.0f / 32768.0f), r = (float)v.y * (1.0f / 32768.0f);
    return __fmul_rn(__fadd_rn(__fadd_rn(0.0f, l), r), 0.5f);
}

// ---- phase 1, interior steps: every tap comes unchecked from the stage ----
template <int KIND, bool STAGED>
__device__ __forceinline__ void resample_step_fast(FusedSmem &sm, const StreamDev &s, uint32_t tile_off, uint32_t base,
                                                   int i_begin)
{
    // taps come from the shared-memory stage (first staged mono frame f_lo) or, when the step does not fit the
    // stage (e.g. stereo f32), unchecked from the stream in global memory (f_lo = 0)
    const unsigned char *__restrict__ srcp = STAGED ? sm.stage : reinterpret_cast<const unsigned char *>(s.data);
    const int f_lo = STAGED ? (int)((uint32_t)sm.st_lo / ((KIND == K_F32_2 || KIND == K_I16_2) ? 2u : 1u)) : 0;
    const int tid = threadIdx.x;
    const uint32_t mode = s.mode;
    int i = i_begin + tid;
    if (mode == RS_PASSTHROUGH) {
        for (; i < YLEN; i += FUSED_THREADS) sm.ybuf[ypad(i)] = tap_fast<KIND>(srcp, (int)(base + i) - f_lo);
        return;
    }
    const uint32_t q = s.q;
    if (KIND == K_F32_1 && q == 1 && s.p == 3) {
        // 48 kHz -> 16 kHz mono f32, four outputs per thread: output n reads x[3n-2 .. 3n+1]; for n = 0 mod 4 that
        // is an 8-byte aligned float2 followed by three 16-byte aligned float4 of the stage (14 floats, 13 used)
        const float *stg = reinterpret_cast<const float *>(srcp);
        const int kbase = sm.tile_k + 3 * (int)tile_off - 1 - f_lo;
        // start on a 32-sample boundary of the padded step buffer so that every quarter-warp stores 128 contiguous bytes
        for (int i4 = (i_begin & ~31) + 4 * tid; i4 < YLEN; i4 += 4 * FUSED_THREADS) {
            if (i4 < i_begin) continue;
            const float *px = stg + (kbase + 3 * i4);
            float v[14];
            const float2 h = *reinterpret_cast<const float2 *>(px);
            const float4 a4 = *reinterpret_cast<const float4 *>(px + 2);
            const float4 b4 = *reinterpret_cast<const float4 *>(px + 6);
            const float4 c4 = *reinterpret_cast<const float4 *>(px + 10);
            v[0] = h.x; v[1] = h.y; v[2] = a4.x; v[3] = a4.y; v[4] = a4.z; v[5] = a4.w; v[6] = b4.x; v[7] = b4.y;
            v[8] = b4.z; v[9] = b4.w; v[10] = c4.x; v[11] = c4.y; v[12] = c4.z; v[13] = c4.w;
            float big = 0.0f;
#pragma unroll
            for (int t = 0; t < 13; ++t) big += fabsf(v[t]);                      // NaN / Inf propagate
            float4 y = make_float4(v[1], v[4], v[7], v[10]);
            if (!(big < 1e30f && y.x != 0.0f && y.y != 0.0f && y.z != 0.0f && y.w != 0.0f)) {
                y.x = interp_cubic(0.0f, v[0], v[1], v[2], v[3]);
                y.y = interp_cubic(0.0f, v[3], v[4], v[5], v[6]);
                y.z = interp_cubic(0.0f, v[6], v[7], v[8], v[9]);
                y.w = interp_cubic(0.0f, v[9], v[10], v[11], v[12]);
            }
            *reinterpret_cast<float4 *>(sm.ybuf + ypad(i4)) = y;
        }
        return;
    }
    const uint32_t a = sm.tile_rem + (tile_off + (uint32_t)i) * s.p;
    if (q == 1) {
        // integer step (48 kHz -> 16 kHz): frac == 0 exactly; the cubic returns y1 bit for bit whenever y1 != 0 and
        // the coefficients are finite -- checked per sample, everything else takes the polynomial
        int o = sm.tile_k + (int)a - 1 - f_lo;
        const int inc = (int)sm.inc_k;
        for (; i < YLEN; i += FUSED_THREADS, o += inc) {
            const float y0 = tap_fast<KIND>(srcp, o), y1 = tap_fast<KIND>(srcp, o + 1);
            const float y2 = tap_fast<KIND>(srcp, o + 2), y3 = tap_fast<KIND>(srcp, o + 3);
            const float big = (fabsf(y0) + fabsf(y1)) + (fabsf(y2) + fabsf(y3));   // NaN / Inf propagate
            float v = y1;
            if (!(y1 != 0.0f && big < 1e30f)) v = interp_cubic(0.0f, y0, y1, y2, y3);
            sm.ybuf[ypad(i)] = v;
        }
        return;
    }
    uint32_t dk = a / q, rem = a - dk * q;
    int k = sm.tile_k + (int)dk - 1 - f_lo;
    const uint32_t inc_k = sm.inc_k, inc_rem = sm.inc_rem;
    const float inv_q = 1.0f / (float)q;
    const float *__restrict__ frac_tab = s.frac + base;
    for (; i < YLEN; i += FUSED_THREADS) {
        int o = k;
        float frac;
        if (mode == RS_TABLE) {
            frac = __ldg(frac_tab + i);
            o += __float2int_rn((float)rem * inv_q - frac);           // -1 when the f64 recurrence sits just below an integer
        } else {
            frac = (float)rem * inv_q;                                // q is a power of two: exact
        }
        const float y0 = tap_fast<KIND>(srcp, o), y1 = tap_fast<KIND>(srcp, o + 1);
        const float y2 = tap_fast<KIND>(srcp, o + 2), y3 = tap_fast<KIND>(srcp, o + 3);
        sm.ybuf[ypad(i)] = interp_cubic(frac, y0, y1, y2, y3);
        k += (int)inc_k;
        rem += inc_rem;
        if (rem >= q) { rem -= q; k += 1; }
    }
}

// ---- phase 1: resample the 16 kHz samples [base + i_begin, base + YLEN) of the stream into ybuf ----
template <int KIND>
__device__ __noinline__ void resample_step(FusedSmem &sm, const StreamDev &s, uint32_t tile_off, uint32_t base,
                                              int i_begin)
{
    const int tid = threadIdx.x;
    const uint32_t st_lo = (uint32_t)sm.st_lo, st_hi = (uint32_t)sm.st_hi;
    const uint32_t n_out = s.n_out, mode = s.mode;
    int i = i_begin + tid;
    if (mode == RS_PASSTHROUGH) {
        for (; i < YLEN; i += FUSED_THREADS) {
            const uint32_t n = base + i;
            sm.ybuf[ypad(i)] = n < n_out ? tap<KIND>(sm, s, st_lo, st_hi, (int)n) : 0.0f;
        }
        return;
    }
    // exact integer position of this thread's first output, relative to the tile start
    const uint32_t q = s.q;
    const uint32_t a = sm.tile_rem + (tile_off + (uint32_t)i) * s.p;
    uint32_t dk, rem;
    if (q == 1) { dk = a; rem = 0; }
    else { dk = a / q; rem = a - dk * q; }
    int k = sm.tile_k + (int)dk;
    const uint32_t inc_k = sm.inc_k, inc_rem = sm.inc_rem;
    const float inv_q = 1.0f / (float)q;
    const float *__restrict__ frac_tab = s.frac;
    for (; i < YLEN; i += FUSED_THREADS) {
        const uint32_t n = base + i;
        float v = 0.0f;
        if (n < n_out) {
            int kk = k;
            float frac;
            if (mode == RS_TABLE) {
                frac = __ldg(frac_tab + n);
                kk += __float2int_rn((float)rem * inv_q - frac);      // -1 when the f64 recurrence sits just below an integer
            } else {
                frac = (float)rem * inv_q;                            // q is a power of two: exact
            }
            const float y0 = tap<KIND>(sm, s, st_lo, st_hi, kk - 1);
            const float y1 = tap<KIND>(sm, s, st_lo, st_hi, kk);
            const float y2 = tap<KIND>(sm, s, st_lo, st_hi, kk + 1);
            const float y3 = tap<KIND>(sm, s, st_lo, st_hi, kk + 2);
            // frac == 0 (48 kHz -> 16 kHz): a0 + a1*0 + a2*0 + a3*0 == y1 bit for bit whenever y1 != 0 and the
            // coefficients are finite; only then skip the polynomial
            const float big = (fabsf(y0) + fabsf(y1)) + (fabsf(y2) + fabsf(y3));   // NaN / Inf propagate
            if (frac == 0.0f && y1 != 0.0f && big < 1e30f) v = y1;
            else v = interp_cubic(frac, y0, y1, y2, y3);
        }
        sm.ybuf[ypad(i)] = v;
        k += (int)inc_k;
        rem += inc_rem;
        if (rem >= q) { rem -= q; k += 1; }
    }
}

// ---- phase 2a: one frame per half-warp: window, packed real FFT, power -> pbuf[bin][q] ----
__device__ __forceinline__ void fft_frame(FusedSmem &sm, float *__restrict__ scr, int q, int l, int lane,
                                          const float2 (&tw2r)[8], const float2 (&winr)[13])
{
    float xr[16], xi[16];
    const float *yb = sm.ybuf + 180 * q + 2 * l;      // ypad(160 q + 32 n1 + 2 l) = 180 q + 36 n1 + 2 l
#pragma unroll
    for (int n1 = 0; n1 < 12; ++n1) {
        const float2 v = *reinterpret_cast<const float2 *>(yb + 36 * n1);
        xr[n1] = __fmul_rn(v.x, winr[n1].x);
        xi[n1] = __fmul_rn(v.y, winr[n1].y);
    }
    {
        float2 v = make_float2(0.0f, 0.0f);
        if (l < 8) v = *reinterpret_cast<const float2 *>(yb + 36 * 12);   // samples 384 + 2l (+1) < 400; winr[12] is 0 beyond
        xr[12] = __fmul_rn(v.x, winr[12].x);
        xi[12] = __fmul_rn(v.y, winr[12].y);
    }
#pragma unroll
    for (int n1 = 13; n1 < 16; ++n1) { xr[n1] = 0.0f; xi[n1] = 0.0f; }

    // pass 1: 16-point FFT over n1 (this lane is n2 = l), twiddle W256^(l k1), transposed store
    fft16<true>(xr, xi);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        float ar = xr[fft16_slot(k1)], ai = xi[fft16_slot(k1)];
        if (k1 > 0) {
            const float2 w = sm.fft.tw1[k1 * 16 + l];
            AF_CMUL(ar, ai, w.x, w.y);
        }
        *reinterpret_cast<float2 *>(scr + (k1 * SCR_ROW + l) * 2) = make_float2(ar, ai);
    }
    __syncwarp();
    // pass 2: this lane is k1 = l; read its row (all n2), 16-point FFT over n2
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float4 v = *reinterpret_cast<const float4 *>(scr + (l * SCR_ROW + 2 * u) * 2);
        xr[2 * u] = v.x; xi[2 * u] = v.y; xr[2 * u + 1] = v.z; xi[2 * u + 1] = v.w;
    }
    __syncwarp();
    fft16<false>(xr, xi);
    // Z[l + 16 k2] is in slot(k2).  Hermitian split: pair k = l + 16 r with 256 - k, which lives in
    // lane (16 - l) & 15 at k2 = 15 - r (lane 0 pairs with itself at k2 = (16 - r) & 15).
    const int src = ((16 - l) & 15) | (lane & 16);
    float *pb = sm.pbuf + q;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const float zr = xr[fft16_slot(r)], zi = xi[fft16_slot(r)];
        float pr = __shfl_sync(0xffffffffu, xr[fft16_slot(15 - r)], src);
        float pi = __shfl_sync(0xffffffffu, xi[fft16_slot(15 - r)], src);
        if (l == 0) { pr = xr[fft16_slot((16 - r) & 15)]; pi = xi[fft16_slot((16 - r) & 15)]; }
        const int k = l + 16 * r;
        const float2 w = tw2r[r];
        const float e2r = zr + pr, e2i = zi - pi;      // 2E = Z[k] + conj(Z[256-k])
        const float o2r = zi + pi, o2i = pr - zr;      // 2O = -i (Z[k] - conj(Z[256-k]))
        const float tr = w.x * o2r - w.y * o2i;
        const float ti = w.x * o2i + w.y * o2r;
        const float ar = e2r + tr, ai = e2i + ti;      // 2 X[k]
        const float br = e2r - tr, bi = e2i - ti;      // 2 conj(X[256-k])
        pb[k * PB_ROW] = ar * ar + ai * ai;
        pb[(256 - k) * PB_ROW] = br * br + bi * bi;
    }
    if (l == 0) {                                       // k = 128 pairs with itself
        const float zr = xr[fft16_slot(8)], zi = xi[fft16_slot(8)];
        pb[128 * PB_ROW] = 4.0f * (zr * zr + zi * zi);
    }
}

// ---- phase 2b: VAD warp, lane = frame: calculate_energy (vad.rs:157-168), strictly sequential ----
__device__ __forceinline__ float frame_energy_smem(const float *__restrict__ ybuf, int q)
{
    const float4 *yp = reinterpret_cast<const float4 *>(ybuf) + 45 * q;   // ypad(160 q) / 4
    float sum = 0.0f;
#pragma unroll 4
    for (int s8 = 0; s8 < 12; ++s8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 v = yp[9 * s8 + j];            // 8 float4 of data + 1 of padding per 32 samples
            sum = __fadd_rn(sum, __fmul_rn(v.x, v.x));
            sum = __fadd_rn(sum, __fmul_rn(v.y, v.y));
            sum = __fadd_rn(sum, __fmul_rn(v.z, v.z));
            sum = __fadd_rn(sum, __fmul_rn(v.w, v.w));
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {                       // samples 384..399
        const float4 v = yp[9 * 12 + j];
        sum = __fadd_rn(sum, __fmul_rn(v.x, v.x));
        sum = __fadd_rn(sum, __fmul_rn(v.y, v.y));
        sum = __fadd_rn(sum, __fmul_rn(v.z, v.z));
        sum = __fadd_rn(sum, __fmul_rn(v.w, v.w));
    }
    return __fdiv_rn(sum, (float)WIN);
}

__global__ void __launch_bounds__(FUSED_THREADS, 2) af_fused_kernel(const FusedParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = lane & 15, half = lane >> 4;
    const uint32_t M = P.n_mels;
    const bool is_filler = (tid == VAD_WARP * 32 + 31);             // last lane of the VAD warp (it has no frame)

    // constant tables -> shared memory (once per CTA)
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(P.fft);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sm.fft);
        for (int i = tid; i < (int)(sizeof(FftTables) / 4); i += FUSED_THREADS) dst[i] = src[i];
        if (M) {
            const uint32_t *ms = reinterpret_cast<const uint32_t *>(P.mel);
            uint32_t *md = reinterpret_cast<uint32_t *>(&sm.mel);
            for (int i = tid; i < (int)(sizeof(MelTables) / 4); i += FUSED_THREADS) md[i] = ms[i];
        }
    }
    if (is_filler) {
        mbar_init(&sm.mbar, 1);
        if (blockIdx.x < P.n_tiles) issue_fill(sm, P, blockIdx.x, 0);
    }
    __syncthreads();
    float2 tw2r[8];                                     // W512^(l + 16 r): this lane's post-twiddles, register resident
#pragma unroll
    for (int r = 0; r < 8; ++r) tw2r[r] = lds_f2_pinned(&sm.fft.tw2[l + 16 * r]);
    float2 winr[13];                                    // this lane's window samples w[32 n1 + 2 l], w[32 n1 + 2 l + 1]
#pragma unroll
    for (int n1 = 0; n1 < 13; ++n1) winr[n1] = lds_f2_pinned(sm.fft.window + 32 * n1 + 2 * l);
    const float log_mul = P.log_scale;
    uint32_t fill_parity = 0;

    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const TileDev td = P.tiles[tile];
        if (tid < (int)(sizeof(StreamDev) / 4))
            reinterpret_cast<uint32_t *>(&sm.stream)[tid] = reinterpret_cast<const uint32_t *>(P.streams + td.stream)[tid];
        __syncthreads();
        const StreamDev &s = sm.stream;
        const uint32_t n_tile0 = td.tile * TILE_SAMPLES;
        const uint32_t tile_end = min(n_tile0 + (uint32_t)TILE_SAMPLES, s.n_out);
        const uint32_t f_tile0 = td.tile * TILE_FRAMES;
        const uint32_t n_steps = (tile_end - n_tile0 + STEP_SAMPLES - 1) / STEP_SAMPLES;
        if (tid == 0 && s.mode != RS_PASSTHROUGH) {
            long long k; uint32_t rem;
            resample_pos(n_tile0, s.p, s.q, &k, &rem);
            sm.tile_k = (int)k; sm.tile_rem = rem;
            const uint32_t inc = (uint32_t)FUSED_THREADS * s.p;
            sm.inc_k = inc / s.q; sm.inc_rem = inc % s.q;
        }
        __syncthreads();
        int kind = K_GENERIC;
        if (s.channels == 1) kind = s.format == FMT_F32 ? K_F32_1 : K_I16_1;
        else if (s.channels == 2) kind = s.format == FMT_F32 ? K_F32_2 : K_I16_2;
        float *pcm_row = P.pcm ? P.pcm + (uint64_t)td.stream * P.pcm_stride : nullptr;
        float *lm_row = P.logmel ? P.logmel + (uint64_t)td.stream * P.logmel_stride : nullptr;
        float *en_row = P.energy ? P.energy + (uint64_t)td.stream * P.energy_stride : nullptr;

        for (uint32_t g = 0; g < n_steps; ++g) {
            const uint32_t base = n_tile0 + g * STEP_SAMPLES;       // stream index of ybuf sample 0
            const uint32_t f0 = f_tile0 + g * SF;                   // first frame of the step
            const int n_valid = f0 < s.n_frames ? (int)min((uint32_t)SF, s.n_frames - f0) : 0;

            // ---- phase 1: wait for the staged input, resample (first step of a tile also recomputes the halo) ----
            mbar_wait(&sm.mbar, fill_parity);
            fill_parity ^= 1u;
            const int i_begin = g == 0 ? 0 : CARRY;
            const uint32_t toff = g * STEP_SAMPLES;
            if (sm.st_interior && kind != K_GENERIC) {
                if (sm.st_interior == 1) {
                    switch (kind) {
                    case K_F32_1: resample_step_fast<K_F32_1, true>(sm, s, toff, base, i_begin); break;
                    case K_I16_1: resample_step_fast<K_I16_1, true>(sm, s, toff, base, i_begin); break;
                    case K_F32_2: resample_step_fast<K_F32_2, true>(sm, s, toff, base, i_begin); break;
                    default: resample_step_fast<K_I16_2, true>(sm, s, toff, base, i_begin); break;
                    }
                } else {
                    switch (kind) {
                    case K_F32_1: resample_step_fast<K_F32_1, false>(sm, s, toff, base, i_begin); break;
                    case K_I16_1: resample_step_fast<K_I16_1, false>(sm, s, toff, base, i_begin); break;
                    case K_F32_2: resample_step_fast<K_F32_2, false>(sm, s, toff, base, i_begin); break;
                    default: resample_step_fast<K_I16_2, false>(sm, s, toff, base, i_begin); break;
                    }
                }
            } else {
                switch (kind) {
                case K_F32_1: resample_step<K_F32_1>(sm, s, toff, base, i_begin); break;
                case K_I16_1: resample_step<K_I16_1>(sm, s, toff, base, i_begin); break;
                case K_F32_2: resample_step<K_F32_2>(sm, s, toff, base, i_begin); break;
                case K_I16_2: resample_step<K_I16_2>(sm, s, toff, base, i_begin); break;
                default: resample_step<K_GENERIC>(sm, s, toff, base, i_begin); break;
                }
            }
            __syncthreads();

            // the 400-term energy chains of the VAD warp (>= 1600 cycles) run beside the FFT phase of warps 0..7
            const bool vad_busy = P.do_energy && en_row;
            const int mel_threads = FUSED_THREADS;
            if (warp == VAD_WARP) {
                // ===== warp 8: next stage fill (async bulk copy), then one 400-term energy chain per lane =====
                if (is_filler) {
                    if (g + 1 < n_steps) issue_fill(sm, P, tile, g + 1);
                    else if (tile + gridDim.x < P.n_tiles) issue_fill(sm, P, tile + gridDim.x, 0);
                }
                if (vad_busy && lane < n_valid) en_row[f0 + lane] = frame_energy_smem(sm.ybuf, lane);
                __syncwarp();
                named_bar_arrive(2, FUSED_THREADS);                   // done reading ybuf
            } else if (warp == AUX_WARP) {
                // ===== warp 9: PCM write-out of the step, then carry the 240-sample overlap forward =====
                if (pcm_row) {
                    const uint32_t own_end = min(base + (uint32_t)STEP_SAMPLES, tile_end);
                    if (own_end == base + (uint32_t)STEP_SAMPLES) {
                        // full step: 640 float4, 20 per lane, loads batched ahead of the stores
                        float4 *dst = reinterpret_cast<float4 *>(pcm_row + base);
#pragma unroll
                        for (int it = 0; it < 20; it += 5) {
                            float4 v[5];
#pragma unroll
                            for (int u = 0; u < 5; ++u) v[u] = *reinterpret_cast<const float4 *>(sm.ybuf + ypad(4 * (lane + 32 * (it + u))));
#pragma unroll
                            for (int u = 0; u < 5; ++u) dst[lane + 32 * (it + u)] = v[u];
                        }
                    } else {
                        for (uint32_t i4 = lane * 4; base + i4 < own_end; i4 += 32 * 4) {
                            const float4 v = *reinterpret_cast<const float4 *>(sm.ybuf + ypad((int)i4));
                            const uint32_t n = base + i4;
                            if (n + 4 <= own_end) {
                                *reinterpret_cast<float4 *>(pcm_row + n) = v;
                            } else {
                                if (n < own_end) pcm_row[n] = v.x;
                                if (n + 1 < own_end) pcm_row[n + 1] = v.y;
                                if (n + 2 < own_end) pcm_row[n + 2] = v.z;
                            }
                        }
                    }
                }
                __syncwarp();
                named_bar_sync(2, FUSED_THREADS);                     // FFT warps and the VAD warp are done reading ybuf
                if (g + 1 < n_steps) {
                    for (int i = lane; i < CARRY; i += 32) sm.ybuf[ypad(i)] = sm.ybuf[ypad(i) + ypad(STEP_SAMPLES)];
                }
            } else {
                // ===== warps 0..7: one frame per half-warp =====
                if (M && warp * 2 < n_valid) {                        // warp-uniform: skip fully invalid pairs
                    const int hw = warp * 2 + half;
                    fft_frame(sm, sm.scr + hw * SCR_FLOATS_PER_FRAME, hw, l, lane, tw2r, winr);
                }
                __syncwarp();
                named_bar_arrive(2, FUSED_THREADS);                   // ybuf no longer needed by this warp
            }
            {
                const int mtid = tid;
                named_bar_sync(1, mel_threads);                       // pbuf complete

                // ---- phase 3: mel + log into the stage (thread = filter x 4 frames) ----
                if (M && n_valid > 0) {
                    const int n_items = (int)M * (SF / 4);
                    for (int item = mtid; item < n_items; item += mel_threads) {
                        const int m = item >> 2, fq = item & 3;
                        if (fq * 4 >= n_valid) continue;
                        const int lo = sm.mel.lo[m], cnt = sm.mel.cnt[m];
                        const float *w = sm.mel.w + sm.mel.off[m];
                        const float *pp = sm.pbuf + lo * PB_ROW + 4 * fq;
                        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
                        int j = 0;
                        for (; j + 2 <= cnt; j += 2) {
                            const float w0 = w[j], w1 = w[j + 1];
                            const float4 p0 = *reinterpret_cast<const float4 *>(pp + j * PB_ROW);
                            const float4 p1 = *reinterpret_cast<const float4 *>(pp + (j + 1) * PB_ROW);
                            a0 = fmaf(w0, p0.x, a0); a1 = fmaf(w0, p0.y, a1); a2 = fmaf(w0, p0.z, a2); a3 = fmaf(w0, p0.w, a3);
                            a0 = fmaf(w1, p1.x, a0); a1 = fmaf(w1, p1.y, a1); a2 = fmaf(w1, p1.z, a2); a3 = fmaf(w1, p1.w, a3);
                        }
                        if (j < cnt) {
                            const float w0 = w[j];
                            const float4 p0 = *reinterpret_cast<const float4 *>(pp + j * PB_ROW);
                            a0 = fmaf(w0, p0.x, a0); a1 = fmaf(w0, p0.y, a1); a2 = fmaf(w0, p0.z, a2); a3 = fmaf(w0, p0.w, a3);
                        }
                        // the 8 consecutive filters of a quarter-warp fill one 32-byte sector per frame row
                        if (lm_row) {
                            float *dst = lm_row + (uint64_t)(f0 + 4 * fq) * M + m;
                            const int nv = n_valid - 4 * fq;
                            dst[0] = __log2f(fmaxf(a0, P.log_floor)) * log_mul;
                            if (nv > 1) dst[M] = __log2f(fmaxf(a1, P.log_floor)) * log_mul;
                            if (nv > 2) dst[2 * M] = __log2f(fmaxf(a2, P.log_floor)) * log_mul;
                            if (nv > 3) dst[3 * M] = __log2f(fmaxf(a3, P.log_floor)) * log_mul;
                        }
                    }
                }
            }
            __syncthreads();        // ybuf carried, pbuf consumed: the next phase 1 / FFT may overwrite them
        }
        __syncthreads();
    }
}

size_t fused_smem_bytes() { return sizeof(FusedSmem); }

cudaError_t launch_fused(const FusedParams &P, int n_ctas, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(af_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(FusedSmem));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    af_fused_kernel<<<n_ctas, FUSED_THREADS, sizeof(FusedSmem), st>>>(P);
    return cudaGetLastError();
}

}  // namespace af
