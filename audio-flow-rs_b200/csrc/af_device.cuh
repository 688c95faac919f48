// af_device.cuh -- device helpers shared by the kernels: bit-exact reference arithmetic
// (downmix, cubic interpolation, frame energy, VAD step) and the register FFT.
#pragma once
#include "af_common.cuh"

namespace af {

// ------------------------------------------------------------------------------------------
// Downmix: AudioFrame::to_mono (capture.rs:30-42).  Sequential sum from 0.0, one division.
// Frames outside [0, n_in) read as 0 (the resampler's zero history / flush padding,
// resampler.rs:150-158).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_mono(const void *__restrict__ data, uint64_t n_samples, uint32_t n_in,
                                           uint32_t channels, uint32_t format, int idx)
{
    if (idx < 0 || (uint32_t)idx >= n_in) return 0.0f;
    if (channels == 1) {
        if (format == FMT_F32) return __ldg(reinterpret_cast<const float *>(data) + idx);
        return (float)__ldg(reinterpret_cast<const short *>(data) + idx) * (1.0f / 32768.0f);
    }
    const uint64_t base = (uint64_t)idx * channels;
    if (channels == 2 && base + 2 <= n_samples) {
        float l, r;
        if (format == FMT_F32) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(data) + idx);
            l = v.x; r = v.y;
        } else {
            const short2 v = __ldg(reinterpret_cast<const short2 *>(data) + idx);
            l = (float)v.x * (1.0f / 32768.0f); r = (float)v.y * (1.0f / 32768.0f);
        }
        return __fmul_rn(__fadd_rn(__fadd_rn(0.0f, l), r), 0.5f);   // x / 2 == x * 0.5 exactly
    }
    uint32_t m = channels;
    if (base + m > n_samples) m = (uint32_t)(n_samples - base);     // trailing partial frame
    float sum = 0.0f;
    for (uint32_t c = 0; c < m; ++c) {
        float v = format == FMT_F32 ? __ldg(reinterpret_cast<const float *>(data) + base + c)
                                    : (float)__ldg(reinterpret_cast<const short *>(data) + base + c) * (1.0f / 32768.0f);
        sum = __fadd_rn(sum, v);
    }
    return __fdiv_rn(sum, (float)channels);
}

// ------------------------------------------------------------------------------------------
// rubato 0.16.2 interp_cubic (points at -1, 0, 1, 2), evaluated with the exact operation
// order of the Rust source and no FMA contraction.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float interp_cubic(float x, float y0, float y1, float y2, float y3)
{
    const float c13 = 1.0f / 3.0f, c16 = 1.0f / 6.0f;
    const float a0 = y1;
    const float a1 = __fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(-c13, y0), __fmul_rn(0.5f, y1)), y2), __fmul_rn(c16, y3));
    const float a2 = __fsub_rn(__fmul_rn(0.5f, __fadd_rn(y0, y2)), y1);
    const float a3 = __fadd_rn(__fmul_rn(0.5f, __fsub_rn(y1, y2)), __fmul_rn(c16, __fsub_rn(y3, y0)));
    const float x2 = __fmul_rn(x, x);
    const float x3 = __fmul_rn(x2, x);
    return __fadd_rn(__fadd_rn(__fadd_rn(a0, __fmul_rn(a1, x)), __fmul_rn(a2, x2)), __fmul_rn(a3, x3));
}

// Position of output n of the reference resampler in input-sample units:
//   P_n = -4 + (n + 1) * p / q          (last_index = -4, one step of p/q per output)
// returned as k = floor(P_n) and rem = (P_n - k) * q, computed exactly in integers.
__host__ __device__ __forceinline__ void resample_pos(uint64_t n, uint32_t p, uint32_t q, long long *k, uint32_t *rem)
{
    const uint64_t num = (n + 1) * (uint64_t)p + 4ull * q;   // shifted by +8q to stay non-negative
    *k = (long long)(num / q) - 8;
    *rem = (uint32_t)(num % q);
}

// Plans one tile: the exact position of its first output and the eight stage fills (two per step of 32 frames).
// A fill covers the taps of the outputs [lo_i, hi_i) of its half step: bytes rounded out to 16, never past the
// stream's last full 16 bytes, and only when it fits the stage; `interior` says which resampling path may run.
__host__ __device__ inline void plan_tile(const StreamDev &s, uint32_t stream, uint32_t tile, TileDev *t)
{
    const uint32_t n_tile0 = tile * (uint32_t)TILE_SAMPLES;
    const uint32_t tile_end = s.n_out - n_tile0 < (uint32_t)TILE_SAMPLES ? s.n_out : n_tile0 + (uint32_t)TILE_SAMPLES;
    const int parts = s.staged == 4u ? 4 : 2;
    t->stream = stream; t->tile = tile;
    long long k0 = 0; uint32_t rem0 = 0;
    if (s.mode != RS_PASSTHROUGH) resample_pos(n_tile0, s.p, s.q, &k0, &rem0);
    t->k0 = (int32_t)k0; t->rem0 = rem0;
    t->n_steps = (tile_end - n_tile0 + STEP_SAMPLES - 1) / STEP_SAMPLES;
    t->tile_end = tile_end;
    t->n_frames = s.n_frames; t->parts = (uint32_t)parts;
    t->sdesc = s;
    {
        const uint32_t inc = (uint32_t)RS_THREADS * s.p;
        t->inc_k = s.q ? inc / s.q : 0; t->inc_rem = s.q ? inc % s.q : 0;
    }
    const uint32_t ch = s.channels, bps = s.format == FMT_I16 ? 2u : 4u;
    const uint32_t left = s.n_out - n_tile0;                       // outputs from the tile start to the stream end
    for (int j = 0; j < TILE_FILLS; ++j) {
        const uint32_t g = (uint32_t)j / (uint32_t)parts;
        const int h = j % parts;
        FillDesc d;
        d.src = reinterpret_cast<const char *>(s.data); d.bytes = 0; d.lo = 0; d.hi = 0; d.interior = 0; d.pad_[0] = d.pad_[1] = 0;
        const uint32_t d_first = g * STEP_SAMPLES + (uint32_t)(h == 0 ? (g == 0 ? 0 : CARRY) : part_end(parts, h - 1));
        const uint32_t d_full = g * STEP_SAMPLES + (uint32_t)part_end(parts, h);
        const uint32_t d_last = d_full < left ? d_full : left;
        if (g < t->n_steps && d_first < d_last) {
            // floor(position) of output n_tile0 + d (the host guarantees (TILE_SAMPLES + YLEN) * p + q < 2^32)
            long long ka, kb;
            if (s.mode == RS_PASSTHROUGH) { ka = (long long)n_tile0 + d_first; kb = (long long)n_tile0 + d_last - 1; }
            else {
                ka = k0 + (long long)((rem0 + d_first * s.p) / s.q);
                kb = k0 + (long long)((rem0 + (d_last - 1) * s.p) / s.q);
            }
            // interior half step: no tap (k-2 .. k+2) leaves the stream, no output beyond n_out, whole channel frames
            const bool geom = (ka - 2 >= 0) && (kb + 3 <= (long long)s.n_in) &&
                              ((unsigned long long)(kb + 3) * ch <= s.n_samples) && (d_last == d_full) && ch <= 2;
            if (geom) d.interior = 2;                            // unchecked taps straight from global memory
            long long i_lo = ka - 2, i_hi = kb + 3;
            if (i_lo < 0) i_lo = 0;
            if (i_hi > (long long)s.n_in) i_hi = (long long)s.n_in;
            if (s.staged && i_lo < i_hi) {
                unsigned long long b_lo = ((unsigned long long)i_lo * ch * bps) & ~15ull;
                unsigned long long e_hi = (unsigned long long)i_hi * ch;
                if (e_hi > s.n_samples) e_hi = s.n_samples;
                unsigned long long b_hi = (e_hi * bps + 15ull) & ~15ull;
                const unsigned long long b_end = (s.n_samples * bps) & ~15ull;   // never read past the stream's last full 16 bytes
                if (b_hi > b_end) b_hi = b_end;
                if (b_hi > b_lo && b_hi - b_lo <= (unsigned long long)STAGE_BYTES) {
                    d.bytes = (uint32_t)(b_hi - b_lo);
                    d.lo = (uint32_t)(b_lo / bps); d.hi = (uint32_t)(b_hi / bps);
                    d.src += b_lo;
                    if (geom && (unsigned long long)(kb + 3) * ch <= d.hi) d.interior = 1;   // ... and from the stage
                }
            }
        }
        t->fill[j] = d;
    }
}

// ------------------------------------------------------------------------------------------
// VAD arithmetic (vad.rs:101-153), one frame step on a precomputed mean-square energy.
// The dB comparison `20*log10(e) > threshold_db` is replaced by `e >= e_min` where e_min is the
// smallest f32 that satisfies it under the host libm (NaN when none does); see DESIGN.md.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int vad_step(VadState &v, const VadParams &pr, float energy)
{
    const float sm = __fadd_rn(__fmul_rn(pr.alpha, energy), __fmul_rn(__fsub_rn(1.0f, pr.alpha), v.smoothed));
    v.smoothed = sm;
    const float det = pr.alpha > 0.0f ? sm : energy;
    const bool is_speech = det >= pr.e_min;
    if (v.state == 0) {                       // Silence
        if (is_speech) { v.speech_frames = 1; v.silence_frames = 0; v.state = 1; }
    } else if (v.state == 1) {                // Speech
        if (is_speech) { v.speech_frames += 1; v.silence_frames = 0; }
        else {
            v.silence_frames += 1;
            if (v.silence_frames >= pr.silence_timeout) {
                v.state = v.speech_frames >= pr.min_speech ? 2 : 0;
                v.speech_frames = 0;
            }
        }
    } else {                                  // Ending -> Silence, unconditionally
        v.state = 0;
        v.silence_frames = 0;
    }
    return v.state;
}

// ------------------------------------------------------------------------------------------
// Register FFT building blocks (forward transform, e^{-i...}).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void fft4(float &r0, float &i0, float &r1, float &i1, float &r2, float &i2, float &r3,
                                     float &i3)
{
    const float t0r = r0 + r2, t0i = i0 + i2, t1r = r0 - r2, t1i = i0 - i2;
    const float t2r = r1 + r3, t2i = i1 + i3, t3r = r1 - r3, t3i = i1 - i3;
    r0 = t0r + t2r; i0 = t0i + t2i;
    r2 = t0r - t2r; i2 = t0i - t2i;
    r1 = t1r + t3i; i1 = t1i - t3r;     // t1 - i t3
    r3 = t1r - t3i; i3 = t1i + t3r;     // t1 + i t3
}
// same with input 3 known to be zero
__device__ __forceinline__ void fft4_z3(float &r0, float &i0, float &r1, float &i1, float &r2, float &i2, float &r3,
                                        float &i3)
{
    const float t0r = r0 + r2, t0i = i0 + i2, t1r = r0 - r2, t1i = i0 - i2;
    const float ur = r1, ui = i1;
    r0 = t0r + ur; i0 = t0i + ui;
    r2 = t0r - ur; i2 = t0i - ui;
    r1 = t1r + ui; i1 = t1i - ur;
    r3 = t1r - ui; i3 = t1i + ur;
}

#define AF_CMUL(ar, ai, wr, wi)                       \
    do {                                              \
        const float _tr = (ar) * (wr) - (ai) * (wi);  \
        const float _ti = (ar) * (wi) + (ai) * (wr);  \
        (ar) = _tr; (ai) = _ti;                       \
    } while (0)

// In-register 16-point FFT, 4x4 decomposition n = 4a + b, k = c + 4d.
// Input x[n] natural order; output X[k] is left in slot 4*(k&3) + (k>>2).
// PRUNED: inputs 13, 14, 15 are known zeros (a 400-sample window in a 512-point transform).
template <bool PRUNED>
__device__ __forceinline__ void fft16(float (&xr)[16], float (&xi)[16])
{
    constexpr float C1 = 0.92387953251128674f;   // cos(pi/8)
    constexpr float S1 = 0.38268343236508977f;   // sin(pi/8)
    constexpr float R = 0.70710678118654752f;
    fft4(xr[0], xi[0], xr[4], xi[4], xr[8], xi[8], xr[12], xi[12]);
    if (PRUNED) {
        fft4_z3(xr[1], xi[1], xr[5], xi[5], xr[9], xi[9], xr[13], xi[13]);
        fft4_z3(xr[2], xi[2], xr[6], xi[6], xr[10], xi[10], xr[14], xi[14]);
        fft4_z3(xr[3], xi[3], xr[7], xi[7], xr[11], xi[11], xr[15], xi[15]);
    } else {
        fft4(xr[1], xi[1], xr[5], xi[5], xr[9], xi[9], xr[13], xi[13]);
        fft4(xr[2], xi[2], xr[6], xi[6], xr[10], xi[10], xr[14], xi[14]);
        fft4(xr[3], xi[3], xr[7], xi[7], xr[11], xi[11], xr[15], xi[15]);
    }
    // slot 4c + b holds Y[b][c]; multiply by W16^(b c)
    AF_CMUL(xr[5], xi[5], C1, -S1);          // b=1 c=1 : W^1
    AF_CMUL(xr[9], xi[9], R, -R);            // b=1 c=2 : W^2
    AF_CMUL(xr[13], xi[13], S1, -C1);        // b=1 c=3 : W^3
    AF_CMUL(xr[6], xi[6], R, -R);            // b=2 c=1 : W^2
    { const float t = xr[10]; xr[10] = xi[10]; xi[10] = -t; }   // b=2 c=2 : W^4 = -i
    AF_CMUL(xr[14], xi[14], -R, -R);         // b=2 c=3 : W^6
    AF_CMUL(xr[7], xi[7], S1, -C1);          // b=3 c=1 : W^3
    AF_CMUL(xr[11], xi[11], -R, -R);         // b=3 c=2 : W^6
    AF_CMUL(xr[15], xi[15], -C1, S1);        // b=3 c=3 : W^9
    fft4(xr[0], xi[0], xr[1], xi[1], xr[2], xi[2], xr[3], xi[3]);
    fft4(xr[4], xi[4], xr[5], xi[5], xr[6], xi[6], xr[7], xi[7]);
    fft4(xr[8], xi[8], xr[9], xi[9], xr[10], xi[10], xr[11], xi[11]);
    fft4(xr[12], xi[12], xr[13], xi[13], xr[14], xi[14], xr[15], xi[15]);
}

// ------------------------------------------------------------------------------------------
// Packed variant: sm_100a executes add / mul / fma on register PAIRS (PTX .f32x2, SASS FADD2 / FMUL2 /
// FFMA2): one issue slot for two independent fp32 operations.  A "pack" holds the same component (real or
// imaginary) of two different points, so the +-i rotations of the butterflies stay plain adds between the real
// and the imaginary packs -- no data movement.  Points 2k and 2k+1 of the 16-point transform share pack k.
// ------------------------------------------------------------------------------------------
struct f2 { float x, y; };
__device__ __forceinline__ f2 mk2(float a, float b) { f2 r; r.x = a; r.y = b; return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 r;
    asm("{.reg .b64 a, b, c; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; add.rn.f32x2 c, a, b; mov.b64 {%0, %1}, c;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b)
{
    f2 r;
    asm("{.reg .b64 a, b, c; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; sub.rn.f32x2 c, a, b; mov.b64 {%0, %1}, c;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
    f2 r;
    asm("{.reg .b64 a, b, c; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; mul.rn.f32x2 c, a, b; mov.b64 {%0, %1}, c;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
    f2 r;
    asm("{.reg .b64 a, b, c, d; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; mov.b64 c, {%6, %7}; fma.rn.f32x2 d, a, b, c; "
        "mov.b64 {%0, %1}, d;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
// (r + i i) (c + i s) on both halves
__device__ __forceinline__ void cmul2(f2 &r, f2 &i, f2 c, f2 s)
{
    const f2 tr = fma2(mk2(-i.x, -i.y), s, mul2(r, c));
    const f2 ti = fma2(r, s, mul2(i, c));
    r = tr; i = ti;
}
// ------------------------------------------------------------------------------------------
// rubato's interp_cubic on TWO outputs at once in packed f32x2 arithmetic, bit for bit the scalar interp_cubic above.
// ptxas contracts a packed multiply with the packed addition that follows it into FFMA2 even when both carry .rn (and
// also when the product is written fma(a, b, -0) with a literal zero, and with -fmad=false: tools/ubench/cubic2_check.cu),
// which changes the rounding.  So every product here is fma(a, b, nz) with nz = -0.0f that only the host knows (a kernel
// argument): ptxas cannot fold the addend away, there is nothing left to contract, and a * b + (-0) is RN(a * b) for every
// a, b -- zero products of either sign, subnormals, Inf and NaN included.  11 FFMA2 + 11 FADD2 for two outputs instead
// of 22 FMUL + 22 FADD.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ f2 interp_cubic2(f2 x, f2 y0, f2 y1, f2 y2, f2 y3, float nz)
{
    const float c13 = 1.0f / 3.0f, c16 = 1.0f / 6.0f;
    const f2 z = mk2(nz, nz), h = mk2(0.5f, 0.5f), s6 = mk2(c16, c16);
    const f2 a1 = sub2(add2(sub2(fma2(mk2(-c13, -c13), y0, z), fma2(h, y1, z)), y2), fma2(s6, y3, z));
    const f2 a2 = sub2(fma2(h, add2(y0, y2), z), y1);
    const f2 a3 = add2(fma2(h, sub2(y1, y2), z), fma2(s6, sub2(y3, y0), z));
    const f2 x2 = fma2(x, x, z);
    const f2 x3 = fma2(x2, x, z);
    return add2(add2(add2(y1, fma2(a1, x, z)), fma2(a2, x2, z)), fma2(a3, x3, z));
}

// radix-4 butterfly on two transforms at once (the halves of the packs), in place; output c lands where input c was
__device__ __forceinline__ void fft4p(f2 &r0, f2 &i0, f2 &r1, f2 &i1, f2 &r2, f2 &i2, f2 &r3, f2 &i3)
{
    const f2 t0r = add2(r0, r2), t0i = add2(i0, i2), t1r = sub2(r0, r2), t1i = sub2(i0, i2);
    const f2 t2r = add2(r1, r3), t2i = add2(i1, i3), t3r = sub2(r1, r3), t3i = sub2(i1, i3);
    r0 = add2(t0r, t2r); i0 = add2(t0i, t2i);
    r2 = sub2(t0r, t2r); i2 = sub2(t0i, t2i);
    r1 = add2(t1r, t3i); i1 = sub2(t1i, t3r);     // t1 - i t3
    r3 = sub2(t1r, t3i); i3 = add2(t1i, t3r);     // t1 + i t3
}
// the same with input 3 known to be zero in both halves
__device__ __forceinline__ void fft4p_z3(f2 &r0, f2 &i0, f2 &r1, f2 &i1, f2 &r2, f2 &i2, f2 &r3, f2 &i3)
{
    const f2 t0r = add2(r0, r2), t0i = add2(i0, i2), t1r = sub2(r0, r2), t1i = sub2(i0, i2);
    const f2 ur = r1, ui = i1;
    r0 = add2(t0r, ur); i0 = add2(t0i, ui);
    r2 = sub2(t0r, ur); i2 = sub2(t0i, ui);
    r1 = add2(t1r, ui); i1 = sub2(t1i, ur);
    r3 = sub2(t1r, ui); i3 = add2(t1i, ur);
}
// radix-4 butterfly over the four slots held by two packs A = (s0, s1), B = (s2, s3): first level packed, second scalar
__device__ __forceinline__ void fft4h(f2 &Ar, f2 &Ai, f2 &Br, f2 &Bi)
{
    const f2 Tr = add2(Ar, Br), Ti = add2(Ai, Bi);     // (t0, t2)
    const f2 Ur = sub2(Ar, Br), Ui = sub2(Ai, Bi);     // (t1, t3)
    Ar.x = Tr.x + Tr.y; Ai.x = Ti.x + Ti.y;            // z0
    Br.x = Tr.x - Tr.y; Bi.x = Ti.x - Ti.y;            // z2
    Ar.y = Ur.x + Ui.y; Ai.y = Ui.x - Ur.y;            // z1 = t1 - i t3
    Br.y = Ur.x - Ui.y; Bi.y = Ui.x + Ur.y;            // z3 = t1 + i t3
}
// W16^(b c) for the packs 2c + j (b = 2j, 2j+1), c = 1..3: {cos pair, sin pair} per pack.  In constant memory so
// that the packed multiplies take them as uniform-register operands instead of rebuilding them in registers.
static __constant__ float2 c_fft16_tw[12] = {
    {1.0f, 0.92387953251128674f},                 {0.0f, -0.38268343236508977f},                  // c = 1, b = 0, 1
    {0.70710678118654752f, 0.38268343236508977f}, {-0.70710678118654752f, -0.92387953251128674f}, // c = 1, b = 2, 3
    {1.0f, 0.70710678118654752f},                 {0.0f, -0.70710678118654752f},                  // c = 2, b = 0, 1
    {0.0f, -0.70710678118654752f},                {-1.0f, -0.70710678118654752f},                 // c = 2, b = 2, 3
    {1.0f, 0.38268343236508977f},                 {0.0f, -0.92387953251128674f},                  // c = 3, b = 0, 1
    {-0.70710678118654752f, -0.92387953251128674f}, {-0.70710678118654752f, 0.38268343236508977f} // c = 3, b = 2, 3
};
__device__ __forceinline__ f2 fft16_tw(int i) { return mk2(c_fft16_tw[i].x, c_fft16_tw[i].y); }

// In-register 16-point FFT on packs: point n in pack n >> 1, half n & 1 (same 4 x 4 decomposition and the same
// output placement as fft16: X[k] in slot 4 (k & 3) + (k >> 2), slot s = pack s >> 1, half s & 1).
// PRUNED: points 13, 14, 15 are known zeros.
template <bool PRUNED>
__device__ __forceinline__ void fft16p(f2 (&R)[8], f2 (&I)[8])
{
    // stage A: transforms b = (0, 1) live in packs 0, 2, 4, 6; b = (2, 3) in packs 1, 3, 5, 7
    fft4p(R[0], I[0], R[2], I[2], R[4], I[4], R[6], I[6]);
    if (PRUNED) fft4p_z3(R[1], I[1], R[3], I[3], R[5], I[5], R[7], I[7]);
    else fft4p(R[1], I[1], R[3], I[3], R[5], I[5], R[7], I[7]);
    // slot 4c + b (pack 2c + (b >> 1)) times W16^(b c)
#pragma unroll
    for (int p = 2; p < 8; ++p) cmul2(R[p], I[p], fft16_tw(2 * p - 4), fft16_tw(2 * p - 3));
    // stage B: the four slots of each c
#pragma unroll
    for (int c = 0; c < 4; ++c) fft4h(R[2 * c], I[2 * c], R[2 * c + 1], I[2 * c + 1]);
}
__host__ __device__ constexpr int fft16_slot(int k) { return 4 * (k & 3) + (k >> 2); }

}  // namespace af
