// af_plan.cpp -- host-side planning (see af_plan.h).
#include "af_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

namespace af {

void RsRecurrence::init(uint32_t in, uint32_t out)
{
    in_rate = in; out_rate = out;
    passthrough = (in == out);                       // resampler.rs:33-40
    const uint32_t g = std::gcd(in, out);
    p = in / g; q = out / g;
    const double ratio = (double)out / (double)in;   // intended resample_ratio
    t = 1.0 / ratio;
    end_idx = (long)RS_CHUNK - (RS_POLY + 1) - (long)std::ceil(t);
    exact = (q & (q - 1)) == 0 && q <= (1u << 20) && t == (double)p / (double)q;
    last_index = -(double)(RS_POLY / 2);
    chunks = 0; n_out = 0;
}

uint32_t RsRecurrence::step(std::vector<float> *frac)
{
    double idx = last_index;
    uint32_t n = 0;
    const double end = (double)end_idx;
    while (idx < end) {
        idx += t;
        const double fl = std::floor(idx);
        if (frac) frac->push_back((float)(idx - fl));
        ++n;
    }
    last_index = idx - (double)RS_CHUNK;
    chunks += 1;
    n_out += n;
    return n;
}

uint64_t rs_exact_count(const RsRecurrence &r, uint64_t chunks)
{
    if (chunks == 0) return 0;
    // outputs n >= 0 with -4 + n t < 128 (chunks - 1) + end_idx   <=>   n p < (128 chunks - 128 + end_idx + 4) q
    const uint64_t bound = (uint64_t)((long long)RS_CHUNK * (long long)chunks - RS_CHUNK + r.end_idx + 4);
    const __uint128_t num = (__uint128_t)bound * r.q;
    return (uint64_t)((num + r.p - 1) / r.p);
}

void build_fft_tables(FftTables *t)
{
    std::memset(t, 0, sizeof(*t));
    for (int n = 0; n < WIN; ++n) t->window[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)WIN));
    for (int k1 = 0; k1 < 16; ++k1)
        for (int l = 0; l < 16; ++l) {
            const double a = -2.0 * M_PI * (double)(l * k1) / 256.0;
            t->tw1[k1 * 16 + l] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int k = 0; k < 128; ++k) {
        const double a = -2.0 * M_PI * (double)k / 512.0;
        t->tw2[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
}

static double hz_to_mel_htk(double f) { return 2595.0 * std::log10(1.0 + f / 700.0); }
static double mel_to_hz_htk(double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); }

bool build_mel_tables(uint32_t n_mels, float f_min, float f_max, MelTables *t)
{
    std::memset(t, 0, sizeof(*t));
    if (n_mels == 0 || n_mels > MAX_MELS) return false;
    std::vector<double> pts(n_mels + 2);
    const double m_lo = hz_to_mel_htk((double)f_min), m_hi = hz_to_mel_htk((double)f_max);
    for (uint32_t i = 0; i < n_mels + 2; ++i)
        pts[i] = mel_to_hz_htk(m_lo + (m_hi - m_lo) * (double)i / (double)(n_mels + 1));
    uint32_t off = 0;
    for (uint32_t m = 0; m < n_mels; ++m) {
        int lo = -1, hi = -1;
        std::vector<float> w(NBIN, 0.0f);
        for (int k = 0; k < NBIN; ++k) {
            const double f = (double)k * (double)OUT_RATE / (double)NFFT;
            const double up = (f - pts[m]) / (pts[m + 1] - pts[m]);
            const double dn = (pts[m + 2] - f) / (pts[m + 2] - pts[m + 1]);
            double v = up < dn ? up : dn;
            if (v < 0.0) v = 0.0;
            w[k] = (float)v;
            if (w[k] != 0.0f) { if (lo < 0) lo = k; hi = k + 1; }
        }
        if (lo < 0) { lo = 0; hi = 0; }
        if (off + (uint32_t)(hi - lo) > sizeof(t->w) / sizeof(float)) return false;
        t->lo[m] = (uint16_t)lo; t->cnt[m] = (uint16_t)(hi - lo); t->off[m] = (uint16_t)off;
        for (int k = lo; k < hi; ++k) t->w[off++] = w[k] * 0.25f;    // pbuf holds 4 |X|^2
    }
    t->n_w = (uint16_t)off; t->n_mels = (uint16_t)n_mels;
    return true;
}

static inline float bits_to_f32(uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; }

float vad_energy_threshold(float threshold_db)
{
    // energy_to_dbfs (vad.rs:171-176): e <= 0 -> -inf; else 20 * log10(e).  is_speech = dbfs > thr.
    auto speech = [&](float e) { return 20.0f * log10f(e) > threshold_db; };
    const uint32_t inf_bits = 0x7f800000u;
    if (!speech(bits_to_f32(inf_bits))) return std::numeric_limits<float>::quiet_NaN();
    uint32_t lo = 1, hi = inf_bits;              // invariant: speech(hi) true; everything below lo false
    if (speech(bits_to_f32(lo))) return bits_to_f32(lo);
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (speech(bits_to_f32(mid))) hi = mid; else lo = mid;
    }
    return bits_to_f32(hi);
}

}  // namespace af
