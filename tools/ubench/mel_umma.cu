// mel_umma.cu -- the A/B the north star asks for: the mel projection + log as (S) the sparse banded FMA of the product
// (lane = frame, weights walked filter by filter) against (T) a dense tensor-core contraction on tcgen05.mma
// (kind::tf32, accumulators in TMEM), both standalone on the same power spectra:
//
//     logmel[f][m] = ln(max(sum_k W[m][k] * P[f][k], 1e-10)),   257 bins, 80 (or 128) HTK mel filters.
//
// (T) is written the way it would have to live inside the fused kernel, and given its best shot:
//   * tf32 keeps 11 significant bits -- 4.9e-4 relative, i.e. ~5e-4 on the log, above the 1e-4 tolerance -- so every
//     operand is split hi + lo (hi = the top 11 bits, lo = the exact remainder) and three MMAs are issued per K step:
//     hi*hi + lo*hi + hi*lo (the "3 x tf32" scheme);
//   * the filterbank is banded, so K is cut into chunks of 64 bins and each chunk only multiplies the N columns of the
//     filters that overlap it (48 / 32 / 16 / 16 / 16 instead of 80): the weights of all chunks stay resident in shared
//     memory (65 KB for hi + lo) and the tensor core does 35 % of the dense work;
//   * A = power (M = 128 frames, K-major, no-swizzle canonical layout: 8-row x 16-byte core matrices), written by the
//     threads that hold the spectrum after an in-register hi / lo split -- the fused kernel's FFT lanes would have to do
//     exactly this -- double buffered per chunk; D = 128 lanes x 96 columns of TMEM; one thread issues, tcgen05.commit
//     releases the chunk buffers, four warps run the epilogue (tcgen05.ld -> max -> log -> 16-byte stores).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o mel_umma mel_umma.cu
//   ./mel_umma [n_mels=80] [tiles_per_sm=8]
//
// Prints the error of both against an f64 evaluation, the time per launch and the derived rates.  The instruction and
// shared-memory wavefront counts come from ncu on this binary (tools/gpu_mel_ab.sh -> profiles/r02_mel_ab_*).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); std::exit(3); } } while (0)

constexpr int NBIN = 257, KP = 264;          // bins, padded row length (floats) of the power rows
constexpr int TILE = 128;                    // frames per tile = UMMA M
constexpr int MAXM = 128;                    // filters
constexpr int CHUNK = 64;                    // bins per K chunk
constexpr int MAX_CHUNKS = 5;
constexpr float LOG_FLOOR = 1e-10f;

// ------------------------------------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(100000u) : "memory");
        if (spin > (1u << 22)) __trap();                 // a lost arrival must not hang the box
    }
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
                   "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])::"memory");
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
}
// shared-memory matrix descriptor, no swizzle, K-major: core matrices of 8 rows x 16 bytes; LBO = bytes between the two
// 16-byte K chunks of one instruction, SBO = bytes between 8-row groups; version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
           (1ull << 46);
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------------------
// (S) sparse banded FMA, lane = frame (the product's formulation: af_fused.cu role_mel)
// ------------------------------------------------------------------------------------------------------------------
struct SparseBank {                   // per filter: first bin (a multiple of 4), number of weight quadruples, offset into w
    int start4[MAXM], c4[MAXM], woff[MAXM];
    int n_mels, n_w;
};
constexpr int S_ROW = 268;            // padded power row in shared memory: 268 / 4 is odd -> conflict-free LDS.128 by lane = row

__global__ void __launch_bounds__(TILE) mel_sparse_kernel(const float *__restrict__ P, int n_tiles, const SparseBank *__restrict__ bank,
                                                          const float *__restrict__ wts, float *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *rows = reinterpret_cast<float *>(smem_raw);                       // [TILE][S_ROW]
    float *w = rows + TILE * S_ROW;                                          // compact weights
    __shared__ SparseBank sb;
    const int tid = threadIdx.x;
    for (int i = tid; i < (int)(sizeof(SparseBank) / 4); i += TILE) reinterpret_cast<int *>(&sb)[i] = reinterpret_cast<const int *>(bank)[i];
    __syncthreads();
    for (int i = tid; i < sb.n_w; i += TILE) w[i] = wts[i];
    const int M = sb.n_mels;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        const float4 *src = reinterpret_cast<const float4 *>(P + (size_t)tile * TILE * KP);
        for (int i = tid; i < TILE * (KP / 4); i += TILE) {                  // coalesced rows -> padded shared rows
            const int r = i / (KP / 4), c = i % (KP / 4);
            *reinterpret_cast<float4 *>(rows + r * S_ROW + 4 * c) = __ldcs(src + i);
        }
        __syncthreads();
        const float *row = rows + tid * S_ROW;
        float *dst = out + ((size_t)tile * TILE + tid) * M;
        for (int m0 = 0; m0 < M; m0 += 4) {
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = m0 + u;
                const float4 *x4 = reinterpret_cast<const float4 *>(row + sb.start4[m]);
                const float4 *w4 = reinterpret_cast<const float4 *>(w + sb.woff[m]);
                float a0 = 0.0f, a1 = 0.0f;
                for (int j = 0; j < sb.c4[m]; ++j) {
                    const float4 x = x4[j], ww = w4[j];
                    a0 = fmaf(ww.x, x.x, a0); a1 = fmaf(ww.y, x.y, a1);
                    a0 = fmaf(ww.z, x.z, a0); a1 = fmaf(ww.w, x.w, a1);
                }
                o[u] = __logf(fmaxf(a0 + a1, LOG_FLOOR));
            }
            __stcs(reinterpret_cast<float4 *>(dst + m0), make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// (T) tcgen05.mma kind::tf32, 3 x split, banded K chunks
// ------------------------------------------------------------------------------------------------------------------
struct ChunkPlan {
    int n_chunks;
    int k0[MAX_CHUNKS], klen[MAX_CHUNKS];        // bins [k0, k0 + klen), klen a multiple of 8
    int n0[MAX_CHUNKS], nn[MAX_CHUNKS];          // TMEM columns [n0, n0 + nn) = the filters that overlap the chunk (multiples of 16)
    int woff[MAX_CHUNKS];                        // byte offset of the chunk's W_hi block in the weight area; W_lo follows it
    int w_bytes;                                 // total bytes of the weight area
    int n_mels, d_cols;                          // output filters, TMEM columns in use (multiple of 16)
};
constexpr int A_BYTES = TILE * CHUNK * 4;        // one operand copy (hi or lo) of one chunk: 32 KB

// n_split: 3 = hi*hi + lo*hi + hi*lo (the variant that meets the tolerance); 1 = plain tf32 (hi*hi only), to show what the
// split buys
__global__ void __launch_bounds__(TILE) mel_umma_kernel(const float *__restrict__ P, int n_tiles, const ChunkPlan *__restrict__ plan_g,
                                                        const unsigned char *__restrict__ w_canon, float *__restrict__ out, int n_split)
{
    extern __shared__ __align__(128) unsigned char smem[];
    // [A hi/lo buffer 0][A hi/lo buffer 1][W blocks]
    unsigned char *a_buf = smem;
    unsigned char *w_buf = smem + 4 * A_BYTES;
    __shared__ ChunkPlan plan;
    __shared__ __align__(8) uint64_t bar_empty[2], bar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (int)(sizeof(ChunkPlan) / 4); i += TILE) reinterpret_cast<int *>(&plan)[i] = reinterpret_cast<const int *>(plan_g)[i];
    if (tid == 0) {
        mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1); mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    for (int i = tid; i < plan.w_bytes / 16; i += TILE) reinterpret_cast<uint4 *>(w_buf)[i] = reinterpret_cast<const uint4 *>(w_canon)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
    __syncthreads();

    const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
    uint32_t uses[2] = {0, 0};
    uint32_t done_phase = 0;
    const int M = plan.n_mels;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // zero the accumulator columns (the chunks' column ranges overlap: every MMA accumulates)
        for (int c = 0; c < plan.d_cols; c += 16) tmem_st16_zero(lane_base + c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        const float *src = P + (size_t)tile * TILE * KP;
        for (int c = 0; c < plan.n_chunks; ++c) {
            const int b = c & 1;
            if (uses[b]) mbar_wait(&bar_empty[b], (uses[b] - 1) & 1u);          // the MMAs that read this buffer have completed
            ++uses[b];
            unsigned char *a_hi = a_buf + (2 * b) * A_BYTES, *a_lo = a_hi + A_BYTES;
            const int k0 = plan.k0[c], klen = plan.klen[c], k4n = klen / 4;
            const uint32_t sbo = (uint32_t)k4n * 128u;
            // power chunk -> hi / lo -> canonical K-major layout: element (r, k) at (r / 8) sbo + (r % 8) 16 + (k / 4) 128 + (k % 4) 4
            for (int q = tid; q < TILE * k4n; q += TILE) {
                int r, k4;
                if (k4n >= 4) { r = ((q >> 5) % 16) * 8 + (q & 7); k4 = (q >> 9) * 4 + ((q >> 3) & 3); }
                else { r = (q / (8 * k4n)) * 8 + (q & 7); k4 = (q >> 3) % k4n; }
                const float4 x = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)r * KP + k0 + 4 * k4));
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); lo.x = x.x - hi.x;
                hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); lo.y = x.y - hi.y;
                hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); lo.z = x.z - hi.z;
                hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); lo.w = x.w - hi.w;
                const uint32_t off = (uint32_t)(r >> 3) * sbo + (uint32_t)(r & 7) * 16u + (uint32_t)k4 * 128u;
                *reinterpret_cast<float4 *>(a_hi + off) = hi;
                *reinterpret_cast<float4 *>(a_lo + off) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t idesc = make_idesc(TILE, plan.nn[c]);
                const uint32_t d = tmem + (uint32_t)plan.n0[c];
                const uint32_t wh = smem_u32(w_buf + plan.woff[c]), wl = wh + (uint32_t)(plan.nn[c] * klen * 4);
                const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo);
                for (int kk = 0; kk < klen / 8; ++kk) {
                    const uint64_t dah = make_desc(ah + kk * 256, 128, sbo), dal = make_desc(al + kk * 256, 128, sbo);
                    const uint64_t dwh = make_desc(wh + kk * 256, 128, sbo), dwl = make_desc(wl + kk * 256, 128, sbo);
                    umma_tf32(d, dah, dwh, idesc, 1u);
                    if (n_split == 3) {
                        umma_tf32(d, dal, dwh, idesc, 1u);
                        umma_tf32(d, dah, dwl, idesc, 1u);
                    }
                }
                umma_commit(&bar_empty[b]);
                if (c == plan.n_chunks - 1) umma_commit(&bar_done);
            }
        }
        // epilogue: lane = frame
        mbar_wait(&bar_done, done_phase);
        done_phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float *dst = out + ((size_t)tile * TILE + tid) * M;
        for (int c = 0; c < M; c += 16) {
            float v[16];
            tmem_ld16(lane_base + c, v);
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                __stcs(reinterpret_cast<float4 *>(dst + c + j),
                       make_float4(__logf(fmaxf(v[j], LOG_FLOOR)), __logf(fmaxf(v[j + 1], LOG_FLOOR)), __logf(fmaxf(v[j + 2], LOG_FLOOR)),
                                   __logf(fmaxf(v[j + 3], LOG_FLOOR))));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------------
static double hz2mel(double f) { return 2595.0 * std::log10(1.0 + f / 700.0); }
static double mel2hz(double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); }

int main(int argc, char **argv)
{
    const int n_mels = argc > 1 ? std::atoi(argv[1]) : 80;
    const int tiles_per_sm = argc > 2 ? std::atoi(argv[2]) : 8;
    if (n_mels % 16 || n_mels > MAXM) { std::printf("n_mels must be a multiple of 16 <= 128\n"); return 2; }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int n_tiles = prop.multiProcessorCount * tiles_per_sm, n_frames = n_tiles * TILE;

    // HTK filterbank (the definition of DESIGN.md section 4), dense f32
    std::vector<float> W((size_t)n_mels * KP, 0.0f);
    {
        std::vector<double> pts(n_mels + 2);
        const double lo = hz2mel(0.0), hi = hz2mel(8000.0);
        for (int i = 0; i < n_mels + 2; ++i) pts[i] = mel2hz(lo + (hi - lo) * i / (n_mels + 1));
        for (int m = 0; m < n_mels; ++m)
            for (int k = 0; k < NBIN; ++k) {
                const double f = k * 16000.0 / 512.0;
                const double up = (f - pts[m]) / (pts[m + 1] - pts[m]), dn = (pts[m + 2] - f) / (pts[m + 2] - pts[m + 1]);
                const double v = std::fmax(0.0, std::fmin(up, dn));
                W[(size_t)m * KP + k] = (float)v;
            }
    }
    // power spectra with speech-like dynamic range: a smooth envelope, harmonics, a noise floor 60-90 dB below the peaks
    std::vector<float> P((size_t)n_frames * KP, 0.0f);
    {
        uint64_t s = 0x1234567ull;
        auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)((s >> 11) & 0xfffffffffffffull) / 4503599627370496.0; };
        for (int f = 0; f < n_frames; ++f) {
            const double level = std::pow(10.0, -6.0 * rnd()), f0 = 3.0 + 20.0 * rnd(), tilt = 0.01 + 0.03 * rnd();
            for (int k = 0; k < NBIN; ++k) {
                const double harm = std::pow(std::cos(3.14159265358979 * k / f0), 2.0 * 8);
                const double env = std::exp(-tilt * k);
                const double noise = 1e-8 * (0.2 + rnd());
                P[(size_t)f * KP + k] = (float)(level * (env * harm * (0.5 + rnd()) * 100.0 + noise));
            }
        }
    }
    // f64 reference
    std::vector<float> ref((size_t)n_frames * n_mels);
    std::vector<int> lo_bin(n_mels), hi_bin(n_mels);
    for (int m = 0; m < n_mels; ++m) {
        int lo = NBIN, hi = 0;
        for (int k = 0; k < NBIN; ++k) if (W[(size_t)m * KP + k] != 0.0f) { lo = std::min(lo, k); hi = std::max(hi, k + 1); }
        if (lo > hi) { lo = 0; hi = 0; }
        lo_bin[m] = lo; hi_bin[m] = hi;
    }
    for (int f = 0; f < n_frames; ++f)
        for (int m = 0; m < n_mels; ++m) {
            double acc = 0.0;
            for (int k = lo_bin[m]; k < hi_bin[m]; ++k) acc += (double)W[(size_t)m * KP + k] * (double)P[(size_t)f * KP + k];
            ref[(size_t)f * n_mels + m] = (float)std::log(std::fmax(acc, (double)LOG_FLOOR));
        }

    // ---- (S) tables ----
    SparseBank sbank{};
    std::vector<float> sw;
    sbank.n_mels = n_mels;
    for (int m = 0; m < n_mels; ++m) {
        const int s4 = lo_bin[m] & ~3, e4 = (hi_bin[m] + 3) & ~3;
        sbank.start4[m] = s4; sbank.c4[m] = (e4 - s4) / 4; sbank.woff[m] = (int)sw.size();
        for (int k = s4; k < e4; ++k) sw.push_back(k < KP ? W[(size_t)m * KP + k] : 0.0f);
    }
    sbank.n_w = (int)sw.size();

    // ---- (T) chunk plan + canonical hi / lo weight blocks ----
    ChunkPlan plan{};
    std::vector<unsigned char> wcanon;
    plan.n_mels = n_mels; plan.d_cols = n_mels;
    for (int k0 = 0; k0 < KP; k0 += CHUNK) {
        const int klen = std::min(CHUNK, KP - k0);
        int mlo = n_mels, mhi = 0;
        for (int m = 0; m < n_mels; ++m)
            if (lo_bin[m] < k0 + klen && hi_bin[m] > k0) { mlo = std::min(mlo, m); mhi = std::max(mhi, m + 1); }
        if (mlo >= mhi) continue;
        const int c = plan.n_chunks++;
        int n0 = mlo & ~15, nn = ((mhi - n0) + 15) & ~15;
        if (n0 + nn > n_mels) n0 = n_mels - nn;                      // stay inside the allocated columns
        plan.k0[c] = k0; plan.klen[c] = klen; plan.n0[c] = n0; plan.nn[c] = nn; plan.woff[c] = (int)wcanon.size();
        const int k4n = klen / 4, sbo = k4n * 128;
        std::vector<unsigned char> blk((size_t)2 * nn * klen * 4, 0);
        for (int part = 0; part < 2; ++part)
            for (int n = 0; n < nn; ++n)
                for (int k = 0; k < klen; ++k) {
                    const float w = (k0 + k < KP) ? W[(size_t)(n0 + n) * KP + k0 + k] : 0.0f;
                    uint32_t hb; std::memcpy(&hb, &w, 4); hb &= 0xffffe000u;
                    float hi; std::memcpy(&hi, &hb, 4);
                    const float v = part == 0 ? hi : (w - hi);
                    const size_t off = (size_t)part * nn * klen * 4 + (size_t)(n >> 3) * sbo + (size_t)(n & 7) * 16 + (size_t)(k >> 2) * 128 + (size_t)(k & 3) * 4;
                    std::memcpy(&blk[off], &v, 4);
                }
        wcanon.insert(wcanon.end(), blk.begin(), blk.end());
    }
    plan.w_bytes = (int)wcanon.size();
    std::printf("%s, %d SMs; %d frames x %d bins -> %d mels; UMMA chunks:", prop.name, prop.multiProcessorCount, n_frames, NBIN, n_mels);
    long mma_cols = 0;
    for (int c = 0; c < plan.n_chunks; ++c) { std::printf(" [k %d+%d, n %d+%d]", plan.k0[c], plan.klen[c], plan.n0[c], plan.nn[c]); mma_cols += (long)plan.klen[c] * plan.nn[c]; }
    std::printf("  (%.0f %% of the dense %d x %d; weights %d KB resident)\n", 100.0 * mma_cols / ((double)KP * n_mels), KP, n_mels, plan.w_bytes / 1024);

    float *dP, *dOut, *dSw;
    SparseBank *dBank; ChunkPlan *dPlan; unsigned char *dWc;
    CK(cudaMalloc(&dP, P.size() * 4)); CK(cudaMalloc(&dOut, ref.size() * 4)); CK(cudaMalloc(&dSw, sw.size() * 4));
    CK(cudaMalloc(&dBank, sizeof(SparseBank))); CK(cudaMalloc(&dPlan, sizeof(ChunkPlan))); CK(cudaMalloc(&dWc, wcanon.size()));
    CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dSw, sw.data(), sw.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBank, &sbank, sizeof(sbank), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dPlan, &plan, sizeof(plan), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dWc, wcanon.data(), wcanon.size(), cudaMemcpyHostToDevice));

    const size_t smem_s = (size_t)TILE * S_ROW * 4 + sw.size() * 4, smem_t = (size_t)4 * A_BYTES + wcanon.size();
    CK(cudaFuncSetAttribute(mel_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    CK(cudaFuncSetAttribute(mel_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    const int grid = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<float> got(ref.size());
    const double bytes = (double)n_frames * (KP * 4.0 + n_mels * 4.0);

    for (int which = 0; which < 3; ++which) {
        CK(cudaMemset(dOut, 0xff, ref.size() * 4));
        auto launch = [&]() {
            if (which == 0) mel_sparse_kernel<<<grid, TILE, smem_s>>>(dP, n_tiles, dBank, dSw, dOut);
            else mel_umma_kernel<<<grid, TILE, smem_t>>>(dP, n_tiles, dPlan, dWc, dOut, which == 1 ? 3 : 1);
        };
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        const int reps = 20;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) launch();
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
        CK(cudaMemcpy(got.data(), dOut, got.size() * 4, cudaMemcpyDeviceToHost));
        double max_abs = 0, se = 0, sr = 0; size_t over = 0, nan = 0;
        for (size_t i = 0; i < got.size(); ++i) {
            if (!(got[i] == got[i])) { ++nan; continue; }
            const double d = std::fabs((double)got[i] - (double)ref[i]);
            max_abs = std::fmax(max_abs, d); se += d * d; sr += (double)ref[i] * ref[i]; over += d > 1e-4;
        }
        std::printf("%-34s %8.3f ms  %7.2f M frames/s  %7.1f GB/s (P in + log-mel out)  max |err| %.3e  rel L2 %.3e  bins > 1e-4: %zu  NaN: %zu\n",
                    which == 0 ? "(S) sparse banded FMA, lane = frame" : (which == 1 ? "(T) tcgen05.mma tf32 x3, banded" : "(T1) tcgen05.mma plain tf32"), ms, n_frames / ms / 1e3, bytes / ms / 1e6,
                    max_abs, std::sqrt(se / sr), over, nan);
    }
    return 0;
}
