// af_common.cuh -- constants and device-visible tables shared by the kernels and the host runtime.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace af {

// ---- feature geometry (DESIGN.md "Feature spec"; absent from the reference) ----
constexpr int OUT_RATE = 16000;
constexpr int WIN = 400;          // 25 ms @ 16 kHz
constexpr int HOP = 160;          // 10 ms
constexpr int NFFT = 512;
constexpr int NBIN = NFFT / 2 + 1;   // 257
constexpr int MAX_MELS = 128;

// ---- reference resampler geometry (resampler.rs:47,55; rubato FastFixedIn) ----
constexpr int RS_CHUNK = 128;     // chunk_size
constexpr int RS_POLY = 8;        // rubato POLYNOMIAL_LEN

// ---- fused kernel tiling ----
constexpr int SF = 16;                       // frames per step: one per half-warp of the 8 FFT warps
constexpr int STEP_SAMPLES = SF * HOP;       // 2560 new 16 kHz samples per step
constexpr int CARRY = WIN - HOP;             // 240 samples shared with the next step
constexpr int YLEN = STEP_SAMPLES + CARRY;   // 2800 samples live per step
constexpr int TILE_FRAMES = 128;             // frames per tile (work unit of one CTA)
constexpr int TILE_SAMPLES = TILE_FRAMES * HOP;   // 20480
constexpr int FFT_WARPS = 8;
constexpr int VAD_WARP = FFT_WARPS;          // warp 8: stage fills (TMA) + sequential frame energies
constexpr int AUX_WARP = FFT_WARPS + 1;      // warp 9: PCM write-out + overlap carry
constexpr int FUSED_WARPS = FFT_WARPS + 2;
constexpr int FUSED_THREADS = FUSED_WARPS * 32;   // 320: 640 resample quads, 640 PCM float4, 320 mel items per step

// padded index of 16 kHz sample i inside the step buffer: 4 pad words after every 32 samples so
// that the 32 VAD lanes (frame starts 160 apart) hit distinct bank quads with LDS.128
__host__ __device__ constexpr int ypad(int i) { return i + 4 * (i >> 5); }
constexpr int YBUF_FLOATS = ((ypad(YLEN + 32) + 31) / 32) * 32;

constexpr int SCR_ROW = 18;                  // complex per transposed row (16 + 2 pad -> LDS.128 conflict free)
constexpr int SCR_FLOATS_PER_FRAME = 16 * SCR_ROW * 2;   // 576 floats = 2304 B
constexpr int PB_ROW = SF + 4;               // floats per power row: 16 frames + 4 pad (16-byte aligned rows)
constexpr int PBUF_FLOATS = NBIN * PB_ROW;
constexpr int STAGE_BYTES = 34816;           // TMA-staged raw input of one step (34 KB)

// formats / flags (mirror include/audioflow_gpu.h)
enum : uint16_t { FMT_F32 = 0, FMT_I16 = 1 };

// one stream of a batch, device resident
struct StreamDev {
    const void *data;        // interleaved samples
    uint64_t n_samples;      // total interleaved samples
    uint32_t n_in;           // mono frames = ceil(n_samples / channels)
    uint32_t n_out;          // resampled length (BatchResampler all + flush)
    uint32_t n_frames;       // STFT frames
    uint32_t n_vad_frames;   // VAD frames
    uint16_t channels;
    uint16_t format;
    uint32_t p, q;           // input step per output = p / q (reduced)
    uint32_t mode;           // RS_*
    const float *frac;       // RS_TABLE: f32 fractional offsets by output index
    uint32_t tile_begin;     // first global tile of this stream
    uint32_t n_tiles;
    uint32_t staged;         // 1: the step input fits the shared-memory stage (bulk-copy path)
    uint32_t pad_;
};
enum : uint32_t { RS_PASSTHROUGH = 0, RS_EXACT = 1, RS_TABLE = 2 };

struct TileDev {
    uint32_t stream;
    uint32_t tile;           // tile index inside the stream
};

// mel filterbank in compact form (weights already carry the 1/4 of the unscaled power)
struct MelTables {
    uint16_t lo[MAX_MELS];   // first nonzero bin
    uint16_t cnt[MAX_MELS];  // number of nonzero bins
    uint16_t off[MAX_MELS];  // offset into w
    uint16_t n_w;
    uint16_t n_mels;
    float w[2 * NBIN + 2 * MAX_MELS];
};

// constant tables of the FFT, filled by the host in f64 and rounded once
struct FftTables {
    float window[416];       // periodic Hann, zero beyond 400
    float2 tw1[16 * 16];     // tw1[k1 * 16 + l] = exp(-2 pi i l k1 / 256)
    float2 tw2[128];         // tw2[k] = exp(-2 pi i k / 512)
};

struct VadParams {
    float alpha;             // smoothing_factor
    float e_min;             // smallest f32 energy with 20*log10f(e) > threshold_db (host libm); +inf if none
    unsigned long long silence_timeout;
    unsigned long long min_speech;
};

struct VadState {
    float smoothed;
    int state;
    unsigned long long silence_frames;
    unsigned long long speech_frames;
};

struct FusedParams {
    const StreamDev *streams;
    const TileDev *tiles;
    uint32_t n_tiles;
    const FftTables *fft;
    const MelTables *mel;
    float *pcm;   uint64_t pcm_stride;
    float *logmel; uint64_t logmel_stride;
    float *energy; uint64_t energy_stride;
    uint32_t n_mels;         // 0: skip STFT/mel
    uint32_t do_energy;      // VAD energies on the STFT frames
    float log_floor;
    float log_scale;         // ln(2) (natural log) or log10(2): multiplies log2(mel)
    uint32_t use_stage;      // 0: plain global loads everywhere ("sync" variant)
};

}  // namespace af
