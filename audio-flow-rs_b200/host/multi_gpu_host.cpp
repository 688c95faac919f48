// A host that uses ONLY the C ABI (include/audioflow_gpu.h) -- no CUDA runtime, no PyTorch, no Python: what the Rust host of
// the north star would do to drive all the GPUs of one box from a single process.
//
//   multi_gpu_host [n_gpus] [n_streams] [seconds]
//
// af_init_multi(n) -> a sharded batch of synthetic mixed 44.1 / 48 kHz streams in pinned host memory -> one
// af_sharded_batch_run_host call (every GPU fed by its own host thread) -> the same streams one by one on GPU 0
// (af_pipeline_run) -> every byte of PCM / log-mel / VAD must agree.  Exit code 0 = identical.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "audioflow_gpu.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != AF_OK) {                                                                \
            char msg_[512];                                                                \
            af_last_error(msg_, sizeof(msg_));                                             \
            std::printf("FAIL %s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, rc_, msg_); \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

static uint64_t splitmix(uint64_t &s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char **argv)
{
    int n_gpus = argc > 1 ? std::atoi(argv[1]) : 0;
    const size_t S = argc > 2 ? (size_t)std::atol(argv[2]) : 24;
    const double seconds = argc > 3 ? std::atof(argv[3]) : 3.0;
    int have = 0;
    af_device_count(&have);
    if (have == 0) { std::printf("no CUDA device (no CPU fallback)\n"); return 2; }
    if (n_gpus <= 0 || n_gpus > have) n_gpus = have;
    CHECK(af_init_multi(n_gpus));
    std::printf("%s: %d GPU(s), communicator of %d rank(s)\n", af_version(), n_gpus, af_comm_size());

    // synthetic streams (loud / quiet segments so that the VAD has something to decide), pinned host memory
    std::vector<af_stream_desc> descs(S);
    std::vector<float *> data(S);
    for (size_t i = 0; i < S; ++i) {
        const uint32_t rate = (i & 1) ? 44100u : 48000u;
        const size_t n = (size_t)(seconds * rate) + 37 * i;
        void *p = nullptr;
        CHECK(af_host_alloc(&p, n * sizeof(float)));
        data[i] = static_cast<float *>(p);
        uint64_t seed = 0xA0D10F10ull + i;
        for (size_t k = 0; k < n; ++k) {
            const double t = (double)k / rate;
            const double amp = (((size_t)(t * 2.0) + i) % 3 == 0) ? 0.3 : 0.001;
            const double u = (double)(splitmix(seed) >> 11) / 9007199254740992.0 - 0.5;
            data[i][k] = (float)(amp * std::sin(6.283185307179586 * (220.0 + 13.0 * i) * t) + 0.002 * u);
        }
        descs[i] = af_stream_desc{data[i], n, rate, 1, AF_FMT_F32};
    }
    af_pipeline_config cfg;
    af_pipeline_config_default(&cfg);
    af_pipeline *pipe = nullptr;
    CHECK(af_pipeline_create(&cfg, &pipe));

    af_sharded_batch *sb = nullptr;
    CHECK(af_sharded_batch_create(pipe, descs.data(), S, AF_MEM_HOST, &sb));
    uint64_t pcm_stride = 0, lm_stride = 0, vad_stride = 0;
    for (int r = 0; r < n_gpus; ++r) {
        size_t first = 0, count = 0; int dev = -1;
        CHECK(af_sharded_batch_shard(sb, r, &first, &count, &dev));
        uint64_t a = 0, b = 0, c = 0;
        if (count) CHECK(af_batch_strides(af_sharded_batch_local(sb, r), &a, &b, &c));
        pcm_stride = a > pcm_stride ? a : pcm_stride; lm_stride = b > lm_stride ? b : lm_stride; vad_stride = c > vad_stride ? c : vad_stride;
        std::printf("  rank %d on GPU %d: streams [%zu, %zu)\n", r, dev, first, first + count);
    }
    std::vector<float> pcm(S * pcm_stride), lm(S * lm_stride);
    std::vector<uint8_t> vad(S * vad_stride, 0xff);
    std::vector<af_vad_final> fin(S);
    af_outputs out{pcm.data(), pcm_stride, lm.data(), lm_stride, vad.data(), vad_stride, nullptr, 0, fin.data()};
    CHECK(af_sharded_batch_run_host(sb, &out));            // warm-up (first touch of every GPU)
    const auto t0 = std::chrono::steady_clock::now();
    CHECK(af_sharded_batch_run_host(sb, &out));
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double audio_s = 0;
    for (size_t i = 0; i < S; ++i) audio_s += (double)descs[i].n_samples / descs[i].sample_rate;
    std::printf("  %zu streams, %.1f audio-s in %.2f ms through host buffers on %d GPU(s): %.0f audio-s/s\n", S, audio_s, dt * 1e3, n_gpus,
                audio_s / dt);

    // the same streams one at a time on GPU 0
    CHECK(af_init(0));
    size_t bad = 0, speech = 0;
    for (size_t i = 0; i < S; ++i) {
        af_batch *b1 = nullptr;
        CHECK(af_batch_create(pipe, &descs[i], 1, AF_MEM_HOST, &b1));
        uint32_t n_out = 0, n_feat = 0, n_vad = 0;
        CHECK(af_batch_counts(b1, &n_out, &n_feat, &n_vad));
        uint64_t ps = 0, ls = 0, vs = 0;
        CHECK(af_batch_strides(b1, &ps, &ls, &vs));
        std::vector<float> p1(ps), l1(ls);
        std::vector<uint8_t> v1(vs);
        af_vad_final f1;
        af_outputs o1{p1.data(), ps, l1.data(), ls, v1.data(), vs, nullptr, 0, &f1};
        CHECK(af_batch_run_host(b1, &o1));
        af_batch_destroy(b1);
        if (std::memcmp(p1.data(), &pcm[i * pcm_stride], n_out * sizeof(float))) ++bad;
        if (std::memcmp(l1.data(), &lm[i * lm_stride], (size_t)n_feat * cfg.n_mels * sizeof(float))) ++bad;
        if (std::memcmp(v1.data(), &vad[i * vad_stride], n_vad)) ++bad;
        if (f1.state != fin[i].state || f1.speech_frames != fin[i].speech_frames || std::memcmp(&f1.smoothed_energy, &fin[i].smoothed_energy, 4)) ++bad;
        for (uint32_t f = 0; f < n_vad; ++f) speech += v1[f] == AF_VAD_SPEECH;
    }
    std::printf("  mismatching outputs: %zu, speech frames: %zu\n", bad, speech);
    af_sharded_batch_destroy(sb);
    af_pipeline_destroy(pipe);
    for (float *p : data) af_host_free(p);
    CHECK(af_shutdown());
    if (bad || speech == 0) { std::printf("FAIL\n"); return 1; }
    std::printf("OK\n");
    return 0;
}
