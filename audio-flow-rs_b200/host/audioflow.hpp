// audioflow.hpp -- C++17 host-side mirror of the reference's audio module
// (src-tauri/src/modules/audio/mod.rs:9-11) over the C ABI of libaudioflow_gpu.so.
// The reference is compiled code (Rust) whose toolchain is absent from the build image, so this
// header is the compiled-language stand-in for the Rust shim in ../rust: same type names, method
// names, argument meaning and error behaviour.  Header only; contains no arithmetic of the path.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/audioflow_gpu.h"

namespace audioflow {

// AudioError::ResamplingFailed (src-tauri/src/error.rs:109-110)
struct ResamplingFailed : std::runtime_error {
    explicit ResamplingFailed(const std::string &m) : std::runtime_error("Resampling failed: " + m) {}
};

inline void check(int rc)
{
    if (rc == AF_OK) return;
    char buf[512];
    af_last_error(buf, sizeof(buf));
    throw ResamplingFailed(buf);
}

// capture.rs:11-42
struct AudioFrame {
    std::vector<float> samples;
    uint32_t sample_rate = 0;
    uint16_t channels = 1;
    unsigned __int128 timestamp_ns = 0;

    AudioFrame(std::vector<float> s, uint32_t rate, uint16_t ch, unsigned __int128 ts = 0)
        : samples(std::move(s)), sample_rate(rate), channels(ch), timestamp_ns(ts) {}

    AudioFrame to_mono() const
    {
        if (channels == 1) return *this;
        const size_t frames = (samples.size() + channels - 1) / channels;
        std::vector<float> mono(frames);
        size_t n = 0;
        check(af_to_mono(samples.data(), samples.size(), channels, mono.data(), frames, &n));
        mono.resize(n);
        return AudioFrame(std::move(mono), sample_rate, 1, timestamp_ns);
    }
};

// resampler.rs:12-112
class AudioResampler {
public:
    AudioResampler(uint32_t input_rate, uint32_t output_rate) : in_(input_rate), out_(output_rate)
    {
        check(af_resampler_create(input_rate, output_rate, &h_));
    }
    static AudioResampler create_48k_to_16k() { return AudioResampler(48000, 16000); }
    AudioResampler(AudioResampler &&o) noexcept : h_(o.h_), in_(o.in_), out_(o.out_) { o.h_ = nullptr; }
    AudioResampler(const AudioResampler &) = delete;
    ~AudioResampler() { if (h_) af_resampler_destroy(h_); }

    std::vector<float> process(const std::vector<float> &input)
    {
        std::vector<float> out(std::max(input.size(), af_resample_max_output(in_, out_, 128)));
        size_t n = 0;
        check(af_resampler_process(h_, input.data(), input.size(), out.data(), out.size(), &n));
        out.resize(n);
        return out;
    }
    uint32_t input_rate() const { return in_; }
    uint32_t output_rate() const { return out_; }
    bool needs_resampling() const { return in_ != out_; }

private:
    af_resampler *h_ = nullptr;
    uint32_t in_, out_;
};

// resampler.rs:115-166
class BatchResampler {
public:
    BatchResampler(uint32_t input_rate, uint32_t output_rate) : in_(input_rate), out_(output_rate)
    {
        check(af_batch_resampler_create(input_rate, output_rate, &h_));
    }
    BatchResampler(const BatchResampler &) = delete;
    ~BatchResampler() { if (h_) af_batch_resampler_destroy(h_); }
    std::vector<float> process(const std::vector<float> &input)
    {
        std::vector<float> out(af_resample_max_output(in_, out_, input.size() + 128));
        size_t n = 0;
        check(af_batch_resampler_process(h_, input.data(), input.size(), out.data(), out.size(), &n));
        out.resize(n);
        return out;
    }
    std::vector<float> flush()
    {
        std::vector<float> out(af_resample_max_output(in_, out_, 128));
        size_t n = 0;
        check(af_batch_resampler_flush(h_, out.data(), out.size(), &n));
        out.resize(n);
        return out;
    }

private:
    af_batch_resampler *h_ = nullptr;
    uint32_t in_, out_;
};

// capture.rs:84-161 -- the capture hand-off.  write() keeps one slot free and returns the count written; read() returns
// nothing (the reference's None) on an empty ring, else min(size, available) samples.
class RingBuffer {
public:
    explicit RingBuffer(size_t capacity_samples) { check(af_ring_create(capacity_samples, &h_)); capacity = af_ring_capacity(h_); }
    RingBuffer(const RingBuffer &) = delete;
    ~RingBuffer() { if (h_) af_ring_destroy(h_); }
    size_t write(const std::vector<float> &data) { return af_ring_write(h_, data.data(), data.size()); }
    bool read(size_t size, std::vector<float> *out)
    {
        out->assign(size, 0.0f);
        size_t n = 0;
        const int rc = af_ring_read(h_, out->data(), size, &n);
        if (rc == AF_RING_EMPTY) { out->clear(); return false; }
        check(rc);
        out->resize(n);
        return true;
    }
    size_t available() const { return af_ring_available(h_); }
    void clear() { af_ring_clear(h_); }
    af_ring *handle() { return h_; }
    size_t capacity = 0;

private:
    af_ring *h_ = nullptr;
};

// vad.rs:8-54
enum class VadLevel { Aggressive, Balanced, Relaxed };
enum class VadState { Silence = 0, Speech = 1, Ending = 2 };

struct VadConfig {
    float threshold_db = -50.0f;
    float smoothing_factor = 0.3f;
    size_t silence_timeout_frames = 15;
    size_t min_speech_frames = 3;
};

// vad.rs:60-205
class VoiceActivityDetector {
public:
    explicit VoiceActivityDetector(const VadConfig &c = VadConfig())
    {
        af_vad_config cc{c.threshold_db, c.smoothing_factor, c.silence_timeout_frames, c.min_speech_frames};
        check(af_vad_create(&cc, &h_));
    }
    VoiceActivityDetector(const VoiceActivityDetector &) = delete;
    ~VoiceActivityDetector() { if (h_) af_vad_destroy(h_); }
    VadState detect(const std::vector<float> &frame)
    {
        uint8_t s = 0;
        check(af_vad_detect(h_, frame.data(), frame.size(), &s));
        return static_cast<VadState>(s);
    }
    void reset() { check(af_vad_reset(h_)); }
    VadState state() const { return static_cast<VadState>(af_vad_state(h_)); }
    float energy_db() const { return af_vad_energy_db(h_); }
    bool is_speaking() const { return af_vad_is_speaking(h_) != 0; }
    size_t speech_frame_count() const { return af_vad_speech_frame_count(h_); }
    float calculate_energy(const std::vector<float> &frame) const
    {
        float e = 0;
        check(af_vad_frame_energy(frame.data(), frame.size(), &e));
        return e;
    }

private:
    af_vad *h_ = nullptr;
};

}  // namespace audioflow
