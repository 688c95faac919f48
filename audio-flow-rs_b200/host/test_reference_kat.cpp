// The reference's hot-path unit tests (capture.rs:371-400, resampler.rs:181-204, vad.rs:207-298)
// transcribed against the C++ mirror; built with g++ and run on the GPU box by tests/test_parity_gpu.py.
#include <cstdio>
#include <cstdlib>

#include "audioflow.hpp"

using namespace audioflow;

#define REQUIRE(c)                                                         \
    do {                                                                   \
        if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

int main()
{
    if (af_init(0) != AF_OK) { char b[256]; af_last_error(b, 256); std::printf("no device: %s\n", b); return 2; }
    {   // test_audio_frame_to_mono_single_channel / _stereo
        AudioFrame f({0.5f, -0.5f}, 16000, 1, 1000);
        auto m = f.to_mono();
        REQUIRE(m.channels == 1 && m.samples == std::vector<float>({0.5f, -0.5f}));
        AudioFrame s({0.5f, 0.25f, -0.5f, -0.25f}, 16000, 2, 1000);
        auto ms = s.to_mono();
        REQUIRE(ms.channels == 1 && ms.samples.size() == 2);
        REQUIRE(std::fabs(ms.samples[0] - 0.375f) < 0.001f && std::fabs(ms.samples[1] + 0.375f) < 0.001f);
    }
    {   // test_no_resample_needed / test_resample_rates / test_same_rates_no_resampling
        AudioResampler r(16000, 16000);
        std::vector<float> in{0.1f, 0.2f, 0.3f, 0.4f};
        REQUIRE(r.process(in) == in);
        AudioResampler r2(48000, 16000);
        REQUIRE(r2.input_rate() == 48000 && r2.output_rate() == 16000 && r2.needs_resampling());
        REQUIRE(!AudioResampler(48000, 48000).needs_resampling());
        bool threw = false;
        try { r2.process(std::vector<float>(100, 0.0f)); } catch (const ResamplingFailed &) { threw = true; }
        REQUIRE(threw);
        std::vector<float> x(480000);
        for (size_t i = 0; i < x.size(); ++i) x[i] = std::sin(0.01f * (float)i);
        BatchResampler b(48000, 16000);
        auto y = b.process(x);
        auto tail = b.flush();
        y.insert(y.end(), tail.begin(), tail.end());
        REQUIRE(y.size() == 159998);
        for (size_t n = 1; n < y.size(); ++n) REQUIRE(y[n] == x[3 * n - 1]);
    }
    {   // test_vad_silence_detection / _speech_detection / _state_transitions / _reset / test_energy_calculation
        VadConfig c; c.threshold_db = -50.0f;
        VoiceActivityDetector v1(c), v2(c);
        REQUIRE(v1.detect(std::vector<float>(480, 0.0001f)) == VadState::Silence);
        REQUIRE(v2.detect(std::vector<float>(480, 0.5f)) == VadState::Speech);
        VadConfig t{-50.0f, 0.0f, 2, 1};
        VoiceActivityDetector v(t);
        REQUIRE(v.state() == VadState::Silence);
        std::vector<float> sp(480, 0.5f), si(480, 0.0001f);
        REQUIRE(v.detect(sp) == VadState::Speech);
        REQUIRE(v.detect(si) == VadState::Speech);
        REQUIRE(v.detect(si) == VadState::Ending);
        REQUIRE(v.detect(si) == VadState::Silence);
        VoiceActivityDetector w;
        w.detect(sp);
        REQUIRE(w.is_speaking());
        w.reset();
        REQUIRE(w.state() == VadState::Silence && !w.is_speaking());
        REQUIRE(w.calculate_energy(std::vector<float>(480, 0.0f)) == 0.0f);
        REQUIRE(std::fabs(w.calculate_energy(sp) - 0.25f) < 0.0001f);
    }
    std::printf("reference KATs through the C++ mirror: ok\n");
    return 0;
}
