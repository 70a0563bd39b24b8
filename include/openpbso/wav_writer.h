// openpbso drop-in, headless side: what the reference's PortAudio callback does with a SoundMessage
// (tools/real_time_modal_sound.cpp:207-210) -- out = (float)(data(i) / 1E10), duplicated to two channels --
// written to a RIFF/WAVE file (IEEE float32, stereo, SAMPLE_RATE) instead of the sound card.
#ifndef PBSO_WAV_WRITER_H
#define PBSO_WAV_WRITER_H
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "config.h"

namespace pbso_wav {

constexpr double OUTPUT_DIVISOR = 1E10;   // tools/real_time_modal_sound.cpp:208

class StereoFloatWriter {
    FILE* _f = nullptr;
    uint32_t _frames = 0;
    int _rate;
    static void u32(unsigned char* p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; p[2] = (v >> 16) & 255; p[3] = (v >> 24) & 255; }
    static void u16(unsigned char* p, uint16_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; }
    void header() {
        unsigned char h[58];
        const uint32_t data_bytes = _frames * 8u;
        std::memcpy(h, "RIFF", 4); u32(h + 4, 50 + data_bytes); std::memcpy(h + 8, "WAVE", 4);
        std::memcpy(h + 12, "fmt ", 4); u32(h + 16, 18); u16(h + 20, 3 /* IEEE float */); u16(h + 22, 2);
        u32(h + 24, (uint32_t)_rate); u32(h + 28, (uint32_t)_rate * 8u); u16(h + 32, 8); u16(h + 34, 32); u16(h + 36, 0);
        std::memcpy(h + 38, "fact", 4); u32(h + 42, 4); u32(h + 46, _frames);
        std::memcpy(h + 50, "data", 4); u32(h + 54, data_bytes);
        std::fseek(_f, 0, SEEK_SET);
        std::fwrite(h, 1, sizeof(h), _f);
    }

public:
    explicit StereoFloatWriter(const std::string& path, int sample_rate = SAMPLE_RATE) : _rate(sample_rate) {
        _f = std::fopen(path.c_str(), "wb");
        if (_f) header();
    }
    ~StereoFloatWriter() { close(); }
    bool ok() const { return _f != nullptr; }
    uint32_t frames() const { return _frames; }
    // One audio buffer: y[i] doubles in the solver's units; `volume` is the tool's slider (default 1).
    void write(const double* y, int n, double volume = 1.0) {
        if (!_f) return;
        std::vector<float> out(2 * (size_t)n);
        for (int i = 0; i < n; ++i) {
            const float s = (float)(y[i] / OUTPUT_DIVISOR * volume);
            out[2 * (size_t)i] = s; out[2 * (size_t)i + 1] = s;
        }
        std::fwrite(out.data(), sizeof(float), out.size(), _f);
        _frames += (uint32_t)n;
    }
    void close() {
        if (!_f) return;
        header();
        std::fclose(_f);
        _f = nullptr;
    }
};

}  // namespace pbso_wav
#endif
