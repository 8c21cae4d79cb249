// sdr_b200.hpp — C++ host-side mirror of the reference's operator interface over the C ABI.
//
// The reference is compiled Rust; no Rust toolchain exists in this image, so the host side above
// include/sdr_b200.h is C++ (task brief ②).  Class and method names follow examples/simple_fm.rs:
//   struct DemodConfig :179-185, optimal_settings :189-214, struct Demod :232-427
//   (new, demodulate, rotate_90, low_pass_complex, fm_demod, low_pass_real, fast_atan2),
// and RtlSdr::read_sync (src/lib.rs:153-155) for the source.  Errors the reference reports as
// Result::Err / panics surface as sdr::Error (code + message from sdr_last_error()).
#pragma once

#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sdr_b200.h"

namespace sdr {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

inline long check(long rc) {
    if (rc < 0) throw Error((int)rc, std::string(sdr_last_error()));
    return rc;
}

using DemodConfig = sdr_demod_config;   // examples/simple_fm.rs:179-185
using RadioConfig = sdr_radio_config;   // :173-176
constexpr uint32_t FREQUENCY = 94'900'000, SAMPLE_RATE = 170'000, RATE_RESAMPLE = 32'000;   // :25-27
constexpr size_t DEFAULT_BUF_LENGTH = SDR_DEFAULT_BUF_LENGTH;                                 // src/lib.rs:25

// A caller-owned buffer page-locked for the lifetime of this object (sdr_host_register): calls that are handed a pointer
// inside it copy by DMA straight from it.  The reader's long-lived Box<[u8; DEFAULT_BUF_LENGTH]> (:114) is the use case.
class Registered {
  public:
    Registered(void *p, size_t bytes) : p_(p) { check(sdr_host_register(p, bytes)); }
    ~Registered() { sdr_host_unregister(p_); }
    Registered(const Registered &) = delete;
    Registered &operator=(const Registered &) = delete;

  private:
    void *p_;
};

// optimal_settings(freq, rate) -> (RadioConfig, DemodConfig), :189-214
inline std::pair<RadioConfig, DemodConfig> optimal_settings(uint32_t freq = FREQUENCY, uint32_t rate = SAMPLE_RATE) {
    RadioConfig r{};
    DemodConfig d{};
    check(sdr_optimal_settings(freq, rate, SAMPLE_RATE, RATE_RESAMPLE, &r, &d));
    return {r, d};
}

using Complex32 = std::complex<int32_t>;   // layout-compatible with num_complex::Complex<i32> {re, im}

class Demod {
public:
    explicit Demod(const DemodConfig &config, int cuda_device = 0) : config(config) {   // Demod::new :243
        check(sdr_demod_new(&config, cuda_device, &h_));
    }
    ~Demod() { sdr_demod_free(h_); }
    Demod(const Demod &) = delete;
    Demod &operator=(const Demod &) = delete;

    // demodulate(&mut self, Vec<u8>) -> Vec<i16>, :256-269
    std::vector<int16_t> demodulate(const std::vector<uint8_t> &buf) {
        long cap = check(sdr_demod_out_len(h_, buf.size()));
        std::vector<int16_t> out((size_t)cap);
        long n = check(sdr_demod_demodulate(h_, buf.data(), buf.size(), out.data(), out.size()));
        out.resize((size_t)n);
        return out;
    }
    // n consecutive demodulate() calls in one pipelined submission (bit-identical to n calls)
    std::vector<int16_t> demodulate_batch(const uint8_t *bufs, size_t buf_len, size_t n_bufs) {
        std::vector<int16_t> out(buf_len * n_bufs / 2 / config.downsample + 16);
        long n = check(sdr_demod_demodulate_batch(h_, bufs, buf_len, n_bufs, out.data(), out.size(), nullptr));
        out.resize((size_t)n);
        return out;
    }
    // rotate_90(Vec<u8>) -> Vec<u8>, scalar branch :276-299
    std::vector<uint8_t> rotate_90(std::vector<uint8_t> buf) {
        check(sdr_rotate_90(h_, buf.data(), buf.size()));
        return buf;
    }
    // low_pass_complex(&mut self, Vec<Complex<i32>>) -> Vec<Complex<i32>>, :337-352
    std::vector<Complex32> low_pass_complex(const std::vector<Complex32> &buf) {
        std::vector<Complex32> out(buf.size() / config.downsample + 2);
        long n = check(sdr_low_pass_complex(h_, reinterpret_cast<const int32_t *>(buf.data()), buf.size(),
                                            reinterpret_cast<int32_t *>(out.data()), out.size()));
        out.resize((size_t)n);
        return out;
    }
    // fm_demod(&mut self, Vec<Complex<i32>>) -> Vec<i16>, :355-367
    std::vector<int16_t> fm_demod(const std::vector<Complex32> &buf) {
        std::vector<int16_t> out(buf.size());
        check(sdr_fm_demod(h_, reinterpret_cast<const int32_t *>(buf.data()), buf.size(), out.data(), out.size()));
        return out;
    }
    // low_pass_real(&mut self, Vec<i16>) -> Vec<i16>, :408-426
    std::vector<int16_t> low_pass_real(const std::vector<int16_t> &buf) {
        std::vector<int16_t> out(buf.size() + 2);
        long n = check(sdr_low_pass_real(h_, buf.data(), buf.size(), out.data(), out.size()));
        out.resize((size_t)n);
        return out;
    }
    // fast_atan2(y, x), :383-405 (element-wise)
    std::vector<int32_t> fast_atan2(const std::vector<int32_t> &y, const std::vector<int32_t> &x) {
        std::vector<int32_t> out(y.size());
        check(sdr_fast_atan2(h_, y.data(), x.data(), y.size(), out.data()));
        return out;
    }
    sdr_demod_state state() const {
        sdr_demod_state s{};
        check(sdr_demod_get_state(h_, &s));
        return s;
    }
    void set_state(const sdr_demod_state &s) { check(sdr_demod_set_state(h_, &s)); }
    sdr_demod *raw() { return h_; }

    DemodConfig config;

private:
    sdr_demod *h_ = nullptr;
};

// Persistent ring around a Demod (sdr_demod_ring_*): the reader -> channel -> processor pair of
// examples/simple_fm.rs:55-60 with the channel living in pinned memory and ONE resident kernel behind it.
// acquire/commit belong to the producer thread, collect to the consumer thread.
class Ring {
public:
    Ring(Demod &d, size_t buf_len, uint32_t n_slots = 8) : buf_len_(buf_len) {
        check(sdr_demod_ring_open(d.raw(), buf_len, n_slots, &h_));
        out_cap_ = buf_len / 2 / d.config.downsample + 16;
    }
    ~Ring() { close(); }
    Ring(const Ring &) = delete;
    Ring &operator=(const Ring &) = delete;
    uint8_t *acquire() {   // blocks while every slot is in flight
        uint8_t *p = nullptr;
        check(sdr_ring_acquire(h_, &p));
        return p;
    }
    void commit() { check(sdr_ring_commit(h_)); }
    size_t collect(std::vector<int16_t> &out) {   // audio of the oldest committed buffer
        out.resize(out_cap_);
        long n = check(sdr_ring_collect(h_, out.data(), out.size()));
        out.resize((size_t)n);
        return (size_t)n;
    }
    void close() {
        if (h_) sdr_ring_close(h_);
        h_ = nullptr;
    }
    size_t buf_len() const { return buf_len_; }

private:
    sdr_ring *h_ = nullptr;
    size_t buf_len_ = 0, out_cap_ = 0;
};

// Optional audio post-stages after low_pass_real (sdr_post_*): all off by default.
class AudioPost {
public:
    explicit AudioPost(const sdr_post_config &cfg, int cuda_device = 0) { check(sdr_post_new(&cfg, cuda_device, &h_)); }
    ~AudioPost() { sdr_post_free(h_); }
    AudioPost(const AudioPost &) = delete;
    AudioPost &operator=(const AudioPost &) = delete;
    void process(std::vector<int16_t> &audio, const uint8_t *raw = nullptr, size_t raw_len = 0) {
        check(sdr_post_process(h_, audio.data(), audio.size(), raw, raw_len));
    }

private:
    sdr_post *h_ = nullptr;
};

// Buffer source with the read_sync contract of RtlSdr (src/lib.rs:153-155)
class Source {
public:
    static Source open_file(const std::string &path, bool loop = false) {
        Source s;
        check(sdr_source_open_file(path.c_str(), loop ? 1 : 0, &s.h_));
        return s;
    }
    static Source open_synth(uint64_t seed, uint64_t total_bytes = 0) {
        Source s;
        check(sdr_source_open_synth(seed, total_bytes, &s.h_));
        return s;
    }
    Source(Source &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    ~Source() { sdr_source_close(h_); }
    // read_sync(&self, buf: &mut [u8]) -> Result<usize>
    size_t read_sync(uint8_t *buf, size_t len) { return (size_t)check(sdr_source_read_sync(h_, buf, len)); }
    sdr_source *raw() { return h_; }

private:
    Source() = default;
    sdr_source *h_ = nullptr;
};

}  // namespace sdr
