// simple_fm_b200 — the streaming shell of examples/simple_fm.rs around the GPU Demod.
//
//   reader thread   (receive, :89-132): read_sync 262144-byte buffers from the source
//   processor thread(process, :135-170): demodulated audio -> raw s16le on stdout (output, :430-438), timing
//   main            (:36-85): ctrl-c sets SHUTDOWN; both threads poll it
//
//   ./simple_fm_b200 capture.bin | aplay -r 32000 -f S16_LE        (readme.md:13-18 pipes to `play`)
//   ./simple_fm_b200 --synth 1000 > /dev/null                      (1000 seeded synthetic buffers)
//   ./simple_fm_b200 --sync capture.bin                            (one sdr_demod_demodulate() per buffer)
//   ./simple_fm_b200 --deemph 75 --dc-block capture.bin            (optional rtl_fm-style post-stages, off by default;
//                                                                    also --scale N and, with --sync, --squelch LEVEL)
//
// Default mode is the persistent ring (sdr_demod_ring_*): the mpsc channel of :55 IS the ring of pinned slots —
// the reader acquires a slot, read_sync()s straight into it and commits it (one H2D copy + a 4-byte doorbell, no
// kernel launch); the processor collects the audio of the oldest buffer from host-mapped memory.  `--sync` keeps
// the reference's literal shape: a queue of Vec<u8> and one demodulate() call per buffer (:153).  Both are
// bit-identical to the reference's output.  At exit the processor prints the mean processing time like :162-169
// plus the per-buffer latency distribution (commit -> audio in host memory; --sync: the demodulate() call).
//
// Differences from the reference, all deliberate: EOF ends the stream (the reference's file mode has no EOF check
// and re-demodulates stale bytes forever, :72-83).  Logging goes to stderr only — stdout is the audio (:37).
// Exit status: 0 clean end of stream, 1 a library error on either thread, 2 usage.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <csignal>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>

#include "sdr_b200.hpp"

using Clock = std::chrono::steady_clock;

static std::atomic<bool> SHUTDOWN{false};
static std::atomic<int> EXIT_CODE{0};
static void on_sigint(int) { SHUTDOWN.store(true); }

struct Channel {   // mpsc::channel<Vec<u8>> of :55 (the --sync mode keeps the Vec<u8> queue; the ring mode only the counters)
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::vector<uint8_t>> q;
    std::deque<Clock::time_point> sent_at;   // ring mode: commit time of every buffer in flight
    uint64_t sent = 0;
    bool closed = false;
};

static void close_channel(Channel &tx) {
    {
        std::lock_guard<std::mutex> lk(tx.mu);
        tx.closed = true;
    }
    tx.cv.notify_all();
    fprintf(stderr, "Close\n");                                      // :130
}

// true: a full buffer was read into `dst`
static bool read_one(sdr::Source &src, uint8_t *dst) {
    size_t len = 0;
    try {
        len = src.read_sync(dst, sdr::DEFAULT_BUF_LENGTH);           // :116
    } catch (const sdr::Error &e) {
        fprintf(stderr, "Read error: %s\n", e.what());              // :117-120
        EXIT_CODE.store(1);
        return false;
    }
    if (len < sdr::DEFAULT_BUF_LENGTH) {                             // :122-125
        if (len) fprintf(stderr, "Short read (%zu), samples lost, exiting!\n", len);
        return false;
    }
    return true;
}

static void receive_sync(sdr::Source &src, Channel &tx, uint64_t max_bufs) {
    uint64_t n_bufs = 0;
    while (!SHUTDOWN.load() && (max_bufs == 0 || n_bufs < max_bufs)) {
        std::vector<uint8_t> buf(sdr::DEFAULT_BUF_LENGTH);          // alloc_buf(), :114
        if (!read_one(src, buf.data())) break;
        {
            std::lock_guard<std::mutex> lk(tx.mu);
            tx.q.push_back(std::move(buf));                          // tx.send(buf.to_vec()), :127
            tx.sent++;
        }
        tx.cv.notify_one();
        n_bufs++;
    }
    close_channel(tx);
}

static void receive_ring(sdr::Source &src, sdr::Ring &ring, Channel &tx, uint64_t max_bufs) {
    uint64_t n_bufs = 0;
    try {
        while (!SHUTDOWN.load() && (max_bufs == 0 || n_bufs < max_bufs)) {
            uint8_t *slot = ring.acquire();                          // blocks while all slots are in flight
            if (!read_one(src, slot)) break;                         // an acquired, uncommitted slot is simply dropped
            const auto t = Clock::now();
            ring.commit();
            {
                std::lock_guard<std::mutex> lk(tx.mu);
                tx.sent_at.push_back(t);
                tx.sent++;
            }
            tx.cv.notify_one();
            n_bufs++;
        }
    } catch (const sdr::Error &e) {
        fprintf(stderr, "error %d: %s\n", e.code, e.what());
        EXIT_CODE.store(1);
        SHUTDOWN.store(true);
    }
    close_channel(tx);
}

struct Stats {
    std::vector<double> lat_us;
    std::chrono::duration<double> total_time{0};
    uint64_t loop_count = 0, audio_samples = 0;
    void report(const char *mode) {
        if (!loop_count) return;
        fprintf(stderr, "Average processing time: %.4fms (%llu loops, %llu audio samples)\n",   // :162-169
                1e3 * total_time.count() / (double)loop_count, (unsigned long long)loop_count,
                (unsigned long long)audio_samples);
        std::sort(lat_us.begin(), lat_us.end());
        auto pct = [&](double p) { return lat_us[std::min(lat_us.size() - 1, (size_t)(p * (double)lat_us.size()))]; };
        fprintf(stderr, "Per-buffer latency (%s): p50 %.1f us, p90 %.1f us, p99 %.1f us, max %.1f us\n", mode, pct(0.50),
                pct(0.90), pct(0.99), lat_us.back());
    }
};

static sdr::AudioPost *POST = nullptr;   // optional post-stages (null: the reference's output, untouched)

static void emit(const std::vector<int16_t> &audio) {
    fwrite(audio.data(), sizeof(int16_t), audio.size(), stdout);     // output(), :430-438 (s16le on x86)
    fflush(stdout);
}

static void process_sync(sdr::Demod &demod, Channel &rx) {
    Stats st;
    for (;;) {
        std::vector<uint8_t> buf;
        {
            std::unique_lock<std::mutex> lk(rx.mu);
            rx.cv.wait(lk, [&] { return !rx.q.empty() || rx.closed || SHUTDOWN.load(); });
            if (rx.q.empty()) break;
            buf = std::move(rx.q.front());                           // rx.recv(), :150
            rx.q.pop_front();
        }
        const auto t0 = Clock::now();                                // :152
        std::vector<int16_t> result = demod.demodulate(buf);         // :153
        if (POST) POST->process(result, buf.data(), buf.size());
        const std::chrono::duration<double> dt = Clock::now() - t0;
        st.total_time += dt;
        st.lat_us.push_back(dt.count() * 1e6);
        st.loop_count++;
        st.audio_samples += result.size();
        emit(result);
    }
    st.report("one sdr_demod_demodulate() call per buffer");
}

static void process_ring(sdr::Ring &ring, Channel &rx) {
    Stats st;
    std::vector<int16_t> result;
    uint64_t collected = 0;
    const auto t_begin = Clock::now();
    for (;;) {
        Clock::time_point sent_at;
        {
            std::unique_lock<std::mutex> lk(rx.mu);
            rx.cv.wait(lk, [&] { return rx.sent > collected || rx.closed; });
            if (rx.sent == collected) break;                          // closed and drained
            sent_at = rx.sent_at.front();
            rx.sent_at.pop_front();
        }
        ring.collect(result);                                        // spins on the buffer's completion word
        if (POST) POST->process(result);
        st.lat_us.push_back(std::chrono::duration<double>(Clock::now() - sent_at).count() * 1e6);
        collected++;
        st.loop_count++;
        st.audio_samples += result.size();
        emit(result);
    }
    st.total_time = Clock::now() - t_begin;
    st.report("persistent ring: commit -> audio in host memory, buffers pipelined");
}

int main(int argc, char **argv) {
    std::signal(SIGINT, on_sigint);                                  // ctrlc::set_handler, :42-45
    const char *path = nullptr;
    uint64_t synth_bufs = 0;
    int device = 0;
    uint32_t slots = 8;
    bool sync_mode = false;
    sdr_post_config post_cfg{};
    double deemph_us = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--synth") && i + 1 < argc) synth_bufs = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--slots") && i + 1 < argc) slots = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--sync")) sync_mode = true;
        else if (!strcmp(argv[i], "--deemph") && i + 1 < argc) deemph_us = atof(argv[++i]);
        else if (!strcmp(argv[i], "--dc-block")) post_cfg.dc_block = 1;
        else if (!strcmp(argv[i], "--scale") && i + 1 < argc) post_cfg.output_scale = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--squelch") && i + 1 < argc) post_cfg.squelch_level = (uint32_t)atoi(argv[++i]);
        else path = argv[i];
    }
    if (!path && !synth_bufs) {
        fprintf(stderr, "usage: %s <capture.bin> | --synth <n_buffers>  [--sync] [--slots N] [--device N]\n"
                        "       optional post-stages (off by default): [--deemph <us>] [--dc-block] [--scale N] [--squelch LEVEL (--sync only)]\n", argv[0]);
        return 2;
    }
    try {
        auto settings = sdr::optimal_settings(sdr::FREQUENCY, sdr::SAMPLE_RATE);   // :48
        fprintf(stderr, "downsample: %u\nrate_in: %u capture_rate: %u\ncapture_freq: %u\n", settings.second.downsample,
                settings.second.rate_in, settings.first.capture_rate, settings.first.capture_freq);
        fprintf(stderr, "Buffer size: %.2fms\n", 1000.0 * 0.5 * sdr::DEFAULT_BUF_LENGTH / settings.first.capture_rate);   // :101-104
        sdr::Source src = path ? sdr::Source::open_file(path) : sdr::Source::open_synth(0xB2000001ull);
        sdr::Demod demod(settings.second, device);                                  // :137
        fprintf(stderr, "Oversampling input by: %ux\nOutput at %u Hz\nOutput scale: %u\n", demod.config.downsample,
                demod.config.rate_in, demod.config.output_scale);                  // :138-140
        std::unique_ptr<sdr::AudioPost> post;
        if (deemph_us > 0) post_cfg.deemph_a = sdr_post_deemph_a(settings.second.rate_resample, deemph_us);
        if (post_cfg.deemph_a || post_cfg.dc_block || post_cfg.output_scale > 1 || post_cfg.squelch_level) {
            post.reset(new sdr::AudioPost(post_cfg, device));
            POST = post.get();
            fprintf(stderr, "Post-stages: scale %u, squelch %u, de-emphasis a = %u, DC block %u\n", post_cfg.output_scale,
                    post_cfg.squelch_level, post_cfg.deemph_a, post_cfg.dc_block);
        }
        Channel ch;
        auto guarded = [&](auto &&fn) {
            try {
                fn();
            } catch (const sdr::Error &e) {
                fprintf(stderr, "error %d: %s\n", e.code, e.what());
                EXIT_CODE.store(1);
                SHUTDOWN.store(true);
                close_channel(ch);
            }
        };
        if (sync_mode) {
            std::thread receive_thread([&] { receive_sync(src, ch, synth_bufs); });      // :58
            std::thread process_thread([&] { guarded([&] { process_sync(demod, ch); }); });   // :60
            process_thread.join();
            receive_thread.join();
        } else {
            sdr::Ring ring(demod, sdr::DEFAULT_BUF_LENGTH, slots);
            std::thread receive_thread([&] { receive_ring(src, ring, ch, synth_bufs); });
            std::thread process_thread([&] { guarded([&] { process_ring(ring, ch); }); });
            process_thread.join();
            SHUTDOWN.store(true);   // a failed processor must not leave the reader blocked in acquire() for ever
            receive_thread.join();
            ring.close();
        }
    } catch (const sdr::Error &e) {
        fprintf(stderr, "error %d: %s\n", e.code, e.what());
        return 1;
    }
    return EXIT_CODE.load();
}
