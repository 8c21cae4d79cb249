// simple_fm_b200 — the streaming shell of examples/simple_fm.rs around the GPU Demod.
//
//   reader thread   (receive, :89-132): read_sync 262144-byte buffers from the source into a queue
//   processor thread(process, :135-170): drain the queue, demodulate on the GPU, write raw s16le audio to
//                                        stdout (output, :430-438), keep the running mean of the time
//   main            (:36-85): ctrl-c sets SHUTDOWN; both threads poll it
//
//   ./simple_fm_b200 capture.bin | aplay -r 32000 -f S16_LE        (readme.md:13-18 pipes to `play`)
//   ./simple_fm_b200 --synth 1000 > /dev/null                      (1000 seeded synthetic buffers)
//
// Differences from the reference, all deliberate: EOF ends the stream (the reference's file mode has no
// EOF check and re-demodulates stale bytes forever, :72-83); the processor hands ALL queued buffers to
// sdr_demod_demodulate_batch in one submission, which is bit-identical to one demodulate() per buffer.
// Logging goes to stderr only — stdout is the audio (:37).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <csignal>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>

#include "sdr_b200.hpp"

static std::atomic<bool> SHUTDOWN{false};
static void on_sigint(int) { SHUTDOWN.store(true); }

struct Channel {   // mpsc::channel<Vec<u8>> of :55
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::vector<uint8_t>> q;
    bool closed = false;
};

static void receive(sdr::Source &src, Channel &tx, uint64_t max_bufs) {
    uint64_t n_bufs = 0;
    while (!SHUTDOWN.load() && (max_bufs == 0 || n_bufs < max_bufs)) {
        std::vector<uint8_t> buf(sdr::DEFAULT_BUF_LENGTH);          // alloc_buf(), :114
        size_t len = 0;
        try {
            len = src.read_sync(buf.data(), buf.size());             // :116
        } catch (const sdr::Error &e) {
            fprintf(stderr, "Read error: %s\n", e.what());          // :117-120
            break;
        }
        if (len < sdr::DEFAULT_BUF_LENGTH) {                         // :122-125
            if (len) fprintf(stderr, "Short read (%zu), samples lost, exiting!\n", len);
            break;
        }
        {
            std::lock_guard<std::mutex> lk(tx.mu);
            tx.q.push_back(std::move(buf));                          // tx.send(buf.to_vec()), :127
        }
        tx.cv.notify_one();
        n_bufs++;
    }
    {
        std::lock_guard<std::mutex> lk(tx.mu);
        tx.closed = true;
    }
    tx.cv.notify_all();
    fprintf(stderr, "Close\n");                                      // :130
}

static void process(const sdr::DemodConfig &cfg, Channel &rx, int device) {
    sdr::Demod demod(cfg, device);                                   // :137
    fprintf(stderr, "Oversampling input by: %ux\nOutput at %u Hz\nOutput scale: %u\n", demod.config.downsample,
            demod.config.rate_in, demod.config.output_scale);        // :138-140
    std::chrono::duration<double> total_time{0};
    uint64_t loop_count = 0, audio_samples = 0;
    std::vector<uint8_t> batch;
    for (;;) {
        std::deque<std::vector<uint8_t>> got;
        {
            std::unique_lock<std::mutex> lk(rx.mu);
            rx.cv.wait(lk, [&] { return !rx.q.empty() || rx.closed || SHUTDOWN.load(); });
            if (rx.q.empty() && (rx.closed || SHUTDOWN.load())) break;
            got.swap(rx.q);                                          // drain everything that is waiting
        }
        batch.resize(got.size() * sdr::DEFAULT_BUF_LENGTH);
        for (size_t i = 0; i < got.size(); i++) memcpy(batch.data() + i * sdr::DEFAULT_BUF_LENGTH, got[i].data(), got[i].size());
        auto t0 = std::chrono::steady_clock::now();                  // :152
        std::vector<int16_t> result = demod.demodulate_batch(batch.data(), sdr::DEFAULT_BUF_LENGTH, got.size());   // :153
        total_time += std::chrono::steady_clock::now() - t0;
        loop_count += got.size();
        audio_samples += result.size();
        fwrite(result.data(), sizeof(int16_t), result.size(), stdout);   // output(), :430-438 (s16le on x86)
        fflush(stdout);
    }
    if (loop_count)                                                  // :162-169
        fprintf(stderr, "Average processing time: %.4fms (%llu loops, %llu audio samples)\n",
                1e3 * total_time.count() / (double)loop_count, (unsigned long long)loop_count,
                (unsigned long long)audio_samples);
}

int main(int argc, char **argv) {
    std::signal(SIGINT, on_sigint);                                  // ctrlc::set_handler, :42-45
    const char *path = nullptr;
    uint64_t synth_bufs = 0;
    int device = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--synth") && i + 1 < argc) synth_bufs = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else path = argv[i];
    }
    if (!path && !synth_bufs) {
        fprintf(stderr, "usage: %s <capture.bin> | --synth <n_buffers>  [--device N]\n", argv[0]);
        return 2;
    }
    try {
        auto settings = sdr::optimal_settings(sdr::FREQUENCY, sdr::SAMPLE_RATE);   // :48
        fprintf(stderr, "downsample: %u\nrate_in: %u capture_rate: %u\ncapture_freq: %u\n", settings.second.downsample,
                settings.second.rate_in, settings.first.capture_rate, settings.first.capture_freq);
        fprintf(stderr, "Buffer size: %.2fms\n", 1000.0 * 0.5 * sdr::DEFAULT_BUF_LENGTH / settings.first.capture_rate);   // :101-104
        sdr::Source src = path ? sdr::Source::open_file(path) : sdr::Source::open_synth(0xB2000001ull);
        Channel ch;
        std::thread receive_thread([&] { receive(src, ch, synth_bufs); });          // :58
        std::thread process_thread([&] {                                            // :60
            try {
                process(settings.second, ch, device);
            } catch (const sdr::Error &e) {
                fprintf(stderr, "error %d: %s\n", e.code, e.what());
                SHUTDOWN.store(true);
            }
        });
        process_thread.join();
        receive_thread.join();
    } catch (const sdr::Error &e) {
        fprintf(stderr, "error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
