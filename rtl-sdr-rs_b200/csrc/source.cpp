// source.cpp — the buffer-source surface that feeds the hot path.
//
// Replaces (shape only; no USB on a GPU box):
//   RtlSdr::read_sync(&self, buf: &mut [u8]) -> Result<usize>      src/lib.rs:153-155
//     -> Sdr::read_sync  src/rtlsdr.rs:409-411 -> Device::bulk_transfer  src/device/mod.rs:141-143
// and the reader-thread -> channel -> processor hand-off of examples/simple_fm.rs:55-60,108-128,145-156
// (sdr_source_read_async; the reference only has a TODO for an async API, src/lib.rs:147).
#include <arpa/inet.h>
#include <netdb.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/sdr_b200.h"

namespace sdr {
int fail(int code, const char *fmt, ...);
}
using sdr::fail;

struct sdr_source {
    enum Kind { FILE_SRC, SYNTH, RTL_TCP } kind = SYNTH;
    FILE *fp = nullptr;
    int sock = -1;                          // rtl_tcp client socket
    uint32_t tuner_type = 0, gain_count = 0;   // from the 12-byte "RTL0" greeting
    bool loop = false;
    uint64_t seed = 0, total = 0, pos = 0;   // synth: stream byte position
    std::atomic<bool> cancel{false};
    std::atomic<bool> async_active{false};
};

namespace {
inline uint64_t mix64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

long read_once(sdr_source *s, uint8_t *buf, size_t len) {
    if (s->kind == sdr_source::RTL_TCP) {
        size_t got = 0;
        while (got < len) {
            ssize_t r = recv(s->sock, buf + got, len - got, 0);
            if (r < 0) return fail(SDR_E_IO, "rtl_tcp recv failed");
            if (r == 0) break;              // server closed: short count, like a short USB read
            got += (size_t)r;
        }
        return (long)got;
    }
    if (s->kind == sdr_source::FILE_SRC) {
        size_t got = 0;
        while (got < len) {
            size_t r = fread(buf + got, 1, len - got, s->fp);
            got += r;
            if (got == len) break;
            if (ferror(s->fp)) return fail(SDR_E_IO, "read error on file source");
            if (!s->loop || (r == 0 && ftell(s->fp) == 0)) break;   // EOF (or empty file): short count
            rewind(s->fp);
        }
        return (long)got;
    }
    uint64_t avail = s->total ? (s->pos < s->total ? s->total - s->pos : 0) : (uint64_t)len;
    size_t n = (size_t)(avail < len ? avail : len);
    size_t i = 0;
    uint64_t b = s->pos;
    while (i < n && (b & 7)) {
        buf[i++] = (uint8_t)(mix64(s->seed, b >> 3) >> (8 * (b & 7)));
        b++;
    }
    while (i + 8 <= n) {
        uint64_t w = mix64(s->seed, b >> 3);
        memcpy(buf + i, &w, 8);   // little-endian host: byte k of the word is stream byte b+k
        i += 8;
        b += 8;
    }
    while (i < n) {
        buf[i++] = (uint8_t)(mix64(s->seed, b >> 3) >> (8 * (b & 7)));
        b++;
    }
    s->pos = b;
    return (long)n;
}
}  // namespace

extern "C" {

int sdr_source_open_file(const char *path, int loop, sdr_source **out) {
    if (!path || !out) return fail(SDR_E_ARG, "sdr_source_open_file: null argument");
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail(SDR_E_IO, "cannot open %s", path);
    sdr_source *s = new sdr_source();
    s->kind = sdr_source::FILE_SRC;
    s->fp = fp;
    s->loop = loop != 0;
    *out = s;
    return SDR_OK;
}

int sdr_source_open_synth(uint64_t seed, uint64_t total_bytes, sdr_source **out) {
    if (!out) return fail(SDR_E_ARG, "sdr_source_open_synth: null argument");
    sdr_source *s = new sdr_source();
    s->kind = sdr_source::SYNTH;
    s->seed = seed;
    s->total = total_bytes;
    *out = s;
    return SDR_OK;
}

// rtl_tcp wire format (examples/rtl_tcp.rs): the server greets with 12 bytes "RTL0" + tuner type (u32 BE) +
// gain count (u32 BE) (send_handshake, :691-697) and then streams raw interleaved u8 IQ; the client may send
// 5-byte commands: 1 byte id + u32 big-endian parameter (command_loop, :633-689; 0x01 frequency, 0x02 sample
// rate, 0x03 gain mode, 0x04 gain, 0x05 ppm, ... 0x0e bias tee).
int sdr_source_open_rtl_tcp(const char *host, uint16_t port, sdr_source **out) {
    if (!host || !out) return fail(SDR_E_ARG, "sdr_source_open_rtl_tcp: null argument");
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    char portstr[16];
    snprintf(portstr, sizeof(portstr), "%u", (unsigned)port);
    if (getaddrinfo(host, portstr, &hints, &res) != 0 || !res) return fail(SDR_E_IO, "cannot resolve %s", host);
    int fd = -1;
    for (addrinfo *ai = res; ai; ai = ai->ai_next) {
        fd = socket(ai->ai_family, ai->ai_socktype, ai->ai_protocol);
        if (fd < 0) continue;
        if (connect(fd, ai->ai_addr, ai->ai_addrlen) == 0) break;
        close(fd);
        fd = -1;
    }
    freeaddrinfo(res);
    if (fd < 0) return fail(SDR_E_IO, "cannot connect to %s:%u", host, (unsigned)port);
    uint8_t hdr[12];
    size_t got = 0;
    while (got < sizeof(hdr)) {
        ssize_t r = recv(fd, hdr + got, sizeof(hdr) - got, 0);
        if (r <= 0) break;
        got += (size_t)r;
    }
    if (got != sizeof(hdr) || memcmp(hdr, "RTL0", 4) != 0) {
        close(fd);
        return fail(SDR_E_IO, "%s:%u did not send the rtl_tcp \"RTL0\" greeting", host, (unsigned)port);
    }
    sdr_source *s = new sdr_source();
    s->kind = sdr_source::RTL_TCP;
    s->sock = fd;
    s->tuner_type = ((uint32_t)hdr[4] << 24) | ((uint32_t)hdr[5] << 16) | ((uint32_t)hdr[6] << 8) | hdr[7];
    s->gain_count = ((uint32_t)hdr[8] << 24) | ((uint32_t)hdr[9] << 16) | ((uint32_t)hdr[10] << 8) | hdr[11];
    *out = s;
    return SDR_OK;
}

int sdr_source_rtl_tcp_info(const sdr_source *s, uint32_t *tuner_type, uint32_t *gain_count) {
    if (!s || s->kind != sdr_source::RTL_TCP) return fail(SDR_E_STATE, "not an rtl_tcp source");
    if (tuner_type) *tuner_type = s->tuner_type;
    if (gain_count) *gain_count = s->gain_count;
    return SDR_OK;
}

int sdr_source_rtl_tcp_command(sdr_source *s, uint8_t cmd, uint32_t param) {
    if (!s || s->kind != sdr_source::RTL_TCP) return fail(SDR_E_STATE, "not an rtl_tcp source");
    const uint8_t msg[5] = {cmd, (uint8_t)(param >> 24), (uint8_t)(param >> 16), (uint8_t)(param >> 8), (uint8_t)param};
    size_t sent = 0;
    while (sent < sizeof(msg)) {
        ssize_t r = send(s->sock, msg + sent, sizeof(msg) - sent, MSG_NOSIGNAL);
        if (r <= 0) return fail(SDR_E_IO, "rtl_tcp command send failed");
        sent += (size_t)r;
    }
    return SDR_OK;
}

long sdr_source_read_sync(sdr_source *s, uint8_t *buf, size_t len) {
    if (!s || (!buf && len)) return fail(SDR_E_ARG, "sdr_source_read_sync: null argument");
    if (s->async_active.load()) return fail(SDR_E_STATE, "read_sync while read_async is active");
    return read_once(s, buf, len);
}

// Reader thread fills a ring of buf_num buffers and hands full ones to the caller's thread, which
// runs cb (the `process` role, examples/simple_fm.rs:135-160).  A short read ends the stream like
// "Short read, samples lost, exiting!" (:122-125).  Blocks until the source ends or is cancelled.
int sdr_source_read_async(sdr_source *s, sdr_read_async_cb cb, void *ctx, uint32_t buf_num, uint32_t buf_len) {
    if (!s || !cb) return fail(SDR_E_ARG, "sdr_source_read_async: null argument");
    if (buf_num == 0) buf_num = 15;                       // librtlsdr's default
    if (buf_len == 0) buf_len = SDR_DEFAULT_BUF_LENGTH;   // src/lib.rs:25
    if (buf_len % 8) return fail(SDR_E_LEN, "buf_len must be a multiple of 8");
    bool expected = false;
    if (!s->async_active.compare_exchange_strong(expected, true)) return fail(SDR_E_STATE, "read_async already active");
    s->cancel.store(false);

    std::vector<std::vector<uint8_t>> ring(buf_num, std::vector<uint8_t>(buf_len));
    std::mutex mu;
    std::condition_variable cv_full, cv_free;
    std::deque<uint32_t> full;       // indices ready for the consumer
    uint32_t n_free = buf_num;
    bool eof = false;
    long io_err = 0;

    std::thread reader([&] {
        uint32_t idx = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_free.wait(lk, [&] { return n_free > 0 || s->cancel.load(); });
                if (s->cancel.load()) break;
                n_free--;
            }
            long r = read_once(s, ring[idx].data(), buf_len);
            std::unique_lock<std::mutex> lk(mu);
            if (r < 0) io_err = r;
            if (r < (long)buf_len) break;   // error or short read: stop (samples would be lost)
            full.push_back(idx);
            cv_full.notify_one();
            idx = (idx + 1) % buf_num;
        }
        std::unique_lock<std::mutex> lk(mu);
        eof = true;
        cv_full.notify_all();
    });

    for (;;) {
        uint32_t idx;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_full.wait(lk, [&] { return !full.empty() || eof; });
            if (full.empty()) break;
            idx = full.front();
            full.pop_front();
        }
        if (!s->cancel.load()) cb(ring[idx].data(), buf_len, ctx);
        {
            std::unique_lock<std::mutex> lk(mu);
            n_free++;
            cv_free.notify_one();
        }
    }
    {
        std::unique_lock<std::mutex> lk(mu);
        cv_free.notify_all();
    }
    reader.join();
    s->async_active.store(false);
    if (io_err < 0) return (int)io_err;
    return SDR_OK;
}

int sdr_source_cancel_async(sdr_source *s) {
    if (!s) return fail(SDR_E_ARG, "null source");
    s->cancel.store(true);
    // a reader blocked in recv() on an rtl_tcp socket would never look at the flag: shut the read side down, which
    // makes recv() return 0 (end of stream) at once
    if (s->kind == sdr_source::RTL_TCP && s->sock >= 0) shutdown(s->sock, SHUT_RD);
    return SDR_OK;
}

void sdr_source_close(sdr_source *s) {
    if (!s) return;
    if (s->fp) fclose(s->fp);
    if (s->sock >= 0) close(s->sock);
    delete s;
}

}  // extern "C"
