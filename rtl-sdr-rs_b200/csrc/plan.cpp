// plan.cpp — closed-form stream bookkeeping (pure host arithmetic; needs no CUDA device).
//
// These are the same formulas the kernels use; exposing them lets callers size
// output buffers up front and lets a multi-rank job cut one stream into per-rank time slices that
// tile it exactly.  The reference has no equivalent: its loops discover the counts by running
// (examples/simple_fm.rs:337-352, :408-426).
#include <cstdint>
#include <cstring>

#include "../../include/sdr_b200.h"

namespace sdr {
int fail(int code, const char *fmt, ...);
}
using sdr::fail;

extern "C" {

int sdr_fmrx_plan(const sdr_fmrx_config *cfg, uint64_t n_in0, size_t n_samples, uint64_t *y0, size_t *n_y,
                  uint64_t *a0, size_t *n_audio) {
    if (!cfg || cfg->decim < 1) return fail(SDR_E_ARG, "sdr_fmrx_plan: bad config");
    if (cfg->n_taps2 && (cfg->up < 1 || cfg->down < 1)) return fail(SDR_E_ARG, "sdr_fmrx_plan: bad resampler ratio");
    const uint64_t D = cfg->decim;
    const uint64_t yy0 = n_in0 / D;                      // outputs whose last sample (m+1)D-1 < n_in0
    const uint64_t yy1 = (n_in0 + n_samples) / D;
    uint64_t aa0 = yy0, aa1 = yy1;
    if (cfg->n_taps2) {                                  // audio i exists once d[floor(iM/L)] does: count = ceil(P*L/M)
        const uint64_t L = cfg->up, M = cfg->down;
        aa0 = (yy0 * L + M - 1) / M;
        aa1 = (yy1 * L + M - 1) / M;
    }
    if (y0) *y0 = yy0;
    if (n_y) *n_y = (size_t)(yy1 - yy0);
    if (a0) *a0 = aa0;
    if (n_audio) *n_audio = (size_t)(aa1 - aa0);
    return SDR_OK;
}

int sdr_demod_plan(const sdr_demod_config *cfg, const sdr_demod_state *st, size_t buf_len, size_t n_bufs,
                   size_t *n_lowpassed, size_t *n_audio, sdr_demod_state *after) {
    if (!cfg || cfg->downsample < 1 || cfg->rate_resample < 1 || cfg->rate_resample > cfg->rate_out)
        return fail(SDR_E_ARG, "sdr_demod_plan: bad config");
    if (buf_len % 8) return fail(SDR_E_LEN, "buffer length must be a multiple of 8 (examples/simple_fm.rs:284-295)");
    const uint64_t D = cfg->downsample, fast = cfg->rate_out, slow = cfg->rate_resample;
    const uint64_t p0 = st ? st->prev_index : 0, q0 = st ? (uint64_t)st->prev_lpr_index : 0;
    if (p0 >= D || q0 >= fast) return fail(SDR_E_ARG, "sdr_demod_plan: state out of range");
    const uint64_t n = (uint64_t)(buf_len / 2) * n_bufs;
    const uint64_t L = (p0 + n) / D;                                  // low_pass_complex :337-352
    const unsigned __int128 t = (unsigned __int128)L * slow + q0;     // low_pass_real :408-426
    const uint64_t E = (uint64_t)(t / fast);
    if (n_lowpassed) *n_lowpassed = (size_t)L;
    if (n_audio) *n_audio = (size_t)E;
    if (after) {
        if (st) *after = *st; else memset(after, 0, sizeof(*after));
        after->prev_index = (p0 + n) % D;
        after->prev_lpr_index = (int32_t)(uint64_t)(t % fast);
        // lp_now / demod_pre / now_lpr are data dependent: not predicted here
    }
    return SDR_OK;
}

int sdr_shard_range(uint64_t total, uint32_t world, uint32_t rank, uint64_t align, uint64_t *lo, uint64_t *hi) {
    if (!lo || !hi || world < 1 || rank >= world) return fail(SDR_E_ARG, "sdr_shard_range: bad argument");
    if (align < 1) align = 1;
    const uint64_t units = total / align;                 // whole aligned units; the remainder goes to the last rank
    const uint64_t base = units / world, extra = units % world;
    const uint64_t u_lo = (uint64_t)rank * base + (rank < extra ? rank : extra);
    const uint64_t u_hi = u_lo + base + (rank < extra ? 1 : 0);
    *lo = u_lo * align;
    *hi = (rank == world - 1) ? total : u_hi * align;
    return SDR_OK;
}

}  // extern "C"
