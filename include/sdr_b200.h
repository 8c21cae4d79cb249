/*
 * sdr_b200.h — C ABI of the B200-native IQ-sample DSP hot path.
 *
 * Drop-in boundary for the `Demod` pipeline of ccostes/rtl-sdr-rs's examples/simple_fm.rs and
 * the `RtlSdr::read_sync` buffer surface that feeds it.  Every entry point names the reference
 * interface it replaces (file:line relative to the reference root).  Plain pointers and sizes
 * only: no C++/torch types cross this boundary.  All compute runs in hand-written CUDA kernels
 * for sm_100a; there is NO CPU fallback — without a CUDA device every compute entry point
 * returns SDR_E_CUDA.
 *
 * Conventions (reference: `Result<T, RtlsdrError>`, src/error.rs:40-53; panics in the example):
 *   - `int`/`long` returns: >= 0 success (element counts where stated), < 0 an SDR_E_* code;
 *     sdr_last_error() returns a thread-local message for the last failure.
 *   - A handle is single-caller-at-a-time ("Send, not Sync", like `&mut self`); distinct handles
 *     are independent (own CUDA stream).
 *   - Host buffers are caller-owned.  *_dev variants take device pointers obtained from
 *     sdr_dev_alloc() (16-byte aligned, with head/tail room the kernels may read) and are
 *     asynchronous on the handle's stream until sdr_*_sync().
 *   - Complex<i32> is {re:i32, im:i32} => interleaved int32_t pairs; complex f32 => float pairs.
 */
#ifndef SDR_B200_H
#define SDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDR_B200_ABI_VERSION 1

/* ---- errors ------------------------------------------------------------------------------ */
enum {
    SDR_OK = 0,
    SDR_E_ARG = -1,   /* null pointer / invalid parameter */
    SDR_E_LEN = -2,   /* length the reference would panic on: len % 8 != 0 (rotate_90 indexes
                         i+7, examples/simple_fm.rs:284-295) or < 2 lowpassed samples (:356) */
    SDR_E_CAP = -3,   /* output capacity too small (never truncates) */
    SDR_E_CUDA = -4,  /* CUDA runtime error / no device */
    SDR_E_NCCL = -5,  /* NCCL error / library not loadable */
    SDR_E_IO = -6,    /* source I/O error (reference: RtlsdrError::Usb, src/error.rs:42) */
    SDR_E_STATE = -7  /* call not valid in the handle's current state */
};
const char *sdr_last_error(void);
int sdr_abi_version(void);
/* Number of visible CUDA devices (0 if none); never fails. */
int sdr_device_count(void);
/* Fills name (<= cap bytes), SM count and global-memory bytes of a device. */
int sdr_device_info(int device, char *name, size_t cap, int *sm_count, uint64_t *mem_bytes);
/* Total kernels launched by this library in this process (bench.py's gpu_launches claim). */
uint64_t sdr_kernel_launch_count(void);

/* ---- device / pinned memory ---------------------------------------------------------------- */
void *sdr_dev_alloc(int device, size_t bytes);          /* 256-B aligned, 4 KiB head+tail room */
void sdr_dev_free(int device, void *p);
void *sdr_host_alloc(size_t bytes);                     /* pinned host memory */
void sdr_host_free(void *p);
/* Page-lock / release a buffer the CALLER owns (the reader's long-lived `Box<[u8; DEFAULT_BUF_LENGTH]>`,
 * examples/simple_fm.rs:114, or a batch Vec<u8>): calls given a pointer inside a registered buffer copy by DMA straight
 * from it, as from sdr_host_alloc memory, instead of through the pageable staging.  Registering costs about as much as
 * touching every page once; do it once per buffer, not per call, and unregister before freeing the memory.
 * Registering twice / unregistering an unknown pointer is SDR_OK.  SDR_E_STATE while a persistent ring is open. */
int sdr_host_register(void *p, size_t bytes);
int sdr_host_unregister(void *p);
int sdr_memcpy_h2d(int device, void *dst, const void *src, size_t bytes);
int sdr_memcpy_d2h(int device, void *dst, const void *src, size_t bytes);
int sdr_dev_memset(int device, void *dst, int value, size_t bytes);
/* Counter-based synthetic IQ on the device: byte i = mix64(seed, i>>3) >> 8*(i&7) (uniform). */
int sdr_synth_fill_dev(int device, uint8_t *d_buf, size_t bytes, uint64_t seed, uint64_t byte_offset);
int sdr_device_sync(int device);

/* ============================================================================================
 * 1. Reference-exact integer path: `Demod` (examples/simple_fm.rs:232-427)
 * ========================================================================================== */

/* DemodConfig, examples/simple_fm.rs:179-185, field for field. */
typedef struct {
    uint32_t rate_in, rate_out, rate_resample, downsample, output_scale;
} sdr_demod_config;

/* RadioConfig, examples/simple_fm.rs:173-176. */
typedef struct {
    uint32_t capture_freq, capture_rate;
} sdr_radio_config;

/* The five carried fields of struct Demod, examples/simple_fm.rs:234-238. */
typedef struct {
    uint64_t prev_index;
    int32_t now_lpr, prev_lpr_index;
    int32_t lp_now_re, lp_now_im;
    int32_t demod_pre_re, demod_pre_im;
} sdr_demod_state;

typedef struct sdr_demod sdr_demod;

/* optimal_settings(freq, rate), examples/simple_fm.rs:189-214 (host arithmetic, no device).
 * sample_rate / rate_resample play the SAMPLE_RATE / RATE_RESAMPLE constants (:26-27). */
int sdr_optimal_settings(uint32_t freq, uint32_t rate, uint32_t sample_rate, uint32_t rate_resample,
                         sdr_radio_config *radio, sdr_demod_config *demod);

/* Demod::new, :243-252.  Requires 1 <= downsample, 1 <= rate_resample <= rate_out < 2^31. */
int sdr_demod_new(const sdr_demod_config *cfg, int cuda_device, sdr_demod **out);
void sdr_demod_free(sdr_demod *d);
int sdr_demod_get_state(const sdr_demod *d, sdr_demod_state *st);
int sdr_demod_set_state(sdr_demod *d, const sdr_demod_state *st);
/* Audio samples one demodulate(len) call will return from the current state (closed form). */
long sdr_demod_out_len(const sdr_demod *d, size_t len);

/* Demod::demodulate(&mut self, Vec<u8>) -> Vec<i16>, :256-269.  One reference call: the first
 * discriminator sample uses the f64 atan2 against demod_pre (:359), state is carried.  Fused
 * kernel: rotate_90 + (-127) + boxcar + discriminator + resampler in one launch. */
long sdr_demod_demodulate(sdr_demod *d, const uint8_t *buf, size_t len, int16_t *out, size_t out_cap);
/* n_bufs consecutive demodulate() calls of buf_len bytes each (the reader->processor stream of
 * examples/simple_fm.rs:108-128,145-160) in one pipelined submission.  Bit-identical to calling
 * sdr_demod_demodulate n_bufs times.  out_lens (optional, n_bufs entries) receives the per-call
 * audio counts.  Returns total audio samples. */
long sdr_demod_demodulate_batch(sdr_demod *d, const uint8_t *buf, size_t buf_len, size_t n_bufs,
                                int16_t *out, size_t out_cap, uint32_t *out_lens);
/* Same, device-resident input and output (asynchronous; sdr_demod_sync() to wait). */
long sdr_demod_demodulate_batch_dev(sdr_demod *d, const uint8_t *d_buf, size_t buf_len, size_t n_bufs,
                                    int16_t *d_out, size_t out_cap);
int sdr_demod_sync(sdr_demod *d);
/* Device time (ms, CUDA events on the handle's stream) of the kernels of the last *_dev or
 * batch submission, and the number of kernels it launched. */
int sdr_demod_last_timing(const sdr_demod *d, float *kernel_ms, uint32_t *n_launches);
/* Device-time span on the handle's stream (CUDA events): begin records an event, end records a
 * second one, waits for it and returns the elapsed milliseconds between the two. */
int sdr_demod_span_begin(sdr_demod *d);
int sdr_demod_span_end(sdr_demod *d, float *ms);

/* Persistent ring: successive USB-sized buffers stream through ONE resident kernel — no relaunch per buffer.
 * Mirrors the reader -> channel -> processor pair of examples/simple_fm.rs:55-60,108-128,145-160: the
 * producer thread acquires the next pinned slot (blocks while all n_slots are in flight), fills it (read_sync
 * straight into it) and commits it; the consumer thread collects the audio of the oldest buffer.  Output is
 * bit-identical to one sdr_demod_demodulate(buf_len) call per buffer.  While a ring is open the Demod handle
 * is owned by it (its other entry points return SDR_E_STATE); sdr_ring_close retires the kernel and hands the
 * carried state back to the handle, and sdr_demod_free closes a ring that is still open.  Other handles on the
 * same GPU keep working while a ring is resident: the library never waits for the whole device then — memory it
 * frees is parked until the last ring on that device closes, and sdr_device_sync() returns SDR_E_STATE. */
typedef struct sdr_ring sdr_ring;
int sdr_demod_ring_open(sdr_demod *d, size_t buf_len, uint32_t n_slots /* 2..64 */, sdr_ring **out);
int sdr_ring_acquire(sdr_ring *r, uint8_t **buf);                 /* producer */
int sdr_ring_commit(sdr_ring *r);                                 /* producer: H2D copy + doorbell */
long sdr_ring_collect(sdr_ring *r, int16_t *out, size_t cap);     /* consumer: audio of the oldest buffer */
int sdr_ring_close(sdr_ring *r);

/* Stage entry points (stage-level parity against the reference's three known-answer tests). */
/* Demod::rotate_90(Vec<u8>) -> Vec<u8>, scalar branch :276-299; in place, len % 8 == 0. */
long sdr_rotate_90(sdr_demod *d, uint8_t *buf, size_t len);
/* `buf.iter().map(|v| *v as i16 - 127)` :258 followed by buf_to_complex :441-450:
 * u8[len] -> Complex<i32>[len/2]. */
long sdr_buf_to_complex(sdr_demod *d, const uint8_t *buf, size_t len, int32_t *out_pairs, size_t cap_pairs);
/* Demod::low_pass_complex, :337-352 (state: prev_index, lp_now). */
long sdr_low_pass_complex(sdr_demod *d, const int32_t *iq_pairs, size_t n, int32_t *out_pairs, size_t cap_pairs);
/* Demod::fm_demod, :355-367 (state: demod_pre). n < 2 => SDR_E_LEN. */
long sdr_fm_demod(sdr_demod *d, const int32_t *iq_pairs, size_t n, int16_t *out, size_t cap);
/* Demod::low_pass_real, :408-426 (state: now_lpr, prev_lpr_index). */
long sdr_low_pass_real(sdr_demod *d, const int16_t *in, size_t n, int16_t *out, size_t cap);
/* Demod::fast_atan2(y, x), :383-405, vectorised: out[i] = fast_atan2(y[i], x[i]). */
long sdr_fast_atan2(sdr_demod *d, const int32_t *y, const int32_t *x, size_t n, int32_t *out);
/* Demod::polar_discriminant (:370-374, fast=0) / polar_discriminant_fast (:377-380, fast=1):
 * out[i] = f(a[i], b[i]) for Complex<i32> pairs. */
long sdr_polar_discriminant(sdr_demod *d, const int32_t *a_pairs, const int32_t *b_pairs, size_t n,
                            int fast, int32_t *out);

/* Optional audio post-stages after low_pass_real, ALL OFF BY DEFAULT: the reference computes
 * `output_scale` (examples/simple_fm.rs:184,197-200) and never uses it, and has no de-emphasis, DC block or squelch — its
 * output is the raw low_pass_real stream and so is the golden hash.  The stages restate what rtl_fm (the program the example
 * was ported from) runs after its own low_pass_real, in its order and integer arithmetic (csrc/post.cu; "parity unpinned").
 * One sdr_post_process() call = one block = the audio of one demodulate() call; state (de-emphasis and DC averages) is carried. */
typedef struct {
    uint32_t output_scale;    /* 0 or 1 = off; audio * scale, saturated to i16 */
    uint32_t squelch_level;   /* 0 = off; a block is zeroed when rms((raw - 127.5) * 16) of its raw bytes is below the level */
    uint32_t deemph_a;        /* 0 = off; rtl_fm deemph_filter constant, see sdr_post_deemph_a() */
    uint32_t dc_block;        /* 0 = off, 1 = rtl_fm dc_block_filter */
} sdr_post_config;
typedef struct sdr_post sdr_post;
/* round(1 / (1 - exp(-1 / (rate * tau)))): rate = audio rate in Hz, tau in microseconds (75 in the Americas, 50 elsewhere). */
uint32_t sdr_post_deemph_a(uint32_t rate, double tau_us);
int sdr_post_new(const sdr_post_config *cfg, int cuda_device, sdr_post **out);
void sdr_post_free(sdr_post *p);
/* In place on the caller's audio block (host memory).  raw / raw_len: the raw IQ bytes the block was demodulated from (only
 * read when the squelch is on; may be NULL).  Returns n. */
long sdr_post_process(sdr_post *p, int16_t *audio, size_t n, const uint8_t *raw, size_t raw_len);

/* ============================================================================================
 * 2. f32 tap'd-FIR receiver (BASELINE.json configs 2-3; extension — no reference implementation,
 *    parity against the f64 oracle).  Stage names follow north_star: low_pass / fm_demod / resample.
 *      low_pass : y[m] = sum_{k<T} h[k] * (x[(m+1)D-1-k] - 127),  x[n<0] = 127   (complex f32)
 *      fm_demod : d[m] = gain * atan2(Im(y[m] conj y[m-1]), Re(..)),  y[-1] = 0   (f32)
 *      resample : a[i] = sum_p g[iM - pL] * d[p]   (rational L/M polyphase FIR, f32)
 *    T=D, h=1 reduces low_pass to the reference boxcar (:337-352) without rotate_90.
 * ========================================================================================== */
typedef struct {
    uint32_t n_taps;     /* T >= 1 */
    uint32_t decim;      /* D >= 1 */
    uint32_t n_taps2;    /* T2 (0 = no resample stage: audio = discriminator output) */
    uint32_t up, down;   /* L, M of the resampler */
    float gain;          /* discriminator gain; 0 => 16384/pi (the reference's i16 scale) */
} sdr_fmrx_config;

typedef struct sdr_fmrx sdr_fmrx;

int sdr_fmrx_new(const sdr_fmrx_config *cfg, const float *taps, const float *taps2, int cuda_device,
                 sdr_fmrx **out);
void sdr_fmrx_free(sdr_fmrx *r);
int sdr_fmrx_reset(sdr_fmrx *r);
/* Counts a process() call of n_samples will produce from the current state. */
int sdr_fmrx_out_lens(const sdr_fmrx *r, size_t n_samples, size_t *n_y, size_t *n_audio);
/* Whole chain, host buffers, streaming state carried (history, y[m-1], discriminator tail).
 * iq: 2*n_samples bytes.  y_pairs / demod may be NULL (then they are never written to HBM:
 * the fused convert+FIR+demod kernel keeps y on chip).  Returns audio samples written. */
long sdr_fmrx_process(sdr_fmrx *r, const uint8_t *iq, size_t n_samples, float *y_pairs, size_t y_cap,
                      float *demod, size_t demod_cap, float *audio, size_t audio_cap);
/* Device-resident variant (asynchronous). */
long sdr_fmrx_process_dev(sdr_fmrx *r, const uint8_t *d_iq, size_t n_samples, float *d_y_pairs,
                          float *d_demod, float *d_audio, size_t audio_cap);
/* Stage entry points (north_star names), host buffers, state carried per stage. */
long sdr_fmrx_low_pass(sdr_fmrx *r, const uint8_t *iq, size_t n_samples, float *y_pairs, size_t cap_pairs);
long sdr_fmrx_fm_demod(sdr_fmrx *r, const float *y_pairs, size_t n, float *out, size_t cap);
long sdr_fmrx_resample(sdr_fmrx *r, const float *d, size_t n, float *out, size_t cap);
int sdr_fmrx_sync(sdr_fmrx *r);
/* Per-kernel device time of the last process*(): [0] fused convert+FIR(+demod), [1] resampler,
 * [2] everything else; and launches.  Which kernel variant ran: 1 = specialised, 0 = generic. */
int sdr_fmrx_last_timing(const sdr_fmrx *r, float ms[3], uint32_t *n_launches, int *specialised);
/* Sums of the per-kernel device times ([0] fused FIR, [1] resampler, [2] rest) over all timed
 * process*() calls since the last reset, harvested from a ring of CUDA events (no per-call sync). */
int sdr_fmrx_timing_totals(sdr_fmrx *r, double sums_ms[3], uint64_t *n_calls, int reset);
/* Which convert+FIR(+demod) kernel the handle runs: 0 = generic (one warp per output), 1 = pre-compiled
 * k_fir_fast (the BASELINE.json shapes), 2 = k_fir_fast compiled at sdr_fmrx_new() time for this (n_taps, decim)
 * by NVRTC (same source as the pre-compiled instances, bit-identical arithmetic), 3 = k_fir_slide, the output-owner kernel
 * compiled by NVRTC for shapes with more than 16 taps per decimation step (n_taps <= 640, decim <= 64) and for decimations
 * up to 4 with 8 or more (SDR_FIR_SLIDE=0 / 1: never / wherever it exists).  When 0 and a specialised kernel
 * was wanted, *note (optional, valid until the handle is freed) says why it could not be had.
 * SDR_FIR_RTC=0 in the environment disables run-time compilation. */
int sdr_fmrx_kernel_kind(const sdr_fmrx *r, const char **note);
/* Diagnostic, needs no GPU: compile the kernel for (n_taps, decim) with NVRTC exactly as sdr_fmrx_new() would and
 * return the cubin bytes (shape[4] = {blocks per thread, threads per CTA, bytes per load + 256 * row padding, instantiations compiled
 * now rather than taken from the on-disk cache});
 * < 0: SDR_E_ARG (shape outside the kernel's range), SDR_E_STATE (no libnvrtc / compile error, see sdr_last_error()). */
long sdr_rtc_selftest(uint32_t n_taps, uint32_t decim, int shape[4]);
/* The CTA shape sdr_fmrx_new() would compile for (n_taps, decim), without compiling: shape[4] = {blocks per thread,
 * threads per CTA, bytes per load, row padding}.  1 = a shape was picked and satisfies the kernel's compile-time
 * requirements, 0 = outside the specialised kernel's range (the generic kernel runs), < 0 = the picker chose a shape the
 * kernel would reject (a bug; sdr_last_error() says which requirement). */
int sdr_rtc_pick_shape(uint32_t n_taps, uint32_t decim, int shape[4]);
/* Diagnostic / build step, needs no GPU: compile the persistent-ring kernel of a run-time-specialised (n_taps, decim) shape
 * exactly as sdr_fmrx_ring_open() would (it holds every load-phase variant of the FIR tile, so it takes several times as long
 * as one FIR kernel); the cubin lands in the on-disk cache.  Returns its size, or a negative code. */
long sdr_rtc_compile_ring(uint32_t n_taps, uint32_t decim);
/* Same for the output-owner FIR kernel (k_fir_slide: shapes with more than 16 lags per sample, T <= 640, decim <= 64, and the
 * decimations up to 4): shape[2] = {outputs per thread, threads per CTA}.  Returns the cubin size, or a negative code
 * (SDR_E_ARG: the shape is outside that kernel's range and sdr_fmrx_new() would run the generic kernel). */
long sdr_rtc_compile_slide(uint32_t n_taps, uint32_t decim, int shape[2]);
/* Persistent ring for the f32 receiver: the reader -> channel -> processor pair of examples/simple_fm.rs:55-60,108-128,
 * 145-160 with ONE resident kernel behind it (same protocol as sdr_demod_ring_*: the producer acquires a pinned slot, fills
 * it and commits it — one H2D copy plus a 4-byte doorbell, no kernel launch; the consumer collects the audio of the oldest
 * buffer from host-mapped memory).  Output is bit-identical to one sdr_fmrx_process(buf_len / 2 samples) call per buffer.
 * buf_len: bytes per buffer, a multiple of 16, holding at least the filter history and producing at least one audio sample
 * (any USB-sized buffer does).  Needs a specialised FIR kernel (sdr_fmrx_kernel_kind() 1 or 2).  While the ring is open the
 * handle's other stream entry points return SDR_E_STATE; close hands the stream position and history back to the handle. */
typedef struct sdr_fmrx_ring sdr_fmrx_ring;
int sdr_fmrx_ring_open(sdr_fmrx *r, size_t buf_len, uint32_t n_slots /* 2..64 */, sdr_fmrx_ring **out);
int sdr_fmrx_ring_acquire(sdr_fmrx_ring *g, uint8_t **buf);              /* producer */
int sdr_fmrx_ring_commit(sdr_fmrx_ring *g);                              /* producer: H2D copy + doorbell */
long sdr_fmrx_ring_collect(sdr_fmrx_ring *g, float *audio, size_t cap);  /* consumer: audio of the oldest buffer */
int sdr_fmrx_ring_close(sdr_fmrx_ring *g);
int sdr_fmrx_span_begin(sdr_fmrx *r);
int sdr_fmrx_span_end(sdr_fmrx *r, float *ms);
/* Reposition a fresh stream at global sample index n (history = mid-scale): lets a rank that owns
 * the time slice [n, ...) of one stream prime its carry by first processing the samples before n. */
int sdr_fmrx_seek(sdr_fmrx *r, uint64_t global_sample_index);

/* ---- closed-form stream bookkeeping (pure host arithmetic: no device needed) -------------------
 * The reference discovers these counts by running its loops (:337-352, :408-426); here they are
 * formulas, so callers can size buffers up front and a multi-rank job can cut ONE stream into
 * per-rank time slices that tile it exactly. */
/* For a stream positioned at global sample n_in0: index of the first FIR output / audio sample a
 * call of n_samples produces, and how many. */
int sdr_fmrx_plan(const sdr_fmrx_config *cfg, uint64_t n_in0, size_t n_samples, uint64_t *y0, size_t *n_y,
                  uint64_t *a0, size_t *n_audio);
/* Integer Demod: lowpassed / audio counts of n_bufs calls of buf_len bytes from state `st` (NULL =
 * fresh) and the index part of the state afterwards (prev_index, prev_lpr_index). */
int sdr_demod_plan(const sdr_demod_config *cfg, const sdr_demod_state *st, size_t buf_len, size_t n_bufs,
                   size_t *n_lowpassed, size_t *n_audio, sdr_demod_state *after);
/* Contiguous slice [lo, hi) of a stream of `total` samples owned by `rank` of `world`, cut on
 * multiples of `align` samples (the remainder goes to the last rank). */
int sdr_shard_range(uint64_t total, uint32_t world, uint32_t rank, uint64_t align, uint64_t *lo, uint64_t *hi);

/* ============================================================================================
 * 3. Wideband channeliser (configs 4-5): per channel c an NCO mix by a 32-bit phase word
 *    (theta_c(n) = 2*pi*((fw_c*n) mod 2^32)/2^32), the same decimating FIR, the discriminator.
 *    Multi-GPU: channels are partitioned across ranks; the raw u8 slab is broadcast from rank 0
 *    with one ncclBroadcast per slab and nothing else.
 * ========================================================================================== */
typedef struct {
    uint32_t n_channels;
    uint32_t n_taps, decim;
    float gain;
} sdr_chan_config;
typedef struct sdr_chan sdr_chan;

int sdr_chan_new(const sdr_chan_config *cfg, const float *taps, const uint32_t *freq_words,
                 int cuda_device, sdr_chan **out);
void sdr_chan_free(sdr_chan *c);
int sdr_chan_reset(sdr_chan *c);
/* Host buffers.  y_pairs: [C][M][2] or NULL; demod: [C][M].  M = outputs per channel (returned). */
long sdr_chan_process(sdr_chan *c, const uint8_t *iq, size_t n_samples, float *y_pairs, float *demod,
                      size_t cap_per_channel);
long sdr_chan_process_dev(sdr_chan *c, const uint8_t *d_iq, size_t n_samples, float *d_y_pairs,
                          float *d_demod, size_t cap_per_channel);
int sdr_chan_sync(sdr_chan *c);
int sdr_chan_last_timing(const sdr_chan *c, float *kernel_ms, uint32_t *n_launches);
/* Which channeliser kernel the handle runs: 0 = k_chan_fir (taps in shared memory, n_taps > 255), 1 = k_chan_fir_u (direct
 * form, taps as uniform operands), 2 = k_chan_bank (two-stage polyphase bank: the channels sit on a uniform grid
 * fw_c = f0 + c * 2^32/K; info = {K, K1, K2, 64-channel groups}).  SDR_CHAN_BANK=0 in the environment keeps the direct form. */
int sdr_chan_kernel_kind(const sdr_chan *c, uint32_t info[4]);
/* Diagnostic, needs no GPU: the bank plan sdr_chan_new() would pick for (cfg, freq_words) and, if `tables` is given, the
 * coefficient blobs the kernel would receive (7680 floats per 64-channel group, layout in csrc/chan_bank.cuh).  Returns the
 * number of groups, 0 when the channels are not a uniform bank (sdr_last_error() says why), < 0 on bad arguments. */
long sdr_chan_bank_plan(const sdr_chan_config *cfg, const float *taps, const uint32_t *freq_words, uint32_t info[4],
                        float *tables, size_t cap_floats);

/* NCCL plumbing for the slab broadcast (libnccl is dlopen()ed lazily; absent => SDR_E_NCCL). */
#define SDR_NCCL_ID_BYTES 128
typedef struct sdr_comm sdr_comm;
int sdr_comm_unique_id(uint8_t id[SDR_NCCL_ID_BYTES]);                  /* rank 0 */
int sdr_comm_init(int cuda_device, int rank, int world, const uint8_t id[SDR_NCCL_ID_BYTES],
                  sdr_comm **out);
/* ncclBroadcast(d_buf, d_buf, bytes, ncclUint8, root) on the comm's own stream; returns after
 * enqueue.  sdr_comm_wait_on(chan) makes the channeliser's stream wait for the last broadcast. */
int sdr_comm_bcast_u8(sdr_comm *c, uint8_t *d_buf, size_t bytes, int root);
int sdr_comm_chan_wait(sdr_comm *c, sdr_chan *ch);   /* chan stream waits for last bcast */
int sdr_comm_wait_chan(sdr_comm *c, sdr_chan *ch);   /* bcast stream waits for chan's work */
/* Slab-granular form of the same ordering, for a ring of slab buffers that is reused step after step: mark(slot)
 * records "everything enqueued on the channeliser so far" under `slot` (< 32); wait_mark(slot) makes the NEXT
 * broadcast wait for that mark only — not for the slabs channelised since (a never-marked slot waits for nothing). */
int sdr_comm_mark_chan(sdr_comm *c, sdr_chan *ch, uint32_t slot);
int sdr_comm_wait_mark(sdr_comm *c, uint32_t slot);
int sdr_comm_sync(sdr_comm *c);
void sdr_comm_free(sdr_comm *c);

/* ============================================================================================
 * 4. Buffer source surface: RtlSdr::read_sync(&self, buf: &mut [u8]) -> Result<usize>
 *    (src/lib.rs:153-155 -> src/rtlsdr.rs:409-411 -> src/device/mod.rs:141-143).  No USB here:
 *    file and seeded-synthetic sources with the same caller-owned-buffer contract.
 * ========================================================================================== */
#define SDR_DEFAULT_BUF_LENGTH (16 * 16384)   /* DEFAULT_BUF_LENGTH, src/lib.rs:25 */
typedef struct sdr_source sdr_source;
int sdr_source_open_file(const char *path, int loop, sdr_source **out);
int sdr_source_open_synth(uint64_t seed, uint64_t total_bytes /* 0 = endless */, sdr_source **out);
/* rtl_tcp client: a dongle elsewhere feeds this box without USB.  Wire format of examples/rtl_tcp.rs: 12-byte
 * greeting "RTL0" + tuner type + gain count (u32 BE, send_handshake :691-697), then raw u8 IQ; commands are
 * 1 byte id + u32 BE parameter (command_loop :633-689: 0x01 frequency, 0x02 sample rate, ... 0x0e bias tee). */
int sdr_source_open_rtl_tcp(const char *host, uint16_t port, sdr_source **out);
int sdr_source_rtl_tcp_info(const sdr_source *s, uint32_t *tuner_type, uint32_t *gain_count);
int sdr_source_rtl_tcp_command(sdr_source *s, uint8_t cmd, uint32_t param);
/* Blocking; fills buf; returns bytes written (short count at end of data, like a short USB
 * read, examples/simple_fm.rs:121-125) or SDR_E_IO. */
long sdr_source_read_sync(sdr_source *s, uint8_t *buf, size_t len);
/* Extension the reference only has as a TODO (src/lib.rs:147): librtlsdr-style async reads.
 * Spawns the reader thread of examples/simple_fm.rs:89-132, which cycles buf_num buffers of buf_len bytes; cb
 * runs on the CALLING thread (the `process` role, :135-160) for each full buffer; returns when the source ends
 * or is cancelled.  sdr_source_cancel_async may be called from cb or from any other thread; it also wakes a
 * reader blocked on an rtl_tcp socket. */
typedef void (*sdr_read_async_cb)(const uint8_t *buf, size_t len, void *ctx);
int sdr_source_read_async(sdr_source *s, sdr_read_async_cb cb, void *ctx, uint32_t buf_num, uint32_t buf_len);
int sdr_source_cancel_async(sdr_source *s);
void sdr_source_close(sdr_source *s);

#ifdef __cplusplus
}
#endif
#endif /* SDR_B200_H */
