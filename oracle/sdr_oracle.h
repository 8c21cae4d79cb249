/*
 * sdr_oracle.h — CPU ORACLE for the IQ-sample DSP hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is a plain-C restatement of the algorithm in the reference's
 * examples/simple_fm.rs (struct Demod, :232-427) plus an f64 definition of the
 * tap'd-FIR extension path (BASELINE.json configs 2-5, which have no reference
 * implementation).  It exists to CHECK the CUDA product path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link, load or execute anything in this directory; the product library
 * (rtl-sdr-rs_b200/csrc) never does and has no CPU fallback.
 *
 * Pinning status:
 *   - integer path (orc_demod_*): PINNED by the reference's three known-answer tests
 *     (examples/simple_fm.rs:466-555) and by the capture.bin hashes of SURVEY §8c
 *     (tests/test_oracle.py re-derives them).
 *   - f64 FIR / discriminator / resampler / channeliser (orc_fx_*): "parity unpinned" —
 *     the reference contains no tap'd FIR; the definition is this file's.
 *     It is cross-checked against the pinned integer path where the two coincide
 *     (boxcar taps) and against scipy in tests/test_oracle.py.
 *
 * The reference cannot be compiled here (Rust; no rustc/cargo in the image, no network),
 * so there is no oracle/_ref.
 */
#ifndef SDR_ORACLE_H
#define SDR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- integer path: examples/simple_fm.rs ------------------------------------------- */

/* DemodConfig, examples/simple_fm.rs:179-185 (field for field). */
typedef struct {
    uint32_t rate_in, rate_out, rate_resample, downsample, output_scale;
} orc_demod_config;

/* RadioConfig, examples/simple_fm.rs:173-176. */
typedef struct {
    uint32_t capture_freq, capture_rate;
} orc_radio_config;

/* struct Demod, examples/simple_fm.rs:232-239. */
typedef struct {
    orc_demod_config config;
    uint64_t prev_index;       /* usize */
    int32_t now_lpr;
    int32_t prev_lpr_index;
    int32_t lp_now_re, lp_now_im;
    int32_t demod_pre_re, demod_pre_im;
} orc_demod;

/* optimal_settings, :189-214.  sample_rate_const is the SAMPLE_RATE constant (:26) the
 * reference writes into rate_in/rate_out regardless of `rate`; resample is RATE_RESAMPLE. */
void orc_optimal_settings(uint32_t freq, uint32_t rate, uint32_t sample_rate_const,
                          uint32_t rate_resample, orc_radio_config *radio, orc_demod_config *cfg);

void orc_demod_init(orc_demod *d, const orc_demod_config *cfg);               /* :243-252 */
void orc_rotate_90(uint8_t *buf, size_t len);                                  /* :276-299 scalar */
void orc_centre(const uint8_t *buf, size_t len, int16_t *out);                 /* :258 */
size_t orc_buf_to_complex(const int16_t *buf, size_t len, int32_t *out_pairs); /* :441-450 */
size_t orc_low_pass_complex(orc_demod *d, const int32_t *in_pairs, size_t n,
                            int32_t *out_pairs);                               /* :337-352 */
int32_t orc_fast_atan2(int32_t y, int32_t x);                                  /* :383-405 */
int32_t orc_polar_discriminant(int32_t are, int32_t aim, int32_t bre, int32_t bim);      /* :370-374 */
int32_t orc_polar_discriminant_fast(int32_t are, int32_t aim, int32_t bre, int32_t bim); /* :377-380 */
/* returns n, or -1 if n < 2 (the reference asserts, :356) */
long orc_fm_demod(orc_demod *d, const int32_t *in_pairs, size_t n, int16_t *out);        /* :355-367 */
size_t orc_low_pass_real(orc_demod *d, const int16_t *in, size_t n, int16_t *out);       /* :408-426 */
/* demodulate, :256-269.  len must be a multiple of 8 (rotate_90 indexes i+7).  Returns
 * audio samples written, or -1 on the reference's panics (len%8, <2 lowpassed).  If
 * lp_out / dm_out are non-NULL the intermediate streams are also stored (capacity len/2
 * pairs and len/2 samples) and *n_lp receives their count. */
long orc_demodulate(orc_demod *d, const uint8_t *buf, size_t len, int16_t *out,
                    int32_t *lp_out, int16_t *dm_out, size_t *n_lp);
/* Same arithmetic, but structured like the reference (one fresh heap vector per stage,
 * :256-269) — this is the "reference CPU path" leg that bench.py times. */
long orc_demodulate_ref_like(orc_demod *d, const uint8_t *buf, size_t len, int16_t *out);
/* Same call as ONE fused pass with no intermediate vectors (BASELINE.md §3 "oracle_fused" timing leg);
 * bit-identical to orc_demodulate. */
long orc_demodulate_fused(orc_demod *d, const uint8_t *buf, size_t len, int16_t *out);
/* n_bufs consecutive demodulate() calls of buf_len bytes each, `threads` pthreads
 * each running an INDEPENDENT Demod over its own slice of buffers (timing leg only: the
 * output differs from a single sequential Demod at the slice seams). Returns total audio. */
long orc_demodulate_many_mt(const orc_demod_config *cfg, const uint8_t *buf, size_t buf_len,
                            size_t n_bufs, int16_t *out, size_t out_cap, int threads);

/* fused != 0: every thread runs orc_demodulate_fused instead of orc_demodulate_ref_like. */
long orc_demodulate_many_mt2(const orc_demod_config *cfg, const uint8_t *buf, size_t buf_len,
                             size_t n_bufs, int16_t *out, size_t out_cap, int threads, int fused);

/* ---- optional audio post-stages (no reference implementation: rtl_fm's deemph_filter / dc_block_filter
 * restated, output_scale and the raw-byte squelch as defined in include/sdr_b200.h) ------------------------------------ */
typedef struct {
    uint32_t output_scale, squelch_level, deemph_a, dc_block;
    int32_t deemph_avg, dc_avg;
} orc_post;
void orc_post_init(orc_post *p, uint32_t output_scale, uint32_t squelch_level, uint32_t deemph_a, uint32_t dc_block);
void orc_post_process(orc_post *p, int16_t *audio, size_t n, const uint8_t *raw, size_t raw_len);

/* ---- f64 extension path ("parity unpinned" by the reference) ------------ */

/* Streaming state of the f64 chain: raw-byte history for the FIR, last FIR output for the
 * discriminator, discriminator history for the resampler.  All counters are global. */
typedef struct {
    uint32_t n_taps, decim;          /* T, D */
    uint32_t up, down, n_taps2;      /* resampler L, M, T2 */
    double gain;                     /* discriminator gain (16384/pi by default) */
    uint64_t n_in;                   /* complex samples consumed so far */
    uint64_t n_y;                    /* FIR outputs produced so far */
    uint64_t n_a;                    /* audio outputs produced so far */
    double *taps;                    /* T (copied from f32) */
    double *taps2;                   /* T2 */
    double *hist_re, *hist_im;       /* last T-1 centred samples */
    double prev_re, prev_im;         /* y[m-1] */
    double *dhist;                   /* last (T2-1)/L + 1 discriminator outputs (ring as array) */
    size_t dhist_len;
} orc_fx;

int orc_fx_init(orc_fx *s, const float *taps, uint32_t n_taps, uint32_t decim,
                const float *taps2, uint32_t n_taps2, uint32_t up, uint32_t down, double gain);
void orc_fx_free(orc_fx *s);
/* x1: y[m] = sum_k h[k]*(x[(m+1)D-1-k]-127), x[n<0]=127.  out_pairs: (re,im) doubles.
 * Returns outputs written (those m whose last sample lies in this call). */
size_t orc_fx_low_pass(orc_fx *s, const uint8_t *iq, size_t n_samples, double *out_pairs);
/* x2: d[m] = gain*atan2(Im(y[m]conj(y[m-1])), Re(..)), y[-1]=0. */
size_t orc_fx_fm_demod(orc_fx *s, const double *y_pairs, size_t n, double *out);
/* x3: rational L/M polyphase FIR: a[i] = sum_p g[iM-pL]*d[p], emitted when d[floor(iM/L)] exists. */
size_t orc_fx_resample(orc_fx *s, const double *d, size_t n, double *out);
/* Whole chain; any of y_out/d_out may be NULL.  Returns audio outputs written. */
size_t orc_fx_process(orc_fx *s, const uint8_t *iq, size_t n_samples, double *y_out,
                      double *d_out, double *a_out, size_t *n_y_out);

/* x4: channeliser, direct definition (mix by a 32-bit-phase NCO, then FIR/decimate, then
 * discriminator).  Stateless over one block starting at global sample n0 with zero history
 * (the test harness feeds whole streams).  y_out: [C][M][2], d_out: [C][M]. Returns M. */
size_t orc_fx_channelise(const uint8_t *iq, size_t n_samples, const float *taps, uint32_t n_taps,
                         uint32_t decim, const uint32_t *freq_words, uint32_t n_chan, double gain,
                         double *y_out, double *d_out);

/* Optimised single-precision CPU port of the x1+x2+x3 chain, pthreads over output blocks, for
 * the cpu_baseline / --impl reference timing legs (not used for parity).  Zero history.
 * Returns audio outputs written. */
size_t orc_fx_process_f32_mt(const uint8_t *iq, size_t n_samples, const float *taps,
                             uint32_t n_taps, uint32_t decim, const float *taps2, uint32_t n_taps2,
                             uint32_t up, uint32_t down, float gain, float *audio, size_t cap,
                             int threads);

/* Counter-based synthetic IQ generator shared (by definition) with the CUDA library's
 * sdr_synth_fill: byte i of stream `seed` = mix64(seed, i>>3) >> (8*(i&7)).  kind 0 = uniform. */
void orc_synth_fill(uint8_t *buf, size_t len, uint64_t seed, uint64_t byte_offset);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
