/*
 * sdr_oracle.c — CPU ORACLE (test infrastructure only; see sdr_oracle.h for the rules).
 *
 * Integer path: a function-by-function restatement of examples/simple_fm.rs (reference
 * @ 8c32c118).  Rust release-mode integer semantics are reproduced explicitly: `as` casts
 * wrap, i32 arithmetic wraps (done here in uint32_t to stay defined in C), `/` truncates
 * toward zero.  The f64 path is this project's own definition of the tap'd-FIR extension.
 */
#include "sdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* minimal static-schedule parallel-for on pthreads (no OpenMP runtime dependency) */
typedef void (*orc_range_fn)(long lo, long hi, int tid, void *ctx);
typedef struct {
    orc_range_fn fn;
    long lo, hi;
    int tid;
    void *ctx;
} orc_job;
static void *orc_job_main(void *p) {
    orc_job *j = (orc_job *)p;
    j->fn(j->lo, j->hi, j->tid, j->ctx);
    return NULL;
}
static void orc_par_for(int threads, long n, orc_range_fn fn, void *ctx) {
    if (threads < 1) threads = 1;
    if ((long)threads > n) threads = (int)(n > 0 ? n : 1);
    if (threads == 1) {
        fn(0, n, 0, ctx);
        return;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    orc_job *jobs = (orc_job *)malloc(sizeof(orc_job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t].fn = fn;
        jobs[t].lo = n * t / threads;
        jobs[t].hi = n * (t + 1) / threads;
        jobs[t].tid = t;
        jobs[t].ctx = ctx;
        pthread_create(&th[t], NULL, orc_job_main, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* ------------------------------------------------------------------------------------ */
/* helpers: wrapping i32 arithmetic                                                       */
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
/* Rust `/` on i32: truncating; the reference would panic on /0 and MIN/-1.  Those cannot
 * occur for downsample <= 256; the oracle (and the CUDA path) define them as 0 / MIN. */
static inline int32_t tdiv(int32_t a, int32_t b) {
    if (b == 0) return 0;
    if (a == INT32_MIN && b == -1) return INT32_MIN;
    return a / b;
}

/* ------------------------------------------------------------------------------------ */
/* optimal_settings — examples/simple_fm.rs:189-214                                      */
void orc_optimal_settings(uint32_t freq, uint32_t rate, uint32_t sample_rate_const,
                          uint32_t rate_resample, orc_radio_config *radio, orc_demod_config *cfg) {
    uint32_t downsample = (1000000u / rate) + 1u;          /* :190 */
    uint32_t capture_rate = downsample * rate;              /* :192 */
    uint32_t capture_freq = freq + capture_rate / 4u;       /* :195 offset tuning */
    uint32_t output_scale = (1u << 15) / (128u * downsample); /* :197 */
    if (output_scale < 1u) output_scale = 1u;               /* :198-200 */
    if (radio) {
        radio->capture_freq = capture_freq;
        radio->capture_rate = capture_rate;
    }
    if (cfg) {
        cfg->rate_in = sample_rate_const;                   /* :207 (the constant, not `rate`) */
        cfg->rate_out = sample_rate_const;                  /* :208 */
        cfg->rate_resample = rate_resample;                 /* :209 */
        cfg->downsample = downsample;
        cfg->output_scale = output_scale;
    }
}

/* Demod::new — :243-252: everything zero. */
void orc_demod_init(orc_demod *d, const orc_demod_config *cfg) {
    memset(d, 0, sizeof(*d));
    d->config = *cfg;
}

/* Demod::rotate_90, scalar branch — :281-298.  "negation" is 255 - x in the u8 domain. */
void orc_rotate_90(uint8_t *buf, size_t len) {
    for (size_t i = 0; i + 7 < len; i += 8) {
        uint8_t tmp;
        tmp = (uint8_t)(255 - buf[i + 3]);
        buf[i + 3] = buf[i + 2];
        buf[i + 2] = tmp;
        buf[i + 4] = (uint8_t)(255 - buf[i + 4]);
        buf[i + 5] = (uint8_t)(255 - buf[i + 5]);
        tmp = (uint8_t)(255 - buf[i + 6]);
        buf[i + 6] = buf[i + 7];
        buf[i + 7] = tmp;
    }
}

/* `*val as i16 - 127` — :258 */
void orc_centre(const uint8_t *buf, size_t len, int16_t *out) {
    for (size_t i = 0; i < len; i++) out[i] = (int16_t)((int16_t)buf[i] - 127);
}

/* buf_to_complex — :441-450: windows(2).step_by(2); a trailing odd element is dropped. */
size_t orc_buf_to_complex(const int16_t *buf, size_t len, int32_t *out_pairs) {
    size_t n = len / 2;
    for (size_t i = 0; i < n; i++) {
        out_pairs[2 * i] = (int32_t)buf[2 * i];
        out_pairs[2 * i + 1] = (int32_t)buf[2 * i + 1];
    }
    return n;
}

/* Demod::low_pass_complex — :337-352: boxcar sum of `downsample` samples, raw sum out. */
size_t orc_low_pass_complex(orc_demod *d, const int32_t *in_pairs, size_t n, int32_t *out_pairs) {
    size_t w = 0;
    for (size_t orig = 0; orig < n; orig++) {
        d->lp_now_re = wadd(d->lp_now_re, in_pairs[2 * orig]);
        d->lp_now_im = wadd(d->lp_now_im, in_pairs[2 * orig + 1]);
        d->prev_index += 1;
        if (d->prev_index < (uint64_t)d->config.downsample) continue;
        out_pairs[2 * w] = d->lp_now_re;
        out_pairs[2 * w + 1] = d->lp_now_im;
        w++;
        d->lp_now_re = 0;
        d->lp_now_im = 0;
        d->prev_index = 0;
    }
    return w;
}

/* Demod::fast_atan2 — :383-405.  NB the cast binds tighter than the divide: the i64
 * product is wrapped to i32 first, THEN divided (truncating). */
int32_t orc_fast_atan2(int32_t y, int32_t x) {
    const int32_t pi4 = 1 << 12;
    const int32_t pi34 = 3 * (1 << 12);
    if (x == 0 && y == 0) return 0;
    int32_t yabs = y;
    if (yabs < 0) yabs = wsub(0, yabs);
    int32_t angle;
    if (x >= 0) {
        int32_t num = (int32_t)(uint32_t)((int64_t)pi4 * (int64_t)wsub(x, yabs));
        angle = wsub(pi4, tdiv(num, wadd(x, yabs)));
    } else {
        int32_t num = (int32_t)(uint32_t)((int64_t)pi4 * (int64_t)wadd(x, yabs));
        angle = wsub(pi34, tdiv(num, wsub(yabs, x)));
    }
    if (y < 0) return wsub(0, angle);
    return angle;
}

/* a * b.conj() for Complex<i32> (num-complex 0.4 Mul + conj), wrapping. */
static inline void cmul_conj(int32_t are, int32_t aim, int32_t bre, int32_t bim, int32_t *cre,
                             int32_t *cim) {
    *cre = wadd(wmul(are, bre), wmul(aim, bim));
    *cim = wsub(wmul(aim, bre), wmul(are, bim));
}

/* Demod::polar_discriminant — :370-374: real f64 atan2, `as i32` truncates toward zero. */
int32_t orc_polar_discriminant(int32_t are, int32_t aim, int32_t bre, int32_t bim) {
    int32_t cre, cim;
    cmul_conj(are, aim, bre, bim, &cre, &cim);
    double angle = atan2((double)cim, (double)cre);
    double v = angle / 3.14159265358979323846264338327950288 * (double)(1 << 14);
    return (int32_t)v; /* |v| <= 16384: in range, C truncation == Rust `as i32` */
}

/* Demod::polar_discriminant_fast — :377-380 */
int32_t orc_polar_discriminant_fast(int32_t are, int32_t aim, int32_t bre, int32_t bim) {
    int32_t cre, cim;
    cmul_conj(are, aim, bre, bim, &cre, &cim);
    return orc_fast_atan2(cim, cre);
}

/* Demod::fm_demod — :355-367: sample 0 of every call uses the f64 path against demod_pre. */
long orc_fm_demod(orc_demod *d, const int32_t *in, size_t n, int16_t *out) {
    if (n < 2) return -1; /* assert!(buf.len() > 1) */
    int32_t pcm = orc_polar_discriminant(in[0], in[1], d->demod_pre_re, d->demod_pre_im);
    out[0] = (int16_t)(uint16_t)(uint32_t)pcm;
    for (size_t i = 1; i < n; i++) {
        pcm = orc_polar_discriminant_fast(in[2 * i], in[2 * i + 1], in[2 * i - 2], in[2 * i - 1]);
        out[i] = (int16_t)(uint16_t)(uint32_t)pcm;
    }
    d->demod_pre_re = in[2 * (n - 1)];
    d->demod_pre_im = in[2 * (n - 1) + 1];
    return (long)n;
}

/* Demod::low_pass_real — :408-426: fractional boxcar, always divides by fast/slow (u32 div). */
size_t orc_low_pass_real(orc_demod *d, const int16_t *in, size_t n, int16_t *out) {
    uint32_t slow = d->config.rate_resample;
    uint32_t fast = d->config.rate_out;
    int32_t div = (int32_t)(slow ? fast / slow : 0u);
    size_t w = 0;
    for (size_t i = 0; i < n; i++) {
        d->now_lpr = wadd(d->now_lpr, (int32_t)in[i]);
        d->prev_lpr_index = wadd(d->prev_lpr_index, (int32_t)slow);
        if (d->prev_lpr_index < (int32_t)fast) continue;
        out[w++] = (int16_t)(uint16_t)(uint32_t)tdiv(d->now_lpr, div);
        d->prev_lpr_index = wsub(d->prev_lpr_index, (int32_t)fast);
        d->now_lpr = 0;
    }
    return w;
}

/* Demod::demodulate — :256-269, fused loop form (no per-stage vectors). */
long orc_demodulate(orc_demod *d, const uint8_t *buf, size_t len, int16_t *out, int32_t *lp_out,
                    int16_t *dm_out, size_t *n_lp) {
    if (len % 8 != 0) return -1;
    size_t ns = len / 2;
    uint8_t *rot = (uint8_t *)malloc(len ? len : 1);
    int32_t *cx = (int32_t *)malloc((ns ? ns : 1) * 2 * sizeof(int32_t));
    int32_t *lp = lp_out ? lp_out : (int32_t *)malloc((ns ? ns : 1) * 2 * sizeof(int32_t));
    int16_t *dm = dm_out ? dm_out : (int16_t *)malloc((ns ? ns : 1) * sizeof(int16_t));
    memcpy(rot, buf, len);
    orc_rotate_90(rot, len);
    for (size_t i = 0; i < ns; i++) {
        cx[2 * i] = (int32_t)rot[2 * i] - 127;
        cx[2 * i + 1] = (int32_t)rot[2 * i + 1] - 127;
    }
    size_t L = orc_low_pass_complex(d, cx, ns, lp);
    long r = -1;
    if (orc_fm_demod(d, lp, L, dm) >= 0) r = (long)orc_low_pass_real(d, dm, L, out);
    if (n_lp) *n_lp = L;
    free(rot);
    free(cx);
    if (!lp_out) free(lp);
    if (!dm_out) free(dm);
    return r;
}

/* Same arithmetic with the reference's structure: one freshly allocated vector per stage
 * (:257-268).  This is the leg timed as "the reference's CPU path (C restatement)". */
long orc_demodulate_ref_like(orc_demod *d, const uint8_t *buf, size_t len, int16_t *out) {
    if (len % 8 != 0) return -1;
    size_t ns = len / 2;
    uint8_t *v0 = (uint8_t *)malloc(len ? len : 1); /* buf.to_vec() at the call site :80/:127 */
    memcpy(v0, buf, len);
    orc_rotate_90(v0, len);                          /* :257 */
    int16_t *v1 = (int16_t *)malloc((len ? len : 1) * sizeof(int16_t));
    orc_centre(v0, len, v1);                         /* :258 */
    free(v0);
    int32_t *v2 = (int32_t *)malloc((ns ? ns : 1) * 2 * sizeof(int32_t));
    orc_buf_to_complex(v1, len, v2);                 /* :259 */
    free(v1);
    int32_t *v3 = (int32_t *)malloc((ns ? ns : 1) * 2 * sizeof(int32_t));
    size_t L = orc_low_pass_complex(d, v2, ns, v3);  /* :261 */
    free(v2);
    int16_t *v4 = (int16_t *)malloc((L ? L : 1) * sizeof(int16_t));
    long r = -1;
    if (orc_fm_demod(d, v3, L, v4) >= 0)             /* :264 */
        r = (long)orc_low_pass_real(d, v4, L, out);  /* :267 */
    free(v3);
    free(v4);
    return r;
}

/* BASELINE.md §3 leg 2, "oracle_fused": the same call (:256-269) as ONE pass over the buffer with no intermediate
 * vectors — rotate_90 (:285-295) and the centring (:258) are folded into the per-phase sample formulas, every
 * completed boxcar window (:337-352) goes straight through the discriminator (:355-367; f64 path for the first window
 * of the call) into the fractional boxcar (:408-426).  Bit-identical to orc_demodulate (tests/test_oracle.py). */
long orc_demodulate_fused(orc_demod *d, const uint8_t *buf, size_t len, int16_t *out) {
    if (len % 8 != 0) return -1;
    const size_t ns = len / 2;
    const uint64_t D = d->config.downsample;
    if (D == 0 || (d->prev_index + ns) / D < 2) return -1;   /* fm_demod's assert, :356 */
    const uint32_t slow = d->config.rate_resample, fast = d->config.rate_out;
    const int32_t div = (int32_t)(slow ? fast / slow : 0u);
    int32_t re = d->lp_now_re, im = d->lp_now_im, pre = d->demod_pre_re, pim = d->demod_pre_im;
    int32_t now_lpr = d->now_lpr, lpr_idx = d->prev_lpr_index;
    uint64_t pi = d->prev_index;
    int first = 1;
    size_t w = 0;
    for (size_t g = 0; g < len; g += 8) {
        const uint8_t *b = buf + g;
        /* rotated, centred samples of this 8-byte group */
        const int32_t sr[4] = {(int32_t)b[0] - 127, 128 - (int32_t)b[3], 128 - (int32_t)b[4], (int32_t)b[7] - 127};
        const int32_t si[4] = {(int32_t)b[1] - 127, (int32_t)b[2] - 127, 128 - (int32_t)b[5], 128 - (int32_t)b[6]};
        for (int k = 0; k < 4; k++) {
            re = wadd(re, sr[k]);
            im = wadd(im, si[k]);
            if (++pi < D) continue;
            pi = 0;
            int32_t pcm = first ? orc_polar_discriminant(re, im, pre, pim) : orc_polar_discriminant_fast(re, im, pre, pim);
            first = 0;
            pre = re;
            pim = im;
            re = im = 0;
            now_lpr = wadd(now_lpr, (int32_t)(int16_t)(uint16_t)(uint32_t)pcm);
            lpr_idx = wadd(lpr_idx, (int32_t)slow);
            if (lpr_idx < (int32_t)fast) continue;
            out[w++] = (int16_t)(uint16_t)(uint32_t)tdiv(now_lpr, div);
            lpr_idx = wsub(lpr_idx, (int32_t)fast);
            now_lpr = 0;
        }
    }
    d->lp_now_re = re, d->lp_now_im = im, d->demod_pre_re = pre, d->demod_pre_im = pim;
    d->now_lpr = now_lpr, d->prev_lpr_index = lpr_idx, d->prev_index = pi;
    return (long)w;
}

typedef struct {
    const orc_demod_config *cfg;
    const uint8_t *buf;
    size_t buf_len, n_bufs, out_cap;
    int16_t *out;
    int threads;
    long *totals;
    int *failed;
    int fused;
} orc_many_ctx;

static void orc_many_range(long lo, long hi, int tid, void *p) {
    orc_many_ctx *c = (orc_many_ctx *)p;
    orc_demod d;
    orc_demod_init(&d, c->cfg);
    int16_t *tmp = (int16_t *)malloc((c->buf_len / 2 + 1) * sizeof(int16_t));
    long total = 0;
    for (long b = lo; b < hi; b++) {
        long r = c->fused ? orc_demodulate_fused(&d, c->buf + (size_t)b * c->buf_len, c->buf_len, tmp)
                          : orc_demodulate_ref_like(&d, c->buf + (size_t)b * c->buf_len, c->buf_len, tmp);
        if (r < 0) {
            c->failed[tid] = 1;
            break;
        }
        /* keep the result observable without serialising the threads on `out` */
        if (c->out && c->out_cap && r > 0) c->out[((size_t)b * 7) % c->out_cap] = tmp[r - 1];
        total += r;
    }
    c->totals[tid] = total;
    free(tmp);
}

long orc_demodulate_many_mt(const orc_demod_config *cfg, const uint8_t *buf, size_t buf_len,
                            size_t n_bufs, int16_t *out, size_t out_cap, int threads) {
    return orc_demodulate_many_mt2(cfg, buf, buf_len, n_bufs, out, out_cap, threads, 0);
}

long orc_demodulate_many_mt2(const orc_demod_config *cfg, const uint8_t *buf, size_t buf_len,
                             size_t n_bufs, int16_t *out, size_t out_cap, int threads, int fused) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n_bufs) threads = (int)(n_bufs ? n_bufs : 1);
    long *totals = (long *)calloc((size_t)threads, sizeof(long));
    int *failed = (int *)calloc((size_t)threads, sizeof(int));
    orc_many_ctx c = {cfg, buf, buf_len, n_bufs, out_cap, out, threads, totals, failed, fused};
    orc_par_for(threads, (long)n_bufs, orc_many_range, &c);
    long total = 0;
    int bad = 0;
    for (int t = 0; t < threads; t++) {
        total += totals[t];
        bad |= failed[t];
    }
    free(totals);
    free(failed);
    return bad ? -1 : total;
}

/* ------------------------------------------------------------------------------------ */
/* Optional audio post-stages (the reference has none — "parity unpinned"): rtl_fm's deemph_filter and
 * dc_block_filter restated, plus this project's output_scale / raw-byte squelch (definitions in include/sdr_b200.h).   */
void orc_post_init(orc_post *p, uint32_t output_scale, uint32_t squelch_level, uint32_t deemph_a, uint32_t dc_block) {
    memset(p, 0, sizeof(*p));
    p->output_scale = output_scale;
    p->squelch_level = squelch_level;
    p->deemph_a = deemph_a;
    p->dc_block = dc_block;
}

void orc_post_process(orc_post *p, int16_t *x, size_t n, const uint8_t *raw, size_t raw_len) {
    if (n == 0) return;
    int gate_open = 1;
    if (p->squelch_level && raw && raw_len) {
        unsigned long long ss = 0;
        for (size_t i = 0; i < raw_len; i++) {
            int v = 2 * (int)raw[i] - 255;
            ss += (unsigned long long)(v * v);
        }
        gate_open = (unsigned __int128)ss * 64u >= (unsigned __int128)p->squelch_level * p->squelch_level * raw_len;
    }
    const int scale = p->output_scale ? (int)p->output_scale : 1;
    if (scale != 1 || !gate_open)
        for (size_t i = 0; i < n; i++) {
            int v = gate_open ? (int)x[i] * scale : 0;
            x[i] = (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
        }
    if (p->deemph_a) {   /* rtl_fm deemph_filter */
        const int a = (int)p->deemph_a;
        int avg = p->deemph_avg;
        for (size_t i = 0; i < n; i++) {
            int d = (int)x[i] - avg;
            avg += d > 0 ? (d + a / 2) / a : (d - a / 2) / a;
            x[i] = (int16_t)avg;
        }
        p->deemph_avg = avg;
    }
    if (p->dc_block) {   /* rtl_fm dc_block_filter */
        long long sum = 0;
        for (size_t i = 0; i < n; i++) sum += x[i];
        int avg = (int)(sum / (long long)n);
        avg = (avg + p->dc_avg * 9) / 10;
        for (size_t i = 0; i < n; i++) x[i] = (int16_t)(uint16_t)(uint32_t)((int)x[i] - avg);
        p->dc_avg = avg;
    }
}

/* ------------------------------------------------------------------------------------ */
/* f64 extension path (own definition — parity unpinned by the reference)               */

static const double ORC_PI = 3.14159265358979323846264338327950288;

int orc_fx_init(orc_fx *s, const float *taps, uint32_t n_taps, uint32_t decim, const float *taps2,
                uint32_t n_taps2, uint32_t up, uint32_t down, double gain) {
    memset(s, 0, sizeof(*s));
    if (n_taps < 1 || decim < 1) return -1;
    if (n_taps2 && (up < 1 || down < 1)) return -1;
    s->n_taps = n_taps;
    s->decim = decim;
    s->n_taps2 = n_taps2;
    s->up = up ? up : 1;
    s->down = down ? down : 1;
    s->gain = gain;
    s->taps = (double *)malloc(n_taps * sizeof(double));
    for (uint32_t k = 0; k < n_taps; k++) s->taps[k] = (double)taps[k];
    if (n_taps2) {
        s->taps2 = (double *)malloc(n_taps2 * sizeof(double));
        for (uint32_t k = 0; k < n_taps2; k++) s->taps2[k] = (double)taps2[k];
    }
    s->hist_re = (double *)calloc(n_taps, sizeof(double));
    s->hist_im = (double *)calloc(n_taps, sizeof(double));
    s->dhist = NULL;
    s->dhist_len = 0;
    return 0;
}

void orc_fx_free(orc_fx *s) {
    free(s->taps);
    free(s->taps2);
    free(s->hist_re);
    free(s->hist_im);
    free(s->dhist);
    memset(s, 0, sizeof(*s));
}

/* x1.  Global sample index g = s->n_in + i.  hist_* holds xc[g0-(T-1) .. g0-1]. */
size_t orc_fx_low_pass(orc_fx *s, const uint8_t *iq, size_t n, double *out_pairs) {
    const uint32_t T = s->n_taps, D = s->decim;
    const size_t H = T - 1;
    double *xr = (double *)malloc((H + n + 1) * sizeof(double));
    double *xi = (double *)malloc((H + n + 1) * sizeof(double));
    memcpy(xr, s->hist_re, H * sizeof(double));
    memcpy(xi, s->hist_im, H * sizeof(double));
    for (size_t i = 0; i < n; i++) {
        xr[H + i] = (double)iq[2 * i] - 127.0;
        xi[H + i] = (double)iq[2 * i + 1] - 127.0;
    }
    size_t w = 0;
    const uint64_t g0 = s->n_in;
    /* first m with (m+1)D-1 >= g0 is floor(g0/D) */
    uint64_t m = g0 / D;
    for (;; m++) {
        uint64_t last = (m + 1) * (uint64_t)D - 1;
        if (last >= g0 + n) break;
        size_t li = (size_t)(last - g0) + H; /* index of the newest sample in xr/xi */
        double ar = 0.0, ai = 0.0;
        for (uint32_t k = 0; k < T; k++) {
            ar += s->taps[k] * xr[li - k];
            ai += s->taps[k] * xi[li - k];
        }
        out_pairs[2 * w] = ar;
        out_pairs[2 * w + 1] = ai;
        w++;
    }
    /* new history = last T-1 samples of [hist | new] */
    if (H) {
        memmove(s->hist_re, xr + n, H * sizeof(double));
        memmove(s->hist_im, xi + n, H * sizeof(double));
    }
    s->n_in += n;
    s->n_y += w;
    free(xr);
    free(xi);
    return w;
}

/* x2 */
size_t orc_fx_fm_demod(orc_fx *s, const double *y, size_t n, double *out) {
    for (size_t i = 0; i < n; i++) {
        double ar = y[2 * i], ai = y[2 * i + 1];
        double cre = ar * s->prev_re + ai * s->prev_im;
        double cim = ai * s->prev_re - ar * s->prev_im;
        /* zero predecessor (stream start) gives 0 by definition (never +-pi from a signed zero) */
        out[i] = (cre == 0.0 && cim == 0.0) ? 0.0 : s->gain * atan2(cim, cre);
        s->prev_re = ar;
        s->prev_im = ai;
    }
    return n;
}

/* x3.  The oracle keeps the whole discriminator stream (it is test infrastructure). */
size_t orc_fx_resample(orc_fx *s, const double *d, size_t n, double *out) {
    const uint64_t L = s->up, M = s->down;
    const uint32_t T2 = s->n_taps2;
    uint64_t P0 = s->dhist_len;
    s->dhist = (double *)realloc(s->dhist, (P0 + n + 1) * sizeof(double));
    memcpy(s->dhist + P0, d, n * sizeof(double));
    s->dhist_len = P0 + n;
    const uint64_t P1 = P0 + n;
    /* outputs i with floor(iM/L) in [P0, P1)  <=>  i in [ceil(P0 L/M), ceil(P1 L/M)) */
    uint64_t i0 = (P0 * L + M - 1) / M, i1 = (P1 * L + M - 1) / M;
    size_t w = 0;
    for (uint64_t i = i0; i < i1; i++) {
        uint64_t t = i * M; /* index in the zero-stuffed stream */
        double acc = 0.0;
        /* p from floor(t/L) downwards while t - pL < T2 */
        uint64_t p = t / L;
        for (;;) {
            uint64_t k = t - p * L;
            if (k >= T2) break;
            acc += s->taps2[k] * s->dhist[p];
            if (p == 0) break;
            p--;
        }
        out[w++] = acc;
    }
    s->n_a += w;
    return w;
}

size_t orc_fx_process(orc_fx *s, const uint8_t *iq, size_t n, double *y_out, double *d_out,
                      double *a_out, size_t *n_y_out) {
    size_t cap = n / s->decim + 2;
    double *y = y_out ? y_out : (double *)malloc(cap * 2 * sizeof(double));
    double *d = d_out ? d_out : (double *)malloc(cap * sizeof(double));
    size_t ny = orc_fx_low_pass(s, iq, n, y);
    orc_fx_fm_demod(s, y, ny, d);
    size_t na = 0;
    if (s->n_taps2 && a_out) na = orc_fx_resample(s, d, ny, a_out);
    if (n_y_out) *n_y_out = ny;
    if (!y_out) free(y);
    if (!d_out) free(d);
    return na;
}

/* x4: direct definition.  theta_c(n) = 2*pi*((fw_c*n) mod 2^32)/2^32, mix by e^{-j theta}. */
size_t orc_fx_channelise(const uint8_t *iq, size_t n, const float *taps, uint32_t T, uint32_t D,
                         const uint32_t *fw, uint32_t C, double gain, double *y_out, double *d_out) {
    size_t M = n / D;
    double *mr = (double *)malloc((n + 1) * sizeof(double));
    double *mi = (double *)malloc((n + 1) * sizeof(double));
    for (uint32_t c = 0; c < C; c++) {
        for (size_t i = 0; i < n; i++) {
            uint32_t ph = (uint32_t)((uint64_t)fw[c] * (uint64_t)i); /* mod 2^32 */
            double th = 2.0 * ORC_PI * ((double)ph / 4294967296.0);
            double cr = cos(th), sr = sin(th);
            double xr = (double)iq[2 * i] - 127.0, xi = (double)iq[2 * i + 1] - 127.0;
            /* (xr + j xi) * (cos - j sin) */
            mr[i] = xr * cr + xi * sr;
            mi[i] = xi * cr - xr * sr;
        }
        double pr = 0.0, pi_ = 0.0;
        for (size_t m = 0; m < M; m++) {
            size_t last = (m + 1) * (size_t)D - 1;
            double ar = 0.0, ai = 0.0;
            for (uint32_t k = 0; k < T && k <= last; k++) {
                ar += (double)taps[k] * mr[last - k];
                ai += (double)taps[k] * mi[last - k];
            }
            if (y_out) {
                y_out[((size_t)c * M + m) * 2] = ar;
                y_out[((size_t)c * M + m) * 2 + 1] = ai;
            }
            if (d_out) {
                double cre = ar * pr + ai * pi_, cim = ai * pr - ar * pi_;
                d_out[(size_t)c * M + m] = (cre == 0.0 && cim == 0.0) ? 0.0 : gain * atan2(cim, cre);
            }
            pr = ar;
            pi_ = ai;
        }
    }
    free(mr);
    free(mi);
    return M;
}

/* Optimised f32 CPU port for the timing legs: zero history, pthreads over FIR outputs. */
typedef struct {
    const uint8_t *iq;
    const float *rt, *taps2;
    float *yr, *yi, *d, *audio;
    uint32_t T, D, T2;
    uint64_t L, M;
    float gain;
} orc_f32_ctx;

static void orc_f32_fir(long lo, long hi, int tid, void *p) {
    (void)tid;
    orc_f32_ctx *c = (orc_f32_ctx *)p;
    const uint32_t T = c->T;
    for (long m = lo; m < hi; m++) {
        long last = (m + 1) * (long)c->D - 1;
        long first = last - (long)T + 1;
        uint32_t k0 = 0;
        if (first < 0) {
            k0 = (uint32_t)(-first);
            first = 0;
        }
        const uint8_t *q = c->iq + 2 * first;
        const float *rt = c->rt + k0;
        uint32_t cnt = T - k0;
        float ar = 0.f, ai = 0.f;
        for (uint32_t k = 0; k < cnt; k++) {
            float h = rt[k];
            ar += h * ((float)q[2 * k] - 127.f);
            ai += h * ((float)q[2 * k + 1] - 127.f);
        }
        c->yr[m] = ar;
        c->yi[m] = ai;
    }
}
static void orc_f32_disc(long lo, long hi, int tid, void *p) {
    (void)tid;
    orc_f32_ctx *c = (orc_f32_ctx *)p;
    for (long m = lo; m < hi; m++) {
        float pr = m ? c->yr[m - 1] : 0.f, pi_ = m ? c->yi[m - 1] : 0.f;
        float cre = c->yr[m] * pr + c->yi[m] * pi_, cim = c->yi[m] * pr - c->yr[m] * pi_;
        c->d[m] = (cre == 0.f && cim == 0.f) ? 0.f : c->gain * atan2f(cim, cre);
    }
}
static void orc_f32_res(long lo, long hi, int tid, void *p) {
    (void)tid;
    orc_f32_ctx *c = (orc_f32_ctx *)p;
    for (long i = lo; i < hi; i++) {
        uint64_t t = (uint64_t)i * c->M;
        float acc = 0.f;
        uint64_t q = t / c->L;
        for (;;) {
            uint64_t k = t - q * c->L;
            if (k >= c->T2) break;
            acc += c->taps2[k] * c->d[q];
            if (q == 0) break;
            q--;
        }
        c->audio[i] = acc;
    }
}

size_t orc_fx_process_f32_mt(const uint8_t *iq, size_t n, const float *taps, uint32_t T, uint32_t D,
                             const float *taps2, uint32_t T2, uint32_t up, uint32_t down, float gain,
                             float *audio, size_t cap, int threads) {
    size_t M = n / D;
    if (threads < 1) threads = 1;
    float *yr = (float *)malloc((M + 1) * sizeof(float));
    float *yi = (float *)malloc((M + 1) * sizeof(float));
    float *d = (float *)malloc((M + 1) * sizeof(float));
    /* reversed taps so the inner loop walks memory forwards */
    float *rt = (float *)malloc(T * sizeof(float));
    for (uint32_t k = 0; k < T; k++) rt[k] = taps[T - 1 - k];
    orc_f32_ctx c = {iq, rt, taps2, yr, yi, d, audio, T, D, T2, up ? up : 1, down ? down : 1, gain};
    orc_par_for(threads, (long)M, orc_f32_fir, &c);
    orc_par_for(threads, (long)M, orc_f32_disc, &c);
    size_t na = 0;
    if (T2) {
        na = (size_t)(((uint64_t)M * c.L + c.M - 1) / c.M);
        if (na > cap) na = cap;
        orc_par_for(threads, (long)na, orc_f32_res, &c);
    } else {
        na = M < cap ? M : cap;
        memcpy(audio, d, na * sizeof(float));
    }
    free(yr);
    free(yi);
    free(d);
    free(rt);
    return na;
}

/* ------------------------------------------------------------------------------------ */
static inline uint64_t orc_mix64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void orc_synth_fill(uint8_t *buf, size_t len, uint64_t seed, uint64_t byte_offset) {
    for (size_t i = 0; i < len; i++) {
        uint64_t b = byte_offset + i;
        buf[i] = (uint8_t)(orc_mix64(seed, b >> 3) >> (8 * (b & 7)));
    }
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
