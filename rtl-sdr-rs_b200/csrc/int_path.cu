// int_path.cu — reference-exact integer path of `Demod` (examples/simple_fm.rs:232-427) on sm_100a.
//
// The reference is a chain of sequential, stateful loops.  Every piece of its state is either a
// pure function of the running sample count (prev_index :234, prev_lpr_index :236) or a partial
// sum / last value at a cut (lp_now :237, demod_pre :238, now_lpr :235), so every output has a
// closed-form window and the whole chain parallelises:
//
//   lowpassed w  = sum of samples [w*D - p0, (w+1)*D - p0)            (low_pass_complex :337-352)
//   demod k      = disc(lp[k], lp[k-1]);  k first-in-call -> f64 atan2 (fm_demod :355-367)
//   audio e      = (sum of demod [J(e-1), J(e))) / (fast/slow),
//                  J(e) = ceil(((e+1)*fast - q0)/slow)                (low_pass_real :408-426)
//
// k_demod_fused does rotate_90 (:276-299) + `- 127` (:258) + all three stages for one tile of
// audio outputs per CTA: the tile's raw bytes come in with one 1-D bulk async copy (TMA engine,
// SASS UBLKCP) into shared memory, the intermediate streams never leave the SM, and only the i16
// audio goes back to HBM (2.06 algorithmic bytes per complex input sample).
#include <cmath>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace sdr {

struct IntState {   // the data-dependent part of struct Demod, device resident between chunks
    int32_t lp_now_re, lp_now_im, demod_pre_re, demod_pre_im, now_lpr, pad0, pad1, pad2;
};

// (angle/PI*16384) as i32 for the eight exact octant directions and (0,0), evaluated once on the
// host with the platform libm exactly like examples/simple_fm.rs:370-374 does (Rust's f64::atan2
// is the platform libm).  Index: see octant_index().
struct OctTable {
    int32_t v[9];
};

// n / d for n < 2^31 with a host-computed round-up magic (d = rate_resample is fixed per handle)
struct UDiv {
    uint32_t magic, shift;   // shift == 0xffffffff => d == 1
};

// n / d for n < 2^63 (d fixed per handle / per batch): q = umul64hi(n, magic) >> shift
struct UDiv64 {
    unsigned long long magic;
    uint32_t shift;   // 0xffffffff => d == 1
};

struct FusedArgs {
    UDiv div_slow, div_audio;
    UDiv64 d64_slow, d64_S, d64_D;
    const uint8_t *in;
    int16_t *out;
    const IntState *st_in;
    IntState *st_out;
    unsigned long long n_samples, Ltot, Etot;
    uint32_t S, D, p0, q0, fast, slow;
    int32_t div;
    uint32_t EB, lp_cap, tile_cap, dm_off;   // dm_off: byte offset of the demodulated-sample array in shared memory
    OctTable oct;
};

// ---- device arithmetic, bit-exact with the reference's wrapping semantics ---------------------
__device__ __forceinline__ int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
__device__ __forceinline__ int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
__device__ __forceinline__ int32_t tdiv(int32_t a, int32_t b) {
    if (b == 0) return 0;                              // the reference would panic; same as oracle
    if (a == INT32_MIN && b == -1) return INT32_MIN;
    return a / b;
}

// Demod::fast_atan2, examples/simple_fm.rs:383-405 — the i64 product is wrapped to i32 BEFORE the divide.
// Slow form: every corner the reference can reach (den == 0, wrapped |y|, INT32_MIN / -1) with the hardware-emulated
// ~28-instruction integer division.
__device__ __noinline__ int32_t d_fast_atan2_slow(int32_t y, int32_t x) {
    const int32_t pi4 = 1 << 12, pi34 = 3 * (1 << 12);
    if (x == 0 && y == 0) return 0;
    const int32_t yabs = y < 0 ? wsub(0, y) : y;
    const bool xpos = x >= 0;
    const int32_t num = (int32_t)((uint32_t)(xpos ? wsub(x, yabs) : wadd(x, yabs)) << 12);
    const int32_t den = xpos ? wadd(x, yabs) : wsub(yabs, x);
    const int32_t angle = wsub(xpos ? pi4 : pi34, tdiv(num, den));
    return y < 0 ? wsub(0, angle) : angle;
}

// Both branches of :396-400 share ONE divide (operands selected first).  (pi4 as i64 * v as i64) as i32 with
// pi4 = 2^12 is the low 32 bits of v << 12, so |num| = m * 4096 with m <= 2^19: EXACT in f32.  When 0 < den < 2^24
// (den = |x| + |y| unwrapped, also exact in f32) the true quotient Q is at most 2^21 — either den >= 2^10, or
// den < 2^10 and then |v| <= den means nothing wrapped and Q <= 4096.  The estimate
//     e = fma(f32|num|, rcp.approx(f32 den), -0.5)
// carries an absolute error below 2^21 * (2^-23 [rcp] + 2^-24 [fma rounding]) = 0.375, so Q - 0.875 < e < Q - 0.125 and
// trunc(e) (a negative e saturates to 0) is floor(Q) or floor(Q) - 1: ONE exact remainder test (`rem >= den`) repairs it.
// BOUNDED = the caller guarantees |x| + |y| < 2^30, so that nothing wraps BEFORE the shift (true for every product of two
// boxcar sums up to downsample 128: 2 * (128 D)^2 <= 2^29); then den == 0 iff x == y == 0 (result 0, :384-386) and there
// is no out-of-line path at all.  Beyond den = 2^24 the f32 image of den is rounded (relative 2^-24), but there
// Q = |num| / den <= 2^31 / 2^24 = 128, so the estimate's absolute error is below 128 * 2^-22 and the same one-step repair
// holds; the repair compares the exact 32-bit remainder (q * den <= |num| never wraps).  Otherwise everything the
// estimate does not cover (den <= 0: x = y = 0 and the wrapped INT32_MIN cases; |den| >= 2^24) takes the slow form.
template <bool BOUNDED>
__device__ __forceinline__ int32_t d_fast_atan2_t(int32_t y, int32_t x) {
    const int32_t pi4 = 1 << 12, pi34 = 3 * (1 << 12);
    const bool xpos = x >= 0;
    int32_t num, den;
    if (BOUNDED) {   // nothing wraps before the shift: x - |y| = |x| - |y| (x >= 0), x + |y| = -(|x| - |y|) (x < 0)
        const int32_t ax = abs(x), ay = abs(y), m = ax - ay;
        den = ax + ay;
        num = (int32_t)((uint32_t)(xpos ? m : -m) << 12);
    } else {
        const int32_t yabs = y < 0 ? wsub(0, y) : y;
        num = (int32_t)((uint32_t)(xpos ? wsub(x, yabs) : wadd(x, yabs)) << 12);
        den = xpos ? wadd(x, yabs) : wsub(yabs, x);
        if ((uint32_t)(den - 1) >= (1u << 24) - 1u) return d_fast_atan2_slow(y, x);
    }
    const uint32_t an = num < 0 ? 0u - (uint32_t)num : (uint32_t)num;
    // |num| goes to f32 as fabs(f32(num)): converting the integer abs() would let ptxas pick a SIGNED conversion
    // (I2FP.F32.S32 after IABS), which turns |INT32_MIN| = 2^31 into -2^31
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__uint2float_rn((uint32_t)den)));
    uint32_t q = __float2uint_rz(fmaf(fabsf(__int2float_rn(num)), r, -0.5f));
    // q is floor(Q) or floor(Q) - 1 (never more: q * den <= |num| < 2^32, so the remainder is exact in 32 bits)
    if (an - q * (uint32_t)den >= (uint32_t)den) q++;
    const int32_t quo = num < 0 ? (int32_t)(0u - q) : (int32_t)q;
    const int32_t angle = (xpos ? pi4 : pi34) - quo;
    const int32_t res = y < 0 ? -angle : angle;
    return (BOUNDED && den == 0) ? 0 : res;   // den == 0: rcp = inf, e = NaN -> q = 0, then repaired to 1; overridden here
}
__device__ __forceinline__ int32_t d_fast_atan2(int32_t y, int32_t x) { return d_fast_atan2_t<false>(y, x); }

// a * b.conj() for Complex<i32>, wrapping (examples/simple_fm.rs:371,378)
__device__ __forceinline__ void d_cmul_conj(int2 a, int2 b, int32_t &cre, int32_t &cim) {
    cre = wadd(wmul(a.x, b.x), wmul(a.y, b.y));
    cim = wsub(wmul(a.y, b.x), wmul(a.x, b.y));
}

// Demod::polar_discriminant, :370-374.  Exact octant directions come from the host-libm table
// (they are the only inputs whose f64 result sits exactly on an integer, where a 1-ulp
// difference between libm implementations would change the truncated value).
__device__ __forceinline__ int32_t d_polar_f64(int32_t cre, int32_t cim, const OctTable &oct) {
    if (cre == 0 && cim == 0) return oct.v[8];
    int64_t ax = cre < 0 ? -(int64_t)cre : cre, ay = cim < 0 ? -(int64_t)cim : cim;
    if (cim == 0) return oct.v[cre > 0 ? 0 : 4];
    if (cre == 0) return oct.v[cim > 0 ? 2 : 6];
    if (ax == ay) return oct.v[cim > 0 ? (cre > 0 ? 1 : 3) : (cre > 0 ? 7 : 5)];
    double angle = atan2((double)cim, (double)cre);
    double v = angle / 3.14159265358979323846264338327950288 * 16384.0;
    return (int32_t)v;   // truncation toward zero == Rust `as i32` in range
}

// u8 x s8 dot product of four byte lanes with 32-bit accumulate (SASS IDP.4A)
__device__ __forceinline__ int32_t dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int32_t c) {
    int32_t d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}

// rotate_90 (:285-295) + `- 127` (:258) + boxcar over the tile samples [pos, end), two samples per
// 32-bit word.  A word holds samples (n, n+1) with n even, so its rotation phase pair is n%4 in {0,2}:
//   n%4==0: re += I0 - 127 + 128 - Q1,  im += Q0 - 127 + I1 - 127  -> coef re [+1,0,0,-1], im [0,+1,+1,0]
//   n%4==2: re += 128 - I0 + Q1 - 127,  im += 128 - Q0 + 128 - I1  -> coef re [-1,0,0,+1], im [0,-1,-1,0]
// (the "+1 asymmetry" of 255-x then -127 lives in the constants).  Odd window edges use half-word masks.
__device__ __forceinline__ void boxcar_rot(const uint32_t *w32, int pos, int end, int32_t &re, int32_t &im) {
    if (pos >= end) return;
    if (pos & 1) {                  // head: upper half of its word; phase 1 or 3
        const uint32_t w = w32[pos >> 1];
        const bool p2 = pos & 2;
        re = dp4a_us(w, p2 ? 0x01000000u : 0xFF000000u, re) + (p2 ? -127 : 128);
        im = dp4a_us(w, p2 ? 0x00FF0000u : 0x00010000u, im) + (p2 ? 128 : -127);
        pos++;
    }
    // whole words: peel one so that the rest are (phase-0 word, phase-2 word) pairs with fixed coefficients
    int wi = pos >> 1;
    const int wend = end >> 1;            // exclusive: words fully inside the window
    if (wi < wend && (wi & 1)) {          // a phase-2 word first
        const uint32_t w = w32[wi];
        re = dp4a_us(w, 0x010000FFu, re) + 1;
        im = dp4a_us(w, 0x00FFFF00u, im) + 256;
        wi++;
    }
    int npairs = 0;
    for (; wi + 2 <= wend; wi += 2, npairs++) {   // wi even: the pair is one aligned 8-byte load
        const uint2 w = *reinterpret_cast<const uint2 *>(w32 + wi);
        re = dp4a_us(w.x, 0xFF000001u, re);
        im = dp4a_us(w.x, 0x00010100u, im);
        re = dp4a_us(w.y, 0x010000FFu, re);
        im = dp4a_us(w.y, 0x00FFFF00u, im);
    }
    re += 2 * npairs;                     // (+1) + (+1) per pair
    im += 2 * npairs;                     // (-254) + (+256) per pair
    if (wi < wend) {                      // a trailing phase-0 word
        const uint32_t w = w32[wi];
        re = dp4a_us(w, 0xFF000001u, re) + 1;
        im = dp4a_us(w, 0x00010100u, im) - 254;
        wi++;
    }
    pos = wi << 1;
    if (pos < end) {                // tail: lower half of its word; phase 0 or 2
        const uint32_t w = w32[pos >> 1];
        const bool p2 = pos & 2;
        re = dp4a_us(w, p2 ? 0x000000FFu : 0x00000001u, re) + (p2 ? 128 : -127);
        im = dp4a_us(w, p2 ? 0x0000FF00u : 0x00000100u, im) + (p2 ? 128 : -127);
    }
}

// Sum of the per-sample constants of a DT-sample window that starts at sample phase ph (the `255 - x` then `- 127`
// asymmetry of rotate_90 + centring lives here; the dp4a coefficients carry the +-1 factors).
template <int DT>
struct BoxK {   // constants of a DT-sample window that starts at sample phase ph: sum of the per-sample constants
    static constexpr int re(int ph) {
        int k = 0;
        for (int j = 0; j < DT; j++) {
            const int p = (ph + j) & 3;
            k += (p == 1 || p == 2) ? 128 : -127;   // phase 0: I-127, 1: 128-Q, 2: 128-I, 3: Q-127
        }
        return k;
    }
    static constexpr int im(int ph) {
        int k = 0;
        for (int j = 0; j < DT; j++) {
            const int p = (ph + j) & 3;
            k += (p == 2 || p == 3) ? 128 : -127;   // phase 0: Q-127, 1: I-127, 2: 128-Q, 3: 128-I
        }
        return k;
    }
};

__device__ __forceinline__ unsigned long long udiv64(unsigned long long n, UDiv64 d) {
    return d.shift == 0xffffffffu ? n : (__umul64hi(n, d.magic) >> d.shift);
}
__device__ __forceinline__ uint32_t udiv(uint32_t n, UDiv d) {
    return d.shift == 0xffffffffu ? n : (__umulhi(n, d.magic) >> d.shift);
}

// Exact lowpassed window i of the tile (any D), straight from the raw bytes: the D = 6 path keeps no window
// array in shared memory, so its rare fix-ups and the carried-state outputs recompute what they need.
// (re0, im0) = the carried partial window (lp_now) when i is window 0 of the stream, else 0.
__device__ __noinline__ int2 lp_window_exact(const uint32_t *w32, int32_t off0, uint32_t i, uint32_t D, int32_t re0, int32_t im0) {
    const int32_t base = off0 + (int32_t)(i * D);
    int32_t re = re0, im = im0;
    boxcar_rot(w32, base < 0 ? 0 : base, base + (int32_t)D, re, im);
    return make_int2(re, im);
}

// ---- even downsample, even window start (every stream whose calls hold a multiple of 4 samples, i.e. every length
// rotate_90 accepts): rotate_90 + centre + boxcar + discriminator in one register-resident pass ---------------------
// A lane owns KW consecutive windows = KW * DT * 2 raw bytes (a multiple of 16: 48 bytes for the reference's DT = 6),
// fetched as aligned 128-bit loads (for DT = 6 the 48-byte lane stride is bank-conflict free per quarter warp).  `a0` =
// byte offset (16-B aligned, may be -16) of the chunk that holds the first word of window 0; S = word shift of that word
// inside its chunk (tile-uniform -> template parameter, so every register index, every dp4a coefficient and every
// window constant below is a compile-time literal).  A window is DT/2 words and the rotate_90 phase alternates per
// word, so word S + j*DT/2 + k of the lane's row is a phase-2 word iff that index is odd.  The predecessor of a lane's
// first window comes from the DT/2 words before its row: recomputed by every lane in the direct form, taken from the
// neighbour lane by shuffle in the staged form (lane 0 recomputes it).  The KW results leave as one store.  Groups past the tile and the predecessor of window 0 produce
// values nobody reads (dm[0] is either predecessor-only or repaired by the fix-up loop).
constexpr int kMaxFusedDT = 32;
template <int DT>
struct PassGeom {   // windows per lane: the row must be whole 16-byte (even DT) / 8-byte (odd DT) chunks and fit in registers
    static constexpr bool ODD = (DT & 1) != 0;
    // even DT: the fewest windows whose bytes are whole 16-byte chunks (from 14 up; the hand-tuned counts below that)
    static constexpr int KW = ODD ? 4 : DT >= 14 ? (DT % 8 == 0 ? 1 : DT % 4 == 0 ? 2 : 4) : (DT == 2 ? 8 : DT == 12 ? 2 : 4);
    static constexpr int ROW_WORDS = KW * DT / 2;
    // a lane recomputes the predecessor of its first window itself (direct form) only where that is a small share of
    // its row; rows of one or two windows take it from the neighbour lane by shuffle
    static constexpr bool RECOMPUTE_PRED = !ODD && DT != 8 && (DT <= 13 || KW >= 4);
    static_assert(DT >= 2 && DT <= kMaxFusedDT, "register-resident pass: downsample 2..32");
    static_assert((ROW_WORDS * 4) % (ODD ? 8 : 16) == 0, "a lane's row is a whole number of load chunks");
    // the bounded atan2 needs |x| + |y| <= 2 * (128 * DT)^2 < 2^30
    static_assert(2ll * 128 * 128 * DT * DT < (1ll << 30), "products of two boxcar sums must stay below 2^30");
};
constexpr bool has_fused_pass(int DT) { return DT >= 2 && DT <= kMaxFusedDT; }
// windows per lane (PassGeom<DT>::KW) for the host's tile geometry
constexpr int fused_pass_kw(int DT) { return (DT & 1) ? 4 : DT >= 14 ? (DT % 8 == 0 ? 1 : DT % 4 == 0 ? 2 : 4) : (DT == 2 ? 8 : DT == 12 ? 2 : 4); }

template <int DT, int S, int NTH, bool GLOBAL>
__device__ __forceinline__ void dn_pass_even(const unsigned char *tile, const int32_t a0, const uint32_t ngroups, int16_t *dm) {
    using PG = PassGeom<DT>;
    constexpr int KW = PG::KW, HW = DT / 2, NCH = (PG::ROW_WORDS + S + 3) / 4;   // windows per lane, words per window, chunks
    constexpr uint32_t CRE0 = 0xFF000001u, CIM0 = 0x00010100u;   // phase-0 word: re [+1,0,0,-1], im [0,+1,+1,0]
    constexpr uint32_t CRE2 = 0x010000FFu, CIM2 = 0x00FFFF00u;   // phase-2 word: negated
    if constexpr (GLOBAL && PG::RECOMPUTE_PRED) {   // (DT = 8: the extra chunk costs registers the 40-register budget does not have, 4.4 vs 4.1 TB/s)
        // direct form: every lane computes the predecessor of its first window itself from the chunk(s) before its row
        // (one more 128-bit load, DT more dp4a): no shuffle, no divergent lane-0 branch, no warp-uniform loop —
        // 0.428 -> 0.416 ms on cfg1
        constexpr int PC = HW > S ? (HW - S + 3) / 4 : 0, O = PC * 4;
        for (uint32_t g = threadIdx.x; g < ngroups; g += NTH) {
            const uint4 *p4 = reinterpret_cast<const uint4 *>(tile + a0 + (PG::ROW_WORDS * 4) * (int32_t)g);
            uint32_t v[(PC + NCH) * 4];
#pragma unroll
            for (int c = -PC; c < NCH; c++) {
                const uint4 q = __ldg(p4 + c);   // (an L2 evict-first policy on these loads measured no difference)
                v[O + 4 * c] = q.x, v[O + 4 * c + 1] = q.y, v[O + 4 * c + 2] = q.z, v[O + 4 * c + 3] = q.w;
            }
            int32_t re[KW + 1], im[KW + 1];   // [0] = the window before the row
#pragma unroll
            for (int j = -1; j < KW; j++) {
                const bool neg = (S + (j + 2) * HW) & 1;   // parity of S + j*HW (HW*2 is even)
                int32_t r = neg ? BoxK<DT>::re(2) : BoxK<DT>::re(0), i = neg ? BoxK<DT>::im(2) : BoxK<DT>::im(0);
#pragma unroll
                for (int k = 0; k < HW; k++) {
                    const bool wn = neg != (bool)(k & 1);
                    r = dp4a_us(v[O + S + HW * j + k], wn ? CRE2 : CRE0, r);
                    i = dp4a_us(v[O + S + HW * j + k], wn ? CIM2 : CIM0, i);
                }
                re[j + 1] = r;
                im[j + 1] = i;
            }
            uint32_t o[KW];
#pragma unroll
            for (int j = 0; j < KW; j++) {
                int32_t cre, cim;
                d_cmul_conj(make_int2(re[j + 1], im[j + 1]), make_int2(re[j], im[j]), cre, cim);
                o[j] = (uint32_t)d_fast_atan2_t<true>(cim, cre);
            }
            uint32_t pk[KW / 2];
#pragma unroll
            for (int j = 0; j < KW / 2; j++) pk[j] = __byte_perm(o[2 * j], o[2 * j + 1], 0x5410);
            if (KW == 2) *reinterpret_cast<uint32_t *>(dm + 2 * g) = pk[0];
            if (KW == 4) *reinterpret_cast<uint2 *>(dm + 4 * g) = make_uint2(pk[0], pk[KW / 2 - 1]);
            if (KW == 8) *reinterpret_cast<uint4 *>(dm + 8 * g) = make_uint4(pk[0], pk[KW / 8], pk[KW / 4], pk[KW / 2 - 1]);
        }
        return;
    }
    // staged form (shared-memory tile) and DT = 8: predecessor by shuffle, lane 0 recomputes it
    const int lane = threadIdx.x & 31;
    for (uint32_t gb = threadIdx.x & ~31u; gb < ngroups; gb += NTH) {
        const uint32_t g = gb + lane;
        const unsigned char *p = tile + a0 + (PG::ROW_WORDS * 4) * (int32_t)(g < ngroups ? g : ngroups - 1);
        uint32_t v[NCH * 4];
        {
            // GLOBAL: the chunks come straight from the input buffer (read-only path; lanes that share a 32-byte
            // sector meet in L1), otherwise from the shared-memory tile
            const uint4 *p4 = reinterpret_cast<const uint4 *>(p);
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const uint4 q = GLOBAL ? __ldg(p4 + c) : p4[c];
                v[4 * c] = q.x, v[4 * c + 1] = q.y, v[4 * c + 2] = q.z, v[4 * c + 3] = q.w;
            }
        }
        int32_t re[KW], im[KW];
#pragma unroll
        for (int j = 0; j < KW; j++) {
            const bool neg = (S + j * HW) & 1;
            int32_t r = neg ? BoxK<DT>::re(2) : BoxK<DT>::re(0), i = neg ? BoxK<DT>::im(2) : BoxK<DT>::im(0);
#pragma unroll
            for (int k = 0; k < HW; k++) {
                const bool wn = neg != (bool)(k & 1);
                r = dp4a_us(v[S + HW * j + k], wn ? CRE2 : CRE0, r);
                i = dp4a_us(v[S + HW * j + k], wn ? CIM2 : CIM0, i);
            }
            re[j] = r;
            im[j] = i;
        }
        int32_t pre = __shfl_up_sync(0xffffffffu, re[KW - 1], 1), pim = __shfl_up_sync(0xffffffffu, im[KW - 1], 1);
        if (lane == 0 && g > 0) {   // window KW*g - 1: the HW words before word S of this row
            const uint32_t *w = reinterpret_cast<const uint32_t *>(p) + (S - HW);
            const bool neg = (S + HW) & 1;   // same parity as S - HW
            pre = neg ? BoxK<DT>::re(2) : BoxK<DT>::re(0);
            pim = neg ? BoxK<DT>::im(2) : BoxK<DT>::im(0);
#pragma unroll
            for (int k = 0; k < HW; k++) {
                const bool wn = neg != (bool)(k & 1);
                pre = dp4a_us(w[k], wn ? CRE2 : CRE0, pre);
                pim = dp4a_us(w[k], wn ? CIM2 : CIM0, pim);
            }
        }
        uint32_t o[KW];
#pragma unroll
        for (int j = 0; j < KW; j++) {
            int32_t cre, cim;
            d_cmul_conj(make_int2(re[j], im[j]), j ? make_int2(re[j - 1], im[j - 1]) : make_int2(pre, pim), cre, cim);
            o[j] = (uint32_t)d_fast_atan2_t<true>(cim, cre);
        }
        if (g < ngroups) {
            if constexpr (KW == 1) {
                dm[g] = (int16_t)(uint16_t)o[0];
            } else {
                uint32_t pk[KW / 2];
#pragma unroll
                for (int j = 0; j < KW / 2; j++) pk[j] = __byte_perm(o[2 * j], o[2 * j + 1], 0x5410);
                if (KW == 2) *reinterpret_cast<uint32_t *>(dm + 2 * g) = pk[0];
                if (KW == 4) *reinterpret_cast<uint2 *>(dm + 4 * g) = make_uint2(pk[0], pk[KW / 2 - 1]);
                if (KW == 8) *reinterpret_cast<uint4 *>(dm + 8 * g) = make_uint4(pk[0], pk[KW / 8], pk[KW / 4], pk[KW / 2 - 1]);
            }
        }
    }
}

// ---- odd downsample: windows alternate between even and odd start samples, so the unit is a PAIR of windows =
// 2*DT samples = DT words starting on an even sample: window A = (DT-1)/2 whole words + the low half of the middle
// word, window B = the high half of the middle word + (DT-1)/2 whole words (half-masked dp4a coefficients).  A lane
// owns two pairs (four windows, 8*DT bytes, 64-bit loads); E = 1 when window 0 of the tile starts on an odd sample —
// it is then the B window of a pair whose A window lies before the tile (computed from whatever bytes are there and
// never stored).  S = word shift of the first pair inside its 8-byte chunk.  The rotate_90 phase of row word x is its
// parity, as in the even form (a lane's row starts on an even word).  Direct (global-memory) form only.
template <int DT, int S, int E, int NTH>
__device__ __forceinline__ void dn_pass_odd(const unsigned char *tile, const int32_t a0, const uint32_t ngroups,
                                            const int32_t last_w, int16_t *dm) {
    constexpr int HW = (DT - 1) / 2, RW = 2 * DT, NCH = (RW + S + 1) / 2;   // whole words per window, row words, 8-byte chunks
    constexpr uint32_t CRE0 = 0xFF000001u, CIM0 = 0x00010100u, CRE2 = 0x010000FFu, CIM2 = 0x00FFFF00u;
    constexpr uint32_t LO = 0x0000FFFFu, HI = 0xFFFF0000u;
    const int lane = threadIdx.x & 31;
    for (uint32_t gb = threadIdx.x & ~31u; gb < ngroups; gb += NTH) {
        const uint32_t g = gb + lane;
        const unsigned char *p = tile + a0 + (RW * 4) * (int32_t)(g < ngroups ? g : ngroups - 1);
        uint32_t v[NCH * 2];
        {
            const uint2 *p2 = reinterpret_cast<const uint2 *>(p);
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const uint2 q = __ldg(p2 + c);
                v[2 * c] = q.x, v[2 * c + 1] = q.y;
            }
        }
        int32_t re[4], im[4];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int f = S + q * DT;   // row index of the pair's first word
            // window A: starts on the first sample of word f
            {
                const int ph = 2 * (f & 1);
                int32_t r = BoxK<DT>::re(ph), i = BoxK<DT>::im(ph);
#pragma unroll
                for (int k = 0; k <= HW; k++) {
                    const bool wn = (f + k) & 1;
                    const uint32_t m = k == HW ? LO : 0xFFFFFFFFu;
                    r = dp4a_us(v[f + k], (wn ? CRE2 : CRE0) & m, r);
                    i = dp4a_us(v[f + k], (wn ? CIM2 : CIM0) & m, i);
                }
                re[2 * q] = r;
                im[2 * q] = i;
            }
            // window B: starts on the second sample of word f + HW
            {
                const int ph = (2 * ((f + HW) & 1) + 1) & 3;
                int32_t r = BoxK<DT>::re(ph), i = BoxK<DT>::im(ph);
#pragma unroll
                for (int k = HW; k < DT; k++) {
                    const bool wn = (f + k) & 1;
                    const uint32_t m = k == HW ? HI : 0xFFFFFFFFu;
                    r = dp4a_us(v[f + k], (wn ? CRE2 : CRE0) & m, r);
                    i = dp4a_us(v[f + k], (wn ? CIM2 : CIM0) & m, i);
                }
                re[2 * q + 1] = r;
                im[2 * q + 1] = i;
            }
        }
        int32_t pre = __shfl_up_sync(0xffffffffu, re[3], 1), pim = __shfl_up_sync(0xffffffffu, im[3], 1);
        if (lane == 0 && g > 0) {   // the B window of the pair before this row: row indices S - DT + HW .. S - 1
            const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
            const int f = S - DT;
            const int ph = (2 * ((f + HW) & 1) + 1) & 3;
            pre = BoxK<DT>::re(ph);
            pim = BoxK<DT>::im(ph);
#pragma unroll
            for (int k = HW; k < DT; k++) {
                const bool wn = (f + k) & 1;
                const uint32_t m = k == HW ? HI : 0xFFFFFFFFu;
                pre = dp4a_us(__ldg(w + (f + k)), (wn ? CRE2 : CRE0) & m, pre);
                pim = dp4a_us(__ldg(w + (f + k)), (wn ? CIM2 : CIM0) & m, pim);
            }
        }
        const int32_t w0 = 4 * (int32_t)g - E;   // tile-relative index of the row's first window
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int32_t cre, cim;
            d_cmul_conj(make_int2(re[j], im[j]), j ? make_int2(re[j - 1], im[j - 1]) : make_int2(pre, pim), cre, cim);
            const int32_t o = d_fast_atan2_t<true>(cim, cre);
            const int32_t w = w0 + j;
            if (g < ngroups && (E == 0 || j > 0 || w >= 0) && w <= last_w) dm[w] = (int16_t)(uint16_t)(uint32_t)o;
        }
    }
}

// ---- odd downsample from 15 up, STAGED form: the direct form's rows of two pairs (4 DT words per lane) no longer fit a
// sensible register budget, and narrower global loads cost more L1 sector requests than the memory pipe has.  Here the
// tile's bytes are in shared memory (one bulk copy) and a lane owns ONE pair of windows = DT words, read with 32-bit loads
// at a lane stride of DT words — odd, hence bank-conflict free.  The first word of a lane's pair alternates between even
// and odd word indices from lane to lane, so the rotate_90 phase of a word is not a compile-time constant: the dp4a sums
// of the even-offset and the odd-offset words are kept apart (always with the phase-0 coefficients; the phase-2 ones are
// their negation) and combined with the lane's sign afterwards.  E as in the direct form.
constexpr bool has_staged_odd_pass(int DT) { return (DT & 1) && DT >= 15 && DT <= kMaxFusedDT; }
// ... and the even downsamples with an odd number of words per window (14, 18, 22, 26, 30): in the direct form their rows
// are four windows (28-60 words); staged, a lane owns ONE window = DT/2 words at an odd (conflict-free) lane stride.
constexpr bool has_staged_even_pass(int DT) { return DT % 4 == 2 && DT >= 14 && DT <= kMaxFusedDT; }
constexpr bool has_staged_pass(int DT) { return has_staged_odd_pass(DT) || has_staged_even_pass(DT); }

template <int DT, int E, int NTH>
__device__ __forceinline__ void dn_pass_odd_staged(const unsigned char *tile, const int32_t byte0, const uint32_t npairs,
                                                   const int32_t last_w, int16_t *dm) {
    constexpr int HW = (DT - 1) / 2;
    constexpr uint32_t CRE0 = 0xFF000001u, CIM0 = 0x00010100u, LO = 0x0000FFFFu, HI = 0xFFFF0000u;
    const int lane = threadIdx.x & 31;
    // sums of one window: words [k0, k0 + HW] of the pair, the half-masked one first (B) or last (A)
    auto window = [&](const uint32_t *v, const int k0, const bool is_b, const uint32_t fpar, int32_t &re, int32_t &im) {
        int32_t r[2] = {0, 0}, i[2] = {0, 0};   // by parity of the word's offset inside the pair
#pragma unroll
        for (int k = 0; k <= HW; k++) {
            const uint32_t m = (is_b ? k == 0 : k == HW) ? (is_b ? HI : LO) : 0xFFFFFFFFu;
            r[(k0 + k) & 1] = dp4a_us(v[k0 + k], CRE0 & m, r[(k0 + k) & 1]);
            i[(k0 + k) & 1] = dp4a_us(v[k0 + k], CIM0 & m, i[(k0 + k) & 1]);
        }
        // absolute word parity = fpar ^ offset parity: phase-0 words count +, phase-2 words -
        const int32_t dr = r[0] - r[1], di = i[0] - i[1];
        // start phase of the window: A starts on the first sample of word k0 (phase 0 or 2), B on the second sample of
        // word k0 (phase 1 or 3)
        const uint32_t wpar = (fpar ^ (uint32_t)k0) & 1u;
        const int ph = is_b ? (wpar ? 3 : 1) : (wpar ? 2 : 0), pho = is_b ? (wpar ? 1 : 3) : (wpar ? 0 : 2);
        (void)pho;
        re = (fpar ? -dr : dr) + (wpar ? (is_b ? BoxK<DT>::re(3) : BoxK<DT>::re(2)) : (is_b ? BoxK<DT>::re(1) : BoxK<DT>::re(0)));
        im = (fpar ? -di : di) + (wpar ? (is_b ? BoxK<DT>::im(3) : BoxK<DT>::im(2)) : (is_b ? BoxK<DT>::im(1) : BoxK<DT>::im(0)));
        (void)ph;
    };
    const uint32_t *w0 = reinterpret_cast<const uint32_t *>(tile + byte0);   // byte0 is a multiple of 4 (may be negative)
    const uint32_t basepar = (uint32_t)(byte0 >> 2) & 1u;
    for (uint32_t gb = threadIdx.x & ~31u; gb < npairs; gb += NTH) {
        const uint32_t g = gb + lane;
        const uint32_t gc = g < npairs ? g : npairs - 1;
        const uint32_t *w = w0 + (size_t)DT * gc;
        const uint32_t fpar = (basepar + gc) & 1u;   // DT is odd: the parity of the pair's first word alternates
        uint32_t v[DT];
#pragma unroll
        for (int k = 0; k < DT; k++) v[k] = w[k];
        int32_t re[2], im[2];
        window(v, 0, false, fpar, re[0], im[0]);
        window(v, HW, true, fpar, re[1], im[1]);
        int32_t pre = __shfl_up_sync(0xffffffffu, re[1], 1), pim = __shfl_up_sync(0xffffffffu, im[1], 1);
        if (lane == 0 && g > 0) {   // the B window of the pair before this one: its words HW .. DT-1
            uint32_t pv[HW + 1];
#pragma unroll
            for (int k = 0; k <= HW; k++) pv[k] = w[k - DT + HW];
            // (pv is indexed from 0, the window from offset HW of ITS pair: pass the pair parity that makes the word
            // parities come out right — first word of that pair has parity fpar ^ 1, its word HW has (fpar ^ 1 ^ HW))
            int32_t r[2] = {0, 0}, i[2] = {0, 0};
#pragma unroll
            for (int k = 0; k <= HW; k++) {
                const uint32_t m = k == 0 ? HI : 0xFFFFFFFFu;
                r[(HW + k) & 1] = dp4a_us(pv[k], CRE0 & m, r[(HW + k) & 1]);
                i[(HW + k) & 1] = dp4a_us(pv[k], CIM0 & m, i[(HW + k) & 1]);
            }
            const uint32_t fp = fpar ^ 1u, wpar = (fp ^ (uint32_t)HW) & 1u;
            const int32_t dr = r[0] - r[1], di = i[0] - i[1];
            pre = (fp ? -dr : dr) + (wpar ? BoxK<DT>::re(3) : BoxK<DT>::re(1));
            pim = (fp ? -di : di) + (wpar ? BoxK<DT>::im(3) : BoxK<DT>::im(1));
        }
        const int32_t wfirst = 2 * (int32_t)g - E;   // tile-relative index of the pair's A window
#pragma unroll
        for (int j = 0; j < 2; j++) {
            int32_t cre, cim;
            d_cmul_conj(make_int2(re[j], im[j]), j ? make_int2(re[0], im[0]) : make_int2(pre, pim), cre, cim);
            const int32_t o = d_fast_atan2_t<true>(cim, cre);
            const int32_t wi = wfirst + j;
            if (g < npairs && wi >= 0 && wi <= last_w) dm[wi] = (int16_t)(uint16_t)(uint32_t)o;
        }
    }
}

// Even downsample with DT/2 odd, even window start: one window (DT/2 whole words) per lane, same sign-split sums.
template <int DT, int NTH>
__device__ __forceinline__ void dn_pass_even_staged(const unsigned char *tile, const int32_t byte0, const uint32_t nwin, int16_t *dm) {
    constexpr int HW = DT / 2;
    constexpr uint32_t CRE0 = 0xFF000001u, CIM0 = 0x00010100u;
    static_assert(HW & 1, "an odd number of words per window (conflict-free lane stride)");
    const int lane = threadIdx.x & 31;
    auto window = [&](const uint32_t *w, const uint32_t fpar, int32_t &re, int32_t &im) {
        int32_t r[2] = {0, 0}, i[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < HW; k++) {
            const uint32_t v = w[k];
            r[k & 1] = dp4a_us(v, CRE0, r[k & 1]);
            i[k & 1] = dp4a_us(v, CIM0, i[k & 1]);
        }
        const int32_t dr = r[0] - r[1], di = i[0] - i[1];
        re = (fpar ? -dr : dr) + (fpar ? BoxK<DT>::re(2) : BoxK<DT>::re(0));
        im = (fpar ? -di : di) + (fpar ? BoxK<DT>::im(2) : BoxK<DT>::im(0));
    };
    const uint32_t *w0 = reinterpret_cast<const uint32_t *>(tile + byte0);   // byte0: first byte of window 0, a multiple of 4
    const uint32_t basepar = (uint32_t)(byte0 >> 2) & 1u;
    for (uint32_t gb = threadIdx.x & ~31u; gb < nwin; gb += NTH) {
        const uint32_t g = gb + lane;
        const uint32_t gc = g < nwin ? g : nwin - 1;
        const uint32_t *w = w0 + (size_t)HW * gc;
        const uint32_t fpar = (basepar + gc) & 1u;   // HW is odd: the parity of the window's first word alternates
        int32_t re, im;
        window(w, fpar, re, im);
        int32_t pre = __shfl_up_sync(0xffffffffu, re, 1), pim = __shfl_up_sync(0xffffffffu, im, 1);
        if (lane == 0 && g > 0) window(w - HW, fpar ^ 1u, pre, pim);   // the window before this warp's first
        int32_t cre, cim;
        d_cmul_conj(make_int2(re, im), make_int2(pre, pim), cre, cim);
        const int32_t o = d_fast_atan2_t<true>(cim, cre);
        if (g < nwin) dm[g] = (int16_t)(uint16_t)(uint32_t)o;
    }
}

// ================================================================================================
// Tile = EB consecutive audio outputs.  Shared pieces of the one-CTA-per-tile kernel and the persistent ring.
// ================================================================================================
struct TileInfo {   // geometry of one tile, derived by ONE thread with 64-bit math; everything after it is 32-bit
    unsigned long long wlo, jlo, jhi, e0;   // first lowpassed window held, demod range [jlo, jhi), first audio output
    unsigned long long clo, chi;            // calls whose first window can lie in [jlo, jhi)
    unsigned long long b_lo;                // byte offset of the tile's raw bytes in the input (16-B aligned)
    uint32_t bytes, ne, rb, nlp, tail_from, ntail;
    int32_t off0;                           // sample offset of window wlo inside the tile (negative: starts inside lp_now)
    uint32_t pad;
};

__device__ __forceinline__ void tile_setup(const FusedArgs &a, const uint32_t tile_idx, const bool last, TileInfo &ti) {
    const unsigned long long fast = a.fast, slow = a.slow;
    unsigned long long e0 = (unsigned long long)tile_idx * a.EB;
    unsigned long long e1 = e0 + a.EB < a.Etot ? e0 + a.EB : a.Etot;
    if (e0 > a.Etot) e0 = a.Etot;
    // J(e) = ceil(((e+1)*fast - q0)/slow), J(-1) = 0
    unsigned long long jlo = e0 ? udiv64(e0 * fast - a.q0 + slow - 1, a.d64_slow) : 0ull;
    unsigned long long jhi = last ? a.Ltot : (e1 ? udiv64(e1 * fast - a.q0 + slow - 1, a.d64_slow) : 0ull);
    unsigned long long wlo = jlo ? jlo - 1 : 0ull;
    // relative form: J(e0-1+u) = jlo + ceil((u*fast - rb)/slow) for u >= 1
    uint32_t rb = e0 ? (uint32_t)(jlo * slow - (e0 * fast - a.q0)) : a.q0;
    long long s_lo = (long long)(wlo * a.D) - (long long)a.p0;
    if (s_lo < 0) s_lo = 0;
    unsigned long long s_hi = last ? a.n_samples : jhi * a.D - a.p0;
    unsigned long long b_lo = (2ull * (unsigned long long)s_lo) & ~15ull;
    unsigned long long b_hi = (2ull * s_hi + 15ull) & ~15ull;
    ti.wlo = wlo;
    ti.jlo = jlo;
    ti.jhi = jhi;
    ti.e0 = e0;
    ti.b_lo = b_lo;
    ti.bytes = (uint32_t)(b_hi - b_lo);
    ti.ne = (uint32_t)(e1 - e0);
    ti.rb = rb;
    ti.nlp = (uint32_t)(jhi - wlo);
    ti.off0 = (int32_t)((long long)(wlo * a.D) - (long long)a.p0 - (long long)(b_lo >> 1));
    // tail samples (after the last complete window) feed lp_now' — last tile only
    ti.tail_from = (uint32_t)((a.Ltot * a.D - a.p0) - (b_lo >> 1));
    ti.ntail = (uint32_t)(a.n_samples - (a.Ltot * a.D - a.p0));
}
// second half (not needed to start the copy): the calls whose first window can lie in [jlo, jhi)
__device__ __forceinline__ void tile_setup_calls(const FusedArgs &a, TileInfo &ti) {
    ti.clo = ti.jhi > ti.jlo ? udiv64((ti.jlo + 1) * a.D - a.p0 - 1, a.d64_S) : 1ull;
    ti.chi = ti.jhi > ti.jlo ? udiv64(ti.jhi * a.D - a.p0 - 1, a.d64_S) : 0ull;
}

// Fix-ups after the register-resident pass (rare): dm[i] recomputed exactly for the first sample of each call (fm_demod :359 uses the f64
// polar_discriminant there) and for the two samples that see the carried state on the first tile.
__device__ __forceinline__ void pass_fixups(const FusedArgs &a, const IntState &st, const TileInfo &ti, const uint32_t *w32,
                                          int16_t *dm, const uint32_t tid, const uint32_t nthreads) {
    const unsigned long long c_lo = ti.clo, c_hi = ti.chi, wlo = ti.wlo, jlo = ti.jlo, jhi = ti.jhi;
    const bool tile0 = wlo == 0;
    const uint32_t D = a.D, nlp = ti.nlp;
    const unsigned long long ncall = c_lo <= c_hi ? c_hi - c_lo + 1 : 0ull;
    for (unsigned long long k = tid; k < ncall + (tile0 ? 2u : 0u); k += nthreads) {
        uint32_t i;
        if (k < ncall) {
            const unsigned long long w = udiv64((c_lo + k) * a.S + a.p0, a.d64_D);
            if (w < jlo || w >= jhi) continue;
            i = (uint32_t)(w - wlo);
        } else {
            i = (uint32_t)(k - ncall);
            if (i >= nlp) continue;
        }
        const bool w0 = tile0 && i == 0, w1 = tile0 && i == 1;
        const int2 cur = lp_window_exact(w32, ti.off0, i, D, w0 ? st.lp_now_re : 0, w0 ? st.lp_now_im : 0);
        const int2 prev = w0 ? make_int2(st.demod_pre_re, st.demod_pre_im)
                             : lp_window_exact(w32, ti.off0, i - 1, D, w1 ? st.lp_now_re : 0, w1 ? st.lp_now_im : 0);
        int32_t cre, cim;
        d_cmul_conj(cur, prev, cre, cim);
        // window w holds the first sample of call c iff w*D <= c*S + p0 < (w+1)*D for some c >= 0
        const unsigned long long lo = (wlo + i) * D;
        const unsigned long long c = lo > a.p0 ? udiv64(lo - a.p0 + a.S - 1, a.d64_S) : 0ull;
        const bool first = c * a.S + a.p0 < lo + D;
        dm[i] = (int16_t)(uint16_t)(uint32_t)(first ? d_polar_f64(cre, cim, a.oct) : d_fast_atan2(cim, cre));
    }
}

// Main pass over one tile (rotate_90 + centre + boxcar + discriminator -> dm) for the downsamples that have the
// register-resident form.
template <int DT, int NTH, bool GLOBAL>
__device__ __forceinline__ void dn_pass(const unsigned char *tile, const TileInfo &ti, int16_t *dm, const int tid) {
    const int32_t off0 = ti.off0;
    const uint32_t nlp = ti.nlp;
    if constexpr (PassGeom<DT>::ODD && !GLOBAL) {
        static_assert(has_staged_odd_pass(DT), "staged odd-downsample pass: 15..31");
        const int e = off0 & 1;                                   // window 0 starts on an odd sample: it is a B window
        const int32_t byte0 = 2 * (off0 - e * DT);                // first byte of the pair that holds window 0
        const uint32_t npairs = (nlp + (uint32_t)e + 1) / 2;
        if (e) dn_pass_odd_staged<DT, 1, NTH>(tile, byte0, npairs, (int32_t)nlp - 1, dm);
        else dn_pass_odd_staged<DT, 0, NTH>(tile, byte0, npairs, (int32_t)nlp - 1, dm);
        return;
    } else if constexpr (PassGeom<DT>::ODD) {
        const int e = off0 & 1;                                   // window 0 starts on an odd sample
        const int32_t byte0 = 2 * (off0 - e * DT);                // first byte of the pair that holds window 0
        const int32_t a0 = byte0 & ~7;
        const uint32_t ngroups = (nlp + (uint32_t)e + 3) / 4;
        const int32_t last_w = (int32_t)nlp - 1;
        switch (((byte0 >> 2) & 1) * 2 + e) {
        case 0: dn_pass_odd<DT, 0, 0, NTH>(tile, a0, ngroups, last_w, dm); break;
        case 1: dn_pass_odd<DT, 0, 1, NTH>(tile, a0, ngroups, last_w, dm); break;
        case 2: dn_pass_odd<DT, 1, 0, NTH>(tile, a0, ngroups, last_w, dm); break;
        default: dn_pass_odd<DT, 1, 1, NTH>(tile, a0, ngroups, last_w, dm); break;
        }
        return;
    } else if constexpr (!GLOBAL && has_staged_even_pass(DT)) {
        // even window starts only (the host sends an odd one to the generic kernel)
        dn_pass_even_staged<DT, NTH>(tile, 2 * off0, nlp, dm);
        return;
    } else {
        // even downsample: the host launches this form for even window starts only (an odd one needs an odd prev_index
        // set by hand, sdr_demod_set_state, and takes the generic three-phase kernel)
        constexpr int KW = PassGeom<DT>::KW;
        const int32_t byte0 = 2 * off0;   // first byte of window 0 (negative on a tile that starts inside it)
        const int32_t a0 = byte0 & ~15;
        const uint32_t ngroups = (nlp + KW - 1) / KW;
        switch ((byte0 >> 2) & 3) {
        case 0: dn_pass_even<DT, 0, NTH, GLOBAL>(tile, a0, ngroups, dm); break;
        case 1: dn_pass_even<DT, 1, NTH, GLOBAL>(tile, a0, ngroups, dm); break;
        case 2: dn_pass_even<DT, 2, NTH, GLOBAL>(tile, a0, ngroups, dm); break;
        default: dn_pass_even<DT, 3, NTH, GLOBAL>(tile, a0, ngroups, dm); break;
        }
    }
}

// low_pass_real (:408-426) over one tile: audio output t of the tile sums dm[dbase + r(t) .. dbase + r(t+1)), with
// r(t) = ceil((t*fast - rb)/slow) by magic division, and divides by fast/slow (truncating, :421).
__device__ __forceinline__ void resample_tile(const FusedArgs &a, const IntState &st, const TileInfo &ti, const int16_t *dm,
                                              const uint32_t tid, const uint32_t nthreads) {
    const uint32_t ne = ti.ne, rb = ti.rb, fast = a.fast, slow = a.slow;
    const uint32_t dbase = (uint32_t)(ti.jlo - ti.wlo);   // dm index of demod sample jlo
    const bool e0zero = ti.e0 == 0;
    int16_t *outp = a.out + ti.e0;
    if (a.div == 5 && a.div_slow.shift != 0xffffffffu && rb < slow) {
        // The reference's ratio (170k -> 32k, and any other with fast / slow == 5): every window holds 5 or 6 samples
        // (rb < slow: no unusual carried prev_lpr_index in this tile), n / 5 == umulhi(n, 0xCCCCCCCD) >> 2 for every
        // 32-bit n — straight-line code, no per-output branches.
        const uint32_t mg = a.div_slow.magic, sh = a.div_slow.shift;
        for (uint32_t t = tid; t < ne; t += nthreads) {
            const uint32_t n0 = t * fast - rb + slow - 1;
            const uint32_t r0 = t ? __umulhi(n0, mg) >> sh : 0u;
            const uint32_t r1 = __umulhi(n0 + fast, mg) >> sh;
            const int16_t *dp = dm + dbase + r0;
            int32_t sum = (int32_t)dp[0] + (int32_t)dp[1] + (int32_t)dp[2] + (int32_t)dp[3] + (int32_t)dp[4];
            if (r1 - r0 == 6u) sum += (int32_t)dp[5];
            if (e0zero && t == 0) sum = wadd(sum, st.now_lpr);
            const uint32_t mag = sum < 0 ? (uint32_t)0 - (uint32_t)sum : (uint32_t)sum;
            const uint32_t qm = __umulhi(mag, 0xCCCCCCCDu) >> 2;
            outp[t] = (int16_t)(uint16_t)(sum < 0 ? (uint32_t)0 - qm : qm);
        }
        return;
    }
    if (a.div_slow.shift != 0xffffffffu && a.div_audio.shift != 0xffffffffu && rb < slow && (uint32_t)a.div <= 64u) {
        // any other ratio without an unusual carried prev_lpr_index: every window holds floor(fast/slow) = a.div samples
        // or one more — a tile-uniform count of plain loads and ONE predicated one, no per-load compare, no special
        // cases in the two magic divisions.  (|sum| <= 65 * 32768 except where the carried now_lpr is added.)
        const uint32_t mg = a.div_slow.magic, sh = a.div_slow.shift, base = (uint32_t)a.div;
        for (uint32_t t = tid; t < ne; t += nthreads) {
            const uint32_t n0 = t * fast - rb + slow - 1;
            const uint32_t r0 = t ? __umulhi(n0, mg) >> sh : 0u;
            const uint32_t r1 = __umulhi(n0 + fast, mg) >> sh;
            const int16_t *dp = dm + dbase + r0;
            int32_t sum = 0;
            for (uint32_t j = 0; j < base; j++) sum += (int32_t)dp[j];
            if (r1 - r0 > base) sum += (int32_t)dp[base];
            uint32_t mag, qm;
            if (e0zero && t == 0) {   // the carried partial sum can be anything
                sum = wadd(sum, st.now_lpr);
                mag = sum < 0 ? (uint32_t)0 - (uint32_t)sum : (uint32_t)sum;
                qm = mag >> 31 ? (uint32_t)((int64_t)mag / a.div) : udiv(mag, a.div_audio);
            } else {
                mag = sum < 0 ? (uint32_t)0 - (uint32_t)sum : (uint32_t)sum;
                qm = __umulhi(mag, a.div_audio.magic) >> a.div_audio.shift;
            }
            outp[t] = (int16_t)(uint16_t)(sum < 0 ? (uint32_t)0 - qm : qm);
        }
        return;
    }
    for (uint32_t t = tid; t < ne; t += nthreads) {
        const uint32_t r0 = t ? udiv(t * fast - rb + slow - 1, a.div_slow) : 0u;
        const uint32_t r1 = udiv((t + 1) * fast - rb + slow - 1, a.div_slow);
        int32_t sum = (e0zero && t == 0) ? st.now_lpr : 0;
        const int16_t *dp = dm + dbase + r0;
        const uint32_t cnt = r1 - r0;
#pragma unroll
        for (uint32_t j = 0; j < 8; j++)   // cnt is floor or ceil of fast/slow: eight predicated loads cover the common ratios
            if (j < cnt) sum = wadd(sum, (int32_t)dp[j]);
        for (uint32_t j = 8; j < cnt; j++) sum = wadd(sum, (int32_t)dp[j]);
        // truncating sum / (fast/slow): divide the magnitude with the magic, restore the sign
        const uint32_t mag = sum < 0 ? (uint32_t)0 - (uint32_t)sum : (uint32_t)sum;
        const uint32_t qm = mag >> 31 ? (uint32_t)((int64_t)mag / a.div) : udiv(mag, a.div_audio);
        outp[t] = (int16_t)(uint16_t)(sum < 0 ? (uint32_t)0 - qm : qm);
    }
}

// carried state after the last tile (one thread): now_lpr' = the demodulated samples after the last audio window,
// demod_pre' = the last lowpassed sample
__device__ __forceinline__ void tile_state_out(const FusedArgs &a, const IntState &st, const TileInfo &ti, const int16_t *dm,
                                               const int2 lastlp) {
    const uint32_t dbase = (uint32_t)(ti.jlo - ti.wlo);
    const uint32_t r0 = ti.ne ? udiv(ti.ne * a.fast - ti.rb + a.slow - 1, a.div_slow) : 0u;
    int32_t sum = (a.Etot == 0) ? st.now_lpr : 0;
    for (uint32_t j = dbase + r0; j < ti.nlp; j++) sum = wadd(sum, (int32_t)dm[j]);
    a.st_out->now_lpr = sum;
    a.st_out->demod_pre_re = lastlp.x;
    a.st_out->demod_pre_im = lastlp.y;
}

// One tile per call.  `first_use`: this CTA has not initialised its mbarrier yet; `parity`: phase of the mbarrier
// for this use (a persistent CTA flips it per tile).  Ends with a __syncthreads so the shared-memory tile can be reused.
// DIRECT (batch kernel for the even downsamples up to 12): no shared-memory tile at all — the pass reads its 16-byte chunks straight from the input
// buffer, so a CTA costs only the 6 KB dm array, the SM holds as many CTAs as the register file allows and the HBM
// latency is hidden by warps, not by a per-CTA copy-then-compute phase.  The ring keeps the staged form (its input slots
// are written by the copy engine while the kernel is resident; the mbarrier orders those bytes).
template <int DT, bool DIRECT = false, int NTH = 256>
__device__ __forceinline__ void demod_tile(const FusedArgs &a, const uint32_t tile_idx, const uint32_t n_tiles,
                                           const uint32_t parity, const bool first_use) {
    static_assert(!DIRECT || has_fused_pass(DT), "the direct form exists for the downsamples with a register-resident pass");
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ TileInfo sh_ti;

    // D = 6: the aligned-chunk pass may read the 16 bytes before the tile; staged odd pass: the pair that holds window 0 may
    // start up to 2*DT bytes before it (that half is computed from whatever is there and never stored)
    unsigned char *tile_s = smem + (DT == 6 ? 16 : has_staged_pass(DT) ? 128 : 0);
    int2 *lp = reinterpret_cast<int2 *>(smem + a.tile_cap);
    int16_t *dm = reinterpret_cast<int16_t *>(smem + (DIRECT ? 0 : a.dm_off));
    uint8_t *flag = smem + a.tile_cap + (((size_t)a.lp_cap * 10 + 15) & ~size_t(15));   // 16-byte aligned

    const int tid = threadIdx.x;
    const bool last = tile_idx == n_tiles - 1;
    // a reference, not a copy: the five words are needed by a handful of threads of the first and last tile only, and
    // eight registers held across the pass would be spilled at the 32-register budget of the direct kernel
    const IntState &st = *a.st_in;

    if (tid == 0) {
        if (!DIRECT && first_use) {
            mbar_init(&bar, 1);
            fence_barrier_init();
        }
        tile_setup(a, tile_idx, last, sh_ti);
        if (!DIRECT) {
            if (sh_ti.bytes) {
                mbar_arrive_expect_tx(&bar, sh_ti.bytes);
                bulk_g2s_stream(tile_s, a.in + sh_ti.b_lo, sh_ti.bytes, &bar);
            } else {
                mbar_arrive(&bar);
            }
        }
        tile_setup_calls(a, sh_ti);   // while the copy is in flight
    }
    __syncthreads();
    const TileInfo &ti = sh_ti;
    const unsigned char *tile = DIRECT ? a.in + ti.b_lo : tile_s;
    const unsigned long long wlo = ti.wlo, jlo = ti.jlo, jhi = ti.jhi;
    const uint32_t nlp = ti.nlp;
    const bool tile0 = wlo == 0;                      // this tile holds lowpassed window 0 / demod 0
    const uint32_t skip = (uint32_t)(jlo - wlo);      // 1 if element 0 is only the predecessor of demod jlo
    const uint32_t D = a.D;
    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(tile);

    int2 lastlp = make_int2(st.demod_pre_re, st.demod_pre_im);   // lp[nlp-1] for the carried state (last tile)
    if constexpr (DT == 6 || DIRECT || has_staged_pass(DT)) {
        // ---- D = 6 (optimal_settings :189-190): boxcar + discriminator fused, no window array ---------------
        if (!DIRECT) mbar_wait(&bar, parity);
        dn_pass<DT, NTH, DIRECT>(tile, ti, dm, tid);
        if (last && tid == NTH - 1) {
            int32_t re = 0, im = 0;
            boxcar_rot(w32, (int)ti.tail_from, (int)(ti.tail_from + ti.ntail), re, im);
            a.st_out->lp_now_re = re;
            a.st_out->lp_now_im = im;
        }
        if (last && tid == 0 && nlp) {
            const bool w0 = tile0 && nlp == 1;
            lastlp = lp_window_exact(w32, ti.off0, nlp - 1, D, w0 ? st.lp_now_re : 0, w0 ? st.lp_now_im : 0);
        }
        __syncthreads();
        if (tile0 || ti.clo <= ti.chi) {
            pass_fixups(a, st, ti, w32, dm, tid, blockDim.x);
            __syncthreads();
        }
    } else {
        // ---- first-of-call flags (while the bulk copy is in flight) --------------------------------
        {
            uint32_t *f32 = reinterpret_cast<uint32_t *>(flag);
            for (uint32_t i = tid; i < (nlp + 3) / 4; i += blockDim.x) f32[i] = 0;
        }
        __syncthreads();
        {
            const unsigned long long c_hi = ti.chi;
            for (unsigned long long c = ti.clo + tid; c <= c_hi; c += blockDim.x) {
                unsigned long long w = udiv64(c * a.S + a.p0, a.d64_D);
                if (w >= jlo && w < jhi) flag[w - wlo] = 1;
            }
        }

        // ---- phase 1: rotate_90 + centre + boxcar over D samples ------------------------------------
        mbar_wait(&bar, parity);
        const int32_t off0 = ti.off0;
        for (uint32_t i = tid; i < nlp; i += blockDim.x) {
            int32_t base = off0 + (int32_t)(i * D);
            int32_t re = 0, im = 0;
            if (tile0 && i == 0) {
                re = st.lp_now_re;
                im = st.lp_now_im;
            }
            // base < 0 only for window 0 with p0 > 0: those samples are already in lp_now
            boxcar_rot(w32, base < 0 ? 0 : base, base + (int32_t)D, re, im);
            lp[i] = make_int2(re, im);
        }
        if (last && tid == NTH - 1) {
            int32_t re = 0, im = 0;
            boxcar_rot(w32, (int)ti.tail_from, (int)(ti.tail_from + ti.ntail), re, im);
            a.st_out->lp_now_re = re;
            a.st_out->lp_now_im = im;
        }
        __syncthreads();

        // ---- phase 2: polar discriminator --------------------------------------------------------------
        for (uint32_t i = tid; i < nlp; i += blockDim.x) {
            if (i < skip) continue;   // the predecessor-only element
            int2 cur = lp[i];
            int2 prev = (tile0 && i == 0) ? make_int2(st.demod_pre_re, st.demod_pre_im) : lp[i - 1];
            int32_t cre, cim;
            d_cmul_conj(cur, prev, cre, cim);
            int32_t pcm = flag[i] ? d_polar_f64(cre, cim, a.oct) : d_fast_atan2(cim, cre);
            dm[i] = (int16_t)(uint16_t)(uint32_t)pcm;
        }
        __syncthreads();

        if (last && tid == 0 && nlp) lastlp = lp[nlp - 1];
    }

    // ---- phase 3: fractional boxcar resampler ------------------------------------------------------
    resample_tile(a, st, ti, dm, tid, blockDim.x);
    if (last && tid == 0) tile_state_out(a, st, ti, dm, lastlp);
    if (last) __threadfence();   // the state words are consumed by another CTA in ring mode
    __syncthreads();   // the shared-memory tile (and the mbarrier phase) may now be reused
}

// One launch per batch of calls: one CTA per tile.
// DT = compile-time downsample (0 = any): the reference's own setting, 6 (optimal_settings :189-190), is specialised.
#ifndef SDR_INT_MINB
#define SDR_INT_MINB 5
#endif
template <int DT>
__global__ void __launch_bounds__(256, SDR_INT_MINB) k_demod_fused(const FusedArgs a) {
    demod_tile<DT>(a, blockIdx.x, gridDim.x, 0, true);
}
// staged kernel of the odd downsamples from 15 up (dn_pass_odd_staged): a lane holds DT words
template <int DT>
__global__ void __launch_bounds__(256, DT <= 21 || !(DT & 1) ? 4 : 3) k_demod_staged_odd(const FusedArgs a) {
    demod_tile<DT>(a, blockIdx.x, gridDim.x, 0, true);
}
// CTAs per SM the register allocation aims for: 8 (32 registers) while a lane's row is at most 16 words, fewer for
// the wider rows of DT = 8 / 10 (a spilled row costs more than the lost occupancy)
#ifndef SDR_INT_DIRECT_MINB
#define SDR_INT_DIRECT_MINB 8
#endif
constexpr int direct_row_words(int DT) { return (DT & 1) ? 2 * DT : fused_pass_kw(DT) * DT / 2; }
// from 14 up the register budget follows the row a lane holds (wide rows: fewer, fatter CTAs — the per-sample instruction
// count falls with the downsample, so fewer warps keep the issue slots and the memory pipe busy)
constexpr int direct_min_blocks(int DT) {
    return DT >= 14 ? (direct_row_words(DT) <= 12 ? 5 : direct_row_words(DT) <= 20 ? 4 : direct_row_words(DT) <= 32 ? 3 : 2)
                    : DT == 13 ? 4 : (DT == 10 || DT == 11) ? 5 : (DT == 8 || DT == 9) ? 6 : SDR_INT_DIRECT_MINB;
}
#ifndef SDR_INT_DIRECT_NTH
#define SDR_INT_DIRECT_NTH 256   // threads per CTA of the direct kernel (128: 0.420 vs 0.426 ms on cfg1, 64 and 512 slower)
#endif
constexpr int kDirectNth = SDR_INT_DIRECT_NTH;
template <int DT>
__global__ void __launch_bounds__(kDirectNth, direct_min_blocks(DT) * 256 / kDirectNth) k_demod_direct(const FusedArgs a) {
    demod_tile<DT, true, kDirectNth>(a, blockIdx.x, gridDim.x, 0, true);
}

// ================================================================================================
// Persistent ring: successive USB-sized buffers stream through ONE resident kernel, no relaunch.
//
// The host copies buffer k into device slot k % n_slots and then (stream-ordered) writes the doorbell
// seq_ready[slot] = k + 1.  Every CTA walks the buffers in order, spins on the doorbell, and takes the tiles
// t = blockIdx.x, + gridDim.x, ... of that buffer (each buffer is exactly one reference demodulate() call).
// The three data-dependent state words hop from the last tile of buffer k to tile 0 of buffer k+1 through
// `state_gen` (release/acquire); the index state is closed-form in k.  The CTA that finishes a buffer's last
// outstanding tile publishes seq_done[slot] = k + 1 in host-mapped memory, where the audio already is.
// ================================================================================================
struct RingCtl {                       // device memory
    unsigned int seq_ready[64];        // doorbells, written by the copy engine
    unsigned int done_tiles[64];       // tiles finished per slot
    unsigned int stop;                 // host sets 1 to retire the kernel
    unsigned int pad;
    unsigned long long state_gen;      // buffers whose final state has been published
    IntState states[66];               // states[k % (n_slots+1)] = state at the START of buffer k
};

struct RingArgs {
    FusedArgs proto;                   // constants (config, magics, tile geometry); per-buffer fields are filled in
    RingCtl *ctl;
    volatile unsigned int *seq_done;   // host-mapped [n_slots]
    const uint8_t *d_in;               // device slots [n_slots][buf_len]
    int16_t *h_out;                    // host-mapped audio slots [n_slots][out_stride]
    unsigned long long buf_len, out_stride, slot_stride;
    unsigned int n_slots;
    unsigned int p0, q0;               // index state at ring open
    UDiv64 d64_fast;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <int DT>
__global__ void __launch_bounds__(256) k_demod_ring(const RingArgs r) {
    __shared__ FusedArgs sh_a;
    __shared__ unsigned int sh_go, sh_ntiles;
    const int tid = threadIdx.x;
    const unsigned long long S = r.buf_len / 2, D = r.proto.D, fast = r.proto.fast, slow = r.proto.slow;
    uint32_t uses = 0;
    for (unsigned long long k = 0;; k++) {
        const unsigned int slot = (unsigned int)(k % r.n_slots);
        if (tid == 0) {
            // wait for the doorbell of buffer k (or for stop)
            unsigned int go = 2;
            while (go == 2) {
                if (ld_acquire_u32(&r.ctl->seq_ready[slot]) == (unsigned int)(k + 1)) go = 1;
                else if (ld_acquire_u32(&r.ctl->stop))
                    // the doorbell of a buffer committed just before close was written BEFORE stop (same copy stream):
                    // it may have landed between the two loads above, so look once more before retiring
                    go = ld_acquire_u32(&r.ctl->seq_ready[slot]) == (unsigned int)(k + 1) ? 1 : 0;
                else __nanosleep(100);
            }
            sh_go = go;
            if (go) {
                // closed-form index state at the start of buffer k
                const unsigned long long n0 = r.p0 + k * S;
                const unsigned long long Lstart = udiv64(n0, r.proto.d64_D);
                const unsigned long long pk = n0 - Lstart * D;
                const unsigned long long t0 = r.q0 + Lstart * slow;
                const unsigned long long qk = t0 - udiv64(t0, r.d64_fast) * fast;
                FusedArgs a = r.proto;
                a.in = r.d_in + (size_t)slot * r.slot_stride;   // 16-byte aligned device slots (bulk-copy source)
                a.out = r.h_out + (size_t)slot * r.out_stride;
                a.st_in = &r.ctl->states[k % (r.n_slots + 1)];
                a.st_out = &r.ctl->states[(k + 1) % (r.n_slots + 1)];
                a.n_samples = S;
                a.p0 = (uint32_t)pk;
                a.q0 = (uint32_t)qk;
                a.Ltot = udiv64(pk + S, r.proto.d64_D);
                a.Etot = udiv64(qk + a.Ltot * slow, r.d64_fast);
                sh_a = a;
                sh_ntiles = a.Etot ? (unsigned int)((a.Etot + a.EB - 1) / a.EB) : 1u;
            }
        }
        __syncthreads();
        if (!sh_go) return;
        const unsigned int n_tiles = sh_ntiles;
        for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            if (t == 0) {   // tile 0 consumes the state the previous buffer's last tile produced
                if (tid == 0)
                    while (ld_acquire_u64(&r.ctl->state_gen) < k) __nanosleep(50);
                __syncthreads();
            }
            demod_tile<DT>(sh_a, t, n_tiles, uses & 1, uses == 0);
            uses++;
            if (tid == 0) {
                if (t == n_tiles - 1) st_release_u64(&r.ctl->state_gen, k + 1);   // st_out written above (+ __syncthreads)
                __threadfence_system();                                              // audio stores reach host memory
                if (atomicAdd(&r.ctl->done_tiles[slot], 1u) == n_tiles - 1) {
                    r.ctl->done_tiles[slot] = 0;
                    __threadfence_system();
                    r.seq_done[slot] = (unsigned int)(k + 1);
                }
            }
        }
        __syncthreads();   // sh_a is rewritten for the next buffer
    }
}

// ================================================================================================
// Stage kernels (stage-level parity against the reference's known-answer tests)
// ================================================================================================

// Demod::rotate_90 scalar branch, :281-298: one 8-byte group per thread, two PRMTs.
__global__ void k_rotate_90(uint2 *buf, size_t n_groups) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        uint2 w = buf[g];
        uint2 o;
        o.x = __byte_perm(w.x, ~w.x, 0x2710);   // [b0, b1, 255-b3, b2]
        o.y = __byte_perm(w.y, ~w.y, 0x6354);   // [255-b4, 255-b5, b7, 255-b6]
        buf[g] = o;
    }
}

// `*val as i16 - 127` (:258) + buf_to_complex (:441-450): u8 pairs -> Complex<i32>
__global__ void k_buf_to_complex(const uint16_t *in, size_t n_pairs, int2 *out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
        uint32_t v = in[i];
        out[i] = make_int2((int32_t)(v & 255u) - 127, (int32_t)(v >> 8) - 127);
    }
}

// Demod::low_pass_complex, :337-352.  Thread w < L sums its window; thread L sums the tail into st.
__global__ void k_low_pass_complex(const int2 *in, unsigned long long n, uint32_t D, uint32_t p0,
                                   unsigned long long L, const IntState *st_in, IntState *st_out, int2 *out) {
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w <= L; w += stride) {
        long long s0 = (long long)(w * D) - (long long)p0;
        long long s1 = (w < L) ? s0 + D : (long long)n;
        int32_t re = 0, im = 0;
        if (w == 0) {
            re = st_in->lp_now_re;
            im = st_in->lp_now_im;
        }
        for (long long s = s0 < 0 ? 0 : s0; s < s1; s++) {
            int2 v = in[s];
            re = wadd(re, v.x);
            im = wadd(im, v.y);
        }
        if (w < L) {
            out[w] = make_int2(re, im);
        } else {
            st_out->lp_now_re = re;
            st_out->lp_now_im = im;
        }
    }
}

// Demod::fm_demod, :355-367
__global__ void k_fm_demod(const int2 *in, unsigned long long n, const IntState *st_in, IntState *st_out,
                           OctTable oct, int16_t *out) {
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int2 cur = in[i];
        int2 prev = i ? in[i - 1] : make_int2(st_in->demod_pre_re, st_in->demod_pre_im);
        int32_t cre, cim;
        d_cmul_conj(cur, prev, cre, cim);
        int32_t pcm = i ? d_fast_atan2(cim, cre) : d_polar_f64(cre, cim, oct);
        out[i] = (int16_t)(uint16_t)(uint32_t)pcm;
        if (i == n - 1) {
            st_out->demod_pre_re = cur.x;
            st_out->demod_pre_im = cur.y;
        }
    }
}

// Demod::low_pass_real, :408-426.  Thread e < E emits output e; thread E sums the tail.
__global__ void k_low_pass_real(const int16_t *in, unsigned long long n, uint32_t q0, uint32_t fast,
                                uint32_t slow, int32_t div, unsigned long long E, const IntState *st_in,
                                IntState *st_out, int16_t *out) {
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e <= E; e += stride) {
        unsigned long long j0 = e ? (e * fast - q0 + slow - 1) / slow : 0ull;
        unsigned long long j1 = (e < E) ? ((e + 1) * fast - q0 + slow - 1) / slow : n;
        int32_t sum = (e == 0) ? st_in->now_lpr : 0;
        for (unsigned long long j = j0; j < j1; j++) sum = wadd(sum, (int32_t)in[j]);
        if (e < E)
            out[e] = (int16_t)(uint16_t)(uint32_t)tdiv(sum, div);
        else
            st_out->now_lpr = sum;
    }
}

__global__ void k_fast_atan2_v(const int32_t *y, const int32_t *x, size_t n, int32_t *out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        // the BOUNDED form (used by the D = 6 pass) is exercised on its whole domain, the general one elsewhere
        const int64_t mag = llabs((long long)y[i]) + llabs((long long)x[i]);
        out[i] = mag < (1ll << 30) ? d_fast_atan2_t<true>(y[i], x[i]) : d_fast_atan2(y[i], x[i]);
    }
}

__global__ void k_polar_v(const int2 *a, const int2 *b, size_t n, int fast, OctTable oct, int32_t *out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t cre, cim;
        d_cmul_conj(a[i], b[i], cre, cim);
        out[i] = fast ? d_fast_atan2(cim, cre) : d_polar_f64(cre, cim, oct);
    }
}

#define SDR_K(x) ((const void *)(x))
static const KernelList kIntKernels{
    SDR_K(k_demod_fused<0>), SDR_K(k_demod_fused<6>), SDR_K(k_demod_ring<0>), SDR_K(k_demod_ring<6>),
    SDR_K(k_demod_staged_odd<15>), SDR_K(k_demod_staged_odd<17>), SDR_K(k_demod_staged_odd<19>), SDR_K(k_demod_staged_odd<21>),
    SDR_K(k_demod_staged_odd<23>), SDR_K(k_demod_staged_odd<25>), SDR_K(k_demod_staged_odd<27>), SDR_K(k_demod_staged_odd<29>),
    SDR_K(k_demod_staged_odd<31>), SDR_K(k_demod_staged_odd<14>), SDR_K(k_demod_staged_odd<18>), SDR_K(k_demod_staged_odd<22>),
    SDR_K(k_demod_staged_odd<26>), SDR_K(k_demod_staged_odd<30>),
    SDR_K(k_demod_direct<2>), SDR_K(k_demod_direct<3>), SDR_K(k_demod_direct<4>), SDR_K(k_demod_direct<5>),
    SDR_K(k_demod_direct<6>), SDR_K(k_demod_direct<7>), SDR_K(k_demod_direct<8>), SDR_K(k_demod_direct<9>),
    SDR_K(k_demod_direct<10>), SDR_K(k_demod_direct<11>), SDR_K(k_demod_direct<12>), SDR_K(k_demod_direct<13>),
    SDR_K(k_demod_direct<14>), SDR_K(k_demod_direct<15>), SDR_K(k_demod_direct<16>), SDR_K(k_demod_direct<17>),
    SDR_K(k_demod_direct<18>), SDR_K(k_demod_direct<19>), SDR_K(k_demod_direct<20>), SDR_K(k_demod_direct<21>),
    SDR_K(k_demod_direct<22>), SDR_K(k_demod_direct<23>), SDR_K(k_demod_direct<24>), SDR_K(k_demod_direct<25>),
    SDR_K(k_demod_direct<26>), SDR_K(k_demod_direct<27>), SDR_K(k_demod_direct<28>), SDR_K(k_demod_direct<29>),
    SDR_K(k_demod_direct<30>), SDR_K(k_demod_direct<31>), SDR_K(k_demod_direct<32>),
    SDR_K(k_rotate_90), SDR_K(k_buf_to_complex), SDR_K(k_low_pass_complex), SDR_K(k_fm_demod), SDR_K(k_low_pass_real),
    SDR_K(k_fast_atan2_v), SDR_K(k_polar_v)};
#undef SDR_K

}  // namespace sdr

using namespace sdr;

// ================================================================================================
// Host side
// ================================================================================================
// Tile geometry: EB audio outputs per tile and the shared-memory layout that follows from it.
struct Geom {
    uint32_t EB = 0, lp_cap = 0, tile_cap = 0, dm_off = 0;
    size_t smem_staged = 0;   // raw tile | [generic D: lowpassed windows (int2)] | demodulated samples (i16) | [generic D: flags]
    size_t smem_direct = 0;   // demodulated samples only (D = 6 direct kernel)
};

// n_lp_target = lowpassed windows per tile to aim for.  False: rate_out too large for the kernel's 32-bit relative math.
// fused_layout: the staged D = 6 kernel keeps no window array and no flags in shared memory.
static bool make_geom(uint64_t D, uint64_t fast, uint64_t slow, uint64_t n_lp_target, Geom &g, bool fused_layout) {
    if (n_lp_target < 4) n_lp_target = 4;
    if (n_lp_target > 8192) n_lp_target = 8192;
    uint64_t EB = n_lp_target * slow / fast;
    if (EB < 1) EB = 1;
    if (EB > 4096) EB = 4096;
    if (EB >= 8) EB &= ~3ull;   // tiles start on 8-byte boundaries of the i16 output
    // the kernel's relative index math and its magic division need (EB+1)*fast + slow < 2^31
    while (EB > 1 && (EB + 1) * fast + slow >= (1ull << 31)) EB /= 2;
    if ((EB + 1) * fast + slow >= (1ull << 31)) return false;
    const bool d6 = fused_layout;
    const uint64_t per = (fast + slow - 1) / slow;
    const uint64_t lp_cap = (EB * fast + slow - 1) / slow + per + 4;
    // slack: pad in front (16 B for D = 6, 128 B for the staged odd pass), bulk-copy rounding, chunk over-read
    const uint64_t tile_cap = ((2 * (lp_cap * D + D) + 15) & ~15ull) + 96 + (fused_layout ? 128 : 0);
    g.EB = (uint32_t)EB;
    g.lp_cap = (uint32_t)lp_cap;
    g.tile_cap = (uint32_t)tile_cap;
    g.dm_off = (uint32_t)(d6 ? tile_cap : tile_cap + lp_cap * 8);
    // + 80: slack after the demodulated samples (whole-group stores of the D = 6 pass)
    g.smem_staged = (size_t)(d6 ? tile_cap + ((lp_cap * 2 + 15) & ~15ull) + 80
                                : tile_cap + ((lp_cap * 10 + 15) & ~15ull) + ((lp_cap + 19) & ~15ull) + 80);
    g.smem_direct = (size_t)(((lp_cap * 2 + 15) & ~15ull) + 80);
    return true;
}

struct sdr_demod {
    sdr_demod_config cfg{};
    int device = 0;
    sdr_demod_state st{};
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_h2d[2]{}, ev_done[2]{}, ev_t0 = nullptr, ev_t1 = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
    DevBuf d_in[2], d_out[2], d_state, d_a, d_b, d_c;
    PinBuf h_state;
    // small synchronous calls (one USB buffer per demodulate(), the reference's call pattern): pinned staging both
    // ways, ONE H2D (input [+ state]), one launch, ONE D2H (audio + state), one stream.  The carried state stays on
    // the device between such calls (slot of d_small_out that holds it, -1 = the host copy is the newer one).
    PinBuf h_small_in, h_small_out;
    DevBuf d_small_in, d_small_out[2];
    int dev_state_slot = -1;
    OctTable oct{};
    Geom geo;                // staged-tile geometry of the handle's usual kernel (k_demod_fused<6> for D = 6, else <0>)
    Geom geo_gen;            // D = 6 only: geometry of the generic kernel, which takes the odd window starts
    Geom geo_odd;            // odd D from 15 up: tiles of the staged register-resident pass (k_demod_staged_odd)
    bool staged_odd = false;
    Geom geo_ring;           // the persistent ring's tiles (D = 6: SDR_RING_PASSES passes, default 1 — one USB buffer is
                             // only ~21845 windows, small tiles spread it over more CTAs)
    Geom geo_direct[4];      // direct kernel: tiles of 1, 2, 4, 8 passes (256 lanes x KW windows each); the launch picks by batch size
    int n_direct = 0;        // 0: direct kernel disabled (SDR_INT_DIRECT=0 or a downsample without the register-resident pass)
    size_t smem_bytes = 0;
    bool pending = false;      // async *_dev submission whose state has not been committed yet
    int pending_slot = 0;
    float last_ms = 0.f;
    uint32_t last_launches = 0;
    bool timing_valid = false;
    H2DStager stager;          // pageable caller buffers go through pinned pieces (common.cuh)
    bool ring_open = false;    // a persistent ring owns the handle until sdr_ring_close()
    struct sdr_ring *ring = nullptr;   // that ring (sdr_demod_free closes it first)
};

namespace {

constexpr size_t kChunkBytes = size_t(32) << 20;

inline int grid_for(size_t n, int device, int per_block = 256) {
    size_t b = (n + per_block - 1) / per_block;
    size_t cap = (size_t)sm_count(device) * 16;
    if (b < 1) b = 1;
    return (int)(b < cap ? b : cap);
}

void fill_oct_table(OctTable &t) {
    // (angle / PI * (1 << 14) as f64) as i32 — examples/simple_fm.rs:372-373, platform libm
    const double PI = 3.14159265358979323846264338327950288;
    const int ys[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    const int xs[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    for (int i = 0; i < 8; i++) t.v[i] = (int32_t)(std::atan2((double)ys[i], (double)xs[i]) / PI * 16384.0);
    t.v[8] = (int32_t)(std::atan2(0.0, 0.0) / PI * 16384.0);
}

void to_dev_state(const sdr_demod_state &s, IntState &d) {
    memset(&d, 0, sizeof(d));
    d.lp_now_re = s.lp_now_re;
    d.lp_now_im = s.lp_now_im;
    d.demod_pre_re = s.demod_pre_re;
    d.demod_pre_im = s.demod_pre_im;
    d.now_lpr = s.now_lpr;
}

// closed-form bookkeeping of a run of `n_bufs` calls of S samples from index state (p0, q0)
struct Plan {
    uint64_t n_samples, Ltot, Etot;
    uint32_t p1, q1;
};
Plan make_plan(const sdr_demod_config &c, uint32_t p0, uint32_t q0, uint64_t S, uint64_t n_bufs) {
    Plan p;
    p.n_samples = S * n_bufs;
    p.Ltot = (p0 + p.n_samples) / c.downsample;
    p.p1 = (uint32_t)((p0 + p.n_samples) % c.downsample);
    unsigned __int128 t = (unsigned __int128)p.Ltot * c.rate_resample + q0;
    p.Etot = (uint64_t)(t / c.rate_out);
    p.q1 = (uint32_t)(t % c.rate_out);
    return p;
}

int validate_lens(const sdr_demod *d, size_t buf_len, size_t n_bufs) {
    if (buf_len == 0 || n_bufs == 0) return fail(SDR_E_LEN, "empty buffer (the reference asserts >1 lowpassed samples, examples/simple_fm.rs:356)");
    if (buf_len % 8 != 0)
        return fail(SDR_E_LEN, "buffer length %zu is not a multiple of 8 (rotate_90 indexes i+7, examples/simple_fm.rs:284-295)", buf_len);
    const uint64_t S = buf_len / 2, D = d->cfg.downsample;
    uint64_t p = d->st.prev_index;
    size_t lim = n_bufs < D ? n_bufs : (size_t)D;   // p_c is periodic with period <= D
    for (size_t c = 0; c < lim; c++) {
        if ((p + S) / D < 2)
            return fail(SDR_E_LEN, "call %zu would produce %llu lowpassed samples; fm_demod asserts len > 1 (examples/simple_fm.rs:356)", c,
                        (unsigned long long)((p + S) / D));
        p = (p + S) % D;
    }
    return SDR_OK;
}

int commit_pending(sdr_demod *d) {
    if (d->ring_open) return fail(SDR_E_STATE, "handle is owned by an open ring (sdr_ring_close first)");
    if (!d->pending) return SDR_OK;
    SDR_CUDA_TRY(cudaStreamSynchronize(d->stream));
    const IntState *h = d->h_state.as<IntState>();
    d->st.lp_now_re = h->lp_now_re;
    d->st.lp_now_im = h->lp_now_im;
    d->st.demod_pre_re = h->demod_pre_re;
    d->st.demod_pre_im = h->demod_pre_im;
    d->st.now_lpr = h->now_lpr;
    d->pending = false;
    if (d->timing_valid) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, d->ev_t0, d->ev_t1) == cudaSuccess) d->last_ms = ms;
    }
    return SDR_OK;
}

// Round-up magics (Granlund-Montgomery): q = umulhi(n, m) >> shift, exact for n < 2^31 (32-bit) / n < 2^63 (64-bit).
UDiv magic32(uint32_t dv) {
    if (dv <= 1) return UDiv{0u, 0xffffffffu};
    uint32_t s = 0;
    while ((1ull << s) < dv) s++;                       // s = ceil(log2 dv) >= 1
    uint64_t m = ((1ull << (31 + s)) / dv) + 1;          // < 2^32 because dv > 2^(s-1)
    return UDiv{(uint32_t)m, s - 1};
}
UDiv64 magic64(uint64_t dv) {
    if (dv <= 1) return UDiv64{0ull, 0xffffffffu};
    uint32_t sh = 0;
    while (sh < 63 && (1ull << sh) < dv) sh++;
    unsigned __int128 m = (((unsigned __int128)1 << (63 + sh)) / dv) + 1;   // < 2^64 because dv > 2^(sh-1)
    return UDiv64{(unsigned long long)m, sh - 1};
}

// Launch the fused kernel for one chunk that starts on a call boundary.
int launch_fused(sdr_demod *d, const uint8_t *d_in, uint64_t S, uint64_t n_calls, uint32_t p0, uint32_t q0,
                 const Plan &pl, const IntState *st_in, IntState *st_out, int16_t *d_out) {
    FusedArgs a{};
    a.in = d_in;
    a.out = d_out;
    a.st_in = st_in;
    a.st_out = st_out;
    a.n_samples = pl.n_samples;
    a.Ltot = pl.Ltot;
    a.Etot = pl.Etot;
    a.S = (uint32_t)S;
    a.D = d->cfg.downsample;
    a.p0 = p0;
    a.q0 = q0;
    a.fast = d->cfg.rate_out;
    a.slow = d->cfg.rate_resample;
    a.div = (int32_t)(d->cfg.rate_out / d->cfg.rate_resample);
    a.div_slow = magic32(d->cfg.rate_resample);
    a.div_audio = magic32(d->cfg.rate_out / d->cfg.rate_resample);
    a.d64_slow = magic64(d->cfg.rate_resample);
    a.d64_S = magic64(S);
    a.d64_D = magic64(d->cfg.downsample);
    a.oct = d->oct;
    (void)n_calls;
    // D = 6 with an even window start (every stream that was not given an odd prev_index by hand): direct kernel, with
    // the largest tile that still leaves `kDirectWaves` full waves of 8 CTAs per SM (SDR_INT_DIRECT_PASSES pins it)
    const bool odd_start_d6 = d->cfg.downsample == 6 && (p0 & 1);   // prev_index set by hand: generic kernel
    const bool staged = d->staged_odd && ((d->cfg.downsample & 1) || !(p0 & 1));   // even D: even window starts only
    const Geom *g = odd_start_d6 ? &d->geo_gen : staged ? &d->geo_odd : &d->geo;
    const bool direct = d->n_direct > 0 && ((d->cfg.downsample & 1) || !(p0 & 1));   // odd downsamples take any window start
    if (direct) {
        constexpr uint64_t kDirectWaves = 4;
        int k = d->n_direct - 1;
        const char *ep = getenv("SDR_INT_DIRECT_PASSES");
        if (ep) {
            const int want = atoi(ep);
            k = want >= 8 ? 3 : want >= 4 ? 2 : want >= 2 ? 1 : 0;
            if (k > d->n_direct - 1) k = d->n_direct - 1;
        } else {
            while (k > 0 && pl.Etot / d->geo_direct[k].EB < kDirectWaves * 8 * (uint64_t)sm_count(d->device)) k--;
        }
        g = &d->geo_direct[k];
    }
    a.EB = g->EB;
    a.lp_cap = g->lp_cap;
    a.tile_cap = g->tile_cap;
    a.dm_off = g->dm_off;
    uint64_t blocks = pl.Etot ? (pl.Etot + g->EB - 1) / g->EB : 1;
    if (blocks > 0x7fffffffull) return fail(SDR_E_ARG, "batch too large for one launch");
    if (direct) {
#define SDR_DIRECT_CASE(DT_) \
    case DT_: k_demod_direct<DT_><<<(unsigned)blocks, kDirectNth, g->smem_direct, d->stream>>>(a); break;
        switch (d->cfg.downsample) {
            SDR_DIRECT_CASE(2) SDR_DIRECT_CASE(3) SDR_DIRECT_CASE(4) SDR_DIRECT_CASE(5) SDR_DIRECT_CASE(6) SDR_DIRECT_CASE(7)
            SDR_DIRECT_CASE(8) SDR_DIRECT_CASE(9) SDR_DIRECT_CASE(10) SDR_DIRECT_CASE(11) SDR_DIRECT_CASE(12) SDR_DIRECT_CASE(13)
            SDR_DIRECT_CASE(14) SDR_DIRECT_CASE(15) SDR_DIRECT_CASE(16) SDR_DIRECT_CASE(17) SDR_DIRECT_CASE(18) SDR_DIRECT_CASE(19)
            SDR_DIRECT_CASE(20) SDR_DIRECT_CASE(21) SDR_DIRECT_CASE(22) SDR_DIRECT_CASE(23) SDR_DIRECT_CASE(24) SDR_DIRECT_CASE(25)
            SDR_DIRECT_CASE(26) SDR_DIRECT_CASE(27) SDR_DIRECT_CASE(28) SDR_DIRECT_CASE(29) SDR_DIRECT_CASE(30) SDR_DIRECT_CASE(31)
            default: k_demod_direct<32><<<(unsigned)blocks, kDirectNth, g->smem_direct, d->stream>>>(a); break;
        }
#undef SDR_DIRECT_CASE
    } else if (staged) {
#define SDR_ODD_CASE(DT_) \
    case DT_: k_demod_staged_odd<DT_><<<(unsigned)blocks, 256, d->smem_bytes, d->stream>>>(a); break;
        switch (d->cfg.downsample) {
            SDR_ODD_CASE(15) SDR_ODD_CASE(17) SDR_ODD_CASE(19) SDR_ODD_CASE(21) SDR_ODD_CASE(23) SDR_ODD_CASE(25) SDR_ODD_CASE(27)
            SDR_ODD_CASE(29) SDR_ODD_CASE(31) SDR_ODD_CASE(14) SDR_ODD_CASE(18) SDR_ODD_CASE(22) SDR_ODD_CASE(26) SDR_ODD_CASE(30)
        }
#undef SDR_ODD_CASE
    } else if (d->cfg.downsample == 6 && !(p0 & 1))
        k_demod_fused<6><<<(unsigned)blocks, 256, d->smem_bytes, d->stream>>>(a);
    else
        k_demod_fused<0><<<(unsigned)blocks, 256, d->smem_bytes, d->stream>>>(a);
    SDR_LAUNCH_CHECK();
    d->last_launches++;
    return SDR_OK;
}

constexpr size_t kSmallCallBytes = size_t(1) << 20;

// One synchronous call of n_bufs buffers whose bytes fit in kSmallCallBytes.  `total` = the plan of the whole call.
int demodulate_small(sdr_demod *d, const uint8_t *buf, size_t buf_len, size_t n_bufs, const Plan &total, int16_t *out) {
    const size_t in_bytes = buf_len * n_bufs;
    const size_t in_pad = (in_bytes + 255) & ~size_t(255);          // state rides behind the input, clear of over-reads
    const size_t out_cap_b = (kSmallCallBytes / 2 / std::max<uint32_t>(1, d->cfg.downsample) + 64) * sizeof(int16_t);
    const size_t st_off = (out_cap_b + 15) & ~size_t(15);            // fixed place of the state behind the audio
    int rc;
    if ((rc = d->h_small_in.reserve(kSmallCallBytes + 512)) || (rc = d->h_small_out.reserve(st_off + sizeof(IntState))) ||
        (rc = d->d_small_in.reserve(kSmallCallBytes + 512)))
        return rc;
    for (int i = 0; i < 2; i++)
        if ((rc = d->d_small_out[i].reserve(st_off + sizeof(IntState)))) return rc;
    if (total.Etot * sizeof(int16_t) > out_cap_b) return SDR_E_STATE;   // caller falls back to the general path
    uint8_t *hin = d->h_small_in.as<uint8_t>();
    memcpy(hin, buf, in_bytes);
    size_t h2d = in_bytes;
    const IntState *st_in;
    if (d->dev_state_slot < 0) {
        to_dev_state(d->st, *reinterpret_cast<IntState *>(hin + in_pad));
        h2d = in_pad + sizeof(IntState);
        st_in = reinterpret_cast<const IntState *>(d->d_small_in.as<uint8_t>() + in_pad);
    } else {
        st_in = reinterpret_cast<const IntState *>(d->d_small_out[d->dev_state_slot].as<uint8_t>() + st_off);
    }
    const int nslot = d->dev_state_slot < 0 ? 0 : d->dev_state_slot ^ 1;
    uint8_t *dout = d->d_small_out[nslot].as<uint8_t>();
    d->dev_state_slot = -1;   // until the call has completed
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_small_in.p, hin, h2d, cudaMemcpyHostToDevice, d->stream));
    d->last_launches = 0;
    d->timing_valid = false;
    if ((rc = launch_fused(d, d->d_small_in.as<uint8_t>(), buf_len / 2, n_bufs, (uint32_t)d->st.prev_index,
                           (uint32_t)d->st.prev_lpr_index, total, st_in, reinterpret_cast<IntState *>(dout + st_off),
                           reinterpret_cast<int16_t *>(dout))))
        return rc;
    // audio and state in one copy: [0, Etot*2) and [st_off, st_off + 32); the gap between them is a few KB at most
    const size_t audio_b = (total.Etot * sizeof(int16_t) + 15) & ~size_t(15);
    uint8_t *hout = d->h_small_out.as<uint8_t>();
    if (st_off - audio_b <= 16384) {
        SDR_CUDA_TRY(cudaMemcpyAsync(hout, dout, st_off + sizeof(IntState), cudaMemcpyDeviceToHost, d->stream));
    } else {
        if (audio_b) SDR_CUDA_TRY(cudaMemcpyAsync(hout, dout, audio_b, cudaMemcpyDeviceToHost, d->stream));
        SDR_CUDA_TRY(cudaMemcpyAsync(hout + st_off, dout + st_off, sizeof(IntState), cudaMemcpyDeviceToHost, d->stream));
    }
    SDR_CUDA_TRY(cudaStreamSynchronize(d->stream));
    memcpy(out, hout, total.Etot * sizeof(int16_t));
    const IntState *h = reinterpret_cast<const IntState *>(hout + st_off);
    d->st.lp_now_re = h->lp_now_re;
    d->st.lp_now_im = h->lp_now_im;
    d->st.demod_pre_re = h->demod_pre_re;
    d->st.demod_pre_im = h->demod_pre_im;
    d->st.now_lpr = h->now_lpr;
    d->st.prev_index = total.p1;
    d->st.prev_lpr_index = (int32_t)total.q1;
    d->dev_state_slot = nslot;
    return SDR_OK;
}

}  // namespace

extern "C" {

int sdr_optimal_settings(uint32_t freq, uint32_t rate, uint32_t sample_rate, uint32_t rate_resample,
                         sdr_radio_config *radio, sdr_demod_config *demod) {
    if (rate == 0) return fail(SDR_E_ARG, "rate must be non-zero");
    uint32_t downsample = (1000000u / rate) + 1u;           // examples/simple_fm.rs:190
    uint32_t capture_rate = downsample * rate;               // :192
    uint32_t capture_freq = freq + capture_rate / 4u;        // :195
    uint32_t output_scale = (1u << 15) / (128u * downsample);  // :197
    if (output_scale < 1u) output_scale = 1u;                // :198-200
    if (radio) {
        radio->capture_freq = capture_freq;
        radio->capture_rate = capture_rate;
    }
    if (demod) {
        demod->rate_in = sample_rate;                        // :207
        demod->rate_out = sample_rate;                       // :208
        demod->rate_resample = rate_resample;                // :209
        demod->downsample = downsample;
        demod->output_scale = output_scale;
    }
    return SDR_OK;
}

int sdr_demod_new(const sdr_demod_config *cfg, int cuda_device, sdr_demod **out) {
    if (!cfg || !out) return fail(SDR_E_ARG, "sdr_demod_new: null argument");
    if (cfg->downsample < 1) return fail(SDR_E_ARG, "downsample must be >= 1");
    if (cfg->rate_resample < 1 || cfg->rate_resample > cfg->rate_out || cfg->rate_out >= (1u << 31))
        return fail(SDR_E_ARG, "need 1 <= rate_resample <= rate_out < 2^31 (low_pass_real divides by rate_out/rate_resample, examples/simple_fm.rs:421)");
    int rc = use_device(cuda_device);
    if (rc) return rc;
    sdr_demod *d = new sdr_demod();
    d->cfg = *cfg;
    d->device = cuda_device;
    fill_oct_table(d->oct);
    // tile geometry.  Generic D: ~16 KB of raw bytes per CTA.  D = 6: 256 lanes x 4 windows per pass; the staged form
    // (ring, odd window starts) uses SDR_INT_PASSES passes per tile (3: five CTAs per SM), the direct batch kernel picks
    // 1, 2, 4 or 8 passes per tile at launch time from the batch size (large tiles amortise the per-tile code, small
    // ones keep a single 262144-byte call spread over many SMs).
    const uint64_t D = cfg->downsample, fast = cfg->rate_out, slow = cfg->rate_resample;
    const bool d6 = D == 6;
    uint64_t n_lp_target = 8192 / D;
    if (d6) {
        const char *env = getenv("SDR_INT_PASSES");
        int passes = env ? atoi(env) : 3;
        if (passes < 1 || passes > 8) passes = 3;
        n_lp_target = 1024ull * passes - 2;
    }
    if (!make_geom(D, fast, slow, n_lp_target, d->geo, d6) || !make_geom(D, fast, slow, 8192 / D, d->geo_gen, false)) {
        delete d;
        return fail(SDR_E_ARG, "rate_out too large");
    }
    d->geo_ring = d->geo;
    if (d6) {
        const char *er = getenv("SDR_RING_PASSES");
        int rp = er ? atoi(er) : 1;
        if (rp < 1 || rp > 8) rp = 1;
        if (!make_geom(D, fast, slow, 1024ull * rp - 2, d->geo_ring, true)) d->geo_ring = d->geo;
    }
    {
        const char *eo = getenv("SDR_INT_STAGED_ODD");
        if (has_staged_pass((int)D) && !(eo && atoi(eo) == 0)) {
            const char *ep = getenv("SDR_INT_ODD_PASSES");
            int op = ep ? atoi(ep) : 3;   // odd: 512 windows per pass (one pair per lane); even: 256 (one window per lane)   // 512 windows (one pair per lane) per pass; measured D = 15 / 21 / 31: 3.2 / 4.1 / 4.4 TB/s
                                          // with 1 pass, 5.1 / 5.4 / 5.3 with 3, 5.0 / 4.7 / 3.6 with 4
            if (op < 1 || op > 8) op = 1;
            d->staged_odd = make_geom(D, fast, slow, 512ull * op - 2, d->geo_odd, true);
        }
    }
    const size_t smem = std::max(std::max(std::max(d->geo.smem_staged, d->geo_gen.smem_staged), d->geo_ring.smem_staged),
                                 d->staged_odd ? d->geo_odd.smem_staged : 0);
    if (smem > 200 * 1024) {
        delete d;
        return fail(SDR_E_ARG, "downsample %u too large for the fused kernel's shared-memory tile", cfg->downsample);
    }
    d->smem_bytes = smem;
    {
        const char *ed = getenv("SDR_INT_DIRECT");
        if (has_fused_pass((int)D) && !d->staged_odd && !(ed && atoi(ed) == 0)) {
            const uint64_t kw = (uint64_t)fused_pass_kw((int)D);   // PassGeom<D>::KW: windows per lane
            for (int k = 0; k < 4; k++)
                if (make_geom(D, fast, slow, ((kDirectNth * kw) << k) - 2, d->geo_direct[k], true) && d->geo_direct[k].smem_direct <= 48 * 1024)
                    d->n_direct = k + 1;
                else
                    break;
        }
    }
    cudaError_t e = raise_dyn_smem(k_demod_fused<0>, smem);
    if (e == cudaSuccess) e = raise_dyn_smem(k_demod_fused<6>, smem);
    if (e == cudaSuccess && d->staged_odd) {
#define SDR_ODD_CASE(DT_) case DT_: e = raise_dyn_smem(k_demod_staged_odd<DT_>, smem); break;
        switch (D) {
            SDR_ODD_CASE(15) SDR_ODD_CASE(17) SDR_ODD_CASE(19) SDR_ODD_CASE(21) SDR_ODD_CASE(23) SDR_ODD_CASE(25) SDR_ODD_CASE(27)
            SDR_ODD_CASE(29) SDR_ODD_CASE(31) SDR_ODD_CASE(14) SDR_ODD_CASE(18) SDR_ODD_CASE(22) SDR_ODD_CASE(26) SDR_ODD_CASE(30)
        }
#undef SDR_ODD_CASE
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&d->ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&d->ev_t0);
    if (e == cudaSuccess) e = cudaEventCreate(&d->ev_t1);
    if (e == cudaSuccess) e = cudaEventCreate(&d->ev_s0);
    if (e == cudaSuccess) e = cudaEventCreate(&d->ev_s1);
    if (e != cudaSuccess) {
        sdr_demod_free(d);
        return fail(SDR_E_CUDA, "sdr_demod_new: %s", cudaGetErrorString(e));
    }
    if ((rc = d->d_state.reserve(4 * sizeof(IntState))) || (rc = d->h_state.reserve(4 * sizeof(IntState)))) {
        sdr_demod_free(d);
        return rc;
    }
    *out = d;
    return SDR_OK;
}

void sdr_demod_free(sdr_demod *d) {
    if (!d) return;
    cudaSetDevice(d->device);
    // a ring that is still open (Python GC order, a forgotten close) would keep its kernel resident for ever and leave
    // a dangling back-pointer: retire it first
    if (d->ring_open && d->ring) sdr_ring_close(d->ring);
    if (d->stream) cudaStreamSynchronize(d->stream);
    if (d->copy_stream) cudaStreamSynchronize(d->copy_stream);
    for (int i = 0; i < 2; i++) {
        d->d_in[i].release();
        d->d_out[i].release();
        d->d_small_out[i].release();
        if (d->ev_h2d[i]) cudaEventDestroy(d->ev_h2d[i]);
        if (d->ev_done[i]) cudaEventDestroy(d->ev_done[i]);
    }
    d->stager.release();
    d->d_state.release();
    d->d_a.release();
    d->d_b.release();
    d->d_c.release();
    d->h_state.release();
    d->d_small_in.release();
    d->h_small_in.release();
    d->h_small_out.release();
    if (d->ev_t0) cudaEventDestroy(d->ev_t0);
    if (d->ev_t1) cudaEventDestroy(d->ev_t1);
    if (d->ev_s0) cudaEventDestroy(d->ev_s0);
    if (d->ev_s1) cudaEventDestroy(d->ev_s1);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    delete d;
}

int sdr_demod_get_state(const sdr_demod *d, sdr_demod_state *st) {
    if (!d || !st) return fail(SDR_E_ARG, "null argument");
    int rc = commit_pending(const_cast<sdr_demod *>(d));
    if (rc) return rc;
    *st = d->st;
    return SDR_OK;
}

int sdr_demod_set_state(sdr_demod *d, const sdr_demod_state *st) {
    if (!d || !st) return fail(SDR_E_ARG, "null argument");
    int rc = commit_pending(d);
    if (rc) return rc;
    if (st->prev_index >= d->cfg.downsample) return fail(SDR_E_ARG, "prev_index must be < downsample");
    if (st->prev_lpr_index < 0 || (uint32_t)st->prev_lpr_index >= d->cfg.rate_out)
        return fail(SDR_E_ARG, "prev_lpr_index must be in [0, rate_out)");
    d->st = *st;
    d->dev_state_slot = -1;
    return SDR_OK;
}

long sdr_demod_out_len(const sdr_demod *d, size_t len) {
    if (!d) return fail(SDR_E_ARG, "null handle");
    int rc = validate_lens(d, len, 1);
    if (rc) return rc;
    Plan p = make_plan(d->cfg, (uint32_t)d->st.prev_index, (uint32_t)d->st.prev_lpr_index, len / 2, 1);
    return (long)p.Etot;
}

long sdr_demod_demodulate_batch(sdr_demod *d, const uint8_t *buf, size_t buf_len, size_t n_bufs, int16_t *out,
                                size_t out_cap, uint32_t *out_lens) {
    if (!d || !buf || !out) return fail(SDR_E_ARG, "sdr_demod_demodulate: null argument");
    int rc = use_device(d->device);
    if (rc) return rc;
    if ((rc = commit_pending(d))) return rc;
    if ((rc = validate_lens(d, buf_len, n_bufs))) return rc;
    const uint64_t S = buf_len / 2;
    const uint32_t p0 = (uint32_t)d->st.prev_index, q0 = (uint32_t)d->st.prev_lpr_index;
    Plan total = make_plan(d->cfg, p0, q0, S, n_bufs);
    if (total.Etot > out_cap)
        return fail(SDR_E_CAP, "output capacity %zu < %llu audio samples", out_cap, (unsigned long long)total.Etot);
    if (out_lens) {
        uint32_t p = p0, q = q0;
        for (size_t c = 0; c < n_bufs; c++) {
            Plan one = make_plan(d->cfg, p, q, S, 1);
            out_lens[c] = (uint32_t)one.Etot;
            p = one.p1;
            q = one.q1;
        }
    }
    if (buf_len * n_bufs <= kSmallCallBytes && !getenv("SDR_INT_NO_SMALL_PATH")) {
        rc = demodulate_small(d, buf, buf_len, n_bufs, total, out);
        if (rc == SDR_OK) return (long)total.Etot;
        if (rc != SDR_E_STATE) return rc;   // SDR_E_STATE: does not fit the small buffers, take the general path
    }
    d->dev_state_slot = -1;   // the general path leaves the newest state on the host
    size_t per_chunk = kChunkBytes / buf_len;
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > n_bufs) per_chunk = n_bufs;
    const size_t chunk_bytes = per_chunk * buf_len;
    const size_t chunk_out_cap = (size_t)(make_plan(d->cfg, 0, d->cfg.rate_out - 1, S, per_chunk).Etot + 2);
    for (int i = 0; i < 2; i++) {
        if ((rc = d->d_in[i].reserve(chunk_bytes + 64))) return rc;
        if ((rc = d->d_out[i].reserve(chunk_out_cap * sizeof(int16_t)))) return rc;
    }
    IntState *dst = d->d_state.as<IntState>();
    IntState hs;
    to_dev_state(d->st, hs);
    *d->h_state.as<IntState>() = hs;
    SDR_CUDA_TRY(cudaMemcpyAsync(&dst[0], d->h_state.p, sizeof(IntState), cudaMemcpyHostToDevice, d->stream));
    d->last_launches = 0;
    d->timing_valid = false;

    uint32_t p = p0, q = q0;
    size_t done = 0, out_off = 0;
    int chunk = 0;
    while (done < n_bufs) {
        const int slot = chunk & 1;
        const size_t nb = (n_bufs - done < per_chunk) ? n_bufs - done : per_chunk;
        Plan pl = make_plan(d->cfg, p, q, S, nb);
        // the copy engine may not overwrite a slot until the kernel that read it has finished
        if (chunk >= 2) SDR_CUDA_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_done[slot], 0));
        if ((rc = d->stager.copy(d->d_in[slot].p, buf + done * buf_len, nb * buf_len, d->copy_stream))) return rc;
        SDR_CUDA_TRY(cudaEventRecord(d->ev_h2d[slot], d->copy_stream));
        SDR_CUDA_TRY(cudaStreamWaitEvent(d->stream, d->ev_h2d[slot], 0));
        if ((rc = launch_fused(d, d->d_in[slot].as<uint8_t>(), S, nb, p, q, pl, &dst[chunk & 1], &dst[(chunk + 1) & 1],
                               d->d_out[slot].as<int16_t>())))
            return rc;
        if (pl.Etot)
            SDR_CUDA_TRY(cudaMemcpyAsync(out + out_off, d->d_out[slot].p, pl.Etot * sizeof(int16_t),
                                         cudaMemcpyDeviceToHost, d->stream));
        SDR_CUDA_TRY(cudaEventRecord(d->ev_done[slot], d->stream));
        p = pl.p1;
        q = pl.q1;
        out_off += pl.Etot;
        done += nb;
        chunk++;
    }
    SDR_CUDA_TRY(cudaMemcpyAsync(d->h_state.p, &dst[chunk & 1], sizeof(IntState), cudaMemcpyDeviceToHost, d->stream));
    d->st.prev_index = p;
    d->st.prev_lpr_index = (int32_t)q;
    d->pending = true;
    if ((rc = commit_pending(d))) return rc;
    SDR_CUDA_TRY(cudaStreamSynchronize(d->copy_stream));
    return (long)total.Etot;
}

long sdr_demod_demodulate(sdr_demod *d, const uint8_t *buf, size_t len, int16_t *out, size_t out_cap) {
    return sdr_demod_demodulate_batch(d, buf, len, 1, out, out_cap, nullptr);
}

long sdr_demod_demodulate_batch_dev(sdr_demod *d, const uint8_t *d_buf, size_t buf_len, size_t n_bufs,
                                    int16_t *d_out, size_t out_cap) {
    if (!d || !d_buf || !d_out) return fail(SDR_E_ARG, "sdr_demod_demodulate_batch_dev: null argument");
    if (reinterpret_cast<uintptr_t>(d_buf) & 15) return fail(SDR_E_ARG, "device input must be 16-byte aligned (use sdr_dev_alloc)");
    int rc = use_device(d->device);
    if (rc) return rc;
    if ((rc = commit_pending(d))) return rc;
    if ((rc = validate_lens(d, buf_len, n_bufs))) return rc;
    const uint64_t S = buf_len / 2;
    const uint32_t p0 = (uint32_t)d->st.prev_index, q0 = (uint32_t)d->st.prev_lpr_index;
    Plan pl = make_plan(d->cfg, p0, q0, S, n_bufs);
    if (pl.Etot > out_cap)
        return fail(SDR_E_CAP, "output capacity %zu < %llu audio samples", out_cap, (unsigned long long)pl.Etot);
    d->dev_state_slot = -1;
    IntState *dst = d->d_state.as<IntState>();
    IntState hs;
    to_dev_state(d->st, hs);
    *d->h_state.as<IntState>() = hs;
    SDR_CUDA_TRY(cudaMemcpyAsync(&dst[0], d->h_state.p, sizeof(IntState), cudaMemcpyHostToDevice, d->stream));
    d->last_launches = 0;
    SDR_CUDA_TRY(cudaEventRecord(d->ev_t0, d->stream));
    if ((rc = launch_fused(d, d_buf, S, n_bufs, p0, q0, pl, &dst[0], &dst[1], d_out))) return rc;
    SDR_CUDA_TRY(cudaEventRecord(d->ev_t1, d->stream));
    d->timing_valid = true;
    SDR_CUDA_TRY(cudaMemcpyAsync(d->h_state.p, &dst[1], sizeof(IntState), cudaMemcpyDeviceToHost, d->stream));
    d->st.prev_index = pl.p1;
    d->st.prev_lpr_index = (int32_t)pl.q1;
    d->pending = true;
    return (long)pl.Etot;
}

int sdr_demod_sync(sdr_demod *d) {
    if (!d) return fail(SDR_E_ARG, "null handle");
    int rc = use_device(d->device);
    if (rc) return rc;
    if ((rc = commit_pending(d))) return rc;
    SDR_CUDA_TRY(cudaStreamSynchronize(d->stream));
    return SDR_OK;
}

int sdr_demod_last_timing(const sdr_demod *d, float *kernel_ms, uint32_t *n_launches) {
    if (!d) return fail(SDR_E_ARG, "null handle");
    if (kernel_ms) *kernel_ms = d->last_ms;
    if (n_launches) *n_launches = d->last_launches;
    return SDR_OK;
}


// ---- persistent ring (host side) --------------------------------------------------------------------------
struct sdr_ring {
    sdr_demod *d = nullptr;
    size_t buf_len = 0, out_stride = 0, slot_stride = 0;
    uint32_t n_slots = 0;
    uint8_t *h_in = nullptr;              // pinned [n_slots][buf_len]
    DevBuf d_in, d_ctl;
    int16_t *h_out = nullptr, *h_out_dev = nullptr;              // host-mapped audio slots
    volatile unsigned int *h_seq_done = nullptr;
    unsigned int *h_seq_done_dev = nullptr;
    unsigned int *h_doorbell = nullptr;   // pinned [n_slots] + stop word at [n_slots]
    cudaStream_t ring_stream = nullptr, copy_stream = nullptr;
    std::atomic<uint64_t> head{0}, tail{0};   // committed / collected buffers
    bool acquired = false;
    uint32_t p0 = 0, q0 = 0;
};

static void ring_release(sdr_ring *r) {
    if (!r) return;
    if (r->d && r->d->ring_open && r->d->ring == r) {   // this ring's kernel has retired (or never started)
        r->d->ring_open = false;
        r->d->ring = nullptr;
        ring_closed(r->d->device);
    }
    if (r->ring_stream) cudaStreamDestroy(r->ring_stream);
    if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
    host_free_or_park(r->h_in);
    host_free_or_park(r->h_out);
    host_free_or_park((void *)r->h_seq_done);
    host_free_or_park(r->h_doorbell);
    r->d_in.release();
    r->d_ctl.release();
    delete r;
}

int sdr_demod_ring_open(sdr_demod *d, size_t buf_len, uint32_t n_slots, sdr_ring **out) {
    if (!d || !out) return fail(SDR_E_ARG, "sdr_demod_ring_open: null argument");
    if (n_slots < 2 || n_slots > 64) return fail(SDR_E_ARG, "n_slots must be in [2, 64]");
    int rc = use_device(d->device);
    if (rc) return rc;
    if ((rc = commit_pending(d))) return rc;
    if ((rc = validate_lens(d, buf_len, d->cfg.downsample))) return rc;   // every phase a buffer can start on
    sdr_ring *r = new sdr_ring();
    r->d = d;
    r->buf_len = buf_len;
    r->slot_stride = (buf_len + 15) & ~size_t(15);
    r->n_slots = n_slots;
    const uint64_t S = buf_len / 2, D = d->cfg.downsample, fast = d->cfg.rate_out, slow = d->cfg.rate_resample;
    r->out_stride = (size_t)((((S + D - 1) / D + 1) * slow + fast - 1) / fast + 8);
    r->p0 = (uint32_t)d->st.prev_index;
    r->q0 = (uint32_t)d->st.prev_lpr_index;
    cudaError_t e = cudaHostAlloc((void **)&r->h_in, (size_t)n_slots * buf_len, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&r->h_out, (size_t)n_slots * r->out_stride * 2, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&r->h_out_dev, r->h_out, 0);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&r->h_seq_done, 64 * sizeof(unsigned int), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&r->h_seq_done_dev, (void *)r->h_seq_done, 0);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&r->h_doorbell, 72 * sizeof(unsigned int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->ring_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        ring_release(r);
        return fail(SDR_E_CUDA, "sdr_demod_ring_open: %s", cudaGetErrorString(e));
    }
    if ((rc = r->d_in.reserve((size_t)n_slots * r->slot_stride + 64)) || (rc = r->d_ctl.reserve(sizeof(RingCtl)))) {
        ring_release(r);
        return rc;
    }
    for (int i = 0; i < 64; i++) r->h_seq_done[i] = 0;
    RingCtl *h_ctl = new RingCtl();
    memset(h_ctl, 0, sizeof(RingCtl));
    to_dev_state(d->st, h_ctl->states[0]);
    // stream-ordered and waited for on this ring's own stream: no device-wide wait (another ring may be resident)
    e = cudaMemcpyAsync(r->d_ctl.p, h_ctl, sizeof(RingCtl), cudaMemcpyHostToDevice, r->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->copy_stream);
    delete h_ctl;
    if (e != cudaSuccess) {
        ring_release(r);
        return fail(SDR_E_CUDA, "sdr_demod_ring_open: %s", cudaGetErrorString(e));
    }
    // nothing may be loaded lazily once a ring kernel is resident (see KernelList in common.cuh)
    if ((rc = preload_kernels())) {
        ring_release(r);
        return rc;
    }
    RingArgs a{};
    Plan pl = make_plan(d->cfg, 0, 0, S, 1);
    // constants of the per-buffer FusedArgs (launch_fused fills the same fields for the one-shot kernel)
    FusedArgs &f = a.proto;
    f.S = (uint32_t)S;
    f.D = d->cfg.downsample;
    f.fast = d->cfg.rate_out;
    f.slow = d->cfg.rate_resample;
    f.div = (int32_t)(d->cfg.rate_out / d->cfg.rate_resample);
    f.div_slow = magic32(d->cfg.rate_resample);
    f.div_audio = magic32(d->cfg.rate_out / d->cfg.rate_resample);
    f.d64_slow = magic64(d->cfg.rate_resample);
    f.d64_S = magic64(S);
    f.d64_D = magic64(d->cfg.downsample);
    const bool d6 = d->cfg.downsample == 6 && !(r->p0 & 1);   // an odd window start takes the generic kernel
    const Geom &rg = (d->cfg.downsample == 6 && !d6) ? d->geo_gen : d->geo_ring;
    f.EB = rg.EB;
    f.lp_cap = rg.lp_cap;
    f.tile_cap = rg.tile_cap;
    f.dm_off = rg.dm_off;
    f.oct = d->oct;
    a.ctl = r->d_ctl.as<RingCtl>();
    a.seq_done = r->h_seq_done_dev;
    a.d_in = r->d_in.as<uint8_t>();
    a.h_out = r->h_out_dev;
    a.buf_len = buf_len;
    a.slot_stride = r->slot_stride;
    a.out_stride = r->out_stride;
    a.n_slots = n_slots;
    a.p0 = r->p0;
    a.q0 = r->q0;
    a.d64_fast = magic64(d->cfg.rate_out);
    unsigned tiles = (unsigned)((pl.Etot + 1 + rg.EB - 1) / rg.EB + 1);
    unsigned grid = std::max(1u, std::min(tiles, (unsigned)sm_count(d->device) / 2));
    e = d6 ? raise_dyn_smem(k_demod_ring<6>, d->smem_bytes) : raise_dyn_smem(k_demod_ring<0>, d->smem_bytes);
    if (e == cudaSuccess) {
        if (d6)
            k_demod_ring<6><<<grid, 256, rg.smem_staged, r->ring_stream>>>(a);
        else
            k_demod_ring<0><<<grid, 256, rg.smem_staged, r->ring_stream>>>(a);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        ring_release(r);
        return fail(SDR_E_CUDA, "sdr_demod_ring_open: launch failed: %s", cudaGetErrorString(e));
    }
    count_launch();
    d->ring_open = true;
    d->ring = r;
    ring_opened(d->device);
    *out = r;
    return SDR_OK;
}

int sdr_ring_acquire(sdr_ring *r, uint8_t **buf) {
    if (!r || !buf) return fail(SDR_E_ARG, "sdr_ring_acquire: null argument");
    if (r->acquired) return fail(SDR_E_STATE, "a slot is already acquired (commit it first)");
    while (r->head.load() - r->tail.load() >= r->n_slots) std::this_thread::yield();   // ring full: wait for collect()
    *buf = r->h_in + (size_t)(r->head.load() % r->n_slots) * r->buf_len;
    r->acquired = true;
    return SDR_OK;
}

int sdr_ring_commit(sdr_ring *r) {
    if (!r) return fail(SDR_E_ARG, "null ring");
    if (!r->acquired) return fail(SDR_E_STATE, "no slot acquired");
    int rc = use_device(r->d->device);
    if (rc) return rc;
    const uint64_t k = r->head.load();
    const uint32_t slot = (uint32_t)(k % r->n_slots);
    SDR_CUDA_TRY(cudaMemcpyAsync(r->d_in.as<uint8_t>() + (size_t)slot * r->slot_stride, r->h_in + (size_t)slot * r->buf_len,
                                 r->buf_len, cudaMemcpyHostToDevice, r->copy_stream));
    r->h_doorbell[slot] = (unsigned int)(k + 1);
    SDR_CUDA_TRY(cudaMemcpyAsync(&r->d_ctl.as<RingCtl>()->seq_ready[slot], &r->h_doorbell[slot], sizeof(unsigned int),
                                 cudaMemcpyHostToDevice, r->copy_stream));   // stream order: data first, then the doorbell
    r->acquired = false;
    r->head.store(k + 1);
    return SDR_OK;
}

long sdr_ring_collect(sdr_ring *r, int16_t *out, size_t cap) {
    if (!r || !out) return fail(SDR_E_ARG, "sdr_ring_collect: null argument");
    const uint64_t k = r->tail.load();
    if (k == r->head.load()) return fail(SDR_E_STATE, "nothing outstanding");
    const uint32_t slot = (uint32_t)(k % r->n_slots);
    // closed-form audio count of buffer k
    const uint64_t S = r->buf_len / 2, D = r->d->cfg.downsample;
    const uint64_t n0 = r->p0 + k * S, Lstart = n0 / D, pk = n0 % D;
    const unsigned __int128 t0 = (unsigned __int128)Lstart * r->d->cfg.rate_resample + r->q0;
    const uint32_t qk = (uint32_t)(t0 % r->d->cfg.rate_out);
    Plan pl = make_plan(r->d->cfg, (uint32_t)pk, qk, S, 1);
    if (pl.Etot > cap) return fail(SDR_E_CAP, "output capacity %zu < %llu audio samples", cap, (unsigned long long)pl.Etot);
    uint64_t spins = 0;
    while (r->h_seq_done[slot] != (unsigned int)(k + 1)) {
        if ((++spins & 0xfff) == 0) {
            cudaError_t e = cudaStreamQuery(r->ring_stream);
            if (e != cudaErrorNotReady) {
                (void)cudaGetLastError();
                return fail(SDR_E_CUDA, "ring kernel is not running (%s)", cudaGetErrorString(e));
            }
            std::this_thread::yield();
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(out, r->h_out + (size_t)slot * r->out_stride, pl.Etot * sizeof(int16_t));
    r->tail.store(k + 1);
    return (long)pl.Etot;
}

int sdr_ring_close(sdr_ring *r) {
    if (!r) return fail(SDR_E_ARG, "null ring");
    sdr_demod *d = r->d;
    int rc = use_device(d->device);
    if (rc) return rc;
    r->h_doorbell[64] = 1;
    cudaError_t e = cudaMemcpyAsync(&r->d_ctl.as<RingCtl>()->stop, &r->h_doorbell[64], sizeof(unsigned int),
                                    cudaMemcpyHostToDevice, r->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->ring_stream);   // every committed buffer is processed before stop is seen
    const uint64_t n = r->head.load();
    IntState hs{};
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(&hs, &r->d_ctl.as<RingCtl>()->states[n % (r->n_slots + 1)], sizeof(IntState), cudaMemcpyDeviceToHost,
                            r->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->copy_stream);
    if (e == cudaSuccess) {
        const uint64_t S = r->buf_len / 2;
        Plan pl = make_plan(d->cfg, r->p0, r->q0, S, n);
        d->dev_state_slot = -1;
        d->st.prev_index = pl.p1;
        d->st.prev_lpr_index = (int32_t)pl.q1;
        d->st.lp_now_re = hs.lp_now_re;
        d->st.lp_now_im = hs.lp_now_im;
        d->st.demod_pre_re = hs.demod_pre_re;
        d->st.demod_pre_im = hs.demod_pre_im;
        d->st.now_lpr = hs.now_lpr;
    }
    ring_release(r);
    if (e != cudaSuccess) return fail(SDR_E_CUDA, "sdr_ring_close: %s", cudaGetErrorString(e));
    return SDR_OK;
}

int sdr_demod_span_begin(sdr_demod *d) {
    if (!d) return fail(SDR_E_ARG, "null handle");
    int rc = use_device(d->device);
    if (rc) return rc;
    SDR_CUDA_TRY(cudaEventRecord(d->ev_s0, d->stream));
    return SDR_OK;
}

int sdr_demod_span_end(sdr_demod *d, float *ms) {
    if (!d || !ms) return fail(SDR_E_ARG, "null argument");
    int rc = use_device(d->device);
    if (rc) return rc;
    SDR_CUDA_TRY(cudaEventRecord(d->ev_s1, d->stream));
    SDR_CUDA_TRY(cudaEventSynchronize(d->ev_s1));
    SDR_CUDA_TRY(cudaEventElapsedTime(ms, d->ev_s0, d->ev_s1));
    return SDR_OK;
}

// ---- stage entry points ---------------------------------------------------------------------------

long sdr_rotate_90(sdr_demod *d, uint8_t *buf, size_t len) {
    if (!d || !buf) return fail(SDR_E_ARG, "null argument");
    if (len % 8) return fail(SDR_E_LEN, "rotate_90 needs len %% 8 == 0 (examples/simple_fm.rs:284-295)");
    int rc = use_device(d->device);
    if (rc) return rc;
    if (len == 0) return 0;
    if ((rc = d->d_a.reserve(len))) return rc;
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_a.p, buf, len, cudaMemcpyHostToDevice, d->stream));
    k_rotate_90<<<grid_for(len / 8, d->device), 256, 0, d->stream>>>(d->d_a.as<uint2>(), len / 8);
    SDR_LAUNCH_CHECK();
    SDR_CUDA_TRY(cudaMemcpyAsync(buf, d->d_a.p, len, cudaMemcpyDeviceToHost, d->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(d->stream));
    return (long)len;
}

long sdr_buf_to_complex(sdr_demod *d, const uint8_t *buf, size_t len, int32_t *out_pairs, size_t cap_pairs) {
    if (!d || !buf || !out_pairs) return fail(SDR_E_ARG, "null argument");
    size_t n = len / 2;   // windows(2).step_by(2): a trailing odd byte is dropped (:441-450)
    if (n > cap_pairs) return fail(SDR_E_CAP, "capacity %zu < %zu pairs", cap_pairs, n);
    int rc = use_device(d->device);
    if (rc) return rc;
    if (n == 0) return 0;
    if ((rc = d->d_a.reserve(len + 16)) || (rc = d->d_b.reserve(n * 8))) return rc;
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_a.p, buf, n * 2, cudaMemcpyHostToDevice, d->stream));
    k_buf_to_complex<<<grid_for(n, d->device), 256, 0, d->stream>>>(d->d_a.as<uint16_t>(), n, d->d_b.as<int2>());
    SDR_LAUNCH_CHECK();
    SDR_CUDA_TRY(cudaMemcpyAsync(out_pairs, d->d_b.p, n * 8, cudaMemcpyDeviceToHost, d->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(d->stream));
    return (long)n;
}

static int upload_state(sdr_demod *d) {
    IntState hs;
    to_dev_state(d->st, hs);
    d->h_state.as<IntState>()[0] = hs;
    d->h_state.as<IntState>()[1] = hs;
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_state.p, d->h_state.p, 2 * sizeof(IntState), cudaMemcpyHostToDevice, d->stream));
    return SDR_OK;
}
static int download_state(sdr_demod *d, IntState *hs) {
    SDR_CUDA_TRY(cudaMemcpyAsync(d->h_state.p, d->d_state.as<IntState>() + 1, sizeof(IntState), cudaMemcpyDeviceToHost,
                                 d->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(d->stream));
    *hs = *d->h_state.as<IntState>();
    return SDR_OK;
}

long sdr_low_pass_complex(sdr_demod *d, const int32_t *iq, size_t n, int32_t *out_pairs, size_t cap_pairs) {
    if (!d || (!iq && n) || !out_pairs) return fail(SDR_E_ARG, "null argument");
    int rc = use_device(d->device);
    if (rc) return rc;
    if ((rc = commit_pending(d))) return rc;
    const uint32_t D = d->cfg.downsample, p0 = (uint32_t)d->st.prev_index;
    const uint64_t L = (p0 + (uint64_t)n) / D;
    if (L > cap_pairs) return fail(SDR_E_CAP, "capacity %zu < %llu pairs", cap_pairs, (unsigned long long)L);
    if (n == 0) return 0;
    if ((rc = d->d_a.reserve(n * 8)) || (rc = d->d_b.reserve((L + 1) * 8))) return rc;
    if ((rc = upload_state(d))) return rc;
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_a.p, iq, n * 8, cudaMemcpyHostToDevice, d->stream));
    IntState *ds = d->d_state.as<IntState>();
    k_low_pass_complex<<<grid_for(L + 1, d->device), 256, 0, d->stream>>>(d->d_a.as<int2>(), n, D, p0, L, ds, ds + 1,
                                                                            d->d_b.as<int2>());
    SDR_LAUNCH_CHECK();
    if (L) SDR_CUDA_TRY(cudaMemcpyAsync(out_pairs, d->d_b.p, L * 8, cudaMemcpyDeviceToHost, d->stream));
    IntState hs;
    if ((rc = download_state(d, &hs))) return rc;
    d->dev_state_slot = -1;
    d->st.lp_now_re = hs.lp_now_re;
    d->st.lp_now_im = hs.lp_now_im;
    d->st.prev_index = (p0 + (uint64_t)n) % D;
    return (long)L;
}

long sdr_fm_demod(sdr_demod *d, const int32_t *iq, size_t n, int16_t *out, size_t cap) {
    if (!d || !iq || !out) return fail(SDR_E_ARG, "null argument");
    if (n < 2) return fail(SDR_E_LEN, "fm_demod asserts buf.len() > 1 (examples/simple_fm.rs:356)");
    if (n > cap) return fail(SDR_E_CAP, "capacity %zu < %zu", cap, n);
    int rc = use_device(d->device);
    if (rc) return rc;
    if ((rc = commit_pending(d))) return rc;
    if ((rc = d->d_a.reserve(n * 8)) || (rc = d->d_b.reserve(n * 2))) return rc;
    if ((rc = upload_state(d))) return rc;
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_a.p, iq, n * 8, cudaMemcpyHostToDevice, d->stream));
    IntState *ds = d->d_state.as<IntState>();
    k_fm_demod<<<grid_for(n, d->device), 256, 0, d->stream>>>(d->d_a.as<int2>(), n, ds, ds + 1, d->oct, d->d_b.as<int16_t>());
    SDR_LAUNCH_CHECK();
    SDR_CUDA_TRY(cudaMemcpyAsync(out, d->d_b.p, n * 2, cudaMemcpyDeviceToHost, d->stream));
    IntState hs;
    if ((rc = download_state(d, &hs))) return rc;
    d->dev_state_slot = -1;
    d->st.demod_pre_re = hs.demod_pre_re;
    d->st.demod_pre_im = hs.demod_pre_im;
    return (long)n;
}

long sdr_low_pass_real(sdr_demod *d, const int16_t *in, size_t n, int16_t *out, size_t cap) {
    if (!d || (!in && n) || !out) return fail(SDR_E_ARG, "null argument");
    int rc = use_device(d->device);
    if (rc) return rc;
    if ((rc = commit_pending(d))) return rc;
    const uint32_t fast = d->cfg.rate_out, slow = d->cfg.rate_resample, q0 = (uint32_t)d->st.prev_lpr_index;
    unsigned __int128 t = (unsigned __int128)n * slow + q0;
    const uint64_t E = (uint64_t)(t / fast);
    if (E > cap) return fail(SDR_E_CAP, "capacity %zu < %llu", cap, (unsigned long long)E);
    if (n == 0) return 0;
    if ((rc = d->d_a.reserve(n * 2)) || (rc = d->d_b.reserve((E + 1) * 2))) return rc;
    if ((rc = upload_state(d))) return rc;
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_a.p, in, n * 2, cudaMemcpyHostToDevice, d->stream));
    IntState *ds = d->d_state.as<IntState>();
    k_low_pass_real<<<grid_for(E + 1, d->device), 256, 0, d->stream>>>(d->d_a.as<int16_t>(), n, q0, fast, slow,
                                                                         (int32_t)(fast / slow), E, ds, ds + 1,
                                                                         d->d_b.as<int16_t>());
    SDR_LAUNCH_CHECK();
    if (E) SDR_CUDA_TRY(cudaMemcpyAsync(out, d->d_b.p, E * 2, cudaMemcpyDeviceToHost, d->stream));
    IntState hs;
    if ((rc = download_state(d, &hs))) return rc;
    d->dev_state_slot = -1;
    d->st.now_lpr = hs.now_lpr;
    d->st.prev_lpr_index = (int32_t)(uint32_t)(t % fast);
    return (long)E;
}

long sdr_fast_atan2(sdr_demod *d, const int32_t *y, const int32_t *x, size_t n, int32_t *out) {
    if (!d || !y || !x || !out) return fail(SDR_E_ARG, "null argument");
    int rc = use_device(d->device);
    if (rc) return rc;
    if (n == 0) return 0;
    if ((rc = d->d_a.reserve(n * 4)) || (rc = d->d_b.reserve(n * 4)) || (rc = d->d_c.reserve(n * 4))) return rc;
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_a.p, y, n * 4, cudaMemcpyHostToDevice, d->stream));
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_b.p, x, n * 4, cudaMemcpyHostToDevice, d->stream));
    k_fast_atan2_v<<<grid_for(n, d->device), 256, 0, d->stream>>>(d->d_a.as<int32_t>(), d->d_b.as<int32_t>(), n,
                                                                    d->d_c.as<int32_t>());
    SDR_LAUNCH_CHECK();
    SDR_CUDA_TRY(cudaMemcpyAsync(out, d->d_c.p, n * 4, cudaMemcpyDeviceToHost, d->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(d->stream));
    return (long)n;
}

long sdr_polar_discriminant(sdr_demod *d, const int32_t *a, const int32_t *b, size_t n, int fast, int32_t *out) {
    if (!d || !a || !b || !out) return fail(SDR_E_ARG, "null argument");
    int rc = use_device(d->device);
    if (rc) return rc;
    if (n == 0) return 0;
    if ((rc = d->d_a.reserve(n * 8)) || (rc = d->d_b.reserve(n * 8)) || (rc = d->d_c.reserve(n * 4))) return rc;
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_a.p, a, n * 8, cudaMemcpyHostToDevice, d->stream));
    SDR_CUDA_TRY(cudaMemcpyAsync(d->d_b.p, b, n * 8, cudaMemcpyHostToDevice, d->stream));
    k_polar_v<<<grid_for(n, d->device), 256, 0, d->stream>>>(d->d_a.as<int2>(), d->d_b.as<int2>(), n, fast, d->oct,
                                                               d->d_c.as<int32_t>());
    SDR_LAUNCH_CHECK();
    SDR_CUDA_TRY(cudaMemcpyAsync(out, d->d_c.p, n * 4, cudaMemcpyDeviceToHost, d->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(d->stream));
    return (long)n;
}

}  // extern "C"
