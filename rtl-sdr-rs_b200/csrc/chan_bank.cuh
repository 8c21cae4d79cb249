// chan_bank.cuh — k_chan_bank<K2>: the channeliser for UNIFORMLY SPACED channels (BASELINE.json configs 4-5), as a
// two-stage polyphase filter bank on CUDA cores.
//
// The direct form (k_chan_fir_u) spends 4*C*T/D FMAs per input sample: every channel runs its own T-tap complex FIR.
// When the channel centres sit on a uniform grid  fw_c = f0 + c * 2^32/K  (K bins; cfg4: K = 100, the interleaved cfg5
// plan: K = 64) the C filters share almost all of their work.  With tap index k, r = k mod K = r1 + K1*r2 (K = K1*K2):
//
//   S_c[m] = sum_k h[k] e^{+j theta_c(k)} xc[n_m - k]                       (n_m = (m+1)D - 1, xc = x - 127)
//          = sum_{r1 < K1} E[c][r1] * A[r1][c mod K2],                       E[c][r1] = e^{+j 2 pi c r1 / K}
//   A[r1][b2] = sum_{k = r1 (mod K1)} G[b2][k] xc[n_m - k],                 G[b2][k] = h[k] e^{+j 2 pi (f0 k / 2^32 + b2 r2(k) / K2)}
//
// i.e. stage 1 is K "sub-filters" of T/K1 complex taps each (T*K2 complex MACs per output time, shared by ALL channels),
// stage 2 a K1-term complex combination per channel (C*K1 MACs) — 2555 complex MACs instead of 16320 for cfg4
// (K1 = 20, K2 = 5).  Both stages have the shape k_chan_fir_u already runs at 76 % of the FP32 peak: a lane owns one
// output time, every coefficient is a scalar-broadcast uniform-register operand of a packed FFMA2 fed straight from the
// kernel-parameter constant bank (no shared-memory coefficient traffic), the only per-lane load is its raw sample.
// The grid is exact up to the rounding of the 32-bit NCO words the caller passes: the host checks that every tap phase
// of every channel is within 2e-6 rad of the ideal grid (else the direct kernel runs), and the per-output de-rotation
// uses the exact words.
//
// The discriminator is fused: lane t owns output out0 - 1 + t, so the predecessor S[m-1] is one shuffle away (lane 0 of
// a CTA is a halo lane that only supplies it).  Because theta_c(n_m) - theta_c(n_{m-1}) = fw_c * D is a per-channel
// constant, arg(y[m] conj y[m-1]) = arg(S[m] conj S[m-1]) - phi_c: no sincos is needed unless the caller asks for y.
#pragma once
#include "ptx_helpers.cuh"

namespace sdr {

#ifndef SDR_BANK_NT
#define SDR_BANK_NT 128
#endif
#ifndef SDR_BANK_MINB
#define SDR_BANK_MINB 3
#endif
#ifndef SDR_BANK_UNROLL
#define SDR_BANK_UNROLL 4
#endif
constexpr int kBankUnroll = SDR_BANK_UNROLL;
constexpr int kBankThreads = SDR_BANK_NT;   // lanes per CTA (one output time each; lane 0 is the predecessor halo)
constexpr int kBankCH = 64;            // channels per launch (packed accumulators held in registers)
constexpr int kBankTabEntries = 3840;  // float2 entries of the coefficient parameter (30 KB of the 32 KB limit)
constexpr int kBankMaxK2 = 8;

// Every residue r1 gets the same number of taps NJ = ceil(T / K1): the filter is zero-padded to Tp = K1*NJ taps.
// [0, Tp*K2): G, sub-filter taps in the order they are consumed: for r1, for j (k = r1 + j*K1), for b2
// [eoff = Tp*K2 rounded up to even, eoff + K1*64): E[r1][c], per channel pair (re c, re c+1), (im c, im c+1);   [.., + 64): (phi_c, 0)
struct BankTab {
    float2 v[kBankTabEntries];
};

struct BankArgs {
    const uint8_t *x;
    const uint8_t *carry_end;
    const uint32_t *fw;        // [C] the caller's exact NCO words (de-rotation when y is wanted)
    float2 *y_out;             // [C][cap] or nullptr
    float *d_out;              // [C][cap] or nullptr
    const float2 *prev_in;     // [C] S of the last output before this call
    float2 *prev_out;          // [C] S of the last output of this call
    long long n_samples, n_out, cap;
    uint32_t r, n0_lo;
    int ch0, n_ch, Tp, D, K1, NJ;   // Tp = K1 * NJ: taps after zero padding
    int eoff;                       // first E entry (host-computed: keeps the index arithmetic on the uniform datapath)
    float gain;
};

__device__ __forceinline__ unsigned long long bk_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void bk_fma2(unsigned long long &acc, float s, unsigned long long x) {
    const unsigned long long ss = bk_pack(s, s);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ss), "l"(x));
}
__device__ __forceinline__ float2 bk_unpack(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

// One tile's raw bytes: samples [s0, s1) (call-local; negative = carry) -> xb.  One thread.  (Same staging as chan.cu.)
__device__ __forceinline__ uint32_t bank_load_tile(unsigned char *xb, const BankArgs &a, long long s0, long long s1, uint64_t *bar) {
    const uint32_t soff = (uint32_t)((2 * s0) & 15);
    uint32_t carry_bytes = 0, x_bytes = 0;
    const long long x_lo = s0 > 0 ? s0 : 0;
    if (s0 < 0) {
        const long long c_hi = s1 < 0 ? s1 : 0;
        carry_bytes = (uint32_t)((2 * (c_hi - s0) + soff + 15) & ~15ll);
    }
    const long long b_lo = (2 * x_lo) & ~15ll;
    if (s1 > 0) x_bytes = (uint32_t)(((2 * s1 + 15) & ~15ll) - b_lo);
    mbar_arrive_expect_tx(bar, carry_bytes + x_bytes);
    if (carry_bytes) bulk_g2s(xb, a.carry_end + 2 * s0 - soff, carry_bytes, bar);
    if (x_bytes) bulk_g2s(xb + soff + (b_lo - 2 * s0), a.x + b_lo, x_bytes, bar);
    return soff;
}

// K1T / NJT != 0: the split is a compile-time constant (the BASELINE.json shapes) — the tap loop is fully unrolled, every
// sample load carries an immediate offset and there is no loop bookkeeping; 0 / 0: any split, run-time loops.
template <int K2, int K1T = 0, int NJT = 0>
__global__ void __launch_bounds__(kBankThreads, SDR_BANK_MINB) k_chan_bank(const BankArgs a, const __grid_constant__ BankTab tab) {
    constexpr int CH = kBankCH, NT = kBankThreads, OUT = NT - 1;
    static_assert(K2 >= 1 && K2 <= kBankMaxK2, "stage-1 accumulators live in registers");
    static_assert((K1T == 0) == (NJT == 0), "either both compile-time or both run-time");
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t sh_soff;
    __shared__ float2 xch[NT / 32][CH];   // S of each warp's last lane: the predecessor of the next warp's lane 0
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long out0 = (long long)blockIdx.x * OUT;   // first OWNED output; thread t computes output out0 - 1 + t
    const long long n_here = a.n_out - out0 < OUT ? a.n_out - out0 : OUT;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        const long long s0 = (out0 - 1) * a.D - (long long)a.r - (a.Tp - 1);
        const long long s1 = (out0 + n_here) * a.D - (long long)a.r;
        sh_soff = bank_load_tile(smem, a, s0, s1, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    // newest sample of this thread's output, relative to the tile's first sample
    const uint16_t *t16 = reinterpret_cast<const uint16_t *>(smem + sh_soff) + (tid + 1) * a.D + a.Tp - 2;

    unsigned long long Y[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) Y[c] = 0ull;
    // opaque per-thread conversion constants (fir_fast.cuh: FHADD takes no immediate / uniform operand)
    float bias;
    uint32_t h1024;
    asm volatile(
        "{\n.reg .u32 t;\nmov.u32 t, %%tid.x;\nshr.u32 t, t, 31;\nor.b32 %0, t, 0xC48FE000;\nor.b32 %1, t, 0x64646464;\n}\n"
        : "=f"(bias), "=r"(h1024));

    const int K1 = K1T ? K1T : a.K1, NJ = NJT ? NJT : a.NJ;
    const int eoff = a.eoff;
    int gidx = 0, eidx = eoff;
#pragma unroll 1
    for (int r1 = 0; r1 < K1; r1++) {
        // ---- stage 1: the K2 sub-filters of residue r1 (taps k = r1 + j*K1) -----------------------------------------
        unsigned long long A[K2], B[K2];   // A += Re(G) * x, B += Im(G) * x;  result = (A.re - B.im, A.im + B.re)
#pragma unroll
        for (int b = 0; b < K2; b++) A[b] = B[b] = 0ull;
        const uint16_t *p = t16 - r1;
        auto tap = [&](const uint16_t raw, const int gi) {
            const uint32_t pair = __byte_perm((uint32_t)raw, h1024, 0x4140u);   // half2 (1024+I, 1024+Q)
            float xr, xi;
            asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(xr) : "h"((unsigned short)(pair & 0xffffu)), "f"(bias));
            asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(xi) : "h"((unsigned short)(pair >> 16)), "f"(bias));
            const unsigned long long x2 = bk_pack(xr, xi);
#pragma unroll
            for (int b = 0; b < K2; b++) {
                const float2 g = tab.v[gi + b];
                bk_fma2(A[b], g.x, x2);
                bk_fma2(B[b], g.y, x2);
            }
        };
        if constexpr (NJT != 0) {
#pragma unroll
            for (int j = 0; j < NJT; j++) tap(p[-j * K1T], gidx + j * K2);
            gidx += NJT * K2;
        } else {
            int koff = 0;
#pragma unroll kBankUnroll
            for (int j = 0; j < NJ; j++, koff += K1, gidx += K2) tap(p[-koff], gidx);
        }
        unsigned long long a2[K2], a2r[K2];   // the sub-filter output a and j*a
#pragma unroll
        for (int b = 0; b < K2; b++) {
            const float2 pa = bk_unpack(A[b]), pb = bk_unpack(B[b]);
            const float ar = pa.x - pb.y, ai = pa.y + pb.x;
            a2[b] = bk_pack(ar, ai);
            a2r[b] = bk_pack(-ai, ar);
        }
        // ---- stage 2: every channel adds E[c][r1] * A[r1][c mod K2] ---------------------------------------------------
        // eight channels at a time: four 16-byte uniform loads (two coefficients each), then the eight real-part FMAs,
        // then the eight imaginary-part FMAs — the two FMAs into one accumulator are never back to back
        // E rows are stored per channel PAIR as (re c, re c+1), (im c, im c+1): the two halves of every 64-bit uniform load
        // are consumed together (by two independent accumulators), so the load lands in the register pair the FMAs read.
        // With (re, im) pairs the halves had different lifetimes and two of three loads cost two UMOVs on top: the loop body
        // went from 601 to 479 instructions for (K2, K1, NJ) = (5, 20, 13).  The two FMAs into one accumulator are never
        // back to back.
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 4) {
            float4 e4[2];
#pragma unroll
            for (int q = 0; q < 2; q++) e4[q] = *reinterpret_cast<const float4 *>(&tab.v[eidx + c0 + 2 * q]);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                bk_fma2(Y[c0 + 2 * q], e4[q].x, a2[(c0 + 2 * q) % K2]);
                bk_fma2(Y[c0 + 2 * q + 1], e4[q].y, a2[(c0 + 2 * q + 1) % K2]);
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                bk_fma2(Y[c0 + 2 * q], e4[q].z, a2r[(c0 + 2 * q) % K2]);
                bk_fma2(Y[c0 + 2 * q + 1], e4[q].w, a2r[(c0 + 2 * q + 1) % K2]);
            }
        }
        eidx += CH;
    }

    // ---- epilogue: discriminator against the predecessor output, optional de-rotated y ------------------------------
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < CH; c++) xch[warp][c] = bk_unpack(Y[c]);
    }
    __syncthreads();
    const long long i = out0 - 1 + tid;                    // call-local output index of this thread
    const bool owns = tid >= 1 && (long long)(tid - 1) < n_here;
    const bool first_of_call = i == 0;                     // its predecessor is the carried state
    const bool from_xch = lane == 0 && warp > 0;           // its predecessor sits in the previous warp
    const int phib = eoff + K1 * CH;
    const size_t row0 = (size_t)a.ch0 * (size_t)a.cap + (size_t)(owns ? i : 0);
    float *dp = a.d_out ? a.d_out + row0 : nullptr;
    const bool want_d = owns && dp != nullptr;
    const float2 *pin = a.prev_in + a.ch0;
#pragma unroll
    for (int c = 0; c < CH; c++) {
        const float2 s = bk_unpack(Y[c]);
        float2 pv;
        pv.x = __shfl_up_sync(0xffffffffu, s.x, 1);
        pv.y = __shfl_up_sync(0xffffffffu, s.y, 1);
        if (from_xch) pv = xch[warp - 1][c];
        if (first_of_call) pv = pin[c];
        const float cre = fmaf(s.x, pv.x, s.y * pv.y);        // Re(S conj P)
        const float cim = fmaf(s.y, pv.x, -(s.x * pv.y));     // Im(S conj P)
        float t = poly_atan2(cim, cre) - tab.v[phib + c].x;
        t -= t > 3.14159265358979324f ? 6.28318530717958648f : 0.f;
        t += t <= -3.14159265358979324f ? 6.28318530717958648f : 0.f;
        // zero predecessor (stream start): 0 by definition (the polynomial's 0/0 is discarded here)
        const float d = (cre == 0.f && cim == 0.f) ? 0.f : a.gain * t;
        if (want_d && c < a.n_ch) dp[(size_t)c * (size_t)a.cap] = d;
    }
    if (a.y_out && owns) {   // the caller wants the de-rotated channel samples as well
        const uint32_t nm = a.n0_lo + (uint32_t)((i + 1) * a.D - 1) - a.r;   // global n_m mod 2^32
        float2 *yp = a.y_out + row0;
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (c >= a.n_ch) break;
            const float2 s = bk_unpack(Y[c]);
            float si, co;   // e^{+j theta_c(n_m)} from the top 24 phase bits; de-rotate with its conjugate
            __sincosf((float)(int32_t)(a.fw[a.ch0 + c] * nm) * (3.14159265358979324f / 2147483648.0f), &si, &co);
            yp[(size_t)c * (size_t)a.cap] = make_float2(s.x * co + s.y * si, s.y * co - s.x * si);
        }
    }
    if (owns && i == a.n_out - 1) {   // the last output of the call is the next call's predecessor
#pragma unroll
        for (int c = 0; c < CH; c++)
            if (c < a.n_ch) a.prev_out[a.ch0 + c] = bk_unpack(Y[c]);
    }
}

}  // namespace sdr
