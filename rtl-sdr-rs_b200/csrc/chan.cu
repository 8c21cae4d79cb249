// chan.cu — wideband channeliser (BASELINE.json configs 4-5) for sm_100a, and the NCCL slab broadcast.
//
// Per channel c:  y_c[m] = sum_k h[k] * xc[n_m - k] * e^{-j theta_c(n_m - k)},   n_m = (m+1)D - 1,
//                 theta_c(n) = 2*pi * ((fw_c * n) mod 2^32) / 2^32              (32-bit phase NCO)
// Folding the NCO into the taps (exact, because the phase arithmetic is modular):
//                 y_c[m] = e^{-j theta_c(n_m)} * sum_k g_c[k] * xc[n_m - k],    g_c[k] = h[k] e^{+j theta_c(k)}
// so the inner loop has no sincos: it is C*T complex MACs per output column on CUDA cores
// (~4*C*T/D FMAs per input sample: FP32-FMA-bound, not HBM-bound).
//
// k_chan_fir: persistent CTAs (one per SM).  The 64 channels' folded taps stay resident in shared
// memory ([k][channel], 130 KB for T=255); raw bytes of the next output tile arrive with a bulk
// async copy while the current tile computes; bytes are converted to f32 once per tile.  A lane owns
// 2 channels x MR outputs (register tile), taps come in as one LDS.128 per tap (2 channels), samples
// as warp-broadcast loads, so shared-memory wavefronts stay at ~50% of the FMA issue time.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "chan_bank.cuh"

namespace sdr {

constexpr int kChanGroup = 64;   // channels per CTA (lane <-> 2 channels)

struct ChanArgs {
    const uint8_t *x;
    const uint8_t *carry_end;
    long long n_samples, n_out;
    uint32_t r;                   // samples of the current decimation block consumed before x
    uint32_t n0_lo;               // low 32 bits of the global index of x[0]
    const float2 *gtaps;          // [groups][T4][64]
    const uint32_t *fw;           // [groups*64]
    float2 *y_out;                // [C_pad][cap]
    long long cap;
    int T4, D, pad;               // taps padded to a multiple of 4; pad: xs origin shift (alignment)
    int n_tiles;
    int n_ch;                     // channels that exist: rows >= n_ch of the last group are never stored
    uint32_t sm_taps, sm_xs, sm_xb;   // byte sizes of the smem regions
};

__device__ __forceinline__ void cis_phase(uint32_t phase, float &c, float &s) {
    // e^{+j 2 pi phase/2^32}: split so both sincospif arguments are exact in f32
    float hi = (float)(phase >> 16) * (1.0f / 32768.0f);              // 2*phase_hi/2^16, exact
    float lo = (float)(phase & 0xffffu) * (1.0f / 2147483648.0f);     // 2*phase_lo/2^32, exact
    float ch, sh, cl, sl;
    sincospif(hi, &sh, &ch);
    sincospif(lo, &sl, &cl);
    c = ch * cl - sh * sl;
    s = sh * cl + ch * sl;
}

// Stage one tile's raw bytes: samples [s0, s1) (call-local; negative = carry) -> xb.  One thread.
__device__ __forceinline__ uint32_t chan_load_tile(unsigned char *xb, const ChanArgs &a, long long s0, long long s1,
                                                   uint64_t *bar) {
    const uint32_t soff = (uint32_t)((2 * s0) & 15);
    uint32_t carry_bytes = 0, x_bytes = 0;
    long long x_lo = s0 > 0 ? s0 : 0;
    if (s0 < 0) {
        long long c_hi = s1 < 0 ? s1 : 0;
        carry_bytes = (uint32_t)((2 * (c_hi - s0) + soff + 15) & ~15ll);
    }
    long long b_lo = (2 * x_lo) & ~15ll;
    if (s1 > 0) x_bytes = (uint32_t)(((2 * s1 + 15) & ~15ll) - b_lo);
    mbar_arrive_expect_tx(bar, carry_bytes + x_bytes);
    if (carry_bytes) bulk_g2s(xb, a.carry_end + 2 * s0 - soff, carry_bytes, bar);
    if (x_bytes) bulk_g2s(xb + soff + (b_lo - 2 * s0), a.x + b_lo, x_bytes, bar);
    return soff;
}

// packed FP32 (Blackwell FFMA2): acc(lo,hi) += s * (x.lo, x.hi), scalar s broadcast to both lanes
__device__ __forceinline__ unsigned long long c_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void c_fma2(unsigned long long &acc, float s, unsigned long long x) {
    const unsigned long long ss = c_pack(s, s);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ss), "l"(x));
}
__device__ __forceinline__ float2 c_unpack(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

template <int W, int MR, bool ALIGNED>
__global__ void __launch_bounds__(W * 32, 1) k_chan_fir(const ChanArgs a) {
    constexpr int MT = W * MR;   // outputs per tile: W warps x MR outputs
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t sh_soff[2];
    float2 *gt = reinterpret_cast<float2 *>(smem);                       // [T4][64]
    float2 *xs = reinterpret_cast<float2 *>(smem + a.sm_taps);           // converted samples / y staging
    unsigned char *xb0 = smem + a.sm_taps + a.sm_xs;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = blockIdx.y;
    const int T4 = a.T4, D = a.D;
    const int NS = MT * D + T4 - 1 + a.pad;      // samples per tile (incl. history)

    // resident folded taps for this channel group
    {
        const float4 *src = reinterpret_cast<const float4 *>(a.gtaps + (size_t)group * T4 * kChanGroup);
        float4 *dst = reinterpret_cast<float4 *>(gt);
        for (int i = tid; i < T4 * kChanGroup / 2; i += blockDim.x) dst[i] = src[i];
    }
    const uint32_t fw0 = a.fw[group * kChanGroup + 2 * lane], fw1 = a.fw[group * kChanGroup + 2 * lane + 1];

    auto tile_s0 = [&](int tile) { return (long long)tile * MT * D - (long long)a.r - (T4 - 1) - a.pad; };
    auto issue = [&](int tile, int buf) {
        long long s0 = tile_s0(tile);
        long long n_here = a.n_out - (long long)tile * MT < MT ? a.n_out - (long long)tile * MT : MT;
        long long s1 = ((long long)tile * MT + n_here) * D - (long long)a.r;
        sh_soff[buf] = chan_load_tile(xb0 + (size_t)buf * a.sm_xb, a, s0, s1, &bar[buf]);
    };

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
        if ((int)blockIdx.x < a.n_tiles) issue(blockIdx.x, 0);
    }
    __syncthreads();

    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, it++) {
        const int buf = it & 1;
        // prefetch the next tile's bytes into the other buffer (its previous contents were consumed
        // before the __syncthreads that ended the previous iteration's conversion phase)
        if (tid == 0 && tile + (int)gridDim.x < a.n_tiles) issue(tile + gridDim.x, buf ^ 1);
        mbar_wait(&bar[buf], (it >> 1) & 1);
        // ---- convert u8 -> centred f32 once per tile ---------------------------------------------
        {
            const uint16_t *t16 = reinterpret_cast<const uint16_t *>(xb0 + (size_t)buf * a.sm_xb + sh_soff[buf]);
            for (int j = tid; j < NS; j += blockDim.x) {
                uint32_t v = t16[j];
                uint32_t bi = __byte_perm(v, 0x4B000000u, 0x7440u), bq = __byte_perm(v, 0x4B000000u, 0x7441u);
                xs[j] = make_float2(__uint_as_float(bi) - 8388735.0f, __uint_as_float(bq) - 8388735.0f);
            }
        }
        __syncthreads();

        // ---- register-tiled complex MAC: lane = 2 channels, warp = MR outputs ------------------------
        float ar0[MR], ai0[MR], ar1[MR], ai1[MR];
#pragma unroll
        for (int o = 0; o < MR; o++) ar0[o] = ai0[o] = ar1[o] = ai1[o] = 0.f;
        // newest sample of output o (tile-local) sits at xs[(o+1)*D + T4 - 2 + pad]
        const int base0 = (warp * MR + 1) * D + T4 - 2 + a.pad;
        const float4 *g4 = reinterpret_cast<const float4 *>(gt) + lane;   // [k][32 lanes] float4 = 2 channels
        for (int k0 = 0; k0 < T4; k0 += 4) {
            float4 g[4];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) g[kk] = g4[(k0 + kk) * 32];
#pragma unroll
            for (int o = 0; o < MR; o++) {
                const int j = base0 + o * D - k0;   // sample for tap k0; taps k0+1.. use j-1..
                float2 x[4];
                if (ALIGNED) {
                    // j is odd: (j-1, j) and (j-3, j-2) are 16-byte aligned pairs
                    float4 p = *reinterpret_cast<const float4 *>(&xs[j - 1]);
                    float4 q = *reinterpret_cast<const float4 *>(&xs[j - 3]);
                    x[0] = make_float2(p.z, p.w);
                    x[1] = make_float2(p.x, p.y);
                    x[2] = make_float2(q.z, q.w);
                    x[3] = make_float2(q.x, q.y);
                } else {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) x[kk] = xs[j - kk];
                }
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    // (gr + j gi) * (xr + j xi)
                    ar0[o] = fmaf(g[kk].x, x[kk].x, ar0[o]);
                    ar0[o] = fmaf(-g[kk].y, x[kk].y, ar0[o]);
                    ai0[o] = fmaf(g[kk].x, x[kk].y, ai0[o]);
                    ai0[o] = fmaf(g[kk].y, x[kk].x, ai0[o]);
                    ar1[o] = fmaf(g[kk].z, x[kk].x, ar1[o]);
                    ar1[o] = fmaf(-g[kk].w, x[kk].y, ar1[o]);
                    ai1[o] = fmaf(g[kk].z, x[kk].y, ai1[o]);
                    ai1[o] = fmaf(g[kk].w, x[kk].x, ai1[o]);
                }
            }
        }
        __syncthreads();   // everyone is done reading xs: reuse it to transpose the y tile

        // ---- de-rotate by the NCO phase at n_m and stage [channel][output] ------------------------------
        float2 *ys = xs;   // [64][MT]
#pragma unroll
        for (int o = 0; o < MR; o++) {
            const long long i = (long long)tile * MT + warp * MR + o;            // call-local output index
            const uint32_t nm = a.n0_lo + (uint32_t)((i + 1) * D - 1) - a.r;     // global n_m mod 2^32
            float c, s;
            cis_phase(fw0 * nm, c, s);   // e^{+j theta}; we need e^{-j theta}: (yr + j yi)(c - j s)
            ys[(2 * lane) * MT + warp * MR + o] = make_float2(ar0[o] * c + ai0[o] * s, ai0[o] * c - ar0[o] * s);
            cis_phase(fw1 * nm, c, s);
            ys[(2 * lane + 1) * MT + warp * MR + o] = make_float2(ar1[o] * c + ai1[o] * s, ai1[o] * c - ar1[o] * s);
        }
        __syncthreads();
        {
            const long long i0 = (long long)tile * MT;
            const int n_here = (int)(a.n_out - i0 < MT ? a.n_out - i0 : MT);
            for (int e = tid; e < kChanGroup * MT; e += blockDim.x) {
                const int c = e / MT, o = e % MT;
                if (o < n_here && group * kChanGroup + c < a.n_ch) a.y_out[(size_t)(group * kChanGroup + c) * a.cap + i0 + o] = ys[e];
            }
        }
        __syncthreads();   // xs free for the next tile's conversion
    }
}

// =================================================================================================
// k_chan_fir_u — the fast channeliser kernel (n_taps <= 255): lanes own OUTPUTS, taps are UNIFORM.
//
// k_chan_fir above is limited by shared-memory wavefronts (every warp re-reads the 130 KB tap table) and
// reaches ~53 % of the FP32 peak.  Here a lane owns one output column and CH = 16 channels live in its
// registers; the folded taps of those 16 channels are a 32 KB __grid_constant__ kernel parameter, so for
// tap k every lane needs the SAME 16 (gr, gi) pairs: they arrive as scalar-broadcast uniform registers
// (LDCU.128 from the constant bank) inside packed FFMA2s — no shared-memory tap traffic at all.  The only
// per-lane load is its own raw sample (one LDS.U16 + PRMT + 2 FHADD per tap, amortised over 32 FFMA2s).
//   A_c += gr_c * (xr, xi);  B_c += gi_c * (xr, xi);   y_c = (A.re - B.im) + j (A.im + B.re)
// (no negation / swap in the loop).  One launch per 16 channels; each re-reads the raw bytes (2 B/sample
// against ~650 FMA/sample: irrelevant).
// =================================================================================================
constexpr int kChanUCh = 8, kChanUMo = 2, kChanUThreads = 128, kChanUMaxT = 255;
constexpr int kChanUOut = kChanUThreads * kChanUMo;   // outputs per CTA
struct TapsU {
    float2 g[kChanUMaxT * kChanUCh];   // [k][channel]: (gr, gi) of the folded tap
};
struct ChanUArgs {
    const uint8_t *x;
    const uint8_t *carry_end;
    const uint32_t *fw;      // [C_pad] NCO phase words
    float2 *y_out;           // [C_pad][cap]
    long long n_samples, n_out, cap;
    uint32_t r, n0_lo;
    int ch0, n_ch, T, D;     // first channel of this launch, channels that exist (< CH for the last group)
};

// A lane owns kChanUMo outputs (tid and tid + 128 of the CTA's 256) x kChanUCh channels: every uniform tap
// load (LDCU.64 of one (gr, gi)) feeds 2*MO FFMA2s, which keeps the constant/MIO queue below the FMA time.
__global__ void __launch_bounds__(kChanUThreads) k_chan_fir_u(const ChanUArgs a, const __grid_constant__ TapsU taps) {
    constexpr int CH = kChanUCh, MO = kChanUMo;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t sh_soff;
    const int tid = threadIdx.x;
    const long long out0 = (long long)blockIdx.x * kChanUOut;
    const long long n_here = a.n_out - out0 < kChanUOut ? a.n_out - out0 : kChanUOut;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        ChanArgs la{};   // chan_load_tile only needs x / carry_end
        la.x = a.x;
        la.carry_end = a.carry_end;
        const long long s0 = out0 * a.D - (long long)a.r - (a.T - 1);
        const long long s1 = (out0 + n_here) * a.D - (long long)a.r;
        sh_soff = chan_load_tile(smem, la, s0, s1, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    // newest sample of output out0 + tid + 128*o, relative to the tile's first sample s0
    const uint16_t *t16[MO];
#pragma unroll
    for (int o = 0; o < MO; o++)
        t16[o] = reinterpret_cast<const uint16_t *>(smem + sh_soff) + (tid + o * kChanUThreads + 1) * a.D + a.T - 2;
    unsigned long long A[MO][CH], B[MO][CH];
#pragma unroll
    for (int o = 0; o < MO; o++)
#pragma unroll
        for (int c = 0; c < CH; c++) A[o][c] = B[o][c] = 0ull;
    // opaque per-thread conversion constants (see fx_path.cu: FHADD takes no immediate / uniform operand)
    float bias;
    uint32_t h1024;
    asm volatile(
        "{\n.reg .u32 t;\nmov.u32 t, %%tid.x;\nshr.u32 t, t, 31;\nor.b32 %0, t, 0xC48FE000;\nor.b32 %1, t, 0x64646464;\n}\n"
        : "=f"(bias), "=r"(h1024));
#pragma unroll 2
    for (int k = 0; k < a.T; k++) {
        unsigned long long x2[MO];
#pragma unroll
        for (int o = 0; o < MO; o++) {
            const uint32_t pair = __byte_perm((uint32_t)t16[o][-k], h1024, 0x4140u);   // half2 (1024+I, 1024+Q)
            float xr, xi;
            asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(xr) : "h"((unsigned short)(pair & 0xffffu)), "f"(bias));
            asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(xi) : "h"((unsigned short)(pair >> 16)), "f"(bias));
            x2[o] = c_pack(xr, xi);
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const float2 g = taps.g[k * CH + c];
#pragma unroll
            for (int o = 0; o < MO; o++) {
                c_fma2(A[o][c], g.x, x2[o]);
                c_fma2(B[o][c], g.y, x2[o]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < MO; o++) {
        const long long i = out0 + tid + o * kChanUThreads;
        if (i >= a.n_out) continue;
        const uint32_t nm = a.n0_lo + (uint32_t)((i + 1) * a.D - 1) - a.r;   // global n_m mod 2^32
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (c >= a.n_ch) break;
            const float2 pa = c_unpack(A[o][c]), pb = c_unpack(B[o][c]);
            const float yr = pa.x - pb.y, yi = pa.y + pb.x;
            // e^{+j theta} from the top 24 phase bits (angle error <= 2 pi 2^-25: far inside the 1e-5 bar);
            // de-rotate with its conjugate
            float co, si;
            sincospif((float)(int32_t)(a.fw[a.ch0 + c] * nm) * (1.0f / 2147483648.0f), &si, &co);
            a.y_out[(size_t)(a.ch0 + c) * a.cap + i] = make_float2(yr * co + yi * si, yi * co - yr * si);
        }
    }
}

// Folded taps: g[group][k][c] = h[k] * e^{+j theta_c(k)} (double precision, setup only).
__global__ void k_chan_fold_taps(const float *h, int T, int T4, const uint32_t *fw, int C, int C_pad, float2 *g) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int total = C_pad * T4;
    if (idx >= total) return;
    int group = idx / (T4 * kChanGroup), rem = idx % (T4 * kChanGroup);
    int k = rem / kChanGroup, cl = rem % kChanGroup, c = group * kChanGroup + cl;
    float2 v = make_float2(0.f, 0.f);
    if (c < C && k < T) {
        uint32_t ph = fw[c] * (uint32_t)k;
        double s, co;
        sincospi(2.0 * ((double)ph / 4294967296.0), &s, &co);
        v = make_float2((float)((double)h[k] * co), (float)((double)h[k] * s));
    }
    g[idx] = v;
}

__device__ __forceinline__ float chan_dop(float a, float b, float c, float d) {
    float cd = c * d;
    float err = fmaf(-c, d, cd);
    return fmaf(a, b, -cd) + err;
}

// Discriminator over [C][cap]; prev[c] carries y_c[m-1] across calls.
__global__ void k_chan_demod(const float2 *y, long long n_out, long long cap, int C, float2 *prev, float gain, float *d) {
    const int c = blockIdx.y;
    if (c >= C) return;
    const float2 *yc = y + (size_t)c * cap;
    const float2 p0 = prev[c];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
        float2 a = yc[i], b = i ? yc[i - 1] : p0;
        float cre = chan_dop(a.x, b.x, -a.y, b.y), cim = chan_dop(a.y, b.x, a.x, b.y);
        d[(size_t)c * cap + i] = (cre == 0.f && cim == 0.f) ? 0.f : gain * atan2f(cim, cre);
    }
}
__global__ void k_chan_store_prev(const float2 *y, long long n_out, long long cap, int C, float2 *prev) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C && n_out > 0) prev[c] = y[(size_t)c * cap + n_out - 1];
}

__global__ void k_chan_update_carry(const uint16_t *old_carry, const uint16_t *x, long long n, int cs, uint16_t *new_carry) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cs; i += gridDim.x * blockDim.x) {
        long long p = n - cs + i;
        new_carry[i] = p >= 0 ? x[p] : old_carry[cs + p];
    }
}

#define SDR_K(x) ((const void *)(x))
static const KernelList kChanKernels{
    SDR_K((k_chan_fir<16, 4, true>)), SDR_K((k_chan_fir<16, 4, false>)), SDR_K((k_chan_fir<16, 2, true>)),
    SDR_K((k_chan_fir<16, 2, false>)), SDR_K((k_chan_fir<16, 1, true>)), SDR_K((k_chan_fir<16, 1, false>)),
    SDR_K((k_chan_fir<8, 8, true>)), SDR_K((k_chan_fir<8, 8, false>)), SDR_K((k_chan_fir<8, 4, true>)),
    SDR_K((k_chan_fir<8, 4, false>)), SDR_K((k_chan_fir<8, 2, true>)), SDR_K((k_chan_fir<8, 2, false>)),
    SDR_K((k_chan_fir<8, 1, true>)), SDR_K((k_chan_fir<8, 1, false>)), SDR_K(k_chan_fir_u), SDR_K(k_chan_bank<1>), SDR_K(k_chan_bank<2>), SDR_K(k_chan_bank<3>),
    SDR_K(k_chan_bank<4>), SDR_K(k_chan_bank<5>), SDR_K(k_chan_bank<6>), SDR_K(k_chan_bank<7>), SDR_K(k_chan_bank<8>),
    SDR_K((k_chan_bank<5, 20, 13>)), SDR_K((k_chan_bank<4, 16, 16>)), SDR_K(k_chan_fold_taps),
    SDR_K(k_chan_demod), SDR_K(k_chan_store_prev), SDR_K(k_chan_update_carry)};
#undef SDR_K

}  // namespace sdr

using namespace sdr;

// =================================================================================================
// Uniform-bank planner (host, f64): is fw[] a uniform grid fw_c = f0 + c * 2^32/K ?  which K = K1*K2 split is cheapest ?
// =================================================================================================
namespace {

struct BankPlan {
    uint32_t K = 0, K1 = 0, K2 = 0, groups = 0;
    double f0 = 0.0;          // ideal NCO word of channel 0 (real-valued: the mean rounding offset is taken out)
    double max_tap_phase_err = 0.0;
    const char *why = "";
};

constexpr double kTwo32 = 4294967296.0;
inline double wrap_words(double x) { return x - kTwo32 * std::nearbyint(x / kTwo32); }   // into [-2^31, 2^31]

// true: the bank kernel applies and beats the direct form; p is filled in.
bool plan_bank(const sdr_chan_config &cfg, const uint32_t *fw, BankPlan &p) {
    const uint32_t C = cfg.n_channels, T = cfg.n_taps;
    if (C < 8) return p.why = "fewer than 8 channels", false;
    const uint32_t step = fw[1] - fw[0];
    uint32_t K = 0;
    for (uint32_t k = 2; k <= 1024 && !K; k++) {
        // k * step == 2^32 (mod 2^32) up to the rounding of the two words involved (each +-0.5 -> the product +-k)
        const uint64_t t = ((uint64_t)k * step) & 0xffffffffull;
        const uint64_t dist = t < (1ull << 32) - t ? t : (1ull << 32) - t;
        if (dist > k) continue;
        // bins must ascend one per channel: s = round(k * step / 2^32) mod k == 1
        const uint64_t s = (uint64_t)std::llround((double)k * (double)step / kTwo32) % k;
        if (s == 1 % k) K = k;
    }
    if (!K) return p.why = "channel spacing is not 2^32/K for any K <= 1024 (one bin per channel)", false;
    // every channel against the ideal grid through channel 0; then take the mean offset out
    double mean = 0.0, worst = 0.0;
    std::vector<double> diff(C);
    for (uint32_t c = 0; c < C; c++) {
        diff[c] = wrap_words((double)(uint32_t)(fw[c] - fw[0]) - (double)c * (kTwo32 / K));
        if (std::fabs(diff[c]) > 4.0) return p.why = "channels are not on one uniform grid", false;
        mean += diff[c] / C;
    }
    for (uint32_t c = 0; c < C; c++) worst = std::max(worst, std::fabs(diff[c] - mean));
    p.max_tap_phase_err = worst * (T > 1 ? T - 1 : 1) * (2.0 * 3.14159265358979323846 / kTwo32);
    if (p.max_tap_phase_err > 2e-6) return p.why = "NCO words deviate from the grid by more than 2e-6 rad over the taps", false;
    p.f0 = (double)fw[0] + mean;
    // cheapest split: T*K2 (stage 1, shared) + 64*K1 (stage 2) complex MACs per output time and 64-channel group
    const uint32_t groups = (C + kBankCH - 1) / kBankCH;
    uint64_t best = ~0ull;
    for (uint32_t k2 = 1; k2 <= (uint32_t)kBankMaxK2; k2++) {
        if (K % k2) continue;
        const uint32_t k1 = K / k2, nj = (T + k1 - 1) / k1;   // every residue gets nj taps (zero padded)
        const uint64_t cost = (uint64_t)k1 * nj * k2 + (uint64_t)kBankCH * k1;
        if (cost + kBankCH + 1 > (uint64_t)kBankTabEntries) continue;
        if (cost < best) best = cost, p.K1 = k1, p.K2 = k2;
    }
    if (best == ~0ull) return p.why = "coefficient tables exceed the 30 KB kernel-parameter blob", false;
    if ((double)best * groups > 0.6 * (double)C * T) return p.why = "no saving over the direct form", false;
    p.K = K;
    p.groups = groups;
    return true;
}

// Coefficient blob of channel group g (channels 64g ...): layout in chan_bank.cuh.
void build_bank_tab(const sdr_chan_config &cfg, const float *taps, const uint32_t *fw, const BankPlan &p, uint32_t g, BankTab &tab) {
    const double PI2 = 2.0 * 3.14159265358979323846;
    const uint32_t T = cfg.n_taps, K = p.K, K1 = p.K1, K2 = p.K2, NJ = (T + K1 - 1) / K1;
    memset(&tab, 0, sizeof(tab));
    const double f0g = p.f0 + (double)g * kBankCH * (kTwo32 / K);   // ideal word of the group's first channel
    size_t idx = 0;
    for (uint32_t r1 = 0; r1 < K1; r1++)
        for (uint32_t j = 0; j < NJ; j++) {
            const uint32_t k = r1 + j * K1;
            const uint32_t r2 = (k % K) / K1;
            const double base = std::fmod(f0g * (double)k / kTwo32, 1.0);
            for (uint32_t b2 = 0; b2 < K2; b2++) {
                const double ph = PI2 * (base + (double)((b2 * r2) % K2) / K2);
                const double h = k < T ? (double)taps[k] : 0.0;   // zero padding up to Tp = K1*NJ taps
                tab.v[idx++] = make_float2((float)(h * std::cos(ph)), (float)(h * std::sin(ph)));
            }
        }
    idx = (idx + 1) & ~size_t(1);   // E starts on an even entry (16-byte aligned pairs)
    // E[r1][c] = e^{j 2 pi c r1 / K}, stored per channel pair as (re c, re c+1), (im c, im c+1) — see stage 2 of k_chan_bank
    for (uint32_t r1 = 0; r1 < K1; r1++)
        for (uint32_t c = 0; c < (uint32_t)kBankCH; c += 2) {
            const double ph0 = PI2 * (double)(((uint64_t)c * r1) % K) / K, ph1 = PI2 * (double)(((uint64_t)(c + 1) * r1) % K) / K;
            tab.v[idx++] = make_float2((float)std::cos(ph0), (float)std::cos(ph1));
            tab.v[idx++] = make_float2((float)std::sin(ph0), (float)std::sin(ph1));
        }
    for (uint32_t c = 0; c < (uint32_t)kBankCH; c++) {
        const uint32_t ch = g * kBankCH + c;
        float phi = 0.f;
        if (ch < cfg.n_channels)   // theta_c(n_m) - theta_c(n_{m-1}) = fw_c * D (mod 2^32), as an angle in (-pi, pi]
            phi = (float)((double)(int32_t)(fw[ch] * cfg.decim) * (PI2 / kTwo32));
        tab.v[idx++] = make_float2(phi, 0.f);
    }
}

}  // namespace

struct sdr_chan {
    sdr_chan_config cfg{};
    int device = 0;
    int C_pad = 0, groups = 0, T4 = 0, pad = 0, mr = 0, warps = 8;
    bool aligned = false;
    bool use_uniform = false;                  // k_chan_fir_u path (n_taps <= 255)
    // uniformly spaced channels: the two-stage polyphase bank (chan_bank.cuh); one coefficient blob per 64 channels
    bool use_bank = false;
    int bank_K = 0, bank_K1 = 0, bank_K2 = 0;
    std::vector<BankTab> bank_tabs;
    size_t smem_bank = 0, smem_bank_tile = 0;
    H2DStager stager;                          // pageable caller buffers go through pinned pieces (common.cuh)
    DevBuf d_prev2;                            // bank path: the carried S[m-1] ping-pongs between d_prev and d_prev2
    int prev_cur = 0;
    std::vector<TapsU> taps_u;                 // one 16 KB parameter blob per 8 channels
    size_t smem_u = 0;
    uint32_t sm_taps = 0, sm_xs = 0, sm_xb = 0;
    size_t smem = 0;
    int cs = 0, carry_cur = 0;
    float gain = 0.f;
    uint64_t n_in = 0;
    DevBuf d_carry[2], d_taps, d_gt, d_fw, d_prev, d_x, d_y, d_d;
    size_t y_cap = 0;
    cudaStream_t stream = nullptr;
    // the channel-group launches of one call are independent: groups 1.. go round-robin to side streams (fork / join
    // on events) so that the partial last wave of one launch is filled by the next launch's first CTAs
    static constexpr int kSide = 3;
    cudaStream_t side[kSide]{};
    cudaEvent_t ev_fork = nullptr, ev_join[kSide]{};
    cudaEvent_t ev[3]{};
    float last_ms = 0.f;
    uint32_t last_launches = 0;
    bool timing_pending = false;
};

namespace {

using ChanKernel = void (*)(const ChanArgs);
// (warps, outputs per warp): the tile is warps*mr outputs; 8x8 and 16x4 cover the same 64-output tile with
// different register / latency-hiding trade-offs (SDR_CHAN_W selects the warp count when tuning).
ChanKernel pick_kernel(int warps, int mr, bool aligned) {
    if (warps == 16) {
        switch (mr) {
            case 4: return aligned ? k_chan_fir<16, 4, true> : k_chan_fir<16, 4, false>;
            case 2: return aligned ? k_chan_fir<16, 2, true> : k_chan_fir<16, 2, false>;
            default: return aligned ? k_chan_fir<16, 1, true> : k_chan_fir<16, 1, false>;
        }
    }
    switch (mr) {
        case 8: return aligned ? k_chan_fir<8, 8, true> : k_chan_fir<8, 8, false>;
        case 4: return aligned ? k_chan_fir<8, 4, true> : k_chan_fir<8, 4, false>;
        case 2: return aligned ? k_chan_fir<8, 2, true> : k_chan_fir<8, 2, false>;
        default: return aligned ? k_chan_fir<8, 1, true> : k_chan_fir<8, 1, false>;
    }
}

int chan_reset_state(sdr_chan *c) {
    for (int i = 0; i < 2; i++) SDR_CUDA_TRY(cudaMemsetAsync(c->d_carry[i].p, 127, (size_t)c->cs * 2, c->stream));
    SDR_CUDA_TRY(cudaMemsetAsync(c->d_prev.p, 0, (size_t)c->C_pad * 8, c->stream));
    if (c->d_prev2.p) SDR_CUDA_TRY(cudaMemsetAsync(c->d_prev2.p, 0, (size_t)c->C_pad * 8, c->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->carry_cur = 0;
    c->prev_cur = 0;
    c->n_in = 0;
    return SDR_OK;
}

// y (and demod) for one call whose input is resident; d_y / d_d are [C][cap].
int chan_run(sdr_chan *c, const uint8_t *d_x, size_t n, float2 *d_y, float *d_d, size_t cap, uint64_t n_out) {
    const int D = (int)c->cfg.decim;
    c->last_launches = 0;
    SDR_CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    if (c->use_bank) {
        if (n_out) {
            BankArgs b{};
            b.x = d_x;
            b.carry_end = c->d_carry[c->carry_cur].as<uint8_t>() + (size_t)c->cs * 2;
            b.fw = c->d_fw.as<uint32_t>();
            b.y_out = d_y;
            b.d_out = d_d;
            b.prev_in = (c->prev_cur ? c->d_prev2 : c->d_prev).as<float2>();
            b.prev_out = (c->prev_cur ? c->d_prev : c->d_prev2).as<float2>();
            b.n_samples = (long long)n;
            b.n_out = (long long)n_out;
            b.cap = (long long)cap;
            b.r = (uint32_t)(c->n_in % D);
            b.n0_lo = (uint32_t)c->n_in;
            b.D = D;
            b.K1 = c->bank_K1;
            b.NJ = ((int)c->cfg.n_taps + b.K1 - 1) / b.K1;
            b.Tp = b.K1 * b.NJ;
            b.eoff = (b.Tp * c->bank_K2 + 1) & ~1;
            b.gain = c->gain;
            const uint64_t tiles = (n_out + (kBankThreads - 1) - 1) / (kBankThreads - 1);
            if (tiles > 0x7fffffffull) return fail(SDR_E_ARG, "call too large");
            const int n_streams = (int)std::min<size_t>(c->bank_tabs.size(), 1 + sdr_chan::kSide);
            if (n_streams > 1) {
                SDR_CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
                for (int s = 1; s < n_streams; s++) SDR_CUDA_TRY(cudaStreamWaitEvent(c->side[s - 1], c->ev_fork, 0));
            }
            for (size_t g = 0; g < c->bank_tabs.size(); g++) {
                b.ch0 = (int)g * kBankCH;
                b.n_ch = std::min<int>(kBankCH, (int)c->cfg.n_channels - b.ch0);
                const int si = (int)(g % n_streams);
                cudaStream_t st = si ? c->side[si - 1] : c->stream;
                const unsigned grid = (unsigned)tiles;
                if (b.K1 == 20 && c->bank_K2 == 5 && b.NJ == 13)        // cfg4: 255 taps on a 100-bin grid
                    k_chan_bank<5, 20, 13><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]);
                else if (b.K1 == 16 && c->bank_K2 == 4 && b.NJ == 16)   // interleaved cfg5: 255 taps on a 64-bin grid
                    k_chan_bank<4, 16, 16><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]);
                else
                switch (c->bank_K2) {
                    case 1: k_chan_bank<1><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]); break;
                    case 2: k_chan_bank<2><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]); break;
                    case 3: k_chan_bank<3><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]); break;
                    case 4: k_chan_bank<4><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]); break;
                    case 5: k_chan_bank<5><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]); break;
                    case 6: k_chan_bank<6><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]); break;
                    case 7: k_chan_bank<7><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]); break;
                    default: k_chan_bank<8><<<grid, kBankThreads, c->smem_bank, st>>>(b, c->bank_tabs[g]); break;
                }
                SDR_LAUNCH_CHECK();
                c->last_launches++;
            }
            for (int s = 1; s < n_streams; s++) {
                SDR_CUDA_TRY(cudaEventRecord(c->ev_join[s - 1], c->side[s - 1]));
                SDR_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join[s - 1], 0));
            }
            c->prev_cur ^= 1;
        }
        SDR_CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        if (n) {
            int nxt = c->carry_cur ^ 1;
            k_chan_update_carry<<<(c->cs + 255) / 256, 256, 0, c->stream>>>(c->d_carry[c->carry_cur].as<uint16_t>(),
                                                                            reinterpret_cast<const uint16_t *>(d_x), (long long)n,
                                                                            c->cs, c->d_carry[nxt].as<uint16_t>());
            SDR_LAUNCH_CHECK();
            c->last_launches++;
            c->carry_cur = nxt;
        }
        SDR_CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        c->timing_pending = true;
        c->n_in += n;
        return SDR_OK;
    }
    if (n_out && c->use_uniform) {
        ChanUArgs u{};
        u.x = d_x;
        u.carry_end = c->d_carry[c->carry_cur].as<uint8_t>() + (size_t)c->cs * 2;
        u.fw = c->d_fw.as<uint32_t>();
        u.y_out = d_y;
        u.n_samples = (long long)n;
        u.n_out = (long long)n_out;
        u.cap = (long long)cap;
        u.r = (uint32_t)(c->n_in % D);
        u.n0_lo = (uint32_t)c->n_in;
        u.T = (int)c->cfg.n_taps;
        u.D = D;
        const uint64_t tiles = (n_out + kChanUOut - 1) / kChanUOut;
        if (tiles > 0x7fffffffull) return fail(SDR_E_ARG, "call too large");
        const int n_streams = (int)std::min<size_t>(c->taps_u.size(), 1 + sdr_chan::kSide);
        if (n_streams > 1) {
            SDR_CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
            for (int s = 1; s < n_streams; s++) SDR_CUDA_TRY(cudaStreamWaitEvent(c->side[s - 1], c->ev_fork, 0));
        }
        for (size_t g = 0; g < c->taps_u.size(); g++) {
            u.ch0 = (int)g * kChanUCh;
            u.n_ch = std::min<int>(kChanUCh, (int)c->cfg.n_channels - u.ch0);
            const int s = (int)(g % n_streams);
            k_chan_fir_u<<<(unsigned)tiles, kChanUThreads, c->smem_u, s ? c->side[s - 1] : c->stream>>>(u, c->taps_u[g]);
            SDR_LAUNCH_CHECK();
            c->last_launches++;
        }
        for (int s = 1; s < n_streams; s++) {
            SDR_CUDA_TRY(cudaEventRecord(c->ev_join[s - 1], c->side[s - 1]));
            SDR_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join[s - 1], 0));
        }
    } else if (n_out) {
        const int MT = c->warps * c->mr;
        ChanArgs a{};
        a.x = d_x;
        a.carry_end = c->d_carry[c->carry_cur].as<uint8_t>() + (size_t)c->cs * 2;
        a.n_samples = (long long)n;
        a.n_out = (long long)n_out;
        a.r = (uint32_t)(c->n_in % D);
        a.n0_lo = (uint32_t)c->n_in;
        a.gtaps = c->d_gt.as<float2>();
        a.fw = c->d_fw.as<uint32_t>();
        a.y_out = d_y;
        a.cap = (long long)cap;
        a.T4 = c->T4;
        a.D = D;
        a.pad = c->pad;
        uint64_t tiles = (n_out + MT - 1) / MT;
        if (tiles > 0x7fffffffull) return fail(SDR_E_ARG, "call too large");
        a.n_tiles = (int)tiles;
        a.n_ch = (int)c->cfg.n_channels;
        a.sm_taps = c->sm_taps;
        a.sm_xs = c->sm_xs;
        a.sm_xb = c->sm_xb;
        int ctas = std::max(1, sm_count(c->device) / c->groups);
        if ((uint64_t)ctas > tiles) ctas = (int)tiles;
        dim3 grid(ctas, c->groups);
        pick_kernel(c->warps, c->mr, c->aligned)<<<grid, c->warps * 32, c->smem, c->stream>>>(a);
        SDR_LAUNCH_CHECK();
        c->last_launches++;
    }
    SDR_CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    if (n_out && d_d) {
        dim3 grid((unsigned)std::min<uint64_t>((n_out + 255) / 256, 64), c->cfg.n_channels);
        k_chan_demod<<<grid, 256, 0, c->stream>>>(d_y, (long long)n_out, (long long)cap, (int)c->cfg.n_channels,
                                                  c->d_prev.as<float2>(), c->gain, d_d);
        SDR_LAUNCH_CHECK();
        c->last_launches++;
    }
    if (n_out) {
        k_chan_store_prev<<<(c->cfg.n_channels + 127) / 128, 128, 0, c->stream>>>(d_y, (long long)n_out, (long long)cap,
                                                                                  (int)c->cfg.n_channels, c->d_prev.as<float2>());
        SDR_LAUNCH_CHECK();
        c->last_launches++;
    }
    if (n) {
        int nxt = c->carry_cur ^ 1;
        k_chan_update_carry<<<(c->cs + 255) / 256, 256, 0, c->stream>>>(c->d_carry[c->carry_cur].as<uint16_t>(),
                                                                        reinterpret_cast<const uint16_t *>(d_x), (long long)n,
                                                                        c->cs, c->d_carry[nxt].as<uint16_t>());
        SDR_LAUNCH_CHECK();
        c->last_launches++;
        c->carry_cur = nxt;
    }
    SDR_CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
    c->timing_pending = true;
    c->n_in += n;
    return SDR_OK;
}

void chan_collect(sdr_chan *c) {
    if (!c->timing_pending) return;
    c->timing_pending = false;
    float ms = 0.f;
    if (cudaEventSynchronize(c->ev[2]) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess)
        c->last_ms = ms;
}

}  // namespace

extern "C" {

int sdr_chan_new(const sdr_chan_config *cfg, const float *taps, const uint32_t *freq_words, int cuda_device, sdr_chan **out) {
    if (!cfg || !taps || !freq_words || !out) return fail(SDR_E_ARG, "sdr_chan_new: null argument");
    if (cfg->n_channels < 1 || cfg->n_taps < 1 || cfg->decim < 1) return fail(SDR_E_ARG, "need n_channels, n_taps, decim >= 1");
    int rc = use_device(cuda_device);
    if (rc) return rc;
    sdr_chan *c = new sdr_chan();
    c->cfg = *cfg;
    c->device = cuda_device;
    c->gain = cfg->gain != 0.f ? cfg->gain : (float)(16384.0 / 3.14159265358979323846);
    c->groups = (int)((cfg->n_channels + kChanGroup - 1) / kChanGroup);
    c->C_pad = c->groups * kChanGroup;
    c->T4 = (int)((cfg->n_taps + 3) & ~3u);
    const int D = (int)cfg->decim, T4 = c->T4;
    c->aligned = (D % 2 == 0);
    // aligned mode wants the newest sample index (o+1)*D + T4 - 2 + pad odd; D even, T4 even => pad = 1
    c->pad = c->aligned ? 1 : 0;
    c->sm_taps = (uint32_t)((size_t)T4 * kChanGroup * 8);
    // pick the largest MR whose tile fits beside the resident taps
    const size_t budget = 225 * 1024;
    c->mr = 0;
    const char *w_env = getenv("SDR_CHAN_W");
    c->warps = (w_env && atoi(w_env) == 8) ? 8 : 16;   // 16 warps x 4 outputs measured ~3% faster than 8 x 8
    for (int mr : {8, 4, 2, 1}) {
        if (c->warps == 16 && mr == 8) continue;
        size_t MT = (size_t)c->warps * mr;
        size_t ns = MT * D + T4 - 1 + c->pad;
        size_t xs = std::max(ns * 8, (size_t)kChanGroup * MT * 8);
        xs = (xs + 15) & ~size_t(15);
        size_t xb = ((ns * 2 + 15) & ~size_t(15)) + 32;
        if (c->sm_taps + xs + 2 * xb <= budget) {
            c->mr = mr;
            c->sm_xs = (uint32_t)xs;
            c->sm_xb = (uint32_t)xb;
            c->smem = c->sm_taps + xs + 2 * xb;
            break;
        }
    }
    if (!c->mr) {
        delete c;
        return fail(SDR_E_ARG, "n_taps=%u / decim=%u do not fit the channeliser's shared-memory tile", cfg->n_taps, cfg->decim);
    }
    c->cs = (int)(((size_t)T4 + D + 16 + 7) & ~size_t(7));
    {
        // uniformly spaced channels: the two-stage polyphase bank (SDR_CHAN_BANK=0 keeps the direct kernels, for A/B runs)
        const char *eb = getenv("SDR_CHAN_BANK");
        BankPlan bp;
        if (!(eb && atoi(eb) == 0) && plan_bank(*cfg, freq_words, bp)) {
            c->use_bank = true;
            c->bank_K = (int)bp.K, c->bank_K1 = (int)bp.K1, c->bank_K2 = (int)bp.K2;
            c->bank_tabs.resize(bp.groups);
            for (uint32_t g = 0; g < bp.groups; g++) build_bank_tab(*cfg, taps, freq_words, bp, g, c->bank_tabs[g]);
            // the halo lane of the first tile reaches back D + r + Tp - 1 samples before the call (r < D, Tp < T + K1)
            const size_t Tp = (size_t)bp.K1 * ((cfg->n_taps + bp.K1 - 1) / bp.K1);
            c->cs = (int)((Tp + 2 * (size_t)D + 16 + 7) & ~size_t(7));
            c->smem_bank_tile = (((size_t)kBankThreads * D + Tp + 16) * 2 + 15 + 32) & ~size_t(15);
            c->smem_bank = c->smem_bank_tile;
            if (c->smem_bank > 200 * 1024) c->use_bank = false;
        }
    }
    cudaError_t e = raise_dyn_smem(pick_kernel(c->warps, c->mr, c->aligned), c->smem);
    if (e == cudaSuccess && c->use_bank) {
        if (e == cudaSuccess) e = raise_dyn_smem(k_chan_bank<5, 20, 13>, c->smem_bank);
        if (e == cudaSuccess) e = raise_dyn_smem(k_chan_bank<4, 16, 16>, c->smem_bank);
        if (e == cudaSuccess)
        switch (c->bank_K2) {
            case 1: e = raise_dyn_smem(k_chan_bank<1>, c->smem_bank); break;
            case 2: e = raise_dyn_smem(k_chan_bank<2>, c->smem_bank); break;
            case 3: e = raise_dyn_smem(k_chan_bank<3>, c->smem_bank); break;
            case 4: e = raise_dyn_smem(k_chan_bank<4>, c->smem_bank); break;
            case 5: e = raise_dyn_smem(k_chan_bank<5>, c->smem_bank); break;
            case 6: e = raise_dyn_smem(k_chan_bank<6>, c->smem_bank); break;
            case 7: e = raise_dyn_smem(k_chan_bank<7>, c->smem_bank); break;
            default: e = raise_dyn_smem(k_chan_bank<8>, c->smem_bank); break;
        }
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 3 && e == cudaSuccess; i++) e = cudaEventCreate(&c->ev[i]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    for (int i = 0; i < sdr_chan::kSide && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        sdr_chan_free(c);
        return fail(SDR_E_CUDA, "sdr_chan_new: %s", cudaGetErrorString(e));
    }
    if ((rc = c->d_carry[0].reserve((size_t)c->cs * 2)) || (rc = c->d_carry[1].reserve((size_t)c->cs * 2)) ||
        (rc = c->d_taps.reserve(cfg->n_taps * 4)) || (rc = c->d_fw.reserve((size_t)c->C_pad * 4)) ||
        (rc = c->d_gt.reserve((size_t)c->C_pad * T4 * 8)) || (rc = c->d_prev.reserve((size_t)c->C_pad * 8)) ||
        (c->use_bank && (rc = c->d_prev2.reserve((size_t)c->C_pad * 8)))) {
        sdr_chan_free(c);
        return rc;
    }
    std::vector<uint32_t> fw(c->C_pad, 0u);
    std::copy(freq_words, freq_words + cfg->n_channels, fw.begin());
    e = cudaMemcpyAsync(c->d_taps.p, taps, cfg->n_taps * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_fw.p, fw.data(), fw.size() * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        sdr_chan_free(c);
        return fail(SDR_E_CUDA, "sdr_chan_new: %s", cudaGetErrorString(e));
    }
    // uniform-tap kernel: folded taps g_c[k] = h[k] e^{+j theta_c(k)} as kernel-parameter blobs (setup, host f64)
    const char *force_smem = getenv("SDR_CHAN_SMEM_TAPS");
    if (cfg->n_taps <= (uint32_t)kChanUMaxT && !(force_smem && atoi(force_smem))) {
        c->use_uniform = true;
        const size_t n_groups = (cfg->n_channels + kChanUCh - 1) / kChanUCh;
        c->taps_u.assign(n_groups, TapsU{});
        for (uint32_t ch = 0; ch < cfg->n_channels; ch++)
            for (uint32_t k = 0; k < cfg->n_taps; k++) {
                const uint32_t ph = freq_words[ch] * k;   // mod 2^32
                const double th = 2.0 * 3.14159265358979323846 * ((double)ph / 4294967296.0);
                c->taps_u[ch / kChanUCh].g[(size_t)k * kChanUCh + ch % kChanUCh] =
                    make_float2((float)((double)taps[k] * std::cos(th)), (float)((double)taps[k] * std::sin(th)));
            }
        c->smem_u = (((size_t)kChanUOut * D + cfg->n_taps + 16) * 2 + 15 + 32) & ~size_t(15);
        if (c->smem_u > 200 * 1024) c->use_uniform = false;
        else e = raise_dyn_smem(k_chan_fir_u, c->smem_u);
        if (e != cudaSuccess) {
            sdr_chan_free(c);
            return fail(SDR_E_CUDA, "sdr_chan_new: %s", cudaGetErrorString(e));
        }
    }
    int total = c->C_pad * T4;
    k_chan_fold_taps<<<(total + 255) / 256, 256, 0, c->stream>>>(c->d_taps.as<float>(), (int)cfg->n_taps, T4, c->d_fw.as<uint32_t>(),
                                                                (int)cfg->n_channels, c->C_pad, c->d_gt.as<float2>());
    if (cudaGetLastError() != cudaSuccess || (rc = chan_reset_state(c))) {
        sdr_chan_free(c);
        return rc ? rc : fail(SDR_E_CUDA, "sdr_chan_new: tap folding kernel failed");
    }
    count_launch();
    *out = c;
    return SDR_OK;
}

void sdr_chan_free(sdr_chan *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 2; i++) c->d_carry[i].release();
    c->d_taps.release();
    c->d_gt.release();
    c->d_fw.release();
    c->stager.release();
    c->d_prev.release();
    c->d_prev2.release();
    c->d_x.release();
    c->d_y.release();
    c->d_d.release();
    for (int i = 0; i < 3; i++)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < sdr_chan::kSide; i++) {
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int sdr_chan_reset(sdr_chan *c) {
    if (!c) return fail(SDR_E_ARG, "null handle");
    int rc = use_device(c->device);
    if (rc) return rc;
    return chan_reset_state(c);
}

long sdr_chan_process(sdr_chan *c, const uint8_t *iq, size_t n_samples, float *y_pairs, float *demod, size_t cap) {
    if (!c || (!iq && n_samples) || !demod) return fail(SDR_E_ARG, "sdr_chan_process: null argument");
    int rc = use_device(c->device);
    if (rc) return rc;
    const uint64_t D = c->cfg.decim;
    const uint64_t n_out = (c->n_in + n_samples) / D - c->n_in / D;
    if (n_out > cap) return fail(SDR_E_CAP, "capacity %zu < %llu outputs per channel", cap, (unsigned long long)n_out);
    if (n_samples == 0) return 0;
    const size_t C = c->cfg.n_channels;
    const bool need_y = y_pairs || !c->use_bank;   // the bank kernel discriminates on chip: y is optional there
    if ((rc = c->d_x.reserve(n_samples * 2 + 64)) || (need_y && (rc = c->d_y.reserve((size_t)c->C_pad * (n_out + 1) * 8))) ||
        (rc = c->d_d.reserve(C * (n_out + 1) * 4)))
        return rc;
    if ((rc = c->stager.copy(c->d_x.p, iq, n_samples * 2, c->stream))) return rc;
    const size_t dcap = n_out ? n_out : 1;
    if ((rc = chan_run(c, c->d_x.as<uint8_t>(), n_samples, need_y ? c->d_y.as<float2>() : nullptr, c->d_d.as<float>(), dcap, n_out)))
        return rc;
    if (n_out) {
        if (y_pairs)
            SDR_CUDA_TRY(cudaMemcpy2DAsync(y_pairs, cap * 8, c->d_y.p, dcap * 8, n_out * 8, C, cudaMemcpyDeviceToHost, c->stream));
        SDR_CUDA_TRY(cudaMemcpy2DAsync(demod, cap * 4, c->d_d.p, dcap * 4, n_out * 4, C, cudaMemcpyDeviceToHost, c->stream));
    }
    SDR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    chan_collect(c);
    return (long)n_out;
}

long sdr_chan_process_dev(sdr_chan *c, const uint8_t *d_iq, size_t n_samples, float *d_y_pairs, float *d_demod, size_t cap) {
    if (!c || (!d_iq && n_samples)) return fail(SDR_E_ARG, "sdr_chan_process_dev: null argument");
    if (reinterpret_cast<uintptr_t>(d_iq) & 15) return fail(SDR_E_ARG, "device input must be 16-byte aligned (use sdr_dev_alloc)");
    int rc = use_device(c->device);
    if (rc) return rc;
    const uint64_t D = c->cfg.decim;
    const uint64_t n_out = (c->n_in + n_samples) / D - c->n_in / D;
    if (n_out > cap) return fail(SDR_E_CAP, "capacity %zu < %llu outputs per channel", cap, (unsigned long long)n_out);
    float2 *d_y = reinterpret_cast<float2 *>(d_y_pairs);
    size_t ycap = cap;
    if (!d_y && !c->use_bank) {   // caller does not want y: keep it in a library buffer with the caller's row stride
        if (c->d_y.cap < (size_t)c->C_pad * cap * 8) SDR_CUDA_TRY(cudaStreamSynchronize(c->stream));   // before a realloc
        if ((rc = c->d_y.reserve((size_t)c->C_pad * cap * 8))) return rc;
        d_y = c->d_y.as<float2>();
    }
    if ((rc = chan_run(c, d_iq, n_samples, d_y, d_demod, ycap, n_out))) return rc;
    return (long)n_out;
}

int sdr_chan_sync(sdr_chan *c) {
    if (!c) return fail(SDR_E_ARG, "null handle");
    int rc = use_device(c->device);
    if (rc) return rc;
    SDR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    chan_collect(c);
    return SDR_OK;
}

int sdr_chan_kernel_kind(const sdr_chan *c, uint32_t info[4]) {
    if (!c) return fail(SDR_E_ARG, "null handle");
    if (info) {
        info[0] = (uint32_t)c->bank_K, info[1] = (uint32_t)c->bank_K1, info[2] = (uint32_t)c->bank_K2;
        info[3] = (uint32_t)c->bank_tabs.size();
    }
    return c->use_bank ? 2 : (c->use_uniform ? 1 : 0);
}

long sdr_chan_bank_plan(const sdr_chan_config *cfg, const float *taps, const uint32_t *freq_words, uint32_t info[4],
                        float *tables, size_t cap_floats) {
    if (!cfg || !taps || !freq_words) return fail(SDR_E_ARG, "sdr_chan_bank_plan: null argument");
    if (cfg->n_channels < 1 || cfg->n_taps < 1 || cfg->decim < 1) return fail(SDR_E_ARG, "need n_channels, n_taps, decim >= 1");
    BankPlan bp;
    if (!plan_bank(*cfg, freq_words, bp)) {
        fail(SDR_OK, "not a uniform bank: %s", bp.why);
        return 0;
    }
    if (info) info[0] = bp.K, info[1] = bp.K1, info[2] = bp.K2, info[3] = bp.groups;
    if (tables) {
        const size_t per = (size_t)kBankTabEntries * 2;
        if (cap_floats < per * bp.groups) return fail(SDR_E_CAP, "table capacity %zu < %zu floats", cap_floats, per * bp.groups);
        BankTab tab;
        for (uint32_t g = 0; g < bp.groups; g++) {
            build_bank_tab(*cfg, taps, freq_words, bp, g, tab);
            memcpy(tables + per * g, tab.v, sizeof(tab.v));
        }
    }
    return (long)bp.groups;
}

int sdr_chan_last_timing(const sdr_chan *c, float *kernel_ms, uint32_t *n_launches) {
    if (!c) return fail(SDR_E_ARG, "null handle");
    if (kernel_ms) *kernel_ms = c->last_ms;
    if (n_launches) *n_launches = c->last_launches;
    return SDR_OK;
}

}  // extern "C"

// =================================================================================================
// NCCL plumbing: one ncclBroadcast(u8) per raw slab and nothing else.  libnccl is dlopen()ed lazily so
// the library loads on boxes without NCCL and shares the copy a host process (e.g. torch) already loaded.
// =================================================================================================
namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclUint8 = 1;   // ncclUint8 / ncclChar group: ncclInt8=0, ncclUint8=1

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
    }
    if (!api.h) return api;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
    api.Broadcast = (decltype(api.Broadcast))dlsym(api.h, "ncclBroadcast");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.GetErrorString;
    return api;
}

int nccl_fail(const char *what, ncclResult_t r) {
    return fail(SDR_E_NCCL, "%s failed: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
}

}  // namespace

struct sdr_comm {
    int device = 0, rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_bcast = nullptr, ev_chan = nullptr;
    static constexpr int kMarks = 32;
    cudaEvent_t ev_mark[kMarks]{};   // per-slab "this slab's memory has been consumed" marks (created on first use)
};

extern "C" {

int sdr_comm_unique_id(uint8_t id[SDR_NCCL_ID_BYTES]) {
    if (!id) return fail(SDR_E_ARG, "null id");
    if (!nccl().ok) return fail(SDR_E_NCCL, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == SDR_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    ncclResult_t r = nccl().GetUniqueId(&u);
    if (r) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id, &u, sizeof(u));
    return SDR_OK;
}

int sdr_comm_init(int cuda_device, int rank, int world, const uint8_t id[SDR_NCCL_ID_BYTES], sdr_comm **out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(SDR_E_ARG, "sdr_comm_init: bad argument");
    if (!nccl().ok) return fail(SDR_E_NCCL, "libnccl.so.2 could not be loaded");
    int rc = use_device(cuda_device);
    if (rc) return rc;
    sdr_comm *c = new sdr_comm();
    c->device = cuda_device;
    c->rank = rank;
    c->world = world;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, u, rank);
    if (r) {
        delete c;
        return nccl_fail("ncclCommInitRank", r);
    }
    // The broadcast kernel's few CTAs must be placed as soon as an SM has room: at equal priority a kernel that arrives
    // while the channeliser's grid is being dispatched gets only that grid's tail, and whether broadcast(s+1) or
    // channeliser(s) arrives first depends on how far the host thread runs ahead (three 8-GPU sessions: 1.8-2.4 ms per
    // step).  Highest priority makes the overlap independent of the arrival order.  SDR_COMM_PRIO=0: default priority.
    int prio_lo = 0, prio_hi = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const char *ep = getenv("SDR_COMM_PRIO");
    if (e == cudaSuccess)
        e = (ep && atoi(ep) == 0) ? cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)
                                  : cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_bcast, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_chan, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        sdr_comm_free(c);
        return fail(SDR_E_CUDA, "sdr_comm_init: %s", cudaGetErrorString(e));
    }
    *out = c;
    return SDR_OK;
}

int sdr_comm_bcast_u8(sdr_comm *c, uint8_t *d_buf, size_t bytes, int root) {
    if (!c || !d_buf) return fail(SDR_E_ARG, "sdr_comm_bcast_u8: null argument");
    int rc = use_device(c->device);
    if (rc) return rc;
    ncclResult_t r = nccl().Broadcast(d_buf, d_buf, bytes, kNcclUint8, root, c->comm, c->stream);
    if (r) return nccl_fail("ncclBroadcast", r);
    SDR_CUDA_TRY(cudaEventRecord(c->ev_bcast, c->stream));
    return SDR_OK;
}

int sdr_comm_chan_wait(sdr_comm *c, sdr_chan *ch) {
    if (!c || !ch) return fail(SDR_E_ARG, "null argument");
    SDR_CUDA_TRY(cudaStreamWaitEvent(ch->stream, c->ev_bcast, 0));
    return SDR_OK;
}

int sdr_comm_wait_chan(sdr_comm *c, sdr_chan *ch) {
    if (!c || !ch) return fail(SDR_E_ARG, "null argument");
    SDR_CUDA_TRY(cudaEventRecord(c->ev_chan, ch->stream));
    SDR_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_chan, 0));
    return SDR_OK;
}

int sdr_comm_mark_chan(sdr_comm *c, sdr_chan *ch, uint32_t slot) {
    if (!c || !ch || slot >= (uint32_t)sdr_comm::kMarks) return fail(SDR_E_ARG, "sdr_comm_mark_chan: bad argument");
    int rc = use_device(c->device);
    if (rc) return rc;
    if (!c->ev_mark[slot]) SDR_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_mark[slot], cudaEventDisableTiming));
    SDR_CUDA_TRY(cudaEventRecord(c->ev_mark[slot], ch->stream));
    return SDR_OK;
}

int sdr_comm_wait_mark(sdr_comm *c, uint32_t slot) {
    if (!c || slot >= (uint32_t)sdr_comm::kMarks) return fail(SDR_E_ARG, "sdr_comm_wait_mark: bad argument");
    if (!c->ev_mark[slot]) return SDR_OK;   // never marked: nothing to wait for
    SDR_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_mark[slot], 0));
    return SDR_OK;
}

int sdr_comm_sync(sdr_comm *c) {
    if (!c) return fail(SDR_E_ARG, "null handle");
    int rc = use_device(c->device);
    if (rc) return rc;
    SDR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SDR_OK;
}

void sdr_comm_free(sdr_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    if (c->ev_bcast) cudaEventDestroy(c->ev_bcast);
    if (c->ev_chan) cudaEventDestroy(c->ev_chan);
    for (cudaEvent_t e : c->ev_mark)
        if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

}  // extern "C"
