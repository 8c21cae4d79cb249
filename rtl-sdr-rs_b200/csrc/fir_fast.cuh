// fir_fast.cuh — the fused u8->f32 convert + decimating FIR + discriminator kernel k_fir_fast<T,D,B,NT,WB,PH>.
//
// This file is compiled twice: by nvcc into libsdr_b200.so for the shapes in fx_path.cu's table, and at run time by
// NVRTC (csrc/rtc.cpp) for any other (taps, decimation) a caller asks for — same source, same arithmetic, so a
// run-time-compiled shape is bit-identical to a pre-compiled one.  Keep it free of host-only headers.
#pragma once
#include "ptx_helpers.cuh"

namespace sdr {

struct FirArgs {
    const uint8_t *x;          // call input, 16-B aligned
    const uint8_t *carry_end;  // one past the last carried byte (16-B aligned); carry holds the samples before x
    long long n_samples;       // samples in x
    long long n_out;           // outputs this call produces
    uint32_t r;                // samples of the current decimation block already consumed before x
    float gain;
    float2 *y_out;             // optional [n_out]
    float *d_out;              // optional [n_out]
    float2 *last_y;            // y of the last output of the call (stage-level fm_demod state)
    // folded bookkeeping (either may be null): the stream carry for the next call = the last `cs` samples of
    // (old carry ++ x), written by the last CTA; the resampler history of the next call = the last `h2` discriminator
    // outputs, written next to d_out by the threads that produce them (needs n_out >= h2)
    uint16_t *carry_out;
    float *hist_out;
    int cs, h2;
};

// new_carry = last cs samples of (old_carry ++ x[0..n)); both carries hold exactly cs samples (u16 each).  Executed by
// one CTA of the FIR kernel (the carry it writes is the OTHER ping-pong buffer: no CTA of this launch reads it).
__device__ __forceinline__ void fold_carry_update(const FirArgs &a, const int tid, const int nthreads) {
    const uint16_t *x16 = reinterpret_cast<const uint16_t *>(a.x);
    const uint16_t *old = reinterpret_cast<const uint16_t *>(a.carry_end) - a.cs;
    for (int i = tid; i < a.cs; i += nthreads) {
        const long long p = a.n_samples - a.cs + i;
        a.carry_out[i] = p >= 0 ? x16[p] : old[a.cs + p];
    }
}

template <int T>
struct Taps {
    float h[T];
};

// Accurate 2x2 determinant / dot (Kahan): a*b - c*d with one rounding error of the result.
__device__ __forceinline__ float diff_of_products(float a, float b, float c, float d) {
    float cd = c * d;
    float err = fmaf(-c, d, cd);
    float dop = fmaf(a, b, -cd);
    return dop + err;
}
// Polar discriminator.  Plain fma products (one rounding each: the angle they feed is good to ~1e-7 rad, two orders inside
// the 1e-5 bar — the compensated products and libm atan2f of round 1 cost 58 instructions per output, 9 % of the fused
// kernel) and a polynomial atan2 (ptx_helpers.cuh, 1.2e-7 rad).
__device__ __forceinline__ float discriminate(float2 y, float2 p, float gain) {
    const float cre = fmaf(y.x, p.x, y.y * p.y);         // y.re*p.re + y.im*p.im
    const float cim = fmaf(y.y, p.x, -(y.x * p.y));      // y.im*p.re - y.re*p.im
    if (cre == 0.f && cim == 0.f) return 0.f;            // zero predecessor (stream start): 0 by definition, not +-pi
    return gain * poly_atan2(cim, cre);
}

// Stage one CTA tile [s0, s1) (call-local sample indices, s0 may be negative = carry) into smem.
// Returns the byte offset of sample s0 inside `tile`.  Executed by one thread.
__device__ __forceinline__ uint32_t load_tile(unsigned char *tile, const FirArgs &a, long long s0, long long s1,
                                              uint64_t *bar) {
    const uint32_t soff = (uint32_t)((2 * s0) & 15);   // two's complement & 15 == positive mod 16
    uint32_t total = 0;
    long long x_lo = s0 > 0 ? s0 : 0;
    uint32_t carry_bytes = 0, x_bytes = 0;
    if (s0 < 0) {
        long long c_hi = s1 < 0 ? s1 : 0;                       // carry part is [s0, c_hi)
        carry_bytes = (uint32_t)((2 * (c_hi - s0) + soff + 15) & ~15ll);
    }
    if (s1 > 0) {
        long long b_lo = (2 * x_lo) & ~15ll;
        long long b_hi = (2 * s1 + 15) & ~15ll;
        x_bytes = (uint32_t)(b_hi - b_lo);
    }
    total = carry_bytes + x_bytes;
    mbar_arrive_expect_tx(bar, total);
    if (carry_bytes) bulk_g2s(tile, a.carry_end + 2 * s0 - soff, carry_bytes, bar);
    if (x_bytes) {
        long long b_lo = (2 * x_lo) & ~15ll;
        // smem position of x byte b_lo: soff + (b_lo - 2*s0)
        bulk_g2s_stream(tile + soff + (b_lo - 2 * s0), a.x + b_lo, x_bytes, bar);
    }
    return soff;
}

// Same tile with PAD bytes of shared memory skipped after every ROW bytes (a row = the bytes one thread owns): when ROW is a
// multiple of 128 every thread's loads would otherwise hit the same bank.  Every thread of the CTA moves 16-byte pieces
// with cp.async (LDGSTS) and then lets the mbarrier count its own copies (cp.async.mbarrier.arrive.noinc: the barrier is
// initialised with NT arrivals and no byte count).  ROW and carry_bytes are multiples of 16, so every piece lies wholly in
// one row and wholly in the carry or in the call input.  (Round 1 issued one bulk copy per row from warp 0: with 32-byte
// rows the whole CTA waited behind 288 serialised copies — (129,/16) 0.65 -> 1.92 TB/s, (201,/64) 3.2 -> 4.2.)
template <int ROW, int PAD, int NT>
__device__ __forceinline__ uint32_t load_tile_rows_async(unsigned char *tile, const FirArgs &a, long long s0, long long s1,
                                                         uint64_t *bar, const int tid) {
    static_assert(ROW % 16 == 0 && PAD % 16 == 0, "16-byte pieces");
    const uint32_t soff = (uint32_t)((2 * s0) & 15);
    long long x_lo = s0 > 0 ? s0 : 0;
    uint32_t carry_bytes = 0, x_bytes = 0;
    if (s0 < 0) {
        long long c_hi = s1 < 0 ? s1 : 0;
        carry_bytes = (uint32_t)((2 * (c_hi - s0) + soff + 15) & ~15ll);
    }
    long long b_lo = 0;
    if (s1 > 0) {
        b_lo = (2 * x_lo) & ~15ll;
        x_bytes = (uint32_t)(((2 * s1 + 15) & ~15ll) - b_lo);
    }
    const uint32_t total = carry_bytes + x_bytes;
    const unsigned char *csrc = a.carry_end + 2 * s0 - soff;
    const unsigned char *xsrc = a.x + b_lo - carry_bytes;
    for (uint32_t f = (uint32_t)tid * 16u; f < total; f += (uint32_t)NT * 16u) {
        const uint32_t row = f / ROW;
        cp_async_16(tile + f + row * PAD, (f < carry_bytes ? csrc : xsrc) + f);
    }
    cp_async_mbar_arrive_noinc(bar);
    return soff;
}

// =================================================================================================
// Specialised kernel
// =================================================================================================
// halo blocks: >= Q so that y[m-1] of the first owned output is complete, and such that the tile stride (OUT*D
// samples) is a whole number of load units, which keeps the load phase CTA-uniform
__host__ __device__ constexpr int fast_pick_hb(int Q, int NBLK, int D, int SPL) {
    int hb = Q;
    while (((NBLK - hb) * D) % SPL != 0) hb++;
    return hb;
}
// partial sums are stored [block][lag] with this row length: an even lag count from 4 up gets one unused slot, so that
// the combine (lane stride = one row) and the per-thread stores stop colliding on the same banks
__host__ __device__ constexpr int fast_qp(int Q) { return (Q >= 4 && Q % 2 == 0) ? Q + 1 : Q; }

template <int T, int D, int B, int NT, int WB, int PAD = 0>
struct FastGeom {
    static constexpr int Q = (T + D - 1) / D;           // lags: outputs a sample contributes to
    static constexpr int NBLK = NT * B;                 // decimation blocks per CTA tile
    static constexpr int SPL = WB / 2;                  // samples per shared-memory load (LDS.32 / LDS.64)
    static constexpr int HB = fast_pick_hb(Q, NBLK, D, SPL);
    static constexpr int OUT = NBLK - HB;               // outputs owned per CTA
    static constexpr int TILE_BYTES = NBLK * D * 2;
    static constexpr int ROW = B * D * 2;                // bytes one thread owns
    static constexpr int SM_TILE = ((TILE_BYTES + 15) / 16) * 16 + 32 + (NT + 1) * PAD;
    static constexpr int QP = fast_qp(Q);               // row length of the partial-sum array
    static constexpr int SM_PART = NBLK * QP * 8;       // float2 partial per (block, lag)
    static constexpr bool ROWS_ASYNC = PAD != 0;        // padded tiles are staged by cp.async from every thread
    static constexpr int SM_Y = NBLK * 8;
    static constexpr int SMEM = SM_TILE + SM_PART + SM_Y;
    static_assert(WB == 4 || WB == 8, "LDS.32 or LDS.64");
    static_assert((B * D) % SPL == 0, "thread span must be a whole number of load units");
    static_assert(OUT > HB, "tile too small");
    static_assert(PAD == 0 || (ROW % 16 == 0 && PAD % 16 == 0), "padded rows are staged by 16-byte bulk copies");
};

// One PRMT builds the half2 (1024+I, 1024+Q) (fp16 0x64bb == 1024+bb exactly); the Blackwell
// mixed-precision add (PTX add.rn.f32.f16, SASS FHADD) then yields the centred f32 sample in one
// instruction per component: 3 instructions per complex sample, all exact.
struct CvtConst {
    float bias;        // -(1024 + 127)
    uint32_t h1024;    // 0x64646464: fp16 exponent byte of 1024 for PRMT
};
__device__ __forceinline__ CvtConst cvt_consts() {
    // Both constants are made opaque AND per-thread (tid >> 31 == 0) so that they live in ordinary vector
    // registers: as literals / uniform values the compiler re-materialises them with one MOV per use
    // (FHADD takes no immediate or uniform operand), which costs more than the conversion itself.
    CvtConst c;
    asm volatile(
        "{\n"
        ".reg .u32 t;\n"
        "mov.u32 t, %%tid.x;\n"
        "shr.u32 t, t, 31;\n"
        "or.b32 %0, t, 0xC48FE000;\n"
        "or.b32 %1, t, 0x64646464;\n"
        "}\n"
        : "=f"(c.bias), "=r"(c.h1024));
    return c;
}
// Packed FP32: Blackwell's FFMA2 (PTX fma.rn.f32x2) does the re and im lanes of one tap in ONE issue slot;
// ptxas turns the duplicated tap {h, h} into a scalar-broadcast uniform operand (FFMA2 R, R.F32x2, UR.F32, R)
// fed by one LDCU.128 per four taps.  Each half is an ordinary IEEE fma, so results are bit-identical to fmaf.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void fma_f32x2(unsigned long long &acc, float h, unsigned long long x) {
    const unsigned long long hh = pack_f32x2(h, h);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(hh), "l"(x));
}
__device__ __forceinline__ float2 unpack_f32x2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

__device__ __forceinline__ void cvt_iq(uint32_t w, int half, const CvtConst &c, float &xr, float &xi) {
    const uint32_t pair = __byte_perm(w, c.h1024, half ? 0x4342u : 0x4140u);
    asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(xr) : "h"((unsigned short)(pair & 0xffffu)), "f"(c.bias));
    asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(xi) : "h"((unsigned short)(pair >> 16)), "f"(c.bias));
}

// One CTA tile of G::OUT outputs.  `blk` = tile index inside the call; `first_use` / `parity`: a persistent CTA (the ring
// kernel below) re-arms the same mbarrier once per tile and flips its phase; PERSIST adds the trailing barrier that makes
// the shared-memory tile reusable.
template <int T, int D, int B, int NT, int WB, int PH, int PAD, bool PERSIST>
__device__ __forceinline__ void fir_fast_tile(const FirArgs &a, const Taps<T> &taps, const long long blk, const bool last_blk,
                                              const uint32_t parity, const bool first_use, uint64_t &bar, uint32_t &sh_soff) {
    using G = FastGeom<T, D, B, NT, WB, PAD>;
    constexpr int Q = G::Q, SPL = G::SPL;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char *tile = smem;
    float2 *part = reinterpret_cast<float2 *>(smem + G::SM_TILE);
    float2 *ysm = reinterpret_cast<float2 *>(smem + G::SM_TILE + G::SM_PART);

    const int tid = threadIdx.x;
    const long long out0 = blk * G::OUT;          // first owned output (call-local)
    // tile = blocks [out0 - HB, out0 - HB + NBLK); block b covers samples [b*D - r, (b+1)*D - r)
    if constexpr (G::ROWS_ASYNC) {
        if (first_use) {   // CTA-uniform
            if (tid == 0) {
                mbar_init(&bar, NT);
                fence_barrier_init();
            }
            __syncthreads();
        }
        const long long s0 = (out0 - G::HB) * D - (long long)a.r;
        const long long last_out = out0 + G::OUT < a.n_out ? out0 + G::OUT : a.n_out;   // exclusive
        const long long s1 = last_out * D - (long long)a.r;                              // end of the last needed block
        const uint32_t so = load_tile_rows_async<G::ROW, PAD, NT>(tile, a, s0, s1, &bar, tid);
        if (tid == 0) sh_soff = so;
    } else if (tid == 0) {
        if (first_use) {
            mbar_init(&bar, 1);
            fence_barrier_init();
        }
        long long s0 = (out0 - G::HB) * D - (long long)a.r;
        long long last_out = out0 + G::OUT < a.n_out ? out0 + G::OUT : a.n_out;   // exclusive
        long long s1 = last_out * D - (long long)a.r;                              // end of the last needed block
        sh_soff = load_tile(tile, a, s0, s1, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, parity);

    // ---- convert once, accumulate per (block, lag) ----------------------------------------------
    // The thread's first sample sits PH samples into load unit (soff / WB) + tid * (B*D/SPL); with PAD every thread's
    // row of U load units is followed by PAD unused bytes.
    constexpr int U = B * D / SPL;
    const uint32_t ubias = sh_soff / WB;   // < 16 / WB
    const unsigned char *ubase = tile + (size_t)tid * (G::ROW + PAD) + (size_t)ubias * WB;
    unsigned long long acc[B][Q];   // packed (re, im) accumulators
#pragma unroll
    for (int b = 0; b < B; b++)
#pragma unroll
        for (int q = 0; q < Q; q++) acc[b][q] = 0ull;

    constexpr int NU = (PH + B * D + SPL - 1) / SPL;
    const CvtConst bias = cvt_consts();
#pragma unroll
    for (int u = 0; u < NU; u++) {
        uint32_t words[WB / 4];
        // unit u of this thread is unit ubias + u of its row; only the last few can spill into the next row
        const unsigned char *up = ubase + u * WB;
        if (PAD && u + 16 / WB > U) up += (ubias + u >= (uint32_t)U) ? PAD : 0;
        if (WB == 8) {
            const uint2 v = *reinterpret_cast<const uint2 *>(up);
            words[0] = v.x;
            words[WB / 4 - 1] = v.y;
        } else {
            words[0] = *reinterpret_cast<const uint32_t *>(up);
        }
#pragma unroll
        for (int i = 0; i < SPL; i++) {
            const int j = u * SPL + i - PH;   // sample index inside the thread's span
            if (j < 0 || j >= B * D) continue;
            float xr, xi;
            cvt_iq(words[i >> 1], i & 1, bias, xr, xi);
            const unsigned long long x2 = pack_f32x2(xr, xi);
            const int bb = j / D, jj = j % D;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                const int k = q * D + (D - 1 - jj);
                if (k < T) fma_f32x2(acc[bb][q], taps.h[k], x2);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < B; b++)
#pragma unroll
        for (int q = 0; q < Q; q++) part[(tid * B + b) * G::QP + q] = unpack_f32x2(acc[b][q]);
    __syncthreads();

    // ---- combine partials oldest block first: y[g] = P[g-Q+1][Q-1] + ... + P[g][0] -----------------
#pragma unroll
    for (int u = 0; u < B; u++) {
        const int g = tid + u * NT;
        float yr = 0.f, yi = 0.f;
        if (g >= Q - 1) {
#pragma unroll
            for (int q = Q - 1; q >= 0; q--) {
                float2 p = part[(g - q) * G::QP + q];
                yr += p.x;
                yi += p.y;
            }
        }
        ysm[g] = make_float2(yr, yi);
    }
    __syncthreads();

    // ---- discriminator + stores ----------------------------------------------------------------------
    // 32-bit indices relative to the tile: outputs [0, n_here) of this tile exist; the resampler history of the next call
    // starts hoff outputs into the tile (beyond it for all but the last tiles)
    const long long left = a.n_out - out0;
    const int n_here = left < G::OUT ? (int)left : G::OUT;
    const long long hfrom = left - a.h2;                       // tile-relative index of the first history value
    const int hoff = (a.hist_out == nullptr || hfrom >= G::OUT) ? G::OUT : (hfrom < 0 ? 0 : (int)hfrom);
    float2 *yp = a.y_out ? a.y_out + out0 : nullptr;
    float *dp = a.d_out ? a.d_out + out0 : nullptr;
#pragma unroll
    for (int u = 0; u < B; u++) {
        const int g = tid + u * NT;
        const int o = g - G::HB;
        if (o < 0 || o >= n_here) continue;
        const float2 y = ysm[g];
        if (yp) yp[o] = y;
        if (dp) {
            const float dv = discriminate(y, ysm[g - 1], a.gain);
            dp[o] = dv;
            if (o >= hoff) a.hist_out[o - (int)hfrom] = dv;
        }
        if (last_blk && o == n_here - 1) *a.last_y = y;
    }
    if (a.carry_out && last_blk) fold_carry_update(a, tid, NT);
    if (PERSIST) __syncthreads();   // the tile, the partial sums and ysm are rewritten by the next tile
}

template <int T, int D, int B, int NT, int WB, int PH, int PAD = 0>
__global__ void __launch_bounds__(NT) k_fir_fast(const FirArgs a, const __grid_constant__ Taps<T> taps) {
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t sh_soff;
    fir_fast_tile<T, D, B, NT, WB, PH, PAD, false>(a, taps, (long long)blockIdx.x, blockIdx.x == gridDim.x - 1, 0u, true, bar, sh_soff);
}



// =================================================================================================
// "Output-owner" kernel for the shapes the block-owner form serves badly or not at all: many lags per sample (more than 16:
// T > 16 D) or a very small decimation, where one thread's few samples cannot amortise the per-thread partial-sum traffic.
// Every sample of the CTA tile is converted ONCE into a packed (re, im) f32 tile in shared memory; a thread then owns R
// consecutive outputs and slides over its (R-1) D + T samples: one LDS.64 per sample feeds up to R FFMA2 (tap = uniform
// register from the __grid_constant__ parameter, compile-time index).  Per-output sum order: oldest sample first, fixed,
// so results do not depend on how the stream is tiled or chunked.  The float2 tile skips one slot after every R D samples
// when R D is even: the lane stride in 8-byte units is odd, LDS.64 is conflict-free.  Local output 0 of a CTA is the
// predecessor of its first owned output (discriminator), recomputed.
// =================================================================================================
template <int T, int D, int R, int NT>
struct SlideGeom {
    static constexpr int NOUT = NT * R;               // outputs computed per CTA
    static constexpr int OPC = NOUT - 1;              // outputs owned per CTA
    static constexpr int NS = T + (NOUT - 1) * D;     // samples in the tile
    static constexpr int RD = R * D;
    static constexpr int PADF = (RD % 2 == 0) ? 1 : 0;
    static constexpr int W = (R - 1) * D + T;         // samples one thread reads
    static constexpr int SM_RAW = ((NS * 2 + 15) / 16) * 16 + 32;
    static constexpr int NF = NS + (NS / RD + 1) * PADF;
    static constexpr int SM_F = NF * 8;
    static constexpr int SMEM = SM_RAW + SM_F + NOUT * 8;
    static_assert(NT % 32 == 0 && R >= 1, "whole warps");
};
__host__ __device__ constexpr int slide_smem(int T, int D, int R, int NT) {
    const int nout = NT * R, ns = T + (nout - 1) * D, rd = R * D, padf = (rd % 2 == 0) ? 1 : 0;
    return ((ns * 2 + 15) / 16) * 16 + 32 + (ns + (ns / rd + 1) * padf) * 8 + nout * 8;
}

template <int T, int D, int R, int NT>
__global__ void __launch_bounds__(NT) k_fir_slide(const FirArgs a, const __grid_constant__ Taps<T> taps) {
    using G = SlideGeom<T, D, R, NT>;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t sh_soff;
    unsigned char *raw = smem;
    unsigned long long *xf = reinterpret_cast<unsigned long long *>(smem + G::SM_RAW);
    float2 *ysm = reinterpret_cast<float2 *>(smem + G::SM_RAW + G::SM_F);
    const int tid = threadIdx.x;
    const long long out0 = (long long)blockIdx.x * G::OPC;
    const long long left = a.n_out - out0;
    const int n_here = left < G::OPC ? (int)left : G::OPC;   // owned outputs that exist; local output l is out0 - 1 + l
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        const long long s0 = out0 * D - (long long)a.r - T;   // oldest sample of local output 0
        const long long s1 = (out0 + n_here) * D - (long long)a.r;
        sh_soff = load_tile(raw, a, s0, s1, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    // ---- convert every sample once ------------------------------------------------------------------
    const uint16_t *t16 = reinterpret_cast<const uint16_t *>(raw + sh_soff);
    const CvtConst cb = cvt_consts();
    const int ns = T + n_here * D;   // samples that exist in this tile
    for (int i = tid; i < G::NS; i += NT) {
        float xr = 0.f, xi = 0.f;
        if (i < ns) cvt_iq(t16[i], 0, cb, xr, xi);
        xf[i + (i / G::RD) * G::PADF] = pack_f32x2(xr, xi);
    }
    __syncthreads();

    // ---- R outputs per thread, sliding over the window ------------------------------------------------
    unsigned long long acc[R];
#pragma unroll
    for (int rr = 0; rr < R; rr++) acc[rr] = 0ull;
    const unsigned long long *xw = xf + tid * (G::RD + G::PADF);
#pragma unroll
    for (int j = 0; j < G::W; j++) {
        const unsigned long long x2 = xw[j + (j / G::RD) * G::PADF];
#pragma unroll
        for (int rr = 0; rr < R; rr++) {
            const int k = T - 1 - (j - rr * D);   // output rr of this thread starts rr * D samples later
            if (k >= 0 && k < T) fma_f32x2(acc[rr], taps.h[k], x2);
        }
    }
#pragma unroll
    for (int rr = 0; rr < R; rr++) ysm[tid * R + rr] = unpack_f32x2(acc[rr]);
    __syncthreads();

    // ---- discriminator + stores (32-bit indices relative to the tile, as in fir_fast_tile) ---------------------------
    const long long hfrom = left - a.h2;                       // tile-relative index of the first history value
    const int hoff = (a.hist_out == nullptr || hfrom >= G::OPC) ? G::OPC : (hfrom < 0 ? 0 : (int)hfrom);
    float2 *yp = a.y_out ? a.y_out + out0 : nullptr;
    float *dp = a.d_out ? a.d_out + out0 : nullptr;
    const bool last_cta = blockIdx.x == gridDim.x - 1;
    for (int o = tid; o < n_here; o += NT) {
        const float2 y = ysm[o + 1];
        if (yp) yp[o] = y;
        if (dp) {
            const float dv = discriminate(y, ysm[o], a.gain);
            dp[o] = dv;
            if (o >= hoff) a.hist_out[o - (int)hfrom] = dv;
        }
        if (last_cta && o == n_here - 1) *a.last_y = y;
    }
    if (a.carry_out && blockIdx.x == gridDim.x - 1) fold_carry_update(a, tid, NT);
}

// =================================================================================================
// Persistent ring for the f32 receiver: successive USB-sized buffers stream through ONE resident kernel
// (reader -> channel -> processor of examples/simple_fm.rs:55-60,108-128,145-160; same protocol as k_demod_ring).
//
// The host copies buffer k into device slot k % m (m = n_slots + 1) and then, stream-ordered, writes the doorbell
// seq_ready[k % n_slots] = k + 1.  Every CTA walks the buffers in order and takes FIR tiles blk = blockIdx.x, +gridDim.x ...
// of that buffer — the same fir_fast_tile code as the one-shot kernel, so every output bit is the same.  No state hops
// between buffers: the carried history of buffer k IS the tail of device slot k-1 (still intact: a device slot is
// rewritten only after the buffer after it has been collected), and y[m-1] of a buffer's first output is recomputed from
// it by the halo blocks.  The discriminator values go to a device-resident [history | new] buffer per slot (the FIR tiles of
// buffer k also write the head of buffer k+1's); when the last FIR tile of a buffer has finished (a counter, then a
// release-store of fir_gen) the audio tiles of that buffer run — the polyphase FIR a[i] = sum_j gp[phase][j] d[p_i - j]
// with the same ascending-j fma chain as the one-shot audio kernels — straight into host-mapped memory; the CTA that
// finishes the last one publishes seq_done[slot] = k + 1 there.  Host cost per buffer: two cudaMemcpyAsync enqueues.
// =================================================================================================
struct FxRingCtl {                 // device memory
    unsigned int seq_ready[64];    // doorbells, written by the copy engine
    unsigned int fir_cnt[72];      // FIR tiles finished, per device slot
    unsigned int fir_gen[72];      // k + 1 once every FIR tile of buffer k has finished
    unsigned int aud_cnt[72];      // audio tiles finished, per device slot
    unsigned int stop;             // host sets 1 to retire the kernel
};

struct FxRingArgs {
    FxRingCtl *ctl;
    volatile unsigned int *seq_done;   // host-mapped [n_slots]
    const uint8_t *d_in;               // device slots [m][slot_stride], each 16-B aligned
    const uint8_t *carry0_end;         // one past the handle's carried samples (history of buffer 0)
    float *dbuf;                       // [m][dbuf_stride]: h2 history values, then the buffer's discriminator values
    float *h_out;                      // host-mapped audio slots [n_slots][out_stride]
    const float *gp;                   // polyphase taps [L][J]
    float2 *last_y;
    unsigned long long S, n_in0, n_y0, slot_stride, dbuf_stride, out_stride;
    unsigned int n_slots, m, D, L, M, J, h2, has_res;
    float gain;
};

__device__ __forceinline__ unsigned int ring_ld_acquire(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ring_st_release(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int kRingAudioTile = 256;   // audio outputs per audio tile

template <int T, int D, int B, int NT, int WB, int PAD = 0>
__global__ void __launch_bounds__(NT) k_fmrx_ring(const FxRingArgs r, const __grid_constant__ Taps<T> taps) {
    using G = FastGeom<T, D, B, NT, WB, PAD>;
    __shared__ __align__(8) uint64_t bar;   // ONE mbarrier for every load-phase instantiation below: `uses` counts its phases
    __shared__ uint32_t sh_soff;
    __shared__ FirArgs sh_a;
    __shared__ unsigned long long sh_a0, sh_p0;
    __shared__ unsigned int sh_go, sh_ntiles, sh_natiles, sh_na, sh_phase, sh_ph0;
    const int tid = threadIdx.x;
    uint32_t uses = 0;
    for (unsigned long long k = 0;; k++) {
        const unsigned int hs = (unsigned int)(k % r.n_slots), ds = (unsigned int)(k % r.m);
        if (tid == 0) {
            unsigned int go = 2;
            while (go == 2) {
                if (ring_ld_acquire(&r.ctl->seq_ready[hs]) == (unsigned int)(k + 1)) go = 1;
                else if (ring_ld_acquire(&r.ctl->stop))   // a doorbell rung just before close was written before stop
                    go = ring_ld_acquire(&r.ctl->seq_ready[hs]) == (unsigned int)(k + 1) ? 1 : 0;
                else __nanosleep(100);
            }
            sh_go = go;
            if (go) {
                // closed-form stream position of buffer k
                const unsigned long long n0 = r.n_in0 + k * r.S;
                const unsigned long long y0 = n0 / D, y1 = (n0 + r.S) / D;   // global FIR output range [y0, y1)
                FirArgs a;
                a.x = r.d_in + (size_t)ds * r.slot_stride;
                a.carry_end = k ? r.d_in + (size_t)((k - 1) % r.m) * r.slot_stride + 2 * r.S : r.carry0_end;
                a.n_samples = (long long)r.S;
                a.n_out = (long long)(y1 - y0);
                a.r = (uint32_t)(n0 - y0 * D);
                a.gain = r.gain;
                a.y_out = nullptr;
                a.last_y = r.last_y;
                a.carry_out = nullptr;
                a.cs = 0;
                a.h2 = (int)r.h2;
                unsigned long long a0 = y0, a1 = y1;   // audio range; without a resample stage the audio IS d
                if (r.has_res) {
                    a0 = (y0 * r.L + r.M - 1) / r.M;
                    a1 = (y1 * r.L + r.M - 1) / r.M;
                    a.d_out = r.dbuf + (size_t)ds * r.dbuf_stride + r.h2;
                    a.hist_out = r.dbuf + (size_t)((k + 1) % r.m) * r.dbuf_stride;
                } else {
                    a.d_out = r.h_out + (size_t)hs * r.out_stride;
                    a.hist_out = nullptr;
                }
                sh_a = a;
                sh_a0 = a0;
                sh_na = (unsigned int)(a1 - a0);
                sh_ntiles = (unsigned int)((a.n_out + G::OUT - 1) / G::OUT);
                sh_natiles = r.has_res ? (sh_na + kRingAudioTile - 1) / kRingAudioTile : 0u;
                // load phase of this buffer's tiles (fx_path.cu launch_fir): sample s0 = -(HB*D + r) inside its 16-byte line
                const long long s0 = -((long long)G::HB * D + (long long)a.r);
                sh_phase = (((uint32_t)((2 * s0) & 15)) % (uint32_t)WB) / 2u;
            }
        }
        __syncthreads();
        if (!sh_go) return;
        const unsigned int n_tiles = sh_ntiles, n_atiles = sh_natiles;
        // ---- FIR + discriminator tiles -------------------------------------------------------------------------------
        for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const bool last = t == n_tiles - 1;
            switch (sh_phase) {
                case 0: fir_fast_tile<T, D, B, NT, WB, 0, PAD, true>(sh_a, taps, t, last, uses & 1, uses == 0, bar, sh_soff); break;
                case 1: fir_fast_tile<T, D, B, NT, WB, 1, PAD, true>(sh_a, taps, t, last, uses & 1, uses == 0, bar, sh_soff); break;
                case 2: fir_fast_tile<T, D, B, NT, WB, (WB == 8 ? 2 : 0), PAD, true>(sh_a, taps, t, last, uses & 1, uses == 0, bar, sh_soff); break;
                default: fir_fast_tile<T, D, B, NT, WB, (WB == 8 ? 3 : 1), PAD, true>(sh_a, taps, t, last, uses & 1, uses == 0, bar, sh_soff); break;
            }
            uses++;
            if (tid == 0) {   // (fir_fast_tile ended with a __syncthreads: every store of the tile has been issued)
                if (!r.has_res) __threadfence_system();   // d went straight to host memory
                __threadfence();
                if (atomicAdd(&r.ctl->fir_cnt[ds], 1u) == n_tiles - 1) {
                    r.ctl->fir_cnt[ds] = 0;
                    ring_st_release(&r.ctl->fir_gen[ds], (unsigned int)(k + 1));
                    if (!r.has_res) {
                        __threadfence_system();
                        r.seq_done[hs] = (unsigned int)(k + 1);
                    }
                }
            }
        }
        // ---- audio tiles: wait for this buffer's (and, for the history head, the previous buffer's) FIR tiles ------------
        for (unsigned int t = blockIdx.x; t < n_atiles; t += gridDim.x) {
            if (tid == 0) {
                while (ring_ld_acquire(&r.ctl->fir_gen[ds]) != (unsigned int)(k + 1)) __nanosleep(50);
                if (k)
                    while (ring_ld_acquire(&r.ctl->fir_gen[(k - 1) % r.m]) != (unsigned int)k) __nanosleep(50);
                const unsigned long long t0 = (sh_a0 + (unsigned long long)t * kRingAudioTile) * r.M;
                const unsigned long long p0 = t0 / r.L;
                sh_p0 = p0;
                sh_ph0 = (unsigned int)(t0 - p0 * r.L);
            }
            __syncthreads();
            {
                const unsigned long long n0 = r.n_in0 + k * r.S;
                const unsigned long long y0 = n0 / D;
                // dbuf index of d[p0]: history h2, then the buffer's values from global output y0
                const float *dslot = r.dbuf + (size_t)ds * r.dbuf_stride;
                const long long xbase = (long long)(sh_p0 - y0) + (long long)r.h2;
                float *outp = r.h_out + (size_t)hs * r.out_stride + (size_t)t * kRingAudioTile;
                const unsigned int n_here = sh_na - t * kRingAudioTile < (unsigned int)kRingAudioTile ? sh_na - t * kRingAudioTile : (unsigned int)kRingAudioTile;
                for (unsigned int o = tid; o < n_here; o += NT) {
                    const unsigned int trel = sh_ph0 + o * r.M;
                    const unsigned int dp = trel / r.L, ph = trel - dp * r.L;
                    const float *g = r.gp + (size_t)ph * r.J;
                    const float *x = dslot + (xbase + dp);
                    float acc = 0.f;
                    for (unsigned int j = 0; j < r.J; j++) acc = fmaf(__ldg(g + j), __ldcg(x - j), acc);
                    outp[o] = acc;
                }
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence_system();   // audio stores reach host memory
                if (atomicAdd(&r.ctl->aud_cnt[ds], 1u) == n_atiles - 1) {
                    r.ctl->aud_cnt[ds] = 0;
                    __threadfence_system();
                    r.seq_done[hs] = (unsigned int)(k + 1);
                }
            }
        }
        __syncthreads();   // sh_a is rewritten for the next buffer
    }
}

}  // namespace sdr
