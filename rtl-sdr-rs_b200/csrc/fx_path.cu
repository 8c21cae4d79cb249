// fx_path.cu — f32 tap'd-FIR receiver (BASELINE.json configs 2-3) for sm_100a.
//
//   low_pass : y[m] = sum_{k<T} h[k] * (x[(m+1)D-1-k] - 127)          complex f32, x[n<0] = 127
//   fm_demod : d[m] = gain * atan2(Im(y[m] conj y[m-1]), Re(..))       f32
//   resample : a[i] = sum_p g[iM - pL] * d[p]                           rational L/M polyphase FIR
//
// Hot kernel: k_fir_fast<T,D,B,NT,WB,PH> — fused u8->f32 convert + decimating FIR + discriminator.
// "Block-owner" polyphase form: the stream is cut into decimation blocks of D samples; a thread owns
// B consecutive blocks, converts every byte exactly once (one PRMT to a half2 (1024+I, 1024+Q), two
// mixed-precision adds FHADD -> centred f32) and feeds the sample to the ceil(T/D) outputs it contributes
// to with packed-FP32 FFMA2 (re and im lanes in one issue slot; the tap is a scalar-broadcast uniform
// register loaded four at a time from the kernel-parameter constant bank — no per-thread tap loads).
// Per-(block,lag) partial sums are combined in a fixed order, so results are independent of how the
// stream is tiled or chunked.  Each CTA's raw bytes arrive with one 1-D bulk async copy (TMA engine) and
// are read once from HBM: 2 B in + 4/D B out per complex sample.  y never leaves the SM unless asked for.
//
// Fallback for arbitrary (T,D): k_fir_generic — one warp per output, lanes stride the taps,
// warp-shuffle reduction.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "fir_fast.cuh"
#include "rtc.h"

namespace sdr {

// =================================================================================================
// Generic kernel: one warp per output, lanes stride the taps, butterfly reduction.
// =================================================================================================
struct GenArgs {
    FirArgs f;
    const float *taps;   // [T] in global memory
    int T, D, OPC;       // OPC = outputs owned per CTA
    uint32_t sm_tile;    // bytes reserved for the tile
};

__global__ void __launch_bounds__(128) k_fir_generic(const GenArgs g) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t sh_soff;
    unsigned char *tile = smem;
    float *htap = reinterpret_cast<float *>(smem + g.sm_tile);
    float2 *ysm = reinterpret_cast<float2 *>(smem + g.sm_tile + (size_t)((g.T + 3) & ~3) * 4);
    const FirArgs &a = g.f;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long out0 = (long long)blockIdx.x * g.OPC;
    long long n_here = a.n_out - out0 < g.OPC ? a.n_out - out0 : g.OPC;
    // local output l in [0, n_here] is call-local output out0 - 1 + l (l = 0 is the predecessor for the discriminator)
    const long long s0 = out0 * g.D - (long long)a.r - g.T;   // first sample of output out0-1
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        long long s1 = (out0 + n_here) * g.D - (long long)a.r;
        sh_soff = load_tile(tile, a, s0, s1, &bar);
    }
    for (int k = tid; k < g.T; k += blockDim.x) htap[k] = g.taps[k];
    __syncthreads();
    mbar_wait(&bar, 0);
    const uint16_t *t16 = reinterpret_cast<const uint16_t *>(tile + sh_soff);
    const CvtConst gbias = cvt_consts();
    for (long long l = warp; l <= n_here; l += 4) {
        // newest sample of this output, relative to s0
        const int newest = (int)((out0 - 1 + l + 1) * g.D - 1 - (long long)a.r - s0);
        float ar = 0.f, ai = 0.f;
        for (int k = lane; k < g.T; k += 32) {
            float xr, xi;
            cvt_iq(t16[newest - k], 0, gbias, xr, xi);
            const float h = htap[k];
            ar = fmaf(h, xr, ar);
            ai = fmaf(h, xi, ai);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, o);
            ai += __shfl_xor_sync(0xffffffffu, ai, o);
        }
        if (lane == 0) ysm[l] = make_float2(ar, ai);
    }
    __syncthreads();
    for (long long l = 1 + tid; l <= n_here; l += blockDim.x) {
        const long long i = out0 + l - 1;
        const float2 y = ysm[l];
        if (a.y_out) a.y_out[i] = y;
        if (a.d_out) {
            const float dv = discriminate(y, ysm[l - 1], a.gain);
            a.d_out[i] = dv;
            if (a.hist_out && i >= a.n_out - a.h2) a.hist_out[i - (a.n_out - a.h2)] = dv;
        }
        if (i == a.n_out - 1) *a.last_y = y;
    }
    if (a.carry_out && blockIdx.x == gridDim.x - 1) fold_carry_update(a, tid, blockDim.x);
}

// =================================================================================================
// Small kernels: carry update, stage-level discriminator, resampler
// =================================================================================================

// new_carry = last cs samples of (old_carry ++ x[0..n)); both carries hold exactly cs samples (u16 each).
__global__ void k_update_carry(const uint16_t *old_carry, const uint16_t *x, long long n, int cs, uint16_t *new_carry) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cs; i += gridDim.x * blockDim.x) {
        long long p = n - cs + i;
        new_carry[i] = p >= 0 ? x[p] : old_carry[cs + p];
    }
}

// Discriminator history: cur = [hist (h2) | new d (n)]; the last h2 values become the head of the NEXT buffer (the
// buffers rotate, so nothing is moved in place).  Only calls with fewer than h2 new values need this kernel: otherwise
// the FIR kernel writes the next head itself (FirArgs::hist_out).
__global__ void k_hist_move(const float *cur, long long n, int h2, float *next) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h2; i += gridDim.x * blockDim.x) next[i] = cur[n + i];
}

__global__ void k_fm_demod_f32(const float2 *y, long long n, float2 *prev_state, float gain, float *out) {
    long long stride = (long long)gridDim.x * blockDim.x;
    const float2 prev0 = *prev_state;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = discriminate(y[i], i ? y[i - 1] : prev0, gain);
}
__global__ void k_store_prev(const float2 *y, long long n, float2 *prev_state) {
    if (n > 0) *prev_state = y[n - 1];
}

// Resampler a[i] = sum_p g[iM - pL] d[p] in polyphase form: phase = (iM) mod L, p_i = (iM) div L,
//   a[i] = sum_{j<J} gp[phase][j] * d[p_i - j],   gp[phase][j] = g[phase + jL] (zero padded), J = ceil(T2/L).
// dbuf[h2 + (p - P0)] = d[p] (history of h2 >= J values in front); outputs i in [i0, i0 + n_out).

// L = M = 1 (plain real FIR at the output rate).  Register tile of kFirR = 16 outputs x 16 taps per step: a
// 32-float sliding window (8 LDS.128) and 16 taps (4 broadcast LDS.128) feed 256 FMAs, which keeps shared-
// memory wavefronts (36 per 256 warp-FMAs) under the FMA issue time — an 8x8 tile is shared-memory-bound.
constexpr int kFirR = 16, kFirTC = 16, kFirThreads = 128, kFirOblk = kFirR * kFirThreads;
inline size_t fir_real_smem(int Jp) {   // taps + padded window (one pad chunk per 8 chunks)
    return ((size_t)Jp + ((size_t)(kFirOblk + Jp) * 9) / 8 + 16) * sizeof(float);
}
__global__ void __launch_bounds__(kFirThreads) k_fir_real_r8(const float *dbuf, int h2, const float *gp, int Jp,
                                                             long long n_out, long long n_valid, float *out) {
    extern __shared__ __align__(16) float fsm[];
    float *taps = fsm;            // [Jp], Jp a multiple of 16
    // window: logical win[idx] = dbuf[h2 + o0 - Jp + idx], idx < kFirOblk + Jp, stored in 16-byte chunks
    // with one pad chunk after every 8 (chunk c lives at c + c/8): lanes read chunks 4t+k, and the pad
    // turns that stride-4 pattern into 8 distinct bank groups per quarter warp (conflict-free LDS.128).
    float4 *win4 = reinterpret_cast<float4 *>(fsm + Jp);
    float *win = fsm + Jp;
    const long long o0 = (long long)blockIdx.x * kFirOblk;
    for (int k = threadIdx.x; k < Jp; k += blockDim.x) taps[k] = gp[k];
    const long long gbase = (long long)h2 + o0 - Jp;   // h2 = Jp + 16, o0 % 2048 == 0: a multiple of 16 floats
    const int n_chunks = (kFirOblk + Jp) / 4;
    if ((gbase & 3) == 0 && gbase >= 0 && gbase + kFirOblk + Jp <= n_valid) {
        const float4 *src = reinterpret_cast<const float4 *>(dbuf + gbase);   // whole tile valid: 16-byte loads
        for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) win4[c + (c >> 3)] = src[c];
    } else {
        for (int idx = threadIdx.x; idx < kFirOblk + Jp; idx += blockDim.x) {
            long long gi = gbase + idx;
            win[idx + (idx >> 5) * 4] = (gi >= 0 && gi < n_valid) ? dbuf[gi] : 0.f;
        }
    }
    __syncthreads();
    float acc[kFirR];
#pragma unroll
    for (int r = 0; r < kFirR; r++) acc[r] = 0.f;
    const int t0 = threadIdx.x * kFirR;
    for (int j0 = 0; j0 < Jp; j0 += kFirTC) {
        // outputs u = t0 + r (r < 16), taps j = j0 + jj (jj < 16): sample win[Jp + u - j]; the 32 floats
        // w[c] = win[Jp + t0 - j0 - 16 + c] cover it: sample(r, jj) = w[16 + r - jj]
        float w[kFirR + kFirTC], g[kFirTC];
        const int c0 = (Jp + t0 - j0 - kFirTC) >> 2;   // first logical chunk (all terms are multiples of 16)
        const float4 *gq = reinterpret_cast<const float4 *>(taps + j0);
#pragma unroll
        for (int c = 0; c < (kFirR + kFirTC) / 4; c++) {
            float4 v = win4[(c0 + c) + ((c0 + c) >> 3)];
            w[4 * c] = v.x, w[4 * c + 1] = v.y, w[4 * c + 2] = v.z, w[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < kFirTC / 4; c++) {
            float4 v = gq[c];
            g[4 * c] = v.x, g[4 * c + 1] = v.y, g[4 * c + 2] = v.z, g[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int jj = 0; jj < kFirTC; jj++)
#pragma unroll
            for (int r = 0; r < kFirR; r++) acc[r] = fmaf(g[jj], w[kFirTC + r - jj], acc[r]);
    }
    // 16 consecutive outputs per thread: four 16-byte stores when the row is whole and aligned
    float *dst = out + o0 + t0;
    if (o0 + t0 + kFirR <= n_out && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int c = 0; c < kFirR / 4; c++)
            reinterpret_cast<float4 *>(dst)[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
    } else {
#pragma unroll
        for (int r = 0; r < kFirR; r++)
            if (o0 + t0 + r < n_out) dst[r] = acc[r];
    }
}

// L = M = 1, packed form.  Two consecutive outputs share every sample with taps one apart:
//   out[n] += g[j] * x[n-j],   out[n+1] += g[j+1] * x[n-j]
// so ONE packed FFMA2 does both with the sample as the scalar-broadcast operand and the tap PAIR (g[j], g[j+1]) as the
// vector operand — no register-pair alignment problem (the reason a window-pair formulation needs the window twice).
// The overlapping tap pairs P[k] = (G[k], G[k+1]), G[0] = 0, G[k] = g[k-1], are a kernel parameter (uniform loads).  Per
// 16 outputs x 16 taps: 128 FFMA2 + 8 LDS.128 + 16 uniform loads instead of 256 FFMA + 12 LDS.  Each output is still the
// ascending-tap fma chain of k_fir_real_r8 / k_resample_poly / the ring's audio tiles (the extra first and last products
// have a zero tap), so all of them agree bit for bit.
constexpr int kFirPMaxJ = 2048;
struct TapsP {
    float2 p[kFirPMaxJ];
};
__global__ void __launch_bounds__(kFirThreads) k_fir_real_p2(const float *dbuf, int h2, int Jp, long long n_out, long long n_valid,
                                                             float *out, const __grid_constant__ TapsP taps) {
    extern __shared__ __align__(16) float fsm[];
    float4 *win4 = reinterpret_cast<float4 *>(fsm);   // window only (same padded chunk layout as k_fir_real_r8)
    float *win = fsm;
    const int n_chunks = (kFirOblk + Jp) / 4;
    const int t0 = threadIdx.x * kFirR;
    // grid-stride over tiles: the launch caps the CTAs per SM so that this kernel trickles along UNDER the next call's
    // fused FIR kernel instead of flooding the SMs for its whole duration (fx_path.cu, launch_resample)
    for (long long o0 = (long long)blockIdx.x * kFirOblk; o0 < n_out; o0 += (long long)gridDim.x * kFirOblk) {
        const long long gbase = (long long)h2 + o0 - Jp;
        if ((gbase & 3) == 0 && gbase >= 0 && gbase + kFirOblk + Jp <= n_valid) {
            const float4 *src = reinterpret_cast<const float4 *>(dbuf + gbase);
            for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) win4[c + (c >> 3)] = src[c];
        } else {
            for (int idx = threadIdx.x; idx < kFirOblk + Jp; idx += blockDim.x) {
                long long gi = gbase + idx;
                win[idx + (idx >> 5) * 4] = (gi >= 0 && gi < n_valid) ? dbuf[gi] : 0.f;
            }
        }
        __syncthreads();
        unsigned long long acc[kFirR / 2];
#pragma unroll
        for (int q = 0; q < kFirR / 2; q++) acc[q] = 0ull;
        for (int k0 = 0; k0 < Jp; k0 += kFirTC) {
            // pair q (outputs t0 + 2q, + 1), tap pair k = k0 + kk: sample win[Jp + t0 + 2q + 1 - k] = w[17 + 2q - kk]
            float w[kFirR + kFirTC];
            const int c0 = (Jp + t0 - k0 - kFirTC) >> 2;
#pragma unroll
            for (int c = 0; c < (kFirR + kFirTC) / 4; c++) {
                float4 v = win4[(c0 + c) + ((c0 + c) >> 3)];
                w[4 * c] = v.x, w[4 * c + 1] = v.y, w[4 * c + 2] = v.z, w[4 * c + 3] = v.w;
            }
#pragma unroll
            for (int kk = 0; kk < kFirTC; kk++) {
                const float2 pk = taps.p[k0 + kk];
                const unsigned long long pp = pack_f32x2(pk.x, pk.y);
#pragma unroll
                for (int q = 0; q < kFirR / 2; q++) {
                    const float sm = w[kFirTC + 1 + 2 * q - kk];
                    const unsigned long long ss = pack_f32x2(sm, sm);
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[q]) : "l"(ss), "l"(pp));
                }
            }
        }
        float a[kFirR];
#pragma unroll
        for (int q = 0; q < kFirR / 2; q++) {
            const float2 v = unpack_f32x2(acc[q]);
            a[2 * q] = v.x, a[2 * q + 1] = v.y;
        }
        float *dst = out + o0 + t0;
        if (o0 + t0 + kFirR <= n_out && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int c = 0; c < kFirR / 4; c++)
                reinterpret_cast<float4 *>(dst)[c] = make_float4(a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
        } else {
#pragma unroll
            for (int r = 0; r < kFirR; r++)
                if (o0 + t0 + r < n_out) dst[r] = a[r];
        }
        __syncthreads();   // the window is rewritten by the next tile
    }
}

// Generic rational L/M: one output per thread, polyphase taps in shared memory (row stride Js = J|1 so the
// L phases of a warp hit different banks), the d window of a 256-output tile staged once, 32-bit index
// math relative to a per-tile 64-bit base computed by one thread.
__global__ void __launch_bounds__(256) k_resample_poly(const float *dbuf, int h2, unsigned long long P0, const float *gp,
                                                       int J, uint32_t L, uint32_t M, unsigned long long i0,
                                                       long long n_out, int taps_in_smem, int win_cap, long long n_valid,
                                                       float *out) {
    extern __shared__ __align__(16) float fsm[];
    __shared__ unsigned long long sh_p0;
    __shared__ uint32_t sh_ph0;
    const int Js = J | 1;
    const float *tp = gp;
    int tstride = J;
    float *xw = fsm;                       // staged window of d (when win_cap > 0), after the taps
    if (taps_in_smem) {
        for (int k = threadIdx.x; k < (int)L * J; k += blockDim.x) fsm[(k / J) * Js + (k % J)] = gp[k];
        tp = fsm;
        tstride = Js;
        xw = fsm + ((L * Js + 3) & ~3u);
    }
    for (long long ob = (long long)blockIdx.x * 256; ob < n_out; ob += (long long)gridDim.x * 256) {
        __syncthreads();                   // taps staged / previous tile's readers are done
        if (threadIdx.x == 0) {
            const unsigned long long t0 = (i0 + (unsigned long long)ob) * M;
            const unsigned long long p0 = t0 / L;
            sh_p0 = p0;
            sh_ph0 = (uint32_t)(t0 - p0 * L);
        }
        __syncthreads();
        const unsigned long long p0 = sh_p0;
        const uint32_t ph0 = sh_ph0;
        const long long o = ob + threadIdx.x;
        const uint32_t trel = ph0 + threadIdx.x * M;
        const uint32_t dp = trel / L, ph = trel - dp * L;
        const long long xbase = (long long)(p0 - P0) + h2;       // dbuf index of d[p0]
        const float *g = tp + (size_t)ph * tstride;
        float acc = 0.f;
        if (win_cap) {
            // the 256 outputs of this tile read d[p0 - (J-1) .. p0 + (ph0 + 255*M)/L]: stage it once, coalesced
            const int span = (int)((ph0 + 255u * M) / L) + J;
            for (int k = threadIdx.x; k < span; k += blockDim.x) {
                const long long gi = xbase - (J - 1) + k;
                xw[k] = (gi >= 0 && gi < n_valid) ? dbuf[gi] : 0.f;
            }
            __syncthreads();
            if (o < n_out) {
                const float *x = xw + (J - 1) + dp;
                for (int j = 0; j < J; j++) acc = fmaf(g[j], x[-j], acc);
            }
        } else if (o < n_out) {
            const float *x = dbuf + (xbase + dp);
            for (int j = 0; j < J; j++) acc = fmaf(g[j], x[-j], acc);
        }
        if (o < n_out) out[o] = acc;
    }
}

static const KernelList kFxKernels{(const void *)k_fir_generic, (const void *)k_update_carry, (const void *)k_hist_move,
                                   (const void *)k_fm_demod_f32, (const void *)k_store_prev, (const void *)k_fir_real_r8, (const void *)k_fir_real_p2,
                                   (const void *)k_resample_poly};

}  // namespace sdr

using namespace sdr;

// =================================================================================================
// Host side
// =================================================================================================
namespace {

struct FastVariant {
    int T, D;
    int out_per_cta, hb, smem, nt, wb, b;
    void (*launch)(const FirArgs &, const float *taps, int phase, int grid, int smem, cudaStream_t);
    cudaError_t (*prepare)(int smem);
    cudaError_t (*launch_ring)(const FxRingArgs &, const float *taps, int grid, int smem, cudaStream_t) = nullptr;
    const RtcModule *rtc = nullptr;   // run-time-compiled instance (launch/prepare unused): fns[PH], PH < wb / 2
    int pad = 0;                      // bytes of shared memory skipped after every thread's row (run-time instances)
};

template <int T, int D, int B, int NT, int WB, int PH>
void launch_one(const FirArgs &a, const Taps<T> &t, int grid, int smem, cudaStream_t st) {
    k_fir_fast<T, D, B, NT, WB, PH><<<grid, NT, smem, st>>>(a, t);
}
template <int T, int D, int B, int NT, int WB>
void launch_fast(const FirArgs &a, const float *taps, int phase, int grid, int smem, cudaStream_t st) {
    Taps<T> t;
    memcpy(t.h, taps, sizeof(float) * T);
    switch (phase) {
        case 0: launch_one<T, D, B, NT, WB, 0>(a, t, grid, smem, st); break;
        case 1: launch_one<T, D, B, NT, WB, 1>(a, t, grid, smem, st); break;
        case 2: launch_one<T, D, B, NT, WB, (WB == 8 ? 2 : 0)>(a, t, grid, smem, st); break;
        default: launch_one<T, D, B, NT, WB, (WB == 8 ? 3 : 1)>(a, t, grid, smem, st); break;
    }
}
template <int T, int D, int B, int NT, int WB>
cudaError_t prepare_fast(int smem) {
    cudaError_t e = raise_dyn_smem(k_fir_fast<T, D, B, NT, WB, 0>, smem);
    if (e == cudaSuccess) e = raise_dyn_smem(k_fir_fast<T, D, B, NT, WB, 1>, smem);
    if (WB == 8) {
        if (e == cudaSuccess) e = raise_dyn_smem(k_fir_fast<T, D, B, NT, WB, (WB == 8 ? 2 : 0)>, smem);
        if (e == cudaSuccess) e = raise_dyn_smem(k_fir_fast<T, D, B, NT, WB, (WB == 8 ? 3 : 1)>, smem);
    }
    return e;
}
template <int T, int D, int B, int NT, int WB>
cudaError_t launch_ring_fast(const FxRingArgs &a, const float *taps, int grid, int smem, cudaStream_t st) {
    Taps<T> t;
    memcpy(t.h, taps, sizeof(float) * T);
    cudaError_t e = raise_dyn_smem(k_fmrx_ring<T, D, B, NT, WB>, (size_t)smem);
    if (e != cudaSuccess) return e;
    k_fmrx_ring<T, D, B, NT, WB><<<grid, NT, smem, st>>>(a, t);
    return cudaGetLastError();
}
template <int T, int D, int B, int NT, int WB>
FastVariant make_variant() {
    using G = FastGeom<T, D, B, NT, WB>;
    // the load-phase instantiations of this shape join the preload list (see KernelList in common.cuh)
    static const KernelList kl{(const void *)k_fir_fast<T, D, B, NT, WB, 0>, (const void *)k_fir_fast<T, D, B, NT, WB, 1>,
                               (const void *)k_fir_fast<T, D, B, NT, WB, (WB == 8 ? 2 : 0)>,
                               (const void *)k_fir_fast<T, D, B, NT, WB, (WB == 8 ? 3 : 1)>,
                               (const void *)k_fmrx_ring<T, D, B, NT, WB>};
    return FastVariant{T, D, G::OUT, G::HB, G::SMEM, NT, WB, B, &launch_fast<T, D, B, NT, WB>, &prepare_fast<T, D, B, NT, WB>,
                       &launch_ring_fast<T, D, B, NT, WB>};
}

// Specialised (taps, decimation) shapes: BASELINE.json configs[1] (127, /75) and configs[2] (255, /100),
// plus the reference's own boxcar shape (6, /6) used by the cross-path tests.  D = 100 gives a 200-byte
// thread stride, so 64-bit shared loads are aligned and bank-conflict free; D = 75 (300-byte stride) is
// conflict free with 32-bit loads.
const FastVariant *find_variant(uint32_t T, uint32_t D) {
    // first entry of a shape = default CTA size; SDR_FIR_NT=<threads> selects another compiled size (tuning)
    static const FastVariant table[] = {
        make_variant<127, 75, 2, 32, 4>(),     // 10 KB/CTA: many small CTAs overlap load and compute best
        make_variant<127, 75, 2, 64, 4>(),
        make_variant<127, 75, 2, 128, 4>(),
        make_variant<255, 100, 1, 160, 8>(),   // 37 KB/CTA -> 6 CTAs = 960 threads per SM (measured best)
        make_variant<255, 100, 1, 128, 8>(),
        make_variant<255, 100, 1, 192, 8>(),
        make_variant<255, 100, 1, 64, 8>(),
        make_variant<6, 6, 4, 128, 4>(),
    };
    if (getenv("SDR_FORCE_GENERIC")) return nullptr;
    const char *nt_env = getenv("SDR_FIR_NT"), *b_env = getenv("SDR_FIR_B");
    const int want_nt = nt_env ? atoi(nt_env) : 0, want_b = b_env ? atoi(b_env) : 0;
    const FastVariant *first = nullptr;
    for (const auto &v : table)
        if ((uint32_t)v.T == T && (uint32_t)v.D == D) {
            if (!first) first = &v;
            if (want_nt && v.nt == want_nt && (!want_b || v.b == want_b)) return &v;
        }
    return first;
}

// ---- run-time specialisation: k_fir_fast for a (taps, decimation) shape that has no pre-compiled instance --------
// (B, NT, WB) are chosen the way the table above was tuned by hand:
//   * a thread owns ~128 samples (B blocks), with at most 16 packed accumulators (B * Q);
//   * the thread stride in load units decides the shared-memory bank conflicts of the per-thread loads: a warp's
//     LDS.32 hits gcd(stride, 32) ways, an LDS.64 gcd(stride, 16) ways; 16 bytes of padding per thread row (staged by
//     one bulk copy per row instead of one per tile) shifts the stride off the powers of two — minimise wavefronts
//     per sample;
//   * the CTA is the smallest multiple of 32 threads whose halo (>= Q blocks of NT*B) costs <= 3 % and whose tile is
//     >= 6 KB, within 56 KB of shared memory (several CTAs per SM overlap copy and compute).
bool pick_fast_params(int T, int D, int *Bo, int *NTo, int *WBo, int *PADo) {
    // D <= 256 and <= 16 packed accumulators keep the fully unrolled body (B*D samples x Q lags) small enough for
    // ptxas to finish in seconds
    if (T < 1 || D < 1 || T > 4096 || D > 256) return false;
    const int Q = (T + D - 1) / D;
    if (Q > 16) return false;   // too many lags per sample for register accumulators: generic kernel
    if ((long)((D & 1) ? 2 : 1) * D * Q > 640) return false;   // smallest possible body already takes ptxas the better part of a minute
    const int bmax = std::max(1, std::min(16 / Q, std::max(160 / D, (D & 1) ? 2 : 1)));   // odd D: B must be even
    int best_b = 0, best_wb = 0, best_pad = 0;
    double best_cost = 1e30;
    for (int B = 1; B <= bmax; B++) {
        const int span = B * D;
        if ((long)span * Q > 640 && B > ((D & 1) ? 2 : 1)) break;
        for (int WB : {8, 4}) {
            const int spl = WB / 2;
            if (span % spl) continue;
            // 16 bytes of padding per row (rows must then be 16-byte multiples) changes the stride by 16 / WB units
            for (int pad : {0, 16}) {
                if (pad && (span * 2) % 16) continue;
                const int stride = span * 2 / WB + pad / WB;
                const int ways = WB == 8 ? std::__gcd(stride, 16) : std::__gcd(stride, 32);
                // wavefronts per sample; then a small charge for the row-wise staging, and the distance of the span
                // from ~128 samples as the tie-break
                const double cost = (double)ways / spl + (pad ? 0.05 : 0.0) + 1e-3 * std::abs(span - 128) / 128.0 +
                                    (span < 48 ? 0.5 : 0.0);
                if (cost < best_cost) best_cost = cost, best_b = B, best_wb = WB, best_pad = pad;
            }
        }
    }
    if (!best_b) return false;
    const int B = best_b, WB = best_wb, PAD = best_pad, spl = WB / 2;
    int pick = 0;
    for (int NT = 32; NT <= 512; NT += 32) {
        const int nblk = NT * B, hb = fast_pick_hb(Q, nblk, D, spl);
        const long smem = (((long)nblk * D * 2 + 15) / 16) * 16 + 32 + (long)(NT + 1) * PAD + (long)nblk * fast_qp(Q) * 8 + (long)nblk * 8;
        if (nblk - hb <= hb) continue;
        if (smem > 56 * 1024) break;
        pick = NT;
        if (nblk >= 32 * hb && (long)nblk * D * 2 >= 6 * 1024) break;
    }
    if (!pick) return false;
    *Bo = B, *NTo = pick, *WBo = WB, *PADo = PAD;
    return true;
}

std::string fast_name_expr(int T, int D, int B, int NT, int WB, int PH, int PAD) {
    char buf[128];
    snprintf(buf, sizeof buf, "sdr::k_fir_fast<%d,%d,%d,%d,%d,%d,%d>", T, D, B, NT, WB, PH, PAD);
    return buf;
}

// Build (or fetch) the run-time-compiled variant of a shape.  nullptr + *why when it cannot be had.
const FastVariant *rtc_variant(int device, uint32_t T, uint32_t D, int B, int NT, int WB, int PAD, std::string *why) {
    static std::mutex mu;
    static std::map<std::string, FastVariant *> cache;
    char key[96];
    snprintf(key, sizeof key, "fir_fast<%u,%u,%d,%d,%d,%d>", T, D, B, NT, WB, PAD);
    std::lock_guard<std::mutex> lk(mu);
    const std::string full = std::to_string(device) + "|" + key;
    auto it = cache.find(full);
    if (it != cache.end()) return it->second;
    const int Q = ((int)T + (int)D - 1) / (int)D, spl = WB / 2, nblk = NT * B;
    const int hb = fast_pick_hb(Q, nblk, (int)D, spl);
    const int smem = ((nblk * (int)D * 2 + 15) / 16) * 16 + 32 + (NT + 1) * PAD + nblk * fast_qp(Q) * 8 + nblk * 8;
    std::vector<std::string> names;
    for (int ph = 0; ph < spl; ph++) names.push_back(fast_name_expr((int)T, (int)D, B, NT, WB, ph, PAD));
    const RtcModule *mod = nullptr;
    if (rtc_get_module(device, key, names, smem, &mod)) {
        if (why) *why = err_buf();
        return nullptr;
    }
    FastVariant *v = new FastVariant{(int)T, (int)D, nblk - hb, hb, smem, NT, WB, B, nullptr, nullptr, nullptr, mod, PAD};
    cache[full] = v;
    return v;
}


// ---- output-owner kernel (k_fir_slide, fir_fast.cuh): shapes the block-owner form cannot take (more than 16 lags per
// sample, unrolled body too large) and, measured, the very small decimations ---------------------------------------
// shapes that HAVE a block-owner instance but run faster on the output-owner kernel (measured)
bool prefer_slide(int T, int D) {
    // decimations up to 4 with 8 or more lags per sample: (31,/2) 0.75 vs 0.36 TB/s, (63,/4) 0.85 vs 0.58, (8,/1) 0.63 vs 0.30;
    // with few lags the block-owner form stays ahead ((15,/4) 1.6 vs 1.1), and from /5 up it always is ((129,/16) 1.9 vs 1.0)
    return D <= 4 && (T + D - 1) / D >= 8;
}
struct SlideVariant {
    int T, D, R, nt, smem, opc;
    const RtcModule *rtc;
};
// R outputs per thread: an unrolled body of about 768 FFMA2 at most (ptxas time), a tile of at most 72 KB
bool pick_slide_params(int T, int D, int *Ro, int *NTo) {
    if (T < 1 || D < 1 || T > 640 || D > 64) return false;
    int R = std::max(1, std::min(8, 768 / T));
    for (;;) {
        for (int NT : {128, 64, 32})
            if (slide_smem(T, D, R, NT) <= 72 * 1024) {
                *Ro = R, *NTo = NT;
                return true;
            }
        if (R == 1) return false;
        R--;
    }
}
const SlideVariant *rtc_slide_variant(int device, uint32_t T, uint32_t D, int R, int NT, std::string *why) {
    static std::mutex mu;
    static std::map<std::string, SlideVariant *> cache;
    char key[96], name[128];
    snprintf(key, sizeof key, "fir_slide<%u,%u,%d,%d>", T, D, R, NT);
    snprintf(name, sizeof name, "sdr::k_fir_slide<%u,%u,%d,%d>", T, D, R, NT);
    std::lock_guard<std::mutex> lk(mu);
    const std::string full = std::to_string(device) + "|" + key;
    auto it = cache.find(full);
    if (it != cache.end()) return it->second;
    const int smem = slide_smem((int)T, (int)D, R, NT);
    const RtcModule *mod = nullptr;
    if (rtc_get_module(device, key, {std::string(name)}, smem, &mod)) {
        if (why) *why = err_buf();
        return nullptr;
    }
    SlideVariant *v = new SlideVariant{(int)T, (int)D, R, NT, smem, NT * R - 1, mod};
    cache[full] = v;
    return v;
}

}  // namespace
namespace sdr {
// the shape table is a function-local static: make sure it exists (and has listed its kernels) before a preload
void fx_touch_variants() { (void)find_variant(0, 0); }
}  // namespace sdr
namespace {

constexpr size_t kFxChunkSamples = size_t(16) << 20;   // 32 MiB of IQ per pipelined chunk (host API)
constexpr size_t kFxSmallCallBytes = size_t(1) << 20;  // calls up to 1 MiB take the one-stream path

}  // namespace

struct sdr_fmrx {
    sdr_fmrx_config cfg{};
    int device = 0;
    std::vector<float> taps, taps2;
    float gain = 0.f;
    const FastVariant *fast = nullptr;
    int kernel_kind = 0;     // 0 generic (k_fir_generic), 1 pre-compiled k_fir_fast, 2 run-time-compiled k_fir_fast, 3 run-time-compiled k_fir_slide
    const struct SlideVariant *slide = nullptr;
    std::string rtc_note;    // why run-time compilation was not used, when it was wanted
    // generic geometry
    int gen_opc = 0;
    uint32_t gen_sm_tile = 0, gen_smem = 0;
    // carry: last `cs` samples of the stream (ping-pong), right-aligned so carry_end is 16-B aligned
    int cs = 0;
    DevBuf d_carry[2];
    int carry_cur = 0;
    int h2 = 0;              // discriminator history length kept for the resampler
    int J = 0, Jp = 0;       // polyphase taps per phase (J = ceil(T2/L)); Jp = J rounded up to 8
    DevBuf d_taps, d_taps2, d_state;
    // discriminator buffers [history h2 | new values]: three of them rotate — call k's FIR kernel fills buffer k%3 and
    // writes the head (history) of buffer (k+1)%3, while the audio kernel of call k reads buffer k%3 on its own stream
    // UNDER the FIR kernel of call k+1 (the FIR kernel is HBM-bound and leaves the FMA pipes idle; the audio FIR works
    // out of L2).  A buffer is written again two calls later, after an event wait on its audio kernel.
    DevBuf d_x[2], d_dbuf[3], d_audio[2], d_y[2], d_tmp;
    int dcur = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, audio_stream = nullptr;
    cudaEvent_t ev_fir[3]{}, ev_aud[3]{}, ev_join = nullptr;
    bool aud_used[3] = {false, false, false};
    bool audio_serial = false;
    int audio_ctas_per_sm = 2;   // resident CTAs per SM of an audio kernel that runs under a fused FIR kernel
    H2DStager stager;        // pageable caller buffers go through pinned pieces (common.cuh)
    struct sdr_fmrx_ring *ring = nullptr;   // a persistent ring owns the handle until sdr_fmrx_ring_close()
    TapsP *taps_p = nullptr;   // L = M = 1: overlapping tap pairs of k_fir_real_p2 (null: k_fir_real_r8)
    int Jpp = 0;               // taps of that kernel: J + 1 rounded up to 16
    static constexpr int kRing = 64;   // per-call kernel timings are harvested lazily from this ring
    cudaEvent_t ev_h2d[2]{}, ev_done[2]{}, ev_ring[kRing][4]{}, ev_s[2]{};
    cudaEvent_t *ev_t = ev_ring[0];
    bool ring_used[kRing]{};
    uint64_t ring_next = 0, sum_calls = 0;
    double sum_ms[3] = {0, 0, 0};
    // closed-form stream position
    uint64_t n_in = 0, n_y = 0, n_a = 0;
    // stage-level resampler position (sdr_fmrx_resample keeps its own history in d_dbuf as well)
    float last_ms[3] = {0, 0, 0};
    uint32_t last_launches = 0;
    bool timing_pending = false;
};

namespace {

int fx_reset_device_state(sdr_fmrx *r) {
    for (int i = 0; i < 2; i++) SDR_CUDA_TRY(cudaMemsetAsync(r->d_carry[i].p, 127, (size_t)r->cs * 2, r->stream));
    SDR_CUDA_TRY(cudaMemsetAsync(r->d_state.p, 0, 64, r->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(r->audio_stream));
    for (int i = 0; i < 3; i++) {
        SDR_CUDA_TRY(cudaMemsetAsync(r->d_dbuf[i].p, 0, r->d_dbuf[i].cap, r->stream));
        r->aud_used[i] = false;
    }
    SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    r->carry_cur = 0;
    r->dcur = 0;
    r->n_in = r->n_y = r->n_a = 0;
    return SDR_OK;
}

inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// entry of every call that advances or rewinds the stream: not while a persistent ring owns the handle
int fx_enter(sdr_fmrx *r) {
    if (r->ring) return fail(SDR_E_STATE, "handle is owned by an open ring (sdr_fmrx_ring_close first)");
    return use_device(r->device);
}

struct CallPlan {
    uint64_t n_y, n_a, a0;   // FIR outputs, audio outputs, first audio index
    uint32_t r;
};
CallPlan plan_call(const sdr_fmrx *r, uint64_t n_in0, uint64_t n_y0, size_t n) {
    CallPlan p;
    const uint64_t D = r->cfg.decim;
    p.r = (uint32_t)(n_in0 % D);
    p.n_y = (n_in0 + n) / D - n_in0 / D;
    if (r->cfg.n_taps2) {
        const uint64_t L = r->cfg.up, M = r->cfg.down;
        p.a0 = ceil_div(n_y0 * L, M);
        p.n_a = ceil_div((n_y0 + p.n_y) * L, M) - p.a0;
    } else {
        p.a0 = n_y0;
        p.n_a = p.n_y;
    }
    return p;
}

int ensure_dbuf(sdr_fmrx *r, size_t n_y) {
    size_t need = ((size_t)r->h2 + n_y + 8) * sizeof(float);
    if (need <= r->d_dbuf[0].cap) return SDR_OK;
    // grow all three while preserving the history at the front of the current one; everything in flight first
    SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(r->audio_stream));
    for (int i = 0; i < 3; i++) {
        DevBuf nb;
        int rc = nb.reserve(need * 2);
        if (rc) return rc;
        SDR_CUDA_TRY(cudaMemsetAsync(nb.p, 0, nb.cap, r->stream));
        if (i == r->dcur && r->d_dbuf[i].p)
            SDR_CUDA_TRY(cudaMemcpyAsync(nb.p, r->d_dbuf[i].p, (size_t)r->h2 * sizeof(float), cudaMemcpyDeviceToDevice, r->stream));
        SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
        r->d_dbuf[i].release();
        r->d_dbuf[i] = nb;
        r->aud_used[i] = false;
    }
    return SDR_OK;
}

int launch_carry_update(sdr_fmrx *r, const uint8_t *d_x, size_t n);

// Launch FIR(+demod) for one call-chunk whose input is resident at d_x.  d_demod_target: where the
// discriminator output goes (nullptr = none).
// hist_out: head of the next discriminator buffer (the FIR kernel writes the last h2 values there; needs n_out >= h2), or
// nullptr.  The stream carry of the next call is always written by the kernel's last CTA (ping-pong buffer).
int launch_fir(sdr_fmrx *r, const uint8_t *d_x, size_t n, uint32_t rphase, uint64_t n_out, float2 *d_y, float *d_d,
               float *hist_out = nullptr) {
    if (n_out == 0) return launch_carry_update(r, d_x, n);   // nothing to filter: only the carry moves on
    FirArgs a{};
    a.carry_out = r->d_carry[r->carry_cur ^ 1].as<uint16_t>();
    a.cs = r->cs;
    a.hist_out = hist_out;
    a.h2 = r->h2;
    a.x = d_x;
    a.carry_end = r->d_carry[r->carry_cur].as<uint8_t>() + (size_t)r->cs * 2;
    a.n_samples = (long long)n;
    a.n_out = (long long)n_out;
    a.r = rphase;
    a.gain = r->gain;
    a.y_out = d_y;
    a.d_out = d_d;
    a.last_y = r->d_state.as<float2>();
    if (r->fast) {
        const FastVariant *v = r->fast;
        uint64_t grid = ceil_div(n_out, (uint64_t)v->out_per_cta);
        if (grid > 0x7fffffffull) return fail(SDR_E_ARG, "call too large for one launch");
        // load phase of every tile: sample s0 = -(HB*D + r) sits at byte (2*s0 mod 16) of its 16-byte line
        const long long s0 = -((long long)v->hb * r->cfg.decim + rphase);
        const uint32_t soff = (uint32_t)((2 * s0) & 15);
        const int phase = (int)((soff % (uint32_t)v->wb) / 2);
        if (v->rtc) {
            void *params[2] = {&a, (void *)r->taps.data()};   // (FirArgs, Taps<T>) by value
            int rc = rtc_launch(v->rtc->fns[phase % (v->wb / 2)], (unsigned)grid, (unsigned)v->nt, (unsigned)v->smem, r->stream, params);
            if (rc) return rc;
            r->last_launches++;
            r->carry_cur ^= 1;
            return SDR_OK;
        }
        v->launch(a, r->taps.data(), phase, (int)grid, v->smem, r->stream);
    } else if (r->slide) {
        const SlideVariant *v = r->slide;
        uint64_t grid = ceil_div(n_out, (uint64_t)v->opc);
        if (grid > 0x7fffffffull) return fail(SDR_E_ARG, "call too large for one launch");
        void *params[2] = {&a, (void *)r->taps.data()};   // (FirArgs, Taps<T>) by value
        int rc = rtc_launch(v->rtc->fns[0], (unsigned)grid, (unsigned)v->nt, (unsigned)v->smem, r->stream, params);
        if (rc) return rc;
        r->last_launches++;
        r->carry_cur ^= 1;
        return SDR_OK;
    } else {
        GenArgs g{};
        g.f = a;
        g.taps = r->d_taps.as<float>();
        g.T = (int)r->cfg.n_taps;
        g.D = (int)r->cfg.decim;
        g.OPC = r->gen_opc;
        g.sm_tile = r->gen_sm_tile;
        uint64_t grid = ceil_div(n_out, (uint64_t)g.OPC);
        if (grid > 0x7fffffffull) return fail(SDR_E_ARG, "call too large for one launch");
        k_fir_generic<<<(int)grid, 128, r->gen_smem, r->stream>>>(g);
    }
    SDR_LAUNCH_CHECK();
    r->last_launches++;
    r->carry_cur ^= 1;
    return SDR_OK;
}

int launch_carry_update(sdr_fmrx *r, const uint8_t *d_x, size_t n) {
    if (n == 0) return SDR_OK;
    int nxt = r->carry_cur ^ 1;
    int blocks = (r->cs + 255) / 256;
    k_update_carry<<<blocks, 256, 0, r->stream>>>(r->d_carry[r->carry_cur].as<uint16_t>(),
                                                  reinterpret_cast<const uint16_t *>(d_x), (long long)n, r->cs,
                                                  r->d_carry[nxt].as<uint16_t>());
    SDR_LAUNCH_CHECK();
    r->last_launches++;
    r->carry_cur = nxt;
    return SDR_OK;
}

// Audio stage over dbuf = [hist h2 | d[P0 .. P0+n_new)] on stream `st`.
int launch_resample(sdr_fmrx *r, const float *dbuf, uint64_t P0, uint64_t n_new, uint64_t a0, uint64_t n_a, float *d_audio,
                    cudaStream_t st) {
    if (n_a && r->cfg.up == 1 && r->cfg.down == 1) {
        uint64_t blocks = ceil_div(n_a, (uint64_t)kFirOblk);
        if (blocks > 0x7fffffffull) return fail(SDR_E_ARG, "call too large for one launch");
        size_t sm = fir_real_smem(r->Jp);
        if (r->taps_p) {
            // on its own stream the kernel runs under the next call's fused kernel: cap its CTAs per SM (SDR_FMRX_AUDIO_CTAS)
            const uint64_t cap = st == r->audio_stream ? (uint64_t)sm_count(r->device) * r->audio_ctas_per_sm : blocks;
            k_fir_real_p2<<<(int)std::min(blocks, std::max<uint64_t>(cap, 1)), kFirThreads, fir_real_smem(r->Jpp), st>>>(
                dbuf, r->h2, r->Jpp, (long long)n_a, (long long)(r->h2 + n_new), d_audio, *r->taps_p);
        }
        else
            k_fir_real_r8<<<(int)blocks, kFirThreads, sm, st>>>(dbuf, r->h2, r->d_taps2.as<float>(), r->Jp, (long long)n_a,
                                                                (long long)(r->h2 + n_new), d_audio);
        SDR_LAUNCH_CHECK();
        r->last_launches++;
    } else if (n_a) {
        int blocks = (int)std::min<uint64_t>(ceil_div(n_a, 256),
                                             (uint64_t)sm_count(r->device) * (st == r->audio_stream ? 2 * r->audio_ctas_per_sm : 16));
        size_t tb = (((size_t)r->cfg.up * (r->J | 1) + 3) & ~size_t(3)) * sizeof(float);
        int in_smem = tb <= 32 * 1024;
        // window of d one 256-output tile touches; staged in shared memory when it is small
        size_t span = ((size_t)(r->cfg.up - 1) + 255ull * r->cfg.down) / r->cfg.up + r->J + 1;
        int win_cap = (in_smem && span <= 3072) ? (int)span : 0;
        size_t sm = (in_smem ? tb : 0) + (size_t)win_cap * sizeof(float);
        k_resample_poly<<<blocks, 256, sm, st>>>(dbuf, r->h2, P0, r->d_taps2.as<float>(), r->J, r->cfg.up, r->cfg.down, a0,
                                                 (long long)n_a, in_smem, win_cap, (long long)(r->h2 + n_new), d_audio);
        SDR_LAUNCH_CHECK();
        r->last_launches++;
    }
    return SDR_OK;
}

// The history of the next call when the FIR kernel could not write it (fewer than h2 new values): tail of the current
// buffer -> head of the next one, on the main stream.
int launch_hist_move(sdr_fmrx *r, int cur, int nxt, uint64_t n_new) {
    k_hist_move<<<(r->h2 + 255) / 256, 256, 0, r->stream>>>(r->d_dbuf[cur].as<float>(), (long long)n_new, r->h2,
                                                            r->d_dbuf[nxt].as<float>());
    SDR_LAUNCH_CHECK();
    r->last_launches++;
    return SDR_OK;
}

void harvest_slot(sdr_fmrx *r, int slot) {
    if (!r->ring_used[slot]) return;
    r->ring_used[slot] = false;
    if (cudaEventSynchronize(r->ev_ring[slot][3]) != cudaSuccess || cudaEventSynchronize(r->ev_ring[slot][1]) != cudaSuccess) return;
    // events 0-1 bracket the FIR kernel on the main stream, 2-3 the audio kernel on the audio stream (it may run under
    // the next call's FIR kernel); [2] = bookkeeping kernels, folded into the FIR kernel by now
    for (int i = 0; i < 2; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r->ev_ring[slot][2 * i], r->ev_ring[slot][2 * i + 1]) == cudaSuccess) {
            r->last_ms[i] = ms;
            r->sum_ms[i] += ms;
        }
    }
    r->last_ms[2] = 0.f;
    r->sum_calls++;
}

// Harvest every outstanding timing slot in submission order (the last one harvested is the latest call).
void collect_timing(sdr_fmrx *r) {
    r->timing_pending = false;
    for (uint64_t k = r->ring_next >= sdr_fmrx::kRing ? r->ring_next - sdr_fmrx::kRing : 0; k < r->ring_next; k++)
        harvest_slot(r, (int)(k % sdr_fmrx::kRing));
}

// Claim the next ring slot for a timed call.
void next_timing_slot(sdr_fmrx *r) {
    int slot = (int)(r->ring_next % sdr_fmrx::kRing);
    harvest_slot(r, slot);
    r->ev_t = r->ev_ring[slot];
    r->ring_used[slot] = true;
    r->ring_next++;
}

// One chunk, everything resident.  d_demod / d_y optional; d_audio required when the chain has a tail.  The audio stage
// (if any) is enqueued on the audio stream; `audio_done` (optional) is recorded behind it there.
int run_chunk(sdr_fmrx *r, const uint8_t *d_x, size_t n, float2 *d_y, float *d_demod, float *d_audio,
              const CallPlan &pl, bool timed, bool serial_audio = false) {
    int rc;
    const bool has_res = r->cfg.n_taps2 != 0;
    if ((rc = ensure_dbuf(r, pl.n_y))) return rc;
    const int cur = r->dcur, nxt = (cur + 1) % 3;
    float *dnew = r->d_dbuf[cur].as<float>() + r->h2;
    // without a resample stage the discriminator output IS the audio
    float *d_target = has_res ? dnew : (d_audio ? d_audio : dnew);
    // this call writes the body of `cur` and the head of `nxt`; the last reader of either is the audio kernel of the
    // call two back (buffer index nxt), which runs on the other stream
    // serial: the audio kernel follows the FIR kernel on the main stream (small synchronous calls: one stream, one wait;
    // SDR_FMRX_AUDIO_STREAM=serial pins it for A/B runs)
    const bool serial = serial_audio || r->audio_serial;
    cudaStream_t ast = serial ? r->stream : r->audio_stream;
    if (has_res && r->aud_used[nxt]) {
        SDR_CUDA_TRY(cudaStreamWaitEvent(r->stream, r->ev_aud[nxt], 0));
        r->aud_used[nxt] = false;
    }
    if (has_res && serial && r->aud_used[cur]) {   // this buffer's last reader was an overlapped audio kernel three calls back
        SDR_CUDA_TRY(cudaStreamWaitEvent(r->stream, r->ev_aud[cur], 0));
        r->aud_used[cur] = false;
    }
    if (timed) {
        next_timing_slot(r);
        SDR_CUDA_TRY(cudaEventRecord(r->ev_t[0], r->stream));
    }
    const bool fold_hist = has_res && pl.n_y >= (uint64_t)r->h2;
    if ((rc = launch_fir(r, d_x, n, pl.r, pl.n_y, d_y, d_target, fold_hist ? r->d_dbuf[nxt].as<float>() : nullptr))) return rc;
    if (timed) SDR_CUDA_TRY(cudaEventRecord(r->ev_t[1], r->stream));
    if (has_res) {
        if (!fold_hist && (rc = launch_hist_move(r, cur, nxt, pl.n_y))) return rc;
        if (!serial) {
            SDR_CUDA_TRY(cudaEventRecord(r->ev_fir[cur], r->stream));
            SDR_CUDA_TRY(cudaStreamWaitEvent(ast, r->ev_fir[cur], 0));
        }
        if (timed) SDR_CUDA_TRY(cudaEventRecord(r->ev_t[2], ast));
        if ((rc = launch_resample(r, r->d_dbuf[cur].as<float>(), r->n_y, pl.n_y, pl.a0, pl.n_a, d_audio, ast))) return rc;
        if (timed) SDR_CUDA_TRY(cudaEventRecord(r->ev_t[3], ast));
        if (!serial) {
            SDR_CUDA_TRY(cudaEventRecord(r->ev_aud[cur], ast));
            r->aud_used[cur] = true;
        }
        r->dcur = nxt;
    } else if (timed) {
        SDR_CUDA_TRY(cudaEventRecord(r->ev_t[2], r->stream));
        SDR_CUDA_TRY(cudaEventRecord(r->ev_t[3], r->stream));
    }
    if (d_demod && d_demod != d_target && pl.n_y)
        SDR_CUDA_TRY(cudaMemcpyAsync(d_demod, d_target, pl.n_y * sizeof(float), cudaMemcpyDeviceToDevice, r->stream));
    if (timed) r->timing_pending = true;
    r->n_in += n;
    r->n_y += pl.n_y;
    r->n_a += pl.n_a;
    return SDR_OK;
}

// main stream waits for everything enqueued on the audio stream so far
int join_audio(sdr_fmrx *r) {
    SDR_CUDA_TRY(cudaEventRecord(r->ev_join, r->audio_stream));
    SDR_CUDA_TRY(cudaStreamWaitEvent(r->stream, r->ev_join, 0));
    return SDR_OK;
}

}  // namespace

extern "C" {

int sdr_fmrx_new(const sdr_fmrx_config *cfg, const float *taps, const float *taps2, int cuda_device, sdr_fmrx **out) {
    if (!cfg || !taps || !out) return fail(SDR_E_ARG, "sdr_fmrx_new: null argument");
    if (cfg->n_taps < 1 || cfg->decim < 1) return fail(SDR_E_ARG, "need n_taps >= 1 and decim >= 1");
    if (cfg->n_taps2 && (!taps2 || cfg->up < 1 || cfg->down < 1)) return fail(SDR_E_ARG, "resampler needs taps2, up >= 1, down >= 1");
    if (cfg->n_taps > (1u << 20) || cfg->decim > (1u << 20)) return fail(SDR_E_ARG, "n_taps/decim too large");
    int rc = use_device(cuda_device);
    if (rc) return rc;
    sdr_fmrx *r = new sdr_fmrx();
    r->cfg = *cfg;
    r->device = cuda_device;
    r->taps.assign(taps, taps + cfg->n_taps);
    if (cfg->n_taps2) r->taps2.assign(taps2, taps2 + cfg->n_taps2);
    r->gain = cfg->gain != 0.f ? cfg->gain : (float)(16384.0 / 3.14159265358979323846);
    r->fast = find_variant(cfg->n_taps, cfg->decim);
    r->kernel_kind = r->fast ? 1 : 0;
    {
        // SDR_FIR_RTC=0: never compile at run time; SDR_FIR_RTC=force: also re-compile the pre-compiled shapes (tests)
        const char *er = getenv("SDR_FIR_RTC");
        const bool off = (er && !strcmp(er, "0")) || getenv("SDR_FORCE_GENERIC"), force = er && !strcmp(er, "force");
        int B = 0, NT = 0, WB = 0, PAD = 0;
        bool have = false;
        if (r->fast && force) B = r->fast->b, NT = r->fast->nt, WB = r->fast->wb, have = true;
        else if (!r->fast && !off) have = pick_fast_params((int)cfg->n_taps, (int)cfg->decim, &B, &NT, &WB, &PAD);
        if (have) {
            const FastVariant *v = rtc_variant(cuda_device, cfg->n_taps, cfg->decim, B, NT, WB, PAD, &r->rtc_note);
            if (v) r->fast = v, r->kernel_kind = 2;
        } else if (!r->fast && !off) {
            r->rtc_note = "shape outside the block-owner kernel's range (decim > 256, more than 16 lags per sample, or an unrolled body of more than 640 tap-samples)";
        }
        // output-owner kernel: where the block-owner form has no instance, and (SDR_FIR_SLIDE=1) wherever it exists
        const char *es = getenv("SDR_FIR_SLIDE");
        const bool slide_force = es && atoi(es) == 1, slide_off = es && !strcmp(es, "0");
        int R = 0, SNT = 0;
        if (!off && !slide_off && (slide_force || r->kernel_kind == 0 || prefer_slide((int)cfg->n_taps, (int)cfg->decim)) &&
            r->kernel_kind != 1 && pick_slide_params((int)cfg->n_taps, (int)cfg->decim, &R, &SNT)) {
            std::string why;
            const SlideVariant *sv = rtc_slide_variant(cuda_device, cfg->n_taps, cfg->decim, R, SNT, &why);
            if (sv) r->slide = sv, r->fast = nullptr, r->kernel_kind = 3, r->rtc_note.clear();
            else if (r->kernel_kind == 0) r->rtc_note = why;
        }
    }
    const uint64_t T = cfg->n_taps, D = cfg->decim;
    // generic geometry: ~32 KB of raw bytes per CTA
    uint64_t opc = (16384 > T ? (16384 - T) : 0) / D;
    if (opc < 1) opc = 1;
    if (opc > 64) opc = 64;
    r->gen_opc = (int)opc;
    uint64_t gen_tile = (((opc + 1) * D + T) * 2 + 15 + 32) & ~15ull;
    r->gen_sm_tile = (uint32_t)gen_tile;
    uint64_t gen_smem = gen_tile + ((T + 3) & ~3ull) * 4 + (opc + 2) * 8;
    r->gen_smem = (uint32_t)gen_smem;
    if (gen_smem > 200 * 1024) {
        delete r;
        return fail(SDR_E_ARG, "n_taps/decim too large for the shared-memory tile (%llu bytes)", (unsigned long long)gen_smem);
    }
    // carry holds enough history for either kernel: T + D (generic), (HB+1)*D (fast); multiple of 8 samples
    uint64_t cs = T + D + 8;
    if (r->fast && (uint64_t)(r->fast->hb + 1) * D + 8 > cs) cs = (uint64_t)(r->fast->hb + 1) * D + 8;
    cs = (cs + 7) & ~7ull;
    r->cs = (int)cs;
    if (cfg->n_taps2) {
        r->J = (int)((cfg->n_taps2 + cfg->up - 1) / cfg->up);
        r->Jp = (r->J + 15) & ~15;
        r->h2 = r->Jp + 16;
    }
    cudaError_t e = raise_dyn_smem(k_fir_generic, gen_smem);
    if (e == cudaSuccess && r->fast && !r->fast->rtc) e = r->fast->prepare(r->fast->smem);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        // the audio kernel of one call runs under the FIR kernel of the next: highest priority, so that its CTAs are
        // placed as soon as FIR CTAs retire
        int lo = 0, hi = 0;
        e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        // SDR_FMRX_AUDIO_STREAM: "serial" = the audio kernel follows the FIR kernel on the main stream, "low" = own stream
        // at the lowest priority, default = own stream at the highest priority (A/B runs)
        const char *ea = getenv("SDR_FMRX_AUDIO_STREAM");
        r->audio_serial = ea && !strcmp(ea, "serial");
        if (const char *ec = getenv("SDR_FMRX_AUDIO_CTAS")) r->audio_ctas_per_sm = std::max(1, atoi(ec));
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&r->audio_stream, cudaStreamNonBlocking, (ea && !strcmp(ea, "low")) ? lo : hi);
    }
    for (int i = 0; i < 3 && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&r->ev_fir[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_aud[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&r->ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_done[i], cudaEventDisableTiming);
    }
    for (int k = 0; k < sdr_fmrx::kRing && e == cudaSuccess; k++)
        for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&r->ev_ring[k][i]);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreate(&r->ev_s[i]);
    if (e != cudaSuccess) {
        sdr_fmrx_free(r);
        return fail(SDR_E_CUDA, "sdr_fmrx_new: %s", cudaGetErrorString(e));
    }
    if ((rc = r->d_carry[0].reserve(cs * 2)) || (rc = r->d_carry[1].reserve(cs * 2)) || (rc = r->d_state.reserve(64)) ||
        (rc = r->d_taps.reserve(T * 4)) || (rc = r->d_taps2.reserve(cfg->n_taps2 ? ((size_t)cfg->up * r->J + r->Jp + 16) * 4 : 4)) ||
        (rc = r->d_dbuf[0].reserve(((size_t)r->h2 + 4096) * sizeof(float))) ||
        (rc = r->d_dbuf[1].reserve(((size_t)r->h2 + 4096) * sizeof(float))) ||
        (rc = r->d_dbuf[2].reserve(((size_t)r->h2 + 4096) * sizeof(float)))) {
        sdr_fmrx_free(r);
        return rc;
    }
    e = cudaMemcpyAsync(r->d_taps.p, taps, T * 4, cudaMemcpyHostToDevice, r->stream);
    if (e == cudaSuccess && cfg->n_taps2) {
        // polyphase layout gp[phase][j] = g[phase + j*L] (zero padded); for L = 1 this is g padded to Jp
        std::vector<float> gp((size_t)cfg->up * r->J + r->Jp + 16, 0.f);
        for (uint32_t ph = 0; ph < cfg->up; ph++)
            for (int j = 0; j < r->J; j++) {
                size_t k = ph + (size_t)j * cfg->up;
                if (k < cfg->n_taps2) gp[(size_t)ph * r->J + j] = taps2[k];
            }
        e = cudaMemcpyAsync(r->d_taps2.p, gp.data(), gp.size() * 4, cudaMemcpyHostToDevice, r->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(r->stream);   // gp is a local
        size_t sm = fir_real_smem(r->Jp);
        if (e == cudaSuccess && sm > 48 * 1024)
            e = raise_dyn_smem(k_fir_real_r8, sm);
        // packed audio FIR (L = M = 1): tap pairs P[k] = (G[k], G[k+1]), G[0] = 0, G[k] = g[k-1]
        const char *ep = getenv("SDR_FIR_AUDIO_P2");
        if (cfg->up == 1 && cfg->down == 1 && !(ep && atoi(ep) == 0)) {
            const int Jpp = (int)((cfg->n_taps2 + 1 + 15) & ~15u);
            if (Jpp <= kFirPMaxJ && Jpp <= r->h2) {
                r->taps_p = new TapsP();
                r->Jpp = Jpp;
                auto G = [&](int k) { return (k >= 1 && k <= (int)cfg->n_taps2) ? taps2[k - 1] : 0.f; };
                for (int k = 0; k < kFirPMaxJ; k++) r->taps_p->p[k] = k < Jpp ? make_float2(G(k), G(k + 1)) : make_float2(0.f, 0.f);
                if (e == cudaSuccess && fir_real_smem(Jpp) > 48 * 1024) e = raise_dyn_smem(k_fir_real_p2, fir_real_smem(Jpp));
            }
        }
    }
    if (e != cudaSuccess) {
        sdr_fmrx_free(r);
        return fail(SDR_E_CUDA, "sdr_fmrx_new: %s", cudaGetErrorString(e));
    }
    if ((rc = fx_reset_device_state(r))) {
        sdr_fmrx_free(r);
        return rc;
    }
    *out = r;
    return SDR_OK;
}

void sdr_fmrx_free(sdr_fmrx *r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->ring) sdr_fmrx_ring_close(r->ring);   // retire a ring that is still open
    if (r->stream) cudaStreamSynchronize(r->stream);
    if (r->copy_stream) cudaStreamSynchronize(r->copy_stream);
    if (r->audio_stream) cudaStreamSynchronize(r->audio_stream);
    for (int i = 0; i < 3; i++) {
        r->d_dbuf[i].release();
        if (r->ev_fir[i]) cudaEventDestroy(r->ev_fir[i]);
        if (r->ev_aud[i]) cudaEventDestroy(r->ev_aud[i]);
    }
    if (r->ev_join) cudaEventDestroy(r->ev_join);
    if (r->audio_stream) cudaStreamDestroy(r->audio_stream);
    r->stager.release();
    delete r->taps_p;
    for (int i = 0; i < 2; i++) {
        r->d_carry[i].release();
        r->d_x[i].release();
        r->d_audio[i].release();
        r->d_y[i].release();
        if (r->ev_h2d[i]) cudaEventDestroy(r->ev_h2d[i]);
        if (r->ev_done[i]) cudaEventDestroy(r->ev_done[i]);
    }
    for (int k = 0; k < sdr_fmrx::kRing; k++)
        for (int i = 0; i < 4; i++)
            if (r->ev_ring[k][i]) cudaEventDestroy(r->ev_ring[k][i]);
    for (int i = 0; i < 2; i++)
        if (r->ev_s[i]) cudaEventDestroy(r->ev_s[i]);
    r->d_taps.release();
    r->d_taps2.release();
    r->d_state.release();
    r->d_tmp.release();
    if (r->stream) cudaStreamDestroy(r->stream);
    if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
    delete r;
}

int sdr_fmrx_reset(sdr_fmrx *r) {
    if (!r) return fail(SDR_E_ARG, "null handle");
    int rc = fx_enter(r);
    if (rc) return rc;
    return fx_reset_device_state(r);
}

int sdr_fmrx_out_lens(const sdr_fmrx *r, size_t n_samples, size_t *n_y, size_t *n_audio) {
    if (!r) return fail(SDR_E_ARG, "null handle");
    CallPlan p = plan_call(r, r->n_in, r->n_y, n_samples);
    if (n_y) *n_y = (size_t)p.n_y;
    if (n_audio) *n_audio = (size_t)p.n_a;
    return SDR_OK;
}

long sdr_fmrx_process(sdr_fmrx *r, const uint8_t *iq, size_t n_samples, float *y_pairs, size_t y_cap, float *demod,
                      size_t demod_cap, float *audio, size_t audio_cap) {
    if (!r || (!iq && n_samples)) return fail(SDR_E_ARG, "sdr_fmrx_process: null argument");
    int rc = fx_enter(r);
    if (rc) return rc;
    CallPlan total = plan_call(r, r->n_in, r->n_y, n_samples);
    if (y_pairs && total.n_y > y_cap) return fail(SDR_E_CAP, "y capacity %zu < %llu", y_cap, (unsigned long long)total.n_y);
    if (demod && total.n_y > demod_cap) return fail(SDR_E_CAP, "demod capacity %zu < %llu", demod_cap, (unsigned long long)total.n_y);
    if (!audio && total.n_a) return fail(SDR_E_ARG, "audio buffer required");
    if (total.n_a > audio_cap) return fail(SDR_E_CAP, "audio capacity %zu < %llu", audio_cap, (unsigned long long)total.n_a);
    r->last_launches = 0;
    if (n_samples * 2 <= kFxSmallCallBytes && !getenv("SDR_FMRX_NO_SMALL_PATH")) {
        // one USB-sized buffer per call (the reference's call pattern, examples/simple_fm.rs:80,153): everything on ONE
        // stream — copy in, FIR+demod, audio stage, copies out, one wait
        const size_t max_y = n_samples / r->cfg.decim + 8;
        if ((rc = r->d_x[0].reserve(kFxSmallCallBytes + 64)) || (rc = r->d_audio[0].reserve((total.n_a + 8) * 4)) ||
            (y_pairs && (rc = r->d_y[0].reserve(max_y * 8))) || (demod && (rc = r->d_tmp.reserve(max_y * 4))))
            return rc;
        SDR_CUDA_TRY(cudaMemcpyAsync(r->d_x[0].p, iq, n_samples * 2, cudaMemcpyHostToDevice, r->stream));
        if ((rc = run_chunk(r, r->d_x[0].as<uint8_t>(), n_samples, y_pairs ? r->d_y[0].as<float2>() : nullptr,
                            demod ? r->d_tmp.as<float>() : nullptr, r->d_audio[0].as<float>(), total, true, true)))
            return rc;
        if (y_pairs && total.n_y) SDR_CUDA_TRY(cudaMemcpyAsync(y_pairs, r->d_y[0].p, total.n_y * 8, cudaMemcpyDeviceToHost, r->stream));
        if (demod && total.n_y) SDR_CUDA_TRY(cudaMemcpyAsync(demod, r->d_tmp.p, total.n_y * 4, cudaMemcpyDeviceToHost, r->stream));
        if (total.n_a) SDR_CUDA_TRY(cudaMemcpyAsync(audio, r->d_audio[0].p, total.n_a * 4, cudaMemcpyDeviceToHost, r->stream));
        SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
        collect_timing(r);
        return (long)total.n_a;
    }
    // chunk on a multiple of 8 samples so that every chunk starts 16-byte aligned in the caller's buffer
    const size_t chunk = kFxChunkSamples;
    size_t done = 0, y_off = 0, a_off = 0;
    int ci = 0;
    while (done < n_samples) {
        const int slot = ci & 1;
        const size_t n = std::min(chunk, n_samples - done);
        CallPlan pl = plan_call(r, r->n_in, r->n_y, n);
        if ((rc = r->d_x[slot].reserve(std::min(chunk, n_samples) * 2 + 64))) return rc;
        const size_t max_y = std::min(chunk, n_samples) / r->cfg.decim + 8;
        const size_t max_a = r->cfg.n_taps2 ? (size_t)ceil_div((uint64_t)max_y * r->cfg.up, r->cfg.down) + 8 : max_y;
        if ((rc = r->d_audio[slot].reserve(max_a * 4))) return rc;
        if (y_pairs && (rc = r->d_y[slot].reserve((chunk / r->cfg.decim + 8) * 8))) return rc;
        if (demod && (rc = r->d_tmp.reserve((chunk / r->cfg.decim + 8) * 4 * 2))) return rc;
        if (ci >= 2) SDR_CUDA_TRY(cudaStreamWaitEvent(r->copy_stream, r->ev_done[slot], 0));
        if ((rc = r->stager.copy(r->d_x[slot].p, iq + done * 2, n * 2, r->copy_stream))) return rc;
        SDR_CUDA_TRY(cudaEventRecord(r->ev_h2d[slot], r->copy_stream));
        SDR_CUDA_TRY(cudaStreamWaitEvent(r->stream, r->ev_h2d[slot], 0));
        float *d_dem = demod ? r->d_tmp.as<float>() + (size_t)slot * (chunk / r->cfg.decim + 8) : nullptr;
        if ((rc = run_chunk(r, r->d_x[slot].as<uint8_t>(), n, y_pairs ? r->d_y[slot].as<float2>() : nullptr, d_dem,
                            r->d_audio[slot].as<float>(), pl, done + n >= n_samples)))
            return rc;
        if (y_pairs && pl.n_y)
            SDR_CUDA_TRY(cudaMemcpyAsync(y_pairs + 2 * y_off, r->d_y[slot].p, pl.n_y * 8, cudaMemcpyDeviceToHost, r->stream));
        if (demod && pl.n_y)
            SDR_CUDA_TRY(cudaMemcpyAsync(demod + y_off, d_dem, pl.n_y * 4, cudaMemcpyDeviceToHost, r->stream));
        if (pl.n_a)   // behind the kernel that produced it: the audio stream when there is a resample stage
            SDR_CUDA_TRY(cudaMemcpyAsync(audio + a_off, r->d_audio[slot].p, pl.n_a * 4, cudaMemcpyDeviceToHost,
                                         (r->cfg.n_taps2 && !r->audio_serial) ? r->audio_stream : r->stream));
        SDR_CUDA_TRY(cudaEventRecord(r->ev_done[slot], r->stream));
        done += n;
        y_off += pl.n_y;
        a_off += pl.n_a;
        ci++;
    }
    SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(r->copy_stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(r->audio_stream));
    collect_timing(r);
    return (long)total.n_a;
}

long sdr_fmrx_process_dev(sdr_fmrx *r, const uint8_t *d_iq, size_t n_samples, float *d_y_pairs, float *d_demod,
                          float *d_audio, size_t audio_cap) {
    if (!r || (!d_iq && n_samples)) return fail(SDR_E_ARG, "sdr_fmrx_process_dev: null argument");
    if (reinterpret_cast<uintptr_t>(d_iq) & 15) return fail(SDR_E_ARG, "device input must be 16-byte aligned (use sdr_dev_alloc)");
    int rc = fx_enter(r);
    if (rc) return rc;
    CallPlan pl = plan_call(r, r->n_in, r->n_y, n_samples);
    if (!d_audio && pl.n_a) return fail(SDR_E_ARG, "audio buffer required");
    if (pl.n_a > audio_cap) return fail(SDR_E_CAP, "audio capacity %zu < %llu", audio_cap, (unsigned long long)pl.n_a);
    r->last_launches = 0;
    if ((rc = run_chunk(r, d_iq, n_samples, reinterpret_cast<float2 *>(d_y_pairs), d_demod, d_audio, pl, true))) return rc;
    return (long)pl.n_a;
}

long sdr_fmrx_low_pass(sdr_fmrx *r, const uint8_t *iq, size_t n_samples, float *y_pairs, size_t cap_pairs) {
    if (!r || (!iq && n_samples) || !y_pairs) return fail(SDR_E_ARG, "sdr_fmrx_low_pass: null argument");
    int rc = fx_enter(r);
    if (rc) return rc;
    CallPlan pl = plan_call(r, r->n_in, r->n_y, n_samples);
    if (pl.n_y > cap_pairs) return fail(SDR_E_CAP, "capacity %zu < %llu pairs", cap_pairs, (unsigned long long)pl.n_y);
    if (n_samples == 0) return 0;
    if ((rc = r->d_x[0].reserve(n_samples * 2 + 64)) || (rc = r->d_y[0].reserve((pl.n_y + 8) * 8))) return rc;
    r->last_launches = 0;
    if ((rc = r->stager.copy(r->d_x[0].p, iq, n_samples * 2, r->stream))) return rc;
    if ((rc = launch_fir(r, r->d_x[0].as<uint8_t>(), n_samples, pl.r, pl.n_y, r->d_y[0].as<float2>(), nullptr))) return rc;
    if (pl.n_y) SDR_CUDA_TRY(cudaMemcpyAsync(y_pairs, r->d_y[0].p, pl.n_y * 8, cudaMemcpyDeviceToHost, r->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    r->n_in += n_samples;
    // n_y (the discriminator/resampler stream position) is advanced by the stages that consume y
    return (long)pl.n_y;
}

long sdr_fmrx_fm_demod(sdr_fmrx *r, const float *y_pairs, size_t n, float *out, size_t cap) {
    if (!r || (!y_pairs && n) || !out) return fail(SDR_E_ARG, "sdr_fmrx_fm_demod: null argument");
    if (n > cap) return fail(SDR_E_CAP, "capacity %zu < %zu", cap, n);
    int rc = fx_enter(r);
    if (rc) return rc;
    if (n == 0) return 0;
    if ((rc = r->d_y[0].reserve(n * 8)) || (rc = r->d_tmp.reserve(n * 4))) return rc;
    SDR_CUDA_TRY(cudaMemcpyAsync(r->d_y[0].p, y_pairs, n * 8, cudaMemcpyHostToDevice, r->stream));
    int blocks = (int)std::min<uint64_t>(ceil_div(n, 256), (uint64_t)sm_count(r->device) * 16);
    k_fm_demod_f32<<<blocks, 256, 0, r->stream>>>(r->d_y[0].as<float2>(), (long long)n, r->d_state.as<float2>(), r->gain,
                                                  r->d_tmp.as<float>());
    SDR_LAUNCH_CHECK();
    k_store_prev<<<1, 1, 0, r->stream>>>(r->d_y[0].as<float2>(), (long long)n, r->d_state.as<float2>());
    SDR_LAUNCH_CHECK();
    SDR_CUDA_TRY(cudaMemcpyAsync(out, r->d_tmp.p, n * 4, cudaMemcpyDeviceToHost, r->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    return (long)n;
}

long sdr_fmrx_resample(sdr_fmrx *r, const float *d, size_t n, float *out, size_t cap) {
    if (!r || (!d && n) || !out) return fail(SDR_E_ARG, "sdr_fmrx_resample: null argument");
    if (!r->cfg.n_taps2) return fail(SDR_E_STATE, "handle was created without a resample stage");
    int rc = fx_enter(r);
    if (rc) return rc;
    const uint64_t L = r->cfg.up, M = r->cfg.down;
    uint64_t a0 = ceil_div(r->n_y * L, M), n_a = ceil_div((r->n_y + n) * L, M) - a0;
    if (n_a > cap) return fail(SDR_E_CAP, "capacity %zu < %llu", cap, (unsigned long long)n_a);
    if (n == 0) return 0;
    if ((rc = ensure_dbuf(r, n)) || (rc = r->d_audio[0].reserve((n_a + 8) * 4))) return rc;
    SDR_CUDA_TRY(cudaStreamSynchronize(r->audio_stream));   // stage entry points are synchronous: one stream from here
    const int cur = r->dcur, nxt = (cur + 1) % 3;
    SDR_CUDA_TRY(cudaMemcpyAsync(r->d_dbuf[cur].as<float>() + r->h2, d, n * 4, cudaMemcpyHostToDevice, r->stream));
    if ((rc = launch_resample(r, r->d_dbuf[cur].as<float>(), r->n_y, n, a0, n_a, r->d_audio[0].as<float>(), r->stream))) return rc;
    if ((rc = launch_hist_move(r, cur, nxt, n))) return rc;
    r->dcur = nxt;
    if (n_a) SDR_CUDA_TRY(cudaMemcpyAsync(out, r->d_audio[0].p, n_a * 4, cudaMemcpyDeviceToHost, r->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    r->n_y += n;
    r->n_a += n_a;
    return (long)n_a;
}

int sdr_fmrx_sync(sdr_fmrx *r) {
    if (!r) return fail(SDR_E_ARG, "null handle");
    int rc = use_device(r->device);
    if (rc) return rc;
    SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(r->audio_stream));
    collect_timing(r);
    return SDR_OK;
}

int sdr_fmrx_span_begin(sdr_fmrx *r) {
    if (!r) return fail(SDR_E_ARG, "null handle");
    int rc = use_device(r->device);
    if (rc) return rc;
    if ((rc = join_audio(r))) return rc;   // the span starts when everything submitted before it is done
    SDR_CUDA_TRY(cudaEventRecord(r->ev_s[0], r->stream));
    return SDR_OK;
}

int sdr_fmrx_span_end(sdr_fmrx *r, float *ms) {
    if (!r || !ms) return fail(SDR_E_ARG, "null argument");
    int rc = use_device(r->device);
    if (rc) return rc;
    if ((rc = join_audio(r))) return rc;   // ... and ends when the last audio kernel has finished
    SDR_CUDA_TRY(cudaEventRecord(r->ev_s[1], r->stream));
    SDR_CUDA_TRY(cudaEventSynchronize(r->ev_s[1]));
    SDR_CUDA_TRY(cudaEventElapsedTime(ms, r->ev_s[0], r->ev_s[1]));
    collect_timing(r);
    return SDR_OK;
}

int sdr_fmrx_seek(sdr_fmrx *r, uint64_t global_sample_index) {
    if (!r) return fail(SDR_E_ARG, "null handle");
    int rc = fx_enter(r);
    if (rc) return rc;
    if ((rc = fx_reset_device_state(r))) return rc;
    r->n_in = global_sample_index;
    r->n_y = global_sample_index / r->cfg.decim;
    r->n_a = r->cfg.n_taps2 ? ceil_div(r->n_y * r->cfg.up, r->cfg.down) : r->n_y;
    return SDR_OK;
}

int sdr_fmrx_last_timing(const sdr_fmrx *r, float ms[3], uint32_t *n_launches, int *specialised) {
    if (!r) return fail(SDR_E_ARG, "null handle");
    if (ms)
        for (int i = 0; i < 3; i++) ms[i] = r->last_ms[i];
    if (n_launches) *n_launches = r->last_launches;
    if (specialised) *specialised = (r->fast || r->slide) ? 1 : 0;
    return SDR_OK;
}

int sdr_fmrx_kernel_kind(const sdr_fmrx *r, const char **note) {
    if (!r) return fail(SDR_E_ARG, "null handle");
    if (note) *note = r->rtc_note.c_str();
    return r->kernel_kind;
}

int sdr_rtc_pick_shape(uint32_t n_taps, uint32_t decim, int shape[4]) {
    int B = 0, NT = 0, WB = 0, PAD = 0;
    if (n_taps > 4096 || decim > 256 || !pick_fast_params((int)n_taps, (int)decim, &B, &NT, &WB, &PAD)) return 0;
    if (shape) shape[0] = B, shape[1] = NT, shape[2] = WB, shape[3] = PAD;
    // the kernel's own compile-time requirements (FastGeom's static_asserts), re-stated on the host
    const int T = (int)n_taps, D = (int)decim, Q = (T + D - 1) / D, spl = WB / 2, nblk = NT * B;
    const int hb = fast_pick_hb(Q, nblk, D, spl), row = B * D * 2;
    const long smem = (((long)nblk * D * 2 + 15) / 16) * 16 + 32 + (long)(NT + 1) * PAD + (long)nblk * fast_qp(Q) * 8 + (long)nblk * 8;
    if (WB != 4 && WB != 8) return fail(SDR_E_STATE, "picked load width %d", WB);
    if ((B * D) % spl) return fail(SDR_E_STATE, "thread span %d is not a whole number of %d-sample loads", B * D, spl);
    if (nblk - hb <= hb) return fail(SDR_E_STATE, "tile of %d blocks is all halo (%d)", nblk, hb);
    if (PAD && (row % 16 || PAD % 16)) return fail(SDR_E_STATE, "padded rows of %d bytes cannot be bulk-copied", row);
    if (NT % 32 || NT < 32 || NT > 1024) return fail(SDR_E_STATE, "CTA of %d threads", NT);
    if (smem > 227 * 1024) return fail(SDR_E_STATE, "%ld bytes of shared memory", smem);
    if (B * Q > 16) return fail(SDR_E_STATE, "%d packed accumulators", B * Q);
    return 1;
}

long sdr_rtc_selftest(uint32_t n_taps, uint32_t decim, int shape[4]) {
    int B = 0, NT = 0, WB = 0, PAD = 0;
    if (n_taps > 4096 || decim > 256 || !pick_fast_params((int)n_taps, (int)decim, &B, &NT, &WB, &PAD))
        return fail(SDR_E_ARG, "(%u taps, /%u) is outside the block-owner kernel's range", n_taps, decim);
    if (shape) shape[0] = B, shape[1] = NT, shape[2] = WB + 256 * PAD;
    std::vector<std::string> names;
    for (int ph = 0; ph < WB / 2; ph++) names.push_back(fast_name_expr((int)n_taps, (int)decim, B, NT, WB, ph, PAD));
    std::vector<std::vector<char>> cubins;
    std::vector<std::string> lowered;
    int n_compiled = 0;
    int rc = rtc_compile_cubins(names, &cubins, &lowered, &n_compiled);
    if (rc) return rc;
    if (shape) shape[3] = n_compiled;
    size_t total = 0;
    for (const auto &c : cubins) total += c.size();
    return (long)total;
}

long sdr_rtc_compile_ring(uint32_t n_taps, uint32_t decim) {
    int B = 0, NT = 0, WB = 0, PAD = 0;
    if (n_taps > 4096 || decim > 256 || !pick_fast_params((int)n_taps, (int)decim, &B, &NT, &WB, &PAD))
        return fail(SDR_E_ARG, "(%u taps, /%u) is outside the block-owner kernel's range", n_taps, decim);
    char name[160];
    snprintf(name, sizeof name, "sdr::k_fmrx_ring<%u,%u,%d,%d,%d,%d>", n_taps, decim, B, NT, WB, PAD);
    std::vector<std::vector<char>> cubins;
    std::vector<std::string> lowered;
    int rc = rtc_compile_cubins({name}, &cubins, &lowered, nullptr);
    if (rc) return rc;
    return (long)cubins[0].size();
}

long sdr_rtc_compile_slide(uint32_t n_taps, uint32_t decim, int shape[2]) {
    int R = 0, NT = 0;
    if (!pick_slide_params((int)n_taps, (int)decim, &R, &NT))
        return fail(SDR_E_ARG, "(%u taps, /%u) is outside the output-owner kernel's range", n_taps, decim);
    if (shape) shape[0] = R, shape[1] = NT;
    char name[160];
    snprintf(name, sizeof name, "sdr::k_fir_slide<%u,%u,%d,%d>", n_taps, decim, R, NT);
    std::vector<std::vector<char>> cubins;
    std::vector<std::string> lowered;
    int rc = rtc_compile_cubins({name}, &cubins, &lowered, nullptr);
    if (rc) return rc;
    return (long)cubins[0].size();
}

int sdr_fmrx_timing_totals(sdr_fmrx *r, double sums_ms[3], uint64_t *n_calls, int reset) {
    if (!r) return fail(SDR_E_ARG, "null handle");
    int rc = use_device(r->device);
    if (rc) return rc;
    collect_timing(r);
    if (sums_ms)
        for (int i = 0; i < 3; i++) sums_ms[i] = r->sum_ms[i];
    if (n_calls) *n_calls = r->sum_calls;
    if (reset) {
        r->sum_ms[0] = r->sum_ms[1] = r->sum_ms[2] = 0.0;
        r->sum_calls = 0;
    }
    return SDR_OK;
}

}  // extern "C"

// =================================================================================================
// Persistent ring of the f32 receiver (host side; kernel: k_fmrx_ring in fir_fast.cuh)
// =================================================================================================
struct sdr_fmrx_ring {
    sdr_fmrx *r = nullptr;
    size_t buf_len = 0, slot_stride = 0, out_stride = 0, dbuf_stride = 0;
    uint64_t S = 0, n_in0 = 0, n_y0 = 0;
    uint32_t n_slots = 0, m = 0;
    uint8_t *h_in = nullptr;               // pinned [n_slots][buf_len]
    DevBuf d_in, d_ctl, d_dbuf;
    float *h_out = nullptr, *h_out_dev = nullptr;   // host-mapped audio slots
    volatile unsigned int *h_seq_done = nullptr;
    unsigned int *h_seq_done_dev = nullptr;
    unsigned int *h_doorbell = nullptr;    // pinned [n_slots] + stop word at [64]
    cudaStream_t ring_stream = nullptr, copy_stream = nullptr;
    std::atomic<uint64_t> head{0}, tail{0};   // committed / collected buffers
    bool acquired = false, launched = false;
};

namespace {

void fx_ring_release(sdr_fmrx_ring *g) {
    if (!g) return;
    if (g->launched) ring_closed(g->r->device);
    if (g->r && g->r->ring == g) g->r->ring = nullptr;
    if (g->ring_stream) cudaStreamDestroy(g->ring_stream);
    if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
    host_free_or_park(g->h_in);
    host_free_or_park(g->h_out);
    host_free_or_park((void *)g->h_seq_done);
    host_free_or_park(g->h_doorbell);
    g->d_in.release();
    g->d_ctl.release();
    g->d_dbuf.release();
    delete g;
}

// audio range [a0, a1) of buffer k (global audio indices), closed form
void fx_ring_audio_range(const sdr_fmrx_ring *g, uint64_t k, uint64_t *a0, uint64_t *a1) {
    const sdr_fmrx *r = g->r;
    const uint64_t D = r->cfg.decim, n0 = g->n_in0 + k * g->S, y0 = n0 / D, y1 = (n0 + g->S) / D;
    if (r->cfg.n_taps2) {
        *a0 = ceil_div(y0 * r->cfg.up, r->cfg.down);
        *a1 = ceil_div(y1 * r->cfg.up, r->cfg.down);
    } else {
        *a0 = y0;
        *a1 = y1;
    }
}

}  // namespace

extern "C" {

int sdr_fmrx_ring_open(sdr_fmrx *r, size_t buf_len, uint32_t n_slots, sdr_fmrx_ring **out) {
    if (!r || !out) return fail(SDR_E_ARG, "sdr_fmrx_ring_open: null argument");
    if (n_slots < 2 || n_slots > 64) return fail(SDR_E_ARG, "n_slots must be in [2, 64]");
    if (buf_len == 0 || buf_len % 16) return fail(SDR_E_LEN, "buffer length %zu is not a positive multiple of 16 bytes", buf_len);
    int rc = fx_enter(r);
    if (rc) return rc;
    if (!r->fast) return fail(SDR_E_STATE, "the ring runs the specialised FIR kernel; this (taps, decimation) shape has none (%s)", r->rtc_note.c_str());
    const uint64_t S = buf_len / 2, D = r->cfg.decim, L = r->cfg.up, M = r->cfg.down;
    const bool has_res = r->cfg.n_taps2 != 0;
    if (S < (uint64_t)r->cs) return fail(SDR_E_LEN, "a ring buffer must hold at least the %d carried samples", r->cs);
    if (S / D < (has_res ? (uint64_t)r->h2 : 1) || (has_res && (S / D) * L < M))
        return fail(SDR_E_LEN, "a ring buffer must produce at least %d FIR outputs and one audio sample", has_res ? r->h2 : 1);
    SDR_CUDA_TRY(cudaStreamSynchronize(r->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(r->audio_stream));
    sdr_fmrx_ring *g = new sdr_fmrx_ring();
    g->r = r;
    g->buf_len = buf_len;
    g->S = S;
    g->n_slots = n_slots;
    g->m = n_slots + 1;
    g->slot_stride = buf_len;            // multiple of 16: every slot (and its end, the next buffer's carry) is aligned
    g->n_in0 = r->n_in;
    g->n_y0 = r->n_y;
    const uint64_t max_y = S / D + 1;
    g->dbuf_stride = (size_t)((r->h2 + max_y + 8 + 3) & ~3ull);
    g->out_stride = (size_t)(((has_res ? ceil_div(max_y * L, M) + 1 : max_y) + 3) & ~3ull);
    cudaError_t e = cudaHostAlloc((void **)&g->h_in, (size_t)n_slots * buf_len, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&g->h_out, (size_t)n_slots * g->out_stride * sizeof(float), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&g->h_out_dev, g->h_out, 0);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&g->h_seq_done, 64 * sizeof(unsigned int), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&g->h_seq_done_dev, (void *)g->h_seq_done, 0);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&g->h_doorbell, 72 * sizeof(unsigned int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->ring_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        fx_ring_release(g);
        return fail(SDR_E_CUDA, "sdr_fmrx_ring_open: %s", cudaGetErrorString(e));
    }
    if ((rc = g->d_in.reserve((size_t)g->m * g->slot_stride + 64)) || (rc = g->d_ctl.reserve(sizeof(FxRingCtl))) ||
        (rc = g->d_dbuf.reserve((size_t)g->m * g->dbuf_stride * sizeof(float)))) {
        fx_ring_release(g);
        return rc;
    }
    for (int i = 0; i < 64; i++) g->h_seq_done[i] = 0;
    e = cudaMemsetAsync(g->d_ctl.p, 0, sizeof(FxRingCtl), g->copy_stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->d_dbuf.p, 0, (size_t)g->m * g->dbuf_stride * sizeof(float), g->copy_stream);
    // the resampler history the handle carries becomes the head of buffer 0's discriminator slot
    if (e == cudaSuccess && has_res)
        e = cudaMemcpyAsync(g->d_dbuf.p, r->d_dbuf[r->dcur].p, (size_t)r->h2 * sizeof(float), cudaMemcpyDeviceToDevice, g->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->copy_stream);
    if (e != cudaSuccess) {
        fx_ring_release(g);
        return fail(SDR_E_CUDA, "sdr_fmrx_ring_open: %s", cudaGetErrorString(e));
    }
    // nothing may be loaded lazily once a ring kernel is resident (see KernelList in common.cuh)
    const FastVariant *v = r->fast;
    const RtcModule *ring_mod = nullptr;
    if (v->rtc) {
        char key[112], name[160];
        snprintf(key, sizeof key, "fmrx_ring<%d,%d,%d,%d,%d,%d>", v->T, v->D, v->b, v->nt, v->wb, v->pad);
        snprintf(name, sizeof name, "sdr::k_fmrx_ring<%d,%d,%d,%d,%d,%d>", v->T, v->D, v->b, v->nt, v->wb, v->pad);
        if ((rc = rtc_get_module(r->device, key, {name}, v->smem, &ring_mod))) {
            fx_ring_release(g);
            return rc;
        }
    }
    if ((rc = preload_kernels())) {
        fx_ring_release(g);
        return rc;
    }
    FxRingArgs a{};
    a.ctl = g->d_ctl.as<FxRingCtl>();
    a.seq_done = g->h_seq_done_dev;
    a.d_in = g->d_in.as<uint8_t>();
    a.carry0_end = r->d_carry[r->carry_cur].as<uint8_t>() + (size_t)r->cs * 2;
    a.dbuf = g->d_dbuf.as<float>();
    a.h_out = g->h_out_dev;
    a.gp = r->d_taps2.as<float>();
    a.last_y = r->d_state.as<float2>();
    a.S = S;
    a.n_in0 = g->n_in0;
    a.n_y0 = g->n_y0;
    a.slot_stride = g->slot_stride;
    a.dbuf_stride = g->dbuf_stride;
    a.out_stride = g->out_stride;
    a.n_slots = n_slots;
    a.m = g->m;
    a.D = (unsigned)D;
    a.L = has_res ? (unsigned)L : 1u;
    a.M = has_res ? (unsigned)M : 1u;
    a.J = (unsigned)r->J;
    a.h2 = (unsigned)r->h2;
    a.has_res = has_res ? 1u : 0u;
    a.gain = r->gain;
    // one CTA per FIR tile of a buffer, at most one per SM: every CTA is resident (audio tiles wait for FIR tiles)
    const uint64_t fir_tiles = ceil_div(max_y, (uint64_t)v->out_per_cta), aud_tiles = ceil_div(g->out_stride, (uint64_t)kRingAudioTile);
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(std::max(fir_tiles, aud_tiles), (uint64_t)sm_count(r->device)));
    if (v->rtc) {
        void *params[2] = {&a, (void *)r->taps.data()};
        rc = rtc_launch(ring_mod->fns[0], (unsigned)grid, (unsigned)v->nt, (unsigned)v->smem, g->ring_stream, params);
        if (rc) {
            fx_ring_release(g);
            return rc;
        }
    } else {
        e = v->launch_ring(a, r->taps.data(), grid, v->smem, g->ring_stream);
        if (e != cudaSuccess) {
            fx_ring_release(g);
            return fail(SDR_E_CUDA, "sdr_fmrx_ring_open: launch failed: %s", cudaGetErrorString(e));
        }
        count_launch();   // (rtc_launch counts its own)
    }
    g->launched = true;
    ring_opened(r->device);
    r->ring = g;
    *out = g;
    return SDR_OK;
}

int sdr_fmrx_ring_acquire(sdr_fmrx_ring *g, uint8_t **buf) {
    if (!g || !buf) return fail(SDR_E_ARG, "sdr_fmrx_ring_acquire: null argument");
    if (g->acquired) return fail(SDR_E_STATE, "a slot is already acquired (commit it first)");
    while (g->head.load() - g->tail.load() >= g->n_slots) std::this_thread::yield();   // ring full: wait for collect()
    *buf = g->h_in + (size_t)(g->head.load() % g->n_slots) * g->buf_len;
    g->acquired = true;
    return SDR_OK;
}

int sdr_fmrx_ring_commit(sdr_fmrx_ring *g) {
    if (!g) return fail(SDR_E_ARG, "null ring");
    if (!g->acquired) return fail(SDR_E_STATE, "no slot acquired");
    int rc = use_device(g->r->device);
    if (rc) return rc;
    const uint64_t k = g->head.load();
    const uint32_t hs = (uint32_t)(k % g->n_slots), ds = (uint32_t)(k % g->m);
    SDR_CUDA_TRY(cudaMemcpyAsync(g->d_in.as<uint8_t>() + (size_t)ds * g->slot_stride, g->h_in + (size_t)hs * g->buf_len, g->buf_len,
                                 cudaMemcpyHostToDevice, g->copy_stream));
    g->h_doorbell[hs] = (unsigned int)(k + 1);
    SDR_CUDA_TRY(cudaMemcpyAsync(&g->d_ctl.as<FxRingCtl>()->seq_ready[hs], &g->h_doorbell[hs], sizeof(unsigned int),
                                 cudaMemcpyHostToDevice, g->copy_stream));   // stream order: data first, then the doorbell
    g->acquired = false;
    g->head.store(k + 1);
    return SDR_OK;
}

long sdr_fmrx_ring_collect(sdr_fmrx_ring *g, float *audio, size_t cap) {
    if (!g || !audio) return fail(SDR_E_ARG, "sdr_fmrx_ring_collect: null argument");
    const uint64_t k = g->tail.load();
    if (k == g->head.load()) return fail(SDR_E_STATE, "nothing outstanding");
    const uint32_t hs = (uint32_t)(k % g->n_slots);
    uint64_t a0 = 0, a1 = 0;
    fx_ring_audio_range(g, k, &a0, &a1);
    const uint64_t na = a1 - a0;
    if (na > cap) return fail(SDR_E_CAP, "output capacity %zu < %llu audio samples", cap, (unsigned long long)na);
    uint64_t spins = 0;
    while (g->h_seq_done[hs] != (unsigned int)(k + 1)) {
        if ((++spins & 0xfff) == 0) {
            cudaError_t e = cudaStreamQuery(g->ring_stream);
            if (e != cudaErrorNotReady) {
                (void)cudaGetLastError();
                return fail(SDR_E_CUDA, "ring kernel is not running (%s)", cudaGetErrorString(e));
            }
            std::this_thread::yield();
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(audio, g->h_out + (size_t)hs * g->out_stride, na * sizeof(float));
    g->tail.store(k + 1);
    return (long)na;
}

int sdr_fmrx_ring_close(sdr_fmrx_ring *g) {
    if (!g) return fail(SDR_E_ARG, "null ring");
    sdr_fmrx *r = g->r;
    int rc = use_device(r->device);
    if (rc) return rc;
    g->h_doorbell[64] = 1;
    cudaError_t e = cudaMemcpyAsync(&g->d_ctl.as<FxRingCtl>()->stop, &g->h_doorbell[64], sizeof(unsigned int), cudaMemcpyHostToDevice,
                                    g->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->ring_stream);   // every committed buffer is processed before stop is honoured
    const uint64_t N = g->head.load();
    if (e == cudaSuccess && N) {
        // hand the stream back: position (closed form), the raw carry = tail of the last buffer's slot, the resampler
        // history = head of the slot the next buffer would have used
        const uint64_t D = r->cfg.decim;
        const uint64_t n1 = g->n_in0 + N * g->S;
        uint64_t a_dummy = 0, a1 = 0;
        fx_ring_audio_range(g, N - 1, &a_dummy, &a1);
        const uint8_t *last = g->d_in.as<uint8_t>() + (size_t)((N - 1) % g->m) * g->slot_stride;
        e = cudaMemcpyAsync(r->d_carry[r->carry_cur].p, last + 2 * (g->S - (uint64_t)r->cs), (size_t)r->cs * 2, cudaMemcpyDeviceToDevice,
                            g->copy_stream);
        if (e == cudaSuccess && r->cfg.n_taps2)
            e = cudaMemcpyAsync(r->d_dbuf[r->dcur].p, g->d_dbuf.as<float>() + (size_t)(N % g->m) * g->dbuf_stride,
                                (size_t)r->h2 * sizeof(float), cudaMemcpyDeviceToDevice, g->copy_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g->copy_stream);
        if (e == cudaSuccess) {
            r->n_in = n1;
            r->n_y = n1 / D;
            r->n_a = a1;
        }
    }
    fx_ring_release(g);
    if (e != cudaSuccess) return fail(SDR_E_CUDA, "sdr_fmrx_ring_close: %s", cudaGetErrorString(e));
    return SDR_OK;
}

}  // extern "C"
