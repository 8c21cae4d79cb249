// post.cu — optional audio post-stages after Demod::low_pass_real.  ALL OFF BY DEFAULT: the reference
// computes `output_scale` (examples/simple_fm.rs:184,197-200) and never applies it, and has no de-emphasis, DC block or
// squelch at all — its output (and the golden hash) is the raw low_pass_real stream.  These stages restate what rtl_fm, the
// program simple_fm.rs was ported from, does after its own low_pass_real (integer arithmetic, same order):
//
//   output_scale : audio * scale, saturated to i16                          (this project's definition; rtl_fm never applies it either)
//   squelch      : the call's audio is zeroed when the RMS deviation of its raw bytes from mid-scale is below the level
//                  (rtl_fm gates on the rms of its lowpassed block; the raw-byte form needs nothing from inside the fused kernel)
//   de-emphasis  : avg += round_half_away((x - avg) / a);  x = avg        rtl_fm deemph_filter; a = round(1 / (1 - exp(-1 / (rate * tau))))
//   DC block     : avg = (mean(block) + 9 * dc_avg) / 10;  x -= avg;  dc_avg = avg      rtl_fm dc_block_filter (one block = one call)
//
// De-emphasis is a nonlinear (integer-rounded) recurrence, so it runs as one sequential thread — at audio rate that is a
// few thousand samples per 128 ms buffer (~60 us); the other stages are one small parallel kernel.  "Parity unpinned":
// the test suite checks them against an independent CPU restatement of the same definitions.
#include <cmath>

#include "common.cuh"

namespace sdr {

struct PostState {
    int32_t deemph_avg, dc_avg;
};

// scale (saturating), squelch gate, block mean for the DC blocker.  One CTA.
__global__ void __launch_bounds__(256) k_post_scale_gate(int16_t *x, int n, int scale, int gate_open, long long *sum_out) {
    __shared__ long long red[256];
    long long s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int v = gate_open ? (int)x[i] * scale : 0;
        v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
        x[i] = (int16_t)v;
        s += v;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *sum_out = red[0];
}

// rtl_fm deemph_filter: sequential by construction (one thread).
__global__ void k_post_deemph(int16_t *x, int n, int a, PostState *st) {
    int avg = st->deemph_avg;
    for (int i = 0; i < n; i++) {
        const int d = (int)x[i] - avg;
        avg += d > 0 ? (d + a / 2) / a : (d - a / 2) / a;
        x[i] = (int16_t)avg;
    }
    st->deemph_avg = avg;
}

// rtl_fm dc_block_filter over one block; `sum` = sum of the block as it stands when this stage is reached.
__global__ void __launch_bounds__(256) k_post_dc_block(int16_t *x, int n, PostState *st) {
    __shared__ long long red[256];
    __shared__ int sh_avg;
    long long s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int avg = (int)(red[0] / n);                 // C truncating division, like rtl_fm
        avg = (avg + st->dc_avg * 9) / 10;
        st->dc_avg = avg;
        sh_avg = avg;
    }
    __syncthreads();
    const int avg = sh_avg;
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = (int16_t)((int)x[i] - avg);   // wraps like rtl_fm's int16 store
}

// sum of squared deviations of the raw bytes from mid-scale, in units of (1/2)^2: sum (2*b - 255)^2
__global__ void __launch_bounds__(256) k_post_raw_power(const uint8_t *raw, size_t n, unsigned long long *out) {
    __shared__ unsigned long long red[256];
    unsigned long long s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int v = 2 * (int)raw[i] - 255;
        s += (unsigned long long)(v * v);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

static const KernelList kPostKernels{(const void *)k_post_scale_gate, (const void *)k_post_deemph, (const void *)k_post_dc_block,
                                     (const void *)k_post_raw_power};

}  // namespace sdr

using namespace sdr;

struct sdr_post {
    sdr_post_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    DevBuf d_audio, d_raw, d_state, d_sum;
    PinBuf h_tmp;
};

extern "C" {

uint32_t sdr_post_deemph_a(uint32_t rate, double tau_us) {
    if (rate == 0 || tau_us <= 0) return 0;
    return (uint32_t)std::lround(1.0 / (1.0 - std::exp(-1.0 / ((double)rate * tau_us * 1e-6))));
}

int sdr_post_new(const sdr_post_config *cfg, int cuda_device, sdr_post **out) {
    if (!cfg || !out) return fail(SDR_E_ARG, "sdr_post_new: null argument");
    if (cfg->output_scale > 32767) return fail(SDR_E_ARG, "output_scale must be < 32768");
    int rc = use_device(cuda_device);
    if (rc) return rc;
    sdr_post *p = new sdr_post();
    p->cfg = *cfg;
    p->device = cuda_device;
    cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete p;
        return fail(SDR_E_CUDA, "sdr_post_new: %s", cudaGetErrorString(e));
    }
    if ((rc = p->d_state.reserve(sizeof(PostState))) || (rc = p->d_sum.reserve(16)) || (rc = p->h_tmp.reserve(64))) {
        sdr_post_free(p);
        return rc;
    }
    e = cudaMemsetAsync(p->d_state.p, 0, sizeof(PostState), p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    if (e != cudaSuccess) {
        sdr_post_free(p);
        return fail(SDR_E_CUDA, "sdr_post_new: %s", cudaGetErrorString(e));
    }
    *out = p;
    return SDR_OK;
}

void sdr_post_free(sdr_post *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    p->d_audio.release();
    p->d_raw.release();
    p->d_state.release();
    p->d_sum.release();
    p->h_tmp.release();
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

long sdr_post_process(sdr_post *p, int16_t *audio, size_t n, const uint8_t *raw, size_t raw_len) {
    if (!p || (!audio && n)) return fail(SDR_E_ARG, "sdr_post_process: null argument");
    if (n > (size_t)0x7fffffff) return fail(SDR_E_ARG, "block too large");
    int rc = use_device(p->device);
    if (rc) return rc;
    if (n == 0) return 0;
    const sdr_post_config &c = p->cfg;
    if ((rc = p->d_audio.reserve(n * 2 + 16))) return rc;
    int gate_open = 1;
    if (c.squelch_level && raw && raw_len) {
        if ((rc = p->d_raw.reserve(raw_len))) return rc;
        SDR_CUDA_TRY(cudaMemcpyAsync(p->d_raw.p, raw, raw_len, cudaMemcpyHostToDevice, p->stream));
        SDR_CUDA_TRY(cudaMemsetAsync(p->d_sum.p, 0, 8, p->stream));
        const int blocks = (int)std::min<size_t>((raw_len + 255) / 256, (size_t)sm_count(p->device) * 4);
        k_post_raw_power<<<blocks, 256, 0, p->stream>>>(p->d_raw.as<uint8_t>(), raw_len, p->d_sum.as<unsigned long long>());
        SDR_LAUNCH_CHECK();
        SDR_CUDA_TRY(cudaMemcpyAsync(p->h_tmp.p, p->d_sum.p, 8, cudaMemcpyDeviceToHost, p->stream));
        SDR_CUDA_TRY(cudaStreamSynchronize(p->stream));
        // rms of (b - 127.5) * 16 against the level, compared as squares in integers: sum (2b-255)^2 * 64 < level^2 * raw_len
        const unsigned __int128 lhs = (unsigned __int128)(*p->h_tmp.as<unsigned long long>()) * 64u;
        const unsigned __int128 rhs = (unsigned __int128)c.squelch_level * c.squelch_level * raw_len;
        gate_open = lhs >= rhs;
    }
    SDR_CUDA_TRY(cudaMemcpyAsync(p->d_audio.p, audio, n * 2, cudaMemcpyHostToDevice, p->stream));
    const int scale = c.output_scale ? (int)c.output_scale : 1;
    if (scale != 1 || !gate_open) {
        k_post_scale_gate<<<1, 256, 0, p->stream>>>(p->d_audio.as<int16_t>(), (int)n, scale, gate_open, p->d_sum.as<long long>() + 1);
        SDR_LAUNCH_CHECK();
    }
    if (c.deemph_a) {
        k_post_deemph<<<1, 1, 0, p->stream>>>(p->d_audio.as<int16_t>(), (int)n, (int)c.deemph_a, p->d_state.as<PostState>());
        SDR_LAUNCH_CHECK();
    }
    if (c.dc_block) {
        k_post_dc_block<<<1, 256, 0, p->stream>>>(p->d_audio.as<int16_t>(), (int)n, p->d_state.as<PostState>());
        SDR_LAUNCH_CHECK();
    }
    SDR_CUDA_TRY(cudaMemcpyAsync(audio, p->d_audio.p, n * 2, cudaMemcpyDeviceToHost, p->stream));
    SDR_CUDA_TRY(cudaStreamSynchronize(p->stream));
    return (long)n;
}

}  // extern "C"
