// util.cu — error plumbing, device/pinned memory, synthetic IQ generator.
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace sdr {

std::atomic<uint64_t> g_launches{0};

char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return fail(SDR_E_CUDA, "no CUDA device available (%s); libsdr_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(SDR_E_ARG, "cuda device %d out of range [0,%d)", device, n);
    SDR_CUDA_TRY(cudaSetDevice(device));
    return SDR_OK;
}

int sm_count(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 148;
    return v > 0 ? v : 148;
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap && p) return SDR_OK;
    release();
    size_t want = bytes + 2 * kDevRoom;
    void *q = nullptr;
    SDR_CUDA_TRY(cudaMalloc(&q, want));
    // cudaMemset runs on the legacy stream, which the handles' non-blocking streams do NOT order against:
    // wait for it, or it could land on top of data a handle stream copies into the new buffer.
    SDR_CUDA_TRY(cudaMemset(q, 127, want));
    SDR_CUDA_TRY(cudaDeviceSynchronize());
    p = static_cast<char *>(q) + kDevRoom;
    cap = bytes;
    return SDR_OK;
}
void DevBuf::release() {
    if (p) cudaFree(static_cast<char *>(p) - kDevRoom);
    p = nullptr;
    cap = 0;
}
int PinBuf::reserve(size_t bytes) {
    if (bytes <= cap && p) return SDR_OK;
    release();
    SDR_CUDA_TRY(cudaMallocHost(&p, bytes ? bytes : 16));
    cap = bytes;
    return SDR_OK;
}
void PinBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// byte i of the stream = mix64(seed, i>>3) >> 8*(i&7).  One 64-bit word per thread-iteration.
__global__ void k_synth_fill(uint8_t *buf, size_t bytes, uint64_t seed, uint64_t byte_offset) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    // head/tail bytes not aligned to the 8-byte word grid of the *stream* are handled bytewise
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i * 8 < bytes + 8; i += stride) {
        uint64_t first = byte_offset & ~7ull;            // stream byte of word 0
        uint64_t wbyte = first + i * 8;                  // stream byte index of this word
        uint64_t w = mix64(seed, wbyte >> 3);
        long long local = (long long)(wbyte - byte_offset);  // local byte index of the word's byte 0
        if (local >= 0 && (size_t)local + 8 <= bytes && ((reinterpret_cast<uintptr_t>(buf) + local) & 7) == 0) {
            *reinterpret_cast<uint64_t *>(buf + local) = w;
        } else {
            for (int b = 0; b < 8; b++) {
                long long l = local + b;
                if (l >= 0 && (size_t)l < bytes) buf[l] = (uint8_t)(w >> (8 * b));
            }
        }
    }
}

}  // namespace sdr

using namespace sdr;

extern "C" {

const char *sdr_last_error(void) { return err_buf(); }
int sdr_abi_version(void) { return SDR_B200_ABI_VERSION; }

int sdr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int sdr_device_info(int device, char *name, size_t cap, int *smc, uint64_t *mem_bytes) {
    int rc = use_device(device);
    if (rc) return rc;
    cudaDeviceProp p;
    SDR_CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (name && cap) {
        strncpy(name, p.name, cap - 1);
        name[cap - 1] = 0;
    }
    if (smc) *smc = p.multiProcessorCount;
    if (mem_bytes) *mem_bytes = (uint64_t)p.totalGlobalMem;
    return SDR_OK;
}

uint64_t sdr_kernel_launch_count(void) { return g_launches.load(); }

void *sdr_dev_alloc(int device, size_t bytes) {
    if (use_device(device)) return nullptr;
    void *q = nullptr;
    size_t want = ((bytes + 255) & ~size_t(255)) + 2 * kDevRoom;
    if (cudaMalloc(&q, want) != cudaSuccess) {
        fail(SDR_E_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    cudaMemset(q, 127, want);   // head/tail room reads as mid-scale (centred zero)
    cudaDeviceSynchronize();    // the legacy-stream memset must not race with non-blocking handle streams
    return static_cast<char *>(q) + kDevRoom;
}
void sdr_dev_free(int device, void *p) {
    if (!p || use_device(device)) return;
    cudaFree(static_cast<char *>(p) - kDevRoom);
}
void *sdr_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) {
        fail(SDR_E_CUDA, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void sdr_host_free(void *p) {
    if (p) cudaFreeHost(p);
}
int sdr_memcpy_h2d(int device, void *dst, const void *src, size_t bytes) {
    int rc = use_device(device);
    if (rc) return rc;
    SDR_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return SDR_OK;
}
int sdr_memcpy_d2h(int device, void *dst, const void *src, size_t bytes) {
    int rc = use_device(device);
    if (rc) return rc;
    SDR_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return SDR_OK;
}
int sdr_dev_memset(int device, void *dst, int value, size_t bytes) {
    int rc = use_device(device);
    if (rc) return rc;
    SDR_CUDA_TRY(cudaMemset(dst, value, bytes));
    return SDR_OK;
}
int sdr_synth_fill_dev(int device, uint8_t *d_buf, size_t bytes, uint64_t seed, uint64_t byte_offset) {
    int rc = use_device(device);
    if (rc) return rc;
    if (!d_buf) return fail(SDR_E_ARG, "sdr_synth_fill_dev: null buffer");
    if (bytes == 0) return SDR_OK;
    int blocks = sm_count(device) * 8;
    k_synth_fill<<<blocks, 256>>>(d_buf, bytes, seed, byte_offset);
    SDR_LAUNCH_CHECK();
    SDR_CUDA_TRY(cudaDeviceSynchronize());
    return SDR_OK;
}
int sdr_device_sync(int device) {
    int rc = use_device(device);
    if (rc) return rc;
    SDR_CUDA_TRY(cudaDeviceSynchronize());
    return SDR_OK;
}

}  // extern "C"
