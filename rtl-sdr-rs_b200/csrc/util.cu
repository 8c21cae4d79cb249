// util.cu — error plumbing, device/pinned memory, synthetic IQ generator.
#include <emmintrin.h>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

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

namespace {
constexpr int kMaxDev = 64;
std::mutex g_mu;
std::unordered_map<uint64_t, size_t> g_dyn_smem;          // (device, kernel) -> largest size set so far
int g_rings[kMaxDev] = {0};                               // open rings per device
std::vector<void *> g_parked_dev[kMaxDev], g_parked_host;
cudaStream_t g_fill_stream[kMaxDev] = {nullptr};

int cur_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 0;
    return dev;
}
}  // namespace

cudaError_t raise_dyn_smem_raw(const void *func, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    const uint64_t key = ((uint64_t)cur_device() << 56) ^ (uint64_t)reinterpret_cast<uintptr_t>(func);
    auto it = g_dyn_smem.find(key);
    // sizes up to the 48 KB default need no attribute at all
    const size_t have = it == g_dyn_smem.end() ? 48 * 1024 : it->second;
    if (bytes <= have) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) g_dyn_smem[key] = bytes;
    return e;
}

void ring_opened(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (device >= 0 && device < kMaxDev) g_rings[device]++;
}
bool ring_is_open(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (device >= 0) return device < kMaxDev && g_rings[device] > 0;
    for (int i = 0; i < kMaxDev; i++)
        if (g_rings[i] > 0) return true;
    return false;
}
void ring_closed(int device) {
    std::vector<void *> dev_list, host_list;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (device < 0 || device >= kMaxDev || g_rings[device] <= 0) return;
        if (--g_rings[device] > 0) return;
        dev_list.swap(g_parked_dev[device]);
        bool any = false;
        for (int i = 0; i < kMaxDev; i++) any = any || g_rings[i] > 0;
        if (!any) host_list.swap(g_parked_host);
    }
    int prev = cur_device();
    cudaSetDevice(device);
    for (void *q : dev_list) cudaFree(q);
    for (void *q : host_list) cudaFreeHost(q);
    cudaSetDevice(prev);
}
void dev_free_or_park(void *raw) {
    if (!raw) return;
    const int dev = cur_device();
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_rings[dev] > 0) {
            g_parked_dev[dev].push_back(raw);
            return;
        }
    }
    cudaFree(raw);
}
void host_free_or_park(void *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (int i = 0; i < kMaxDev; i++)
            if (g_rings[i] > 0) {
                g_parked_host.push_back(p);
                return;
            }
    }
    cudaFreeHost(p);
}
namespace {
std::vector<const void *> &kernel_registry() {
    static std::vector<const void *> v;   // function-local: safe to use from other units' static initialisers
    return v;
}
bool g_preloaded[kMaxDev] = {false};
}  // namespace

KernelList::KernelList(std::initializer_list<const void *> fns) {
    auto &v = kernel_registry();
    v.insert(v.end(), fns.begin(), fns.end());
}

int preload_kernels() {
    const int dev = cur_device();
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_preloaded[dev]) return SDR_OK;
    }
    fx_touch_variants();
    for (const void *fn : kernel_registry()) {
        cudaFuncAttributes attr;
        SDR_CUDA_TRY(cudaFuncGetAttributes(&attr, fn));
    }
    // the memset kernel of dev_fill() as well
    void *q = nullptr;
    SDR_CUDA_TRY(cudaMalloc(&q, 256));
    int rc = dev_fill(q, 0, 256);
    cudaFree(q);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    g_preloaded[dev] = true;
    return SDR_OK;
}

int dev_fill(void *p, int value, size_t bytes) {
    const int dev = cur_device();
    cudaStream_t st;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_fill_stream[dev]) SDR_CUDA_TRY(cudaStreamCreateWithFlags(&g_fill_stream[dev], cudaStreamNonBlocking));
        st = g_fill_stream[dev];
    }
    // NOT the legacy stream (whose memset the handles' non-blocking streams would not order against) and no
    // device-wide wait (which a resident ring kernel would never let return)
    SDR_CUDA_TRY(cudaMemsetAsync(p, value, bytes, st));
    SDR_CUDA_TRY(cudaStreamSynchronize(st));
    return SDR_OK;
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap && p) return SDR_OK;
    release();
    size_t want = bytes + 2 * kDevRoom;
    void *q = nullptr;
    SDR_CUDA_TRY(cudaMalloc(&q, want));
    int rc = dev_fill(q, 127, want);
    if (rc) {
        dev_free_or_park(q);
        return rc;
    }
    p = static_cast<char *>(q) + kDevRoom;
    cap = bytes;
    return SDR_OK;
}
void DevBuf::release() {
    if (p) dev_free_or_park(static_cast<char *>(p) - kDevRoom);
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
    if (p) host_free_or_park(p);
    p = nullptr;
    cap = 0;
}

// ---- pageable-source staging ------------------------------------------------------------------------------------------
namespace {

class CopyPool {
public:
    static CopyPool &get() {
        static CopyPool *p = new CopyPool();   // leaked on purpose: worker threads must not be joined from a static destructor
        return *p;
    }
    void run(char *dst, const char *src, size_t bytes) {
        const size_t n = workers_.size() + 1;
        if (n == 1 || bytes < (size_t(1) << 20)) {
            memcpy(dst, src, bytes);
            return;
        }
        std::unique_lock<std::mutex> lk(mu_);
        const size_t part = ((bytes + n - 1) / n + 4095) & ~size_t(4095);
        dst_ = dst, src_ = src, bytes_ = bytes, part_ = part;
        pending_ = (int)workers_.size();
        gen_++;
        cv_.notify_all();
        lk.unlock();
        copy_part(0);   // the caller is worker 0
        lk.lock();
        done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    CopyPool() {
        int n = 0;
        if (const char *e = getenv("SDR_STAGE_THREADS")) n = atoi(e);
        if (n <= 0) n = (int)std::min<unsigned>(12, std::max<unsigned>(2, std::thread::hardware_concurrency() * 3 / 4));   // measured on a 16-vCPU box: 11 / 18 / 26 / 33 / 36 GB/s with 1 / 2 / 4 / 8 / 12 threads
        for (int i = 1; i < n; i++) workers_.emplace_back([this, i] { loop(i); });
        for (auto &t : workers_) t.detach();
    }
    // The destination is a pinned staging piece that only the copy engine will read: streaming (non-temporal) stores
    // skip the read-for-ownership of every destination line, i.e. a third of the memory traffic of a cached copy.
    static void stream_copy(char *dst, const char *src, size_t n) {
        size_t i = 0;
        if (((uintptr_t)dst & 15) == 0) {
            for (; i + 64 <= n; i += 64) {
                const __m128i a = _mm_loadu_si128((const __m128i *)(src + i)), b = _mm_loadu_si128((const __m128i *)(src + i + 16));
                const __m128i c = _mm_loadu_si128((const __m128i *)(src + i + 32)), d = _mm_loadu_si128((const __m128i *)(src + i + 48));
                _mm_stream_si128((__m128i *)(dst + i), a);
                _mm_stream_si128((__m128i *)(dst + i + 16), b);
                _mm_stream_si128((__m128i *)(dst + i + 32), c);
                _mm_stream_si128((__m128i *)(dst + i + 48), d);
            }
            _mm_sfence();
        }
        if (i < n) memcpy(dst + i, src + i, n - i);
    }
    void copy_part(size_t i) {
        const size_t lo = i * part_;
        if (lo < bytes_) stream_copy(dst_ + lo, src_ + lo, std::min(part_, bytes_ - lo));
    }
    void loop(int idx) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            lk.unlock();
            copy_part((size_t)idx);
            lk.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, part_ = 0;
    int pending_ = 0;
    uint64_t gen_ = 0;
};

constexpr size_t kStagePiece = size_t(8) << 20;

}  // namespace

void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    static std::mutex one_at_a_time;   // the pool runs one job at a time; handles on different threads queue up here
    std::lock_guard<std::mutex> lk(one_at_a_time);
    CopyPool::get().run(static_cast<char *>(dst), static_cast<const char *>(src), bytes);
}

bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

int H2DStager::copy(void *d_dst, const void *h_src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return SDR_OK;
    const char *env = getenv("SDR_STAGE_PAGEABLE");
    if (host_ptr_is_pinned(h_src) || bytes < (size_t(1) << 20) || (env && atoi(env) == 0)) {
        SDR_CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
        return SDR_OK;
    }
    for (size_t off = 0; off < bytes; off += kStagePiece) {
        const int s = next;
        next ^= 1;
        const size_t n = std::min(kStagePiece, bytes - off);
        int rc = piece[s].reserve(kStagePiece);
        if (rc) return rc;
        if (!ev[s]) SDR_CUDA_TRY(cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming));
        if (used[s]) SDR_CUDA_TRY(cudaEventSynchronize(ev[s]));   // the copy engine is done with this piece
        parallel_memcpy(piece[s].p, static_cast<const char *>(h_src) + off, n);
        SDR_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(d_dst) + off, piece[s].p, n, cudaMemcpyHostToDevice, stream));
        SDR_CUDA_TRY(cudaEventRecord(ev[s], stream));
        used[s] = true;
    }
    return SDR_OK;
}

void H2DStager::release() {
    for (int s = 0; s < 2; s++) {
        if (ev[s]) {
            cudaEventSynchronize(ev[s]);
            cudaEventDestroy(ev[s]);
            ev[s] = nullptr;
        }
        piece[s].release();
        used[s] = false;
    }
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

static const KernelList kUtilKernels{(const void *)k_synth_fill};

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
    if (dev_fill(q, 127, want)) {   // head/tail room reads as mid-scale (centred zero)
        dev_free_or_park(q);
        return nullptr;
    }
    return static_cast<char *>(q) + kDevRoom;
}
void sdr_dev_free(int device, void *p) {
    if (!p || use_device(device)) return;
    dev_free_or_park(static_cast<char *>(p) - kDevRoom);
}
void *sdr_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) {
        fail(SDR_E_CUDA, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void sdr_host_free(void *p) { host_free_or_park(p); }
// Page-lock a buffer the caller owns (a long-lived Vec<u8> / Box<[u8; N]>): every call that is handed a pointer inside it
// afterwards copies by DMA straight from it, like memory from sdr_host_alloc, instead of through the staging pieces.
int sdr_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return fail(SDR_E_ARG, "null / empty buffer");
    if (ring_is_open(-1)) return fail(SDR_E_STATE, "a persistent ring is open on this process: register buffers before opening it");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        (void)cudaGetLastError();
        return SDR_OK;
    }
    if (e != cudaSuccess) return fail(SDR_E_CUDA, "cudaHostRegister(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return SDR_OK;
}
int sdr_host_unregister(void *p) {
    if (!p) return SDR_OK;
    if (ring_is_open(-1)) return fail(SDR_E_STATE, "a persistent ring is open on this process: close it before unregistering");
    cudaError_t e = cudaHostUnregister(p);
    if (e == cudaErrorHostMemoryNotRegistered) {
        (void)cudaGetLastError();
        return SDR_OK;
    }
    if (e != cudaSuccess) return fail(SDR_E_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SDR_OK;
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
    cudaStream_t st;
    SDR_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    k_synth_fill<<<blocks, 256, 0, st>>>(d_buf, bytes, seed, byte_offset);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (e != cudaSuccess) return fail(SDR_E_CUDA, "sdr_synth_fill_dev: %s", cudaGetErrorString(e));
    count_launch();
    return SDR_OK;
}
int sdr_device_sync(int device) {
    int rc = use_device(device);
    if (rc) return rc;
    if (ring_is_open(device))
        return fail(SDR_E_STATE, "a persistent ring is resident on device %d: a device-wide wait would never return (close the ring, or sync the handles)", device);
    SDR_CUDA_TRY(cudaDeviceSynchronize());
    return SDR_OK;
}

}  // extern "C"
