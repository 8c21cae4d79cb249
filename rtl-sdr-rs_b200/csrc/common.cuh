// common.cuh — shared host/device helpers for libsdr_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <initializer_list>

#include "../../include/sdr_b200.h"

namespace sdr {

// ---- error plumbing (thread-local message, negative codes; never throws across the ABI) -------
char *err_buf();
int fail(int code, const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define SDR_CUDA_TRY(expr)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return ::sdr::fail(SDR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                          \
    } while (0)

#define SDR_LAUNCH_CHECK()                                                                   \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess)                                                               \
            return ::sdr::fail(SDR_E_CUDA, "kernel launch failed: %s (%s:%d)",                \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                  \
        ::sdr::count_launch();                                                               \
    } while (0)

// Select a device; SDR_E_CUDA if there is none (the product path has no CPU fallback).
int use_device(int device);
int sm_count(int device);

// cudaFuncAttributeMaxDynamicSharedMemorySize is one value per (kernel, device) for the whole process, while every
// handle wants its own size: only ever RAISE it (running maximum under a mutex), so that a later handle with a smaller
// tile cannot make an earlier handle's launches fail.  Applies to the current device.
cudaError_t raise_dyn_smem_raw(const void *func, size_t bytes);
template <typename F>
inline cudaError_t raise_dyn_smem(F *func, size_t bytes) {
    return raise_dyn_smem_raw(reinterpret_cast<const void *>(func), bytes);
}

// Persistent rings keep one kernel resident on a device.  While one is open there, nothing in the library may call
// anything that waits for the whole device (cudaDeviceSynchronize, cudaFree, cudaFreeHost): allocations clear their
// memory on a private stream, frees are parked and carried out when the last ring on the device closes.
void ring_opened(int device);
void ring_closed(int device);          // frees everything parked for the device when its last ring closes
bool ring_is_open(int device);         // device < 0: on any device
void dev_free_or_park(void *raw);      // cudaFree now, or later if a ring is resident on the current device
void host_free_or_park(void *p);       // cudaFreeHost now, or later if a ring is resident on any device
// fill freshly allocated device memory without touching the legacy stream or the other handles' streams
int dev_fill(void *p, int value, size_t bytes);
// CUDA loads kernels lazily, on their first launch, and that load can wait for the whole context — for ever, behind a
// resident ring kernel.  Every translation unit therefore lists its __global__ functions (a static KernelList), and
// the first ring to open on a device loads them all beforehand (cudaFuncGetAttributes forces the load).
struct KernelList {
    KernelList(std::initializer_list<const void *> fns);
};
int preload_kernels();   // current device; cheap after the first call per device
void fx_touch_variants();   // fx_path.cu: instantiates its shape table (whose kernels join the list)

// Growable device scratch buffer owned by a handle.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);   // keeps contents only if no growth is needed
    void release();
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

// Pinned host staging buffer.
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

constexpr size_t kDevRoom = 4096;   // head/tail room of sdr_dev_alloc()

// Host -> device copy of caller memory that may be PAGEABLE (the drop-in caller hands over an ordinary Vec<u8>,
// examples/simple_fm.rs:80,153).  cudaMemcpyAsync from pageable memory is staged by the driver on one CPU thread
// (~10 GB/s, no overlap); here it goes through two pinned pieces filled by a small pool of memcpy threads, so the copy
// engine moves piece p while the CPU fills piece p+1.  Pinned / registered sources take one plain cudaMemcpyAsync.
// Like cudaMemcpyAsync from pageable memory, the call returns once `src` may be reused.
struct H2DStager {
    PinBuf piece[2];
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
    int next = 0;
    int copy(void *d_dst, const void *h_src, size_t bytes, cudaStream_t stream);
    void release();
};
bool host_ptr_is_pinned(const void *p);
// memcpy split over the library's pool of copy threads (SDR_STAGE_THREADS, default min(12, 3/4 of the cores))
void parallel_memcpy(void *dst, const void *src, size_t bytes);

// ---- device-side PTX helpers (mbarrier, bulk async copy): ptx_helpers.cuh -------------------------


}  // namespace sdr

#ifdef __CUDACC__
#include "ptx_helpers.cuh"
#endif
