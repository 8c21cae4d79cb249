// common.cuh — shared host/device helpers for libsdr_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

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

// ---- device-side PTX helpers (mbarrier, bulk async copy): ptx_helpers.cuh -------------------------


}  // namespace sdr

#ifdef __CUDACC__
#include "ptx_helpers.cuh"
#endif
