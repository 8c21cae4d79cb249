// rtc.h — run-time specialisation of the kernel templates in fir_fast.cuh (NVRTC -> cubin -> driver module).
//
// libnvrtc and libcuda are dlopen()ed on first use (like NCCL in chan.cu): the shared library itself still links only
// the static CUDA runtime.  No NVRTC on the machine is not an error of the product path — the caller keeps its
// pre-compiled / generic kernel — but it is reported (sdr_fmrx_kernel_kind(), sdr_last_error()).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace sdr {

struct RtcModule {
    std::vector<void *> modules;   // CUmodule per instantiation (one cubin each)
    std::vector<void *> fns;       // CUfunction per requested name expression, in request order
};

// Compile fir_fast.cuh for sm_100a, one cubin per template instantiation (NVRTC name expressions, e.g.
// "sdr::k_fir_fast<201,64,2,64,8,0>"); cubins already in the on-disk cache are reused, the others are compiled
// concurrently.  Needs libnvrtc only — no GPU, no driver.  SDR_OK, or a negative SDR_E_* code (sdr_last_error()).
int rtc_compile_cubins(const std::vector<std::string> &name_exprs, std::vector<std::vector<char>> *cubins,
                       std::vector<std::string> *lowered, int *n_compiled);

// Compile (or fetch from the in-process cache, keyed by `key` and device), load into the device's primary context
// and raise each function's dynamic shared-memory limit to max_dyn_smem.  SDR_OK or a negative code (sdr_last_error()).
int rtc_get_module(int device, const std::string &key, const std::vector<std::string> &name_exprs, int max_dyn_smem,
                   const RtcModule **out);

// cuLaunchKernel of a 1-D grid; params = array of pointers to the kernel's arguments.
int rtc_launch(void *fn, unsigned grid, unsigned block, unsigned dyn_smem, cudaStream_t stream, void **params);

}  // namespace sdr
