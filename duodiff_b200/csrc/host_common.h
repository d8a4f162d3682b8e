// Host-side helpers shared by the translation units of libduodiff_b200.so (hidden visibility: not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <utility>

#include "../../include/duodiff_b200.h"

#define DDB_HIDDEN __attribute__((visibility("hidden")))

namespace ddb_host {
DDB_HIDDEN int fail_msg(int code, const char* msg);  // stores the thread-local ddb_last_error() text, returns code
DDB_HIDDEN void count_launch();                      // ddb_launch_count()
DDB_HIDDEN int use_pdl();
DDB_HIDDEN bool ae_set_option(const char* name, int value);  // autoencoder.cu's share of ddb_set_option
DDB_HIDDEN int sm100_device(int* num_sms);           // DDB_ERR_CUDA unless the current device is sm_100
// cuTensorMapEncodeTiled for a bf16 tensor of `rank` dims (dims[0] contiguous; strides[i] = byte stride of dim i+1),
// 128-byte swizzle, zero fill outside the tensor
DDB_HIDDEN int encode_bf16_sw128(CUtensorMap* tm, const void* base, int rank, const unsigned long long* dims,
                                 const unsigned long long* strides, const unsigned* box);

// cudaFuncSetAttribute (dynamic shared memory size) is a PER-DEVICE setting: a kernel is configured once per device
// ordinal, not once per process.  Usage: static DeviceOnce once; if (!once.done()) { set attribute; once.mark(); }
struct DeviceOnce {
    std::atomic<unsigned long long> mask[4] = {};  // device ordinals 0..255
    static int ordinal() {
        int dev = 0;
        cudaGetDevice(&dev);
        return dev & 255;
    }
    bool done() const {
        const int d = ordinal();
        return (mask[d >> 6].load(std::memory_order_acquire) >> (d & 63)) & 1ull;
    }
    void mark() {
        const int d = ordinal();
        mask[d >> 6].fetch_or(1ull << (d & 63), std::memory_order_release);
    }
};

inline int failf(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return fail_msg(code, buf);
}

// launch with programmaticStreamSerialization (see csrc/ptx.cuh: every kernel calls pdl_wait() before touching memory)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kfn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = use_pdl() ? 1 : 0;
    cfg.attrs = at, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kfn, std::forward<Args>(args)...);
}
}  // namespace ddb_host
