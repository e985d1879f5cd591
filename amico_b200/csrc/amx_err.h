// Error plumbing shared by the translation units of libamico_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

namespace amx {
// records the calling thread's message for amx_last_error() and returns `code` (defined in amx_api.cu)
int set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3), visibility("hidden")));
}  // namespace amx

#define AMX_CK(call)                                                                                              \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess)                                                                                    \
            return amx::set_error(-2 /* AMX_E_CUDA */, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                                  __LINE__);                                                                      \
    } while (0)
