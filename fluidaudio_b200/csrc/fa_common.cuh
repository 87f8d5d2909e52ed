// Shared definitions for the fluidaudio_b200 CUDA library (sm_100a only).
#pragma once

#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define FA_HD __host__ __device__ __forceinline__
#else
#define FA_HD inline
#endif

// Status codes of the C ABI (include/fluidaudio_b200.h).  Values 0..5 and 255 coincide with
// fastcluster_wrapper_status (reference: Sources/FastClusterWrapper/include/FastClusterWrapper.h:11-19).
enum : int {
    FA_OK = 0,
    FA_INVALID_ARGUMENT = 1,
    FA_INDEX_OVERFLOW = 2,
    FA_OUTPUT_TOO_SMALL = 3,
    FA_ALLOCATION_FAILURE = 4,
    FA_RUNTIME_ERROR = 5,
    FA_NO_DEVICE = 6,
    FA_CUDA_ERROR = 7,
    FA_UNSUPPORTED = 8,
    FA_UNKNOWN_ERROR = 255,
};

namespace fa {

// thread-local last-error text, set by the C ABI on every failure
void set_error(const char *fmt, ...);
const char *last_error();

} // namespace fa
