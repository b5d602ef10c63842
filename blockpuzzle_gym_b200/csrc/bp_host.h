// bp_host.h -- host-side helpers shared by the translation units of libblockpuzzle_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <string>

// records the calling thread's error message (returned by bp_last_error) and passes `code` through
int bp_fail(int code, const std::string& msg);

#define BP_CU(call)                                                                              \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return bp_fail(BP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));     \
    } while (0)
