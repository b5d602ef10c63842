// bp_host.h -- host-side helpers shared by the translation units of libblockpuzzle_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

// records the calling thread's error message (returned by bp_last_error) and passes `code` through
int bp_fail(int code, const std::string& msg);

// bp_replay.cu: the lane-cooperative fast path of bp_her_relabel (rows of 1, 2, 4 or 8 float4 chunks, 16-byte aligned
// tensors); BP_ERR_NOT_IMPLEMENTED when the shape does not qualify -- the caller then uses the generic sampler kernel
int bp_her_relabel_coop(const float* d_ep_ag, const float* d_ep_g, int32_t B, int32_t T, int32_t dimg, int64_t n, float future_p,
                        uint64_t seed, int64_t index_offset, int32_t* d_ep_idx, int32_t* d_t, int32_t* d_future_t, float* d_ag2,
                        float* d_g, float* d_r, void* stream);

#define BP_CU(call)                                                                              \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return bp_fail(BP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));     \
    } while (0)
