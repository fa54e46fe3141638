// Shared host/device helpers for libgenpose_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/genpose_b200.h"

namespace gp {

// ---- host side: error reporting + launch accounting (thread-local, no global mutable state) ----
char *last_error_buf();
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define GP_REQUIRE(cond, ...)             \
    do {                                  \
        if (!(cond)) {                    \
            gp::set_error(__VA_ARGS__);   \
            return GP_ERR_BAD_ARG;        \
        }                                 \
    } while (0)

#define GP_CHECK_LAUNCH(name)                                                         \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            gp::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
            return (int)e__;                                                          \
        }                                                                             \
        gp::count_launch();                                                           \
    } while (0)

#define GP_CUDA(call)                                                                 \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            gp::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
            return (int)e__;                                                          \
        }                                                                             \
    } while (0)

static inline cudaStream_t as_stream(gp_stream_t s) { return (cudaStream_t)s; }
int num_sms();

// ---- device side ----
#ifdef __CUDACC__
// The reference kernels' squared distance as nvcc emits it for sm_100 (verified in the SASS of the
// reference ext): FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.).  Written with explicit intrinsics so
// no compiler version can re-associate it.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ unsigned warp_max_u32(unsigned v) {
    return __reduce_max_sync(0xffffffffu, v);
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// redux.sync with a register member mask (sub-warp groups): the CUDA intrinsic wraps partial masks in a
// branchy helper loop, the instruction itself takes the mask directly
__device__ __forceinline__ unsigned redux_max_u32(unsigned mask, unsigned v) {
    unsigned r;
    asm volatile("redux.sync.max.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(mask));
    return r;
}
#endif

}  // namespace gp
