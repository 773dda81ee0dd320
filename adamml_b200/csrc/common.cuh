// Shared helpers for the adamml_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define ADAMML_OK 0
#define ADAMML_ERR_ARG 1
#define ADAMML_ERR_CUDA 2
#define ADAMML_ERR_UNSUPPORTED 3

#define ADAMML_F32 0
#define ADAMML_BF16 1

#define ADAMML_ACT_NONE 0
#define ADAMML_ACT_RELU 1
#define ADAMML_ACT_RELU6 2

void adamml_set_error(const char* fmt, ...);
int adamml_check_launch(const char* what);

#define ADAMML_REQUIRE(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      adamml_set_error(__VA_ARGS__);              \
      return ADAMML_ERR_ARG;                      \
    }                                             \
  } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ADAMML_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ADAMML_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}
// gradient mask evaluated on the stored post-activation value
// (torch threshold_backward / hardtanh_backward use the in-place result).
__device__ __forceinline__ bool act_pass(float out, int act) {
  if (act == ADAMML_ACT_RELU) return out > 0.f;
  if (act == ADAMML_ACT_RELU6) return out > 0.f && out < 6.f;
  return true;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Dispatch a templated launcher on the activation dtype code.
#define ADAMML_DISPATCH_DTYPE(dtype, T, ...)                      \
  do {                                                            \
    if ((dtype) == ADAMML_F32) { typedef float T; __VA_ARGS__; }  \
    else if ((dtype) == ADAMML_BF16) { typedef bf16 T; __VA_ARGS__; } \
    else { adamml_set_error("bad dtype %d", (int)(dtype)); return ADAMML_ERR_ARG; } \
  } while (0)
