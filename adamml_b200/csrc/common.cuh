// Shared helpers for the adamml_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define ADAMML_OK 0
#define ADAMML_ERR_ARG 1
#define ADAMML_ERR_CUDA 2
#define ADAMML_ERR_UNSUPPORTED 3

#define ADAMML_F32 0
#define ADAMML_BF16 1
// "x2" activations: TWO planes per tensor, hi = bf16(v) and lo = fp16(v - hi)  (about 20 mantissa bits, bf16 range).
// The forward pass of the default precision mode stores every activation this way and multiplies on the tensor
// cores as hi*hi + hi*lo + lo*hi + lo*lo with fp32 accumulation; the backward pass reads the hi planes as bf16.
#define ADAMML_X2 2

#define ADAMML_ACT_NONE 0
#define ADAMML_ACT_RELU 1
#define ADAMML_ACT_RELU6 2

void adamml_set_error(const char* fmt, ...);
int adamml_check_launch(const char* what);
// conv_simt.cu: scalar depthwise fallbacks used by dwconv.cu for ragged channel counts
int adamml_dwconv_fwd_scalar(const void* x, const float* w, void* y, int IMGS, int H, int W, int C, int stride, int Ho,
                             int Wo, int dtype, cudaStream_t stream);
int adamml_dwconv_dgrad_scalar(const void* dy, const float* w, void* dx, const void* addend, int IMGS, int H, int W,
                               int C, int stride, int Ho, int Wo, int dtype, cudaStream_t stream);
int adamml_dwconv_wgrad_scalar(const void* x, const void* dy, float* dw, int IMGS, int H, int W, int C, int stride,
                               int Ho, int Wo, int dtype, cudaStream_t stream);

// Device-side work limit (inference with decision-driven skipping, models/adamml.py:81-86): the batch buffers keep
// their static capacity, but only the first *n clips hold selected (segment, video) pairs; `unit` = images (or GEMM
// rows) per clip at this layer.  Kernels read *n on the device and skip every tile / thread beyond it, so the host
// never learns the count (no D2H copy, no data-dependent launch shape -> the pass is CUDA-graph capturable).
struct LiveLimit {
  const int* n;  // nullptr: everything is live
  int unit;
};
// set by adamml_set_live_clips (api.cu, thread local); imgs = static capacity of the op (images or rows)
LiveLimit adamml_live_limit(long long capacity_items);
__device__ __forceinline__ long long live_count(const LiveLimit& l, long long cap) {
  if (!l.n) return cap;
  const long long v = (long long)(*l.n) * l.unit;
  return v < cap ? v : cap;
}

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

// ---- x2 planes ---------------------------------------------------------------------------------
struct x2_t {};  // tag type of the two-plane activation format
__device__ __forceinline__ void x2_split(float v, bf16& hi, __half& lo) {
  hi = __float2bfloat16_rn(v);
  const float r = v - __bfloat162float(hi);
  lo = __float2half_rn(r == r ? r : 0.f);  // inf - inf: keep the non-finite value in hi only
}
__device__ __forceinline__ float x2_join(bf16 hi, __half lo) { return __bfloat162float(hi) + __half2float(lo); }
struct X2CPtr {
  const bf16* hi;
  const __half* lo;
  __host__ __device__ __forceinline__ X2CPtr operator+(long long i) const { return X2CPtr{hi + i, lo + i}; }
  __host__ __device__ __forceinline__ explicit operator bool() const { return hi != nullptr; }
};
struct X2Ptr {
  bf16* hi;
  __half* lo;
  __host__ __device__ __forceinline__ X2Ptr operator+(long long i) const { return X2Ptr{hi + i, lo + i}; }
  __host__ __device__ __forceinline__ explicit operator bool() const { return hi != nullptr; }
  __host__ __device__ __forceinline__ operator X2CPtr() const { return X2CPtr{hi, lo}; }
};
static inline X2CPtr x2c(const void* hi, const void* lo) { return X2CPtr{(const bf16*)hi, (const __half*)lo}; }
static inline X2Ptr x2m(void* hi, void* lo) { return X2Ptr{(bf16*)hi, (__half*)lo}; }
// scalar element access through either a plain pointer or a plane pair
__device__ __forceinline__ float ld_elem(const float* p, long long i) { return p[i]; }
__device__ __forceinline__ float ld_elem(const bf16* p, long long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ float ld_elem(const X2CPtr& p, long long i) { return x2_join(p.hi[i], p.lo[i]); }
__device__ __forceinline__ void st_elem(float* p, long long i, float v) { p[i] = v; }
__device__ __forceinline__ void st_elem(bf16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st_elem(const X2Ptr& p, long long i, float v) { x2_split(v, p.hi[i], p.lo[i]); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Dispatch a templated launcher on the activation dtype code.
#define ADAMML_DISPATCH_DTYPE(dtype, T, ...)                      \
  do {                                                            \
    if ((dtype) == ADAMML_F32) { typedef float T; __VA_ARGS__; }  \
    else if ((dtype) == ADAMML_BF16) { typedef bf16 T; __VA_ARGS__; } \
    else { adamml_set_error("bad dtype %d", (int)(dtype)); return ADAMML_ERR_ARG; } \
  } while (0)

// ---- 16-byte vector I/O: VEC = 4 fp32 or 8 bf16 elements --------------------------------------
template <typename T> struct VecIO;
template <> struct VecIO<float> {
  static constexpr int N = 4;
  typedef float4 raw;
  __device__ __forceinline__ static raw zero_raw() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ static raw load_raw(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ static void unpack(const raw& t, float (&v)[4]) {
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ static void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct VecIO<bf16> {
  static constexpr int N = 8;
  typedef uint4 raw;
  __device__ __forceinline__ static raw zero_raw() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ static raw load_raw(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ static void unpack(const raw& t, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ static void load(const bf16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ static void store(bf16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// x2: 8 channels per thread = one 16-byte vector of each plane
struct X2Raw { uint4 hi, lo; };
template <> struct VecIO<x2_t> {
  static constexpr int N = 8;
  typedef X2Raw raw;
  __device__ __forceinline__ static raw zero_raw() { return X2Raw{make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)}; }
  __device__ __forceinline__ static raw load_raw(const X2CPtr& p) {
    return X2Raw{*reinterpret_cast<const uint4*>(p.hi), *reinterpret_cast<const uint4*>(p.lo)};
  }
  __device__ __forceinline__ static void unpack(const raw& t, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t.hi);
    const __half2* l = reinterpret_cast<const __half2*>(&t.lo);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      const float2 g = __half22float2(l[i]);
      v[2 * i] = f.x + g.x; v[2 * i + 1] = f.y + g.y;
    }
  }
  __device__ __forceinline__ static void load(const X2CPtr& p, float (&v)[8]) { unpack(load_raw(p), v); }
  __device__ __forceinline__ static void store(const X2Ptr& p, const float (&v)[8]) {
    uint4 th, tl;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&th);
    __half2* l = reinterpret_cast<__half2*>(&tl);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bf16 h0, h1;
      __half l0, l1;
      x2_split(v[2 * i], h0, l0);
      x2_split(v[2 * i + 1], h1, l1);
      h[i] = __halves2bfloat162(h0, h1);
      l[i] = __halves2half2(l0, l1);
    }
    *reinterpret_cast<uint4*>(p.hi) = th;
    *reinterpret_cast<uint4*>(p.lo) = tl;
  }
};
// pointer types of an activation tensor of element type T
template <typename T> struct PtrOf { typedef const T* c; typedef T* m; };
template <> struct PtrOf<x2_t> { typedef X2CPtr c; typedef X2Ptr m; };
