// Depthwise 3x3 convolution (pad 1, stride 1|2) for NHWC activations: forward, data gradient, weight
// gradient.  Reference call sites: the 3x3 `groups=hidden_dim` convolutions of every InvertedResidual in
// models/sound_mobilenet_v2.py:58 and models/policy_net.py:66,80 (+ their autograd).
//
// These layers carry 0.6 % of the MACs but stream the 6x-expanded MobileNetV2 tensors, so they are pure
// HBM streams: algorithmic bytes = (|x| + |y|) x sizeof(T) forward and data gradient, (|x| + |dy|) for the
// weight gradient.  Every thread owns ONE 16-byte channel vector (8 bf16 / 4 fp32 channels) with its 3x3
// weights in registers and walks a strip of output pixels along W, so each input vector is loaded once
// per strip (L1 serves the 3-row overlap) and all global accesses are full 16-byte vectors.
#include "common.cuh"

namespace {

constexpr int DW_THREADS = 128;
#ifndef ADAMML_DW_FWD_PACKED
#define ADAMML_DW_FWD_PACKED 1
#endif
constexpr bool DW_FWD_PACKED = ADAMML_DW_FWD_PACKED != 0;  // packed FFMA2 in the forward / dgrad stencils (wgrad: always)

// weights are TAP-MAJOR fp32 [9][C] (adamml_pack_weight_dw): a thread's V channels of one tap are one or two
// 16-byte loads, coalesced across the warp
template <int V>
__device__ __forceinline__ void load_w9(const float* __restrict__ w, int C, int c0, float (&wr)[9][V]) {
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const float4 f = *reinterpret_cast<const float4*>(w + (long long)t * C + c0 + i);
      wr[t][i] = f.x; wr[t][i + 1] = f.y; wr[t][i + 2] = f.z; wr[t][i + 3] = f.w;
    }
}

// Fused inference BatchNorm + ReLU/ReLU6 of a depthwise ConvBNReLU (sound_mobilenet_v2.py:58, policy_net.py:66,80): with
// running statistics BN is a per-channel affine map, applied to the accumulators before the store.  ss = [C][2] fp32
// (scale, shift) or nullptr.
template <int V>
__device__ __forceinline__ void bn_act_vec(float (&a)[V], const float* __restrict__ ss, int c0, int act) {
  if (!ss) return;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float2 p = *reinterpret_cast<const float2*>(ss + 2 * (c0 + i));
    a[i] = act_apply(fmaf(a[i], p.x, p.y), act);
  }
}

// acc += a * b over V channels with the packed FFMA2 of sm_100 (fma.rn.f32x2: two fp32 FMAs per issued instruction;
// same IEEE results as two fmaf): the depthwise stencils are issue bound, not HBM bound
template <int V, bool PACKED = true>
__device__ __forceinline__ void fma_vec(float (&acc)[V], const float (&a)[V], const float (&b)[V]) {
  static_assert(V % 2 == 0, "channel vectors hold an even number of channels");
  if constexpr (!PACKED) {  // (A/B switch: measured within 3 % of the packed form on the forward / dgrad stencils)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = fmaf(a[i], b[i], acc[i]);
  } else {
#pragma unroll
  for (int i = 0; i < V; i += 2) {
    const float2 r = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(b[i], b[i + 1]),
                                make_float2(acc[i], acc[i + 1]));
    acc[i] = r.x;
    acc[i + 1] = r.y;
  }
  }
}

// Stride-1 stencil shared by forward (FLIP = false) and data gradient (FLIP = true: 180-degree rotated
// weights, + optional addend).  Strip of SW outputs along W.
template <typename T, bool FLIP, int SW>
__global__ void __launch_bounds__(DW_THREADS)
dw_s1_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y, const T* __restrict__ addend,
             int IMGS, int H, int W, int C, int strips) {
  constexpr int V = VecIO<T>::N;
  const int cvecs = C / V;
  const long long total = (long long)IMGS * H * strips * cvecs;
  const long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= total) return;
  const int cv = (int)(iv % cvecs);
  long long rest = iv / cvecs;
  const int w0 = (int)(rest % strips) * SW;
  rest /= strips;
  const int ho = (int)(rest % H);
  const long long img = rest / H;
  const int c0 = cv * V;
  float wr[9][V];
  load_w9<V>(w, C, c0, wr);
  float acc[SW][V];
#pragma unroll
  for (int j = 0; j < SW; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[j][i] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hi = ho + r - 1;
    if (hi < 0 || hi >= H) continue;
    const T* row = x + ((img * H + hi) * W) * C + c0;
    typename VecIO<T>::raw q[SW + 2];
#pragma unroll
    for (int j = 0; j < SW + 2; ++j) {
      const int wi = w0 + j - 1;
      if (wi >= 0 && wi < W) q[j] = VecIO<T>::load_raw(row + wi * C);
    }
#pragma unroll
    for (int j = 0; j < SW + 2; ++j) {
      const int wi = w0 + j - 1;
      if (wi < 0 || wi >= W) continue;
      float v[V];
      VecIO<T>::unpack(q[j], v);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int o = j - s;  // output index inside the strip fed by this column through tap s
        if (o < 0 || o >= SW) continue;
        const int t = FLIP ? (2 - r) * 3 + (2 - s) : r * 3 + s;
#pragma unroll
        for (int i = 0; i < V; ++i) acc[o][i] = fmaf(v[i], wr[t][i], acc[o][i]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < SW; ++j) {
    const int wo = w0 + j;
    if (wo >= W) continue;
    const long long o = ((img * H + ho) * W + wo) * C + c0;
    if (addend) {
      float a[V];
      VecIO<T>::load(addend + o, a);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[j][i] += a[i];
    }
    VecIO<T>::store(y + o, acc[j]);
  }
}

// Vector policy of the band kernel: 4 channels per thread (8-byte bf16 / 16-byte fp32 accesses) keep the 3x3 fp32
// weights (36 registers) + the rolling window + accumulators under ~100 registers, i.e. >= 4 CTAs per SM.
template <typename T> struct DwVec;
template <> struct DwVec<float> : VecIO<float> {};
template <> struct DwVec<bf16> {
  static constexpr int N = 4;
  typedef uint2 raw;
  __device__ __forceinline__ static raw zero_raw() { return make_uint2(0u, 0u); }
  __device__ __forceinline__ static raw load_raw(const bf16* p) { return *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ static void unpack(const raw& t, float (&v)[4]) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  __device__ __forceinline__ static void load(const bf16* p, float (&v)[4]) { unpack(load_raw(p), v); }
  __device__ __forceinline__ static void store(bf16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

struct X2Raw4 { uint2 hi, lo; };
template <> struct DwVec<x2_t> {
  static constexpr int N = 4;
  typedef X2Raw4 raw;
  __device__ __forceinline__ static raw zero_raw() { return X2Raw4{make_uint2(0u, 0u), make_uint2(0u, 0u)}; }
  __device__ __forceinline__ static raw load_raw(const X2CPtr& p) {
    return X2Raw4{*reinterpret_cast<const uint2*>(p.hi), *reinterpret_cast<const uint2*>(p.lo)};
  }
  __device__ __forceinline__ static void unpack(const raw& t, float (&v)[4]) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.hi.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.hi.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&t.lo.x));
    const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&t.lo.y));
    v[0] = a.x + c.x; v[1] = a.y + c.y; v[2] = b.x + d.x; v[3] = b.y + d.y;
  }
  __device__ __forceinline__ static void load(const X2CPtr& p, float (&v)[4]) { unpack(load_raw(p), v); }
  __device__ __forceinline__ static void store(const X2Ptr& p, const float (&v)[4]) {
    bf16 h[4];
    __half l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x2_split(v[i], h[i], l[i]);
    __nv_bfloat162 h0 = __halves2bfloat162(h[0], h[1]), h1 = __halves2bfloat162(h[2], h[3]);
    __half2 l0 = __halves2half2(l[0], l[1]), l1 = __halves2half2(l[2], l[3]);
    uint2 th, tl;
    th.x = *reinterpret_cast<uint32_t*>(&h0); th.y = *reinterpret_cast<uint32_t*>(&h1);
    tl.x = *reinterpret_cast<uint32_t*>(&l0); tl.y = *reinterpret_cast<uint32_t*>(&l1);
    *reinterpret_cast<uint2*>(p.hi) = th;
    *reinterpret_cast<uint2*>(p.lo) = tl;
  }
};

// Stride-1 stencil over a BAND of RH output rows with a rolling 3-row register window: every input vector is
// loaded once per band (+ 2 halo rows), the 3x3 weights once per thread.  Forward (FLIP = false) and data
// gradient (FLIP = true, + optional addend).
template <typename T, bool FLIP, int SW, int RH, typename CP, typename MP>
__device__ __forceinline__ void
dw_s1_band_body(CP x, const float* __restrict__ w, MP y, CP addend, int IMGS, int H, int W, int C, int strips,
                int bands, const float* __restrict__ ss = nullptr, int act = ADAMML_ACT_NONE,
                LiveLimit live = LiveLimit{nullptr, 0}) {
  typedef DwVec<T> VIO;
  constexpr int V = VIO::N;
  constexpr int NC = SW + 2;
  typedef typename VIO::raw raw_t;
  const int cvecs = C / V;
  const long long total = (long long)IMGS * bands * strips * cvecs;
  const long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= total) return;
  const int cv = (int)(iv % cvecs);
  long long rest = iv / cvecs;
  const int w0 = (int)(rest % strips) * SW;
  rest /= strips;
  const int h0 = (int)(rest % bands) * RH;
  const long long img = rest / bands;
  if (img >= live_count(live, IMGS)) return;  // device-side work limit (inference with skipping)
  const int c0 = cv * V;
  const int h1 = h0 + RH < H ? h0 + RH : H;
  float wr[9][V];
  load_w9<V>(w, C, c0, wr);
  const CP xb = x + ((img * H * W) * C + c0);
  auto load_row = [&](int hi, raw_t (&dst)[NC]) {
    const bool rok = hi >= 0 && hi < H;
    const int rbase = (hi * W + w0 - 1) * C;  // 32-bit offsets inside one image (checked by the launcher)
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int wi = w0 + j - 1;
      dst[j] = (rok && wi >= 0 && wi < W) ? VIO::load_raw(xb + (rbase + j * C)) : VIO::zero_raw();
    }
  };
  auto emit = [&](int ho, const raw_t (&ra)[NC], const raw_t (&rb)[NC], const raw_t (&rc)[NC]) {
    float acc[SW][V];
#pragma unroll
    for (int j = 0; j < SW; ++j)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[j][i] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        float v[V];
        VIO::unpack(r == 0 ? ra[j] : (r == 1 ? rb[j] : rc[j]), v);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int o = j - s;
          if (o < 0 || o >= SW) continue;
          const int t = FLIP ? (2 - r) * 3 + (2 - s) : r * 3 + s;
          fma_vec<V, DW_FWD_PACKED>(acc[o], v, wr[t]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < SW; ++j) {
      const int wo = w0 + j;
      if (wo >= W) continue;
      const long long o = img * ((long long)H * W * C) + ((ho * W + wo) * C + c0);  // (32-bit inside the image)
      if (addend) {
        float a[V];
        VIO::load(addend + o, a);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[j][i] += a[i];
      }
      bn_act_vec<V>(acc[j], ss, c0, act);
      VIO::store(y + o, acc[j]);
    }
  };
  raw_t r0[NC], r1[NC], r2[NC];
  load_row(h0 - 1, r0);
  load_row(h0, r1);
  for (int ho = h0; ho < h1; ho += 3) {
    load_row(ho + 1, r2);
    emit(ho, r0, r1, r2);
    if (ho + 1 < h1) {
      load_row(ho + 2, r0);
      emit(ho + 1, r1, r2, r0);
    }
    if (ho + 2 < h1) {
      load_row(ho + 3, r1);
      emit(ho + 2, r2, r0, r1);
    }
  }
}

template <typename T, bool FLIP, int SW, int RH>
__global__ void __launch_bounds__(DW_THREADS, 4)
dw_s1_band_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                  const T* __restrict__ addend, int IMGS, int H, int W, int C, int strips, int bands,
                  const float* __restrict__ ss, int act, LiveLimit live) {
  dw_s1_band_body<T, FLIP, SW, RH, const T*, T*>(x, w, y, addend, IMGS, H, W, C, strips, bands, ss, act, live);
}
template <int SW, int RH>
__global__ void __launch_bounds__(DW_THREADS, 3)
dw_s1_band_x2_kernel(X2CPtr x, const float* __restrict__ w, X2Ptr y, int IMGS, int H, int W, int C, int strips,
                     int bands, const float* __restrict__ ss, int act, LiveLimit live) {
  dw_s1_band_body<x2_t, false, SW, RH, X2CPtr, X2Ptr>(x, w, y, X2CPtr{nullptr, nullptr}, IMGS, H, W, C, strips, bands,
                                                       ss, act, live);
}

// Stride-2 forward: strip of SW outputs needs 2*SW+1 input columns per row.
template <typename T, int SW, typename CP, typename MP>
__device__ __forceinline__ void
dw_s2_fwd_body(CP x, const float* __restrict__ w, MP y, int IMGS, int H, int W, int C, int Ho, int Wo, int strips,
               const float* __restrict__ ss, int act, LiveLimit live) {
  constexpr int V = VecIO<T>::N;
  const int cvecs = C / V;
  const long long total = (long long)IMGS * Ho * strips * cvecs;
  const long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= total) return;
  const int cv = (int)(iv % cvecs);
  long long rest = iv / cvecs;
  const int w0 = (int)(rest % strips) * SW;
  rest /= strips;
  const int ho = (int)(rest % Ho);
  const long long img = rest / Ho;
  if (img >= live_count(live, IMGS)) return;
  const int c0 = cv * V;
  float wr[9][V];
  load_w9<V>(w, C, c0, wr);
  float acc[SW][V];
#pragma unroll
  for (int j = 0; j < SW; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[j][i] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hi = ho * 2 + r - 1;
    if (hi < 0 || hi >= H) continue;
    const CP row = x + (((img * H + hi) * W) * C + c0);
    typename VecIO<T>::raw q[2 * SW + 1];
#pragma unroll
    for (int j = 0; j < 2 * SW + 1; ++j) {
      const int wi = w0 * 2 + j - 1;
      if (wi >= 0 && wi < W) q[j] = VecIO<T>::load_raw(row + wi * C);
    }
#pragma unroll
    for (int j = 0; j < 2 * SW + 1; ++j) {
      const int wi = w0 * 2 + j - 1;
      if (wi < 0 || wi >= W) continue;
      float v[V];
      VecIO<T>::unpack(q[j], v);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if ((j - s) & 1) continue;
        const int o = (j - s) / 2;
        if (j - s < 0 || o >= SW) continue;
        fma_vec<V, DW_FWD_PACKED>(acc[o], v, wr[r * 3 + s]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < SW; ++j) {
    const int wo = w0 + j;
    if (wo >= Wo) continue;
    bn_act_vec<V>(acc[j], ss, c0, act);
    VecIO<T>::store(y + ((img * Ho + ho) * Wo + wo) * C + c0, acc[j]);
  }
}

template <typename T, int SW>
__global__ void __launch_bounds__(DW_THREADS)
dw_s2_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y, int IMGS, int H, int W,
                 int C, int Ho, int Wo, int strips, const float* __restrict__ ss, int act, LiveLimit live) {
  dw_s2_fwd_body<T, SW, const T*, T*>(x, w, y, IMGS, H, W, C, Ho, Wo, strips, ss, act, live);
}
template <int SW>
__global__ void __launch_bounds__(DW_THREADS)
dw_s2_fwd_x2_kernel(X2CPtr x, const float* __restrict__ w, X2Ptr y, int IMGS, int H, int W, int C, int Ho, int Wo,
                    int strips, const float* __restrict__ ss, int act, LiveLimit live) {
  dw_s2_fwd_body<x2_t, SW, X2CPtr, X2Ptr>(x, w, y, IMGS, H, W, C, Ho, Wo, strips, ss, act, live);
}

// Stride-2 data gradient: one thread per 2x2 input quad (rows 2m, 2m+1; cols 2n, 2n+1) and channel vector;
// the quad needs dy[m..m+1][n..n+1] only.
template <typename T>
__global__ void __launch_bounds__(DW_THREADS)
dw_s2_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx,
                   const T* __restrict__ addend, int IMGS, int H, int W, int C, int Ho, int Wo) {
  constexpr int V = VecIO<T>::N;
  const int cvecs = C / V;
  const int QH = (H + 1) / 2, QW = (W + 1) / 2;
  const long long total = (long long)IMGS * QH * QW * cvecs;
  const long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= total) return;
  const int cv = (int)(iv % cvecs);
  long long rest = iv / cvecs;
  const int n = (int)(rest % QW);
  rest /= QW;
  const int m = (int)(rest % QH);
  const long long img = rest / QH;
  const int c0 = cv * V;
  float wr[9][V];
  load_w9<V>(w, C, c0, wr);
  float g[2][2][V];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const bool ok = (m + a) < Ho && (n + b) < Wo;
      if (ok) VecIO<T>::load(dy + ((img * Ho + m + a) * Wo + n + b) * C + c0, g[a][b]);
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) g[a][b][i] = 0.f;
      }
    }
  // dx[2m+p][2n+q]: row taps r with (2m+p+1-r) even: p=0 -> r=1 (ho=m); p=1 -> r=0 (ho=m+1), r=2 (ho=m)
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int hi = 2 * m + p;
    if (hi >= H) continue;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int wi = 2 * n + q;
      if (wi >= W) continue;
      float acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if ((p + 1 - r) & 1) continue;
        const int a = (p + 1 - r) / 2;  // 0 or 1 (p=1,r=0 -> 1)
        if (p + 1 - r < 0) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          if ((q + 1 - s) & 1) continue;
          if (q + 1 - s < 0) continue;
          const int b = (q + 1 - s) / 2;
          fma_vec<V, DW_FWD_PACKED>(acc, g[a][b], wr[r * 3 + s]);
        }
      }
      const long long o = ((img * H + hi) * W + wi) * C + c0;
      if (addend) {
        float ad[V];
        VecIO<T>::load(addend + o, ad);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += ad[i];
      }
      VecIO<T>::store(dx + o, acc);
    }
  }
}

// Weight gradient: thread = (channel vector, item lane); items = (img, ho, strip of SW outputs) dealt
// round-robin over a persistent grid; 9 x V fp32 accumulators per thread, block reduction through smem,
// one atomic per (block, channel, tap).
template <typename T, int STRIDE, int SW>
__global__ void __launch_bounds__(256, 2)
dw_wgrad_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int IMGS, int H, int W,
                    int C, int Ho, int Wo, int strips, int cpb, int k) {
  constexpr int V = VecIO<T>::N;
  constexpr int NC = STRIDE * SW + (3 - STRIDE);  // input columns per strip row: s1 -> SW+2, s2 -> 2*SW+1
  __shared__ float sh[256][V + 1];
  const int cl = threadIdx.x % cpb, rl = threadIdx.x / cpb;
  const int c0 = (blockIdx.y * cpb + cl) * V;
  const bool ok = c0 < C;
  float acc[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[t][i] = 0.f;
  const long long items = (long long)IMGS * Ho * strips;
  if (ok) {
    for (long long it = (long long)blockIdx.x * k + rl; it < items; it += (long long)gridDim.x * k) {
      const int w0 = (int)(it % strips) * SW;
      const long long rest = it / strips;
      const int ho = (int)(rest % Ho);
      const long long img = rest / Ho;
      float g[SW][V];
#pragma unroll
      for (int j = 0; j < SW; ++j) {
        if (w0 + j < Wo) VecIO<T>::load(dy + ((img * Ho + ho) * Wo + w0 + j) * C + c0, g[j]);
        else {
#pragma unroll
          for (int i = 0; i < V; ++i) g[j][i] = 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hi = ho * STRIDE + r - 1;
        if (hi < 0 || hi >= H) continue;
        const T* row = x + ((img * H + hi) * W) * C + c0;
        typename VecIO<T>::raw q[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const int wi = w0 * STRIDE + j - 1;
          if (wi >= 0 && wi < W) q[j] = VecIO<T>::load_raw(row + wi * C);
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const int wi = w0 * STRIDE + j - 1;
          if (wi < 0 || wi >= W) continue;
          float v[V];
          VecIO<T>::unpack(q[j], v);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            if (j - s < 0 || ((j - s) % STRIDE) != 0) continue;
            const int o = (j - s) / STRIDE;
            if (o >= SW) continue;
            fma_vec<V>(acc[r * 3 + s], g[o], v);
          }
        }
      }
    }
  }
#pragma unroll 1
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int i = 0; i < V; ++i) sh[threadIdx.x][i] = acc[t][i];
    __syncthreads();
    if (rl == 0 && ok) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float s = 0.f;
        for (int y = 0; y < k; ++y) s += sh[y * cpb + cl][i];
        atomicAdd(&dw[(long long)t * C + c0 + i], s);
      }
    }
    __syncthreads();
  }
}

// Stride-1 weight gradient with the same rolling-window scheme: a thread owns 4 channels (36 fp32 accumulators),
// walks (image, row band, column strip) items of a persistent grid, and per step streams two new x rows and two
// dy rows (12 independent loads in flight).
template <typename T, int SW, int RH>
__global__ void __launch_bounds__(256, 2)
dw_wgrad_band_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int IMGS, int H, int W,
                     int C, int strips, int bands, int cpb, int k) {
  typedef DwVec<T> VIO;
  constexpr int V = VIO::N;
  constexpr int NC = SW + 2;
  typedef typename VIO::raw raw_t;
  __shared__ float sh[256][V + 1];
  const int cl = threadIdx.x % cpb, rl = threadIdx.x / cpb;
  const int c0 = (blockIdx.y * cpb + cl) * V;
  const bool ok = c0 < C;
  float acc[9][V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[t][i] = 0.f;
  const long long items = (long long)IMGS * bands * strips;
  if (ok) {
    for (long long it = (long long)blockIdx.x * k + rl; it < items; it += (long long)gridDim.x * k) {
      const int w0 = (int)(it % strips) * SW;
      const long long rest = it / strips;
      const int h0 = (int)(rest % bands) * RH;
      const long long img = rest / bands;
      const int h1 = h0 + RH < H ? h0 + RH : H;
      const T* xb = x + (img * H * W) * C + c0;
      const T* gb = dy + (img * H * W) * C + c0;
      auto load_x = [&](int hi, raw_t (&dst)[NC]) {
        const bool rok = hi >= 0 && hi < H;
        const int rbase = (hi * W + w0 - 1) * C;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const int wi = w0 + j - 1;
          dst[j] = (rok && wi >= 0 && wi < W) ? VIO::load_raw(xb + (rbase + j * C)) : VIO::zero_raw();
        }
      };
      auto load_g = [&](int ho, raw_t (&dst)[SW]) {
        const bool rok = ho < h1;
        const int rbase = (ho * W + w0) * C;
#pragma unroll
        for (int j = 0; j < SW; ++j)
          dst[j] = (rok && w0 + j < W) ? VIO::load_raw(gb + (rbase + j * C)) : VIO::zero_raw();
      };
      auto emit = [&](const raw_t (&g)[SW], const raw_t (&ra)[NC], const raw_t (&rb)[NC], const raw_t (&rc)[NC]) {
        float gv[SW][V];
#pragma unroll
        for (int j = 0; j < SW; ++j) VIO::unpack(g[j], gv[j]);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            float v[V];
            VIO::unpack(r == 0 ? ra[j] : (r == 1 ? rb[j] : rc[j]), v);
#pragma unroll
            for (int s_ = 0; s_ < 3; ++s_) {
              const int o = j - s_;
              if (o < 0 || o >= SW) continue;
              fma_vec<V>(acc[r * 3 + s_], gv[o], v);
            }
          }
        }
      };
      raw_t r0[NC], r1[NC], r2[NC], r3[NC], g0[SW], g1[SW];
      load_x(h0 - 1, r0);
      load_x(h0, r1);
      for (int ho = h0; ho < h1; ho += 4) {
        load_x(ho + 1, r2);
        load_x(ho + 2, r3);
        load_g(ho, g0);
        load_g(ho + 1, g1);
        emit(g0, r0, r1, r2);
        emit(g1, r1, r2, r3);  // rows past the band contribute zeros (g1 = 0)
        if (ho + 2 < h1) {
          load_x(ho + 3, r0);
          load_x(ho + 4, r1);
          load_g(ho + 2, g0);
          load_g(ho + 3, g1);
          emit(g0, r2, r3, r0);
          emit(g1, r3, r0, r1);
        }
      }
    }
  }
#pragma unroll 1
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int i = 0; i < V; ++i) sh[threadIdx.x][i] = acc[t][i];
    __syncthreads();
    if (rl == 0 && ok) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float s_ = 0.f;
        for (int y = 0; y < k; ++y) s_ += sh[y * cpb + cl][i];
        atomicAdd(&dw[(long long)t * C + c0 + i], s_);
      }
    }
    __syncthreads();
  }
}

template <typename T>
inline bool dw_vec_ok(int C, const void* a, const void* b, const void* c = nullptr, const void* d = nullptr) {
  if (C % VecIO<T>::N) return false;
  const void* ps[4] = {a, b, c, d};
  for (int i = 0; i < 4; ++i)
    if (ps[i] && ((uintptr_t)ps[i] % 16)) return false;
  return true;
}

inline unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace

extern "C" {

static int dwconv_fwd_impl(const void* x, const float* w, void* y, int IMGS, int H, int W, int C, int stride, int Ho,
                           int Wo, int dtype, const float* ss, int act, cudaStream_t stream) {
  ADAMML_REQUIRE(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2");
  ADAMML_REQUIRE((long long)H * W * C < (1LL << 31), "dwconv: one image must stay below 2^31 elements");
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / stride + 1 && Wo == (W + 2 - 3) / stride + 1, "dwconv: bad Ho/Wo");
  const LiveLimit live = ss ? adamml_live_limit(IMGS) : LiveLimit{nullptr, 0};  // inference launches only
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (!dw_vec_ok<T>(C, x, y)) {
      ADAMML_REQUIRE(!ss, "dwconv: the fused BatchNorm epilogue needs a vectorisable channel count");
      return adamml_dwconv_fwd_scalar(x, w, y, IMGS, H, W, C, stride, Ho, Wo, dtype, stream);
    }
    const int cvecs = C / VecIO<T>::N;
    if (stride == 1) {
      constexpr int SW = 2, RH = 16;
      const int strips = (W + SW - 1) / SW, bands = (H + RH - 1) / RH;
      const long long total = (long long)IMGS * bands * strips * (C / DwVec<T>::N);
      dw_s1_band_kernel<T, false, SW, RH><<<blocks_for(total, DW_THREADS), DW_THREADS, 0, stream>>>(
          (const T*)x, w, (T*)y, nullptr, IMGS, H, W, C, strips, bands, ss, act, live);
    } else {
      constexpr int SW = 2;
      const int strips = (Wo + SW - 1) / SW;
      const long long total = (long long)IMGS * Ho * strips * cvecs;
      dw_s2_fwd_kernel<T, SW><<<blocks_for(total, DW_THREADS), DW_THREADS, 0, stream>>>(
          (const T*)x, w, (T*)y, IMGS, H, W, C, Ho, Wo, strips, ss, act, live);
    }
  });
  return adamml_check_launch("dwconv_fwd");
}

int adamml_dwconv_fwd(const void* x, const float* w, void* y, int IMGS, int H, int W, int C, int stride, int Ho,
                      int Wo, int dtype, cudaStream_t stream) {
  return dwconv_fwd_impl(x, w, y, IMGS, H, W, C, stride, Ho, Wo, dtype, nullptr, ADAMML_ACT_NONE, stream);
}

/* inference-mode depthwise conv + BatchNorm + ReLU/ReLU6 in one pass: y = act(dwconv(x) * scale[c] + shift[c]) */
int adamml_dwconv_bn_act_fwd(const void* x, const float* w, void* y, int IMGS, int H, int W, int C, int stride, int Ho,
                             int Wo, const float* scale_shift, int act, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(scale_shift, "dwconv_bn_act: needs the folded BatchNorm scale / shift");
  return dwconv_fwd_impl(x, w, y, IMGS, H, W, C, stride, Ho, Wo, dtype, scale_shift, act, stream);
}

/* x2 planes (forward pass of the default precision mode); C % 8 == 0 */
static int dwconv_fwd_x2_impl(const void* x_hi, const void* x_lo, const float* w, void* y_hi, void* y_lo, int IMGS,
                              int H, int W, int C, int stride, int Ho, int Wo, const float* ss, int act,
                              cudaStream_t stream) {
  ADAMML_REQUIRE(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2");
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / stride + 1 && Wo == (W + 2 - 3) / stride + 1, "dwconv: bad Ho/Wo");
  ADAMML_REQUIRE(dw_vec_ok<bf16>(C, x_hi, x_lo, y_hi, y_lo), "dwconv_fwd_x2: needs C %% 8 == 0 and aligned planes");
  const LiveLimit live = ss ? adamml_live_limit(IMGS) : LiveLimit{nullptr, 0};
  if (stride == 1) {
    constexpr int SW = 2, RH = 16;
    const int strips = (W + SW - 1) / SW, bands = (H + RH - 1) / RH;
    const long long total = (long long)IMGS * bands * strips * (C / 4);
    dw_s1_band_x2_kernel<SW, RH><<<blocks_for(total, DW_THREADS), DW_THREADS, 0, stream>>>(
        x2c(x_hi, x_lo), w, x2m(y_hi, y_lo), IMGS, H, W, C, strips, bands, ss, act, live);
  } else {
    constexpr int SW = 2;
    const int strips = (Wo + SW - 1) / SW;
    const long long total = (long long)IMGS * Ho * strips * (C / 8);
    dw_s2_fwd_x2_kernel<SW><<<blocks_for(total, DW_THREADS), DW_THREADS, 0, stream>>>(
        x2c(x_hi, x_lo), w, x2m(y_hi, y_lo), IMGS, H, W, C, Ho, Wo, strips, ss, act, live);
  }
  return adamml_check_launch("dwconv_fwd_x2");
}

int adamml_dwconv_fwd_x2(const void* x_hi, const void* x_lo, const float* w, void* y_hi, void* y_lo, int IMGS, int H,
                         int W, int C, int stride, int Ho, int Wo, cudaStream_t stream) {
  return dwconv_fwd_x2_impl(x_hi, x_lo, w, y_hi, y_lo, IMGS, H, W, C, stride, Ho, Wo, nullptr, ADAMML_ACT_NONE, stream);
}

int adamml_dwconv_bn_act_fwd_x2(const void* x_hi, const void* x_lo, const float* w, void* y_hi, void* y_lo, int IMGS,
                                int H, int W, int C, int stride, int Ho, int Wo, const float* scale_shift, int act,
                                cudaStream_t stream) {
  ADAMML_REQUIRE(scale_shift, "dwconv_bn_act_x2: needs the folded BatchNorm scale / shift");
  return dwconv_fwd_x2_impl(x_hi, x_lo, w, y_hi, y_lo, IMGS, H, W, C, stride, Ho, Wo, scale_shift, act, stream);
}

int adamml_dwconv_dgrad(const void* dy, const float* w, void* dx, const void* addend, int IMGS, int H, int W, int C,
                        int stride, int Ho, int Wo, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2");
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / stride + 1 && Wo == (W + 2 - 3) / stride + 1, "dwconv: bad Ho/Wo");
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (!dw_vec_ok<T>(C, dy, dx, addend))
      return adamml_dwconv_dgrad_scalar(dy, w, dx, addend, IMGS, H, W, C, stride, Ho, Wo, dtype, stream);
    const int cvecs = C / VecIO<T>::N;
    if (stride == 1) {
      constexpr int SW = 2, RH = 16;
      const int strips = (W + SW - 1) / SW, bands = (H + RH - 1) / RH;
      const long long total = (long long)IMGS * bands * strips * (C / DwVec<T>::N);
      dw_s1_band_kernel<T, true, SW, RH><<<blocks_for(total, DW_THREADS), DW_THREADS, 0, stream>>>(
          (const T*)dy, w, (T*)dx, (const T*)addend, IMGS, H, W, C, strips, bands, nullptr, ADAMML_ACT_NONE,
          LiveLimit{nullptr, 0});
    } else {
      const long long total = (long long)IMGS * ((H + 1) / 2) * ((W + 1) / 2) * cvecs;
      dw_s2_dgrad_kernel<T><<<blocks_for(total, DW_THREADS), DW_THREADS, 0, stream>>>(
          (const T*)dy, w, (T*)dx, (const T*)addend, IMGS, H, W, C, Ho, Wo);
    }
  });
  return adamml_check_launch("dwconv_dgrad");
}

int adamml_dwconv_wgrad(const void* x, const void* dy, float* dw, int IMGS, int H, int W, int C, int stride, int Ho,
                        int Wo, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2");
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / stride + 1 && Wo == (W + 2 - 3) / stride + 1, "dwconv: bad Ho/Wo");
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (!dw_vec_ok<T>(C, x, dy))
      return adamml_dwconv_wgrad_scalar(x, dy, dw, IMGS, H, W, C, stride, Ho, Wo, dtype, stream);
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * 9, stream);
    if (stride == 1) {
      constexpr int SW = 2, RH = 16;
      const int cv4 = C / DwVec<T>::N;
      const int cch = (cv4 + 255) / 256;
      const int cpb4 = (cv4 + cch - 1) / cch;
      const int k4 = 256 / cpb4;
      const int strips = (W + SW - 1) / SW, bands = (H + RH - 1) / RH;
      const long long items = (long long)IMGS * bands * strips;
      long long gx = (items + k4 - 1) / k4;
      const long long cap = 148LL * 2 * 4 / cch > 0 ? 148LL * 2 * 4 / cch : 1;
      if (gx > cap) gx = cap;
      dim3 grid((unsigned)gx, cch);
      dw_wgrad_band_kernel<T, SW, RH><<<grid, cpb4 * k4, 0, stream>>>((const T*)x, (const T*)dy, dw, IMGS, H, W, C,
                                                                       strips, bands, cpb4, k4);
      return adamml_check_launch("dwconv_wgrad");
    }
    const int cvecs = C / VecIO<T>::N;
    const int cchunks = (cvecs + 255) / 256;
    const int cpb = (cvecs + cchunks - 1) / cchunks;
    const int k = 256 / cpb;
    constexpr int SW1 = 2, SW2 = 2;
    const int strips = stride == 1 ? (Wo + SW1 - 1) / SW1 : (Wo + SW2 - 1) / SW2;
    const long long items = (long long)IMGS * Ho * strips;
    long long gx = (items + k - 1) / k;
    const long long cap = 148LL * 8 / cchunks > 0 ? 148LL * 8 / cchunks : 1;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, cchunks);
    if (stride == 1)
      dw_wgrad_vec_kernel<T, 1, SW1><<<grid, cpb * k, 0, stream>>>((const T*)x, (const T*)dy, dw, IMGS, H, W, C, Ho, Wo,
                                                                   strips, cpb, k);
    else
      dw_wgrad_vec_kernel<T, 2, SW2><<<grid, cpb * k, 0, stream>>>((const T*)x, (const T*)dy, dw, IMGS, H, W, C, Ho, Wo,
                                                                   strips, cpb, k);
  });
  return adamml_check_launch("dwconv_wgrad");
}

}  // extern "C"
