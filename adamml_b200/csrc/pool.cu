// Pooling kernels for NHWC activations.
//  - 3x3/s2/p1 spatial max-pool of the ResNet stem (reference models/resnet.py:141,202)
//  - temporal max/avg pool k3 s2 p1 over the frames of one video (models/common.py:4-33)
//  - global average pool (models/resnet.py:212, sound_mobilenet_v2.py:156, policy_net.py:146)
// Max-pool backward routes the gradient to the FIRST maximum in scan order, which is what
// ATen's max_pool2d / max_pool3d kernels do (strict '>' while scanning).
#include "common.cuh"

namespace {

inline int ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  long long cap = 148LL * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int IMGS, int H, int W, int C, int Ho,
                                   int Wo) {
  long long total = (long long)IMGS * Ho * Wo * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    long long pix = idx / C;
    int wo = (int)(pix % Wo);
    int ho = (int)((pix / Wo) % Ho);
    int img = (int)(pix / ((long long)Wo * Ho));
    float best = -INFINITY;
    for (int r = 0; r < 3; ++r) {
      int hi = ho * 2 + r - 1;
      if (hi < 0 || hi >= H) continue;
      for (int s = 0; s < 3; ++s) {
        int wi = wo * 2 + s - 1;
        if (wi < 0 || wi >= W) continue;
        float v = to_f32(x[(((long long)img * H + hi) * W + wi) * C + c]);
        if (v > best || v != v) best = v;
      }
    }
    y[idx] = from_f32<T>(best);
  }
}

// gather form: each input element looks at the <=4 windows that contain it and recomputes
// that window's first-max position.
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int IMGS,
                                   int H, int W, int C, int Ho, int Wo) {
  long long total = (long long)IMGS * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    long long pix = idx / C;
    int wi = (int)(pix % W);
    int hi = (int)((pix / W) % H);
    int img = (int)(pix / ((long long)W * H));
    float acc = 0.f;
    // windows ho with ho*2-1 <= hi <= ho*2+1
    int ho_lo = (hi - 1 + 1) / 2;  // ceil((hi-1)/2) for hi>=0
    int ho_hi = (hi + 1) / 2;
    int wo_lo = (wi - 1 + 1) / 2;
    int wo_hi = (wi + 1) / 2;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      if (ho < 0 || ho >= Ho) continue;
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        if (wo < 0 || wo >= Wo) continue;
        float best = -INFINITY;
        int bh = -1, bw = -1;
        for (int r = 0; r < 3; ++r) {
          int h2 = ho * 2 + r - 1;
          if (h2 < 0 || h2 >= H) continue;
          for (int s = 0; s < 3; ++s) {
            int w2 = wo * 2 + s - 1;
            if (w2 < 0 || w2 >= W) continue;
            float v = to_f32(x[(((long long)img * H + h2) * W + w2) * C + c]);
            if (v > best || v != v) { best = v; bh = h2; bw = w2; }
          }
        }
        if (bh == hi && bw == wi) acc += to_f32(dy[(((long long)img * Ho + ho) * Wo + wo) * C + c]);
      }
    }
    dx[idx] = from_f32<T>(acc);
  }
}

// x: [V, T, E] -> y: [V, To, E], To = (T + 2 - 3)/2 + 1, window {2t-1, 2t, 2t+1}
template <typename T>
__global__ void tpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long V, int Tn, int To, long long E,
                                 int mode_avg) {
  long long total = V * To * E;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long e = idx % E;
    int to = (int)((idx / E) % To);
    long long v = idx / (E * To);
    float best = -INFINITY, sum = 0.f;
    for (int k = 0; k < 3; ++k) {
      int t = to * 2 + k - 1;
      if (t < 0 || t >= Tn) continue;
      float val = to_f32(x[(v * Tn + t) * E + e]);
      sum += val;
      if (val > best || val != val) best = val;
    }
    // AvgPool3d default count_include_pad=True: divisor is always 3
    y[idx] = from_f32<T>(mode_avg ? sum / 3.f : best);
  }
}

template <typename T>
__global__ void tpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long long V,
                                 int Tn, int To, long long E, int mode_avg) {
  long long total = V * Tn * E;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long e = idx % E;
    int t = (int)((idx / E) % Tn);
    long long v = idx / (E * Tn);
    float acc = 0.f;
    int lo = t / 2;         // ceil((t-1)/2)
    int hi = (t + 1) / 2;
    for (int to = lo; to <= hi; ++to) {
      if (to < 0 || to >= To) continue;
      float g = to_f32(dy[(v * To + to) * E + e]);
      if (mode_avg) { acc += g / 3.f; continue; }
      float best = -INFINITY;
      int bt = -1;
      for (int k = 0; k < 3; ++k) {
        int t2 = to * 2 + k - 1;
        if (t2 < 0 || t2 >= Tn) continue;
        float val = to_f32(x[(v * Tn + t2) * E + e]);
        if (val > best || val != val) { best = val; bt = t2; }
      }
      if (bt == t) acc += g;
    }
    dx[idx] = from_f32<T>(acc);
  }
}

// ---- 16-byte vectorised versions (C % VEC == 0): thread = (pixel, channel vector) ----------------------
// 3x3/s2/p1 max-pool; optionally records the window position (r*3+s, first maximum in scan order) of
// every output element so that the backward pass is a pure gather that never re-reads x.
// PreBN (training stem, resnet.py:197-200): x is the PRE-BatchNorm tensor z and every window element goes through
// act(z * scale + shift) (per BN group = img / imgs_per_group and channel) on its way into the maximum, so the
// full-resolution post-activation tensor -- which nothing but this pool reads -- is never written.
struct PreBN {
  const float* ss;  // [G][C][2] (scale, shift); nullptr = plain max-pool
  int imgs_per_group;
  int act;
};
template <typename T, typename CP, typename MP>
__device__ __forceinline__ void
maxpool_fwd_vec_body(CP x, MP y, unsigned char* __restrict__ pos, int IMGS, int H, int W, int C, int Ho, int Wo,
                     LiveLimit live, PreBN pre = PreBN{nullptr, 1, ADAMML_ACT_NONE}) {
  constexpr int V = VecIO<T>::N;
  const int cvecs = C / V;
  const long long total = (long long)IMGS * Ho * Wo * cvecs;
  const long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= total) return;
  const int cv = (int)(iv % cvecs);
  const long long pix = iv / cvecs;
  const int wo = (int)(pix % Wo);
  const int ho = (int)((pix / Wo) % Ho);
  const long long img = pix / ((long long)Wo * Ho);
  if (img >= live_count(live, IMGS)) return;  // device-side work limit (inference with skipping)
  typename VecIO<T>::raw q[9];
  bool ok[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hi = ho * 2 + r - 1;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int wi = wo * 2 + s - 1;
      ok[r * 3 + s] = hi >= 0 && hi < H && wi >= 0 && wi < W;
      if (ok[r * 3 + s]) q[r * 3 + s] = VecIO<T>::load_raw(x + ((img * H + hi) * W + wi) * C + cv * V);
    }
  }
  float best[V];
  int bp[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { best[i] = -INFINITY; bp[i] = 0; }
  float sc[V], sh[V];
  if (pre.ss) {
    const float* p = pre.ss + ((img / pre.imgs_per_group) * C + cv * V) * 2;
#pragma unroll
    for (int i = 0; i < V; i += 2) {
      const float4 f = *reinterpret_cast<const float4*>(p + 2 * i);
      sc[i] = f.x; sh[i] = f.y; sc[i + 1] = f.z; sh[i + 1] = f.w;
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (ok[t]) {
      float v[V];
      VecIO<T>::unpack(q[t], v);
      if (pre.ss) {
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = act_apply(fmaf(v[i], sc[i], sh[i]), pre.act);
      }
#pragma unroll
      for (int i = 0; i < V; ++i)
        if (v[i] > best[i] || v[i] != v[i]) { best[i] = v[i]; bp[i] = t; }
    }
  }
  VecIO<T>::store(y + pix * C + cv * V, best);
  if (pos) {
    unsigned char* pp = pos + pix * C + cv * V;
    if (V == 8) {
      uint2 pk;
      pk.x = (unsigned)bp[0] | ((unsigned)bp[1] << 8) | ((unsigned)bp[2] << 16) | ((unsigned)bp[3] << 24);
      pk.y = (unsigned)bp[4 % V] | ((unsigned)bp[5 % V] << 8) | ((unsigned)bp[6 % V] << 16) | ((unsigned)bp[7 % V] << 24);
      *reinterpret_cast<uint2*>(pp) = pk;
    } else {
      unsigned pk = (unsigned)bp[0] | ((unsigned)bp[1] << 8) | ((unsigned)bp[2] << 16) | ((unsigned)bp[3] << 24);
      *reinterpret_cast<unsigned*>(pp) = pk;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool_fwd_vec_kernel(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ pos, int IMGS, int H,
                       int W, int C, int Ho, int Wo, LiveLimit live) {
  maxpool_fwd_vec_body<T, const T*, T*>(x, y, pos, IMGS, H, W, C, Ho, Wo, live);
}
// x2 max-pool (plain, or the training stem's maxpool(act(z * scale + shift))): one window ROW (3 taps = 6 x 16 bytes) in
// flight per thread instead of the whole window, so the kernel fits 3 blocks per SM without spilling (the whole-window
// body needs 135 registers -> ONE 256-thread block per SM; ncu: 12 % warps active, 6.46 ms for 12.1 GB at the N=72 stem
// shape), and one output row (img, ho) per block: the row decode is block-uniform and 32-bit, the per-thread decode one
// small division -- with a flat index the five 64-bit divisions cost as many instructions as the nine taps.
// 6.46 -> 3.11 ms.  (IMGS * Ho < 2^31 rows, Wo * C / 8 < 2^31: checked by the launcher)
__global__ void __launch_bounds__(256, 3)
maxpool_rows_x2_kernel(X2CPtr z, X2Ptr y, unsigned char* __restrict__ pos, int IMGS, int H, int W, int C, int Ho, int Wo,
                       LiveLimit live, PreBN pre) {
  constexpr int V = 8;
  const int cvecs = C / V;
  const unsigned rowid = blockIdx.x;
  const int ho = (int)(rowid % (unsigned)Ho);
  const long long img = rowid / (unsigned)Ho;
  if (img >= live_count(live, IMGS)) return;  // device-side work limit (inference with skipping)
  const unsigned it = threadIdx.x + blockIdx.y * blockDim.x;
  if (it >= (unsigned)(Wo * cvecs)) return;
  const int wo = (int)(it / (unsigned)cvecs);
  const int cv = (int)(it - (unsigned)wo * (unsigned)cvecs);
  const long long pix = (long long)rowid * Wo + wo;
  float best[V], sc[V], sh[V];
  unsigned bp = 0;  // 4 bits per channel
#pragma unroll
  for (int i = 0; i < V; ++i) { best[i] = -INFINITY; sc[i] = 1.f; sh[i] = 0.f; }
  if (pre.ss) {
    const float* p = pre.ss + ((img / pre.imgs_per_group) * C + cv * V) * 2;
#pragma unroll
    for (int i = 0; i < V; i += 2) {
      const float4 f = *reinterpret_cast<const float4*>(p + 2 * i);
      sc[i] = f.x; sh[i] = f.y; sc[i + 1] = f.z; sh[i + 1] = f.w;
    }
  }
  const int wi0 = wo * 2 - 1;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hi = ho * 2 + r - 1;
    if (hi < 0 || hi >= H) continue;
    const X2CPtr row = z + ((img * H + hi) * W) * C + cv * V;
    X2Raw q[3];
    bool ok[3];
#pragma unroll
    for (int s_ = 0; s_ < 3; ++s_) {
      const int wi = wi0 + s_;
      ok[s_] = wi >= 0 && wi < W;
      if (ok[s_]) q[s_] = VecIO<x2_t>::load_raw(row + (long long)wi * C);
    }
#pragma unroll
    for (int s_ = 0; s_ < 3; ++s_) {
      if (ok[s_]) {
        float v[V];
        VecIO<x2_t>::unpack(q[s_], v);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float a = pre.ss ? act_apply(fmaf(v[i], sc[i], sh[i]), pre.act) : v[i];
          if (a > best[i] || a != a) {
            best[i] = a;
            bp = (bp & ~(0xFu << (4 * i))) | ((unsigned)(r * 3 + s_) << (4 * i));
          }
        }
      }
    }
  }
  VecIO<x2_t>::store(y + pix * C + cv * V, best);
  if (pos) {
    uint2 pk;
    pk.x = (bp & 0xFu) | ((bp & 0xF0u) << 4) | ((bp & 0xF00u) << 8) | ((bp & 0xF000u) << 12);
    pk.y = ((bp >> 16) & 0xFu) | ((bp >> 12) & 0xF00u) | ((bp >> 8) & 0xF0000u) | ((bp >> 4) & 0xF000000u);
    *reinterpret_cast<uint2*>(pos + pix * C + cv * V) = pk;
  }
}

int launch_maxpool_rows_x2(X2CPtr z, X2Ptr y, unsigned char* pos, int IMGS, int H, int W, int C, int Ho, int Wo,
                           LiveLimit live, PreBN pre, cudaStream_t stream) {
  ADAMML_REQUIRE((long long)IMGS * Ho < (1LL << 31) && (long long)Wo * (C / 8) < (1LL << 31), "maxpool_x2: too large");
  // threads of one output row, split evenly over as few blocks (<= 256 threads, whole warps) as it takes
  const int per_row = Wo * (C / 8);
  const int parts = (per_row + 255) / 256;
  const int threads = (((per_row + parts - 1) / parts) + 31) / 32 * 32;
  ADAMML_REQUIRE(parts <= 65535, "maxpool_x2: row too wide");
  maxpool_rows_x2_kernel<<<dim3((unsigned)((long long)IMGS * Ho), (unsigned)parts), threads, 0, stream>>>(
      z, y, pos, IMGS, H, W, C, Ho, Wo, live, pre);
  return ADAMML_OK;
}

// gather backward from recorded positions, one thread per 2x2 input quad (rows 2m, 2m+1; cols 2n, 2n+1) and channel
// vector: the quad is covered by the windows (m..m+1) x (n..n+1) only, so 4 (position, dy) loads feed 4 outputs.
// Input pixel (hi, wi) sits at row r = hi + 1 - 2*ho, column s = wi + 1 - 2*wo of window (ho, wo).
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_pos_kernel(const unsigned char* __restrict__ pos, const T* __restrict__ dy, T* __restrict__ dx, int IMGS,
                       int H, int W, int C, int Ho, int Wo) {
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
  float g[2][2][V];
  unsigned char pc[2][2][8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const bool ok = (m + a) < Ho && (n + b) < Wo;
      const long long o = ((img * Ho + m + a) * Wo + n + b) * C + cv * V;
      if (ok) {
        VecIO<T>::load(dy + o, g[a][b]);
        if (V == 8) *reinterpret_cast<uint2*>(pc[a][b]) = *reinterpret_cast<const uint2*>(pos + o);
        else *reinterpret_cast<unsigned*>(pc[a][b]) = *reinterpret_cast<const unsigned*>(pos + o);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) { g[a][b][i] = 0.f; pc[a][b][i] = 255; }
      }
    }
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
      // windows containing row hi: ho = m (r = p + 1) and, for odd rows, ho = m + 1 (r = 0)
#pragma unroll
      for (int a = 0; a <= p; ++a) {
        const int r = hi + 1 - 2 * (m + a);
#pragma unroll
        for (int b = 0; b <= q; ++b) {
          const int code = r * 3 + (wi + 1 - 2 * (n + b));
#pragma unroll
          for (int i = 0; i < V; ++i)
            if (pc[a][b][i] == code) acc[i] += g[a][b][i];
        }
      }
      VecIO<T>::store(dx + ((img * H + hi) * W + wi) * C + cv * V, acc);
    }
  }
}

// temporal pool k3 s2 p1, one thread per (video, element vector): all TN frames of that position in registers
template <typename T, int TN, typename CP, typename MP>
__device__ __forceinline__ void tpool_fwd_vec_body(CP x, MP y, long long V_, long long E, int mode_avg,
                                                   LiveLimit live) {
  constexpr int V = VecIO<T>::N;
  constexpr int TO = (TN + 2 - 3) / 2 + 1;
  const long long evecs = E / V;
  const long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= V_ * evecs) return;
  const long long v = iv / evecs, e = (iv % evecs) * V;
  if (v >= live_count(live, V_)) return;
  typename VecIO<T>::raw q[TN];
#pragma unroll
  for (int t = 0; t < TN; ++t) q[t] = VecIO<T>::load_raw(x + (v * TN + t) * E + e);
  float f[TN][V];
#pragma unroll
  for (int t = 0; t < TN; ++t) VecIO<T>::unpack(q[t], f[t]);
#pragma unroll
  for (int to = 0; to < TO; ++to) {
    float best[V], sum[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { best[i] = -INFINITY; sum[i] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      const int t = to * 2 + kk - 1;
      if (t < 0 || t >= TN) continue;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        sum[i] += f[t][i];
        if (f[t][i] > best[i] || f[t][i] != f[t][i]) best[i] = f[t][i];
      }
    }
    if (mode_avg) {
#pragma unroll
      for (int i = 0; i < V; ++i) best[i] = sum[i] / 3.f;
    }
    VecIO<T>::store(y + (v * TO + to) * E + e, best);
  }
}

template <typename T, int TN>
__global__ void __launch_bounds__(256)
tpool_fwd_vec_kernel(const T* __restrict__ x, T* __restrict__ y, long long V_, long long E, int mode_avg,
                     LiveLimit live) {
  tpool_fwd_vec_body<T, TN, const T*, T*>(x, y, V_, E, mode_avg, live);
}
template <int TN>
__global__ void __launch_bounds__(256)
tpool_fwd_vec_x2_kernel(X2CPtr x, X2Ptr y, long long V_, long long E, int mode_avg, LiveLimit live) {
  tpool_fwd_vec_body<x2_t, TN, X2CPtr, X2Ptr>(x, y, V_, E, mode_avg, live);
}

template <typename T, int TN>
__global__ void __launch_bounds__(256)
tpool_bwd_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long long V_, long long E,
                     int mode_avg) {
  constexpr int V = VecIO<T>::N;
  constexpr int TO = (TN + 2 - 3) / 2 + 1;
  const long long evecs = E / V;
  const long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= V_ * evecs) return;
  const long long v = iv / evecs, e = (iv % evecs) * V;
  typename VecIO<T>::raw q[TN], qg[TO];
#pragma unroll
  for (int t = 0; t < TN; ++t) q[t] = VecIO<T>::load_raw(x + (v * TN + t) * E + e);
#pragma unroll
  for (int to = 0; to < TO; ++to) qg[to] = VecIO<T>::load_raw(dy + (v * TO + to) * E + e);
  float f[TN][V], d[TN][V];
#pragma unroll
  for (int t = 0; t < TN; ++t) {
    VecIO<T>::unpack(q[t], f[t]);
#pragma unroll
    for (int i = 0; i < V; ++i) d[t][i] = 0.f;
  }
#pragma unroll
  for (int to = 0; to < TO; ++to) {
    float g[V];
    VecIO<T>::unpack(qg[to], g);
    if (mode_avg) {
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        const int t = to * 2 + kk - 1;
        if (t < 0 || t >= TN) continue;
#pragma unroll
        for (int i = 0; i < V; ++i) d[t][i] += g[i] / 3.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float best = -INFINITY;
        int bt = -1;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int t = to * 2 + kk - 1;
          if (t < 0 || t >= TN) continue;
          if (f[t][i] > best || f[t][i] != f[t][i]) { best = f[t][i]; bt = t; }
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int t = to * 2 + kk - 1;
          if (t < 0 || t >= TN) continue;
          if (bt == t) d[t][i] += g[i];
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < TN; ++t) VecIO<T>::store(dx + (v * TN + t) * E + e, d[t]);
}

template <typename T>
inline bool pool_vec_ok(long long C, const void* a, const void* b = nullptr, const void* c = nullptr) {
  if (C % VecIO<T>::N) return false;
  const void* ps[3] = {a, b, c};
  for (int i = 0; i < 3; ++i)
    if (ps[i] && ((uintptr_t)ps[i] % 16)) return false;
  return true;
}

// y[img][c] = mean over HW. block (32 channels, 8 pixel lanes) per (img, channel tile)
template <typename CP>
__global__ void avgpool_fwd_kernel(CP x, float* __restrict__ y, int HW, int C, long long y_ld, LiveLimit live) {
  __shared__ float sh[8][33];
  int img = blockIdx.x;
  if (img >= live_count(live, gridDim.x)) return;  // (uniform over the block)
  int c = blockIdx.y * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C)
    for (int p = threadIdx.y; p < HW; p += 8) s += ld_elem(x, ((long long)img * HW + p) * C + c);
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) s += sh[i][threadIdx.x];
    y[(long long)img * y_ld + c] = s / (float)HW;
  }
}

template <typename T>
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, T* __restrict__ dx, long long total, int HW, int C,
                                   long long dy_ld) {
  float inv = 1.f / (float)HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    long long img = idx / ((long long)HW * C);
    dx[idx] = from_f32<T>(dy[img * dy_ld + c] * inv);
  }
}

}  // namespace

extern "C" {

int adamml_maxpool3x3s2_fwd(const void* x, void* y, unsigned char* pos, int IMGS, int H, int W, int C, int Ho, int Wo,
                            int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / 2 + 1 && Wo == (W + 2 - 3) / 2 + 1, "maxpool: bad Ho/Wo");
  long long total = (long long)IMGS * Ho * Wo * C;
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (pool_vec_ok<T>(C, x, y) && ((uintptr_t)pos % 8) == 0) {
      long long tv = total / VecIO<T>::N;
      maxpool_fwd_vec_kernel<T><<<(unsigned)((tv + 255) / 256), 256, 0, stream>>>(
          (const T*)x, (T*)y, pos, IMGS, H, W, C, Ho, Wo, pos ? LiveLimit{nullptr, 0} : adamml_live_limit(IMGS));
    } else {
      ADAMML_REQUIRE(pos == nullptr, "maxpool: position output needs C %% 8 == 0 (bf16) / C %% 4 == 0 (fp32)");
      maxpool_fwd_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)x, (T*)y, IMGS, H, W, C, Ho, Wo);
    }
  });
  return adamml_check_launch("maxpool_fwd");
}

int adamml_maxpool3x3s2_bwd(const void* x, const unsigned char* pos, const void* dy, void* dx, int IMGS, int H, int W,
                            int C, int Ho, int Wo, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(x || pos, "maxpool_bwd: needs either the forward input or the recorded positions");
  long long total = (long long)IMGS * H * W * C;
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (pos) {
      ADAMML_REQUIRE(pool_vec_ok<T>(C, dy, dx) && ((uintptr_t)pos % 8) == 0, "maxpool_bwd: unaligned / ragged C");
      long long tv = (long long)IMGS * ((H + 1) / 2) * ((W + 1) / 2) * (C / VecIO<T>::N);
      maxpool_bwd_pos_kernel<T><<<(unsigned)((tv + 255) / 256), 256, 0, stream>>>(pos, (const T*)dy, (T*)dx, IMGS, H,
                                                                                   W, C, Ho, Wo);
    } else {
      maxpool_bwd_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)x, (const T*)dy, (T*)dx, IMGS, H, W, C, Ho,
                                                                 Wo);
    }
  });
  return adamml_check_launch("maxpool_bwd");
}

// x: [V videos, Tn frames, E = H*W*C]
int adamml_tpool_fwd(const void* x, void* y, long long V, int Tn, long long E, int mode_avg, int dtype,
                     cudaStream_t stream) {
  ADAMML_REQUIRE(V > 0 && Tn > 0 && E > 0, "tpool: empty dims");
  int To = (Tn + 2 - 3) / 2 + 1;
  long long total = V * To * E;
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    const bool vec = pool_vec_ok<T>(E, x, y) && (Tn == 2 || Tn == 4 || Tn == 8);
    const long long tv = V * (E / VecIO<T>::N);
    const unsigned vb = (unsigned)((tv + 255) / 256);
    const LiveLimit live = adamml_live_limit(V);
    if (vec && Tn == 8) tpool_fwd_vec_kernel<T, 8><<<vb, 256, 0, stream>>>((const T*)x, (T*)y, V, E, mode_avg, live);
    else if (vec && Tn == 4) tpool_fwd_vec_kernel<T, 4><<<vb, 256, 0, stream>>>((const T*)x, (T*)y, V, E, mode_avg, live);
    else if (vec && Tn == 2) tpool_fwd_vec_kernel<T, 2><<<vb, 256, 0, stream>>>((const T*)x, (T*)y, V, E, mode_avg, live);
    else tpool_fwd_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)x, (T*)y, V, Tn, To, E, mode_avg);
  });
  return adamml_check_launch("tpool_fwd");
}

int adamml_tpool_bwd(const void* x, const void* dy, void* dx, long long V, int Tn, long long E, int mode_avg,
                     int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(V > 0 && Tn > 0 && E > 0, "tpool: empty dims");
  int To = (Tn + 2 - 3) / 2 + 1;
  long long total = V * Tn * E;
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    const bool vec = pool_vec_ok<T>(E, x, dy, dx) && (Tn == 2 || Tn == 4 || Tn == 8);
    const long long tv = V * (E / VecIO<T>::N);
    const unsigned vb = (unsigned)((tv + 255) / 256);
    if (vec && Tn == 8)
      tpool_bwd_vec_kernel<T, 8><<<vb, 256, 0, stream>>>((const T*)x, (const T*)dy, (T*)dx, V, E, mode_avg);
    else if (vec && Tn == 4)
      tpool_bwd_vec_kernel<T, 4><<<vb, 256, 0, stream>>>((const T*)x, (const T*)dy, (T*)dx, V, E, mode_avg);
    else if (vec && Tn == 2)
      tpool_bwd_vec_kernel<T, 2><<<vb, 256, 0, stream>>>((const T*)x, (const T*)dy, (T*)dx, V, E, mode_avg);
    else
      tpool_bwd_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)x, (const T*)dy, (T*)dx, V, Tn, To, E,
                                                               mode_avg);
  });
  return adamml_check_launch("tpool_bwd");
}

int adamml_avgpool_fwd(const void* x, float* y, int IMGS, int HW, int C, long long y_ld, int dtype,
                       cudaStream_t stream) {
  ADAMML_REQUIRE(IMGS > 0 && HW > 0 && C > 0, "avgpool: empty dims");
  dim3 grid(IMGS, ceil_div(C, 32));
  dim3 block(32, 8);
  if (y_ld <= 0) y_ld = C;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    avgpool_fwd_kernel<const T*><<<grid, block, 0, stream>>>((const T*)x, y, HW, C, y_ld, adamml_live_limit(IMGS)));
  return adamml_check_launch("avgpool_fwd");
}

/* ---- x2 planes (forward pass of the default precision mode); channel counts are multiples of 8 ---- */
int adamml_maxpool3x3s2_fwd_x2(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, unsigned char* pos,
                               int IMGS, int H, int W, int C, int Ho, int Wo, cudaStream_t stream) {
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / 2 + 1 && Wo == (W + 2 - 3) / 2 + 1, "maxpool: bad Ho/Wo");
  ADAMML_REQUIRE(pool_vec_ok<bf16>(C, x_hi, x_lo, y_hi) && pool_vec_ok<bf16>(C, y_lo) && ((uintptr_t)pos % 8) == 0,
                 "maxpool_fwd_x2: needs C %% 8 == 0 and aligned planes");
  const int rc = launch_maxpool_rows_x2(x2c(x_hi, x_lo), x2m(y_hi, y_lo), pos, IMGS, H, W, C, Ho, Wo,
                                        pos ? LiveLimit{nullptr, 0} : adamml_live_limit(IMGS),
                                        PreBN{nullptr, 1, ADAMML_ACT_NONE}, stream);
  if (rc) return rc;
  return adamml_check_launch("maxpool_fwd_x2");
}

/* training stem (resnet.py:197-200): y = maxpool3x3s2(act(z * scale + shift)) straight from the pre-BatchNorm planes z;
 * scale_shift = float [G][C][2] of adamml_bn_finalize, group = img / imgs_per_group.  pos as adamml_maxpool3x3s2_fwd. */
int adamml_bn_act_maxpool3x3s2_fwd_x2(const void* z_hi, const void* z_lo, const float* scale_shift, int imgs_per_group,
                                      int act, void* y_hi, void* y_lo, unsigned char* pos, int IMGS, int H, int W,
                                      int C, int Ho, int Wo, cudaStream_t stream) {
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / 2 + 1 && Wo == (W + 2 - 3) / 2 + 1, "bn_act_maxpool: bad Ho/Wo");
  ADAMML_REQUIRE(scale_shift && imgs_per_group > 0 && IMGS % imgs_per_group == 0, "bn_act_maxpool: bad BN groups");
  ADAMML_REQUIRE(pool_vec_ok<bf16>(C, z_hi, z_lo, y_hi) && pool_vec_ok<bf16>(C, y_lo) && ((uintptr_t)pos % 8) == 0 &&
                     ((uintptr_t)scale_shift % 16) == 0,
                 "bn_act_maxpool_fwd_x2: needs C %% 8 == 0 and aligned planes");
  const int rc = launch_maxpool_rows_x2(x2c(z_hi, z_lo), x2m(y_hi, y_lo), pos, IMGS, H, W, C, Ho, Wo,
                                        LiveLimit{nullptr, 0}, PreBN{scale_shift, imgs_per_group, act}, stream);
  if (rc) return rc;
  return adamml_check_launch("bn_act_maxpool_fwd_x2");
}

int adamml_tpool_fwd_x2(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, long long V, int Tn, long long E,
                        int mode_avg, cudaStream_t stream) {
  ADAMML_REQUIRE(V > 0 && E > 0, "tpool: empty dims");
  ADAMML_REQUIRE(Tn == 2 || Tn == 4 || Tn == 8, "tpool_fwd_x2: 2, 4 or 8 frames");
  ADAMML_REQUIRE(pool_vec_ok<bf16>(E, x_hi, x_lo, y_hi) && pool_vec_ok<bf16>(E, y_lo),
                 "tpool_fwd_x2: needs E %% 8 == 0 and aligned planes");
  const long long tv = V * (E / 8);
  const unsigned vb = (unsigned)((tv + 255) / 256);
  const X2CPtr x = x2c(x_hi, x_lo);
  const X2Ptr y = x2m(y_hi, y_lo);
  const LiveLimit live = adamml_live_limit(V);
  if (Tn == 8) tpool_fwd_vec_x2_kernel<8><<<vb, 256, 0, stream>>>(x, y, V, E, mode_avg, live);
  else if (Tn == 4) tpool_fwd_vec_x2_kernel<4><<<vb, 256, 0, stream>>>(x, y, V, E, mode_avg, live);
  else tpool_fwd_vec_x2_kernel<2><<<vb, 256, 0, stream>>>(x, y, V, E, mode_avg, live);
  return adamml_check_launch("tpool_fwd_x2");
}

int adamml_avgpool_fwd_x2(const void* x_hi, const void* x_lo, float* y, int IMGS, int HW, int C, long long y_ld,
                          cudaStream_t stream) {
  ADAMML_REQUIRE(IMGS > 0 && HW > 0 && C > 0, "avgpool: empty dims");
  dim3 grid(IMGS, ceil_div(C, 32));
  dim3 block(32, 8);
  if (y_ld <= 0) y_ld = C;
  avgpool_fwd_kernel<X2CPtr><<<grid, block, 0, stream>>>(x2c(x_hi, x_lo), y, HW, C, y_ld, adamml_live_limit(IMGS));
  return adamml_check_launch("avgpool_fwd_x2");
}

int adamml_avgpool_bwd(const float* dy, void* dx, int IMGS, int HW, int C, long long dy_ld, int dtype,
                       cudaStream_t stream) {
  long long total = (long long)IMGS * HW * C;
  if (dy_ld <= 0) dy_ld = C;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    avgpool_bwd_kernel<T><<<ew_blocks(total), 256, 0, stream>>>(dy, (T*)dx, total, HW, C, dy_ld));
  return adamml_check_launch("avgpool_bwd");
}

}  // extern "C"
