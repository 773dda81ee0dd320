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

// y[img][c] = mean over HW. block (32 channels, 8 pixel lanes) per (img, channel tile)
template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, float* __restrict__ y, int HW, int C, long long y_ld) {
  __shared__ float sh[8][33];
  int img = blockIdx.x;
  int c = blockIdx.y * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C)
    for (int p = threadIdx.y; p < HW; p += 8) s += to_f32(x[((long long)img * HW + p) * C + c]);
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

int adamml_maxpool3x3s2_fwd(const void* x, void* y, int IMGS, int H, int W, int C, int Ho, int Wo, int dtype,
                            cudaStream_t stream) {
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / 2 + 1 && Wo == (W + 2 - 3) / 2 + 1, "maxpool: bad Ho/Wo");
  long long total = (long long)IMGS * Ho * Wo * C;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    maxpool_fwd_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)x, (T*)y, IMGS, H, W, C, Ho, Wo));
  return adamml_check_launch("maxpool_fwd");
}

int adamml_maxpool3x3s2_bwd(const void* x, const void* dy, void* dx, int IMGS, int H, int W, int C, int Ho, int Wo,
                            int dtype, cudaStream_t stream) {
  long long total = (long long)IMGS * H * W * C;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    maxpool_bwd_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)x, (const T*)dy, (T*)dx, IMGS, H, W, C, Ho, Wo));
  return adamml_check_launch("maxpool_bwd");
}

// x: [V videos, Tn frames, E = H*W*C]
int adamml_tpool_fwd(const void* x, void* y, long long V, int Tn, long long E, int mode_avg, int dtype,
                     cudaStream_t stream) {
  ADAMML_REQUIRE(V > 0 && Tn > 0 && E > 0, "tpool: empty dims");
  int To = (Tn + 2 - 3) / 2 + 1;
  long long total = V * To * E;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    tpool_fwd_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)x, (T*)y, V, Tn, To, E, mode_avg));
  return adamml_check_launch("tpool_fwd");
}

int adamml_tpool_bwd(const void* x, const void* dy, void* dx, long long V, int Tn, long long E, int mode_avg,
                     int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(V > 0 && Tn > 0 && E > 0, "tpool: empty dims");
  int To = (Tn + 2 - 3) / 2 + 1;
  long long total = V * Tn * E;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    tpool_bwd_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)x, (const T*)dy, (T*)dx, V, Tn, To, E, mode_avg));
  return adamml_check_launch("tpool_bwd");
}

int adamml_avgpool_fwd(const void* x, float* y, int IMGS, int HW, int C, long long y_ld, int dtype,
                       cudaStream_t stream) {
  ADAMML_REQUIRE(IMGS > 0 && HW > 0 && C > 0, "avgpool: empty dims");
  dim3 grid(IMGS, ceil_div(C, 32));
  dim3 block(32, 8);
  if (y_ld <= 0) y_ld = C;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    avgpool_fwd_kernel<T><<<grid, block, 0, stream>>>((const T*)x, y, HW, C, y_ld));
  return adamml_check_launch("avgpool_fwd");
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
