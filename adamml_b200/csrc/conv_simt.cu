// Generic implicit-GEMM convolution on CUDA cores (fp32 accumulate, fp32 math).
//
// This is the exact-arithmetic engine of adamml_b200: every dense convolution / linear
// layer of the AdaMML hot path (reference: models/resnet.py:35-43,96-111,
// models/sound_mobilenet_v2.py:33-69, models/policy_net.py:38-95,228-231,278-279) can run
// through it in fp32 so that logits match the PyTorch fp32 path to ~1e-5.  The tcgen05
// engine (conv_tc.cu) takes over the bf16 shapes it supports; this file stays the
// fallback for odd shapes (Cin=1/3/15 stems, Cout=2/31 heads) and the parity mode.
//
// Layouts: activations NHWC [IMGS,H,W,C] (pixel stride *_ld elements), weights
// [Cout][R][S][Cin] ("OHWI", row stride w_ld).  No im2col buffer is ever materialised.
#include "common.cuh"

namespace {

struct ConvP {
  int IMGS, H, W, Cin;
  int Cout, R, S, stride, pad;
  int Ho, Wo;
  long long x_ld, y_ld, w_ld;
};

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, NTHREADS = 256;
constexpr int APAD = 4;

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

// ---- element fetchers -------------------------------------------------------------
// FWD   : A(m,k) = x[img, ho*st+r-pad, wo*st+s-pad, ci]   m=(img,ho,wo) k=(r,s,ci)
//         B(k,n) = w[n][k]
// DGRAD : A(m,k) = dy[img, (h+pad-r)/st, (w+pad-s)/st, co] m=(img,h,w)  k=(r,s,co)
//         B(k,n) = w[co][r][s][n]
// WGRAD : A(m,k) = dy[k][m]                                 m=co, k=(img,ho,wo)
//         B(k,n) = x[img, ho*st+r-pad, wo*st+s-pad, ci]     n=(r,s,ci)

template <typename T, int MODE>
__global__ void __launch_bounds__(NTHREADS)
igemm_kernel(const T* __restrict__ Aptr, const T* __restrict__ Bptr, void* __restrict__ Out,
             const T* __restrict__ addend, ConvP p, long long M, int N, long long K,
             long long k_per_split) {
  __shared__ float As[BK][BM + APAD];
  __shared__ float Bs[BK][BN + APAD];

  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long kbeg = (long long)blockIdx.z * k_per_split;
  long long kend = kbeg + k_per_split;
  if (kend > K) kend = K;

  const int tx = tid % 16, ty = tid / 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // Per-thread fixed row decode for the gather side.
  // FWD/DGRAD: A rows (k fastest mapping): m_l = tid/16 + 16*i, k_l = tid%16
  // WGRAD    : B cols (n fastest mapping): k_l = tid/64 + 4*i,  n_l = tid%64
  int a_img[4], a_h[4], a_w[4];
  bool a_ok[4];
  if (MODE == MODE_FWD || MODE == MODE_DGRAD) {
    const int HW = (MODE == MODE_FWD) ? p.Ho * p.Wo : p.H * p.W;
    const int Wd = (MODE == MODE_FWD) ? p.Wo : p.W;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long m = m0 + tid / 16 + 16 * i;
      a_ok[i] = m < M;
      long long mm = a_ok[i] ? m : 0;
      a_img[i] = (int)(mm / HW);
      int rem = (int)(mm % HW);
      a_h[i] = rem / Wd;
      a_w[i] = rem % Wd;
    }
  }
  int b_r = 0, b_s = 0, b_c = 0;
  bool b_ok = false;
  if (MODE == MODE_WGRAD) {
    int n = n0 + tid % 64;
    b_ok = n < N;
    int nn = b_ok ? n : 0;
    b_c = nn % p.Cin;
    int rs = nn / p.Cin;
    b_r = rs / p.S;
    b_s = rs % p.S;
  }

  for (long long k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- load A tile ----
    if (MODE == MODE_FWD) {
      long long k = k0 + tid % 16;
      bool kok = k < kend;
      int kk = kok ? (int)k : 0;
      int ci = kk % p.Cin;
      int rs = kk / p.Cin;
      int r = rs / p.S, s = rs % p.S;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        int hi = a_h[i] * p.stride + r - p.pad;
        int wi = a_w[i] * p.stride + s - p.pad;
        if (kok && a_ok[i] && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
          v = to_f32(Aptr[(((long long)a_img[i] * p.H + hi) * p.W + wi) * p.x_ld + ci]);
        As[tid % 16][tid / 16 + 16 * i] = v;
      }
    } else if (MODE == MODE_DGRAD) {
      long long k = k0 + tid % 16;
      bool kok = k < kend;
      int kk = kok ? (int)k : 0;
      int co = kk % p.Cout;
      int rs = kk / p.Cout;
      int r = rs / p.S, s = rs % p.S;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        int hn = a_h[i] + p.pad - r;
        int wn = a_w[i] + p.pad - s;
        if (kok && a_ok[i] && hn >= 0 && wn >= 0 && (hn % p.stride) == 0 && (wn % p.stride) == 0) {
          int ho = hn / p.stride, wo = wn / p.stride;
          if (ho < p.Ho && wo < p.Wo)
            v = to_f32(Aptr[(((long long)a_img[i] * p.Ho + ho) * p.Wo + wo) * p.y_ld + co]);
        }
        As[tid % 16][tid / 16 + 16 * i] = v;
      }
    } else {  // WGRAD: A(m,k) = dy[k*y_ld + m], m fastest
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int kl = tid / 64 + 4 * i;
        long long k = k0 + kl;
        long long m = m0 + tid % 64;
        float v = 0.f;
        if (k < kend && m < M) v = to_f32(Aptr[k * p.y_ld + m]);
        As[kl][tid % 64] = v;
      }
    }
    // ---- load B tile ----
    if (MODE == MODE_FWD) {
      long long k = k0 + tid % 16;
      bool kok = k < kend;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int n = n0 + tid / 16 + 16 * i;
        float v = 0.f;
        if (kok && n < N) v = to_f32(Bptr[(long long)n * p.w_ld + k]);
        Bs[tid % 16][tid / 16 + 16 * i] = v;
      }
    } else if (MODE == MODE_DGRAD) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int kl = tid / 64 + 4 * i;
        long long k = k0 + kl;
        int n = n0 + tid % 64;
        float v = 0.f;
        if (k < kend && n < N) {
          int co = (int)(k % p.Cout);
          int rs = (int)(k / p.Cout);
          v = to_f32(Bptr[(long long)co * p.w_ld + (long long)rs * p.Cin + n]);
        }
        Bs[kl][tid % 64] = v;
      }
    } else {  // WGRAD: B(k,n) gather from x
      const int HWo = p.Ho * p.Wo;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int kl = tid / 64 + 4 * i;
        long long k = k0 + kl;
        float v = 0.f;
        if (k < kend && b_ok) {
          int img = (int)(k / HWo);
          int rem = (int)(k % HWo);
          int ho = rem / p.Wo, wo = rem % p.Wo;
          int hi = ho * p.stride + b_r - p.pad;
          int wi = wo * p.stride + b_s - p.pad;
          if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
            v = to_f32(Bptr[(((long long)img * p.H + hi) * p.W + wi) * p.x_ld + b_c]);
        }
        Bs[kl][tid % 64] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= N) continue;
      if (MODE == MODE_FWD) {
        reinterpret_cast<T*>(Out)[m * p.y_ld + n] = from_f32<T>(acc[i][j]);
      } else if (MODE == MODE_DGRAD) {
        float v = acc[i][j];
        if (addend) v += to_f32(addend[m * p.x_ld + n]);
        reinterpret_cast<T*>(Out)[m * p.x_ld + n] = from_f32<T>(v);
      } else {
        atomicAdd(&reinterpret_cast<float*>(Out)[m * p.w_ld + n], acc[i][j]);
      }
    }
  }
}

// ---- depthwise 3x3 (pad 1, stride 1|2), weights fp32 tap-major [9][C] (scalar fallbacks) ------
template <typename T>
__global__ void dw_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                              int IMGS, int H, int W, int C, int stride, int Ho, int Wo) {
  long long total = (long long)IMGS * Ho * Wo * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    long long pix = idx / C;
    int wo = (int)(pix % Wo);
    int ho = (int)((pix / Wo) % Ho);
    int img = (int)(pix / ((long long)Wo * Ho));
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int hi = ho * stride + r - 1;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int wi = wo * stride + s - 1;
        if (wi < 0 || wi >= W) continue;
        acc = fmaf(to_f32(x[(((long long)img * H + hi) * W + wi) * C + c]), w[(r * 3 + s) * C + c], acc);
      }
    }
    y[idx] = from_f32<T>(acc);
  }
}

template <typename T>
__global__ void dw_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx,
                                const T* __restrict__ addend,
                                int IMGS, int H, int W, int C, int stride, int Ho, int Wo) {
  long long total = (long long)IMGS * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    long long pix = idx / C;
    int wi = (int)(pix % W);
    int hi = (int)((pix / W) % H);
    int img = (int)(pix / ((long long)W * H));
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int hn = hi + 1 - r;
      if (hn < 0 || (hn % stride) != 0) continue;
      int ho = hn / stride;
      if (ho >= Ho) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int wn = wi + 1 - s;
        if (wn < 0 || (wn % stride) != 0) continue;
        int wo = wn / stride;
        if (wo >= Wo) continue;
        acc = fmaf(to_f32(dy[(((long long)img * Ho + ho) * Wo + wo) * C + c]), w[(r * 3 + s) * C + c], acc);
      }
    }
    if (addend) acc += to_f32(addend[idx]);
    dx[idx] = from_f32<T>(acc);
  }
}

// block = 128 channels x 2 pixel lanes; each block covers PIX_PER_BLOCK output pixels.
constexpr int DW_PIX_PER_BLOCK = 512;
template <typename T>
__global__ void dw_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                                int IMGS, int H, int W, int C, int stride, int Ho, int Wo) {
  int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= C) return;
  long long P = (long long)IMGS * Ho * Wo;
  long long p0 = (long long)blockIdx.y * DW_PIX_PER_BLOCK;
  long long p1 = p0 + DW_PIX_PER_BLOCK;
  if (p1 > P) p1 = P;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  for (long long pix = p0 + threadIdx.y; pix < p1; pix += blockDim.y) {
    int wo = (int)(pix % Wo);
    int ho = (int)((pix / Wo) % Ho);
    int img = (int)(pix / ((long long)Wo * Ho));
    float g = to_f32(dy[pix * C + c]);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int hi = ho * stride + r - 1;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int wi = wo * stride + s - 1;
        if (wi < 0 || wi >= W) continue;
        acc[r * 3 + s] = fmaf(g, to_f32(x[(((long long)img * H + hi) * W + wi) * C + c]), acc[r * 3 + s]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) atomicAdd(&dw[i * C + c], acc[i]);
}

int check_conv(const ConvP& p) {
  ADAMML_REQUIRE(p.IMGS > 0 && p.H > 0 && p.W > 0 && p.Cin > 0 && p.Cout > 0, "conv: empty dims");
  ADAMML_REQUIRE(p.R > 0 && p.S > 0 && p.stride > 0 && p.pad >= 0, "conv: bad filter geometry");
  ADAMML_REQUIRE(p.Ho == (p.H + 2 * p.pad - p.R) / p.stride + 1 && p.Wo == (p.W + 2 * p.pad - p.S) / p.stride + 1,
                 "conv: Ho/Wo inconsistent with H/W/R/S/stride/pad");
  ADAMML_REQUIRE(p.x_ld >= p.Cin && p.y_ld >= p.Cout && p.w_ld >= (long long)p.R * p.S * p.Cin,
                 "conv: leading dimensions too small");
  return ADAMML_OK;
}

ConvP make_p(int IMGS, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int Ho, int Wo,
             long long x_ld, long long y_ld, long long w_ld) {
  ConvP p;
  p.IMGS = IMGS; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S;
  p.stride = stride; p.pad = pad; p.Ho = Ho; p.Wo = Wo;
  p.x_ld = x_ld > 0 ? x_ld : Cin;
  p.y_ld = y_ld > 0 ? y_ld : Cout;
  p.w_ld = w_ld > 0 ? w_ld : (long long)R * S * Cin;
  return p;
}

}  // namespace

extern "C" {

int adamml_simt_conv_fwd(const void* x, const void* w, void* y, int IMGS, int H, int W, int Cin, int Cout,
                         int R, int S, int stride, int pad, int Ho, int Wo, long long x_ld, long long y_ld,
                         long long w_ld, int dtype, cudaStream_t stream) {
  ConvP p = make_p(IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, x_ld, y_ld, w_ld);
  int rc = check_conv(p);
  if (rc) return rc;
  long long M = (long long)IMGS * Ho * Wo;
  long long K = (long long)R * S * Cin;
  dim3 grid(ceil_div(M, BM), ceil_div(Cout, BN), 1);
  ADAMML_DISPATCH_DTYPE(dtype, T,
    igemm_kernel<T, MODE_FWD><<<grid, NTHREADS, 0, stream>>>((const T*)x, (const T*)w, y, nullptr, p, M, Cout, K, K));
  return adamml_check_launch("simt_conv_fwd");
}

int adamml_simt_conv_dgrad(const void* dy, const void* w, void* dx, const void* addend, int IMGS, int H, int W,
                           int Cin, int Cout, int R, int S, int stride, int pad, int Ho, int Wo, long long x_ld,
                           long long y_ld, long long w_ld, int dtype, cudaStream_t stream) {
  ConvP p = make_p(IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, x_ld, y_ld, w_ld);
  int rc = check_conv(p);
  if (rc) return rc;
  long long M = (long long)IMGS * H * W;
  long long K = (long long)R * S * Cout;
  dim3 grid(ceil_div(M, BM), ceil_div(Cin, BN), 1);
  ADAMML_DISPATCH_DTYPE(dtype, T,
    igemm_kernel<T, MODE_DGRAD><<<grid, NTHREADS, 0, stream>>>((const T*)dy, (const T*)w, dx, (const T*)addend, p, M, Cin, K, K));
  return adamml_check_launch("simt_conv_dgrad");
}

// dw: fp32 [Cout][R][S][Cin] with row stride w_ld; overwritten (zeroed here, split-K atomics).
int adamml_simt_conv_wgrad(const void* x, const void* dy, float* dw, int IMGS, int H, int W, int Cin, int Cout,
                           int R, int S, int stride, int pad, int Ho, int Wo, long long x_ld, long long y_ld,
                           long long w_ld, int dtype, cudaStream_t stream) {
  ConvP p = make_p(IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, x_ld, y_ld, w_ld);
  int rc = check_conv(p);
  if (rc) return rc;
  long long M = Cout;
  int N = R * S * Cin;
  long long K = (long long)IMGS * Ho * Wo;
  int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  // enough K-splits to fill the 148 SMs a few times, at least 256 pixels per split
  long long want = (148LL * 8 + tiles - 1) / tiles;
  long long maxsplit = (K + 255) / 256;
  long long splits = want < maxsplit ? want : maxsplit;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long kps = (K + splits - 1) / splits;
  kps = ((kps + BK - 1) / BK) * BK;
  splits = (K + kps - 1) / kps;
  for (int r = 0; r < Cout; ++r) {
    if (p.w_ld == (long long)N) { cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * N, stream); break; }
    cudaMemsetAsync(dw + (size_t)r * p.w_ld, 0, sizeof(float) * (size_t)N, stream);
  }
  dim3 grid(ceil_div(M, BM), ceil_div(N, BN), (unsigned)splits);
  ADAMML_DISPATCH_DTYPE(dtype, T,
    igemm_kernel<T, MODE_WGRAD><<<grid, NTHREADS, 0, stream>>>((const T*)dy, (const T*)x, dw, nullptr, p, M, N, K, kps));
  return adamml_check_launch("simt_conv_wgrad");
}

}  // extern "C"

// scalar fallbacks (C not a multiple of the 16-byte vector); the C-ABI entry points live in dwconv.cu
int adamml_dwconv_fwd_scalar(const void* x, const float* w, void* y, int IMGS, int H, int W, int C, int stride, int Ho,
                      int Wo, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2");
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / stride + 1 && Wo == (W + 2 - 3) / stride + 1, "dwconv: bad Ho/Wo");
  long long total = (long long)IMGS * Ho * Wo * C;
  int blocks = (int)((total + 255) / 256 < 148LL * 32 ? (total + 255) / 256 : 148LL * 32);
  ADAMML_DISPATCH_DTYPE(dtype, T,
    dw_fwd_kernel<T><<<blocks, 256, 0, stream>>>((const T*)x, w, (T*)y, IMGS, H, W, C, stride, Ho, Wo));
  return adamml_check_launch("dwconv_fwd_scalar");
}

int adamml_dwconv_dgrad_scalar(const void* dy, const float* w, void* dx, const void* addend, int IMGS, int H, int W, int C,
                        int stride, int Ho, int Wo, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2");
  long long total = (long long)IMGS * H * W * C;
  int blocks = (int)((total + 255) / 256 < 148LL * 32 ? (total + 255) / 256 : 148LL * 32);
  ADAMML_DISPATCH_DTYPE(dtype, T,
    dw_dgrad_kernel<T><<<blocks, 256, 0, stream>>>((const T*)dy, w, (T*)dx, (const T*)addend, IMGS, H, W, C, stride, Ho, Wo));
  return adamml_check_launch("dwconv_dgrad_scalar");
}

int adamml_dwconv_wgrad_scalar(const void* x, const void* dy, float* dw, int IMGS, int H, int W, int C, int stride, int Ho,
                        int Wo, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(stride == 1 || stride == 2, "dwconv: stride must be 1 or 2");
  cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * 9, stream);
  long long P = (long long)IMGS * Ho * Wo;
  dim3 grid(ceil_div(C, 128), ceil_div(P, DW_PIX_PER_BLOCK));
  dim3 block(128, 2);
  ADAMML_DISPATCH_DTYPE(dtype, T,
    dw_wgrad_kernel<T><<<grid, block, 0, stream>>>((const T*)x, (const T*)dy, dw, IMGS, H, W, C, stride, Ho, Wo));
  return adamml_check_launch("dwconv_wgrad_scalar");
}
