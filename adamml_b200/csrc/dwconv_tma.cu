// Depthwise 3x3 on TMA-staged tiles: (1) backward -- data gradient AND weight gradient in ONE pass (bf16 NHWC);
// (2) x2 training forward, stride 1, with the train-mode BatchNorm statistics of its output fused (dw_fwd_x2_s1_kernel).
//
// Reference call sites: autograd of the 3x3 `groups=hidden_dim` convolutions of every InvertedResidual
// (models/sound_mobilenet_v2.py:58, models/policy_net.py:66,80).  The two gradients read the same dy neighbourhood:
//
//   stride 1:  dx[h][w]      = sum_{r,s} dy[h+1-r][w+1-s] * w[r][s]
//              dW[r][s]     += x[h][w] * dy[h+1-r][w+1-s]
//   stride 2:  the same sums restricted to (h+1-r), (w+1-s) even, dy index halved (one 2x2 input quad <-> dy[m..m+1][n..n+1])
//
// so one kernel streams dy + x once and writes dx: 3 tensor passes instead of the 4 of separate dgrad + wgrad
// launches.  HBM-bound by construction (algorithmic bytes 2 * (|x| + |dx| + |dy|)); what the register-window kernels
// of dwconv.cu lacked was memory-level parallelism (a load -> unpack -> FMA chain per row, ~20 % issue utilisation), so
// here the operands arrive through TMA:
//   * persistent CTAs, 2-stage ring: ONE elected thread issues two 4D boxes per tile ({CB channels, TW(+2), TH(+2), BI
//     images} of dy with its halo and of x) while all 256 threads compute the previous tile out of shared memory;
//   * zero padding = TMA out-of-bounds fill (negative / overhanging box coordinates), no boundary branches on loads;
//   * a thread owns ONE channel pair (packed fma.rn.f32x2 everywhere) and a 2-column strip (stride 1) or one 2x2 quad
//     column (stride 2) and walks the tile rows with a rolling dy window in registers; conflict-free 4-byte LDS
//     (consecutive threads = consecutive channel pairs);
//   * a CTA keeps one channel chunk for its whole life, so the 9 x 2 weight-gradient accumulators stay in registers
//     across tiles; one shared-memory reduction and 9 x CB fp32 atomics per CTA at the end.
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int DWB_THREADS = 256;

struct DwGeom {
  int IMGS, H, W, C, Ho, Wo;       // x / dx: [IMGS,H,W,C]; dy: [IMGS,Ho,Wo,C]
  int CB, CP;                      // channels per tile, channel pairs (= threads per position)
  int TW, BI;                      // stride 1: dx columns per tile (even); stride 2: quad columns per tile; images per tile
  int npos;                        // active positions per tile: stride 1 (TW/2)*BI, stride 2 TW*BI
  int tiles_w, tiles_h, tiles_i, chunks;
  int dy_bytes, x_bytes;           // per stage, padded to 128
};

__device__ __forceinline__ float2 ld_bf2(const bf16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st_bf2(bf16* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
}

// Optional fused BatchNorm-backward REDUCTION of the layer that produced x (the expand / first conv in front of the
// depthwise conv: x = act(bn(z)), no residual): that layer's backward needs gm = dx * act'(x) and, per BN group and
// channel, sum gm and sum gm * xhat.  The kernel already holds dx and x at every position, and wherever act' = 1 the
// saved output IS the normalised value (x = z * scale + shift), so it stores gm instead of dx and accumulates
// (sum gm, sum gm * x); adamml_bn_sums_from_out (bn.cu) turns those into (sum gm, sum gm * xhat).  The separate
// bn_bwd_reduce pass over (dx, z) disappears.  A tile never straddles two BN groups (BI divides imgs_per_group).
struct PreReduce {
  double* sums;  // raw [G][C][2] = (sum gm, sum gm * x), zeroed by the launcher; nullptr = off
  int imgs_per_group;
  int act;
};
constexpr int PRE_FLUSH_TILES = 32;
constexpr int DW_RED_BYTES = 4096;  // [positions][2][CB] fp32 partials of a statistics flush

// ACT (compile time): ADAMML_ACT_RELU | ADAMML_ACT_RELU6 (the mask is evaluated on the stored output, act_pass)
template <int ACT>
__device__ __forceinline__ float2 pre_mask(float2 a, float2 x, float2& A, float2& B) {
  const bool p0 = ACT == ADAMML_ACT_RELU6 ? (x.x > 0.f && x.x < 6.f) : (x.x > 0.f);
  const bool p1 = ACT == ADAMML_ACT_RELU6 ? (x.y > 0.f && x.y < 6.f) : (x.y > 0.f);
  const float2 gm = make_float2(p0 ? a.x : 0.f, p1 ? a.y : 0.f);
  A = __ffma2_rn(gm, make_float2(1.f, 1.f), A);
  B = __ffma2_rn(gm, x, B);
  return gm;
}
// collective of the CTA: per-thread (A, B) partials -> shared memory -> one fp64 atomic per channel and value
template <int CB>
__device__ __forceinline__ void pre_flush(float2& A, float2& B, float* red, int npos, int cp, int pos, bool active,
                                          double* __restrict__ sums, long long grp, int C, int c0) {
  if (active) {
    *reinterpret_cast<float2*>(red + (pos * 2 + 0) * CB + 2 * cp) = A;
    *reinterpret_cast<float2*>(red + (pos * 2 + 1) * CB + 2 * cp) = B;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * CB; i += blockDim.x) {
    const int k = i / CB, c = i - k * CB;
    double a = 0.0;
    for (int p = 0; p < npos; ++p) a += (double)red[(p * 2 + k) * CB + c];
    atomicAdd(sums + (grp * C + c0 + c) * 2 + k, a);
  }
  __syncthreads();
  A = make_float2(0.f, 0.f);
  B = make_float2(0.f, 0.f);
}

// shared-memory reduction of the per-thread weight-gradient accumulators over the positions of the CTA, then one
// fp32 atomic per (tap, channel)
__device__ __forceinline__ void reduce_dw(const float2 (&dW)[9], float* red, const DwGeom& g, int cp, int pos,
                                          bool active, int c0, float* __restrict__ dWg) {
  __syncthreads();  // every thread is done with the stage buffers
  if (active) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
      *reinterpret_cast<float2*>(red + (pos * 9 + t) * g.CB + 2 * cp) = dW[t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * g.CB; i += DWB_THREADS) {
    float s = 0.f;
    for (int p = 0; p < g.npos; ++p) s += red[p * 9 * g.CB + i];
    const int t = i / g.CB, c = i - t * g.CB;
    atomicAdd(dWg + (long long)t * g.C + c0 + c, s);
  }
}

template <int TH, int CB, int PRE>
__global__ void __launch_bounds__(DWB_THREADS, 2)
dw_bwd_s1_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                 const float* __restrict__ w, bf16* __restrict__ dx, float* __restrict__ dWg,
                 const __grid_constant__ DwGeom g, const PreReduce pre) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  float* red = reinterpret_cast<float*>(smem + 128);
  uint8_t* stages = smem + 128 + DW_RED_BYTES;
  const int stage_bytes = g.dy_bytes + g.x_bytes;
  float2 pA = make_float2(0.f, 0.f), pB = make_float2(0.f, 0.f);
  long long pre_g = -1;
  int pre_tiles = 0;

  const int cp = threadIdx.x % (CB / 2), pos = threadIdx.x / (CB / 2);
  const bool active = pos < g.npos;
  const int half_tw = g.TW >> 1;
  const int jp = pos % half_tw, bi = pos / half_tw;
  const int chunk = blockIdx.x % g.chunks;
  const int c0 = chunk * CB;
  const int cta = blockIdx.x / g.chunks, ncta = gridDim.x / g.chunks;
  const int sp_tiles = g.tiles_w * g.tiles_h * g.tiles_i;

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDy)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
  }
  __syncthreads();

  auto issue = [&](int s, int stage) {  // one thread
    const int wt = s % g.tiles_w, ht = (s / g.tiles_w) % g.tiles_h, it = s / (g.tiles_w * g.tiles_h);
    uint8_t* dst = stages + stage * stage_bytes;
    mbar_expect_tx(&full[stage], (uint32_t)(g.BI * ((TH + 2) * (g.TW + 2) + TH * g.TW) * CB * 2));
    tma_load_4d(dst, &tmDy, &full[stage], c0, wt * g.TW - 1, ht * TH - 1, it * g.BI);
    tma_load_4d(dst + g.dy_bytes, &tmX, &full[stage], c0, wt * g.TW, ht * TH, it * g.BI);
  };

  float2 wr[9], dW[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    wr[t] = active ? *reinterpret_cast<const float2*>(w + (long long)t * g.C + c0 + 2 * cp) : make_float2(0.f, 0.f);
    dW[t] = make_float2(0.f, 0.f);
  }

  if (threadIdx.x == 0 && cta < sp_tiles) issue(cta, 0);
  const int dy_pitch = (g.TW + 2) * CB, x_pitch = g.TW * CB;  // elements per tile row
  int it_ = 0;
  for (int s = cta; s < sp_tiles; s += ncta, ++it_) {
    const int stage = it_ & 1;
    if (threadIdx.x == 0 && s + ncta < sp_tiles) issue(s + ncta, stage ^ 1);
    if (PRE) {
      const long long tg = ((s / (g.tiles_w * g.tiles_h)) * g.BI) / pre.imgs_per_group;
      if (tg != pre_g || pre_tiles == PRE_FLUSH_TILES) {
        if (pre_g >= 0) pre_flush<CB>(pA, pB, red, g.npos, cp, pos, active, pre.sums, pre_g, g.C, c0);
        pre_g = tg;
        pre_tiles = 0;
      }
      ++pre_tiles;
    }
    mbar_wait(&full[stage], (uint32_t)((it_ >> 1) & 1));
    if (active) {
      const int wt = s % g.tiles_w, ht = (s / g.tiles_w) % g.tiles_h, ti = s / (g.tiles_w * g.tiles_h);
      const bf16* dyS = reinterpret_cast<const bf16*>(stages + stage * stage_bytes) +
                        ((bi * (TH + 2)) * (g.TW + 2) + 2 * jp) * CB + 2 * cp;
      const bf16* xS = reinterpret_cast<const bf16*>(stages + stage * stage_bytes + g.dy_bytes) +
                       ((bi * TH) * g.TW + 2 * jp) * CB + 2 * cp;
      const int img = ti * g.BI + bi, h0 = ht * TH, wc = wt * g.TW + 2 * jp;
      const bool ok0 = img < g.IMGS && wc < g.W, ok1 = img < g.IMGS && wc + 1 < g.W;
      bf16* dxr = dx + (((long long)img * g.H + h0) * g.W + wc) * g.C + c0 + 2 * cp;
      const long long row_stride = (long long)g.W * g.C;
      float2 D[3][4];  // rolling window: dy rows h-1, h, h+1 (mod 3) x columns wc-1 .. wc+2
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        D[0][j] = ld_bf2(dyS + j * CB);
        D[1][j] = ld_bf2(dyS + dy_pitch + j * CB);
      }
#pragma unroll
      for (int h = 0; h < TH; ++h) {
        float2(&Dm)[4] = D[h % 3];        // dy row h-1
        float2(&Dc)[4] = D[(h + 1) % 3];  // dy row h
        float2(&Dp)[4] = D[(h + 2) % 3];  // dy row h+1 (loaded now)
#pragma unroll
        for (int j = 0; j < 4; ++j) Dp[j] = ld_bf2(dyS + (h + 2) * dy_pitch + j * CB);
        const float2 x0 = ld_bf2(xS + h * x_pitch), x1 = ld_bf2(xS + h * x_pitch + CB);
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          float2(&Dr)[4] = r == 0 ? Dp : (r == 1 ? Dc : Dm);  // dy row h + 1 - r
#pragma unroll
          for (int s_ = 0; s_ < 3; ++s_) {
            const float2 d0 = Dr[2 - s_], d1 = Dr[3 - s_];  // dy column (wc + o) + 1 - s  ->  window index o + 2 - s
            a0 = __ffma2_rn(d0, wr[r * 3 + s_], a0);
            a1 = __ffma2_rn(d1, wr[r * 3 + s_], a1);
            dW[r * 3 + s_] = __ffma2_rn(x0, d0, dW[r * 3 + s_]);
            dW[r * 3 + s_] = __ffma2_rn(x1, d1, dW[r * 3 + s_]);
          }
        }
        // (x is zero outside the image -- TMA fill -- so the activation mask already drops those positions from the
        // fused sums; only the stores are predicated)
        const float2 o0 = PRE ? pre_mask<PRE>(a0, x0, pA, pB) : a0;
        const float2 o1 = PRE ? pre_mask<PRE>(a1, x1, pA, pB) : a1;
        const bool rok = h0 + h < g.H;
        if (rok && ok0) st_bf2(dxr, o0);
        if (rok && ok1) st_bf2(dxr + g.C, o1);
        dxr += row_stride;
      }
    }
    __syncthreads();  // the stage may be refilled by the next iteration's TMA
  }
  if (PRE && pre_g >= 0) pre_flush<CB>(pA, pB, red, g.npos, cp, pos, active, pre.sums, pre_g, g.C, c0);
  reduce_dw(dW, reinterpret_cast<float*>(stages), g, cp, pos, active, c0, dWg);
}

// stride 2: thread = (channel pair, quad column n, image); tile = TH quad rows x TW quad columns; dy tile has one
// extra row / column (dy[m+1], dy[n+1]), x tile is [2 TH][2 TW]
template <int TH, int CB, int PRE>
__global__ void __launch_bounds__(DWB_THREADS, 2)
dw_bwd_s2_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                 const float* __restrict__ w, bf16* __restrict__ dx, float* __restrict__ dWg,
                 const __grid_constant__ DwGeom g, const PreReduce pre) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  float* red = reinterpret_cast<float*>(smem + 128);
  uint8_t* stages = smem + 128 + DW_RED_BYTES;
  const int stage_bytes = g.dy_bytes + g.x_bytes;
  float2 pA = make_float2(0.f, 0.f), pB = make_float2(0.f, 0.f);
  long long pre_g = -1;
  int pre_tiles = 0;

  const int cp = threadIdx.x % (CB / 2), pos = threadIdx.x / (CB / 2);
  const bool active = pos < g.npos;
  const int n = pos % g.TW, bi = pos / g.TW;
  const int chunk = blockIdx.x % g.chunks;
  const int c0 = chunk * CB;
  const int cta = blockIdx.x / g.chunks, ncta = gridDim.x / g.chunks;
  const int sp_tiles = g.tiles_w * g.tiles_h * g.tiles_i;

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDy)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
  }
  __syncthreads();

  auto issue = [&](int s, int stage) {
    const int wt = s % g.tiles_w, ht = (s / g.tiles_w) % g.tiles_h, it = s / (g.tiles_w * g.tiles_h);
    uint8_t* dst = stages + stage * stage_bytes;
    mbar_expect_tx(&full[stage], (uint32_t)(g.BI * ((TH + 1) * (g.TW + 1) + 4 * TH * g.TW) * CB * 2));
    tma_load_4d(dst, &tmDy, &full[stage], c0, wt * g.TW, ht * TH, it * g.BI);
    tma_load_4d(dst + g.dy_bytes, &tmX, &full[stage], c0, wt * g.TW * 2, ht * TH * 2, it * g.BI);
  };

  float2 wr[9], dW[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    wr[t] = active ? *reinterpret_cast<const float2*>(w + (long long)t * g.C + c0 + 2 * cp) : make_float2(0.f, 0.f);
    dW[t] = make_float2(0.f, 0.f);
  }

  if (threadIdx.x == 0 && cta < sp_tiles) issue(cta, 0);
  const int dy_pitch = (g.TW + 1) * CB, x_pitch = 2 * g.TW * CB;
  int it_ = 0;
  for (int s = cta; s < sp_tiles; s += ncta, ++it_) {
    const int stage = it_ & 1;
    if (threadIdx.x == 0 && s + ncta < sp_tiles) issue(s + ncta, stage ^ 1);
    if (PRE) {
      const long long tg = ((s / (g.tiles_w * g.tiles_h)) * g.BI) / pre.imgs_per_group;
      if (tg != pre_g || pre_tiles == PRE_FLUSH_TILES) {
        if (pre_g >= 0) pre_flush<CB>(pA, pB, red, g.npos, cp, pos, active, pre.sums, pre_g, g.C, c0);
        pre_g = tg;
        pre_tiles = 0;
      }
      ++pre_tiles;
    }
    mbar_wait(&full[stage], (uint32_t)((it_ >> 1) & 1));
    if (active) {
      const int wt = s % g.tiles_w, ht = (s / g.tiles_w) % g.tiles_h, ti = s / (g.tiles_w * g.tiles_h);
      const bf16* dyS = reinterpret_cast<const bf16*>(stages + stage * stage_bytes) +
                        ((bi * (TH + 1)) * (g.TW + 1) + n) * CB + 2 * cp;
      const bf16* xS = reinterpret_cast<const bf16*>(stages + stage * stage_bytes + g.dy_bytes) +
                       ((bi * 2 * TH) * 2 * g.TW + 2 * n) * CB + 2 * cp;
      const int img = ti * g.BI + bi, hq = ht * TH, wc = (wt * g.TW + n) * 2;
      const bool ok0 = img < g.IMGS && wc < g.W, ok1 = img < g.IMGS && wc + 1 < g.W;
      bf16* dxr = dx + (((long long)img * g.H + 2 * hq) * g.W + wc) * g.C + c0 + 2 * cp;
      const long long row_stride = (long long)g.W * g.C;
      float2 D[2][2];  // dy rows m, m+1 (mod 2) x columns n, n+1
      D[0][0] = ld_bf2(dyS);
      D[0][1] = ld_bf2(dyS + CB);
#pragma unroll
      for (int m = 0; m < TH; ++m) {
        float2(&D0)[2] = D[m & 1];
        float2(&D1)[2] = D[(m + 1) & 1];
        D1[0] = ld_bf2(dyS + (m + 1) * dy_pitch);
        D1[1] = ld_bf2(dyS + (m + 1) * dy_pitch + CB);
        const bf16* xr = xS + (2 * m) * x_pitch;
        const float2 x00 = ld_bf2(xr), x01 = ld_bf2(xr + CB);
        const float2 x10 = ld_bf2(xr + x_pitch), x11 = ld_bf2(xr + x_pitch + CB);
        const float2 z = make_float2(0.f, 0.f);
        // (2m, 2n): tap (1,1)
        const float2 a00 = __ffma2_rn(D0[0], wr[4], z);
        dW[4] = __ffma2_rn(x00, D0[0], dW[4]);
        // (2m, 2n+1): taps (1,0) <- dy[m][n+1], (1,2) <- dy[m][n]
        float2 a01 = __ffma2_rn(D0[1], wr[3], z);
        a01 = __ffma2_rn(D0[0], wr[5], a01);
        dW[3] = __ffma2_rn(x01, D0[1], dW[3]);
        dW[5] = __ffma2_rn(x01, D0[0], dW[5]);
        // (2m+1, 2n): taps (0,1) <- dy[m+1][n], (2,1) <- dy[m][n]
        float2 a10 = __ffma2_rn(D1[0], wr[1], z);
        a10 = __ffma2_rn(D0[0], wr[7], a10);
        dW[1] = __ffma2_rn(x10, D1[0], dW[1]);
        dW[7] = __ffma2_rn(x10, D0[0], dW[7]);
        // (2m+1, 2n+1): taps (0,0) <- dy[m+1][n+1], (0,2) <- dy[m+1][n], (2,0) <- dy[m][n+1], (2,2) <- dy[m][n]
        float2 a11 = __ffma2_rn(D1[1], wr[0], z);
        a11 = __ffma2_rn(D1[0], wr[2], a11);
        a11 = __ffma2_rn(D0[1], wr[6], a11);
        a11 = __ffma2_rn(D0[0], wr[8], a11);
        dW[0] = __ffma2_rn(x11, D1[1], dW[0]);
        dW[2] = __ffma2_rn(x11, D1[0], dW[2]);
        dW[6] = __ffma2_rn(x11, D0[1], dW[6]);
        dW[8] = __ffma2_rn(x11, D0[0], dW[8]);
        const int h = 2 * (hq + m);
        const float2 o00 = PRE ? pre_mask<PRE>(a00, x00, pA, pB) : a00;
        const float2 o01 = PRE ? pre_mask<PRE>(a01, x01, pA, pB) : a01;
        const float2 o10 = PRE ? pre_mask<PRE>(a10, x10, pA, pB) : a10;
        const float2 o11 = PRE ? pre_mask<PRE>(a11, x11, pA, pB) : a11;
        const bool r0 = h < g.H, r1 = h + 1 < g.H;
        if (r0 && ok0) st_bf2(dxr, o00);
        if (r0 && ok1) st_bf2(dxr + g.C, o01);
        if (r1 && ok0) st_bf2(dxr + row_stride, o10);
        if (r1 && ok1) st_bf2(dxr + row_stride + g.C, o11);
        dxr += 2 * row_stride;
      }
    }
    __syncthreads();
  }
  if (PRE && pre_g >= 0) pre_flush<CB>(pA, pB, red, g.npos, cp, pos, active, pre.sums, pre_g, g.C, c0);
  reduce_dw(dW, reinterpret_cast<float*>(stages), g, cp, pos, active, c0, dWg);
}

// ---- x2 training forward, stride 1, + BatchNorm statistics ---------------------------------------------------------
// z = dwconv3x3(x) on two-plane activations (hi bf16 + lo fp16, common.cuh), written as two planes, and
// sums[g][c] = (sum z, sum z^2) per BN group g = img / imgs_per_group in the same pass (models/sound_mobilenet_v2.py:58,62;
// policy_net.py:66-67,80-81 in train mode): the separate bn_stats_x2 pass over z disappears.  Same tile scheme as the
// backward: both planes of the x tile (with halo) arrive by TMA, a thread owns a channel pair and a 2-column strip,
// joins hi + lo ONCE per loaded value into a rolling fp32 window, 9 packed FMAs per output pair.  A tile never
// straddles two BN groups (BI divides imgs_per_group), so the per-thread fp32 partial sums are flushed (shared-memory
// reduction -> one fp64 atomic per channel and CTA) only when the CTA's group changes or every FWD_FLUSH_TILES tiles.
constexpr int FWD_FLUSH_TILES = 32;

__device__ __forceinline__ float2 ld_x2(const bf16* hi, const __half* lo) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hi));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(lo));
  return make_float2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ void st_x2(bf16* hi, __half* lo, float2 v) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  const float2 hf = __bfloat1622float2(h);
  float rx = v.x - hf.x, ry = v.y - hf.y;
  rx = rx == rx ? rx : 0.f;  // inf - inf: keep the non-finite value in hi only (x2_split)
  ry = ry == ry ? ry : 0.f;
  *reinterpret_cast<__nv_bfloat162*>(hi) = h;
  *reinterpret_cast<__half2*>(lo) = __floats2half2_rn(rx, ry);
}

template <int TH, int CB>
__global__ void __launch_bounds__(DWB_THREADS, 2)
dw_fwd_x2_s1_kernel(const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo,
                    const float* __restrict__ w, bf16* __restrict__ y_hi, __half* __restrict__ y_lo,
                    double* __restrict__ sums, int imgs_per_group, const __grid_constant__ DwGeom g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  float* red = reinterpret_cast<float*>(smem + 128);  // [npos][2][CB] statistics partials
  uint8_t* stages = smem + 128 + 4096;
  const int stage_bytes = 2 * g.dy_bytes;  // (dy_bytes = one halo-tile plane)

  const int cp = threadIdx.x % (CB / 2), pos = threadIdx.x / (CB / 2);
  const bool active = pos < g.npos;
  const int half_tw = g.TW >> 1;
  const int jp = pos % half_tw, bi = pos / half_tw;
  const int chunk = blockIdx.x % g.chunks;
  const int c0 = chunk * CB;
  const int cta = blockIdx.x / g.chunks, ncta = gridDim.x / g.chunks;
  const int sp_tiles = g.tiles_w * g.tiles_h * g.tiles_i;

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmHi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmLo)) : "memory");
  }
  __syncthreads();

  auto issue = [&](int s, int stage) {  // one thread
    const int wt = s % g.tiles_w, ht = (s / g.tiles_w) % g.tiles_h, it = s / (g.tiles_w * g.tiles_h);
    uint8_t* dst = stages + stage * stage_bytes;
    mbar_expect_tx(&full[stage], (uint32_t)(2 * g.BI * (TH + 2) * (g.TW + 2) * CB * 2));
    tma_load_4d(dst, &tmHi, &full[stage], c0, wt * g.TW - 1, ht * TH - 1, it * g.BI);
    tma_load_4d(dst + g.dy_bytes, &tmLo, &full[stage], c0, wt * g.TW - 1, ht * TH - 1, it * g.BI);
  };

  float2 wr[9];
#pragma unroll
  for (int t = 0; t < 9; ++t)
    wr[t] = active ? *reinterpret_cast<const float2*>(w + (long long)t * g.C + c0 + 2 * cp) : make_float2(0.f, 0.f);
  float2 S = make_float2(0.f, 0.f), Q = make_float2(0.f, 0.f);
  int cur_g = -1, since_flush = 0;
  auto flush = [&]() {  // collective of the CTA
    if (active) {
      *reinterpret_cast<float2*>(red + (pos * 2 + 0) * CB + 2 * cp) = S;
      *reinterpret_cast<float2*>(red + (pos * 2 + 1) * CB + 2 * cp) = Q;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CB; i += DWB_THREADS) {
      const int k = i / CB, c = i - k * CB;
      double a = 0.0;
      for (int p = 0; p < g.npos; ++p) a += (double)red[(p * 2 + k) * CB + c];
      atomicAdd(sums + ((long long)cur_g * g.C + c0 + c) * 2 + k, a);
    }
    __syncthreads();
    S = make_float2(0.f, 0.f);
    Q = make_float2(0.f, 0.f);
    since_flush = 0;
  };

  if (threadIdx.x == 0 && cta < sp_tiles) issue(cta, 0);
  const int pitch = (g.TW + 2) * CB;  // elements per tile row
  int it_ = 0;
  for (int s = cta; s < sp_tiles; s += ncta, ++it_) {
    const int stage = it_ & 1;
    if (threadIdx.x == 0 && s + ncta < sp_tiles) issue(s + ncta, stage ^ 1);
    const int wt = s % g.tiles_w, ht = (s / g.tiles_w) % g.tiles_h, ti = s / (g.tiles_w * g.tiles_h);
    if (sums) {
      const int tg = (ti * g.BI) / imgs_per_group;
      if (tg != cur_g || since_flush == FWD_FLUSH_TILES) {
        if (cur_g >= 0) flush();
        cur_g = tg;
      }
      ++since_flush;
    }
    mbar_wait(&full[stage], (uint32_t)((it_ >> 1) & 1));
    if (active) {
      const int off = ((bi * (TH + 2)) * (g.TW + 2) + 2 * jp) * CB + 2 * cp;
      const bf16* hiS = reinterpret_cast<const bf16*>(stages + stage * stage_bytes) + off;
      const __half* loS = reinterpret_cast<const __half*>(stages + stage * stage_bytes + g.dy_bytes) + off;
      const int img = ti * g.BI + bi, h0 = ht * TH, wc = wt * g.TW + 2 * jp;
      const bool ok0 = img < g.IMGS && wc < g.W, ok1 = img < g.IMGS && wc + 1 < g.W;
      const long long o0 = (((long long)img * g.H + h0) * g.W + wc) * g.C + c0 + 2 * cp;
      float2 X[3][4];  // rolling window: x rows h-1, h, h+1 (mod 3) x columns wc-1 .. wc+2
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        X[0][j] = ld_x2(hiS + j * CB, loS + j * CB);
        X[1][j] = ld_x2(hiS + pitch + j * CB, loS + pitch + j * CB);
      }
#pragma unroll
      for (int h = 0; h < TH; ++h) {
        float2(&Xn)[4] = X[(h + 2) % 3];
#pragma unroll
        for (int j = 0; j < 4; ++j) Xn[j] = ld_x2(hiS + (h + 2) * pitch + j * CB, loS + (h + 2) * pitch + j * CB);
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          float2(&Xr)[4] = X[(h + r) % 3];  // x row h + r - 1
#pragma unroll
          for (int s_ = 0; s_ < 3; ++s_) {
            a0 = __ffma2_rn(Xr[s_], wr[r * 3 + s_], a0);      // x column wc + s - 1  ->  window index s
            a1 = __ffma2_rn(Xr[s_ + 1], wr[r * 3 + s_], a1);
          }
        }
        if (h0 + h < g.H) {
          const long long o = o0 + (long long)h * g.W * g.C;
          if (ok0) {
            st_x2(y_hi + o, y_lo + o, a0);
            S = __ffma2_rn(a0, make_float2(1.f, 1.f), S);
            Q = __ffma2_rn(a0, a0, Q);
          }
          if (ok1) {
            st_x2(y_hi + o + g.C, y_lo + o + g.C, a1);
            S = __ffma2_rn(a1, make_float2(1.f, 1.f), S);
            Q = __ffma2_rn(a1, a1, Q);
          }
        }
      }
    }
    __syncthreads();  // the stage may be refilled by the next iteration's TMA
  }
  if (sums && cur_g >= 0) flush();
}


// 4D map over an NHWC bf16 tensor, dims {C, W, H, IMGS}, box {CB, bw, bh, bi}, no swizzle, out-of-bounds = zeros
int make_dw_map(CUtensorMap* map, const void* ptr, int C, int W, int H, int IMGS, int CB, int bw, int bh, int bi) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { adamml_set_error("cuTensorMapEncodeTiled entry point unavailable"); return ADAMML_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)IMGS};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)CB, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bi};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    adamml_set_error("dwconv_bwd: cuTensorMapEncodeTiled failed (%d) C=%d W=%d H=%d I=%d box=%d,%d,%d,%d", (int)r, C, W,
                     H, IMGS, CB, bw, bh, bi);
    return ADAMML_ERR_CUDA;
  }
  return ADAMML_OK;
}

inline int pad128(int v) { return (v + 127) & ~127; }

constexpr int DW_SMEM_LIMIT = 110 * 1024;  // two CTAs per SM

// Positions of a tile = column units (column pairs | quad columns) x images, at most DWB_THREADS / CP of them.  Fewest
// tiles wins (a tile costs one pass of every thread over its rows), then the wider one (narrower halo), among the
// shapes whose two-stage ring fits DW_SMEM_LIMIT.  group_imgs > 0: images per tile divide it (a tile never straddles
// two BatchNorm groups).  smem_of(u, bi) -> dynamic shared memory of the kernel for that tile shape.
template <typename F>
bool pick_dw_tile(int units, int IMGS, int max_pos, int group_imgs, F smem_of, int* out_u, int* out_bi) {
  long long best = -1;
  for (int u = 1; u <= max_pos && u <= 64; ++u) {
    int bi = max_pos / u;
    if (bi > 16) bi = 16;
    if (bi > IMGS) bi = IMGS;
    if (group_imgs > 0) {
      if (bi > group_imgs) bi = group_imgs;
      while (bi > 1 && group_imgs % bi) --bi;
    }
    while (bi > 1 && smem_of(u, bi) > DW_SMEM_LIMIT) --bi;
    if (group_imgs > 0)
      while (bi > 1 && group_imgs % bi) --bi;
    if (bi < 1 || smem_of(u, bi) > DW_SMEM_LIMIT) continue;
    const long long t = (long long)((units + u - 1) / u) * ((IMGS + bi - 1) / bi);
    if (best < 0 || t < best || (t == best && u > *out_u)) { best = t; *out_u = u; *out_bi = bi; }
  }
  return best >= 0;
}

}  // namespace

extern "C" {

/* 1 if adamml_dwconv_bwd handles the shape (else use adamml_dwconv_dgrad + adamml_dwconv_wgrad) */
int adamml_dwconv_bwd_supported(int IMGS, int H, int W, int C, int stride) {
  if (IMGS <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
  if (stride != 1 && stride != 2) return 0;
  if (C % 16) return 0;  // channel chunks of 16 | 32 | 48 | 64 (16-byte TMA rows, even channel pairs)
  return 1;
}

/* Depthwise 3x3 (pad 1) backward, bf16 NHWC: dx = conv_transpose(dy, w) and dw = the weight gradient (fp32 tap-major
 * [9][C], overwritten) from ONE pass over dy and x.  w: fp32 tap-major [9][C] (adamml_pack_weight_dw). */
int adamml_dwconv_bwd(const void* x, const void* dy, const float* w, void* dx, float* dw, double* pre_sums,
                      int pre_imgs_per_group, int pre_act, int IMGS, int H, int W, int C, int stride, int Ho, int Wo,
                      cudaStream_t stream) {
  ADAMML_REQUIRE(adamml_dwconv_bwd_supported(IMGS, H, W, C, stride), "dwconv_bwd: unsupported shape (C %% 16, stride)");
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / stride + 1 && Wo == (W + 2 - 3) / stride + 1, "dwconv_bwd: bad Ho/Wo");
  ADAMML_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)dx % 16) == 0,
                 "dwconv_bwd: tensors must be 16-byte aligned");
  PreReduce pre{pre_sums, pre_imgs_per_group > 0 ? pre_imgs_per_group : IMGS, pre_act};
  if (pre_sums) {
    ADAMML_REQUIRE(pre_act == ADAMML_ACT_RELU || pre_act == ADAMML_ACT_RELU6,
                   "dwconv_bwd: the fused producer reduction needs a ReLU / ReLU6 producer");
    ADAMML_REQUIRE(IMGS % pre.imgs_per_group == 0, "dwconv_bwd: IMGS must be a multiple of pre_imgs_per_group");
    cudaMemsetAsync(pre_sums, 0, sizeof(double) * (size_t)(IMGS / pre.imgs_per_group) * C * 2, stream);
  }
  DwGeom g;
  memset(&g, 0, sizeof(g));
  g.IMGS = IMGS; g.H = H; g.W = W; g.C = C; g.Ho = Ho; g.Wo = Wo;
  g.CB = (C % 64 == 0) ? 64 : ((C % 48 == 0) ? 48 : ((C % 32 == 0) ? 32 : 16));
  g.CP = g.CB / 2;
  g.chunks = C / g.CB;
  const int max_pos = DWB_THREADS / g.CP;
  const int units = (W + 1) / 2;  // column pairs (stride 1) | quad columns (stride 2)
  const int TH1 = (H % 8 != 0 && H % 5 == 0) ? 5 : 8;
  const int Hq_ = (H + 1) / 2;
  const int TH2 = (Hq_ % 4 != 0 && Hq_ % 5 == 0) ? 5 : 4;
  const int CBv = g.CB;
  auto smem_of = [&](int u, int bi) {
    if (stride == 1)
      return 2 * (pad128(bi * (TH1 + 2) * (2 * u + 2) * CBv * 2) + pad128(bi * TH1 * 2 * u * CBv * 2)) + 256 + DW_RED_BYTES;
    return 2 * (pad128(bi * (TH2 + 1) * (u + 1) * CBv * 2) + pad128(bi * 4 * TH2 * u * CBv * 2)) + 256 + DW_RED_BYTES;
  };
  int best_u = 1, best_bi = 1;
  ADAMML_REQUIRE(pick_dw_tile(units, IMGS, max_pos, pre_sums ? pre.imgs_per_group : 0, smem_of, &best_u, &best_bi),
                 "dwconv_bwd: no tile shape fits");
  g.BI = best_bi;
  g.npos = best_u * best_bi;
  g.tiles_i = (IMGS + g.BI - 1) / g.BI;
  CUtensorMap tmDy, tmX;
  int rc;
  cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * 9, stream);
  const int sms2 = num_sms() * 2;
  if (stride == 1) {
    const int TH = TH1;
    g.TW = best_u * 2;
    g.tiles_w = (W + g.TW - 1) / g.TW;
    g.tiles_h = (H + TH - 1) / TH;
    g.dy_bytes = pad128(g.BI * (TH + 2) * (g.TW + 2) * g.CB * 2);
    g.x_bytes = pad128(g.BI * TH * g.TW * g.CB * 2);
    rc = make_dw_map(&tmDy, dy, C, Wo, Ho, IMGS, g.CB, g.TW + 2, TH + 2, g.BI);
    if (rc) return rc;
    rc = make_dw_map(&tmX, x, C, W, H, IMGS, g.CB, g.TW, TH, g.BI);
    if (rc) return rc;
    const int smem = 2 * (g.dy_bytes + g.x_bytes) + 256 + DW_RED_BYTES;
    const long long sp = (long long)g.tiles_w * g.tiles_h * g.tiles_i;
    long long per = sms2 / g.chunks > 0 ? sms2 / g.chunks : 1;
    if (per > sp) per = sp;
    const int grid = (int)per * g.chunks;
    typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const float*, bf16*, float*, const DwGeom,
                           const PreReduce);
    KernFn kern = nullptr;
#define DWB_PICK(THV, CBV)                                                                              \
  if (TH == THV && g.CB == CBV)                                                                         \
    kern = !pre_sums ? dw_bwd_s1_kernel<THV, CBV, 0>                                                    \
                     : (pre_act == ADAMML_ACT_RELU6 ? dw_bwd_s1_kernel<THV, CBV, ADAMML_ACT_RELU6>     \
                                                    : dw_bwd_s1_kernel<THV, CBV, ADAMML_ACT_RELU>);
    DWB_PICK(8, 64) DWB_PICK(8, 48) DWB_PICK(8, 32) DWB_PICK(8, 16)
    DWB_PICK(5, 64) DWB_PICK(5, 48) DWB_PICK(5, 32) DWB_PICK(5, 16)
#undef DWB_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_LIMIT);
    if (e != cudaSuccess) { adamml_set_error("dwconv_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ADAMML_ERR_CUDA; }
    ADAMML_REQUIRE(smem <= DW_SMEM_LIMIT, "dwconv_bwd: tile does not fit shared memory (%d bytes)", smem);
    kern<<<grid, DWB_THREADS, smem, stream>>>(tmDy, tmX, w, (bf16*)dx, dw, g, pre);
  } else {
    const int Hq = Hq_;
    const int TH = TH2;
    g.TW = best_u;
    g.tiles_w = (units + g.TW - 1) / g.TW;
    g.tiles_h = (Hq + TH - 1) / TH;
    g.dy_bytes = pad128(g.BI * (TH + 1) * (g.TW + 1) * g.CB * 2);
    g.x_bytes = pad128(g.BI * 4 * TH * g.TW * g.CB * 2);
    rc = make_dw_map(&tmDy, dy, C, Wo, Ho, IMGS, g.CB, g.TW + 1, TH + 1, g.BI);
    if (rc) return rc;
    rc = make_dw_map(&tmX, x, C, W, H, IMGS, g.CB, 2 * g.TW, 2 * TH, g.BI);
    if (rc) return rc;
    const int smem = 2 * (g.dy_bytes + g.x_bytes) + 256 + DW_RED_BYTES;
    const long long sp = (long long)g.tiles_w * g.tiles_h * g.tiles_i;
    long long per = sms2 / g.chunks > 0 ? sms2 / g.chunks : 1;
    if (per > sp) per = sp;
    const int grid = (int)per * g.chunks;
    typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const float*, bf16*, float*, const DwGeom,
                           const PreReduce);
    KernFn kern = nullptr;
#define DWB_PICK(THV, CBV)                                                                              \
  if (TH == THV && g.CB == CBV)                                                                         \
    kern = !pre_sums ? dw_bwd_s2_kernel<THV, CBV, 0>                                                    \
                     : (pre_act == ADAMML_ACT_RELU6 ? dw_bwd_s2_kernel<THV, CBV, ADAMML_ACT_RELU6>     \
                                                    : dw_bwd_s2_kernel<THV, CBV, ADAMML_ACT_RELU>);
    DWB_PICK(4, 64) DWB_PICK(4, 48) DWB_PICK(4, 32) DWB_PICK(4, 16)
    DWB_PICK(5, 64) DWB_PICK(5, 48) DWB_PICK(5, 32) DWB_PICK(5, 16)
#undef DWB_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_LIMIT);
    if (e != cudaSuccess) { adamml_set_error("dwconv_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ADAMML_ERR_CUDA; }
    ADAMML_REQUIRE(smem <= DW_SMEM_LIMIT, "dwconv_bwd: tile does not fit shared memory (%d bytes)", smem);
    kern<<<grid, DWB_THREADS, smem, stream>>>(tmDy, tmX, w, (bf16*)dx, dw, g, pre);
  }
  return adamml_check_launch("dwconv_bwd");
}

/* x2 training forward of a depthwise 3x3 / stride 1 / pad 1 conv on TMA tiles with the BatchNorm statistics of its
 * output fused: z (two planes) = dwconv(x), sums[G][C][2] (double; overwritten; may be NULL) = per-group (sum, sum of
 * squares) of z, group = img / imgs_per_group.  adamml_dwconv_fwd_stats_x2_supported -> 1 when the shape is handled. */
int adamml_dwconv_fwd_stats_x2_supported(int IMGS, int H, int W, int C, int stride) {
  return stride == 1 && adamml_dwconv_bwd_supported(IMGS, H, W, C, stride);
}

int adamml_dwconv_fwd_stats_x2(const void* x_hi, const void* x_lo, const float* w, void* y_hi, void* y_lo,
                               double* sums, int IMGS, int H, int W, int C, int imgs_per_group, cudaStream_t stream) {
  ADAMML_REQUIRE(adamml_dwconv_fwd_stats_x2_supported(IMGS, H, W, C, 1), "dwconv_fwd_stats_x2: unsupported shape");
  ADAMML_REQUIRE(((uintptr_t)x_hi % 16) == 0 && ((uintptr_t)x_lo % 16) == 0 && ((uintptr_t)y_hi % 16) == 0 &&
                     ((uintptr_t)y_lo % 16) == 0, "dwconv_fwd_stats_x2: planes must be 16-byte aligned");
  if (imgs_per_group <= 0) imgs_per_group = IMGS;
  ADAMML_REQUIRE(!sums || IMGS % imgs_per_group == 0, "dwconv_fwd_stats_x2: IMGS must be a multiple of imgs_per_group");
  DwGeom g;
  memset(&g, 0, sizeof(g));
  g.IMGS = IMGS; g.H = H; g.W = W; g.C = C; g.Ho = H; g.Wo = W;
  g.CB = (C % 64 == 0) ? 64 : ((C % 48 == 0) ? 48 : ((C % 32 == 0) ? 32 : 16));
  g.CP = g.CB / 2;
  g.chunks = C / g.CB;
  const int max_pos = DWB_THREADS / g.CP;
  const int units = (W + 1) / 2;
  const int TH = (H % 8 != 0 && H % 5 == 0) ? 5 : 8;
  const int CBv = g.CB;
  auto smem_of = [&](int u, int bi) { return 4 * pad128(bi * (TH + 2) * (2 * u + 2) * CBv * 2) + 4096 + 256; };
  int best_u = 1, best_bi = 1;
  ADAMML_REQUIRE(pick_dw_tile(units, IMGS, max_pos, imgs_per_group, smem_of, &best_u, &best_bi),
                 "dwconv_fwd_stats_x2: no tile shape fits");
  g.BI = best_bi;
  g.npos = best_u * best_bi;
  g.tiles_i = (IMGS + g.BI - 1) / g.BI;
  g.TW = best_u * 2;
  g.tiles_w = (W + g.TW - 1) / g.TW;
  g.tiles_h = (H + TH - 1) / TH;
  g.dy_bytes = pad128(g.BI * (TH + 2) * (g.TW + 2) * g.CB * 2);
  g.x_bytes = 0;
  CUtensorMap tmHi, tmLo;
  int rc = make_dw_map(&tmHi, x_hi, C, W, H, IMGS, g.CB, g.TW + 2, TH + 2, g.BI);
  if (rc) return rc;
  rc = make_dw_map(&tmLo, x_lo, C, W, H, IMGS, g.CB, g.TW + 2, TH + 2, g.BI);
  if (rc) return rc;
  if (sums) cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)(IMGS / imgs_per_group) * C * 2, stream);
  const int smem = 2 * (2 * g.dy_bytes) + 4096 + 256;
  ADAMML_REQUIRE(smem <= DW_SMEM_LIMIT, "dwconv_fwd_stats_x2: tile does not fit shared memory (%d bytes)", smem);
  const long long sp = (long long)g.tiles_w * g.tiles_h * g.tiles_i;
  const int sms2 = num_sms() * 2;
  long long per = sms2 / g.chunks > 0 ? sms2 / g.chunks : 1;
  if (per > sp) per = sp;
  const int grid = (int)per * g.chunks;
  typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const float*, bf16*, __half*, double*, int, const DwGeom);
  KernFn kern = nullptr;
#define DWF_PICK(THV, CBV) if (TH == THV && g.CB == CBV) kern = dw_fwd_x2_s1_kernel<THV, CBV>;
  DWF_PICK(8, 64) DWF_PICK(8, 48) DWF_PICK(8, 32) DWF_PICK(8, 16)
  DWF_PICK(5, 64) DWF_PICK(5, 48) DWF_PICK(5, 32) DWF_PICK(5, 16)
#undef DWF_PICK
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_LIMIT);
  if (e != cudaSuccess) { adamml_set_error("dwconv_fwd_stats_x2: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ADAMML_ERR_CUDA; }
  kern<<<grid, DWB_THREADS, smem, stream>>>(tmHi, tmLo, w, (bf16*)y_hi, (__half*)y_lo, sums, imgs_per_group, g);
  return adamml_check_launch("dwconv_fwd_stats_x2");
}

}  // extern "C"
