// Depthwise 3x3 backward on TMA-staged tiles: data gradient AND weight gradient in ONE pass (bf16 NHWC).
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

template <int TH, int CB>
__global__ void __launch_bounds__(DWB_THREADS, 2)
dw_bwd_s1_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                 const float* __restrict__ w, bf16* __restrict__ dx, float* __restrict__ dWg,
                 const __grid_constant__ DwGeom g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint8_t* stages = smem + 128;
  const int stage_bytes = g.dy_bytes + g.x_bytes;

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
    mbar_wait(&full[stage], (uint32_t)((it_ >> 1) & 1));
    if (active) {
      const int wt = s % g.tiles_w, ht = (s / g.tiles_w) % g.tiles_h, ti = s / (g.tiles_w * g.tiles_h);
      const bf16* dyS = reinterpret_cast<const bf16*>(stages + stage * stage_bytes) +
                        ((bi * (TH + 2)) * (g.TW + 2) + 2 * jp) * CB + 2 * cp;
      const bf16* xS = reinterpret_cast<const bf16*>(stages + stage * stage_bytes + g.dy_bytes) +
                       ((bi * TH) * g.TW + 2 * jp) * CB + 2 * cp;
      const int img = ti * g.BI + bi, h0 = ht * TH, wc = wt * g.TW + 2 * jp;
      const bool ok0 = img < g.IMGS && wc < g.W, ok1 = img < g.IMGS && wc + 1 < g.W;
      bf16* dxp = dx + (((long long)img * g.H + h0) * g.W + wc) * g.C + c0 + 2 * cp;
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
        if (h0 + h < g.H) {
          if (ok0) st_bf2(dxp + (long long)h * g.W * g.C, a0);
          if (ok1) st_bf2(dxp + (long long)h * g.W * g.C + g.C, a1);
        }
      }
    }
    __syncthreads();  // the stage may be refilled by the next iteration's TMA
  }
  reduce_dw(dW, reinterpret_cast<float*>(stages), g, cp, pos, active, c0, dWg);
}

// stride 2: thread = (channel pair, quad column n, image); tile = TH quad rows x TW quad columns; dy tile has one
// extra row / column (dy[m+1], dy[n+1]), x tile is [2 TH][2 TW]
template <int TH, int CB>
__global__ void __launch_bounds__(DWB_THREADS, 2)
dw_bwd_s2_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                 const float* __restrict__ w, bf16* __restrict__ dx, float* __restrict__ dWg,
                 const __grid_constant__ DwGeom g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint8_t* stages = smem + 128;
  const int stage_bytes = g.dy_bytes + g.x_bytes;

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
    mbar_wait(&full[stage], (uint32_t)((it_ >> 1) & 1));
    if (active) {
      const int wt = s % g.tiles_w, ht = (s / g.tiles_w) % g.tiles_h, ti = s / (g.tiles_w * g.tiles_h);
      const bf16* dyS = reinterpret_cast<const bf16*>(stages + stage * stage_bytes) +
                        ((bi * (TH + 1)) * (g.TW + 1) + n) * CB + 2 * cp;
      const bf16* xS = reinterpret_cast<const bf16*>(stages + stage * stage_bytes + g.dy_bytes) +
                       ((bi * 2 * TH) * 2 * g.TW + 2 * n) * CB + 2 * cp;
      const int img = ti * g.BI + bi, hq = ht * TH, wc = (wt * g.TW + n) * 2;
      const bool ok0 = img < g.IMGS && wc < g.W, ok1 = img < g.IMGS && wc + 1 < g.W;
      bf16* dxp = dx + (((long long)img * g.H + 2 * hq) * g.W + wc) * g.C + c0 + 2 * cp;
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
        bf16* o = dxp + (long long)(2 * m) * g.W * g.C;
        if (h < g.H) {
          if (ok0) st_bf2(o, a00);
          if (ok1) st_bf2(o + g.C, a01);
        }
        if (h + 1 < g.H) {
          if (ok0) st_bf2(o + (long long)g.W * g.C, a10);
          if (ok1) st_bf2(o + (long long)g.W * g.C + g.C, a11);
        }
      }
    }
    __syncthreads();
  }
  reduce_dw(dW, reinterpret_cast<float*>(stages), g, cp, pos, active, c0, dWg);
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
int adamml_dwconv_bwd(const void* x, const void* dy, const float* w, void* dx, float* dw, int IMGS, int H, int W, int C,
                      int stride, int Ho, int Wo, cudaStream_t stream) {
  ADAMML_REQUIRE(adamml_dwconv_bwd_supported(IMGS, H, W, C, stride), "dwconv_bwd: unsupported shape (C %% 16, stride)");
  ADAMML_REQUIRE(Ho == (H + 2 - 3) / stride + 1 && Wo == (W + 2 - 3) / stride + 1, "dwconv_bwd: bad Ho/Wo");
  ADAMML_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)dx % 16) == 0,
                 "dwconv_bwd: tensors must be 16-byte aligned");
  DwGeom g;
  memset(&g, 0, sizeof(g));
  g.IMGS = IMGS; g.H = H; g.W = W; g.C = C; g.Ho = Ho; g.Wo = Wo;
  g.CB = (C % 64 == 0) ? 64 : ((C % 48 == 0) ? 48 : ((C % 32 == 0) ? 32 : 16));
  g.CP = g.CB / 2;
  g.chunks = C / g.CB;
  const int max_pos = DWB_THREADS / g.CP;
  // positions per tile = column units x images: fewest tiles wins, then the narrower halo
  const int units = stride == 1 ? (W + 1) / 2 : (W + 1) / 2;  // column pairs (s1) | quad columns (s2)
  long long best = -1;
  int best_u = 1, best_bi = 1;
  for (int u = 1; u <= max_pos && u <= 64; ++u) {
    int bi = max_pos / u;
    if (bi > IMGS) bi = IMGS;
    if (bi < 1) bi = 1;
    if (bi > 16) bi = 16;
    const long long t = (long long)((units + u - 1) / u) * ((IMGS + bi - 1) / bi);
    if (best < 0 || t < best || (t == best && u > best_u)) { best = t; best_u = u; best_bi = bi; }
  }
  g.BI = best_bi;
  g.npos = best_u * best_bi;
  g.tiles_i = (IMGS + g.BI - 1) / g.BI;
  CUtensorMap tmDy, tmX;
  int rc;
  cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * 9, stream);
  const int sms2 = num_sms() * 2;
  if (stride == 1) {
    const int TH = (H % 8 != 0 && H % 5 == 0) ? 5 : 8;
    g.TW = best_u * 2;
    g.tiles_w = (W + g.TW - 1) / g.TW;
    g.tiles_h = (H + TH - 1) / TH;
    g.dy_bytes = pad128(g.BI * (TH + 2) * (g.TW + 2) * g.CB * 2);
    g.x_bytes = pad128(g.BI * TH * g.TW * g.CB * 2);
    rc = make_dw_map(&tmDy, dy, C, Wo, Ho, IMGS, g.CB, g.TW + 2, TH + 2, g.BI);
    if (rc) return rc;
    rc = make_dw_map(&tmX, x, C, W, H, IMGS, g.CB, g.TW, TH, g.BI);
    if (rc) return rc;
    const int smem = 2 * (g.dy_bytes + g.x_bytes) + 256;
    const long long sp = (long long)g.tiles_w * g.tiles_h * g.tiles_i;
    long long per = sms2 / g.chunks > 0 ? sms2 / g.chunks : 1;
    if (per > sp) per = sp;
    const int grid = (int)per * g.chunks;
    typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const float*, bf16*, float*, const DwGeom);
    KernFn kern = nullptr;
#define DWB_PICK(THV, CBV) if (TH == THV && g.CB == CBV) kern = dw_bwd_s1_kernel<THV, CBV>;
    DWB_PICK(8, 64) DWB_PICK(8, 48) DWB_PICK(8, 32) DWB_PICK(8, 16)
    DWB_PICK(5, 64) DWB_PICK(5, 48) DWB_PICK(5, 32) DWB_PICK(5, 16)
#undef DWB_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) { adamml_set_error("dwconv_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ADAMML_ERR_CUDA; }
    ADAMML_REQUIRE(smem <= 110 * 1024, "dwconv_bwd: tile does not fit shared memory (%d bytes)", smem);
    kern<<<grid, DWB_THREADS, smem, stream>>>(tmDy, tmX, w, (bf16*)dx, dw, g);
  } else {
    const int Hq = (H + 1) / 2;
    const int TH = (Hq % 4 != 0 && Hq % 5 == 0) ? 5 : 4;
    g.TW = best_u;
    g.tiles_w = (units + g.TW - 1) / g.TW;
    g.tiles_h = (Hq + TH - 1) / TH;
    g.dy_bytes = pad128(g.BI * (TH + 1) * (g.TW + 1) * g.CB * 2);
    g.x_bytes = pad128(g.BI * 4 * TH * g.TW * g.CB * 2);
    rc = make_dw_map(&tmDy, dy, C, Wo, Ho, IMGS, g.CB, g.TW + 1, TH + 1, g.BI);
    if (rc) return rc;
    rc = make_dw_map(&tmX, x, C, W, H, IMGS, g.CB, 2 * g.TW, 2 * TH, g.BI);
    if (rc) return rc;
    const int smem = 2 * (g.dy_bytes + g.x_bytes) + 256;
    const long long sp = (long long)g.tiles_w * g.tiles_h * g.tiles_i;
    long long per = sms2 / g.chunks > 0 ? sms2 / g.chunks : 1;
    if (per > sp) per = sp;
    const int grid = (int)per * g.chunks;
    typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const float*, bf16*, float*, const DwGeom);
    KernFn kern = nullptr;
#define DWB_PICK(THV, CBV) if (TH == THV && g.CB == CBV) kern = dw_bwd_s2_kernel<THV, CBV>;
    DWB_PICK(4, 64) DWB_PICK(4, 48) DWB_PICK(4, 32) DWB_PICK(4, 16)
    DWB_PICK(5, 64) DWB_PICK(5, 48) DWB_PICK(5, 32) DWB_PICK(5, 16)
#undef DWB_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) { adamml_set_error("dwconv_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ADAMML_ERR_CUDA; }
    ADAMML_REQUIRE(smem <= 110 * 1024, "dwconv_bwd: tile does not fit shared memory (%d bytes)", smem);
    kern<<<grid, DWB_THREADS, smem, stream>>>(tmDy, tmX, w, (bf16*)dx, dw, g);
  }
  return adamml_check_launch("dwconv_bwd");
}

}  // extern "C"
