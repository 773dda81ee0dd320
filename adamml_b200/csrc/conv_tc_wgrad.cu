// tcgen05 weight-gradient kernel: dW[co][r][s][ci] = sum over output pixels of dy[p][co] * x[p (+) (r,s)][ci].
//
// Reference call sites: autograd of every dense nn.Conv2d on the AdaMML path (models/resnet.py:35-43,
// models/sound_mobilenet_v2.py:55,61, models/policy_net.py:49,76,84) — in the reference this is cuDNN's
// wgrad; here it is an implicit GEMM whose reduction dimension is the PIXEL axis:
//
//      D[m, n] = sum_p  X[p (+) tap(m), ci(m)] * dY[p, n]        m = (tap, ci)  (R*S*Cin rows)
//                                                                n = co         (Cout columns)
//
// Both operands are "MN-major" for the tensor core (the contiguous NHWC channel axis is the M / N axis,
// the pixel axis is K), so each 64-channel x 64-pixel operand block is ONE 4D TMA box
// {64 channels, BW, BH, BI} (BW*BH*BI = 64 pixels) landing in smem as 64 rows of 128 swizzled bytes —
// exactly the canonical SWIZZLE_128B MN-major layout (LBO = 8 KB between 64-channel blocks, SBO = 1 KB
// between 8-pixel groups).  Filter taps are shifted boxes of x (zero padding = TMA OOB fill, stride 2 =
// parity sub-lattice maps), two (tap, ci-chunk) blocks form one M=128 tile.
// The pixel axis is split across CTAs (split-K); partial tiles are reduced with coalesced fp32 atomics.
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int WG_MAX_TAPS = 49;
constexpr int PIX_BLOCK = 64;  // pixels per pipeline stage (4 UMMA K-steps of 16)

struct WgGeom {
  int ntaps, cin, cout;
  int chunks_per_tap;          // ceil(cin / 64)
  int m_blocks;                // ntaps * chunks_per_tap   (64-row blocks of the M axis)
  int m_tiles;                 // ceil(m_blocks / 2)
  int n_tiles;                 // ceil(cout / BLOCK_N)
  int BW, BH, BI;              // pixel box, product 64
  int tiles_w, tiles_h, tiles_i;
  int ksplit;                  // CTAs sharing one output tile
  signed char tap_map[WG_MAX_TAPS], tap_dh[WG_MAX_TAPS], tap_dw[WG_MAX_TAPS];
};
struct WgMaps {
  CUtensorMap x[4];  // parity sub-lattices of x
  CUtensorMap dy;
};

template <int BLOCK_N>
struct WgCfg {
  static constexpr int A_BYTES = 2 * PIX_BLOCK * 128;                 // two 64-channel blocks
  static constexpr int B_BLOCKS = BLOCK_N / 64;
  static constexpr int B_BYTES = B_BLOCKS * PIX_BLOCK * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// MN-major, 128B-swizzled operand: 64 channels (128 B) per row, rows = pixels.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((PIX_BLOCK * 128) >> 4) << 16;     // LBO: next 64-channel block
  d |= (uint64_t)(1024 >> 4) << 32;                  // SBO: next group of 8 pixels
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgGeom geo, float* __restrict__ dW) {
  using Cfg = WgCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pix_tiles = geo.tiles_w * geo.tiles_h * geo.tiles_i;
  const long long num_items = (long long)geo.m_tiles * geo.n_tiles * geo.ksplit;
  const int per_split = (pix_tiles + geo.ksplit - 1) / geo.ksplit;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (m_tile fastest, then n_tile, then ks): the CTAs that run concurrently work on the SAME pixel range with
  // different (tap, channel) row tiles, so x / dy come from DRAM once and are shared through L2 (with the split index
  // fastest the m tiles of one pixel range ran in different waves: ncu 6.97 GB of DRAM reads for 2.31 GB of operands
  // on the 3x3 64->64 layer at 56x56).
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int m_tile = (int)(item % geo.m_tiles);
        const int n_tile = (int)((item / geo.m_tiles) % geo.n_tiles);
        const int ks = (int)(item / ((long long)geo.m_tiles * geo.n_tiles));
        const int p_beg = ks * per_split;
        const int p_end = min(p_beg + per_split, pix_tiles);
        for (int pt = p_beg; pt < p_end; ++pt) {
          const int w0 = (pt % geo.tiles_w) * geo.BW;
          const int h0 = ((pt / geo.tiles_w) % geo.tiles_h) * geo.BH;
          const int i0 = (pt / (geo.tiles_w * geo.tiles_h)) * geo.BI;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int mb = m_tile * 2 + b;
            // blocks past the end re-load block 0 (their rows are masked in the epilogue) so that the
            // expected transaction byte count stays constant
            const int mbc = mb < geo.m_blocks ? mb : 0;
            const int tap = mbc / geo.chunks_per_tap;
            const int c0 = (mbc - tap * geo.chunks_per_tap) * 64;
            tma_load_4d(sa + b * (PIX_BLOCK * 128), &maps.x[geo.tap_map[tap]], &full_bar[stage], c0,
                        w0 + geo.tap_dw[tap], h0 + geo.tap_dh[tap], i0);
          }
#pragma unroll
          for (int b = 0; b < Cfg::B_BLOCKS; ++b)
            tma_load_4d(sb + b * (PIX_BLOCK * 128), &maps.dy, &full_bar[stage], n_tile * BLOCK_N + b * 64, w0, h0, i0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // D = f32, A = B = bf16, both MN-major (bits 15, 16), N = BLOCK_N, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int ks = (int)(item / ((long long)geo.m_tiles * geo.n_tiles));
      const int p_beg = ks * per_split;
      const int p_end = min(p_beg + per_split, pix_tiles);
      if (p_end <= p_beg) continue;  // empty split: producer / epilogue skip it too
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int pt = p_beg; pt < p_end; ++pt) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = make_smem_desc_mn_sw128(sa);
          const uint64_t db = make_smem_desc_mn_sw128(sb);
#pragma unroll
          for (int k = 0; k < PIX_BLOCK / UMMA_K; ++k) {
            // 16 pixels further along K: 16 rows x 128 B = 2048 B = +128 in 16-byte units
            umma_bf16(tmem_d, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc,
                      (pt > p_beg || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (pt == p_end - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const long long rsc = (long long)geo.ntaps * geo.cin;
    for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int m_tile = (int)(item % geo.m_tiles);
      const int n_tile = (int)((item / geo.m_tiles) % geo.n_tiles);
      const int ks = (int)(item / ((long long)geo.m_tiles * geo.n_tiles));
      const int p_beg = ks * per_split;
      const int p_end = min(p_beg + per_split, pix_tiles);
      if (p_end <= p_beg) continue;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      // this thread's row: block (q >> 1) of the tile, channel (q & 1) * 32 + lane inside the block
      const int mb = m_tile * 2 + (q >> 1);
      const int tap = mb / geo.chunks_per_tap;
      const int ci = (mb - tap * geo.chunks_per_tap) * 64 + (q & 1) * 32 + lane;
      const bool row_ok = mb < geo.m_blocks && ci < geo.cin;
      float* drow = dW + (long long)tap * geo.cin + ci;
#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
        const int col0 = n_tile * BLOCK_N + chunk * 32;
        if (col0 >= geo.cout) break;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + chunk * 32), r);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < geo.cout) atomicAdd(drow + (long long)(col0 + j) * rsc, __uint_as_float(r[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

template <int BLOCK_N>
int launch_wgrad(const WgMaps& maps, const WgGeom& geo, float* dW, cudaStream_t stream) {
  using Cfg = WgCfg<BLOCK_N>;
  static bool configured = false;
  auto kern = tc_wgrad_kernel<BLOCK_N>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      adamml_set_error("tc_wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ADAMML_ERR_CUDA;
    }
    configured = true;
  }
  long long items = (long long)geo.m_tiles * geo.n_tiles * geo.ksplit;
  int sms = num_sms();
  int grid = (int)(items < sms ? items : sms);
  kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, geo, dW);
  return adamml_check_launch("tc_wgrad");
}


// fills the tiling / split-K fields of geo for ntaps x cin rows, cout columns and an [IMGS, Ho, Wo] pixel axis
int wg_tiling(WgGeom& geo, int ntaps, int cin, int cout, int Wo, int Ho, int IMGS) {
  geo.ntaps = ntaps;
  geo.cin = cin;
  geo.cout = cout;
  geo.chunks_per_tap = (cin + 63) / 64;
  geo.m_blocks = geo.ntaps * geo.chunks_per_tap;
  geo.m_tiles = (geo.m_blocks + 1) / 2;
  const int block_n = cout <= 64 ? 64 : (cout <= 128 ? 128 : 256);
  geo.n_tiles = (cout + block_n - 1) / block_n;
  pick_box(Wo, Ho, IMGS, PIX_BLOCK, &geo.BW, &geo.BH, &geo.BI);
  geo.tiles_w = (Wo + geo.BW - 1) / geo.BW;
  geo.tiles_h = (Ho + geo.BH - 1) / geo.BH;
  geo.tiles_i = (IMGS + geo.BI - 1) / geo.BI;
  const int pix_tiles = geo.tiles_w * geo.tiles_h * geo.tiles_i;
  // enough splits for ~3 items per SM -- never 3 and a bit: one CTA with a fourth item is a 33 % tail (ncu: average
  // SM active 74 % of the kernel at 445 items on 148 SMs) -- at least 8 pixel tiles per item
  long long base = (long long)geo.m_tiles * geo.n_tiles;
  long long want = (3LL * num_sms()) / base;
  long long cap = (pix_tiles + 7) / 8;
  long long ks = want < cap ? want : cap;
  if (ks < 1) ks = 1;
  // avoid empty trailing splits
  int per = (int)((pix_tiles + ks - 1) / ks);
  ks = (pix_tiles + per - 1) / per;
  geo.ksplit = (int)ks;
  return block_n;
}

int wg_launch(int block_n, const WgMaps& maps, const WgGeom& geo, float* dw, cudaStream_t stream) {
  cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)geo.cout * geo.ntaps * geo.cin, stream);
  if (block_n == 64) return launch_wgrad<64>(maps, geo, dw, stream);
  if (block_n == 128) return launch_wgrad<128>(maps, geo, dw, stream);
  return launch_wgrad<256>(maps, geo, dw, stream);
}

}  // namespace

extern "C" {

int adamml_tc_wgrad_supported(int Cin, int Cout, int R, int S, int stride) {
  if (Cin % 8 || Cout % 8) return 0;
  if (R * S > WG_MAX_TAPS || R < 1 || S < 1) return 0;
  if (stride != 1 && stride != 2) return 0;
  return 1;
}

// x [IMGS,H,W,Cin] bf16, dy [IMGS,Ho,Wo,Cout] bf16 -> dw fp32 [Cout][R][S][Cin] (overwritten)
int adamml_tc_wgrad_bf16(const void* x, const void* dy, float* dw, int IMGS, int H, int W, int Cin, int Cout, int R,
                         int S, int stride, int pad, int Ho, int Wo, cudaStream_t stream) {
  if (!adamml_tc_wgrad_supported(Cin, Cout, R, S, stride)) {
    adamml_set_error("tc_wgrad: Cin=%d Cout=%d R=%d S=%d stride=%d outside the tcgen05 envelope", Cin, Cout, R, S,
                     stride);
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(Ho == (H + 2 * pad - R) / stride + 1 && Wo == (W + 2 * pad - S) / stride + 1,
                 "tc_wgrad: Ho/Wo inconsistent with H/W/R/S/stride/pad");
  ADAMML_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0, "tc_wgrad: operands must be 16-byte aligned");
  WgGeom geo;
  memset(&geo, 0, sizeof(geo));
  const int block_n = wg_tiling(geo, R * S, Cin, Cout, Wo, Ho, IMGS);

  bool used[4] = {false, false, false, false};
  for (int r = 0; r < R; ++r)
    for (int s_ = 0; s_ < S; ++s_) {
      int th = r - pad, tw = s_ - pad;
      int ph = ((th % stride) + stride) % stride, pw = ((tw % stride) + stride) % stride;
      int t = r * S + s_;
      geo.tap_map[t] = (signed char)(ph * stride + pw);
      geo.tap_dh[t] = (signed char)((th - ph) / stride);
      geo.tap_dw[t] = (signed char)((tw - pw) / stride);
      used[ph * stride + pw] = true;
    }
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  const bf16* xb = (const bf16*)x;
  int first_used = -1;
  for (int ph = 0; ph < stride; ++ph)
    for (int pw = 0; pw < stride; ++pw) {
      int id = ph * stride + pw;
      if (!used[id]) continue;
      int Wd = (W - pw + stride - 1) / stride, Hd = (H - ph + stride - 1) / stride;
      if (Wd < 1) Wd = 1;
      if (Hd < 1) Hd = 1;
      int rc = make_map_4d(&maps.x[id], xb + ((long long)ph * W + pw) * Cin, Cin, Wd, Hd, IMGS, (long long)stride * Cin,
                           (long long)stride * W * Cin, (long long)H * W * Cin, geo.BW, geo.BH, geo.BI);
      if (rc) return rc;
      if (first_used < 0) first_used = id;
    }
  int rc = make_map_4d(&maps.dy, dy, Cout, Wo, Ho, IMGS, Cout, (long long)Wo * Cout, (long long)Ho * Wo * Cout, geo.BW,
                       geo.BH, geo.BI);
  if (rc) return rc;
  return wg_launch(block_n, maps, geo, dw, stream);
}

/* Weight gradient of the 7x7/s2 stem on the space-to-depth input (see adamml_tc_stem_conv_bf16):
 * dw fp32 [Cout][4][4*Cs] (overwritten), unpacked to OIHW by adamml_unpack_wgrad_stem. */
int adamml_tc_stem_wgrad_bf16(const void* xs, const void* dy, float* dw, int IMGS, int Hs, int Wp, int Cs, int Cout,
                              int Ho, int Wo, int taps, cudaStream_t stream) {
  if (Cs % 8 || Cout % 8 || (taps != 4 && taps != 2) || taps * Cs > 256) {
    adamml_set_error("tc_stem_wgrad: Cs=%d Cout=%d taps=%d outside the tcgen05 envelope", Cs, Cout, taps);
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(Ho == Hs && Wo + taps - 1 <= Wp, "tc_stem_wgrad: geometry (Ho == Hs, Wp >= Wo + taps - 1)");
  ADAMML_REQUIRE(((uintptr_t)xs % 16) == 0 && ((uintptr_t)dy % 16) == 0, "tc_stem_wgrad: operands must be 16-byte aligned");
  const int VC = taps * Cs;
  WgGeom geo;
  memset(&geo, 0, sizeof(geo));
  const int block_n = wg_tiling(geo, taps, VC, Cout, Wo, Ho, IMGS);
  for (int t = 0; t < taps; ++t) {
    geo.tap_map[t] = 0;
    geo.tap_dh[t] = (signed char)(t - taps / 2);
    geo.tap_dw[t] = 0;
  }
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = make_map_4d(&maps.x[0], xs, VC, Wp - (taps - 1), Hs, IMGS, Cs, (long long)Wp * Cs, (long long)Hs * Wp * Cs,
                       geo.BW, geo.BH, geo.BI);
  if (rc) return rc;
  rc = make_map_4d(&maps.dy, dy, Cout, Wo, Ho, IMGS, Cout, (long long)Wo * Cout, (long long)Ho * Wo * Cout, geo.BW,
                   geo.BH, geo.BI);
  if (rc) return rc;
  return wg_launch(block_n, maps, geo, dw, stream);
}

}  // extern "C"
