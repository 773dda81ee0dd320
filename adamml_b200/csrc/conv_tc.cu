// tcgen05 tensor-core engine: persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   D[M, Ncols] = A[M, K] . B[Ncols, K]^T        A, B bf16 K-major; fp32 accumulators in TMEM
//
// This is the 1x1-convolution / pointwise / linear workhorse of the AdaMML backbones in NHWC
// (reference call sites: models/resnet.py:40-43,96,104 Bottleneck conv1/conv3 + downsample,
// models/sound_mobilenet_v2.py:55,61 and models/policy_net.py:76,84,49 pointwise convs): with
// channels-last activations a 1x1 conv IS this GEMM with M = images*H*W pixels.  dgrad of the
// same layers is the same GEMM with A = dy and B = w^T.
//
// Structure (persistent, one CTA per SM, 64 + 128|256 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2D tiles / shifted 4D boxes (128B swizzle) -> smem ring
//   warp 1      MMA issuer     : one lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16) x4 per stage,
//                                tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..   epilogue       : tcgen05.ld TMEM -> registers -> (+ addend tile fetched by TMA) -> bf16 tile staged in
//                                128B-swizzled smem -> TMA store; optional fused train-mode BatchNorm statistics
//                                (per-column sum / sum of squares of the staged, bf16-rounded tile, accumulated in
//                                fp64 registers across the CTA's tiles, flushed with fp64 atomics per BN group)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// Ragged M / Ncols / K edges are handled by TMA out-of-bounds zero fill on loads and clipping on stores.
#include "tc_common.cuh"

namespace {
using namespace tc;

// Implicit-GEMM geometry: the 128 rows of an M tile are a BW x BH x BI box of output pixels; every filter
// tap is one shifted 4D TMA box {64 channels, BW, BH, BI} of the (sub-lattice of the) NHWC input.
constexpr int MAX_TAPS = 49;
struct ConvGeom {
  int ntaps, cin, kb_per_tap;
  int BW, BH, BI;
  int tiles_w, tiles_h, tiles_i;
  int Ho, Wo, IMGS;
  int imgs_per_group;
  // output lattice: pixel (img, oh, ow) of this launch is stored at (img, oh*out_s + out_ph, ow*out_s + out_pw)
  // of an [IMGS, out_H, out_W, ldd] tensor (the 4 parity classes of a stride-2 data gradient)
  int out_H, out_W, out_s, out_ph, out_pw;
  // addend_sub == 2: `addend` is the compact [IMGS, ceil(Ho/2), ceil(Wo/2), ldd] gradient of a stride-2 1x1
  // branch, added at even (oh, ow) only
  int addend_sub, add_H, add_W;
  int tap_koff[MAX_TAPS];  // K offset of the tap's weight slice inside a row of B
  signed char tap_map[MAX_TAPS], tap_dh[MAX_TAPS], tap_dw[MAX_TAPS];
};
struct ConvMaps {
  CUtensorMap m[4];
};
// x2 mode (two-plane activations, see common.cuh): tensor maps of the lo (fp16 remainder) planes of A / D, the lo conv
// tap maps, the lo plane of a linear-mode output, and the three further weight planes.  tcgen05.mma kind::f16 needs
// both operands of one instruction in the SAME 16-bit format (a mixed bf16 x fp16 descriptor is an illegal
// instruction), so the weight operand comes as FOUR planes (adamml_pack_weight_x2): the bf16 cascade b1 = bf16(w),
// b2 = bf16(w - b1), b3 = bf16(w - b1 - b2) multiplies the bf16 hi plane of the activations (exact to 24 bits of w),
// and f = fp16(w) multiplies their fp16 lo plane.  b[0..2] = {b2, b3, f}; b1 is the kernel's tmB.
struct X2Maps {
  CUtensorMap a, d;
  CUtensorMap b[3];
  CUtensorMap add;  // lo plane of a residual addend (fused BN + residual epilogue)
  ConvMaps c;
  void* dlin_lo;
};
// Fused epilogue of an inference-mode conv + BatchNorm (+ residual) + ReLU/ReLU6 (resnet.py:96-111,
// sound_mobilenet_v2.py:33-40, policy_net.py:38-52): with running statistics BN is a per-channel affine map known before
// the convolution runs, so out = act(acc * scale[c] + shift[c] (+ residual)) leaves the accumulator directly and the
// pre-BN tensor z never exists.  ss = [Ncols][2] fp32 (scale, shift); act applies after the addend.
struct EpiSpec {
  const float* ss;
  int act;
  LiveLimit live = LiveLimit{nullptr, 0};  // inference with skipping: only the first live.n clips are computed
};

// SHALLOW: reductions of one or two K blocks (the MobileNetV2 expand / project-gradient GEMMs, K = 16..96).  Their
// tiles are all epilogue; two stages suffice, and two co-resident CTAs per SM overlap one tile's epilogue chain
// (accumulator wait, tcgen05.ld, staging, store drain, statistics) with the other's.
// X2: the activations and the output are two-plane (hi bf16 + lo fp16) tensors, the weights four-plane (X2Maps); a
// stage holds [A_hi][A_lo][B1][B2][B3][Bf] and every K step issues A_lo*Bf, A_hi*B3, A_hi*B2, A_hi*B1 into the same
// fp32 accumulator.
template <int BLOCK_N, bool SHALLOW = false, int X2 = 0>
struct TcCfg {
  static constexpr int PLANES = X2 ? 2 : 1;              // planes of A and of D
  static constexpr int B_PLANES = X2 ? 4 : 1;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // per plane
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int A_ALL = PLANES * A_BYTES;
  static constexpr int STAGE_BYTES = A_ALL + B_PLANES * B_BYTES;
  static_assert(!SHALLOW || BLOCK_N <= 128, "two CTAs per SM need 2 x (2 x BLOCK_N) <= 512 TMEM columns");
  static_assert(!X2 || (BLOCK_N == 64 && !SHALLOW), "x2: 64-column tiles (64 KB per stage), one CTA per SM");
  // (stage counts that divide by the common K-block counts 1, 2, 4 keep the weight tile resident, see the producer)
  // X2 = 1: reductions of one or two K blocks (2 stages, double-buffered output staging, 8 epilogue warps: the tile is
  //         all epilogue);
  // X2 = 2: deep reductions, load / MMA bound (3 stages keep more loads in flight, single output staging buffer);
  // X2 = 3: 3..8 K blocks under MANY column blocks (the late ResNet 1x1 layers): as 1, but 4 epilogue warps
  //         (measured: 141 k x 1024 x 256  0.75 ms vs 0.86 ms as X2 = 2 and 0.98 ms as X2 = 1)
  static constexpr int STAGES = X2 == 2 ? 3 : (X2 ? 2 : (SHALLOW ? 2 : ((BLOCK_N == 256) ? 3 : (BLOCK_N == 128 ? 4 : 8))));
  static constexpr int CTAS_PER_SM = SHALLOW ? 2 : 1;
  // Epilogue warps: wide tiles (256 columns, issue bound) get TWO warps per TMEM lane quarter that split the
  // accumulator columns; narrow tiles are latency bound and run best with one warp per quarter (measured).
  // (x2 tiles: twice the conversion / statistics work per element -- two warps per quarter hide its latency)
  // (deep reductions are load / MMA bound: extra epilogue warps only steal issue slots from the control warps)
  static constexpr int EPI_WARPS = (BLOCK_N == 256 || X2 == 1) ? 8 : 4;
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int THREADS = 64 + EPI_THREADS;
  static constexpr bool BACKOFF = (BLOCK_N == 256 || X2 == 1);  // control warps share SM sub-partitions with epilogue warps
  static constexpr int SUBTILES = BLOCK_N / 64;              // 64-column (128-byte) output sub-tiles
  // staging tiles of the TMA store: double buffering (store of tile i drains while tile i+1 is staged) measured
  // no faster than a single buffer, which leaves the smem to the operand ring
  static constexpr int OUT_BUFS = ((SHALLOW && BLOCK_N == 64) || X2 == 1 || X2 == 3) ? 2 : 1;
  static constexpr int OUT_PLANE_BYTES = SUBTILES * BLOCK_M * 128;
  static constexpr int OUT_TILE_BYTES = PLANES * OUT_PLANE_BYTES;
  static constexpr int OUT_BYTES = OUT_BUFS * OUT_TILE_BYTES;
  // WIDE (deep-reduction x2 variants): the three bf16 weight planes b1 | b2 | b3 lie back to back in a stage, so
  // x_hi multiplies them as ONE 192-column MMA into three accumulator column groups and x_lo * fp16(w) goes into a
  // fourth; the epilogue adds the four groups (small terms first).  Per K step the tensor core then reads
  // (4 + 6) + (4 + 2) = 16 KB of operands from shared memory instead of 4 x (4 + 2) = 24 KB: these kernels are bound
  // by the SM's 128 B/clk shared-memory port (TMA fills + operand reads), not by the MMA rate (ncu: tensor pipe 42 %
  // active on the 3x3 64 -> 64 layer, MMA warp never waiting for loads).
  static constexpr bool WIDE = X2 >= 2;
  static constexpr int ACC_COLS = WIDE ? 4 * BLOCK_N : BLOCK_N;  // TMEM columns of one accumulator stage
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int RED_BYTES = BLOCK_N * 2 * 8;  // per-CTA fp64 column sums, combined before the global atomics
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + 1024 /*align slack*/ + 1024 /*barriers*/ + RED_BYTES;
};

template <int THREADS>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }
template <bool BACKOFF>
__device__ __forceinline__ void ctl_wait(uint64_t* bar, uint32_t parity) {
  if (BACKOFF) mbar_wait_backoff(bar, parity);
  else mbar_wait(bar, parity);
}

// per-thread running BatchNorm statistics of one (4-column group, row slice).  A flush (BN group or column block
// changes, CTA retires) is a collective of the epilogue threads: the row slices are first combined per column in
// shared memory, then ONE fp64 atomic per (column, sum|sum of squares) leaves the CTA -- narrow layers (16..96
// columns) would otherwise serialise hundreds of thousands of same-address atomics in L2.
struct StatAcc {
  double S[4], Q[4];
  long long g;   // uniform over the CTA
  int col;
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < 4; ++i) { S[i] = 0.0; Q[i] = 0.0; }
  }
  template <int BLOCK_N, int EPI_THREADS>
  __device__ __forceinline__ void flush(double* __restrict__ stats, int Ncols, double* red, int col_base, int st) {
    if (g >= 0) {
      if (col < Ncols) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (col + i < Ncols) {
            atomicAdd(red + (col - col_base + i) * 2 + 0, S[i]);
            atomicAdd(red + (col - col_base + i) * 2 + 1, Q[i]);
          }
      }
      epi_bar_sync<EPI_THREADS>();
      for (int j = st; j < BLOCK_N * 2; j += EPI_THREADS) {
        const double v = red[j];
        if (v != 0.0) {  // (col_base + j/2 < Ncols holds: columns beyond Ncols are never accumulated)
          atomicAdd(stats + (g * Ncols + col_base + (j >> 1)) * 2 + (j & 1), v);
          red[j] = 0.0;
        }
      }
      epi_bar_sync<EPI_THREADS>();
    }
    reset();
  }
};

// Tile order of a persistent CTA.  Default: tiles blockIdx.x, blockIdx.x + grid, ... with the column block fastest
// (neighbouring CTAs share one A tile through L2).  pin_n (wide layers whose weight block stays resident in the ring,
// grid a multiple of the column-block count): the CTA keeps ONE column block for its whole life and walks the row
// blocks c / nb, c / nb + grid / nb, ... -- the weight planes are loaded once per CTA instead of once per tile, and the
// nb CTAs of a row block still run in step and share its A tile through L2.
struct TileSched {
  bool pin;
  int num_n_blks, num_m_blks;
  long long num_tiles;
  __device__ __forceinline__ bool get(int i, int& m_blk, int& n_blk) const {
    if (pin) {
      const long long m = (long long)(blockIdx.x / num_n_blks) + (long long)i * (gridDim.x / num_n_blks);
      if (m >= num_m_blks) return false;
      m_blk = (int)m;
      n_blk = (int)(blockIdx.x % num_n_blks);
      return true;
    }
    const long long t = blockIdx.x + (long long)i * gridDim.x;
    if (t >= num_tiles) return false;
    const unsigned tu = (unsigned)t, q = tu / (unsigned)num_n_blks;  // (tile counts stay far below 2^31)
    n_blk = (int)(tu - q * (unsigned)num_n_blks);
    m_blk = (int)q;
    return true;
  }
};

// EPI: the fused inference epilogue (EpiSpec: BatchNorm scale / shift, activation, x2 residual planes, device-side work
// limit) is compiled in only for the inference instantiations -- the training kernels keep their lean epilogue (with the
// code compiled in but unused, the x2 training GEMMs ran 12-20 % slower: 168 instead of 136 registers per thread).
template <int BLOCK_N, bool CONV, bool SHALLOW = false, int X2 = 0, bool EPI = false>
__global__ void __launch_bounds__(TcCfg<BLOCK_N, SHALLOW, X2>::THREADS, TcCfg<BLOCK_N, SHALLOW, X2>::CTAS_PER_SM)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmAdd,
               const __grid_constant__ ConvMaps cmaps, const __grid_constant__ ConvGeom geo,
               const __grid_constant__ X2Maps x2, const bf16* __restrict__ addend, long long M, int Ncols, int K,
               long long ldd, double* __restrict__ stats, long long rows_per_group, bf16* __restrict__ dlin,
               int pin_n, const EpiSpec epi) {
  using Cfg = TcCfg<BLOCK_N, SHALLOW, X2>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic (keeps the shared address space visible to the compiler: LDS/STS
  // instead of generic LD/ST in the epilogue)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem + Cfg::STAGES * Cfg::STAGE_BYTES;  // 1024-aligned (all sizes are multiples of 1024)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + Cfg::OUT_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* add_bar = tmem_empty + 2;  // dense addend tile landed in the staging buffer (TMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(add_bar + 1);
  double* red = reinterpret_cast<double*>(stage_base + Cfg::OUT_BYTES + 1024);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int num_m_blks = CONV ? geo.tiles_w * geo.tiles_h * geo.tiles_i : (int)((M + BLOCK_M - 1) / BLOCK_M);
  if (EPI && epi.live.n) {  // device-side work limit: row blocks are image-major, so the live ones come first
    if (CONV) {
      const int ti = (int)((live_count(epi.live, geo.IMGS) + geo.BI - 1) / geo.BI);
      num_m_blks = geo.tiles_w * geo.tiles_h * ti;
    } else {
      num_m_blks = (int)((live_count(epi.live, M) + BLOCK_M - 1) / BLOCK_M);
    }
  }
  const int num_n_blks = (Ncols + BLOCK_N - 1) / BLOCK_N;
  const long long num_tiles = (long long)num_m_blks * num_n_blks;
  const int num_kb = CONV ? geo.ntaps * geo.kb_per_tap : (K + BLOCK_K - 1) / BLOCK_K;
  const TileSched sched{(pin_n & 1) != 0, num_n_blks, num_m_blks, num_tiles};

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], Cfg::EPI_WARPS); }
    mbar_init(add_bar, 1);
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
  // dense residual-gradient addend (same pixel lattice as the output): fetched by TMA into the staging tile
  // (x2 launches only carry an addend -- the residual planes -- in the fused inference epilogue)
  // (x2 training launches never carry an addend: their epilogue is compiled without the addend paths)
  constexpr bool HAS_ADD = !X2 || EPI;
  const bool add_tma = HAS_ADD && addend != nullptr && geo.addend_sub != 2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (!CONV) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      if (X2) {
        if (!CONV) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&x2.a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&x2.b[0])) : "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      // One column block and a K-block count that divides the ring: slot s always carries K block s % num_kb, and
      // every tile multiplies by the SAME weight tiles, so each slot receives its weight tile once and later tiles
      // load only their A rows.  (Re-fetching the few weight lines per tile from all 148 SMs hot-spots one L2
      // slice: narrow MobileNetV2 layers ran up to 3x slower with the reload.)
      const bool b_resident = (num_n_blks == 1 || sched.pin) && (Cfg::STAGES % num_kb) == 0;
      int b_loaded = 0;
      int n_blk, m_blk;
      for (int ti = 0; sched.get(ti, m_blk, n_blk); ++ti) {
        int w0 = 0, h0 = 0, i0 = 0;
        if (CONV) {
          w0 = (m_blk % geo.tiles_w) * geo.BW;
          h0 = ((m_blk / geo.tiles_w) % geo.tiles_h) * geo.BH;
          i0 = (m_blk / (geo.tiles_w * geo.tiles_h)) * geo.BI;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          ctl_wait<Cfg::BACKOFF>(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_ALL;
          const bool need_b = !b_resident || b_loaded < Cfg::STAGES;
          if (b_resident && need_b) ++b_loaded;
          mbar_expect_tx(&full_bar[stage], need_b ? Cfg::STAGE_BYTES : Cfg::A_ALL);
          if (CONV) {
            const int tap = kb / geo.kb_per_tap;
            const int c0 = (kb - tap * geo.kb_per_tap) * BLOCK_K;
            tma_load_4d(sa, &cmaps.m[geo.tap_map[tap]], &full_bar[stage], c0, w0 + geo.tap_dw[tap],
                        h0 + geo.tap_dh[tap], i0);
            if (need_b) tma_load_2d(sb, &tmB, &full_bar[stage], geo.tap_koff[tap] + c0, n_blk * BLOCK_N);
            if (X2) {
              tma_load_4d(sa + Cfg::A_BYTES, &x2.c.m[geo.tap_map[tap]], &full_bar[stage], c0, w0 + geo.tap_dw[tap],
                          h0 + geo.tap_dh[tap], i0);
              if (need_b) {
#pragma unroll
                for (int pl = 0; pl < 3; ++pl)
                  tma_load_2d(sb + (pl + 1) * Cfg::B_BYTES, &x2.b[pl], &full_bar[stage], geo.tap_koff[tap] + c0,
                              n_blk * BLOCK_N);
              }
            }
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
            if (need_b) tma_load_2d(sb, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
            if (X2) {
              tma_load_2d(sa + Cfg::A_BYTES, &x2.a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
              if (need_b) {
#pragma unroll
                for (int pl = 0; pl < 3; ++pl)
                  tma_load_2d(sb + (pl + 1) * Cfg::B_BYTES, &x2.b[pl], &full_bar[stage], kb * BLOCK_K,
                              n_blk * BLOCK_N);
              }
            }
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D=f32, A=B=bf16, K-major both, N=BLOCK_N, M=128
    // (a_format bits [7,10), b_format bits [10,13): 1 = bf16, 0 = fp16; both operands of one MMA share the format)
    const uint32_t idesc_f16 = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
    const uint32_t idesc = idesc_f16 | (1u << 7) | (1u << 10);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int n_blk_, m_blk_;
    for (int ti = 0; sched.get(ti, m_blk_, n_blk_); ++ti) {
      ctl_wait<Cfg::BACKOFF>(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
      for (int kb = 0; kb < num_kb; ++kb) {
        ctl_wait<Cfg::BACKOFF>(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_ALL;
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sb);
          if (X2) {
            const uint64_t da_lo = make_smem_desc_sw128(sa + Cfg::A_BYTES);
            const uint64_t db2 = make_smem_desc_sw128(sb + Cfg::B_BYTES);
            const uint64_t db3 = make_smem_desc_sw128(sb + 2 * Cfg::B_BYTES);
            const uint64_t dbf = make_smem_desc_sw128(sb + 3 * Cfg::B_BYTES);
            if (Cfg::WIDE && !(pin_n & 8)) {
              // x_hi * [b1 | b2 | b3] -> columns [0, 3 BLOCK_N), x_lo * fp16(w) -> columns [3 BLOCK_N, 4 BLOCK_N)
              const uint32_t idesc_w = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((3 * BLOCK_N) >> 3) << 17) |
                                       ((uint32_t)(BLOCK_M >> 4) << 24);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                const uint64_t o = (uint64_t)(k * 2);
                const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
                umma_bf16(tmem_d + 3 * BLOCK_N, da_lo + o, dbf + o, idesc_f16, accum);
                umma_bf16(tmem_d, da + o, db + o, idesc_w, accum);
              }
            } else {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                const uint64_t o = (uint64_t)(k * 2);
                // small terms first: lo * fp16(w), hi * b3, hi * b2, hi * b1
                umma_bf16(tmem_d, da_lo + o, dbf + o, idesc_f16, (kb > 0 || k > 0) ? 1u : 0u);
                umma_bf16(tmem_d, da + o, db3 + o, idesc, 1u);
                umma_bf16(tmem_d, da + o, db2 + o, idesc, 1u);
                umma_bf16(tmem_d, da + o, db + o, idesc, 1u);
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 32 bytes (16 bf16) along K inside the swizzle row: +2 in 16-byte units
              umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9, 256 threads) =====================
    // TMEM -> registers (thread = accumulator row) -> (+ addend) -> bf16 -> 128B-swizzled smem staging tile ->
    // one elected thread issues the TMA store(s) (fully coalesced, edge-clipped by the tensor map);
    // train-mode BatchNorm statistics are column sums over the staged (bf16-rounded) tile, kept in registers
    // across the tiles of this persistent CTA.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;  // which share of the 32-column chunks this warp drains (0 when 4 warps)
    const int et = q * 32 + lane;  // == accumulator row inside the tile
    const bool issuer = (warp == 2 && lane == 0);
    constexpr int QUADS = BLOCK_N / 4;          // 4-column (8-byte) groups per tile
    constexpr int RSPLIT = Cfg::EPI_THREADS / QUADS; // row slices per column group
    // statistics role of this thread: threads are numbered by `st` so that consecutive threads take
    // consecutive column groups (conflict-free 8-byte smem reads); thread (squad, srs) sums rows srs, srs+RSPLIT, ...
    const int st = (warp - 2) * 32 + lane;
    const int squad = st % QUADS, srs = st / QUADS;
    StatAcc sa_;
    sa_.reset();
    sa_.g = -1;
    sa_.col = 0;
    int last_n_blk = -1, last_ss_blk = -1;
    float2* ss_s = reinterpret_cast<float2*>(red);  // (scale, shift) of the tile's columns (no statistics in that mode)
    if (stats) {
      for (int j = st; j < BLOCK_N * 2; j += Cfg::EPI_THREADS) red[j] = 0.0;
      epi_bar_sync<Cfg::EPI_THREADS>();
    }
    // Register-resident statistics (shallow x2 training GEMMs: the tile is all epilogue, and re-reading the staged
    // tile for the column sums cost more issue slots than the conversion itself).  A thread owns ONE accumulator row
    // slot and a FIXED 32-column chunk for its whole life (one column block per CTA: a single block or pinned
    // scheduling), so sum / sum of squares are 2 x 16 packed fp32 FMAs per tile on the values it converts anyway; the
    // cross-row reduction (warp butterfly -> shared fp64 -> one global fp64 atomic per column) runs only when the BN
    // group changes or after RS_TILES tiles (bounds the fp32 partial sums).  Tiles that straddle two BN groups take
    // the staged-tile path below.
    constexpr bool REGSTATS = (X2 == 1) && !CONV && !EPI;
    constexpr int RS_N = REGSTATS ? 16 : 1;
    constexpr int RS_TILES = 32;
    float2 rsS[RS_N], rsQ[RS_N];
#pragma unroll
    for (int j = 0; j < RS_N; ++j) { rsS[j] = make_float2(0.f, 0.f); rsQ[j] = make_float2(0.f, 0.f); }
    long long rs_g = -1;
    int rs_tiles = 0;
    long long rs_gcur = 0, rs_gend = rows_per_group;  // group of the current tile's first row and its end row (the
                                                       // tiles of a CTA come in increasing row order: no division)
    const bool rs_on = REGSTATS && stats != nullptr && (num_n_blks == 1 || sched.pin);
    const int rs_nblk = sched.pin ? (int)(blockIdx.x % num_n_blks) : 0;
    auto rs_flush = [&]() {
      if constexpr (REGSTATS) {
        float s[32], qv[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          s[2 * j] = rsS[j].x; s[2 * j + 1] = rsS[j].y;
          qv[2 * j] = rsQ[j].x; qv[2 * j + 1] = rsQ[j].y;
          rsS[j] = make_float2(0.f, 0.f);
          rsQ[j] = make_float2(0.f, 0.f);
        }
        // halving butterfly over the warp's 32 rows: afterwards lane l holds the sums of column l of the chunk
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const float ss_ = up ? s[i] : s[i + off], ks = up ? s[i + off] : s[i];
            const float sq_ = up ? qv[i] : qv[i + off], kq = up ? qv[i + off] : qv[i];
            s[i] = ks + __shfl_xor_sync(0xffffffffu, ss_, off);
            qv[i] = kq + __shfl_xor_sync(0xffffffffu, sq_, off);
          }
        }
        const int cl = chalf * 32 + lane;
        if (rs_nblk * BLOCK_N + cl < Ncols) {
          atomicAdd(red + cl * 2 + 0, (double)s[0]);
          atomicAdd(red + cl * 2 + 1, (double)qv[0]);
        }
        epi_bar_sync<Cfg::EPI_THREADS>();
        for (int j = st; j < BLOCK_N * 2; j += Cfg::EPI_THREADS) {
          const double v = red[j];
          if (v != 0.0) {
            atomicAdd(stats + (rs_g * Ncols + rs_nblk * BLOCK_N + (j >> 1)) * 2 + (j & 1), v);
            red[j] = 0.0;
          }
        }
        epi_bar_sync<Cfg::EPI_THREADS>();
      }
    };
    int acc = 0, obuf = 0;
    uint32_t acc_phase = 0, add_phase = 0;
    int n_blk, m_blk;
    for (int ti = 0; sched.get(ti, m_blk, n_blk); ++ti) {
      uint8_t* stage_out = stage_base + obuf * Cfg::OUT_TILE_BYTES;
      if (Cfg::OUT_BUFS == 2) obuf ^= 1;
      long long arow = 0;
      bool row_ok, add_ok = HAS_ADD && addend != nullptr && !add_tma;
      int w0 = 0, h0 = 0, i0 = 0;
      if (CONV) {
        w0 = (m_blk % geo.tiles_w) * geo.BW;
        h0 = ((m_blk / geo.tiles_w) % geo.tiles_h) * geo.BH;
        i0 = (m_blk / (geo.tiles_w * geo.tiles_h)) * geo.BI;
        const int ow = w0 + et % geo.BW;
        const int oh = h0 + (et / geo.BW) % geo.BH;
        const int img = i0 + et / (geo.BW * geo.BH);
        row_ok = ow < geo.Wo && oh < geo.Ho && img < geo.IMGS;
        arow = ((long long)img * geo.out_H + oh * geo.out_s + geo.out_ph) * geo.out_W + ow * geo.out_s + geo.out_pw;
        if (geo.addend_sub == 2) {
          add_ok = add_ok && !((oh | ow) & 1);
          arow = ((long long)img * geo.add_H + (oh >> 1)) * geo.add_W + (ow >> 1);
        }
      } else {
        const long long row = (long long)m_blk * BLOCK_M + et;
        arow = row;
        row_ok = row < M;
      }
      if (EPI && epi.ss && n_blk != last_ss_blk) {
        // every thread is past the mid-tile barrier of the previous tile, i.e. done reading the previous block's pairs
        for (int j = st; j < BLOCK_N; j += Cfg::EPI_THREADS) {
          const int col = n_blk * BLOCK_N + j;
          ss_s[j] = col < Ncols ? reinterpret_cast<const float2*>(epi.ss)[col] : make_float2(0.f, 0.f);
        }
        last_ss_blk = n_blk;
      }
      bool rs_tile = false;
      if (REGSTATS && rs_on) {
        const long long row0 = (long long)m_blk * BLOCK_M;
        const long long rend = row0 + BLOCK_M <= M ? row0 + BLOCK_M : M;
        while (row0 >= rs_gend) { ++rs_gcur; rs_gend += rows_per_group; }
        rs_tile = rend <= rs_gend;  // the whole tile lies in group rs_gcur
        if (rs_tile) {
          if (rs_gcur != rs_g || rs_tiles == RS_TILES) {  // (uniform over the epilogue threads: a collective)
            if (rs_g >= 0) rs_flush();
            rs_g = rs_gcur;
            rs_tiles = 0;
          }
          ++rs_tiles;
        }
      }
      // the staging tile (and row_group) of the previous tile must have been consumed
      if (issuer) {
        if (Cfg::OUT_BUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      epi_bar_sync<Cfg::EPI_THREADS>();
      if (HAS_ADD && add_tma) {
        if (issuer) {
          int nsub = 0;
#pragma unroll
          for (int sub = 0; sub < Cfg::SUBTILES; ++sub) nsub += (n_blk * BLOCK_N + sub * 64 < Ncols) ? 1 : 0;
          mbar_expect_tx(add_bar, (uint32_t)(nsub * Cfg::PLANES) * (BLOCK_M * 128));
#pragma unroll
          for (int sub = 0; sub < Cfg::SUBTILES; ++sub) {
            const int col = n_blk * BLOCK_N + sub * 64;
            if (col < Ncols) {
              uint8_t* dst = stage_out + sub * (BLOCK_M * 128);
              if (CONV) tma_load_4d(dst, &tmAdd, add_bar, col, w0, h0, i0);
              else tma_load_2d(dst, &tmAdd, add_bar, col, m_blk * BLOCK_M);
              if (X2) {
                if (CONV) tma_load_4d(dst + Cfg::OUT_PLANE_BYTES, &x2.add, add_bar, col, w0, h0, i0);
                else tma_load_2d(dst + Cfg::OUT_PLANE_BYTES, &x2.add, add_bar, col, m_blk * BLOCK_M);
              }
            }
          }
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (HAS_ADD && add_tma) {
        mbar_wait(add_bar, add_phase);
        add_phase ^= 1;
      }
      constexpr int CHUNKS_PER_WARP = (BLOCK_N / 32) / (Cfg::EPI_WARPS / 4);
#pragma unroll 1
      for (int chunk = chalf * CHUNKS_PER_WARP; chunk < (chalf + 1) * CHUNKS_PER_WARP; ++chunk) {
        const int col0 = n_blk * BLOCK_N + chunk * 32;
        if (col0 >= Ncols) break;  // warp-uniform
        float v[32];
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS + chunk * 32);
        if (Cfg::WIDE && !(pin_n & 8)) {
          // four accumulator groups: ((x_lo * f + x_hi * b3) + x_hi * b2) + x_hi * b1
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32_nowait(tacc + 3 * BLOCK_N, r0);
          tmem_ld_32x32b_x32_nowait(tacc + 2 * BLOCK_N, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
          tmem_ld_32x32b_x32_nowait(tacc + BLOCK_N, r0);
          tmem_ld_32x32b_x32_nowait(tacc, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (v[j] + __uint_as_float(r0[j])) + __uint_as_float(r1[j]);
        } else {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tacc, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
        if (!row_ok) {  // edge tiles only: rows outside the tensor are staged as zeros (statistics sum them)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        } else if (EPI && epi.ss) {  // fused inference BatchNorm: per-column scale / shift (broadcast smem reads)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 p = ss_s[chunk * 32 + j];
            v[j] = fmaf(v[j], p.x, p.y);
          }
        }
        if (HAS_ADD && add_ok && row_ok) {
          const bf16* ap = addend + arow * ldd + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < Ncols) {
              uint4 pk = *reinterpret_cast<const uint4*>(ap + j);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float2 f = __bfloat1622float2(h2[t]);
                v[j + 2 * t] += f.x;
                v[j + 2 * t + 1] += f.y;
              }
            }
          }
        }
        // staging address: sub-tile (chunk / 2), row et, logical 16-byte chunk c -> physical c ^ (et & 7);
        // linear mode (dlin): plain row-major rows of Ncols elements, the image of the contiguous output tile
        uint8_t* srow = stage_out + (chunk >> 1) * (BLOCK_M * 128) + et * 128;
        uint8_t* lrow = stage_out + (size_t)et * (size_t)(Ncols * 2) + chunk * 64;
        if (HAS_ADD && add_tma && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int c = (chunk & 1) * 4 + (j >> 3);
            const uint4 pk = *reinterpret_cast<const uint4*>(srow + ((c ^ (et & 7)) << 4));
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float2 f = __bfloat1622float2(h2[t]);
              v[j + 2 * t] += f.x;
              v[j + 2 * t + 1] += f.y;
            }
            if (X2 && EPI) {
              const uint4 pl = *reinterpret_cast<const uint4*>(srow + Cfg::OUT_PLANE_BYTES + ((c ^ (et & 7)) << 4));
              const __half2* l2 = reinterpret_cast<const __half2*>(&pl);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float2 f = __half22float2(l2[t]);
                v[j + 2 * t] += f.x;
                v[j + 2 * t + 1] += f.y;
              }
            }
          }
        }
        if (EPI && epi.act != ADAMML_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = act_apply(v[j], epi.act);
        }
        if (REGSTATS && rs_tile) {
#pragma unroll
          for (int j = 0; j < RS_N; ++j) {
            const float2 vv = make_float2(v[(2 * j) & 31], v[(2 * j + 1) & 31]);
            rsS[j] = __ffma2_rn(vv, make_float2(1.f, 1.f), rsS[j]);
            rsQ[j] = __ffma2_rn(vv, vv, rsQ[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]);
          __nv_bfloat162 p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
          __nv_bfloat162 p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&p0);
          pk.y = *reinterpret_cast<uint32_t*>(&p1);
          pk.z = *reinterpret_cast<uint32_t*>(&p2);
          pk.w = *reinterpret_cast<uint32_t*>(&p3);
          const int c = (chunk & 1) * 4 + (j >> 3);
          if (!CONV && dlin) {
            if (col0 + j < Ncols) *reinterpret_cast<uint4*>(lrow + (j << 1)) = pk;
          } else {
            *reinterpret_cast<uint4*>(srow + ((c ^ (et & 7)) << 4)) = pk;
          }
          if (X2) {  // lo plane: fp16 of the remainder v - bf16(v)
            const float2 h0 = __bfloat1622float2(p0), h1 = __bfloat1622float2(p1);
            const float2 h2 = __bfloat1622float2(p2), h3 = __bfloat1622float2(p3);
            __half2 l0 = __floats2half2_rn(v[j] - h0.x, v[j + 1] - h0.y);
            __half2 l1 = __floats2half2_rn(v[j + 2] - h1.x, v[j + 3] - h1.y);
            __half2 l2 = __floats2half2_rn(v[j + 4] - h2.x, v[j + 5] - h2.y);
            __half2 l3 = __floats2half2_rn(v[j + 6] - h3.x, v[j + 7] - h3.y);
            uint4 pl;
            pl.x = *reinterpret_cast<uint32_t*>(&l0);
            pl.y = *reinterpret_cast<uint32_t*>(&l1);
            pl.z = *reinterpret_cast<uint32_t*>(&l2);
            pl.w = *reinterpret_cast<uint32_t*>(&l3);
            if (!CONV && dlin) {
              if (col0 + j < Ncols) *reinterpret_cast<uint4*>(lrow + Cfg::OUT_PLANE_BYTES + (j << 1)) = pl;
            } else {
              *reinterpret_cast<uint4*>(srow + Cfg::OUT_PLANE_BYTES + ((c ^ (et & 7)) << 4)) = pl;
            }
          }
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      // (every lane has executed tcgen05.wait::ld, so its TMEM reads are complete; the before_thread_sync fence in
      // front of the arrive compiles to a MEMBAR.ALL.CTA in the per-tile critical chain -- pin_n bit 1 drops it)
      if (!(pin_n & 2)) tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      // publish the staged tile to the async proxy and store it
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      epi_bar_sync<Cfg::EPI_THREADS>();
      if (issuer) {
        if (!CONV && dlin) {
          // the tile's rows are one contiguous chunk of the output: a single bulk copy of full 128-byte lines
          // instead of 128 narrow row requests per 64-column sub-tile
          const long long row0 = (long long)m_blk * BLOCK_M;
          const long long rows = (M - row0) < BLOCK_M ? (M - row0) : BLOCK_M;
          bulk_store_1d(dlin + row0 * Ncols, stage_out, (uint32_t)(rows * Ncols * 2));
          if (X2)
            bulk_store_1d(reinterpret_cast<bf16*>(x2.dlin_lo) + row0 * Ncols, stage_out + Cfg::OUT_PLANE_BYTES,
                          (uint32_t)(rows * Ncols * 2));
        } else {
#pragma unroll
          for (int sub = 0; sub < Cfg::SUBTILES; ++sub) {
            const int col = n_blk * BLOCK_N + sub * 64;
            if (col < Ncols) {
              if (CONV) tma_store_4d(&tmD, stage_out + sub * (BLOCK_M * 128), col, w0, h0, i0);
              else tma_store_2d(&tmD, stage_out + sub * (BLOCK_M * 128), col, m_blk * BLOCK_M);
              if (X2) {
                const uint8_t* lo = stage_out + Cfg::OUT_PLANE_BYTES + sub * (BLOCK_M * 128);
                if (CONV) tma_store_4d(&x2.d, lo, col, w0, h0, i0);
                else tma_store_2d(&x2.d, lo, col, m_blk * BLOCK_M);
              }
            }
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (stats && !rs_tile) {
        if (n_blk != last_n_blk) {
          sa_.template flush<BLOCK_N, Cfg::EPI_THREADS>(stats, Ncols, red, last_n_blk * BLOCK_N, st);
          last_n_blk = n_blk;
          sa_.col = n_blk * BLOCK_N + squad * 4;
        }
        const bool col_ok = sa_.col < Ncols;
        {  // every epilogue thread walks the row segments (the flushes inside are collectives)
          // rows of a tile are ordered by BN group, so the tile is a few row segments of constant group;
          // invalid rows were staged as zeros and may be summed into any group
          const bool lin = !CONV && dlin != nullptr;
          const uint8_t* sbase = lin ? stage_out + squad * 8
                                     : stage_out + (squad >> 4) * (BLOCK_M * 128) + ((squad & 1) << 3);
          const int pitch = lin ? Ncols * 2 : 128;
          const int c = (squad & 15) >> 1;
          int r0 = 0;
          while (r0 < BLOCK_M) {
            long long g;
            int r1;
            if (CONV) {
              const int pb = geo.BW * geo.BH;
              const int img = i0 + r0 / pb;
              if (img >= geo.IMGS) break;
              g = img / geo.imgs_per_group;
              const long long e = ((g + 1) * geo.imgs_per_group - i0) * pb;
              r1 = e < BLOCK_M ? (int)e : BLOCK_M;
            } else {
              const long long row0 = (long long)m_blk * BLOCK_M;
              if (row0 + r0 >= M) break;
              // (M < 2^31 for every launch, see adamml_tc_supported: 32-bit division instead of the ~80-instruction
              // 64-bit one, executed by every epilogue thread for every tile)
              g = (long long)((unsigned)(row0 + r0) / (unsigned)rows_per_group);
              const long long e = (g + 1) * rows_per_group - row0;
              r1 = e < BLOCK_M ? (int)e : BLOCK_M;
            }
            if (g != sa_.g) {
              sa_.template flush<BLOCK_N, Cfg::EPI_THREADS>(stats, Ncols, red, n_blk * BLOCK_N, st);
              sa_.g = g;
            }
            float s[4] = {0.f, 0.f, 0.f, 0.f}, qq[4] = {0.f, 0.f, 0.f, 0.f};
            // first row of this thread's slice inside [r0, r1)
            int rw = r0 + ((srs - r0) % RSPLIT + RSPLIT) % RSPLIT;
            if (!col_ok) rw = r1;
#pragma unroll 8
            for (; rw < r1; rw += RSPLIT) {
              const uint8_t* sp = sbase + rw * pitch + (lin ? 0 : ((c ^ (rw & 7)) << 4));
              const uint2 pk = *reinterpret_cast<const uint2*>(sp);
              float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
              float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
              if (X2) {  // statistics of the full-precision value hi + lo
                const uint2 pl = *reinterpret_cast<const uint2*>(sp + Cfg::OUT_PLANE_BYTES);
                const float2 g0 = __half22float2(*reinterpret_cast<const __half2*>(&pl.x));
                const float2 g1 = __half22float2(*reinterpret_cast<const __half2*>(&pl.y));
                f0.x += g0.x; f0.y += g0.y; f1.x += g1.x; f1.y += g1.y;
              }
              s[0] += f0.x; qq[0] = fmaf(f0.x, f0.x, qq[0]);
              s[1] += f0.y; qq[1] = fmaf(f0.y, f0.y, qq[1]);
              s[2] += f1.x; qq[2] = fmaf(f1.x, f1.x, qq[2]);
              s[3] += f1.y; qq[3] = fmaf(f1.y, f1.y, qq[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { sa_.S[i] += (double)s[i]; sa_.Q[i] += (double)qq[i]; }
            r0 = r1;
          }
        }
      }
    }
    if (stats) sa_.template flush<BLOCK_N, Cfg::EPI_THREADS>(stats, Ncols, red, last_n_blk * BLOCK_N, st);
    if (REGSTATS && rs_on && rs_g >= 0) rs_flush();
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

// ---- host side ---------------------------------------------------------------------------
// 2D bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle
int make_map_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { adamml_set_error("cuTensorMapEncodeTiled entry point unavailable"); return ADAMML_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    adamml_set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
    return ADAMML_ERR_CUDA;
  }
  return ADAMML_OK;
}

// TcCfg variant of an x2 launch (see TcCfg::STAGES); ADAMML_B200_X2_VARIANT=1|2|3 forces one (experiments)
int x2_variant(int K, int Ncols) {
  static const int forced = []() { const char* e = getenv("ADAMML_B200_X2_VARIANT"); return e ? atoi(e) : 0; }();
  if (forced >= 1 && forced <= 3) return forced;
  if (K <= 2 * BLOCK_K) return 1;
  if (K <= 8 * BLOCK_K && Ncols >= 4 * 64) return 3;
  return 2;
}

const X2Maps& no_x2() {
  static X2Maps z;
  static bool init = false;
  if (!init) { memset(&z, 0, sizeof(z)); init = true; }
  return z;
}

template <int BLOCK_N, bool CONV, bool SHALLOW, int X2, bool EPI>
int launch_tc_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const CUtensorMap& tmAdd,
              const ConvMaps& cm, const ConvGeom& geo, const void* addend, long long M, int Ncols, int K, long long ldd, double* stats,
              long long rpg, cudaStream_t stream, void* dlin = nullptr, const X2Maps& x2 = no_x2(),
              const EpiSpec& epi = EpiSpec{nullptr, ADAMML_ACT_NONE}) {
  using Cfg = TcCfg<BLOCK_N, SHALLOW, X2>;
  static bool configured = false;
  auto kern = tc_gemm_kernel<BLOCK_N, CONV, SHALLOW, X2, EPI>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      adamml_set_error("tc_gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ADAMML_ERR_CUDA;
    }
    configured = true;
  }
  long long m_blks = CONV ? (long long)geo.tiles_w * geo.tiles_h * geo.tiles_i : (M + BLOCK_M - 1) / BLOCK_M;
  const int n_blks = (Ncols + BLOCK_N - 1) / BLOCK_N;
  long long tiles = m_blks * n_blks;
  const int sms = num_sms() * Cfg::CTAS_PER_SM;
  int grid = (int)(tiles < sms ? tiles : sms);
  // column-block pinning (TileSched): x2 layers with several column blocks -- when the weight block fits the ring it
  // stays resident; with fused statistics pinning pays even when it does not: a CTA that changes its column block
  // every tile flushes its statistics every tile (shared fp64 CAS atomics + 128 global fp64 atomics in the epilogue
  // chain; ncu on 141 k x 1024 x 256: 61 % of the shared-memory wavefronts were conflict replays of that flush,
  // 0.84 -> 0.52 ms pinned; 1x1/s2 256 -> 512 at 56x56: 2.90 -> 1.89 ms), a pinned CTA only when the BatchNorm group
  // changes.  (Column-block counts that divide the grid -- 2, 4 -- never changed block and are unaffected.)
  // ADAMML_B200_TC_PIN=1: resident case only, 0: off.
  static const int pin_mode = []() { const char* e = getenv("ADAMML_B200_TC_PIN"); return e ? atoi(e) : 2; }();
  const int num_kb = CONV ? geo.ntaps * geo.kb_per_tap : (K + BLOCK_K - 1) / BLOCK_K;
  int pin_n = 0;
  const bool pin_fit = (Cfg::STAGES % num_kb) == 0 || (pin_mode >= 2 && stats != nullptr);
  if (X2 && pin_mode > 0 && n_blks > 1 && n_blks <= sms && pin_fit && m_blks >= 4LL * (sms / n_blks)) {
    pin_n = 1;
    grid = (sms / n_blks) * n_blks;
  }
  static const bool nofence = []() { const char* e = getenv("ADAMML_B200_TC_NOFENCE"); return e && e[0] == '1'; }();
  if (nofence) pin_n |= 2;
  // (bit 3: the WIDE variants fall back to four 64-column MMAs per K step -- A/B experiments)
  static const bool nowide = []() { const char* e = getenv("ADAMML_B200_X2_WIDE"); return e && e[0] == '0'; }();
  if (nowide) pin_n |= 8;
  EpiSpec e = epi;
  if (EPI && e.ss && !stats) e.live = adamml_live_limit(CONV ? (long long)geo.IMGS : M);  // inference launches only
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmD, tmAdd, cm, geo, x2, (const bf16*)addend, M,
                                                        Ncols, K, ldd, stats, rpg, (bf16*)dlin, pin_n, e);
  return adamml_check_launch(CONV ? "tc_conv" : "tc_gemm");
}

template <int BLOCK_N, bool CONV, bool SHALLOW = false, int X2 = 0>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const CUtensorMap& tmAdd,
              const ConvMaps& cm, const ConvGeom& geo, const void* addend, long long M, int Ncols, int K, long long ldd,
              double* stats, long long rpg, cudaStream_t stream, void* dlin = nullptr, const X2Maps& x2 = no_x2(),
              const EpiSpec& epi = EpiSpec{nullptr, ADAMML_ACT_NONE}) {
  if (epi.ss)
    return launch_tc_impl<BLOCK_N, CONV, SHALLOW, X2, true>(tmA, tmB, tmD, tmAdd, cm, geo, addend, M, Ncols, K, ldd,
                                                            stats, rpg, stream, dlin, x2, epi);
  return launch_tc_impl<BLOCK_N, CONV, SHALLOW, X2, false>(tmA, tmB, tmD, tmAdd, cm, geo, addend, M, Ncols, K, ldd,
                                                           stats, rpg, stream, dlin, x2, epi);
}

template <bool CONV>
int dispatch_tc(int block_n, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                const CUtensorMap& tmAdd, const ConvMaps& cm, const ConvGeom& geo, const void* addend, long long M, int Ncols, int K, long long ldd, double* stats,
                long long rpg, cudaStream_t stream, void* dlin = nullptr,
                const EpiSpec& epi = EpiSpec{nullptr, ADAMML_ACT_NONE}) {
#define ADAMML_TC_CASE(BN) \
  if (block_n == BN)       \
    return launch_tc<BN, CONV>(tmA, tmB, tmD, tmAdd, cm, geo, addend, M, Ncols, K, ldd, stats, rpg, stream, dlin, \
                               no_x2(), epi);
  ADAMML_TC_CASE(64)
  ADAMML_TC_CASE(128)
  ADAMML_TC_CASE(256)
#undef ADAMML_TC_CASE
  adamml_set_error("tc: bad block_n %d", block_n);
  return ADAMML_ERR_ARG;
}


void set_tiles(ConvGeom& geo, int Wo, int Ho, int IMGS) {
  pick_box(Wo, Ho, IMGS, 128, &geo.BW, &geo.BH, &geo.BI);
  geo.tiles_w = (Wo + geo.BW - 1) / geo.BW;
  geo.tiles_h = (Ho + geo.BH - 1) / geo.BH;
  geo.tiles_i = (IMGS + geo.BI - 1) / geo.BI;
  geo.Ho = Ho; geo.Wo = Wo; geo.IMGS = IMGS;
}

int make_out_map(CUtensorMap* map, const ConvGeom& geo, const void* y, int Cout) {
  return make_map_4d(map, (const bf16*)y + ((long long)geo.out_ph * geo.out_W + geo.out_pw) * Cout, Cout, geo.Wo, geo.Ho,
                     geo.IMGS, (long long)geo.out_s * Cout, (long long)geo.out_s * geo.out_W * Cout,
                     (long long)geo.out_H * geo.out_W * Cout, geo.BW, geo.BH, geo.BI);
}

// the four weight planes [4][Ncols][w_ld] (adamml_pack_weight_x2): plane 0 -> tmB, planes 1..3 -> x2.b
int make_w4_maps(CUtensorMap* tmB, X2Maps& x2, const void* w4, int Ncols, long long w_ld) {
  const bf16* w = (const bf16*)w4;
  int rc = make_map_2d(tmB, w, Ncols, w_ld, w_ld, 64);
  for (int pl = 0; pl < 3 && !rc; ++pl)
    rc = make_map_2d(&x2.b[pl], w + (long long)(pl + 1) * Ncols * w_ld, Ncols, w_ld, w_ld, 64);
  return rc;
}

// x2 variant of run_conv: cm / cm_lo are the tap maps of the hi / lo input planes; w4 the four weight planes
int run_conv_x2(const ConvGeom& geo, const ConvMaps& cm, const ConvMaps& cm_lo, const void* w4, long long w_ld, void* y,
                void* y_lo, int Cout, double* stats, cudaStream_t stream,
                const EpiSpec& epi = EpiSpec{nullptr, ADAMML_ACT_NONE}, const void* res = nullptr,
                const void* res_lo = nullptr) {
  CUtensorMap tmB, tmD;
  X2Maps x2;
  memset(&x2, 0, sizeof(x2));
  x2.c = cm_lo;
  int rc = make_w4_maps(&tmB, x2, w4, Cout, w_ld);
  if (rc) return rc;
  rc = make_out_map(&tmD, geo, y, Cout);
  if (rc) return rc;
  rc = make_out_map(&x2.d, geo, y_lo, Cout);
  if (rc) return rc;
  CUtensorMap tmAdd = tmD;
  if (res) {
    rc = make_out_map(&tmAdd, geo, res, Cout);
    if (!rc) rc = make_out_map(&x2.add, geo, res_lo, Cout);
    if (rc) return rc;
  }
  if (stats) {
    int G = (geo.IMGS + geo.imgs_per_group - 1) / geo.imgs_per_group;
    cudaMemsetAsync(stats, 0, sizeof(double) * (size_t)G * Cout * 2, stream);
  }
  const long long M = (long long)geo.IMGS * geo.Ho * geo.Wo;
  const int K = geo.ntaps * geo.kb_per_tap * BLOCK_K;
  const int variant = x2_variant(K, Cout);
  if (variant == 2)
    return launch_tc<64, true, false, 2>(tmB, tmB, tmD, tmAdd, cm, geo, res, M, Cout, K, Cout, stats, 0, stream,
                                         nullptr, x2, epi);
  if (variant == 3)
    return launch_tc<64, true, false, 3>(tmB, tmB, tmD, tmAdd, cm, geo, res, M, Cout, K, Cout, stats, 0, stream,
                                         nullptr, x2, epi);
  return launch_tc<64, true, false, 1>(tmB, tmB, tmD, tmAdd, cm, geo, res, M, Cout, K, Cout, stats, 0, stream,
                                       nullptr, x2, epi);
}

// launches the implicit-GEMM kernel for a prepared geometry; w is [Cout][w_ld] bf16, y/addend rows have Cout columns
int run_conv(const ConvGeom& geo, const ConvMaps& cm, const void* w, long long w_ld, void* y, const void* addend,
             int Cout, double* stats, cudaStream_t stream, const EpiSpec& epi = EpiSpec{nullptr, ADAMML_ACT_NONE}) {
  const int block_n = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
  CUtensorMap tmB;
  int rc = make_map_2d(&tmB, w, Cout, w_ld, w_ld, block_n);
  if (rc) return rc;
  // output view: the (possibly strided) pixel lattice this launch writes, same {64 ch, BW, BH, BI} boxes as the input
  CUtensorMap tmD;
  rc = make_map_4d(&tmD, (const bf16*)y + ((long long)geo.out_ph * geo.out_W + geo.out_pw) * Cout, Cout, geo.Wo, geo.Ho,
                   geo.IMGS, (long long)geo.out_s * Cout, (long long)geo.out_s * geo.out_W * Cout,
                   (long long)geo.out_H * geo.out_W * Cout, geo.BW, geo.BH, geo.BI);
  if (rc) return rc;
  if (stats) {
    int G = (geo.IMGS + geo.imgs_per_group - 1) / geo.imgs_per_group;
    cudaMemsetAsync(stats, 0, sizeof(double) * (size_t)G * Cout * 2, stream);
  }
  CUtensorMap tmAdd = tmD;
  if (addend && geo.addend_sub != 2) {
    rc = make_map_4d(&tmAdd, (const bf16*)addend + ((long long)geo.out_ph * geo.out_W + geo.out_pw) * Cout, Cout, geo.Wo,
                     geo.Ho, geo.IMGS, (long long)geo.out_s * Cout, (long long)geo.out_s * geo.out_W * Cout,
                     (long long)geo.out_H * geo.out_W * Cout, geo.BW, geo.BH, geo.BI);
    if (rc) return rc;
  }
  return dispatch_tc<true>(block_n, tmB, tmB, tmD, tmAdd, cm, geo, addend, (long long)geo.IMGS * geo.Ho * geo.Wo, Cout,
                           geo.ntaps * geo.kb_per_tap * BLOCK_K, Cout, stats, 0, stream, nullptr, epi);
}

// geometry + tap tensor maps of an RxS / stride 1|2 convolution over x [IMGS,H,W,Cin] (x_lo: optional lo plane -> cm_lo)
int conv_setup(ConvGeom& geo, ConvMaps& cm, ConvMaps* cm_lo, const void* x, const void* x_lo, int IMGS, int H, int W,
               int Cin, int R, int S, int stride, int pad, int Ho, int Wo, int imgs_per_group) {
  memset(&geo, 0, sizeof(geo));
  geo.ntaps = R * S;
  geo.cin = Cin;
  geo.kb_per_tap = (Cin + BLOCK_K - 1) / BLOCK_K;
  set_tiles(geo, Wo, Ho, IMGS);
  geo.out_H = Ho; geo.out_W = Wo; geo.out_s = 1;
  geo.imgs_per_group = imgs_per_group > 0 ? imgs_per_group : IMGS;
  // taps: input coordinate = stride*o + (r - pad) = stride*(o + d) + parity
  bool used[4] = {false, false, false, false};
  for (int r = 0; r < R; ++r)
    for (int s_ = 0; s_ < S; ++s_) {
      int th = r - pad, tw = s_ - pad;
      int ph = ((th % stride) + stride) % stride, pw = ((tw % stride) + stride) % stride;
      int t = r * S + s_;
      geo.tap_koff[t] = t * Cin;
      geo.tap_map[t] = (signed char)(ph * stride + pw);
      geo.tap_dh[t] = (signed char)((th - ph) / stride);
      geo.tap_dw[t] = (signed char)((tw - pw) / stride);
      used[ph * stride + pw] = true;
    }
  memset(&cm, 0, sizeof(cm));
  if (cm_lo) memset(cm_lo, 0, sizeof(*cm_lo));
  for (int ph = 0; ph < stride; ++ph)
    for (int pw = 0; pw < stride; ++pw) {
      int id = ph * stride + pw;
      if (!used[id]) continue;
      int Wd = (W - pw + stride - 1) / stride, Hd = (H - ph + stride - 1) / stride;
      if (Wd <= 0 || Hd <= 0) { Wd = Wd > 0 ? Wd : 1; Hd = Hd > 0 ? Hd : 1; }
      const long long off = ((long long)ph * W + pw) * Cin;
      int rc = make_map_4d(&cm.m[id], (const bf16*)x + off, Cin, Wd, Hd, IMGS, (long long)stride * Cin,
                           (long long)stride * W * Cin, (long long)H * W * Cin, geo.BW, geo.BH, geo.BI);
      if (rc) return rc;
      if (cm_lo) {
        rc = make_map_4d(&cm_lo->m[id], (const bf16*)x_lo + off, Cin, Wd, Hd, IMGS, (long long)stride * Cin,
                         (long long)stride * W * Cin, (long long)H * W * Cin, geo.BW, geo.BH, geo.BI);
        if (rc) return rc;
      }
    }
  return ADAMML_OK;
}

// geometry of a stride-2 first convolution on the space-to-depth operand (see adamml_tc_stem_conv_bf16)
int stem_setup(ConvGeom& geo, ConvMaps& cm, ConvMaps* cm_lo, const void* xs, const void* xs_lo, int IMGS, int Hs, int Wp,
               int Cs, int Ho, int Wo, int taps, int imgs_per_group) {
  const int VC = taps * Cs;  // virtual channels per position: `taps` consecutive s2d columns
  memset(&geo, 0, sizeof(geo));
  geo.ntaps = taps;
  geo.cin = VC;
  geo.kb_per_tap = (VC + BLOCK_K - 1) / BLOCK_K;
  for (int t = 0; t < taps; ++t) {
    geo.tap_koff[t] = t * VC;
    geo.tap_map[t] = 0;
    geo.tap_dh[t] = (signed char)(t - taps / 2);
    geo.tap_dw[t] = 0;
  }
  set_tiles(geo, Wo, Ho, IMGS);
  geo.out_H = Ho; geo.out_W = Wo; geo.out_s = 1;
  geo.imgs_per_group = imgs_per_group > 0 ? imgs_per_group : IMGS;
  memset(&cm, 0, sizeof(cm));
  // positions 0..Wp-taps: position p covers stored columns p..p+taps-1
  int rc = make_map_4d(&cm.m[0], xs, VC, Wp - (taps - 1), Hs, IMGS, Cs, (long long)Wp * Cs, (long long)Hs * Wp * Cs,
                       geo.BW, geo.BH, geo.BI);
  if (rc) return rc;
  if (cm_lo) {
    memset(cm_lo, 0, sizeof(*cm_lo));
    rc = make_map_4d(&cm_lo->m[0], xs_lo, VC, Wp - (taps - 1), Hs, IMGS, Cs, (long long)Wp * Cs,
                     (long long)Hs * Wp * Cs, geo.BW, geo.BH, geo.BI);
  }
  return rc;
}

}  // namespace

extern "C" {

int adamml_tc_supported(long long M, int Ncols, int K, long long lda, long long ldb, long long ldd) {
  if (M <= 0 || Ncols <= 0 || K <= 0) return 0;
  if (lda <= 0) lda = K;
  if (ldb <= 0) ldb = K;
  if (ldd <= 0) ldd = Ncols;
  if (K % 8 || Ncols % 8 || lda % 8 || ldb % 8 || ldd % 8) return 0;  // 16-byte TMA strides / vector stores
  if (lda < K || ldb < K || ldd < Ncols) return 0;
  if (M >= (1LL << 31)) return 0;
  return 1;
}

static int gemm_bf16_impl(const void* A, const void* B, void* D, long long M, int Ncols, int K, long long lda,
                          long long ldb, long long ldd, int d_dtype, double* stats, long long rows_per_group,
                          const EpiSpec& epi, const void* res, cudaStream_t stream) {
  if (lda <= 0) lda = K;
  if (ldb <= 0) ldb = K;
  if (ldd <= 0) ldd = Ncols;
  if (!adamml_tc_supported(M, Ncols, K, lda, ldb, ldd)) {
    adamml_set_error("tc_gemm: shape M=%lld N=%d K=%d lda=%lld ldb=%lld ldd=%lld outside the tcgen05 envelope", M,
                     Ncols, K, lda, ldb, ldd);
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)D % 16) == 0,
                 "tc_gemm: operands must be 16-byte aligned");
  if (d_dtype != ADAMML_BF16) {
    adamml_set_error("tc_gemm: only bf16 output (the epilogue stages bf16 tiles for the TMA store)");
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(!stats || rows_per_group > 0, "tc_gemm: stats need rows_per_group");
  // shallow reductions under narrow tiles: two co-resident CTAs per SM (measured: 64x64 layer1 conv1 0.64 -> 0.39 ms;
  // wider outputs split into 128-column blocks lose against one 256-column CTA, so those keep the deep config)
  static const bool shallow_on = []() { const char* e = getenv("ADAMML_B200_TC_SHALLOW"); return !(e && e[0] == '0'); }();
  const bool shallow = shallow_on && K <= 2 * BLOCK_K && Ncols <= 128;
  const int block_n = Ncols <= 64 ? 64 : (Ncols <= 128 ? 128 : 256);
  CUtensorMap tmA, tmB;
  int rc = make_map_2d(&tmA, A, M, K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_map_2d(&tmB, B, Ncols, K, ldb, block_n);
  if (rc) return rc;
  if (stats) {
    long long G = (M + rows_per_group - 1) / rows_per_group;
    cudaMemsetAsync(stats, 0, sizeof(double) * (size_t)G * Ncols * 2, stream);
  }
  CUtensorMap tmD;
  rc = make_map_2d(&tmD, D, M, Ncols, ldd, BLOCK_M);
  if (rc) return rc;
  ConvMaps cm;
  ConvGeom geo;
  memset(&cm, 0, sizeof(cm));
  memset(&geo, 0, sizeof(geo));
  // contiguous output whose rows are not whole 128-byte lines (N = 16, 24, 32, 96, 144, ...): the TMA unit retires
  // about one box row per 4 cycles whatever its width, so the tile is stored as ONE linear bulk copy instead
  static const bool linear_on = []() { const char* e = getenv("ADAMML_B200_TC_LINEAR"); return !(e && e[0] == '0'); }();
  void* dlin = (linear_on && !res && ldd == Ncols && Ncols <= block_n && (Ncols % 64) != 0) ? D : nullptr;
  CUtensorMap tmAdd = tmD;
  if (res) {  // residual tile of the fused epilogue: same [M, Ncols] lattice as the output
    ADAMML_REQUIRE(((uintptr_t)res % 16) == 0, "tc_gemm: residual must be 16-byte aligned");
    rc = make_map_2d(&tmAdd, res, M, Ncols, ldd, BLOCK_M);
    if (rc) return rc;
  }
  if (shallow) {
    if (block_n == 64)
      return launch_tc<64, false, true>(tmA, tmB, tmD, tmAdd, cm, geo, res, M, Ncols, K, ldd, stats, rows_per_group,
                                        stream, dlin, no_x2(), epi);
    return launch_tc<128, false, true>(tmA, tmB, tmD, tmAdd, cm, geo, res, M, Ncols, K, ldd, stats, rows_per_group,
                                       stream, dlin, no_x2(), epi);
  }
  return dispatch_tc<false>(block_n, tmA, tmB, tmD, tmAdd, cm, geo, res, M, Ncols, K, ldd, stats, rows_per_group,
                            stream, dlin, epi);
}

int adamml_tc_gemm_bf16(const void* A, const void* B, void* D, long long M, int Ncols, int K, long long lda,
                        long long ldb, long long ldd, int d_dtype, double* stats, long long rows_per_group,
                        cudaStream_t stream) {
  return gemm_bf16_impl(A, B, D, M, Ncols, K, lda, ldb, ldd, d_dtype, stats, rows_per_group,
                        EpiSpec{nullptr, ADAMML_ACT_NONE}, nullptr, stream);
}

/* Inference-mode conv + BatchNorm (+ residual) + ReLU/ReLU6 in ONE kernel (resnet.py:96-111 and the MobileNetV2
 * ConvBNReLU stacks): out = act(A.B^T * scale[c] + shift[c] (+ res)); scale_shift = [Ncols][2] fp32 (the folded
 * running statistics of adamml_bn_finalize, group 0), res = optional [M, Ncols] tensor, dense strides. */
int adamml_tc_gemm_bn_act_bf16(const void* A, const void* B, void* D, long long M, int Ncols, int K,
                               const float* scale_shift, int act, const void* res, cudaStream_t stream) {
  ADAMML_REQUIRE(scale_shift, "tc_gemm_bn_act: needs the folded BatchNorm scale / shift");
  return gemm_bf16_impl(A, B, D, M, Ncols, K, 0, 0, 0, ADAMML_BF16, nullptr, 0, EpiSpec{scale_shift, act}, res, stream);
}

int adamml_tc_conv_supported(int Cin, int Cout, int R, int S, int stride) {
  if (Cin % 8 || Cout % 8) return 0;
  if (R * S > MAX_TAPS || R < 1 || S < 1) return 0;
  if (stride != 1 && stride != 2) return 0;
  return 1;
}

static int conv_bf16_impl(const void* x, const void* w, void* y, const void* addend, int IMGS, int H, int W, int Cin,
                          int Cout, int R, int S, int stride, int pad, int Ho, int Wo, double* stats,
                          int imgs_per_group, int addend_sub, const EpiSpec& epi, cudaStream_t stream) {
  if (!adamml_tc_conv_supported(Cin, Cout, R, S, stride)) {
    adamml_set_error("tc_conv: Cin=%d Cout=%d R=%d S=%d stride=%d outside the tcgen05 envelope", Cin, Cout, R, S,
                     stride);
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(Ho == (H + 2 * pad - R) / stride + 1 && Wo == (W + 2 * pad - S) / stride + 1,
                 "tc_conv: Ho/Wo inconsistent with H/W/R/S/stride/pad");
  ADAMML_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)y % 16) == 0 &&
                     ((uintptr_t)addend % 16) == 0,
                 "tc_conv: operands must be 16-byte aligned");
  ADAMML_REQUIRE(!stats || imgs_per_group > 0, "tc_conv: stats need imgs_per_group");
  ADAMML_REQUIRE(addend_sub == 0 || addend_sub == 1 || addend_sub == 2, "tc_conv: addend_sub must be 0, 1 or 2");
  ConvGeom geo;
  ConvMaps cm;
  int rc = conv_setup(geo, cm, nullptr, x, nullptr, IMGS, H, W, Cin, R, S, stride, pad, Ho, Wo, imgs_per_group);
  if (rc) return rc;
  if (addend && addend_sub == 2) { geo.addend_sub = 2; geo.add_H = (Ho + 1) / 2; geo.add_W = (Wo + 1) / 2; }
  return run_conv(geo, cm, w, (long long)R * S * Cin, y, addend, Cout, stats, stream, epi);
}

int adamml_tc_conv_bf16(const void* x, const void* w, void* y, const void* addend, int IMGS, int H, int W, int Cin,
                        int Cout, int R, int S, int stride, int pad, int Ho, int Wo, double* stats,
                        int imgs_per_group, int addend_sub, cudaStream_t stream) {
  return conv_bf16_impl(x, w, y, addend, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, stats, imgs_per_group,
                        addend_sub, EpiSpec{nullptr, ADAMML_ACT_NONE}, stream);
}

int adamml_tc_conv_bn_act_bf16(const void* x, const void* w, void* y, int IMGS, int H, int W, int Cin, int Cout, int R,
                               int S, int stride, int pad, int Ho, int Wo, const float* scale_shift, int act,
                               const void* res, cudaStream_t stream) {
  ADAMML_REQUIRE(scale_shift, "tc_conv_bn_act: needs the folded BatchNorm scale / shift");
  return conv_bf16_impl(x, w, y, res, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, nullptr, 0, res ? 1 : 0,
                        EpiSpec{scale_shift, act}, stream);
}

/* x2 planes: forward convolution of the default precision mode.  x / y are (hi bf16, lo fp16) plane pairs, w4 the
 * four weight planes of adamml_pack_weight_x2; every K step accumulates x_hi*(b1+b2+b3) + x_lo*fp16(w) in the fp32
 * TMEM accumulator; fused BN statistics are taken over the full-precision outputs. */
static int conv_x2_impl(const void* x_hi, const void* x_lo, const void* w4, void* y_hi, void* y_lo, int IMGS, int H,
                        int W, int Cin, int Cout, int R, int S, int stride, int pad, int Ho, int Wo, double* stats,
                        int imgs_per_group, const EpiSpec& epi, const void* res_hi, const void* res_lo,
                        cudaStream_t stream) {
  if (!adamml_tc_conv_supported(Cin, Cout, R, S, stride)) {
    adamml_set_error("tc_conv_x2: Cin=%d Cout=%d R=%d S=%d stride=%d outside the tcgen05 envelope", Cin, Cout, R, S,
                     stride);
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(Ho == (H + 2 * pad - R) / stride + 1 && Wo == (W + 2 * pad - S) / stride + 1,
                 "tc_conv_x2: Ho/Wo inconsistent with H/W/R/S/stride/pad");
  ADAMML_REQUIRE(x_hi && x_lo && w4 && y_hi && y_lo, "tc_conv_x2: every tensor needs all its planes");
  ADAMML_REQUIRE(((uintptr_t)x_hi % 16) == 0 && ((uintptr_t)x_lo % 16) == 0 && ((uintptr_t)w4 % 16) == 0 &&
                     ((uintptr_t)y_hi % 16) == 0 && ((uintptr_t)y_lo % 16) == 0,
                 "tc_conv_x2: operands must be 16-byte aligned");
  ADAMML_REQUIRE(!stats || imgs_per_group > 0, "tc_conv_x2: stats need imgs_per_group");
  ConvGeom geo;
  ConvMaps cm, cm_lo;
  int rc = conv_setup(geo, cm, &cm_lo, x_hi, x_lo, IMGS, H, W, Cin, R, S, stride, pad, Ho, Wo, imgs_per_group);
  if (rc) return rc;
  ADAMML_REQUIRE((!res_hi == !res_lo) && ((uintptr_t)res_hi % 16) == 0 && ((uintptr_t)res_lo % 16) == 0,
                 "tc_conv_x2: the residual needs both planes, 16-byte aligned");
  return run_conv_x2(geo, cm, cm_lo, w4, (long long)R * S * Cin, y_hi, y_lo, Cout, stats, stream, epi, res_hi, res_lo);
}

int adamml_tc_conv_x2(const void* x_hi, const void* x_lo, const void* w4, void* y_hi, void* y_lo, int IMGS, int H,
                      int W, int Cin, int Cout, int R, int S, int stride, int pad, int Ho, int Wo, double* stats,
                      int imgs_per_group, cudaStream_t stream) {
  return conv_x2_impl(x_hi, x_lo, w4, y_hi, y_lo, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, stats,
                      imgs_per_group, EpiSpec{nullptr, ADAMML_ACT_NONE}, nullptr, nullptr, stream);
}

/* x2 planes of the fused inference epilogue (see adamml_tc_gemm_bn_act_bf16): out = act(conv * scale + shift (+ res)) */
int adamml_tc_conv_bn_act_x2(const void* x_hi, const void* x_lo, const void* w4, void* y_hi, void* y_lo, int IMGS,
                             int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int Ho, int Wo,
                             const float* scale_shift, int act, const void* res_hi, const void* res_lo,
                             cudaStream_t stream) {
  ADAMML_REQUIRE(scale_shift, "tc_conv_bn_act_x2: needs the folded BatchNorm scale / shift");
  return conv_x2_impl(x_hi, x_lo, w4, y_hi, y_lo, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, nullptr, 0,
                      EpiSpec{scale_shift, act}, res_hi, res_lo, stream);
}

static int gemm_x2_impl(const void* A_hi, const void* A_lo, const void* B4, void* D_hi, void* D_lo, long long M,
                        int Ncols, int K, double* stats, long long rows_per_group, const EpiSpec& epi,
                        const void* res_hi, const void* res_lo, cudaStream_t stream) {
  if (!adamml_tc_supported(M, Ncols, K, K, K, Ncols)) {
    adamml_set_error("tc_gemm_x2: shape M=%lld N=%d K=%d outside the tcgen05 envelope", M, Ncols, K);
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(A_hi && A_lo && B4 && D_hi && D_lo, "tc_gemm_x2: every tensor needs all its planes");
  ADAMML_REQUIRE(((uintptr_t)A_hi % 16) == 0 && ((uintptr_t)A_lo % 16) == 0 && ((uintptr_t)B4 % 16) == 0 &&
                     ((uintptr_t)D_hi % 16) == 0 && ((uintptr_t)D_lo % 16) == 0 && ((long long)Ncols * K) % 8 == 0,
                 "tc_gemm_x2: operands must be 16-byte aligned");
  ADAMML_REQUIRE(!stats || rows_per_group > 0, "tc_gemm_x2: stats need rows_per_group");
  const int block_n = 64;
  CUtensorMap tmA, tmB, tmD;
  X2Maps x2;
  memset(&x2, 0, sizeof(x2));
  int rc = make_map_2d(&tmA, A_hi, M, K, K, BLOCK_M);
  if (!rc) rc = make_map_2d(&x2.a, A_lo, M, K, K, BLOCK_M);
  if (!rc) rc = make_w4_maps(&tmB, x2, B4, Ncols, K);
  if (!rc) rc = make_map_2d(&tmD, D_hi, M, Ncols, Ncols, BLOCK_M);
  if (!rc) rc = make_map_2d(&x2.d, D_lo, M, Ncols, Ncols, BLOCK_M);
  if (rc) return rc;
  if (stats) {
    long long G = (M + rows_per_group - 1) / rows_per_group;
    cudaMemsetAsync(stats, 0, sizeof(double) * (size_t)G * Ncols * 2, stream);
  }
  ConvMaps cm;
  ConvGeom geo;
  memset(&cm, 0, sizeof(cm));
  memset(&geo, 0, sizeof(geo));
  void* dlin = (!res_hi && Ncols <= block_n && (Ncols % 64) != 0) ? D_hi : nullptr;
  x2.dlin_lo = dlin ? D_lo : nullptr;
  CUtensorMap tmAdd = tmD;
  if (res_hi) {
    ADAMML_REQUIRE(res_lo && ((uintptr_t)res_hi % 16) == 0 && ((uintptr_t)res_lo % 16) == 0,
                   "tc_gemm_x2: the residual needs both planes, 16-byte aligned");
    rc = make_map_2d(&tmAdd, res_hi, M, Ncols, Ncols, BLOCK_M);
    if (!rc) rc = make_map_2d(&x2.add, res_lo, M, Ncols, Ncols, BLOCK_M);
    if (rc) return rc;
  }
  const int variant = x2_variant(K, Ncols);
  if (variant == 2)
    return launch_tc<64, false, false, 2>(tmA, tmB, tmD, tmAdd, cm, geo, res_hi, M, Ncols, K, Ncols, stats,
                                          rows_per_group, stream, dlin, x2, epi);
  if (variant == 3)
    return launch_tc<64, false, false, 3>(tmA, tmB, tmD, tmAdd, cm, geo, res_hi, M, Ncols, K, Ncols, stats,
                                          rows_per_group, stream, dlin, x2, epi);
  return launch_tc<64, false, false, 1>(tmA, tmB, tmD, tmAdd, cm, geo, res_hi, M, Ncols, K, Ncols, stats,
                                        rows_per_group, stream, dlin, x2, epi);
}

int adamml_tc_gemm_x2(const void* A_hi, const void* A_lo, const void* B4, void* D_hi, void* D_lo, long long M,
                      int Ncols, int K, double* stats, long long rows_per_group, cudaStream_t stream) {
  return gemm_x2_impl(A_hi, A_lo, B4, D_hi, D_lo, M, Ncols, K, stats, rows_per_group,
                      EpiSpec{nullptr, ADAMML_ACT_NONE}, nullptr, nullptr, stream);
}

int adamml_tc_gemm_bn_act_x2(const void* A_hi, const void* A_lo, const void* B4, void* D_hi, void* D_lo, long long M,
                             int Ncols, int K, const float* scale_shift, int act, const void* res_hi,
                             const void* res_lo, cudaStream_t stream) {
  ADAMML_REQUIRE(scale_shift, "tc_gemm_bn_act_x2: needs the folded BatchNorm scale / shift");
  return gemm_x2_impl(A_hi, A_lo, B4, D_hi, D_lo, M, Ncols, K, nullptr, 0, EpiSpec{scale_shift, act}, res_hi, res_lo,
                      stream);
}

/* Data gradient of a stride-2 convolution (resnet.py:100 conv2 of the first Bottleneck of layer2-4) as four
 * stride-1 implicit GEMMs, one per parity class (ph, pw) of the input pixel grid: class outputs
 * dx[2a+ph, 2b+pw] = sum over the taps (r, s) with (ph+pad-r), (pw+pad-s) even of
 * dy[a + (ph+pad-r)/2, b + (pw+pad-s)/2] . w[r, s]; every MAC of the forward pass is issued exactly once.
 * w_rot = adamml_pack_weight_dgrad operand [Cin][R][S][Cout] (tap (r,s) of w sits at (R-1-r, S-1-s)). */
int adamml_tc_dgrad_s2_bf16(const void* dy, const void* w_rot, void* dx, int IMGS, int H, int W, int Cin, int Cout,
                            int R, int S, int pad, int Ho, int Wo, cudaStream_t stream) {
  if (Cin % 8 || Cout % 8 || R * S > MAX_TAPS || R < 2 || S < 2) {
    adamml_set_error("tc_dgrad_s2: Cin=%d Cout=%d R=%d S=%d outside the tcgen05 envelope", Cin, Cout, R, S);
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(Ho == (H + 2 * pad - R) / 2 + 1 && Wo == (W + 2 * pad - S) / 2 + 1,
                 "tc_dgrad_s2: Ho/Wo inconsistent with H/W/R/S/pad");
  ADAMML_REQUIRE(((uintptr_t)dy % 16) == 0 && ((uintptr_t)w_rot % 16) == 0 && ((uintptr_t)dx % 16) == 0,
                 "tc_dgrad_s2: operands must be 16-byte aligned");
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const int Hc = (H - ph + 1) / 2, Wc = (W - pw + 1) / 2;
      if (Hc <= 0 || Wc <= 0) continue;
      ConvGeom geo;
      memset(&geo, 0, sizeof(geo));
      geo.cin = Cout;  // reduction channels of the gradient GEMM
      geo.kb_per_tap = (Cout + BLOCK_K - 1) / BLOCK_K;
      int nt = 0;
      for (int r = 0; r < R; ++r) {
        if ((ph + pad - r) & 1) continue;
        for (int s_ = 0; s_ < S; ++s_) {
          if ((pw + pad - s_) & 1) continue;
          geo.tap_koff[nt] = ((R - 1 - r) * S + (S - 1 - s_)) * Cout;
          geo.tap_map[nt] = 0;
          // arithmetic shift == floor division by 2 (numerator is even here)
          geo.tap_dh[nt] = (signed char)((ph + pad - r) / 2);
          geo.tap_dw[nt] = (signed char)((pw + pad - s_) / 2);
          ++nt;
        }
      }
      ADAMML_REQUIRE(nt > 0, "tc_dgrad_s2: parity class (%d,%d) has no taps (R=%d S=%d pad=%d)", ph, pw, R, S, pad);
      geo.ntaps = nt;
      set_tiles(geo, Wc, Hc, IMGS);
      geo.out_H = H; geo.out_W = W; geo.out_s = 2; geo.out_ph = ph; geo.out_pw = pw;
      geo.imgs_per_group = IMGS;
      ConvMaps cm;
      memset(&cm, 0, sizeof(cm));
      int rc = make_map_4d(&cm.m[0], dy, Cout, Wo, Ho, IMGS, Cout, (long long)Wo * Cout, (long long)Ho * Wo * Cout,
                           geo.BW, geo.BH, geo.BI);
      if (rc) return rc;
      rc = run_conv(geo, cm, w_rot, (long long)R * S * Cout, dx, nullptr, Cin, nullptr, stream);
      if (rc) return rc;
    }
  return ADAMML_OK;
}

/* Stride-2 first convolutions on a space-to-depth operand: the 7x7 / pad 3 ResNet stem (resnet.py:138,199; taps = 4)
 * and the 3x3 / pad 1 first conv of the MobileNetV2s (sound_mobilenet_v2.py:120, policy_net.py:117; taps = 2).
 * Described for the stem: 7x7 / stride 2 / pad 3 on the space-to-depth input written by
 * adamml_pack_frames_s2d: xs [IMGS, Hs, Wp, Cs] bf16 (Hs = H/2 rows, Wp = W/2 + 4 columns of which the first two
 * and the last two are zeros, Cs = 4*C padded to a multiple of 16).  In that layout the stem is a 4-tap
 * (dh = -2..1) stride-1 convolution whose "pixel" is the 4*Cs-wide window of four consecutive s2d columns:
 * that window is CONTIGUOUS in memory, so it is one TMA box of a tensor map whose W stride (Cs elements) is
 * smaller than its innermost extent (4*Cs elements) — an overlapping, im2col-free view.  K = 16*Cs.
 * w: adamml_pack_weight_stem operand [Cout][4][4*Cs] bf16. */
static int stem_bf16_impl(const void* xs, const void* w, void* y, int IMGS, int Hs, int Wp, int Cs, int Cout, int Ho,
                          int Wo, int taps, double* stats, int imgs_per_group, const EpiSpec& epi,
                          cudaStream_t stream) {
  if (Cs % 8 || Cout % 8 || (taps != 4 && taps != 2) || taps * Cs > 256) {
    adamml_set_error("tc_stem_conv: Cs=%d Cout=%d taps=%d outside the tcgen05 envelope", Cs, Cout, taps);
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(Ho == Hs && Wo + taps - 1 <= Wp, "tc_stem_conv: geometry (Ho == Hs, Wp >= Wo + taps - 1)");
  ADAMML_REQUIRE(((uintptr_t)xs % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)y % 16) == 0,
                 "tc_stem_conv: operands must be 16-byte aligned");
  ADAMML_REQUIRE(!stats || imgs_per_group > 0, "tc_stem_conv: stats need imgs_per_group");
  ConvGeom geo;
  ConvMaps cm;
  int rc = stem_setup(geo, cm, nullptr, xs, nullptr, IMGS, Hs, Wp, Cs, Ho, Wo, taps, imgs_per_group);
  if (rc) return rc;
  return run_conv(geo, cm, w, (long long)taps * taps * Cs, y, nullptr, Cout, stats, stream, epi);
}

int adamml_tc_stem_conv_bf16(const void* xs, const void* w, void* y, int IMGS, int Hs, int Wp, int Cs, int Cout,
                             int Ho, int Wo, int taps, double* stats, int imgs_per_group, cudaStream_t stream) {
  return stem_bf16_impl(xs, w, y, IMGS, Hs, Wp, Cs, Cout, Ho, Wo, taps, stats, imgs_per_group,
                        EpiSpec{nullptr, ADAMML_ACT_NONE}, stream);
}

int adamml_tc_stem_conv_bn_act_bf16(const void* xs, const void* w, void* y, int IMGS, int Hs, int Wp, int Cs, int Cout,
                                    int Ho, int Wo, int taps, const float* scale_shift, int act,
                                    cudaStream_t stream) {
  ADAMML_REQUIRE(scale_shift, "tc_stem_conv_bn_act: needs the folded BatchNorm scale / shift");
  return stem_bf16_impl(xs, w, y, IMGS, Hs, Wp, Cs, Cout, Ho, Wo, taps, nullptr, 0, EpiSpec{scale_shift, act}, stream);
}

/* x2 planes of the same first convolutions (operands from adamml_pack_frames_s2d_x2 / adamml_nhwc_to_s2d per plane and
 * adamml_pack_weight_x2 with stem = 1). */
static int stem_x2_impl(const void* xs_hi, const void* xs_lo, const void* w4, void* y_hi, void* y_lo, int IMGS,
                        int Hs, int Wp, int Cs, int Cout, int Ho, int Wo, int taps, double* stats,
                        int imgs_per_group, const EpiSpec& epi, cudaStream_t stream) {
  if (Cs % 8 || Cout % 8 || (taps != 4 && taps != 2) || taps * Cs > 256) {
    adamml_set_error("tc_stem_conv_x2: Cs=%d Cout=%d taps=%d outside the tcgen05 envelope", Cs, Cout, taps);
    return ADAMML_ERR_UNSUPPORTED;
  }
  ADAMML_REQUIRE(Ho == Hs && Wo + taps - 1 <= Wp, "tc_stem_conv_x2: geometry (Ho == Hs, Wp >= Wo + taps - 1)");
  ADAMML_REQUIRE(xs_hi && xs_lo && w4 && y_hi && y_lo, "tc_stem_conv_x2: every tensor needs all its planes");
  ADAMML_REQUIRE(((uintptr_t)xs_hi % 16) == 0 && ((uintptr_t)xs_lo % 16) == 0 && ((uintptr_t)w4 % 16) == 0 &&
                     ((uintptr_t)y_hi % 16) == 0 && ((uintptr_t)y_lo % 16) == 0,
                 "tc_stem_conv_x2: operands must be 16-byte aligned");
  ADAMML_REQUIRE(!stats || imgs_per_group > 0, "tc_stem_conv_x2: stats need imgs_per_group");
  ConvGeom geo;
  ConvMaps cm, cm_lo;
  int rc = stem_setup(geo, cm, &cm_lo, xs_hi, xs_lo, IMGS, Hs, Wp, Cs, Ho, Wo, taps, imgs_per_group);
  if (rc) return rc;
  return run_conv_x2(geo, cm, cm_lo, w4, (long long)taps * taps * Cs, y_hi, y_lo, Cout, stats, stream, epi);
}

int adamml_tc_stem_conv_x2(const void* xs_hi, const void* xs_lo, const void* w4, void* y_hi, void* y_lo, int IMGS,
                           int Hs, int Wp, int Cs, int Cout, int Ho, int Wo, int taps, double* stats,
                           int imgs_per_group, cudaStream_t stream) {
  return stem_x2_impl(xs_hi, xs_lo, w4, y_hi, y_lo, IMGS, Hs, Wp, Cs, Cout, Ho, Wo, taps, stats, imgs_per_group,
                      EpiSpec{nullptr, ADAMML_ACT_NONE}, stream);
}

int adamml_tc_stem_conv_bn_act_x2(const void* xs_hi, const void* xs_lo, const void* w4, void* y_hi, void* y_lo,
                                  int IMGS, int Hs, int Wp, int Cs, int Cout, int Ho, int Wo, int taps,
                                  const float* scale_shift, int act, cudaStream_t stream) {
  ADAMML_REQUIRE(scale_shift, "tc_stem_conv_bn_act_x2: needs the folded BatchNorm scale / shift");
  return stem_x2_impl(xs_hi, xs_lo, w4, y_hi, y_lo, IMGS, Hs, Wp, Cs, Cout, Ho, Wo, taps, nullptr, 0,
                      EpiSpec{scale_shift, act}, stream);
}

}  // extern "C"
