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
// Structure (one CTA per SM, 192 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2D tiles (128B swizzle) -> smem ring
//   warp 1      MMA issuer     : one lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16) x4 per stage,
//                                tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..5  epilogue       : tcgen05.ld TMEM -> registers -> (bf16|fp32) global stores, plus the
//                                optional fused train-mode BatchNorm statistics (per-column sum and
//                                sum of squares, reduced with warp shuffles, fp64 atomics)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// Ragged M / Ncols / K edges are handled by TMA out-of-bounds zero fill + masked stores.
#include "common.cuh"
#include <cuda.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                            // leading byte offset (unused: one atom along K)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BLOCK_N>
struct TcCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, bool D_F32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               void* __restrict__ Dptr, long long M, int Ncols, int K, long long ldd, double* __restrict__ stats,
               long long rows_per_group) {
  using Cfg = TcCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m_blks = (int)((M + BLOCK_M - 1) / BLOCK_M);
  const int num_n_blks = (Ncols + BLOCK_N - 1) / BLOCK_N;
  const long long num_tiles = (long long)num_m_blks * num_n_blks;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = (int)(tile % num_n_blks);
        const int m_blk = (int)(tile / num_n_blks);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D=f32, A=B=bf16, K-major both, N=BLOCK_N, M=128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                           ((uint32_t)(BLOCK_M >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) along K inside the swizzle row: +2 in 16-byte units
            umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
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
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_blk = (int)(tile % num_n_blks);
      const int m_blk = (int)(tile / num_n_blks);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const long long row = (long long)m_blk * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < M;
      // BN groups touched by this warp's 32 rows
      long long g_lo = 0, g_hi = 0;
      if (stats) {
        long long r0 = (long long)m_blk * BLOCK_M + q * 32;
        long long r1 = r0 + 31 < M - 1 ? r0 + 31 : M - 1;
        g_lo = r0 / rows_per_group;
        g_hi = r1 / rows_per_group;
      }
      const long long my_g = stats ? (row_ok ? row / rows_per_group : -1) : 0;
#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
        const int col0 = n_blk * BLOCK_N + chunk * 32;
        if (col0 >= Ncols) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + chunk * 32), r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (!D_F32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
        }
        if (row_ok) {
          if (D_F32) {
            float* dst = reinterpret_cast<float*>(Dptr) + row * ldd + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (col0 + j < Ncols)
                *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            bf16* dst = reinterpret_cast<bf16*>(Dptr) + row * ldd + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < Ncols) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&p0);
                pk.y = *reinterpret_cast<uint32_t*>(&p1);
                pk.z = *reinterpret_cast<uint32_t*>(&p2);
                pk.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(dst + j) = pk;
              }
            }
          }
        }
        if (stats) {
          for (long long g = g_lo; g <= g_hi; ++g) {
            // column sums over this warp's rows that belong to group g: recursive-halving
            // transpose-reduce, 31 shuffles per quantity; lane j ends with column j.
            float s[32], qq[32];
            const bool mine = (my_g == g);
#pragma unroll
            for (int j = 0; j < 32; ++j) { s[j] = mine ? v[j] : 0.f; qq[j] = s[j] * s[j]; }
#pragma unroll
            for (int half = 16; half >= 1; half >>= 1) {
              const bool upper = (lane & half) != 0;
#pragma unroll
              for (int j = 0; j < half; ++j) {
                // keep the half of the columns selected by this lane bit, send the other half
                float keep_s = upper ? s[j + half] : s[j];
                float send_s = upper ? s[j] : s[j + half];
                float keep_q = upper ? qq[j + half] : qq[j];
                float send_q = upper ? qq[j] : qq[j + half];
                s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, half);
                qq[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, half);
              }
            }
            // after the loop lane L holds column index L (bits assembled MSB-first)
            const int col = col0 + lane;
            if (col < Ncols) {
              atomicAdd(&stats[(g * Ncols + col) * 2 + 0], (double)s[0]);
              atomicAdd(&stats[(g * Ncols + col) * 2 + 1], (double)qq[0]);
            }
          }
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

// ---- host side ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

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

template <int BLOCK_N, bool D_F32>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, void* D, long long M, int Ncols, int K, long long ldd,
              double* stats, long long rpg, cudaStream_t stream) {
  using Cfg = TcCfg<BLOCK_N>;
  static bool configured = false;
  auto kern = tc_gemm_kernel<BLOCK_N, D_F32>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      adamml_set_error("tc_gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ADAMML_ERR_CUDA;
    }
    configured = true;
  }
  long long tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((Ncols + BLOCK_N - 1) / BLOCK_N);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = (int)(tiles < sms ? tiles : sms);
  kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, D, M, Ncols, K, ldd, stats, rpg);
  return adamml_check_launch("tc_gemm");
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

int adamml_tc_gemm_bf16(const void* A, const void* B, void* D, long long M, int Ncols, int K, long long lda,
                        long long ldb, long long ldd, int d_dtype, double* stats, long long rows_per_group,
                        cudaStream_t stream) {
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
  ADAMML_REQUIRE(d_dtype == ADAMML_F32 || d_dtype == ADAMML_BF16, "tc_gemm: bad output dtype");
  ADAMML_REQUIRE(!stats || rows_per_group > 0, "tc_gemm: stats need rows_per_group");
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
  const bool f32 = d_dtype == ADAMML_F32;
  if (block_n == 64)
    return f32 ? launch_tc<64, true>(tmA, tmB, D, M, Ncols, K, ldd, stats, rows_per_group, stream)
               : launch_tc<64, false>(tmA, tmB, D, M, Ncols, K, ldd, stats, rows_per_group, stream);
  if (block_n == 128)
    return f32 ? launch_tc<128, true>(tmA, tmB, D, M, Ncols, K, ldd, stats, rows_per_group, stream)
               : launch_tc<128, false>(tmA, tmB, D, M, Ncols, K, ldd, stats, rows_per_group, stream);
  return f32 ? launch_tc<256, true>(tmA, tmB, D, M, Ncols, K, ldd, stats, rows_per_group, stream)
             : launch_tc<256, false>(tmA, tmB, D, M, Ncols, K, ldd, stats, rows_per_group, stream);
}

}  // extern "C"
