// Device-side gating for inference with decision-driven skipping.
//
// The reference runs every main backbone on every (segment, video) pair and multiplies its logits by the policy's
// 0/1 decision (models/adamml.py:81-86, models/joint_resnet_mobilenetv2.py:92-94).  In inference mode (running-statistic
// BatchNorm, utils/utils.py:427-507) an unselected pair contributes exactly zero, so its backbone pass can be dropped.
// These kernels do the bookkeeping WITHOUT a host round trip: the decisions are compacted on the device into an
// ascending index list + a count, the selected clips are gathered to the front of a static-capacity batch buffer, every
// forward kernel limits itself to the live prefix by reading the count on the device (LiveLimit, common.cuh), and the
// logits are scattered back.  No D2H copy, no data-dependent launch geometry: the whole pass is one CUDA graph.
#include "common.cuh"

namespace {

constexpr int SEL_THREADS = 1024;

// dec [S][M][N] fp32 in {0, 1}; pair index p = s * N + n (the segment-major clip order of the batch buffers)
__global__ void __launch_bounds__(SEL_THREADS)
select_compact_kernel(const float* __restrict__ dec, int S, int M, int N, int m, int* __restrict__ idx,
                      int* __restrict__ count) {
  __shared__ int part[SEL_THREADS];
  const int SN = S * N;
  const int per = (SN + SEL_THREADS - 1) / SEL_THREADS;
  const int p0 = threadIdx.x * per;
  const int p1 = p0 + per < SN ? p0 + per : SN;
  int c = 0;
  for (int p = p0; p < p1; ++p) {
    const int s = p / N, n = p - s * N;
    c += dec[((long long)s * M + m) * N + n] > 0.f ? 1 : 0;
  }
  part[threadIdx.x] = c;
  __syncthreads();
  // inclusive scan over the per-thread counts (Hillis-Steele, 10 steps)
  for (int off = 1; off < SEL_THREADS; off <<= 1) {
    const int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int o = part[threadIdx.x] - c;
  const int total = part[SEL_THREADS - 1];
  for (int p = p0; p < p1; ++p) {
    const int s = p / N, n = p - s * N;
    if (dec[((long long)s * M + m) * N + n] > 0.f) idx[o++] = p;
  }
  // unused tail: valid indices, never gathered
  for (int j = total + threadIdx.x; j < SN; j += SEL_THREADS) idx[j] = 0;
  if (threadIdx.x == 0) *count = total;
}

// dst row j = src row idx[j] for j < *count; rows of row_vecs 16-byte vectors
__global__ void gather_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ idx,
                                   const int* __restrict__ count, long long row_vecs) {
  const long long total = (long long)(*count) * row_vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long j = i / row_vecs, v = i - j * row_vecs;
    dst[i] = src[(long long)idx[j] * row_vecs + v];
  }
}

// out [rows_out][C] was zeroed; out[idx[j]] = y[j] for j < *count
__global__ void scatter_rows_kernel(const float* __restrict__ y, const int* __restrict__ idx,
                                    const int* __restrict__ count, float* __restrict__ out, int C) {
  const long long total = (long long)(*count) * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long j = i / C;
    out[(long long)idx[j] * C + (i - j * C)] = y[i];
  }
}

}  // namespace

extern "C" {

int adamml_select_compact(const float* decisions, int S, int M, int N, int m, int* idx, int* count,
                          cudaStream_t stream) {
  ADAMML_REQUIRE(decisions && idx && count && S > 0 && M > 0 && N > 0 && m >= 0 && m < M, "select_compact: bad arguments");
  select_compact_kernel<<<1, SEL_THREADS, 0, stream>>>(decisions, S, M, N, m, idx, count);
  return adamml_check_launch("select_compact");
}

int adamml_gather_rows(const void* src, void* dst, const int* idx, const int* count, long long row_bytes, int capacity,
                       cudaStream_t stream) {
  ADAMML_REQUIRE(src && dst && idx && count && capacity > 0, "gather_rows: bad arguments");
  ADAMML_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0,
                 "gather_rows: rows must be 16-byte multiples, buffers 16-byte aligned");
  const long long row_vecs = row_bytes / 16;
  long long blocks = ((long long)capacity * row_vecs + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const uint4*)src, (uint4*)dst, idx, count, row_vecs);
  return adamml_check_launch("gather_rows");
}

int adamml_scatter_rows_f32(const float* y, const int* idx, const int* count, float* out, int rows_out, int C,
                            cudaStream_t stream) {
  ADAMML_REQUIRE(y && idx && count && out && rows_out > 0 && C > 0, "scatter_rows: bad arguments");
  cudaMemsetAsync(out, 0, sizeof(float) * (size_t)rows_out * C, stream);
  long long blocks = ((long long)rows_out * C + 255) / 256;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  scatter_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(y, idx, count, out, C);
  return adamml_check_launch("scatter_rows");
}

}  // extern "C"
