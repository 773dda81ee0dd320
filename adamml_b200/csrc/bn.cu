// Batch-norm statistics / apply / backward for NHWC activations with per-segment groups.
//
// Reference semantics: every nn.BatchNorm2d on the AdaMML path (models/resnet.py:92-111,
// models/sound_mobilenet_v2.py:33-40, models/policy_net.py:38-52,63-86) is called once per
// segment (models/adamml.py:84-86, models/policy_net.py:323-326), i.e. batch statistics
// are taken over ONE segment's images.  Here all S segments are batched into one launch:
// images are ordered segment-major, group g = img / imgs_per_group, statistics are kept per
// (group, channel) and running stats receive the S momentum updates in segment order.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int ROWS_PER_BLOCK = 256;

// sums[g][c][0] += sum z ; sums[g][c][1] += sum z^2     (double accumulation)
template <typename T>
__global__ void bn_stats_kernel(const T* __restrict__ z, double* __restrict__ sums, long long rows_per_group, int C,
                                int blocks_per_group) {
  __shared__ double sh[2][8][33];
  int g = blockIdx.x / blocks_per_group;
  int bg = blockIdx.x % blocks_per_group;
  int c = blockIdx.y * 32 + threadIdx.x;
  long long r0 = (long long)bg * ROWS_PER_BLOCK;
  long long r1 = r0 + ROWS_PER_BLOCK;
  if (r1 > rows_per_group) r1 = rows_per_group;
  double s = 0.0, q = 0.0;
  if (c < C) {
    const T* base = z + (long long)g * rows_per_group * C + c;
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      double v = (double)to_f32(base[r * C]);
      s += v;
      q += v * v;
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { s += sh[0][i][threadIdx.x]; q += sh[1][i][threadIdx.x]; }
    atomicAdd(&sums[((long long)g * C + c) * 2 + 0], s);
    atomicAdd(&sums[((long long)g * C + c) * 2 + 1], q);
  }
}

// One thread per channel; loops groups in order so running stats see S sequential updates.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean_invstd,
                                   float* __restrict__ scale_shift, double count, float momentum, float eps, int C,
                                   int G, int training, int update_running) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float ga = gamma ? gamma[c] : 1.f;
  float be = beta ? beta[c] : 0.f;
  if (training) {
    float rm = running_mean ? running_mean[c] : 0.f;
    float rv = running_var ? running_var[c] : 1.f;
    for (int g = 0; g < G; ++g) {
      double s = sums[((long long)g * C + c) * 2 + 0];
      double q = sums[((long long)g * C + c) * 2 + 1];
      double mean = s / count;
      double var = q / count - mean * mean;
      if (var < 0.0) var = 0.0;
      float invstd = (float)(1.0 / sqrt(var + (double)eps));
      float meanf = (float)mean;
      mean_invstd[((long long)g * C + c) * 2 + 0] = meanf;
      mean_invstd[((long long)g * C + c) * 2 + 1] = invstd;
      float sc = ga * invstd;
      scale_shift[((long long)g * C + c) * 2 + 0] = sc;
      scale_shift[((long long)g * C + c) * 2 + 1] = be - meanf * sc;
      double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      rm = (1.f - momentum) * rm + momentum * meanf;
      rv = (1.f - momentum) * rv + momentum * (float)unbiased;
    }
    if (update_running && running_mean) running_mean[c] = rm;
    if (update_running && running_var) running_var[c] = rv;
  } else {
    float rm = running_mean[c];
    float rv = running_var[c];
    float invstd = 1.f / sqrtf(rv + eps);
    float sc = ga * invstd;
    for (int g = 0; g < G; ++g) {
      mean_invstd[((long long)g * C + c) * 2 + 0] = rm;
      mean_invstd[((long long)g * C + c) * 2 + 1] = invstd;
      scale_shift[((long long)g * C + c) * 2 + 0] = sc;
      scale_shift[((long long)g * C + c) * 2 + 1] = be - rm * sc;
    }
  }
}

// out = act( z*scale+shift  [+ res]  [+ res_z*res_scale+res_shift] )
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ z, const float* __restrict__ ss, const T* __restrict__ res,
                                const T* __restrict__ res_z, const float* __restrict__ res_ss, T* __restrict__ out,
                                long long total, long long elems_per_group, int C, int act) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    int g = (int)(idx / elems_per_group);
    const float* s = ss + ((long long)g * C + c) * 2;
    float v = fmaf(to_f32(z[idx]), s[0], s[1]);
    if (res) v += to_f32(res[idx]);
    if (res_z) {
      const float* rs = res_ss + ((long long)g * C + c) * 2;
      v += fmaf(to_f32(res_z[idx]), rs[0], rs[1]);
    }
    out[idx] = from_f32<T>(act_apply(v, act));
  }
}

// sums[g][c][0] += sum gm ; sums[g][c][1] += sum gm * xhat, gm = dout * mask(out)
template <typename T>
__global__ void bn_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ z,
                                     const float* __restrict__ mean_invstd, double* __restrict__ sums,
                                     long long rows_per_group, int C, int blocks_per_group, int act) {
  __shared__ double sh[2][8][33];
  int g = blockIdx.x / blocks_per_group;
  int bg = blockIdx.x % blocks_per_group;
  int c = blockIdx.y * 32 + threadIdx.x;
  long long r0 = (long long)bg * ROWS_PER_BLOCK;
  long long r1 = r0 + ROWS_PER_BLOCK;
  if (r1 > rows_per_group) r1 = rows_per_group;
  double s = 0.0, q = 0.0;
  if (c < C) {
    float mean = mean_invstd[((long long)g * C + c) * 2 + 0];
    float invstd = mean_invstd[((long long)g * C + c) * 2 + 1];
    long long base = (long long)g * rows_per_group * C + c;
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      long long i = base + r * C;
      float gm = to_f32(dout[i]);
      if (act != ADAMML_ACT_NONE && !act_pass(to_f32(out[i]), act)) gm = 0.f;
      float xhat = (to_f32(z[i]) - mean) * invstd;
      s += (double)gm;
      q += (double)gm * (double)xhat;
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { s += sh[0][i][threadIdx.x]; q += sh[1][i][threadIdx.x]; }
    atomicAdd(&sums[((long long)g * C + c) * 2 + 0], s);
    atomicAdd(&sums[((long long)g * C + c) * 2 + 1], q);
  }
}

// training: dz = gamma*invstd*(gm - sum_g/cnt - xhat*sum_gx/cnt);  eval: dz = gamma*invstd*gm
// dres (optional) = gm
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ z,
                                    const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
                                    const double* __restrict__ sums, T* __restrict__ dz, T* __restrict__ dres,
                                    long long total, long long elems_per_group, int C, double count, int act,
                                    int training) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    int g = (int)(idx / elems_per_group);
    long long gc = (long long)g * C + c;
    float gm = to_f32(dout[idx]);
    if (act != ADAMML_ACT_NONE && !act_pass(to_f32(out[idx]), act)) gm = 0.f;
    if (dres) dres[idx] = from_f32<T>(gm);
    if (dz) {
      float mean = mean_invstd[gc * 2 + 0];
      float invstd = mean_invstd[gc * 2 + 1];
      float ga = gamma ? gamma[c] : 1.f;
      float v;
      if (training) {
        float xhat = (to_f32(z[idx]) - mean) * invstd;
        float m1 = (float)(sums[gc * 2 + 0] / count);
        float m2 = (float)(sums[gc * 2 + 1] / count);
        v = ga * invstd * (gm - m1 - xhat * m2);
      } else {
        v = ga * invstd * gm;
      }
      dz[idx] = from_f32<T>(v);
    }
  }
}

__global__ void bn_param_grad_kernel(const double* __restrict__ sums, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int C, int G, int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int g = 0; g < G; ++g) {
    s += sums[((long long)g * C + c) * 2 + 0];
    q += sums[((long long)g * C + c) * 2 + 1];
  }
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)q;
}


// ---------------------------------------------------------------------------------------------------
// Row-streaming kernels (C % VEC == 0, 16-byte aligned): the [rows, C] NHWC matrix is contiguous, a block
// of cpb*k threads walks it so that every thread keeps ONE channel vector (VEC = 8 bf16 / 4 fp32 channels)
// for its whole life: per-channel coefficients live in registers, there is no index arithmetic in the
// loop, and consecutive threads touch consecutive 16-byte vectors (fully coalesced).  The kernels are
// pure HBM streams: algorithmic bytes = (tensors read + tensors written) x rows x C x sizeof(T).
constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;
template <typename T> struct AccT { typedef float type; };
template <> struct AccT<float> { typedef double type; };

struct RowGeom {
  int cpb;      // channel vectors per block
  int k;        // row lanes per block
  int threads;  // cpb * k
  int cchunks;  // blocks along the channel axis
};
template <typename T>
inline RowGeom row_geom(int C) {
  const int V = VecIO<T>::N;
  RowGeom g;
  int cvecs = C / V;
  g.cchunks = (cvecs + EW_THREADS - 1) / EW_THREADS;
  g.cpb = (cvecs + g.cchunks - 1) / g.cchunks;
  g.k = EW_THREADS / g.cpb;
  if (g.k < 1) g.k = 1;
  g.threads = g.cpb * g.k;
  return g;
}
inline int num_sms_ew() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// out = act( z*scale+shift  [+ res]  [+ res_z*res_scale+res_shift] )
// RESZ = false: compiled without the BN(downsample) residual branch (its 16 coefficient registers)
template <typename T, typename CP, typename MP, bool RESZ = true>
__device__ __forceinline__ void
bn_apply_rows_body(CP z, const float* __restrict__ ss, CP res, CP res_z_, const float* __restrict__ res_ss, MP out,
                   long long rows_per_group, int C, int cpb, int k, int rows_per_block, int blocks_per_group,
                   int act, unsigned char* __restrict__ mask_bits = nullptr) {
  constexpr int V = VecIO<T>::N;
  const int cl = threadIdx.x % cpb, rl = threadIdx.x / cpb;
  const int c0 = (blockIdx.y * cpb + cl) * V;
  if (c0 >= C) return;
  const int g = blockIdx.x / blocks_per_group;
  const int bg = blockIdx.x % blocks_per_group;
  const long long r0 = (long long)bg * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows_per_group) r1 = rows_per_group;
  const CP res_z = RESZ ? res_z_ : CP{};
  float sc[V], sh[V], rsc[RESZ ? V : 1], rsh[RESZ ? V : 1];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float2 p = *reinterpret_cast<const float2*>(ss + ((long long)g * C + c0 + i) * 2);
    sc[i] = p.x; sh[i] = p.y;
    if (RESZ) { rsc[i] = 0.f; rsh[i] = 0.f; }
  }
  if (RESZ && res_z) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float2 p = *reinterpret_cast<const float2*>(res_ss + ((long long)g * C + c0 + i) * 2);
      rsc[i] = p.x; rsh[i] = p.y;
    }
  }
  const long long base = (long long)g * rows_per_group * C + c0;
  for (long long r = r0 + rl; r < r1; r += (long long)k * EW_UNROLL) {
    typename VecIO<T>::raw qz[EW_UNROLL], qr[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const long long rr = r + (long long)u * k;
      if (rr < r1) {
        qz[u] = VecIO<T>::load_raw(z + base + rr * C);
        if (res) qr[u] = VecIO<T>::load_raw(res + base + rr * C);
        else if (res_z) qr[u] = VecIO<T>::load_raw(res_z + base + rr * C);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const long long rr = r + (long long)u * k;
      if (rr < r1) {
        float vo[V], vz[V], vr[V];
        VecIO<T>::unpack(qz[u], vz);
        if (res || res_z) VecIO<T>::unpack(qr[u], vr);
        unsigned m = 0;  // 1-bit activation mask of the 8 channels: what the backward reduction of a residual layer
                         // reads instead of streaming `out` again
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float v = fmaf(vz[i], sc[i], sh[i]);
          if (res) v += vr[i];
          else if (RESZ && res_z) v += fmaf(vr[i], rsc[RESZ ? i : 0], rsh[RESZ ? i : 0]);
          vo[i] = act_apply(v, act);
          m |= act_pass(vo[i], act) ? (1u << i) : 0u;
        }
        VecIO<T>::store(out + base + rr * C, vo);
        if (V == 8 && mask_bits) mask_bits[(base + rr * C) >> 3] = (unsigned char)m;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
bn_apply_rows_kernel(const T* __restrict__ z, const float* __restrict__ ss, const T* __restrict__ res,
                     const T* __restrict__ res_z, const float* __restrict__ res_ss, T* __restrict__ out,
                     long long rows_per_group, int C, int cpb, int k, int rows_per_block, int blocks_per_group,
                     int act) {
  bn_apply_rows_body<T, const T*, T*>(z, ss, res, res_z, res_ss, out, rows_per_group, C, cpb, k, rows_per_block,
                                      blocks_per_group, act);
}
// x2 planes (forward pass of the default precision mode)
template <bool RESZ>
__global__ void __launch_bounds__(EW_THREADS, 2)
bn_apply_rows_x2_kernel(X2CPtr z, const float* __restrict__ ss, X2CPtr res, X2CPtr res_z,
                        const float* __restrict__ res_ss, X2Ptr out, long long rows_per_group, int C, int cpb, int k,
                        int rows_per_block, int blocks_per_group, int act, unsigned char* __restrict__ mask_bits) {
  bn_apply_rows_body<x2_t, X2CPtr, X2Ptr, RESZ>(z, ss, res, res_z, res_ss, out, rows_per_group, C, cpb, k,
                                                rows_per_block, blocks_per_group, act, mask_bits);
}

// sums[g][c][0] += sum z ; sums[g][c][1] += sum z^2 over x2 planes (depthwise-conv outputs: their BN statistics are
// not produced by a GEMM epilogue).  Same persistent row-chunk scheme as bn_reduce_rows_kernel<MODE 0>.
__global__ void __launch_bounds__(EW_THREADS, 2)
bn_stats_rows_x2_kernel(X2CPtr a, double* __restrict__ sums, long long rows_per_group, int C, int cpb, int k,
                        int blocks_per_group) {
  constexpr int V = 8;
  constexpr int UN = 4;  // 8 independent 16-byte loads in flight per thread: one tensor only, so the memory-level
  constexpr int CH = 4;  // parallelism has to come from the unroll (2 -> 4: measured below)
  extern __shared__ double shd[];
  const int cl = threadIdx.x % cpb, rl = threadIdx.x / cpb;
  const int c0 = (blockIdx.y * cpb + cl) * V;
  const bool ok = c0 < C;
  const int g = blockIdx.x / blocks_per_group;
  const int bg = blockIdx.x % blocks_per_group;
  double* sh = shd + (size_t)threadIdx.x * (2 * V);
#pragma unroll
  for (int i = 0; i < 2 * V; ++i) sh[i] = 0.0;
  if (ok) {
    const long long base = (long long)g * rows_per_group * C + c0;
    const long long chunk_rows = (long long)k * UN * CH;
    for (long long rc = (long long)bg * chunk_rows; rc < rows_per_group; rc += (long long)blocks_per_group * chunk_rows) {
      long long r1 = rc + chunk_rows;
      if (r1 > rows_per_group) r1 = rows_per_group;
      float s[V], q[V];
#pragma unroll
      for (int i = 0; i < V; ++i) { s[i] = 0.f; q[i] = 0.f; }
      for (long long r = rc + rl; r < r1; r += (long long)k * UN) {
        X2Raw qa[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const long long rr = r + (long long)u * k;
          if (rr < r1) qa[u] = VecIO<x2_t>::load_raw(a + (base + rr * C));
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const long long rr = r + (long long)u * k;
          if (rr < r1) {
            float va[V];
            VecIO<x2_t>::unpack(qa[u], va);
#pragma unroll
            for (int i = 0; i < V; ++i) { s[i] += va[i]; q[i] = fmaf(va[i], va[i], q[i]); }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) { sh[i] += (double)s[i]; sh[V + i] += (double)q[i]; }
    }
  }
  __syncthreads();
  if (rl == 0 && ok) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      double ds = 0.0, dq = 0.0;
      for (int y = 0; y < k; ++y) {
        const double* o = shd + ((size_t)y * cpb + cl) * (2 * V);
        ds += o[i];
        dq += o[V + i];
      }
      atomicAdd(&sums[((long long)g * C + c0 + i) * 2 + 0], ds);
      atomicAdd(&sums[((long long)g * C + c0 + i) * 2 + 1], dq);
    }
  }
}

// MODE 0: sum z, sum z^2.   MODE 1: sum gm, sum gm*xhat (gm = dout * mask(out)).
// Persistent over row chunks: block b of a group handles chunks b, b+bpg, ...; per-thread partial sums
// in fp32 over <= 16 rows are flushed into wide accumulators, one smem reduction + 2 atomics per
// (block, channel) at the end.
template <typename T, int MODE, bool MASKZ>
__global__ void __launch_bounds__(EW_THREADS, 2)
bn_reduce_rows_kernel(const T* a /*z | dout (may alias gm_out)*/, const T* __restrict__ out, const T* __restrict__ z,
                      const float* __restrict__ mean_invstd, const float* __restrict__ mss,
                      double* __restrict__ sums, T* gm_out, long long rows_per_group, int C, int cpb, int k,
                      int blocks_per_group, int act, const unsigned char* __restrict__ mask_bits = nullptr) {
  constexpr int V = VecIO<T>::N;
  // the wide (fp64) accumulators live in the thread's shared-memory slot, not in registers: that leaves room for
  // 4 independent 16-byte loads per tensor in flight at <= 128 registers (2 CTAs / SM) — the kernel is
  // memory-level-parallelism bound
  constexpr int UN = EW_UNROLL;
  constexpr int CH = 16 / UN;  // row iterations per flush
  extern __shared__ double shd[];
  const int cl = threadIdx.x % cpb, rl = threadIdx.x / cpb;
  const int c0 = (blockIdx.y * cpb + cl) * V;
  const bool ok = c0 < C;
  const int g = blockIdx.x / blocks_per_group;
  const int bg = blockIdx.x % blocks_per_group;
  typedef typename AccT<T>::type acc_t;
  double* sh = shd + (size_t)threadIdx.x * (2 * V);
#pragma unroll
  for (int i = 0; i < 2 * V; ++i) sh[i] = 0.0;
  if (ok) {
    float msc[V], msh[V];  // forward scale/shift: the activation mask is recomputed from z instead of reading `out`
#pragma unroll
    for (int i = 0; i < V; ++i) { msc[i] = 0.f; msh[i] = 0.f; }
    if (MODE == 1) {
      if (MASKZ) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float2 p = *reinterpret_cast<const float2*>(mss + ((long long)g * C + c0 + i) * 2);
          msc[i] = p.x; msh[i] = p.y;
        }
      }
    }
    const long long base = (long long)g * rows_per_group * C + c0;
    const long long chunk_rows = (long long)k * UN * CH;
    for (long long rc = (long long)bg * chunk_rows; rc < rows_per_group; rc += (long long)blocks_per_group * chunk_rows) {
      long long r1 = rc + chunk_rows;
      if (r1 > rows_per_group) r1 = rows_per_group;
      acc_t s[V], q[V];
#pragma unroll
      for (int i = 0; i < V; ++i) { s[i] = 0; q[i] = 0; }
      for (long long r = rc + rl; r < r1; r += (long long)k * UN) {
        typename VecIO<T>::raw qa[UN], qz[UN], qo[UN];
        unsigned mb[UN];  // 1-bit masks written by the forward bn_apply (residual layers): 1 byte instead of 16
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const long long rr = r + (long long)u * k;
          if (rr < r1) {
            qa[u] = VecIO<T>::load_raw(a + base + rr * C);
            if (MODE == 1) {
              qz[u] = VecIO<T>::load_raw(z + base + rr * C);
              if (act != ADAMML_ACT_NONE && !MASKZ) {
                if (V == 8 && mask_bits) mb[u] = mask_bits[(base + rr * C) >> 3];
                else qo[u] = VecIO<T>::load_raw(out + base + rr * C);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const long long rr = r + (long long)u * k;
          if (rr < r1) {
            float va[V], vz[V], vo[V];
            VecIO<T>::unpack(qa[u], va);
            if (MODE == 1) {
              VecIO<T>::unpack(qz[u], vz);
              if (act != ADAMML_ACT_NONE) {
                if (MASKZ) {
#pragma unroll
                  for (int i = 0; i < V; ++i) vo[i] = fmaf(vz[i], msc[i], msh[i]);
                } else if (V == 8 && mask_bits) {  // (a value act_pass accepts for ReLU and ReLU6 alike | rejects)
#pragma unroll
                  for (int i = 0; i < V; ++i) vo[i] = ((mb[u] >> i) & 1u) ? 1.f : 0.f;
                } else {
                  VecIO<T>::unpack(qo[u], vo);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {
              if (MODE == 0) {
                s[i] += (acc_t)va[i];
                q[i] += (acc_t)va[i] * (acc_t)va[i];
              } else {
                float gm = va[i];
                if (act != ADAMML_ACT_NONE && !act_pass(vo[i], act)) gm = 0.f;
                va[i] = gm;
                s[i] += (acc_t)gm;
                q[i] += (acc_t)gm * (acc_t)vz[i];  // sum gm*z; turned into sum gm*xhat at the end (fp64)
              }
            }
            if (MODE == 1 && gm_out) VecIO<T>::store(gm_out + base + rr * C, va);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) { sh[i] += (double)s[i]; sh[V + i] += (double)q[i]; }
    }
  }
  // block reduce over the k row lanes
  __syncthreads();
  if (rl == 0 && ok) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      double ds = 0.0, dq = 0.0;
      for (int y = 0; y < k; ++y) {
        const double* o = shd + ((size_t)y * cpb + cl) * (2 * V);
        ds += o[i];
        dq += o[V + i];
      }
      if (MODE == 1) {  // sum gm*xhat = invstd * (sum gm*z - mean * sum gm)
        const float2 p = *reinterpret_cast<const float2*>(mean_invstd + ((long long)g * C + c0 + i) * 2);
        dq = (double)p.y * (dq - (double)p.x * ds);
      }
      atomicAdd(&sums[((long long)g * C + c0 + i) * 2 + 0], ds);
      atomicAdd(&sums[((long long)g * C + c0 + i) * 2 + 1], dq);
    }
  }
}

// training: dz = gamma*invstd*(gm - sum_g/cnt - xhat*sum_gx/cnt);  eval: dz = gamma*invstd*gm ; dres = gm
template <typename T, bool MASKZ>
__global__ void __launch_bounds__(EW_THREADS, 2)
bn_bwd_apply_rows_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ z,
                         const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
                         const float* __restrict__ mss, const double* __restrict__ sums, T* __restrict__ dz,
                         T* __restrict__ dres, long long rows_per_group, int C, int cpb, int k, int rows_per_block,
                         int blocks_per_group, double count, int act, int training) {
  constexpr int V = VecIO<T>::N;
  const int cl = threadIdx.x % cpb, rl = threadIdx.x / cpb;
  const int c0 = (blockIdx.y * cpb + cl) * V;
  if (c0 >= C) return;
  const int g = blockIdx.x / blocks_per_group;
  const int bg = blockIdx.x % blocks_per_group;
  const long long r0 = (long long)bg * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows_per_group) r1 = rows_per_group;
  const float inv_count = (float)(1.0 / count);
  float mean[V], invstd[V], A[V], m1[V], m2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const long long gc = (long long)g * C + c0 + i;
    const float2 p = *reinterpret_cast<const float2*>(mean_invstd + gc * 2);
    mean[i] = p.x; invstd[i] = p.y;
    A[i] = (gamma ? gamma[c0 + i] : 1.f) * p.y;
    m1[i] = 0.f; m2[i] = 0.f;
    if (training && dz) {
      m1[i] = (float)sums[gc * 2 + 0] * inv_count;
      m2[i] = (float)sums[gc * 2 + 1] * inv_count;
    }
  }
  float msc[V], msh[V];  // forward scale/shift (MASKZ): mask = act'(z*scale+shift)
#pragma unroll
  for (int i = 0; i < V; ++i) { msc[i] = 0.f; msh[i] = 0.f; }
  if (MASKZ) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float2 p = *reinterpret_cast<const float2*>(mss + ((long long)g * C + c0 + i) * 2);
      msc[i] = p.x; msh[i] = p.y;
    }
  }
  const long long base = (long long)g * rows_per_group * C + c0;
  const bool need_z = (training && dz) || (MASKZ && act != ADAMML_ACT_NONE);
  for (long long r = r0 + rl; r < r1; r += (long long)k * EW_UNROLL) {
    typename VecIO<T>::raw qg[EW_UNROLL], qz[EW_UNROLL], qo[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const long long rr = r + (long long)u * k;
      if (rr < r1) {
        qg[u] = VecIO<T>::load_raw(dout + base + rr * C);
        if (act != ADAMML_ACT_NONE && !MASKZ) qo[u] = VecIO<T>::load_raw(out + base + rr * C);
        if (need_z) qz[u] = VecIO<T>::load_raw(z + base + rr * C);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const long long rr = r + (long long)u * k;
      if (rr < r1) {
        float gm[V], vz[V], vo[V];
        VecIO<T>::unpack(qg[u], gm);
        if (need_z) VecIO<T>::unpack(qz[u], vz);
        if (act != ADAMML_ACT_NONE) {
          if (MASKZ) {
#pragma unroll
            for (int i = 0; i < V; ++i) vo[i] = fmaf(vz[i], msc[i], msh[i]);
          } else {
            VecIO<T>::unpack(qo[u], vo);
          }
#pragma unroll
          for (int i = 0; i < V; ++i) if (!act_pass(vo[i], act)) gm[i] = 0.f;
        }
        if (dres) VecIO<T>::store(dres + base + rr * C, gm);
        if (dz) {
          float v[V];
          if (training) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
              const float xhat = (vz[i] - mean[i]) * invstd[i];
              v[i] = A[i] * (gm[i] - m1[i] - xhat * m2[i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = A[i] * gm[i];
          }
          VecIO<T>::store(dz + base + rr * C, v);
        }
      }
    }
  }
}

template <typename T>
inline bool vec_ok(int C, const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr,
                   const void* e = nullptr) {
  if (C % VecIO<T>::N) return false;
  const void* ps[5] = {a, b, c, d, e};
  for (int i = 0; i < 5; ++i)
    if (ps[i] && ((uintptr_t)ps[i] % 16)) return false;
  return true;
}

// streaming launch: ~32 vector iterations per thread
inline void stream_geom(const RowGeom& rg, long long rows_per_group, int* rows_per_block, int* bpg) {
  long long rpb = (long long)rg.k * EW_UNROLL * 8;
  *bpg = (int)((rows_per_group + rpb - 1) / rpb);
  *rows_per_block = (int)rpb;
}
// persistent reduction launch: exactly ONE wave of resident blocks.  (The grid used to be a fixed 4 blocks per SM; at
// 128 registers only 2 x 256 threads are resident per SM, so 595 equal-work blocks ran as 296 + 296 + 3: a third round
// for three blocks.)  `kern` / smem: the kernel about to be launched, for the occupancy query.
template <typename K>
inline int reduce_bpg(K kern, size_t smem, const RowGeom& rg, long long rows_per_group, int G, int chunk_iters = 4) {
  long long chunk_rows = (long long)rg.k * EW_UNROLL * chunk_iters;
  long long chunks = (rows_per_group + chunk_rows - 1) / chunk_rows;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, rg.threads, smem) != cudaSuccess || per_sm < 1)
    per_sm = 2;
  static const int forced = []() { const char* e = getenv("ADAMML_B200_BN_REDUCE_BPSM"); return e ? atoi(e) : 0; }();
  if (forced > 0) per_sm = forced;  // (experiments: 4 = the former fixed grid)
  long long want = ((long long)per_sm * num_sms_ew()) / ((long long)G * rg.cchunks);
  if (want < 1) want = 1;
  return (int)(chunks < want ? chunks : want);
}

// (sum gm, sum gm * out) -> (sum gm, sum gm * xhat) for a layer out = act(z * scale + shift): wherever the activation
// passes the gradient, out IS z * scale + shift, so sum gm * z = (sum gm * out - shift * sum gm) / scale.
__global__ void bn_sums_from_out_kernel(const double* __restrict__ raw, const float* __restrict__ ss,
                                        const float* __restrict__ mean_invstd, double* __restrict__ sums, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double A = raw[2 * i], B = raw[2 * i + 1];
  const double sc = (double)ss[2 * i], sh = (double)ss[2 * i + 1];
  const double mean = (double)mean_invstd[2 * i], invstd = (double)mean_invstd[2 * i + 1];
  // scale == 0 (gamma == 0): the output carries no information about z; dz = gamma * (...) is zero anyway
  const double gz = sc != 0.0 ? (B - sh * A) / sc : mean * A;
  sums[2 * i] = A;
  sums[2 * i + 1] = invstd * (gz - mean * A);
}

inline int ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  long long cap = 148LL * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" {

// z: [G*rows_per_group, C]; sums: double [G][C][2], zeroed here.
int adamml_bn_stats(const void* z, double* sums, long long rows_per_group, int C, int G, int dtype,
                    cudaStream_t stream) {
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_stats: empty dims");
  cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)G * C * 2, stream);
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (vec_ok<T>(C, z)) {
      const RowGeom rg = row_geom<T>(C);
      const size_t sm = sizeof(double) * rg.threads * 2 * VecIO<T>::N;
      const int bpg = reduce_bpg(bn_reduce_rows_kernel<T, 0, false>, sm, rg, rows_per_group, G);
      dim3 vg((unsigned)(bpg * (long long)G), rg.cchunks);
      bn_reduce_rows_kernel<T, 0, false><<<vg, rg.threads, sm, stream>>>((const T*)z, nullptr, nullptr, nullptr, nullptr, sums,
                                                                  nullptr, rows_per_group, C, rg.cpb, rg.k, bpg, 0);
    } else {
      int bpg = ceil_div(rows_per_group, ROWS_PER_BLOCK);
      dim3 grid((unsigned)(bpg * (long long)G), ceil_div(C, 32));
      dim3 block(32, 8);
      bn_stats_kernel<T><<<grid, block, 0, stream>>>((const T*)z, sums, rows_per_group, C, bpg);
    }
  });
  return adamml_check_launch("bn_stats");
}

int adamml_bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float* mean_invstd, float* scale_shift, double count, float momentum,
                       float eps, int C, int G, int training, int update_running, cudaStream_t stream) {
  ADAMML_REQUIRE(C > 0 && G > 0, "bn_finalize: empty dims");
  ADAMML_REQUIRE(training || (running_mean && running_var), "bn_finalize: eval mode needs running stats");
  ADAMML_REQUIRE(!training || sums, "bn_finalize: training mode needs sums");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(sums, gamma, beta, running_mean, running_var, mean_invstd,
                                                          scale_shift, count, momentum, eps, C, G, training,
                                                          update_running);
  return adamml_check_launch("bn_finalize");
}

int adamml_bn_apply(const void* z, const float* scale_shift, const void* res, const void* res_z,
                    const float* res_scale_shift, void* out, long long rows_per_group, int C, int G, int act,
                    int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_apply: empty dims");
  ADAMML_REQUIRE(!res_z || res_scale_shift, "bn_apply: res_z needs res_scale_shift");
  long long epg = rows_per_group * C;
  long long total = epg * G;
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (vec_ok<T>(C, z, res, res_z, out)) {
      const RowGeom rg = row_geom<T>(C);
      int rpb, bpg;
      stream_geom(rg, rows_per_group, &rpb, &bpg);
      dim3 vg((unsigned)(bpg * (long long)G), rg.cchunks);
      bn_apply_rows_kernel<T><<<vg, rg.threads, 0, stream>>>((const T*)z, scale_shift, (const T*)res, (const T*)res_z,
                                                             res_scale_shift, (T*)out, rows_per_group, C, rg.cpb, rg.k,
                                                             rpb, bpg, act);
    } else {
      bn_apply_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)z, scale_shift, (const T*)res,
                                                              (const T*)res_z, res_scale_shift, (T*)out, total, epg, C,
                                                              act);
    }
  });
  return adamml_check_launch("bn_apply");
}

/* x2 planes: z/res/res_z/out are (hi, lo) plane pairs; C must be a multiple of 8 */
int adamml_bn_apply_x2(const void* z_hi, const void* z_lo, const float* scale_shift, const void* res_hi,
                       const void* res_lo, const void* resz_hi, const void* resz_lo, const float* res_scale_shift,
                       void* out_hi, void* out_lo, long long rows_per_group, int C, int G, int act,
                       unsigned char* mask_bits, cudaStream_t stream) {
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_apply_x2: empty dims");
  ADAMML_REQUIRE(!resz_hi || res_scale_shift, "bn_apply_x2: res_z needs res_scale_shift");
  ADAMML_REQUIRE(z_hi && z_lo && out_hi && out_lo && (!res_hi == !res_lo) && (!resz_hi == !resz_lo),
                 "bn_apply_x2: every tensor needs both planes");
  ADAMML_REQUIRE(C % 8 == 0 && vec_ok<bf16>(C, z_hi, z_lo, res_hi, res_lo) && vec_ok<bf16>(C, resz_hi, resz_lo, out_hi, out_lo),
                 "bn_apply_x2: needs C %% 8 == 0 and 16-byte aligned planes");
  const RowGeom rg = row_geom<x2_t>(C);
  int rpb, bpg;
  stream_geom(rg, rows_per_group, &rpb, &bpg);
  dim3 vg((unsigned)(bpg * (long long)G), rg.cchunks);
  if (resz_hi)
    bn_apply_rows_x2_kernel<true><<<vg, rg.threads, 0, stream>>>(
        x2c(z_hi, z_lo), scale_shift, x2c(res_hi, res_lo), x2c(resz_hi, resz_lo), res_scale_shift, x2m(out_hi, out_lo),
        rows_per_group, C, rg.cpb, rg.k, rpb, bpg, act, mask_bits);
  else
    bn_apply_rows_x2_kernel<false><<<vg, rg.threads, 0, stream>>>(
        x2c(z_hi, z_lo), scale_shift, x2c(res_hi, res_lo), x2c(resz_hi, resz_lo), res_scale_shift, x2m(out_hi, out_lo),
        rows_per_group, C, rg.cpb, rg.k, rpb, bpg, act, mask_bits);
  return adamml_check_launch("bn_apply_x2");
}

int adamml_bn_stats_x2(const void* z_hi, const void* z_lo, double* sums, long long rows_per_group, int C, int G,
                       cudaStream_t stream) {
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_stats_x2: empty dims");
  ADAMML_REQUIRE(C % 8 == 0 && vec_ok<bf16>(C, z_hi, z_lo), "bn_stats_x2: needs C %% 8 == 0 and aligned planes");
  cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)G * C * 2, stream);
  const RowGeom rg = row_geom<x2_t>(C);
  // persistent over row chunks, ~8 blocks per SM in total (a single-tensor stream needs more resident warps than the
  // three-tensor backward reduction to cover the HBM latency)
  const size_t sm = sizeof(double) * rg.threads * 2 * 8;
  const int bpg = reduce_bpg(bn_stats_rows_x2_kernel, sm, rg, rows_per_group, G);  // (chunk = k * 4 * 4 rows)
  dim3 vg((unsigned)(bpg * (long long)G), rg.cchunks);
  bn_stats_rows_x2_kernel<<<vg, rg.threads, sm, stream>>>(x2c(z_hi, z_lo), sums, rows_per_group, C, rg.cpb, rg.k, bpg);
  return adamml_check_launch("bn_stats_x2");
}

int adamml_bn_bwd_reduce(const void* dout, const void* out, const void* z, const float* mean_invstd,
                         const float* mask_scale_shift, double* sums, void* gm_out, long long rows_per_group, int C,
                         int G, int act, int dtype, const unsigned char* mask_bits, cudaStream_t stream) {
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_bwd_reduce: empty dims");
  ADAMML_REQUIRE(act == ADAMML_ACT_NONE || out || mask_scale_shift || mask_bits,
                 "bn_bwd_reduce: activation mask needs the saved output, its 1-bit mask or the forward scale/shift");
  ADAMML_REQUIRE(!mask_bits || (dtype == ADAMML_BF16 && C % 8 == 0),
                 "bn_bwd_reduce: 1-bit masks come with bf16 tensors of C %% 8 == 0");
  cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)G * C * 2, stream);
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (vec_ok<T>(C, dout, out, z)) {
      const RowGeom rg = row_geom<T>(C);
      const size_t sm = sizeof(double) * rg.threads * 2 * VecIO<T>::N;
      const bool maskz = mask_scale_shift && act != ADAMML_ACT_NONE;
      const int bpg = maskz ? reduce_bpg(bn_reduce_rows_kernel<T, 1, true>, sm, rg, rows_per_group, G)
                            : reduce_bpg(bn_reduce_rows_kernel<T, 1, false>, sm, rg, rows_per_group, G);
      dim3 vg((unsigned)(bpg * (long long)G), rg.cchunks);
      if (maskz)
        bn_reduce_rows_kernel<T, 1, true><<<vg, rg.threads, sm, stream>>>((const T*)dout, nullptr, (const T*)z,
                                                                          mean_invstd, mask_scale_shift, sums,
                                                                          (T*)gm_out, rows_per_group, C, rg.cpb, rg.k,
                                                                          bpg, act);
      else
        bn_reduce_rows_kernel<T, 1, false><<<vg, rg.threads, sm, stream>>>((const T*)dout, (const T*)out, (const T*)z,
                                                                           mean_invstd, nullptr, sums, (T*)gm_out,
                                                                           rows_per_group, C, rg.cpb, rg.k, bpg, act,
                                                                           mask_bits);
    } else {
      ADAMML_REQUIRE(act == ADAMML_ACT_NONE || out, "bn_bwd_reduce: ragged channel count needs the saved output");
      ADAMML_REQUIRE(!gm_out && !mask_bits, "bn_bwd_reduce: gm_out / mask_bits need a vectorisable channel count");
      int bpg = ceil_div(rows_per_group, ROWS_PER_BLOCK);
      dim3 grid((unsigned)(bpg * (long long)G), ceil_div(C, 32));
      dim3 block(32, 8);
      bn_bwd_reduce_kernel<T><<<grid, block, 0, stream>>>((const T*)dout, (const T*)out, (const T*)z, mean_invstd,
                                                         sums, rows_per_group, C, bpg, act);
    }
  });
  return adamml_check_launch("bn_bwd_reduce");
}

int adamml_bn_bwd_apply(const void* dout, const void* out, const void* z, const float* mean_invstd,
                        const float* gamma, const float* mask_scale_shift, const double* sums, void* dz, void* dres,
                        long long rows_per_group, int C, int G, double count, int act, int training, int dtype,
                        cudaStream_t stream) {
  ADAMML_REQUIRE(act == ADAMML_ACT_NONE || out || mask_scale_shift,
                 "bn_bwd_apply: activation mask needs the saved output or the forward scale/shift");
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_bwd_apply: empty dims");
  ADAMML_REQUIRE(dz || dres, "bn_bwd_apply: nothing to write");
  long long epg = rows_per_group * C;
  long long total = epg * G;
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (vec_ok<T>(C, dout, out, z, dz, dres)) {
      const RowGeom rg = row_geom<T>(C);
      int rpb, bpg;
      stream_geom(rg, rows_per_group, &rpb, &bpg);
      dim3 vg((unsigned)(bpg * (long long)G), rg.cchunks);
      if (mask_scale_shift && act != ADAMML_ACT_NONE)
        bn_bwd_apply_rows_kernel<T, true><<<vg, rg.threads, 0, stream>>>(
            (const T*)dout, nullptr, (const T*)z, mean_invstd, gamma, mask_scale_shift, sums, (T*)dz, (T*)dres,
            rows_per_group, C, rg.cpb, rg.k, rpb, bpg, count, act, training);
      else
        bn_bwd_apply_rows_kernel<T, false><<<vg, rg.threads, 0, stream>>>(
            (const T*)dout, (const T*)out, (const T*)z, mean_invstd, gamma, nullptr, sums, (T*)dz, (T*)dres,
            rows_per_group, C, rg.cpb, rg.k, rpb, bpg, count, act, training);
    } else {
      ADAMML_REQUIRE(act == ADAMML_ACT_NONE || out, "bn_bwd_apply: ragged channel count needs the saved output");
      bn_bwd_apply_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)dout, (const T*)out, (const T*)z,
                                                                  mean_invstd, gamma, sums, (T*)dz, (T*)dres, total,
                                                                  epg, C, count, act, training);
    }
  });
  return adamml_check_launch("bn_bwd_apply");
}

/* raw [G][C][2] = (sum gm, sum gm * out) accumulated by adamml_dwconv_bwd's fused reduction -> sums [G][C][2] =
 * (sum gm, sum gm * xhat), the layout adamml_bn_bwd_reduce produces; scale_shift / mean_invstd: the layer's forward
 * [G][C][2] coefficients.  raw and sums may alias. */
int adamml_bn_sums_from_out(const double* raw, const float* scale_shift, const float* mean_invstd, double* sums, int C,
                            int G, cudaStream_t stream) {
  ADAMML_REQUIRE(raw && scale_shift && mean_invstd && sums && C > 0 && G > 0, "bn_sums_from_out: bad arguments");
  const int n = C * G;
  bn_sums_from_out_kernel<<<(n + 255) / 256, 256, 0, stream>>>(raw, scale_shift, mean_invstd, sums, n);
  return adamml_check_launch("bn_sums_from_out");
}

int adamml_bn_param_grad(const double* sums, float* dgamma, float* dbeta, int C, int G, int accumulate,
                         cudaStream_t stream) {
  bn_param_grad_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(sums, dgamma, dbeta, C, G, accumulate);
  return adamml_check_launch("bn_param_grad");
}

}  // extern "C"
