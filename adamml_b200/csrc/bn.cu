// Batch-norm statistics / apply / backward for NHWC activations with per-segment groups.
//
// Reference semantics: every nn.BatchNorm2d on the AdaMML path (models/resnet.py:92-111,
// models/sound_mobilenet_v2.py:33-40, models/policy_net.py:38-52,63-86) is called once per
// segment (models/adamml.py:84-86, models/policy_net.py:323-326), i.e. batch statistics
// are taken over ONE segment's images.  Here all S segments are batched into one launch:
// images are ordered segment-major, group g = img / imgs_per_group, statistics are kept per
// (group, channel) and running stats receive the S momentum updates in segment order.
#include "common.cuh"

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
// 16-byte vectorised versions (C % VEC == 0): thread = (channel vector, row lane)
constexpr int VROWS_PER_THREAD = 32;
template <typename T> struct AccT { typedef float type; };
template <> struct AccT<float> { typedef double type; };

// MODE 0: sum z, sum z^2.   MODE 1: sum gm, sum gm*xhat (gm = dout * mask(out)).
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
bn_reduce_vec_kernel(const T* __restrict__ a /*z | dout*/, const T* __restrict__ out, const T* __restrict__ z,
                     const float* __restrict__ mean_invstd, double* __restrict__ sums, long long rows_per_group,
                     int C, int blocks_per_group, int rows_per_block, int act) {
  constexpr int V = VecIO<T>::N;
  extern __shared__ double shd[];
  const int TX = blockDim.x, TY = blockDim.y;
  const int g = blockIdx.x / blocks_per_group;
  const int bg = blockIdx.x % blocks_per_group;
  const int cv = blockIdx.y * TX + threadIdx.x;
  const int c0 = cv * V;
  const bool ok = c0 < C;
  long long r0 = (long long)bg * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows_per_group) r1 = rows_per_group;
  // fp32 (parity) mode accumulates in double per thread; bf16 mode in float over <= 32 rows per thread
  typedef typename AccT<T>::type acc_t;
  acc_t s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s[i] = 0; q[i] = 0; }
  if (ok) {
    float mean[V], invstd[V];
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        mean[i] = mean_invstd[((long long)g * C + c0 + i) * 2 + 0];
        invstd[i] = mean_invstd[((long long)g * C + c0 + i) * 2 + 1];
      }
    }
    const long long base = (long long)g * rows_per_group * C + c0;
    for (long long r = r0 + threadIdx.y; r < r1; r += TY) {
      const long long off = base + r * C;
      float va[V];
      VecIO<T>::load(a + off, va);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) { s[i] += (acc_t)va[i]; q[i] += (acc_t)va[i] * (acc_t)va[i]; }
      } else {
        float vz[V];
        VecIO<T>::load(z + off, vz);
        if (act != ADAMML_ACT_NONE) {
          float vo[V];
          VecIO<T>::load(out + off, vo);
#pragma unroll
          for (int i = 0; i < V; ++i) if (!act_pass(vo[i], act)) va[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          s[i] += (acc_t)va[i];
          q[i] += (acc_t)va[i] * (acc_t)((vz[i] - mean[i]) * invstd[i]);
        }
      }
    }
  }
  // block reduce over TY in double
  double* sh = shd + ((size_t)threadIdx.y * TX + threadIdx.x) * (2 * V);
#pragma unroll
  for (int i = 0; i < V; ++i) { sh[i] = (double)s[i]; sh[V + i] = (double)q[i]; }
  __syncthreads();
  if (threadIdx.y == 0 && ok) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      double ds = 0.0, dq = 0.0;
      for (int y = 0; y < TY; ++y) {
        const double* o = shd + ((size_t)y * TX + threadIdx.x) * (2 * V);
        ds += o[i];
        dq += o[V + i];
      }
      atomicAdd(&sums[((long long)g * C + c0 + i) * 2 + 0], ds);
      atomicAdd(&sums[((long long)g * C + c0 + i) * 2 + 1], dq);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_vec_kernel(const T* __restrict__ z, const float* __restrict__ ss, const T* __restrict__ res,
                    const T* __restrict__ res_z, const float* __restrict__ res_ss, T* __restrict__ out,
                    long long total_vec, long long elems_per_group, int C, int act) {
  constexpr int V = VecIO<T>::N;
  for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < total_vec;
       iv += (long long)gridDim.x * blockDim.x) {
    const long long idx = iv * V;
    const int c0 = (int)(idx % C);
    const int g = (int)(idx / elems_per_group);
    const float* s = ss + ((long long)g * C + c0) * 2;
    float vz[V], vo[V];
    VecIO<T>::load(z + idx, vz);
#pragma unroll
    for (int i = 0; i < V; ++i) vo[i] = fmaf(vz[i], s[2 * i], s[2 * i + 1]);
    if (res) {
      float vr[V];
      VecIO<T>::load(res + idx, vr);
#pragma unroll
      for (int i = 0; i < V; ++i) vo[i] += vr[i];
    }
    if (res_z) {
      const float* rs = res_ss + ((long long)g * C + c0) * 2;
      float vr[V];
      VecIO<T>::load(res_z + idx, vr);
#pragma unroll
      for (int i = 0; i < V; ++i) vo[i] += fmaf(vr[i], rs[2 * i], rs[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) vo[i] = act_apply(vo[i], act);
    VecIO<T>::store(out + idx, vo);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_vec_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ z,
                        const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
                        const double* __restrict__ sums, T* __restrict__ dz, T* __restrict__ dres,
                        long long total_vec, long long elems_per_group, int C, double count, int act, int training) {
  constexpr int V = VecIO<T>::N;
  const float inv_count = (float)(1.0 / count);
  for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < total_vec;
       iv += (long long)gridDim.x * blockDim.x) {
    const long long idx = iv * V;
    const int c0 = (int)(idx % C);
    const int g = (int)(idx / elems_per_group);
    const long long gc = (long long)g * C + c0;
    float gm[V];
    VecIO<T>::load(dout + idx, gm);
    if (act != ADAMML_ACT_NONE) {
      float vo[V];
      VecIO<T>::load(out + idx, vo);
#pragma unroll
      for (int i = 0; i < V; ++i) if (!act_pass(vo[i], act)) gm[i] = 0.f;
    }
    if (dres) VecIO<T>::store(dres + idx, gm);
    if (dz) {
      float v[V];
      if (training) {
        float vz[V];
        VecIO<T>::load(z + idx, vz);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float mean = mean_invstd[(gc + i) * 2 + 0];
          const float invstd = mean_invstd[(gc + i) * 2 + 1];
          const float ga = gamma ? gamma[c0 + i] : 1.f;
          const float m1 = (float)sums[(gc + i) * 2 + 0] * inv_count;
          const float m2 = (float)sums[(gc + i) * 2 + 1] * inv_count;
          const float xhat = (vz[i] - mean) * invstd;
          v[i] = ga * invstd * (gm[i] - m1 - xhat * m2);
        }
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float invstd = mean_invstd[(gc + i) * 2 + 1];
          const float ga = gamma ? gamma[c0 + i] : 1.f;
          v[i] = ga * invstd * gm[i];
        }
      }
      VecIO<T>::store(dz + idx, v);
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

// launch geometry of the vectorised reductions
template <typename T>
inline void reduce_geom(long long rows_per_group, int C, dim3* block, int* cgrid, int* rows_per_block, int* bpg,
                        size_t* smem) {
  const int V = VecIO<T>::N;
  int cvecs = C / V;
  int tx = 1;
  while (tx < cvecs && tx < 32) tx *= 2;
  int ty = 256 / tx;
  *block = dim3(tx, ty);
  *cgrid = (cvecs + tx - 1) / tx;
  *rows_per_block = ty * VROWS_PER_THREAD;
  *bpg = (int)((rows_per_group + *rows_per_block - 1) / *rows_per_block);
  *smem = sizeof(double) * 256 * 2 * V;
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
      dim3 vb; int cg, rpb, vbpg; size_t sm;
      reduce_geom<T>(rows_per_group, C, &vb, &cg, &rpb, &vbpg, &sm);
      static bool cfgd = false;
      if (!cfgd) { cudaFuncSetAttribute(bn_reduce_vec_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); cfgd = true; }
      dim3 vg((unsigned)(vbpg * (long long)G), cg);
      bn_reduce_vec_kernel<T, 0><<<vg, vb, sm, stream>>>((const T*)z, nullptr, nullptr, nullptr, sums, rows_per_group,
                                                         C, vbpg, rpb, 0);
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
      long long tv = total / VecIO<T>::N;
      bn_apply_vec_kernel<T><<<ew_blocks(tv), 256, 0, stream>>>((const T*)z, scale_shift, (const T*)res,
                                                                (const T*)res_z, res_scale_shift, (T*)out, tv, epg, C,
                                                                act);
    } else {
      bn_apply_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)z, scale_shift, (const T*)res,
                                                              (const T*)res_z, res_scale_shift, (T*)out, total, epg, C,
                                                              act);
    }
  });
  return adamml_check_launch("bn_apply");
}

int adamml_bn_bwd_reduce(const void* dout, const void* out, const void* z, const float* mean_invstd, double* sums,
                         long long rows_per_group, int C, int G, int act, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_bwd_reduce: empty dims");
  ADAMML_REQUIRE(act == ADAMML_ACT_NONE || out, "bn_bwd_reduce: activation mask needs the saved output");
  cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)G * C * 2, stream);
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (vec_ok<T>(C, dout, out, z)) {
      dim3 vb; int cg, rpb, vbpg; size_t sm;
      reduce_geom<T>(rows_per_group, C, &vb, &cg, &rpb, &vbpg, &sm);
      static bool cfgd = false;
      if (!cfgd) { cudaFuncSetAttribute(bn_reduce_vec_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); cfgd = true; }
      dim3 vg((unsigned)(vbpg * (long long)G), cg);
      bn_reduce_vec_kernel<T, 1><<<vg, vb, sm, stream>>>((const T*)dout, (const T*)out, (const T*)z, mean_invstd, sums,
                                                         rows_per_group, C, vbpg, rpb, act);
    } else {
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
                        const float* gamma, const double* sums, void* dz, void* dres, long long rows_per_group, int C,
                        int G, double count, int act, int training, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_bwd_apply: empty dims");
  ADAMML_REQUIRE(dz || dres, "bn_bwd_apply: nothing to write");
  long long epg = rows_per_group * C;
  long long total = epg * G;
  ADAMML_DISPATCH_DTYPE(dtype, T, {
    if (vec_ok<T>(C, dout, out, z, dz, dres)) {
      long long tv = total / VecIO<T>::N;
      bn_bwd_apply_vec_kernel<T><<<ew_blocks(tv), 256, 0, stream>>>((const T*)dout, (const T*)out, (const T*)z,
                                                                    mean_invstd, gamma, sums, (T*)dz, (T*)dres, tv, epg,
                                                                    C, count, act, training);
    } else {
      bn_bwd_apply_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)dout, (const T*)out, (const T*)z,
                                                                  mean_invstd, gamma, sums, (T*)dz, (T*)dres, total,
                                                                  epg, C, count, act, training);
    }
  });
  return adamml_check_launch("bn_bwd_apply");
}

int adamml_bn_param_grad(const double* sums, float* dgamma, float* dbeta, int C, int G, int accumulate,
                         cudaStream_t stream) {
  bn_param_grad_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(sums, dgamma, dbeta, C, G, accumulate);
  return adamml_check_launch("bn_param_grad");
}

}  // extern "C"
