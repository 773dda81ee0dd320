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
  int bpg = ceil_div(rows_per_group, ROWS_PER_BLOCK);
  dim3 grid((unsigned)(bpg * (long long)G), ceil_div(C, 32));
  dim3 block(32, 8);
  ADAMML_DISPATCH_DTYPE(dtype, T,
    bn_stats_kernel<T><<<grid, block, 0, stream>>>((const T*)z, sums, rows_per_group, C, bpg));
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
  ADAMML_DISPATCH_DTYPE(dtype, T,
    bn_apply_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)z, scale_shift, (const T*)res, (const T*)res_z,
                                                            res_scale_shift, (T*)out, total, epg, C, act));
  return adamml_check_launch("bn_apply");
}

int adamml_bn_bwd_reduce(const void* dout, const void* out, const void* z, const float* mean_invstd, double* sums,
                         long long rows_per_group, int C, int G, int act, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_bwd_reduce: empty dims");
  ADAMML_REQUIRE(act == ADAMML_ACT_NONE || out, "bn_bwd_reduce: activation mask needs the saved output");
  cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)G * C * 2, stream);
  int bpg = ceil_div(rows_per_group, ROWS_PER_BLOCK);
  dim3 grid((unsigned)(bpg * (long long)G), ceil_div(C, 32));
  dim3 block(32, 8);
  ADAMML_DISPATCH_DTYPE(dtype, T,
    bn_bwd_reduce_kernel<T><<<grid, block, 0, stream>>>((const T*)dout, (const T*)out, (const T*)z, mean_invstd, sums,
                                                       rows_per_group, C, bpg, act));
  return adamml_check_launch("bn_bwd_reduce");
}

int adamml_bn_bwd_apply(const void* dout, const void* out, const void* z, const float* mean_invstd,
                        const float* gamma, const double* sums, void* dz, void* dres, long long rows_per_group, int C,
                        int G, double count, int act, int training, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(rows_per_group > 0 && C > 0 && G > 0, "bn_bwd_apply: empty dims");
  ADAMML_REQUIRE(dz || dres, "bn_bwd_apply: nothing to write");
  long long epg = rows_per_group * C;
  long long total = epg * G;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    bn_bwd_apply_kernel<T><<<ew_blocks(total), 256, 0, stream>>>((const T*)dout, (const T*)out, (const T*)z,
                                                                mean_invstd, gamma, sums, (T*)dz, (T*)dres, total,
                                                                epg, C, count, act, training));
  return adamml_check_launch("bn_bwd_apply");
}

int adamml_bn_param_grad(const double* sums, float* dgamma, float* dbeta, int C, int G, int accumulate,
                         cudaStream_t stream) {
  bn_param_grad_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(sums, dgamma, dbeta, C, G, accumulate);
  return adamml_check_launch("bn_param_grad");
}

}  // extern "C"
