// Policy-head and fusion kernels (all fp32; they are <0.1% of the FLOPs but decide the
// bit-exactness of the modality selections).
//  - LSTMCell step + per-modality Linear(256,2) + hard Gumbel-softmax as ONE warp-reduction
//    kernel per segment step (reference models/policy_net.py:283-290,345-365)
//  - gate x logits, learnable late-fusion weights, sum over modalities, mean over segments
//    (reference models/joint_resnet_mobilenetv2.py:92-97,112-127; models/adamml.py:88)
//  - bias/activation helpers for the Linear layers (policy_net.py:228-231, resnet.py:217)
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------
// One CTA per sample.  dynamic smem: hprev[Hd] | u[2M] | gates[4Hd] | h[Hd]
// gx      [N,4Hd]   = W_ih[:, :Fdim] . feat_s   (hoisted GEMM, no bias)
// xin_tail          = &x_in[s][0][Fdim], row stride xin_ld: receives u (prev logits) for wgrad
__global__ void policy_step_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ prev_logits,
                                       const float* __restrict__ h_prev, const float* __restrict__ c_prev,
                                       const float* __restrict__ w_ih, long long w_ih_ld, int Fdim,
                                       const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                       const float* __restrict__ b_hh, const float* __restrict__ fc_w,
                                       const float* __restrict__ fc_b, const float* __restrict__ expo, float tau,
                                       float* __restrict__ gates_out, float* __restrict__ h_out,
                                       float* __restrict__ c_out, float* __restrict__ logits_out,
                                       float* __restrict__ ysoft_out, float* __restrict__ dec_out,
                                       float* __restrict__ xin_tail, long long xin_ld, int N, int M, int Hd) {
  extern __shared__ float sm[];
  float* s_hprev = sm;
  float* s_u = s_hprev + Hd;
  float* s_gates = s_u + 2 * M;
  float* s_h = s_gates + 4 * Hd;
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

  for (int k = tid; k < Hd; k += blockDim.x) s_hprev[k] = h_prev ? h_prev[(long long)n * Hd + k] : 0.f;
  for (int j = tid; j < 2 * M; j += blockDim.x) {
    int m = j >> 1, jj = j & 1;
    float v = prev_logits ? prev_logits[((long long)m * N + n) * 2 + jj] : 0.f;
    s_u[j] = v;
    if (xin_tail) xin_tail[(long long)n * xin_ld + j] = v;
  }
  __syncthreads();

  for (int r = warp; r < 4 * Hd; r += nwarps) {
    float acc = 0.f;
    if (h_prev) {
      const float* wr = w_hh + (long long)r * Hd;
      for (int k = lane; k < Hd; k += 32) acc = fmaf(wr[k], s_hprev[k], acc);
    }
    if (prev_logits && lane < 2 * M) acc = fmaf(w_ih[(long long)r * w_ih_ld + Fdim + lane], s_u[lane], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_gates[r] = acc + gx[(long long)n * 4 * Hd + r] + b_ih[r] + b_hh[r];
  }
  __syncthreads();

  for (int k = tid; k < Hd; k += blockDim.x) {
    float ig = sigmoidf_acc(s_gates[k]);
    float fg = sigmoidf_acc(s_gates[Hd + k]);
    float gg = tanhf(s_gates[2 * Hd + k]);
    float og = sigmoidf_acc(s_gates[3 * Hd + k]);
    float cp = c_prev ? c_prev[(long long)n * Hd + k] : 0.f;
    float c = fg * cp + ig * gg;
    float h = og * tanhf(c);
    long long gb = (long long)n * 4 * Hd;
    gates_out[gb + k] = ig;
    gates_out[gb + Hd + k] = fg;
    gates_out[gb + 2 * Hd + k] = gg;
    gates_out[gb + 3 * Hd + k] = og;
    c_out[(long long)n * Hd + k] = c;
    h_out[(long long)n * Hd + k] = h;
    s_h[k] = h;
  }
  __syncthreads();

  // logits: row j of modality m = fc_w[m][j][:] . h + fc_b[m][j]   (stored in s_gates[0..2M))
  for (int row = warp; row < 2 * M; row += nwarps) {
    const float* wr = fc_w + (long long)row * Hd;
    float acc = 0.f;
    for (int k = lane; k < Hd; k += 32) acc = fmaf(wr[k], s_h[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_gates[row] = acc + fc_b[row];
  }
  __syncthreads();

  if (tid < M) {
    int m = tid;
    float l0 = s_gates[2 * m], l1 = s_gates[2 * m + 1];
    long long o = ((long long)m * N + n) * 2;
    logits_out[o] = l0;
    logits_out[o + 1] = l1;
    // F.gumbel_softmax: g = -log(Exp(1)); y = softmax((l+g)/tau); hard one-hot, straight-through
    float g0 = -logf(expo[o]), g1 = -logf(expo[o + 1]);
    float v0 = (l0 + g0) / tau, v1 = (l1 + g1) / tau;
    float mx = fmaxf(v0, v1);
    float e0 = expf(v0 - mx), e1 = expf(v1 - mx);
    float ssum = e0 + e1;
    float y0 = e0 / ssum, y1 = e1 / ssum;
    ysoft_out[o] = y0;
    ysoft_out[o + 1] = y1;
    float hard1 = (y1 > y0) ? 1.f : 0.f;  // torch.max keeps the first index on ties
    dec_out[(long long)m * N + n] = (hard1 - y1) + y1;
  }
}

// dynamic smem: dl[2M] | dh[Hd] | dgates[4Hd]
__global__ void policy_step_bwd_kernel(const float* __restrict__ d_dec, const float* __restrict__ d_logits_fb,
                                       const float* __restrict__ dh_next, const float* __restrict__ dc_next,
                                       const float* __restrict__ gates, const float* __restrict__ c_cur,
                                       const float* __restrict__ c_prev, const float* __restrict__ ysoft,
                                       const float* __restrict__ w_ih, long long w_ih_ld, int Fdim,
                                       const float* __restrict__ w_hh, const float* __restrict__ fc_w, float tau,
                                       float* __restrict__ dl_out /*[M][rows..] slice for this step: [m*dl_ms + n*2 + j]*/,
                                       long long dl_ms, float* __restrict__ dgates_out, float* __restrict__ dh_prev,
                                       float* __restrict__ dc_prev, float* __restrict__ d_prev_logits, int N, int M,
                                       int Hd) {
  extern __shared__ float sm[];
  float* s_dl = sm;
  float* s_dh = s_dl + 2 * M;
  float* s_dg = s_dh + Hd;
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

  if (tid < M) {
    int m = tid;
    long long o = ((long long)m * N + n) * 2;
    float dd = d_dec ? d_dec[(long long)m * N + n] : 0.f;
    float y0 = ysoft[o], y1 = ysoft[o + 1];
    // ret = y_hard - y_soft.detach() + y_soft  => d ret / d y_soft = I ; only column 1 is used
    float dot = dd * y1;
    float dv0 = y0 * (0.f - dot);
    float dv1 = y1 * (dd - dot);
    float dl0 = dv0 / tau, dl1 = dv1 / tau;
    if (d_logits_fb) { dl0 += d_logits_fb[o]; dl1 += d_logits_fb[o + 1]; }
    s_dl[2 * m] = dl0;
    s_dl[2 * m + 1] = dl1;
    dl_out[(long long)m * dl_ms + (long long)n * 2] = dl0;
    dl_out[(long long)m * dl_ms + (long long)n * 2 + 1] = dl1;
  }
  __syncthreads();
  for (int k = tid; k < Hd; k += blockDim.x) {
    float acc = dh_next ? dh_next[(long long)n * Hd + k] : 0.f;
    for (int row = 0; row < 2 * M; ++row) acc = fmaf(s_dl[row], fc_w[(long long)row * Hd + k], acc);
    s_dh[k] = acc;
  }
  __syncthreads();
  for (int k = tid; k < Hd; k += blockDim.x) {
    long long gb = (long long)n * 4 * Hd;
    float ig = gates[gb + k], fg = gates[gb + Hd + k], gg = gates[gb + 2 * Hd + k], og = gates[gb + 3 * Hd + k];
    float c = c_cur[(long long)n * Hd + k];
    float cp = c_prev ? c_prev[(long long)n * Hd + k] : 0.f;
    float tc = tanhf(c);
    float dh = s_dh[k];
    float d_o = dh * tc;
    float dc = (dc_next ? dc_next[(long long)n * Hd + k] : 0.f) + dh * og * (1.f - tc * tc);
    float di = dc * gg, df = dc * cp, dg = dc * ig;
    if (dc_prev) dc_prev[(long long)n * Hd + k] = dc * fg;
    float a = di * ig * (1.f - ig);
    float b = df * fg * (1.f - fg);
    float cgr = dg * (1.f - gg * gg);
    float d = d_o * og * (1.f - og);
    s_dg[k] = a; s_dg[Hd + k] = b; s_dg[2 * Hd + k] = cgr; s_dg[3 * Hd + k] = d;
    dgates_out[gb + k] = a;
    dgates_out[gb + Hd + k] = b;
    dgates_out[gb + 2 * Hd + k] = cgr;
    dgates_out[gb + 3 * Hd + k] = d;
  }
  __syncthreads();
  if (dh_prev) {
    for (int k = tid; k < Hd; k += blockDim.x) {
      float acc = 0.f;
      for (int r = 0; r < 4 * Hd; ++r) acc = fmaf(s_dg[r], w_hh[(long long)r * Hd + k], acc);
      dh_prev[(long long)n * Hd + k] = acc;
    }
  }
  if (d_prev_logits) {
    for (int j = warp; j < 2 * M; j += nwarps) {
      float acc = 0.f;
      for (int r = lane; r < 4 * Hd; r += 32) acc = fmaf(s_dg[r], w_ih[(long long)r * w_ih_ld + Fdim + j], acc);
      acc = warp_sum(acc);
      if (lane == 0) d_prev_logits[((long long)(j >> 1) * N + n) * 2 + (j & 1)] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float fuse_weight(const float* lf, int m, int M) {
  if (!lf) return 1.f / (float)M;
  if (m < M - 1) return lf[m];
  float s = 0.f;
  for (int i = 0; i < M - 1; ++i) s += lf[i];
  return 1.f - s;
}

// logits [M][S][N][C], dec [S][M][N] (may be null = all ones), out [N][C]
__global__ void fuse_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ dec,
                                const float* __restrict__ lf, float* __restrict__ out, int M, int S, int N, int C) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * C) return;
  int c = idx % C, n = idx / C;
  float acc = 0.f;
  for (int s = 0; s < S; ++s) {
    float seg = 0.f;
    for (int m = 0; m < M; ++m) {
      float d = dec ? dec[((long long)s * M + m) * N + n] : 1.f;
      seg += (logits[(((long long)m * S + s) * N + n) * C + c] * d) * fuse_weight(lf, m, M);
    }
    acc += seg;
  }
  out[idx] = acc / (float)S;
}

// one warp per (s,m,n)
__global__ void fuse_bwd_kernel(const float* __restrict__ g, const float* __restrict__ logits,
                                const float* __restrict__ dec, const float* __restrict__ lf,
                                float* __restrict__ dlogits, float* __restrict__ ddec, float* __restrict__ dlf, int M,
                                int S, int N, int C) {
  int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (wid >= S * M * N) return;
  int n = wid % N;
  int m = (wid / N) % M;
  int s = wid / (N * M);
  float w = fuse_weight(lf, m, M);
  float d = dec ? dec[((long long)s * M + m) * N + n] : 1.f;
  float invS = 1.f / (float)S;
  float t = 0.f;
  for (int c = lane; c < C; c += 32) {
    float gv = g[(long long)n * C + c];
    long long li = (((long long)m * S + s) * N + n) * C + c;
    t = fmaf(gv, logits[li], t);
    if (dlogits) dlogits[li] = gv * w * d * invS;
  }
  t = warp_sum(t);
  if (lane == 0) {
    if (ddec) ddec[((long long)s * M + m) * N + n] = w * t * invS;
    if (dlf && lf) {
      float contrib = d * t * invS;  // d out / d w_m
      if (m < M - 1) atomicAdd(&dlf[m], contrib);
      else for (int i = 0; i < M - 1; ++i) atomicAdd(&dlf[i], -contrib);
    }
  }
}

// ---------------------------------------------------------------------------------------
__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, long long rows, int cols,
                                long long ld, int act) {
  long long total = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % cols);
    long long r = idx / cols;
    float v = y[r * ld + c] + (bias ? bias[c] : 0.f);
    y[r * ld + c] = act_apply(v, act);
  }
}

__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz,
                               long long rows, int cols, long long ld_dy, long long ld_y, long long ld_dz, int act) {
  long long total = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % cols);
    long long r = idx / cols;
    float gv = dy[r * ld_dy + c];
    if (!act_pass(y[r * ld_y + c], act)) gv = 0.f;
    dz[r * ld_dz + c] = gv;
  }
}

// out[c] (+)= sum_r x[r][c]   block (32 cols, 8 row lanes), grid.x over col tiles
__global__ void colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int cols,
                              long long ld, int accumulate) {
  __shared__ float sh[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < cols)
    for (long long r = threadIdx.y; r < rows; r += 8) s += x[r * ld + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    for (int i = 1; i < 8; ++i) s += sh[i][threadIdx.x];
    out[c] = (accumulate ? out[c] : 0.f) + s;
  }
}

__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                           long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x)
    out[idx] = a[idx] * b[idx];
}

// out[v][c] = mean_t x[(v*T+t)][c]
__global__ void frame_mean_kernel(const float* __restrict__ x, float* __restrict__ out, long long V, int Tn, int C,
                                  long long out_ld) {
  long long total = V * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    long long v = idx / C;
    float s = 0.f;
    for (int t = 0; t < Tn; ++t) s += x[(v * Tn + t) * C + c];
    out[v * out_ld + c] = s / (float)Tn;
  }
}

// dx[(v*T+t)][c] = dy[v][c] / T
__global__ void frame_mean_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long V, int Tn, int C,
                                      long long dy_ld) {
  long long total = V * Tn * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    long long v = idx / ((long long)C * Tn);
    dx[idx] = dy[v * dy_ld + c] / (float)Tn;
  }
}

inline int ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  long long cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}


// Stand-alone hard Gumbel-softmax over [R, 2] logits (causality_modeling=None branch, policy_net.py:330-339):
// g = -log(Exp(1)); y = softmax((l+g)/tau); dec = (onehot(argmax) - y + y)[:, 1]
__global__ void gumbel_hard_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ expo, float tau,
                                       float* __restrict__ ysoft, float* __restrict__ dec, long long R) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float l0 = logits[2 * r], l1 = logits[2 * r + 1];
  const float g0 = -logf(expo[2 * r]), g1 = -logf(expo[2 * r + 1]);
  const float v0 = (l0 + g0) / tau, v1 = (l1 + g1) / tau;
  const float mx = fmaxf(v0, v1);
  const float e0 = expf(v0 - mx), e1 = expf(v1 - mx);
  const float ssum = e0 + e1;
  const float y0 = e0 / ssum, y1 = e1 / ssum;
  ysoft[2 * r] = y0;
  ysoft[2 * r + 1] = y1;
  const float hard1 = (y1 > y0) ? 1.f : 0.f;  // torch.max keeps the first index on ties
  dec[r] = (hard1 - y1) + y1;
}
// straight-through gradient: d dec / d l1 = y0*y1/tau, d dec / d l0 = -y0*y1/tau
__global__ void gumbel_hard_bwd_kernel(const float* __restrict__ d_dec, const float* __restrict__ ysoft, float tau,
                                       float* __restrict__ dlogits, long long R) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float t = d_dec[r] * ysoft[2 * r] * ysoft[2 * r + 1] / tau;
  dlogits[2 * r] = -t;
  dlogits[2 * r + 1] = t;
}

}  // namespace

extern "C" {

int adamml_policy_step_fwd(const float* gx, const float* prev_logits, const float* h_prev, const float* c_prev,
                           const float* w_ih, long long w_ih_ld, int Fdim, const float* w_hh, const float* b_ih,
                           const float* b_hh, const float* fc_w, const float* fc_b, const float* expo, float tau,
                           float* gates_out, float* h_out, float* c_out, float* logits_out, float* ysoft_out,
                           float* dec_out, float* xin_tail, long long xin_ld, int N, int M, int Hd,
                           cudaStream_t stream) {
  ADAMML_REQUIRE(N > 0 && M > 0 && M <= 16 && Hd > 0 && Hd <= 2048, "policy_step_fwd: bad dims");
  ADAMML_REQUIRE(tau > 0.f, "policy_step_fwd: temperature must be positive");
  ADAMML_REQUIRE((h_prev == nullptr) == (c_prev == nullptr), "policy_step_fwd: h_prev/c_prev must come together");
  size_t smem = sizeof(float) * (size_t)(Hd + 2 * M + 4 * Hd + Hd);
  policy_step_fwd_kernel<<<N, 256, smem, stream>>>(gx, prev_logits, h_prev, c_prev, w_ih, w_ih_ld, Fdim, w_hh, b_ih,
                                                  b_hh, fc_w, fc_b, expo, tau, gates_out, h_out, c_out, logits_out,
                                                  ysoft_out, dec_out, xin_tail, xin_ld, N, M, Hd);
  return adamml_check_launch("policy_step_fwd");
}

int adamml_policy_step_bwd(const float* d_dec, const float* d_logits_fb, const float* dh_next, const float* dc_next,
                           const float* gates, const float* c_cur, const float* c_prev, const float* ysoft,
                           const float* w_ih, long long w_ih_ld, int Fdim, const float* w_hh, const float* fc_w,
                           float tau, float* dl_out, long long dl_ms, float* dgates_out, float* dh_prev,
                           float* dc_prev, float* d_prev_logits, int N, int M, int Hd, cudaStream_t stream) {
  ADAMML_REQUIRE(N > 0 && M > 0 && M <= 16 && Hd > 0 && Hd <= 2048, "policy_step_bwd: bad dims");
  size_t smem = sizeof(float) * (size_t)(2 * M + Hd + 4 * Hd);
  policy_step_bwd_kernel<<<N, 256, smem, stream>>>(d_dec, d_logits_fb, dh_next, dc_next, gates, c_cur, c_prev, ysoft,
                                                  w_ih, w_ih_ld, Fdim, w_hh, fc_w, tau, dl_out, dl_ms, dgates_out,
                                                  dh_prev, dc_prev, d_prev_logits, N, M, Hd);
  return adamml_check_launch("policy_step_bwd");
}

int adamml_fuse_fwd(const float* logits, const float* dec, const float* lf, float* out, int M, int S, int N, int C,
                    cudaStream_t stream) {
  ADAMML_REQUIRE(M > 0 && S > 0 && N > 0 && C > 0, "fuse_fwd: bad dims");
  fuse_fwd_kernel<<<ceil_div((long long)N * C, 128), 128, 0, stream>>>(logits, dec, lf, out, M, S, N, C);
  return adamml_check_launch("fuse_fwd");
}

int adamml_fuse_bwd(const float* g, const float* logits, const float* dec, const float* lf, float* dlogits,
                    float* ddec, float* dlf, int M, int S, int N, int C, cudaStream_t stream) {
  ADAMML_REQUIRE(M > 0 && S > 0 && N > 0 && C > 0, "fuse_bwd: bad dims");
  if (dlf && M > 1) cudaMemsetAsync(dlf, 0, sizeof(float) * (size_t)(M - 1), stream);
  long long warps = (long long)S * M * N;
  fuse_bwd_kernel<<<ceil_div(warps * 32, 128), 128, 0, stream>>>(g, logits, dec, lf, dlogits, ddec, dlf, M, S, N, C);
  return adamml_check_launch("fuse_bwd");
}

int adamml_bias_act(float* y, const float* bias, long long rows, int cols, long long ld, int act,
                    cudaStream_t stream) {
  ADAMML_REQUIRE(rows > 0 && cols > 0, "bias_act: bad dims");
  if (ld <= 0) ld = cols;
  bias_act_kernel<<<ew_blocks(rows * cols), 256, 0, stream>>>(y, bias, rows, cols, ld, act);
  return adamml_check_launch("bias_act");
}

int adamml_act_bwd(const float* dy, const float* y, float* dz, long long rows, int cols, long long ld_dy,
                   long long ld_y, long long ld_dz, int act, cudaStream_t stream) {
  ADAMML_REQUIRE(rows > 0 && cols > 0, "act_bwd: bad dims");
  if (ld_dy <= 0) ld_dy = cols;
  if (ld_y <= 0) ld_y = cols;
  if (ld_dz <= 0) ld_dz = cols;
  act_bwd_kernel<<<ew_blocks(rows * cols), 256, 0, stream>>>(dy, y, dz, rows, cols, ld_dy, ld_y, ld_dz, act);
  return adamml_check_launch("act_bwd");
}

int adamml_colsum(const float* x, float* out, long long rows, int cols, long long ld, int accumulate,
                  cudaStream_t stream) {
  ADAMML_REQUIRE(rows > 0 && cols > 0, "colsum: bad dims");
  if (ld <= 0) ld = cols;
  colsum_kernel<<<ceil_div(cols, 32), dim3(32, 8), 0, stream>>>(x, out, rows, cols, ld, accumulate);
  return adamml_check_launch("colsum");
}

int adamml_mul(const float* a, const float* b, float* out, long long total, cudaStream_t stream) {
  ADAMML_REQUIRE(total > 0, "mul: bad size");
  mul_kernel<<<ew_blocks(total), 256, 0, stream>>>(a, b, out, total);
  return adamml_check_launch("mul");
}

int adamml_frame_mean(const float* x, float* out, long long V, int Tn, int C, long long out_ld, cudaStream_t stream) {
  ADAMML_REQUIRE(V > 0 && Tn > 0 && C > 0, "frame_mean: bad dims");
  if (out_ld <= 0) out_ld = C;
  frame_mean_kernel<<<ew_blocks(V * C), 256, 0, stream>>>(x, out, V, Tn, C, out_ld);
  return adamml_check_launch("frame_mean");
}

int adamml_frame_mean_bwd(const float* dy, float* dx, long long V, int Tn, int C, long long dy_ld,
                          cudaStream_t stream) {
  ADAMML_REQUIRE(V > 0 && Tn > 0 && C > 0, "frame_mean_bwd: bad dims");
  if (dy_ld <= 0) dy_ld = C;
  frame_mean_bwd_kernel<<<ew_blocks(V * Tn * C), 256, 0, stream>>>(dy, dx, V, Tn, C, dy_ld);
  return adamml_check_launch("frame_mean_bwd");
}


int adamml_gumbel_hard_fwd(const float* logits, const float* expo, float tau, float* ysoft, float* dec, long long R,
                           cudaStream_t stream) {
  ADAMML_REQUIRE(R > 0 && tau > 0.f, "gumbel_hard_fwd: bad arguments");
  gumbel_hard_fwd_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(logits, expo, tau, ysoft, dec, R);
  return adamml_check_launch("gumbel_hard_fwd");
}

int adamml_gumbel_hard_bwd(const float* d_dec, const float* ysoft, float tau, float* dlogits, long long R,
                           cudaStream_t stream) {
  ADAMML_REQUIRE(R > 0 && tau > 0.f, "gumbel_hard_bwd: bad arguments");
  gumbel_hard_bwd_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(d_dec, ysoft, tau, dlogits, R);
  return adamml_check_launch("gumbel_hard_bwd");
}

}  // extern "C"
