// Train-step tail of train_adamml() (utils/utils.py:362-400, train_adamml.py:250-257), §8 f2:
//   * cross-entropy + 'blockdrop' policy loss (utils/utils.py:166-184) forward AND gradients in one launch,
//   * multi-tensor SGD(momentum, weight decay) and Adam(weight decay) over ALL parameter tensors of an optimizer in one
//     launch each (the reference's torch.optim.SGD / Adam loop over ~650 tensors each),
//   * clip_grad_norm_ (utils/utils.py:390-391) over all gradients as two launches without a host round trip.
// Semantics follow torch.optim exactly (dampening 0, no Nesterov, L2 weight decay added to the gradient, Adam with bias
// correction and eps outside the square root, no amsgrad); the step counter lives on the device so that the launches
// capture into a CUDA graph.
#include "common.cuh"

namespace {

constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = 8192;  // elements per block

// table: [4][n] device pointers (param, grad, state1, state2) as unsigned long long; sizes [n]; chunk_tensor /
// chunk_start [chunks]: which tensor and which element offset a block works on
__global__ void __launch_bounds__(OPT_THREADS)
sgd_multi_kernel(const unsigned long long* __restrict__ table, const long long* __restrict__ sizes,
                 const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_start, int n, float lr,
                 float momentum, float weight_decay) {
  const int t = chunk_tensor[blockIdx.x];
  float* __restrict__ p = reinterpret_cast<float*>(table[t]);
  const float* __restrict__ g = reinterpret_cast<const float*>(table[n + t]);
  float* __restrict__ buf = reinterpret_cast<float*>(table[2 * n + t]);
  const long long e0 = chunk_start[blockIdx.x];
  long long e1 = e0 + OPT_CHUNK;
  if (e1 > sizes[t]) e1 = sizes[t];
  for (long long i = e0 + threadIdx.x; i < e1; i += OPT_THREADS) {
    const float w = p[i];
    float d = g[i];
    if (weight_decay != 0.f) d = fmaf(weight_decay, w, d);
    if (momentum != 0.f) {
      d = fmaf(momentum, buf[i], d);  // zero-initialised buffer == torch's "first step: buf = grad"
      buf[i] = d;
    }
    p[i] = fmaf(-lr, d, w);
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
adam_multi_kernel(const unsigned long long* __restrict__ table, const long long* __restrict__ sizes,
                  const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_start, int n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, const long long* __restrict__ step) {
  const int t = chunk_tensor[blockIdx.x];
  float* __restrict__ p = reinterpret_cast<float*>(table[t]);
  const float* __restrict__ g = reinterpret_cast<const float*>(table[n + t]);
  float* __restrict__ m = reinterpret_cast<float*>(table[2 * n + t]);
  float* __restrict__ v = reinterpret_cast<float*>(table[3 * n + t]);
  const double k = (double)(*step + 1);  // 1-based step of THIS update (advanced afterwards by adam_step_kernel)
  const float bias1 = (float)(1.0 - pow((double)beta1, k));
  const float rsq_bias2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, k)));
  const float step_size = lr / bias1;
  const long long e0 = chunk_start[blockIdx.x];
  long long e1 = e0 + OPT_CHUNK;
  if (e1 > sizes[t]) e1 = sizes[t];
  for (long long i = e0 + threadIdx.x; i < e1; i += OPT_THREADS) {
    const float w = p[i];
    float d = g[i];
    if (weight_decay != 0.f) d = fmaf(weight_decay, w, d);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * d);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * d * d);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * rsq_bias2 + eps;
    p[i] = w - step_size * (mi / denom);
  }
}

__global__ void adam_step_kernel(long long* step) { *step += 1; }

// torch.nn.utils.clip_grad_norm_(parameters, max_norm) (utils/utils.py:390-391), L2 norm over ALL gradient tensors:
//   total = sqrt(sum_t sum_i g_t[i]^2);  every gradient *= min(1, max_norm / (total + 1e-6))
// pass 1: sum of squares (fp32 per thread, fp64 across the block and the grid); pass 2: the scaling, with the
// coefficient computed on the device (no host round trip: the pair captures into the step's CUDA graph).
// grads: device array [n] of gradient addresses.
__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_multi_kernel(const unsigned long long* __restrict__ grads, const long long* __restrict__ sizes,
                         const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_start,
                         double* __restrict__ sq) {
  const int t = chunk_tensor[blockIdx.x];
  const float* __restrict__ g = reinterpret_cast<const float*>(grads[t]);
  const long long e0 = chunk_start[blockIdx.x];
  long long e1 = e0 + OPT_CHUNK;
  if (e1 > sizes[t]) e1 = sizes[t];
  float acc = 0.f;
  for (long long i = e0 + threadIdx.x; i < e1; i += OPT_THREADS) acc = fmaf(g[i], g[i], acc);
  double d = (double)acc;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
  __shared__ double part[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) tot += part[w];
    atomicAdd(sq, tot);
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
grad_clip_multi_kernel(const unsigned long long* __restrict__ grads, const long long* __restrict__ sizes,
                       const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_start,
                       const double* __restrict__ sq, float max_norm, float* __restrict__ total_norm) {
  const float total = (float)sqrt(*sq);
  if (blockIdx.x == 0 && threadIdx.x == 0) *total_norm = total;
  const float coef = max_norm / (total + 1e-6f);
  if (!(coef < 1.f)) return;  // torch multiplies by clamp(coef, max=1): a no-op here
  const int t = chunk_tensor[blockIdx.x];
  float* __restrict__ g = reinterpret_cast<float*>(grads[t]);
  const long long e0 = chunk_start[blockIdx.x];
  long long e1 = e0 + OPT_CHUNK;
  if (e1 > sizes[t]) e1 = sizes[t];
  for (long long i = e0 + threadIdx.x; i < e1; i += OPT_THREADS) g[i] *= coef;
}

// One block: CE(logits, target) + blockdrop policy loss and their gradients.
// logits [N][C], target [N] int64, sel [N][S][M] (0/1 decisions with straight-through gradient), cw [M].
// loss = mean_n CE_n + use_policy * ( sum_m cw[m] * mean_n(correct) * mean_n(selbar[n][m]^2) + gamma * mean_n(1 - correct) )
// -- the product of the two means is the reference's [N] x [N,1] -> [N,N] broadcast (utils/utils.py:180), kept as is.
__global__ void __launch_bounds__(256)
loss_tail_kernel(const float* __restrict__ logits, const long long* __restrict__ target, const float* __restrict__ sel,
                 const float* __restrict__ cw, float gamma, int use_policy, int N, int C, int S, int M,
                 float* __restrict__ loss, float* __restrict__ dlogits, float* __restrict__ dsel) {
  __shared__ float s_ce, s_correct;
  __shared__ float s_pol;
  if (threadIdx.x == 0) { s_ce = 0.f; s_correct = 0.f; s_pol = 0.f; }
  __syncthreads();
  // phase 1: per-sample softmax / CE / argmax (one thread per sample, C is small)
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* l = logits + (long long)n * C;
    float mx = l[0];
    int am = 0;
    for (int c = 1; c < C; ++c)
      if (l[c] > mx) { mx = l[c]; am = c; }   // first maximum, like torch.argmax
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(l[c] - mx);
    const float lse = logf(se) + mx;
    const int tg = (int)target[n];
    atomicAdd(&s_ce, lse - l[tg]);
    atomicAdd(&s_correct, am == tg ? 1.f : 0.f);
    const float inv = 1.f / (float)N;
    for (int c = 0; c < C; ++c) dlogits[(long long)n * C + c] = (expf(l[c] - lse) - (c == tg ? 1.f : 0.f)) * inv;
  }
  __syncthreads();
  const float cbar = s_correct / (float)N;
  // phase 2: policy term and its gradient with respect to the decisions
  for (int i = threadIdx.x; i < N * M; i += blockDim.x) {
    const int n = i / M, m = i - n * M;
    float sb = 0.f;
    for (int s = 0; s < S; ++s) sb += sel[((long long)n * S + s) * M + m];
    sb /= (float)S;
    if (use_policy) atomicAdd(&s_pol, cw[m] * sb * sb);
    const float gsel = use_policy ? cw[m] * cbar * 2.f * sb / ((float)S * (float)N) : 0.f;
    for (int s = 0; s < S; ++s) dsel[((long long)n * S + s) * M + m] = gsel;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = s_ce / (float)N;
    if (use_policy) v += cbar * (s_pol / (float)N) + gamma * (1.f - cbar);
    *loss = v;
  }
}

}  // namespace

extern "C" {

int adamml_sgd_multi(const unsigned long long* table, const long long* sizes, const int* chunk_tensor,
                     const long long* chunk_start, int n_tensors, int n_chunks, float lr, float momentum,
                     float weight_decay, cudaStream_t stream) {
  ADAMML_REQUIRE(table && sizes && chunk_tensor && chunk_start && n_tensors > 0 && n_chunks > 0, "sgd_multi: bad arguments");
  sgd_multi_kernel<<<n_chunks, OPT_THREADS, 0, stream>>>(table, sizes, chunk_tensor, chunk_start, n_tensors, lr, momentum,
                                                         weight_decay);
  return adamml_check_launch("sgd_multi");
}

int adamml_adam_multi(const unsigned long long* table, const long long* sizes, const int* chunk_tensor,
                      const long long* chunk_start, int n_tensors, int n_chunks, float lr, float beta1, float beta2,
                      float eps, float weight_decay, long long* step, cudaStream_t stream) {
  ADAMML_REQUIRE(table && sizes && chunk_tensor && chunk_start && step && n_tensors > 0 && n_chunks > 0,
                 "adam_multi: bad arguments");
  adam_multi_kernel<<<n_chunks, OPT_THREADS, 0, stream>>>(table, sizes, chunk_tensor, chunk_start, n_tensors, lr, beta1,
                                                          beta2, eps, weight_decay, step);
  adam_step_kernel<<<1, 1, 0, stream>>>(step);
  return adamml_check_launch("adam_multi");
}

int adamml_opt_chunk(void) { return OPT_CHUNK; }

int adamml_clip_grad_norm_multi(const unsigned long long* grads, const long long* sizes, const int* chunk_tensor,
                                const long long* chunk_start, int n_tensors, int n_chunks, float max_norm,
                                double* sq_scratch, float* total_norm, cudaStream_t stream) {
  ADAMML_REQUIRE(grads && sizes && chunk_tensor && chunk_start && sq_scratch && total_norm && n_tensors > 0 &&
                     n_chunks > 0 && max_norm >= 0.f,
                 "clip_grad_norm_multi: bad arguments");
  cudaMemsetAsync(sq_scratch, 0, sizeof(double), stream);
  grad_sqnorm_multi_kernel<<<n_chunks, OPT_THREADS, 0, stream>>>(grads, sizes, chunk_tensor, chunk_start, sq_scratch);
  grad_clip_multi_kernel<<<n_chunks, OPT_THREADS, 0, stream>>>(grads, sizes, chunk_tensor, chunk_start, sq_scratch,
                                                               max_norm, total_norm);
  return adamml_check_launch("clip_grad_norm_multi");
}

int adamml_loss_tail(const float* logits, const long long* target, const float* selection, const float* cost_weights,
                     float gamma, int use_policy, int N, int C, int S, int M, float* loss, float* dlogits,
                     float* dselection, cudaStream_t stream) {
  ADAMML_REQUIRE(logits && target && selection && loss && dlogits && dselection && N > 0 && C > 0 && S > 0 && M > 0,
                 "loss_tail: bad arguments");
  ADAMML_REQUIRE(!use_policy || cost_weights, "loss_tail: the policy term needs the cost weights");
  loss_tail_kernel<<<1, 256, 0, stream>>>(logits, target, selection, cost_weights, gamma, use_policy, N, C, S, M, loss,
                                          dlogits, dselection);
  return adamml_check_launch("loss_tail");
}

}  // extern "C"
