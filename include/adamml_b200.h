/* adamml_b200.h — C-ABI of libadamml_b200.so (sm_100a CUDA kernels of the AdaMML hot path).
 *
 * The reference (IBM/AdaMML) has no FFI: every operator below replaces a torch.nn call site
 * inside reference models/*.py (cited per entry point).  The drop-in boundary above this ABI
 * is the Python package adamml_b200.models (build_model / MODEL_TABLE / AdaMML.forward), which
 * binds these symbols with ctypes (adamml_b200/_lib.py; INTEGRATION.md shows the stub).
 *
 * Conventions:
 *  - every pointer is a DEVICE pointer owned by the caller (torch's caching allocator);
 *    the library never allocates, frees or retains memory;
 *  - every call is asynchronous on the caller-supplied `stream`; no hidden synchronisation;
 *  - return value 0 = success; non-zero = error, message via adamml_last_error() (thread local);
 *  - activations are NHWC ("channels last"), images ordered segment-major:
 *        img = (segment * N + video) * frames + frame ;  BN group = segment;
 *  - `dtype` selects the activation type: ADAMML_F32 (0) or ADAMML_BF16 (1); all
 *    accumulation, BN statistics and the whole policy head are fp32/fp64;
 *  - dense conv weights are OHWI ([Cout][R][S][Cin], row stride w_ld) in the activation dtype,
 *    produced from torch's OIHW fp32 parameters by adamml_pack_weight;
 *  - *_ld arguments are element strides between consecutive pixels/rows; pass 0 for "dense".
 */
#ifndef ADAMML_B200_H_
#define ADAMML_B200_H_

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADAMML_F32 0
#define ADAMML_BF16 1
#define ADAMML_ACT_NONE 0
#define ADAMML_ACT_RELU 1
#define ADAMML_ACT_RELU6 2

/* ---- library ---- */
const char* adamml_last_error(void);
int adamml_abi_version(void);
unsigned long long adamml_launch_count(void); /* kernels launched by this process so far */

/* ---- data layer: models/adamml.py:42-67 (AdaMML.data_layer) ---- */
/* x NCHW fp32 [N, S*F*C, H, W] -> NHWC [(s*N+n)*F+f, H, W, Cpad]  (adamml.py:53,65) */
int adamml_pack_frames(const float* x, void* out, int N, int S, int F, int C, int H, int W, int Cpad, int dtype,
                       cudaStream_t stream);
/* F.interpolate(bilinear, align_corners=False) to OHxOW + keep frames 0,fstep,.. (adamml.py:59-62) */
int adamml_resize_frames(const float* x, void* out, int N, int S, int F, int C, int H, int W, int OH, int OW,
                         int fstep, int Cpad, int dtype, cudaStream_t stream);
/* The same two passes (and the stem operand below) on decoded uint8 frames: x is the CHW byte tensor
 * ToTorchFormatTensor holds before .float() (utils/video_transforms.py:321-343); the kernels apply
 * x.float().div(255) (:343) and GroupNormalize's t.sub_(mean).div_(std) per channel plane (:62-84) in the reference's
 * fp32 arithmetic while re-laying out, so H2D traffic is 1 byte per sample.  mean/std: device fp32 [C]. */
int adamml_pack_frames_u8(const unsigned char* x, const float* mean, const float* stdv, void* out, int N, int S, int F,
                          int C, int H, int W, int Cpad, int dtype, cudaStream_t stream);
int adamml_resize_frames_u8(const unsigned char* x, const float* mean, const float* stdv, void* out, int N, int S,
                            int F, int C, int H, int W, int OH, int OW, int fstep, int Cpad, int dtype,
                            cudaStream_t stream);
int adamml_pack_frames_s2d_u8(const unsigned char* x, const float* mean, const float* stdv, void* out, int N, int S,
                              int F, int C, int H, int W, int Cs, cudaStream_t stream);
/* nn.Conv2d.weight OIHW fp32 -> OHWI operand (CinPad >= Cin, zero filled) */
int adamml_pack_weight(const float* w_oihw, void* w_ohwi, int Cout, int Cin, int R, int S, int CinPad, int dtype,
                       cudaStream_t stream);
/* Every weight operand of a model in ONE launch (the per-layer adamml_pack_weight* calls of a training step are ~360
 * small launches).  jobs = device array [n_jobs][8] of int64 {src fp32 OIHW address, dst address, Cout, Cin, R, S,
 * CinPad, kind}; kind 0 | 1 | 2 = OHWI operand in fp32 | bf16 | the four x2 planes (adamml_pack_weight,
 * adamml_pack_weight_x2 with stem = 0), 3 | 4 = rotated data-gradient operand in fp32 | bf16
 * (adamml_pack_weight_dgrad), 5 = tap-major depthwise weights (adamml_pack_weight_dw; Cout = C).  chunk_job /
 * chunk_start [n_chunks]: the job and destination element offset each block of adamml_pack_chunk() elements converts.
 * Results are bit-identical to the per-layer entry points. */
int adamml_pack_chunk(void);
int adamml_pack_weights_multi(const long long* jobs, const int* chunk_job, const long long* chunk_start, int n_jobs,
                              int n_chunks, cudaStream_t stream);
/* nn.Conv2d.weight OIHW fp32 -> [Cin][R][S][Cout] rotated by 180 degrees: the operand that turns the
 * stride-1 data gradient into a forward convolution of dy (pad' = R-1-pad) on the tcgen05 engine */
int adamml_pack_weight_dgrad(const float* w_oihw, void* w_ihwo, int Cout, int Cin, int R, int S, int dtype,
                             cudaStream_t stream);
/* OHWI fp32 weight gradient -> OIHW fp32 .grad layout */
int adamml_unpack_wgrad(const float* dw_ohwi, float* dw_oihw, int Cout, int Cin, int R, int S, int CinPad,
                        int accumulate, cudaStream_t stream);
/* Space-to-depth operands of the tensor-core ResNet stem (7x7/s2/p3 conv1, resnet.py:138,199):
 * frames  -> bf16 [(s*N+n)*F+f, H/2, W/2+4, Cs]  (two zero columns on either side, channel (ph*2+pw)*C+c)
 * weights -> bf16 [Cout][T][T][Cs] (T = 4 for R = 7, T = 2 for R = 3); the fp32 gradient of that operand -> OIHW */
int adamml_pack_frames_s2d(const float* x, void* out, int N, int S, int F, int C, int H, int W, int Cs,
                           cudaStream_t stream);
int adamml_pack_weight_stem(const float* w_oihw, void* w_packed, int Cout, int C, int Cs, int R,
                            cudaStream_t stream);
int adamml_unpack_wgrad_stem(const float* dw_packed, float* dw_oihw, int Cout, int C, int Cs, int R,
                             cudaStream_t stream);
/* NHWC bf16 -> space-to-depth bf16 [IMGS, H/2, W/2+padl+padr, Cs] (operand of the stride-2 3x3 first convolutions
 * of the MobileNetV2s: sound_mobilenet_v2.py:120, policy_net.py:117; R = 3 -> T = 2 taps, padl = 1, padr = 0) */
int adamml_nhwc_to_s2d(const void* x, void* out, long long IMGS, int C, int H, int W, int Cs, int padl, int padr,
                       cudaStream_t stream);
int adamml_cast(const void* src, void* dst, long long total, int src_dtype, int dst_dtype, cudaStream_t stream);

/* ---- dense convolution / linear, exact fp32-math engine (CUDA cores) ----
 * nn.Conv2d at models/resnet.py:35-43,138 ; sound_mobilenet_v2.py:36,61 ; policy_net.py:40,49,66-84
 * nn.Linear at resnet.py:159 ; sound_mobilenet_v2.py:134 ; policy_net.py:229-230,278-279 (H=W=1) */
int adamml_simt_conv_fwd(const void* x, const void* w, void* y, int IMGS, int H, int W, int Cin, int Cout, int R,
                         int S, int stride, int pad, int Ho, int Wo, long long x_ld, long long y_ld, long long w_ld,
                         int dtype, cudaStream_t stream);
/* dx = conv_transpose(dy, w) (+ addend) */
int adamml_simt_conv_dgrad(const void* dy, const void* w, void* dx, const void* addend, int IMGS, int H, int W,
                           int Cin, int Cout, int R, int S, int stride, int pad, int Ho, int Wo, long long x_ld,
                           long long y_ld, long long w_ld, int dtype, cudaStream_t stream);
/* dw (fp32, OHWI, overwritten) = sum_pixels dy (x) x */
int adamml_simt_conv_wgrad(const void* x, const void* dy, float* dw, int IMGS, int H, int W, int Cin, int Cout,
                           int R, int S, int stride, int pad, int Ho, int Wo, long long x_ld, long long y_ld,
                           long long w_ld, int dtype, cudaStream_t stream);

/* ---- dense convolution, tcgen05 tensor-core engine (bf16 operands, fp32 TMEM accumulators) ----
 * Same call sites as above.  GEMM view  D[M, Ncols] = A[M, K] . B[Ncols, K]^T  with A and B both
 * K-major bf16 (row strides lda/ldb elements, multiples of 8), D bf16 or fp32 row-major (ldd).
 * Used for 1x1 convolutions forward (A = activations, B = weights) and dgrad (A = dy, B = w^T).
 * Optional fused epilogue: per-(group, column) sum / sum-of-squares of the fp32 accumulators
 * (train-mode BatchNorm statistics, resnet.py:97,101,105) accumulated into `stats` [G][Ncols][2]
 * (double), rows_per_group rows per BN group.  Returns ADAMML_ERR_UNSUPPORTED (3) for shapes
 * outside its envelope so that the caller can route to the exact engine. */
int adamml_tc_gemm_bf16(const void* A, const void* B, void* D, long long M, int Ncols, int K, long long lda,
                        long long ldb, long long ldd, int d_dtype, double* stats, long long rows_per_group,
                        cudaStream_t stream);
int adamml_tc_supported(long long M, int Ncols, int K, long long lda, long long ldb, long long ldd);
/* Implicit-GEMM convolution on the same tcgen05 pipeline (no im2col buffer): the 128 rows of an M tile are
 * a BWxBHxBI box of output pixels and every filter tap (r,s) is ONE shifted 4D TMA box
 * {64 channels, BW, BH, BI} of the NHWC input (zero padding = TMA out-of-bounds fill; stride 2 = four
 * parity sub-lattice tensor maps).  Call sites: resnet.py:35-38,100 (3x3 conv2 of every Bottleneck, stride
 * 1|2) and :164-167 (strided 1x1 downsample); the stride-1 data gradient runs through the same entry point
 * with adamml_pack_weight_dgrad weights and `addend` = the residual-branch gradient.
 * x [IMGS,H,W,Cin] bf16, w [Cout][R][S][Cin] bf16, y/addend [IMGS,Ho,Wo,Cout] bf16, stats double
 * [G][Cout][2] (optional fused BN statistics, G = IMGS / imgs_per_group). */
/* addend_sub = 2: `addend` is the compact [IMGS, ceil(Ho/2), ceil(Wo/2), Cout] gradient of a stride-2 1x1 branch
 * (the Bottleneck downsample, resnet.py:164-167) and is added at even (oh, ow) only; otherwise 0/1 = dense. */
int adamml_tc_conv_bf16(const void* x, const void* w, void* y, const void* addend, int IMGS, int H, int W, int Cin,
                        int Cout, int R, int S, int stride, int pad, int Ho, int Wo, double* stats,
                        int imgs_per_group, int addend_sub, cudaStream_t stream);
int adamml_tc_conv_supported(int Cin, int Cout, int R, int S, int stride);
/* Data gradient of a stride-2 RxS convolution (R,S >= 2) as four stride-1 implicit GEMMs over the parity classes
 * of the input grid; w_rot from adamml_pack_weight_dgrad; dx [IMGS,H,W,Cin] is fully overwritten. */
int adamml_tc_dgrad_s2_bf16(const void* dy, const void* w_rot, void* dx, int IMGS, int H, int W, int Cin, int Cout,
                            int R, int S, int pad, int Ho, int Wo, cudaStream_t stream);
/* ResNet stem on the space-to-depth operands: y [IMGS,Ho,Wo,Cout] bf16 (+ fused BN statistics), and its weight
 * gradient dw fp32 [Cout][4][4*Cs] (unpack with adamml_unpack_wgrad_stem). */
int adamml_tc_stem_conv_bf16(const void* xs, const void* w, void* y, int IMGS, int Hs, int Wp, int Cs, int Cout,
                             int Ho, int Wo, int taps, double* stats, int imgs_per_group, cudaStream_t stream);
int adamml_tc_stem_wgrad_bf16(const void* xs, const void* dy, float* dw, int IMGS, int Hs, int Wp, int Cs, int Cout,
                              int Ho, int Wo, int taps, cudaStream_t stream);
/* Weight gradient on tcgen05: implicit GEMM whose reduction axis is the pixel axis, both operands MN-major
 * (64-channel x 64-pixel 4D TMA boxes of x and dy), split-K over pixel ranges with fp32 atomics.
 * dw fp32 [Cout][R][S][Cin] is overwritten.  Same call sites as adamml_simt_conv_wgrad. */
int adamml_tc_wgrad_bf16(const void* x, const void* dy, float* dw, int IMGS, int H, int W, int Cin, int Cout, int R,
                         int S, int stride, int pad, int Ho, int Wo, cudaStream_t stream);
int adamml_tc_wgrad_supported(int Cin, int Cout, int R, int S, int stride);

/* ---- depthwise 3x3 conv, pad 1, stride 1|2 ----
 * sound_mobilenet_v2.py:58 ; policy_net.py:66,80.  Weights and weight gradients are fp32 TAP-MAJOR [9][C]
 * (coalesced per-thread loads); adamml_pack_weight_dw / adamml_unpack_wgrad_dw convert from / to torch's
 * [C,1,3,3]. */
int adamml_pack_weight_dw(const float* w_c33, float* w_9c, int C, cudaStream_t stream);
int adamml_unpack_wgrad_dw(const float* dw_9c, float* dw_c33, int C, cudaStream_t stream);
int adamml_dwconv_fwd(const void* x, const float* w, void* y, int IMGS, int H, int W, int C, int stride, int Ho,
                      int Wo, int dtype, cudaStream_t stream);
int adamml_dwconv_dgrad(const void* dy, const float* w, void* dx, const void* addend, int IMGS, int H, int W, int C,
                        int stride, int Ho, int Wo, int dtype, cudaStream_t stream);
int adamml_dwconv_wgrad(const void* x, const void* dy, float* dw, int IMGS, int H, int W, int C, int stride, int Ho,
                        int Wo, int dtype, cudaStream_t stream);
/* Fused depthwise backward (bf16 NHWC, TMA-staged tiles): dx = conv_transpose(dy, w) AND dw (fp32 tap-major [9][C],
 * overwritten) from ONE pass over dy and x -- the autograd of the same call sites.  adamml_dwconv_bwd_supported -> 1
 * when the shape is handled (C % 16 == 0), else the two calls above are used. */
int adamml_dwconv_bwd_supported(int IMGS, int H, int W, int C, int stride);
int adamml_dwconv_bwd(const void* x, const void* dy, const float* w, void* dx, float* dw, double* pre_sums,
                      int pre_imgs_per_group, int pre_act, int IMGS, int H, int W, int C, int stride, int Ho, int Wo,
                      cudaStream_t stream);
/* pre_sums != NULL fuses the BatchNorm-backward REDUCTION of the layer that produced x = act(bn(z)) (the expand / first
 * conv in front of the depthwise conv, no residual): dx is written already masked (gm = dx * act'(x), pre_act = that
 * layer's activation) and pre_sums [G][C][2] (double, overwritten) = per BN group (img / pre_imgs_per_group) and
 * channel (sum gm, sum gm * x); adamml_bn_sums_from_out converts them to the (sum gm, sum gm * xhat) that
 * adamml_bn_bwd_reduce would have produced from (dx, z) in a separate pass. */
int adamml_bn_sums_from_out(const double* raw, const float* scale_shift, const float* mean_invstd, double* sums, int C,
                            int G, cudaStream_t stream);

/* ---- BatchNorm2d (+ReLU/ReLU6, + residual) with per-segment groups ----
 * nn.BatchNorm2d at resnet.py:50,53,82-86,139,166 ; sound_mobilenet_v2.py:37,62 ; policy_net.py:41,50,67-85
 * sums: double [G][C][2]; mean_invstd / scale_shift: float [G][C][2] */
int adamml_bn_stats(const void* z, double* sums, long long rows_per_group, int C, int G, int dtype,
                    cudaStream_t stream);
int adamml_bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float* mean_invstd, float* scale_shift, double count, float momentum,
                       float eps, int C, int G, int training, int update_running, cudaStream_t stream);
/* out = act(z*scale+shift [+ res] [+ res_z*res_scale+res_shift])  (resnet.py:104-111 residual + ReLU) */
int adamml_bn_apply(const void* z, const float* scale_shift, const void* res, const void* res_z,
                    const float* res_scale_shift, void* out, long long rows_per_group, int C, int G, int act,
                    int dtype, cudaStream_t stream);
/* mask_scale_shift (optional, float [G][C][2] = the forward scale/shift of a layer WITHOUT residual input): the
 * ReLU/ReLU6 mask is recomputed as act'(z*scale+shift) and `out` is not read (may be NULL when C is a multiple of
 * the 16-byte vector). */
/* gm_out (optional, may alias dout): the masked gradient gm = dout * act'(out) is written back, so that the following
 * bn_bwd_apply (and the residual branch, whose gradient IS gm) run with act = NONE and never read `out`. */
int adamml_bn_bwd_reduce(const void* dout, const void* out, const void* z, const float* mean_invstd,
                         const float* mask_scale_shift, double* sums, void* gm_out, long long rows_per_group, int C,
                         int G, int act, int dtype, const unsigned char* mask_bits, cudaStream_t stream);
/* (mask_bits: optional [rows][C/8] bytes written by adamml_bn_apply_x2 -- bit i of a byte = activation passes the
 * gradient at channel 8*j + i; when given, `out` is not read: 1 bit instead of 16 per element on residual layers) */
int adamml_bn_bwd_apply(const void* dout, const void* out, const void* z, const float* mean_invstd,
                        const float* gamma, const float* mask_scale_shift, const double* sums, void* dz, void* dres,
                        long long rows_per_group, int C, int G, double count, int act, int training, int dtype,
                        cudaStream_t stream);
int adamml_bn_param_grad(const double* sums, float* dgamma, float* dbeta, int C, int G, int accumulate,
                         cudaStream_t stream);

/* ---- sync-BN statistic exchange over NVLink peer memory (train_adamml.py:125-127 SyncBatchNorm) ----
 * One-shot all-reduce (sum in rank order) of n doubles: every rank's partial sums sit at `slot_off` doubles inside a
 * symmetric buffer; peer_bufs / peer_flags are DEVICE arrays [world] holding the peers' mapped base addresses of the
 * data buffer and of the 32-bit flag array [lanes][world]; epoch_ctr [lanes] and err_flag live in local device
 * memory.  Asynchronous on `stream`, capturable in a CUDA graph; a peer that never arrives sets *err_flag. */
int adamml_p2p_allreduce_f64(const unsigned long long* peer_bufs, const unsigned long long* peer_flags,
                             long long slot_off, double* out, int n, int world, int rank, int lane,
                             unsigned* epoch_ctr, unsigned* err_flag, cudaStream_t stream);

/* ---- pooling ---- */
/* nn.MaxPool2d(3,2,1): resnet.py:141,202.  `pos` (optional, uint8 [IMGS,Ho,Wo,C]) records the window
 * position r*3+s of the first maximum; the backward pass then is a gather over (pos, dy) and x may be NULL. */
int adamml_maxpool3x3s2_fwd(const void* x, void* y, unsigned char* pos, int IMGS, int H, int W, int C, int Ho, int Wo,
                            int dtype, cudaStream_t stream);
int adamml_maxpool3x3s2_bwd(const void* x, const unsigned char* pos, const void* dy, void* dx, int IMGS, int H, int W,
                            int C, int Ho, int Wo, int dtype, cudaStream_t stream);
/* TemporalPooling k3 s2 p1 over frames: common.py:4-33 ; x [V videos][Tn frames][E = H*W*C] */
int adamml_tpool_fwd(const void* x, void* y, long long V, int Tn, long long E, int mode_avg, int dtype,
                     cudaStream_t stream);
int adamml_tpool_bwd(const void* x, const void* dy, void* dx, long long V, int Tn, long long E, int mode_avg,
                     int dtype, cudaStream_t stream);
/* nn.AdaptiveAvgPool2d(1): resnet.py:157,212 ; sound_mobilenet_v2.py:156 ; policy_net.py:136,146 */
int adamml_avgpool_fwd(const void* x, float* y, int IMGS, int HW, int C, long long y_ld, int dtype,
                       cudaStream_t stream);
int adamml_avgpool_bwd(const float* dy, void* dx, int IMGS, int HW, int C, long long dy_ld, int dtype,
                       cudaStream_t stream);
/* torch.mean over the remaining frames of a video: resnet.py:218-221 */
int adamml_frame_mean(const float* x, float* out, long long V, int Tn, int C, long long out_ld, cudaStream_t stream);
int adamml_frame_mean_bwd(const float* dy, float* dx, long long V, int Tn, int C, long long dy_ld,
                          cudaStream_t stream);

/* ---- fp32 helpers for nn.Linear bias / ReLU / Dropout-mask ---- */
int adamml_bias_act(float* y, const float* bias, long long rows, int cols, long long ld, int act,
                    cudaStream_t stream);
int adamml_act_bwd(const float* dy, const float* y, float* dz, long long rows, int cols, long long ld_dy,
                   long long ld_y, long long ld_dz, int act, cudaStream_t stream);
int adamml_colsum(const float* x, float* out, long long rows, int cols, long long ld, int accumulate,
                  cudaStream_t stream);
int adamml_mul(const float* a, const float* b, float* out, long long total, cudaStream_t stream);

/* ---- policy head: LSTMCell + Linear(256,2) x M + hard Gumbel-softmax, one segment step ----
 * policy_net.py:283-290 (wrapper_gumbel_softmax), :345-365 (LSTM loop).
 * gx [N,4Hd] = W_ih[:, :Fdim] . feat (hoisted GEMM); prev_logits/logits/ysoft [M][N][2];
 * expo = Exp(1) samples [M*N][2] in torch's draw order; dec [M][N] in {0,1}. */
int adamml_policy_step_fwd(const float* gx, const float* prev_logits, const float* h_prev, const float* c_prev,
                           const float* w_ih, long long w_ih_ld, int Fdim, const float* w_hh, const float* b_ih,
                           const float* b_hh, const float* fc_w, const float* fc_b, const float* expo, float tau,
                           float* gates_out, float* h_out, float* c_out, float* logits_out, float* ysoft_out,
                           float* dec_out, float* xin_tail, long long xin_ld, int N, int M, int Hd,
                           cudaStream_t stream);
int adamml_policy_step_bwd(const float* d_dec, const float* d_logits_fb, const float* dh_next, const float* dc_next,
                           const float* gates, const float* c_cur, const float* c_prev, const float* ysoft,
                           const float* w_ih, long long w_ih_ld, int Fdim, const float* w_hh, const float* fc_w,
                           float tau, float* dl_out, long long dl_ms, float* dgates_out, float* dh_prev,
                           float* dc_prev, float* d_prev_logits, int N, int M, int Hd, cudaStream_t stream);

/* Stand-alone hard Gumbel-softmax over [R,2] logits: the causality_modeling=None policy (policy_net.py:330-339)
 * calls F.gumbel_softmax once on all (modality, segment, video) rows.  dec = straight-through column 1. */
int adamml_gumbel_hard_fwd(const float* logits, const float* expo, float tau, float* ysoft, float* dec, long long R,
                           cudaStream_t stream);
int adamml_gumbel_hard_bwd(const float* d_dec, const float* ysoft, float tau, float* dlogits, long long R,
                           cudaStream_t stream);

/* ---- gate x logits, late-fusion weights, sum over modalities, mean over segments ----
 * joint_resnet_mobilenetv2.py:92-97,112-127 ; adamml.py:88.
 * logits [M][S][N][C]; dec [S][M][N] (NULL = ungated); lf [M-1] (NULL = plain mean); out [N][C] */
int adamml_fuse_fwd(const float* logits, const float* dec, const float* lf, float* out, int M, int S, int N, int C,
                    cudaStream_t stream);
int adamml_fuse_bwd(const float* g, const float* logits, const float* dec, const float* lf, float* dlogits,
                    float* ddec, float* dlf, int M, int S, int N, int C, cudaStream_t stream);

/* ---- inference-mode fused epilogue: conv + BatchNorm (+ residual) + ReLU/ReLU6 in ONE kernel ----
 * resnet.py:96-111 (Bottleneck: conv -> bn -> relu, conv3 -> bn3 -> += identity -> relu), sound_mobilenet_v2.py:33-40
 * and policy_net.py:38-52 (ConvBNReLU6).  With running statistics BatchNorm is a per-channel affine map known before
 * the convolution runs, so out = act(acc * scale[c] + shift[c] (+ res)) leaves the tcgen05 accumulator directly:
 * the pre-BN tensor is never written and there is no separate BN / add / ReLU pass.  scale_shift = [Cout][2] fp32
 * (group 0 of adamml_bn_finalize in eval mode); res = optional tensor of the output's shape (fetched by TMA into the
 * epilogue's staging tile); act applies after the addend.  Training uses batch statistics, which only exist after the
 * convolution: the train path keeps the fused-statistics epilogue + adamml_bn_apply. */
int adamml_tc_gemm_bn_act_bf16(const void* A, const void* B, void* D, long long M, int Ncols, int K,
                               const float* scale_shift, int act, const void* res, cudaStream_t stream);
int adamml_tc_conv_bn_act_bf16(const void* x, const void* w, void* y, int IMGS, int H, int W, int Cin, int Cout, int R,
                               int S, int stride, int pad, int Ho, int Wo, const float* scale_shift, int act,
                               const void* res, cudaStream_t stream);
int adamml_tc_stem_conv_bn_act_bf16(const void* xs, const void* w, void* y, int IMGS, int Hs, int Wp, int Cs, int Cout,
                                    int Ho, int Wo, int taps, const float* scale_shift, int act,
                                    cudaStream_t stream);
int adamml_dwconv_bn_act_fwd(const void* x, const float* w, void* y, int IMGS, int H, int W, int C, int stride, int Ho,
                             int Wo, const float* scale_shift, int act, int dtype, cudaStream_t stream);

/* ---- train-step tail (utils/utils.py:362-400, train_adamml.py:250-257) ----
 * adamml_loss_tail: cross-entropy + 'blockdrop' policy loss (utils/utils.py:166-184, including its [N] x [N,1]
 * broadcast) and both gradients in one launch.  logits [N][C] fp32, target int64 [N], selection [N][S][M] fp32,
 * cost_weights fp32 [M]; loss fp32 [1], dlogits [N][C], dselection [N][S][M] (gradients of the loss itself).
 * adamml_sgd_multi / adamml_adam_multi: torch.optim.SGD (momentum, dampening 0, L2 weight decay) / torch.optim.Adam
 * (bias correction, L2 weight decay, no amsgrad) over ALL tensors of a parameter group in one launch.  table = device
 * array [2 + states][n_tensors] of addresses (param, grad, momentum | exp_avg, exp_avg_sq), sizes [n_tensors] element
 * counts, chunk_tensor / chunk_start [n_chunks] = the tensor and element offset each block of adamml_opt_chunk()
 * elements works on.  The Adam step counter `step` (int64 [1]) lives on the device and is advanced by the call. */
int adamml_loss_tail(const float* logits, const long long* target, const float* selection, const float* cost_weights,
                     float gamma, int use_policy, int N, int C, int S, int M, float* loss, float* dlogits,
                     float* dselection, cudaStream_t stream);
int adamml_sgd_multi(const unsigned long long* table, const long long* sizes, const int* chunk_tensor,
                     const long long* chunk_start, int n_tensors, int n_chunks, float lr, float momentum,
                     float weight_decay, cudaStream_t stream);
int adamml_adam_multi(const unsigned long long* table, const long long* sizes, const int* chunk_tensor,
                      const long long* chunk_start, int n_tensors, int n_chunks, float lr, float beta1, float beta2,
                      float eps, float weight_decay, long long* step, cudaStream_t stream);
int adamml_opt_chunk(void);
/* torch.nn.utils.clip_grad_norm_(parameters, max_norm), L2 (utils/utils.py:390-391, opts.py:75 --clip_gradient): total =
 * sqrt(sum of squares over ALL gradient tensors), every gradient *= min(1, max_norm / (total + 1e-6)).  grads = device
 * array [n_tensors] of fp32 gradient addresses; sizes / chunk_tensor / chunk_start as above; sq_scratch fp64 [1]
 * (overwritten), total_norm fp32 [1] receives the norm BEFORE clipping.  No host synchronisation. */
int adamml_clip_grad_norm_multi(const unsigned long long* grads, const long long* sizes, const int* chunk_tensor,
                                const long long* chunk_start, int n_tensors, int n_chunks, float max_norm,
                                double* sq_scratch, float* total_norm, cudaStream_t stream);

/* ---- device-side gating: inference with decision-driven skipping, no host round trip ----
 * The reference runs every main backbone on every (segment, video) pair and multiplies its logits by the policy's 0/1
 * decision (models/adamml.py:81-86, joint_resnet_mobilenetv2.py:92-94).  With running-statistic BatchNorm an unselected
 * pair contributes exactly zero, so it is skipped -- on the device: adamml_select_compact turns the decisions
 * [S][M][N] of modality m into the ascending list idx[] of selected pairs p = s*N + n and their number *count;
 * adamml_gather_rows moves those clips to the front of a static-capacity batch buffer; adamml_set_live_clips arms a
 * (thread-local) work limit that the following inference launches of the calling thread honour by reading *count on
 * the device (tcgen05 tile loops, depthwise / pooling / s2d threads beyond the live prefix exit); adamml_scatter_rows_f32
 * puts the logits back (zeros elsewhere).  Nothing depends on the count on the host, so the pass captures into a CUDA
 * graph.  adamml_set_live_clips(NULL, 0) disarms the limit. */
int adamml_select_compact(const float* decisions, int S, int M, int N, int m, int* idx, int* count,
                          cudaStream_t stream);
int adamml_gather_rows(const void* src, void* dst, const int* idx, const int* count, long long row_bytes, int capacity,
                       cudaStream_t stream);
int adamml_scatter_rows_f32(const float* y, const int* idx, const int* count, float* out, int rows_out, int C,
                            cudaStream_t stream);
int adamml_set_live_clips(const int* live_clips, int clip_capacity);

/* ---- "x2" forward path: two-plane activations (default precision mode) ----
 * north_star asks for logits within 1e-3 of the reference's fp32 path and bit-exact policy selections; bf16
 * storage (8 mantissa bits) misses that by two orders of magnitude on these 50-layer BatchNorm stacks.  In x2 mode
 * every forward activation, weight operand and conv output is stored as TWO planes, hi = bf16(v) and
 * lo = fp16(v - hi) (about 20 mantissa bits at bf16 range, 4 bytes per element).  A tcgen05.mma needs both operands
 * in the same 16-bit format, so the (small) weight operands are packed as FOUR planes `w4` [4][Cout][R][S][Cin]:
 * the bf16 cascade b1 = bf16(w), b2 = bf16(w - b1), b3 = bf16(w - b1 - b2) and f = fp16(w); every K step issues
 * x_hi*b1 + x_hi*b2 + x_hi*b3 (bf16 x bf16) + x_lo*f (fp16 x fp16) into one fp32 TMEM accumulator.
 * The hi planes are ordinary bf16 tensors: they are what the backward pass keeps and reads (bf16 engine above), the
 * lo planes die with the forward pass.  Same reference call sites as the entry points they mirror; every `_hi` /
 * `_lo` pair has the shape of the bf16 tensor it replaces, channel counts are multiples of 8. */
#define ADAMML_X2 2
int adamml_pack_frames_x2(const void* x, const float* mean, const float* stdv, void* out_hi, void* out_lo, int N, int S,
                          int F, int C, int H, int W, int Cpad, int is_u8, cudaStream_t stream);
int adamml_resize_frames_x2(const void* x, const float* mean, const float* stdv, void* out_hi, void* out_lo, int N,
                            int S, int F, int C, int H, int W, int OH, int OW, int fstep, int Cpad, int is_u8,
                            cudaStream_t stream);
int adamml_pack_frames_s2d_x2(const void* x, const float* mean, const float* stdv, void* out_hi, void* out_lo, int N,
                              int S, int F, int C, int H, int W, int Cs, int is_u8, cudaStream_t stream);
/* OIHW fp32 -> the four OHWI planes w4 (stem = 1: the space-to-depth first-conv operand, Cin = C, CinPad = Cs,
 * R = S = 7 | 3) */
int adamml_pack_weight_x2(const float* w_oihw, void* w4, int Cout, int Cin, int R, int S, int CinPad, int stem,
                          cudaStream_t stream);
int adamml_tc_gemm_x2(const void* A_hi, const void* A_lo, const void* B4, void* D_hi, void* D_lo, long long M,
                      int Ncols, int K, double* stats, long long rows_per_group, cudaStream_t stream);
int adamml_tc_conv_x2(const void* x_hi, const void* x_lo, const void* w4, void* y_hi, void* y_lo, int IMGS, int H,
                      int W, int Cin, int Cout, int R, int S, int stride, int pad, int Ho, int Wo, double* stats,
                      int imgs_per_group, cudaStream_t stream);
int adamml_tc_stem_conv_x2(const void* xs_hi, const void* xs_lo, const void* w4, void* y_hi, void* y_lo, int IMGS,
                           int Hs, int Wp, int Cs, int Cout, int Ho, int Wo, int taps, double* stats,
                           int imgs_per_group, cudaStream_t stream);
int adamml_tc_gemm_bn_act_x2(const void* A_hi, const void* A_lo, const void* B4, void* D_hi, void* D_lo, long long M,
                             int Ncols, int K, const float* scale_shift, int act, const void* res_hi,
                             const void* res_lo, cudaStream_t stream);
int adamml_tc_conv_bn_act_x2(const void* x_hi, const void* x_lo, const void* w4, void* y_hi, void* y_lo, int IMGS,
                             int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int Ho, int Wo,
                             const float* scale_shift, int act, const void* res_hi, const void* res_lo,
                             cudaStream_t stream);
int adamml_tc_stem_conv_bn_act_x2(const void* xs_hi, const void* xs_lo, const void* w4, void* y_hi, void* y_lo,
                                  int IMGS, int Hs, int Wp, int Cs, int Cout, int Ho, int Wo, int taps,
                                  const float* scale_shift, int act, cudaStream_t stream);
int adamml_dwconv_bn_act_fwd_x2(const void* x_hi, const void* x_lo, const float* w, void* y_hi, void* y_lo, int IMGS,
                                int H, int W, int C, int stride, int Ho, int Wo, const float* scale_shift, int act,
                                cudaStream_t stream);
int adamml_dwconv_fwd_x2(const void* x_hi, const void* x_lo, const float* w, void* y_hi, void* y_lo, int IMGS, int H,
                         int W, int C, int stride, int Ho, int Wo, cudaStream_t stream);
/* x2 training forward of a depthwise 3x3 / stride-1 conv on TMA-staged tiles with the train-mode BatchNorm statistics
 * of its output fused (sound_mobilenet_v2.py:58,62 ; policy_net.py:66-67,80-81): sums = double [G][C][2] (sum, sum of
 * squares) per group = img / imgs_per_group, overwritten; NULL = no statistics. */
int adamml_dwconv_fwd_stats_x2_supported(int IMGS, int H, int W, int C, int stride);
int adamml_dwconv_fwd_stats_x2(const void* x_hi, const void* x_lo, const float* w, void* y_hi, void* y_lo,
                               double* sums, int IMGS, int H, int W, int C, int imgs_per_group, cudaStream_t stream);
int adamml_bn_stats_x2(const void* z_hi, const void* z_lo, double* sums, long long rows_per_group, int C, int G,
                       cudaStream_t stream);
int adamml_bn_apply_x2(const void* z_hi, const void* z_lo, const float* scale_shift, const void* res_hi,
                       const void* res_lo, const void* resz_hi, const void* resz_lo, const float* res_scale_shift,
                       void* out_hi, void* out_lo, long long rows_per_group, int C, int G, int act,
                       unsigned char* mask_bits, cudaStream_t stream);
/* training stem: maxpool3x3s2(act(bn(z))) from the pre-BN planes in one pass (resnet.py:197-200: bn1, relu, maxpool);
 * the full-resolution post-activation tensor is never written */
int adamml_bn_act_maxpool3x3s2_fwd_x2(const void* z_hi, const void* z_lo, const float* scale_shift, int imgs_per_group,
                                      int act, void* y_hi, void* y_lo, unsigned char* pos, int IMGS, int H, int W,
                                      int C, int Ho, int Wo, cudaStream_t stream);
int adamml_maxpool3x3s2_fwd_x2(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, unsigned char* pos,
                               int IMGS, int H, int W, int C, int Ho, int Wo, cudaStream_t stream);
int adamml_tpool_fwd_x2(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, long long V, int Tn, long long E,
                        int mode_avg, cudaStream_t stream);
int adamml_avgpool_fwd_x2(const void* x_hi, const void* x_lo, float* y, int IMGS, int HW, int C, long long y_ld,
                          cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ADAMML_B200_H_ */
