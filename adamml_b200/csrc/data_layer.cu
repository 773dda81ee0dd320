// Input re-layout kernels: the AdaMML data_layer (reference models/adamml.py:42-67) and
// weight / weight-gradient re-packing between torch's OIHW parameters and the OHWI
// operand layout of the convolution engines.
#include "common.cuh"

namespace {

inline int ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  long long cap = 148LL * 64;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// four-plane x2 weight operand (see adamml_pack_weight_x2): plane stride n elements
struct W4Ptr {
  bf16* p;
  long long n;
};
using ::st_elem;  // keep the plain / x2 overloads of common.cuh visible next to this one
__device__ __forceinline__ void st_elem(const W4Ptr& w, long long i, float v) {
  const bf16 b1 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(b1);
  const bf16 b2 = __float2bfloat16_rn(r1);
  const bf16 b3 = __float2bfloat16_rn(r1 - __bfloat162float(b2));
  w.p[i] = b1;
  w.p[w.n + i] = b2;
  w.p[2 * w.n + i] = b3;
  reinterpret_cast<__half*>(w.p)[3 * w.n + i] = __float2half_rn(v);
}

// Input element of the data layer.  fp32 clips arrive normalised.  uint8 clips (decoded frames, the CHW byte
// tensor ToTorchFormatTensor holds before .float()) are normalised here with the reference's own fp32 arithmetic:
// x.float().div(255) (utils/video_transforms.py:343) then t.sub_(mean).div_(std) per channel plane (:81-82).
struct InNorm {
  const float* mean;  // [C] per-frame channel, device memory (uint8 input only)
  const float* stdv;
};
__device__ __forceinline__ float in_val(float v, const InNorm&, int) { return v; }
__device__ __forceinline__ float in_val(unsigned char v, const InNorm& nm, int c) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), nm.mean[c]), nm.stdv[c]);
}
template <typename TIn> struct Pair;
template <> struct Pair<float> { using type = float2; };
template <> struct Pair<unsigned char> { using type = uchar2; };

// x: NCHW [N, S*F*C, H, W]  ->  out: NHWC [(s*N+n)*F+f, H, W, Cpad]   (adamml.py:53,65)
template <typename TIn, typename MP>
__global__ void pack_frames_kernel(const TIn* __restrict__ x, InNorm nm, MP out, int N, int S, int F,
                                   int C, int H, int W, int Cpad) {
  long long total = (long long)S * N * F * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int w = (int)(idx % W);
    int h = (int)((idx / W) % H);
    long long img = idx / ((long long)W * H);
    int f = (int)(img % F);
    int n = (int)((img / F) % N);
    int s = (int)(img / ((long long)F * N));
    const TIn* src = x + (((long long)n * S * F * C + ((long long)s * F + f) * C) * H + h) * W + w;
    const MP dst = out + idx * Cpad;
    for (int c = 0; c < C; ++c) st_elem(dst, c, in_val(src[(long long)c * H * W], nm, c));
    for (int c = C; c < Cpad; ++c) st_elem(dst, c, 0.f);
  }
}

// bilinear, align_corners=False, no antialias (F.interpolate at adamml.py:59), keeping
// frames 0, fstep, 2*fstep, ... of each segment (adamml.py:60-62).
template <typename TIn, typename MP>
__global__ void resize_frames_kernel(const TIn* __restrict__ x, InNorm nm, MP out, int N, int S, int F,
                                     int C, int H, int W, int OH, int OW, int fstep, int Fk, int Cpad) {
  const float sh = (float)H / (float)OH;
  const float sw = (float)W / (float)OW;
  long long total = (long long)S * N * Fk * OH * OW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int ow = (int)(idx % OW);
    int oh = (int)((idx / OW) % OH);
    long long img = idx / ((long long)OW * OH);
    int fk = (int)(img % Fk);
    int n = (int)((img / Fk) % N);
    int s = (int)(img / ((long long)Fk * N));
    int f = fk * fstep;
    float hr = sh * ((float)oh + 0.5f) - 0.5f;
    if (hr < 0.f) hr = 0.f;
    float wr = sw * ((float)ow + 0.5f) - 0.5f;
    if (wr < 0.f) wr = 0.f;
    int h0 = (int)hr; if (h0 > H - 1) h0 = H - 1;
    int w0 = (int)wr; if (w0 > W - 1) w0 = W - 1;
    int hp = h0 < H - 1 ? 1 : 0;
    int wp = w0 < W - 1 ? 1 : 0;
    float lh1 = fminf(fmaxf(hr - (float)h0, 0.f), 1.f), lh0 = 1.f - lh1;
    float lw1 = fminf(fmaxf(wr - (float)w0, 0.f), 1.f), lw0 = 1.f - lw1;
    const TIn* src = x + ((long long)n * S * F * C + ((long long)s * F + f) * C) * H * W;
    const MP dst = out + idx * Cpad;
    for (int c = 0; c < C; ++c) {
      const TIn* pl = src + (long long)c * H * W;
      float p00 = in_val(pl[(long long)h0 * W + w0], nm, c);
      float p01 = in_val(pl[(long long)h0 * W + w0 + wp], nm, c);
      float p10 = in_val(pl[(long long)(h0 + hp) * W + w0], nm, c);
      float p11 = in_val(pl[(long long)(h0 + hp) * W + w0 + wp], nm, c);
      float v = lh0 * (lw0 * p00 + lw1 * p01) + lh1 * (lw0 * p10 + lw1 * p11);
      st_elem(dst, c, v);
    }
    for (int c = C; c < Cpad; ++c) st_elem(dst, c, 0.f);
  }
}

// OIHW fp32 -> OHWI T
template <typename MP>
__global__ void pack_weight_kernel(const float* __restrict__ src, MP dst, int Cout, int Cin, int R,
                                   int S, int CinPad) {
  long long total = (long long)Cout * R * S * CinPad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int ci = (int)(idx % CinPad);
    int s = (int)((idx / CinPad) % S);
    int r = (int)((idx / ((long long)CinPad * S)) % R);
    int co = (int)(idx / ((long long)CinPad * S * R));
    float v = ci < Cin ? src[(((long long)co * Cin + ci) * R + r) * S + s] : 0.f;
    st_elem(dst, idx, v);
  }
}

// OIHW fp32 -> dgrad operand [Cin][R][S][Cout] with the filter rotated by 180 degrees:
// dst[ci][r][s][co] = src[co][ci][R-1-r][S-1-s]   (dx = conv(dy, dst, pad = R-1-pad) for stride 1)
template <typename T>
__global__ void pack_weight_dgrad_kernel(const float* __restrict__ src, T* __restrict__ dst, int Cout, int Cin, int R,
                                         int S) {
  long long total = (long long)Cin * R * S * Cout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int co = (int)(idx % Cout);
    int s = (int)((idx / Cout) % S);
    int r = (int)((idx / ((long long)Cout * S)) % R);
    int ci = (int)(idx / ((long long)Cout * S * R));
    dst[idx] = from_f32<T>(src[(((long long)co * Cin + ci) * R + (R - 1 - r)) * S + (S - 1 - s)]);
  }
}

// OHWI fp32 (CinPad channels) -> OIHW fp32 ; dst = (accumulate ? dst : 0) + src
__global__ void unpack_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin, int R,
                                    int S, int CinPad, int accumulate) {
  long long total = (long long)Cout * Cin * R * S;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int s = (int)(idx % S);
    int r = (int)((idx / S) % R);
    int ci = (int)((idx / ((long long)S * R)) % Cin);
    int co = (int)(idx / ((long long)S * R * Cin));
    float v = src[(((long long)co * R + r) * S + s) * CinPad + ci];
    dst[idx] = accumulate ? dst[idx] + v : v;
  }
}


// Space-to-depth input of the ResNet stem (see conv_tc.cu adamml_tc_stem_conv_bf16):
// x NCHW fp32 [N, S*F*C, H, W] -> out bf16 [(s*N+n)*F+f, H/2, W/2+4, Cs], stored column ip holds s2d column
// ip-2 (two zero columns left and right), channel (ph*2+pw)*C + c = x[.., c, 2j+ph, 2i+pw], rest zero.
// CT/CST > 0: compile-time channel counts, so the pixel is assembled in registers and stored as 16-byte vectors
// out_lo != nullptr: x2 planes (hi = bf16, lo = fp16 remainder)
template <typename TIn, int CT, int CST>
__global__ void pack_frames_s2d_kernel(const TIn* __restrict__ x, InNorm nm, bf16* __restrict__ out,
                                       __half* __restrict__ out_lo, int N, int S, int F, int C, int H, int W, int Cs) {
  using P2 = typename Pair<TIn>::type;
  const int Hs = H / 2, Ws = W / 2, Wp = Ws + 4;
  const long long total = (long long)S * N * F * Hs * Wp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ip = (int)(idx % Wp);
    const int j = (int)((idx / Wp) % Hs);
    const long long img = idx / ((long long)Wp * Hs);
    const int f = (int)(img % F);
    const int n = (int)((img / F) % N);
    const int s = (int)(img / ((long long)F * N));
    bf16* dst = out + idx * Cs;
    __half* dlo = out_lo ? out_lo + idx * Cs : nullptr;
    const int i = ip - 2;
    if (i < 0 || i >= Ws) {
      for (int c = 0; c < Cs; c += 8) {
        *reinterpret_cast<uint4*>(dst + c) = make_uint4(0, 0, 0, 0);
        if (dlo) *reinterpret_cast<uint4*>(dlo + c) = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    const TIn* src = x + (((long long)n * S * F * C + ((long long)s * F + f) * C) * H + 2 * j) * W + 2 * i;
    if (CT > 0) {
      constexpr int CSV = CST > 0 ? CST : 8;
      __align__(16) bf16 v[CSV];
      __align__(16) __half vl[CSV];
#pragma unroll
      for (int c = 0; c < CSV; ++c) { v[c] = __float2bfloat16_rn(0.f); vl[c] = __float2half_rn(0.f); }
#pragma unroll
      for (int ph = 0; ph < 2; ++ph)
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          const P2 t = *reinterpret_cast<const P2*>(src + ((long long)c * H + ph) * W);
          x2_split(in_val(t.x, nm, c), v[(ph * 2 + 0) * CT + c], vl[(ph * 2 + 0) * CT + c]);
          x2_split(in_val(t.y, nm, c), v[(ph * 2 + 1) * CT + c], vl[(ph * 2 + 1) * CT + c]);
        }
#pragma unroll
      for (int c = 0; c < CSV; c += 8) {
        *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<const uint4*>(v + c);
        if (dlo) *reinterpret_cast<uint4*>(dlo + c) = *reinterpret_cast<const uint4*>(vl + c);
      }
    } else {
      for (int ph = 0; ph < 2; ++ph)
        for (int c = 0; c < C; ++c) {
          const P2 v = *reinterpret_cast<const P2*>(src + ((long long)c * H + ph) * W);
          bf16 h0, h1;
          __half l0, l1;
          x2_split(in_val(v.x, nm, c), h0, l0);
          x2_split(in_val(v.y, nm, c), h1, l1);
          dst[(ph * 2 + 0) * C + c] = h0;
          dst[(ph * 2 + 1) * C + c] = h1;
          if (dlo) { dlo[(ph * 2 + 0) * C + c] = l0; dlo[(ph * 2 + 1) * C + c] = l1; }
        }
      for (int c = 4 * C; c < Cs; ++c) {
        dst[c] = __float2bfloat16_rn(0.f);
        if (dlo) dlo[c] = __float2half_rn(0.f);
      }
    }
  }
}

// NHWC bf16 [IMGS, H, W, C] -> space-to-depth bf16 [IMGS, H/2, W/2 + padl + padr, Cs] (zero pad columns, channel
// (ph*2+pw)*C + c): operand of the stride-2 first convolutions of the MobileNetV2s on the tensor-core path.
__global__ void nhwc_to_s2d_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long IMGS, int C, int H,
                                   int W, int Cs, int padl, int padr, LiveLimit live) {
  const int Hs = H / 2, Ws = W / 2, Wp = Ws + padl + padr;
  const long long total = live_count(live, IMGS) * Hs * Wp;  // (device-side work limit: inference with skipping)
  const bf16 zero = __float2bfloat16_rn(0.f);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ip = (int)(idx % Wp);
    const int j = (int)((idx / Wp) % Hs);
    const long long img = idx / ((long long)Wp * Hs);
    bf16* dst = out + idx * Cs;
    const int i = ip - padl;
    if (i < 0 || i >= Ws) {
      for (int c = 0; c < Cs; ++c) dst[c] = zero;
      continue;
    }
    const bf16* src = x + ((img * H + 2 * j) * W + 2 * i) * C;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        for (int c = 0; c < C; ++c) dst[(ph * 2 + pw) * C + c] = src[((long long)ph * W + pw) * C + c];
    for (int c = 4 * C; c < Cs; ++c) dst[c] = zero;
  }
}

// stride-2 first-conv weight OIHW fp32 [Cout][C][R][R] (R = 7, pad 3 -> T = 4 taps per axis; R = 3, pad 1 -> T = 2)
// -> bf16 [Cout][T (dh + T/2)][T (dw + T/2)][Cs]: tap (r, s) = (2*dh+ph+pad, 2*dw+pw+pad), channel (ph*2+pw)*C + c;
// everything else zero.
template <typename MP>
__global__ void pack_weight_stem_kernel(const float* __restrict__ src, MP dst, int Cout, int C,
                                        int Cs, int R, int T) {
  const int pad = R / 2;
  const long long total = (long long)Cout * T * T * Cs;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx % Cs);
    const int dwi = (int)((idx / Cs) % T);
    const int dhi = (int)((idx / (Cs * T)) % T);
    const int co = (int)(idx / ((long long)Cs * T * T));
    float v = 0.f;
    if (q < 4 * C) {
      const int c = q % C, pp = q / C, ph = pp >> 1, pw = pp & 1;
      const int r = 2 * (dhi - T / 2) + ph + pad, s_ = 2 * (dwi - T / 2) + pw + pad;
      if (r >= 0 && r < R && s_ >= 0 && s_ < R) v = src[(((long long)co * C + c) * R + r) * R + s_];
    }
    st_elem(dst, idx, v);
  }
}

// inverse of pack_weight_stem for the fp32 gradient: [Cout][T][T][Cs] -> OIHW [Cout][C][R][R]
__global__ void unpack_wgrad_stem_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int C,
                                         int Cs, int R, int T) {
  const int pad = R / 2;
  const long long total = (long long)Cout * C * R * R;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int s_ = (int)(idx % R);
    const int r = (int)((idx / R) % R);
    const int c = (int)((idx / (R * R)) % C);
    const int co = (int)(idx / ((long long)R * R * C));
    const int t = r - pad, u = s_ - pad;
    const int dh = (t >= 0 ? t : t - 1) / 2, dw = (u >= 0 ? u : u - 1) / 2;  // floor division
    const int ph = t - 2 * dh, pw = u - 2 * dw;
    dst[idx] = src[(((long long)co * T + dh + T / 2) * T + dw + T / 2) * Cs + (ph * 2 + pw) * C + c];
  }
}

// depthwise weight [C][3][3] (torch [C,1,3,3]) <-> tap-major [9][C]
__global__ void transpose_dw_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int to_tap_major) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < C * 9; idx += gridDim.x * blockDim.x) {
    const int c = idx / 9, t = idx % 9;
    if (to_tap_major) dst[t * C + c] = src[idx];
    else dst[idx] = src[t * C + c];
  }
}

// ---- every weight operand of a model in ONE launch -------------------------------------------------------------
// The engine re-derives its weight operands from the fp32 OIHW parameters every step (they change with every optimizer
// step): OHWI in the compute precision for the forward pass, the rotated [Cin][R][S][Cout] operand for the data
// gradient, tap-major depthwise weights -- one small launch per layer and kind, ~360 launches per RGB+Audio step.  This
// kernel takes a job table instead: block b converts elements [chunk_start[b], + PACK_CHUNK) of job chunk_job[b].
// job = 8 x int64: {src, dst, Cout, Cin, R, S, CinPad, kind}; element maps are those of pack_weight_kernel,
// pack_weight_dgrad_kernel and transpose_dw_kernel (bit-identical results).
constexpr int PACK_CHUNK = 4096;
enum { PK_OHWI_F32 = 0, PK_OHWI_BF16 = 1, PK_OHWI_X2 = 2, PK_DGRAD_F32 = 3, PK_DGRAD_BF16 = 4, PK_DW = 5 };
__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const long long* __restrict__ jobs, const int* __restrict__ chunk_job,
                          const long long* __restrict__ chunk_start) {
  const long long* j = jobs + (long long)chunk_job[blockIdx.x] * 8;
  const float* __restrict__ src = reinterpret_cast<const float*>(j[0]);
  void* dst = reinterpret_cast<void*>(j[1]);
  const int Cout = (int)j[2], Cin = (int)j[3], R = (int)j[4], S = (int)j[5], CinPad = (int)j[6], kind = (int)j[7];
  long long total;
  if (kind == PK_DW) total = (long long)Cout * 9;
  else if (kind == PK_DGRAD_F32 || kind == PK_DGRAD_BF16) total = (long long)Cin * R * S * Cout;
  else total = (long long)Cout * R * S * CinPad;
  const long long e0 = chunk_start[blockIdx.x];
  long long e1 = e0 + PACK_CHUNK;
  if (e1 > total) e1 = total;
  for (long long idx = e0 + threadIdx.x; idx < e1; idx += 256) {
    if (kind == PK_DW) {  // [C][3][3] -> [9][C]
      const int c = (int)(idx / 9), t = (int)(idx % 9);
      reinterpret_cast<float*>(dst)[(long long)t * Cout + c] = src[idx];
    } else if (kind == PK_DGRAD_F32 || kind == PK_DGRAD_BF16) {
      const int co = (int)(idx % Cout);
      const int s_ = (int)((idx / Cout) % S);
      const int r = (int)((idx / ((long long)Cout * S)) % R);
      const int ci = (int)(idx / ((long long)Cout * S * R));
      const float v = src[(((long long)co * Cin + ci) * R + (R - 1 - r)) * S + (S - 1 - s_)];
      if (kind == PK_DGRAD_F32) reinterpret_cast<float*>(dst)[idx] = v;
      else reinterpret_cast<bf16*>(dst)[idx] = __float2bfloat16_rn(v);
    } else {
      const int ci = (int)(idx % CinPad);
      const int s_ = (int)((idx / CinPad) % S);
      const int r = (int)((idx / ((long long)CinPad * S)) % R);
      const int co = (int)(idx / ((long long)CinPad * S * R));
      const float v = ci < Cin ? src[(((long long)co * Cin + ci) * R + r) * S + s_] : 0.f;
      if (kind == PK_OHWI_F32) reinterpret_cast<float*>(dst)[idx] = v;
      else if (kind == PK_OHWI_BF16) reinterpret_cast<bf16*>(dst)[idx] = __float2bfloat16_rn(v);
      else st_elem(W4Ptr{reinterpret_cast<bf16*>(dst), total}, idx, v);
    }
  }
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ src, TO* __restrict__ dst, long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x)
    dst[idx] = from_f32<TO>(to_f32(src[idx]));
}

}  // namespace

template <typename TIn>
static int pack_frames_any(const TIn* x, InNorm nm, void* out, void* out_lo, int N, int S, int F, int C, int H, int W,
                           int Cpad, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(N > 0 && S > 0 && F > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "pack_frames: bad dims");
  long long total = (long long)S * N * F * H * W;
  if (dtype == ADAMML_X2) {
    ADAMML_REQUIRE(out_lo, "pack_frames: x2 output needs the lo plane");
    pack_frames_kernel<TIn, X2Ptr><<<ew_blocks(total), 256, 0, stream>>>(x, nm, x2m(out, out_lo), N, S, F, C, H, W, Cpad);
    return adamml_check_launch("pack_frames");
  }
  ADAMML_DISPATCH_DTYPE(dtype, T,
    (pack_frames_kernel<TIn, T*><<<ew_blocks(total), 256, 0, stream>>>(x, nm, (T*)out, N, S, F, C, H, W, Cpad)));
  return adamml_check_launch("pack_frames");
}

template <typename TIn>
static int resize_frames_any(const TIn* x, InNorm nm, void* out, void* out_lo, int N, int S, int F, int C, int H, int W,
                             int OH, int OW, int fstep, int Cpad, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(N > 0 && S > 0 && F > 0 && C > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && fstep > 0 && Cpad >= C,
                 "resize_frames: bad dims");
  int Fk = (F + fstep - 1) / fstep;
  long long total = (long long)S * N * Fk * OH * OW;
  if (dtype == ADAMML_X2) {
    ADAMML_REQUIRE(out_lo, "resize_frames: x2 output needs the lo plane");
    resize_frames_kernel<TIn, X2Ptr><<<ew_blocks(total), 256, 0, stream>>>(x, nm, x2m(out, out_lo), N, S, F, C, H, W,
                                                                          OH, OW, fstep, Fk, Cpad);
    return adamml_check_launch("resize_frames");
  }
  ADAMML_DISPATCH_DTYPE(dtype, T,
    (resize_frames_kernel<TIn, T*><<<ew_blocks(total), 256, 0, stream>>>(x, nm, (T*)out, N, S, F, C, H, W, OH, OW,
                                                                         fstep, Fk, Cpad)));
  return adamml_check_launch("resize_frames");
}

template <typename TIn>
static int pack_frames_s2d_any(const TIn* x, InNorm nm, void* out, void* out_lo, int N, int S, int F, int C, int H,
                               int W, int Cs, cudaStream_t stream) {
  ADAMML_REQUIRE(N > 0 && S > 0 && F > 0 && C > 0 && H > 0 && W > 0, "pack_frames_s2d: bad dims");
  ADAMML_REQUIRE(H % 2 == 0 && W % 2 == 0 && Cs % 8 == 0 && Cs >= 4 * C, "pack_frames_s2d: needs even H, W and Cs >= 4C");
  ADAMML_REQUIRE(((uintptr_t)x % (2 * sizeof(TIn))) == 0 && ((uintptr_t)out % 16) == 0,
                 "pack_frames_s2d: unaligned buffers");
  long long total = (long long)S * N * F * (H / 2) * (W / 2 + 4);
  ADAMML_REQUIRE(((uintptr_t)out_lo % 16) == 0, "pack_frames_s2d: unaligned lo plane");
  bf16* o = (bf16*)out;
  __half* ol = (__half*)out_lo;
  if (C == 3 && Cs == 16)
    pack_frames_s2d_kernel<TIn, 3, 16><<<ew_blocks(total), 256, 0, stream>>>(x, nm, o, ol, N, S, F, C, H, W, Cs);
  else
    pack_frames_s2d_kernel<TIn, 0, 0><<<ew_blocks(total), 256, 0, stream>>>(x, nm, o, ol, N, S, F, C, H, W, Cs);
  return adamml_check_launch("pack_frames_s2d");
}

extern "C" {

int adamml_pack_frames(const float* x, void* out, int N, int S, int F, int C, int H, int W, int Cpad, int dtype,
                       cudaStream_t stream) {
  return pack_frames_any<float>(x, InNorm{nullptr, nullptr}, out, nullptr, N, S, F, C, H, W, Cpad, dtype, stream);
}

int adamml_pack_frames_u8(const unsigned char* x, const float* mean, const float* stdv, void* out, int N, int S, int F,
                          int C, int H, int W, int Cpad, int dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(mean && stdv, "pack_frames_u8: needs the per-channel mean and std");
  return pack_frames_any<unsigned char>(x, InNorm{mean, stdv}, out, nullptr, N, S, F, C, H, W, Cpad, dtype, stream);
}

int adamml_resize_frames(const float* x, void* out, int N, int S, int F, int C, int H, int W, int OH, int OW,
                         int fstep, int Cpad, int dtype, cudaStream_t stream) {
  return resize_frames_any<float>(x, InNorm{nullptr, nullptr}, out, nullptr, N, S, F, C, H, W, OH, OW, fstep, Cpad,
                                  dtype, stream);
}

int adamml_resize_frames_u8(const unsigned char* x, const float* mean, const float* stdv, void* out, int N, int S,
                            int F, int C, int H, int W, int OH, int OW, int fstep, int Cpad, int dtype,
                            cudaStream_t stream) {
  ADAMML_REQUIRE(mean && stdv, "resize_frames_u8: needs the per-channel mean and std");
  return resize_frames_any<unsigned char>(x, InNorm{mean, stdv}, out, nullptr, N, S, F, C, H, W, OH, OW, fstep, Cpad,
                                          dtype, stream);
}

int adamml_pack_weight(const float* w_oihw, void* w_ohwi, int Cout, int Cin, int R, int S, int CinPad, int dtype,
                       cudaStream_t stream) {
  ADAMML_REQUIRE(Cout > 0 && Cin > 0 && R > 0 && S > 0 && CinPad >= Cin, "pack_weight: bad dims");
  long long total = (long long)Cout * R * S * CinPad;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    pack_weight_kernel<T*><<<ew_blocks(total), 256, 0, stream>>>(w_oihw, (T*)w_ohwi, Cout, Cin, R, S, CinPad));
  return adamml_check_launch("pack_weight");
}

int adamml_pack_weight_dgrad(const float* w_oihw, void* w_ihwo, int Cout, int Cin, int R, int S, int dtype,
                             cudaStream_t stream) {
  ADAMML_REQUIRE(Cout > 0 && Cin > 0 && R > 0 && S > 0, "pack_weight_dgrad: bad dims");
  long long total = (long long)Cout * R * S * Cin;
  ADAMML_DISPATCH_DTYPE(dtype, T,
    pack_weight_dgrad_kernel<T><<<ew_blocks(total), 256, 0, stream>>>(w_oihw, (T*)w_ihwo, Cout, Cin, R, S));
  return adamml_check_launch("pack_weight_dgrad");
}

int adamml_unpack_wgrad(const float* dw_ohwi, float* dw_oihw, int Cout, int Cin, int R, int S, int CinPad,
                        int accumulate, cudaStream_t stream) {
  ADAMML_REQUIRE(Cout > 0 && Cin > 0 && R > 0 && S > 0 && CinPad >= Cin, "unpack_wgrad: bad dims");
  long long total = (long long)Cout * Cin * R * S;
  unpack_wgrad_kernel<<<ew_blocks(total), 256, 0, stream>>>(dw_ohwi, dw_oihw, Cout, Cin, R, S, CinPad, accumulate);
  return adamml_check_launch("unpack_wgrad");
}

int adamml_pack_frames_s2d(const float* x, void* out, int N, int S, int F, int C, int H, int W, int Cs,
                           cudaStream_t stream) {
  return pack_frames_s2d_any<float>(x, InNorm{nullptr, nullptr}, out, nullptr, N, S, F, C, H, W, Cs, stream);
}

int adamml_pack_frames_s2d_u8(const unsigned char* x, const float* mean, const float* stdv, void* out, int N, int S,
                              int F, int C, int H, int W, int Cs, cudaStream_t stream) {
  ADAMML_REQUIRE(mean && stdv, "pack_frames_s2d_u8: needs the per-channel mean and std");
  return pack_frames_s2d_any<unsigned char>(x, InNorm{mean, stdv}, out, nullptr, N, S, F, C, H, W, Cs, stream);
}

/* ---- x2 planes: the data layer of the default precision mode writes hi (bf16) + lo (fp16 remainder) ---- */
int adamml_pack_frames_x2(const void* x, const float* mean, const float* stdv, void* out_hi, void* out_lo, int N, int S,
                          int F, int C, int H, int W, int Cpad, int is_u8, cudaStream_t stream) {
  ADAMML_REQUIRE(!is_u8 || (mean && stdv), "pack_frames_x2: uint8 input needs the per-channel mean and std");
  if (is_u8)
    return pack_frames_any<unsigned char>((const unsigned char*)x, InNorm{mean, stdv}, out_hi, out_lo, N, S, F, C, H,
                                          W, Cpad, ADAMML_X2, stream);
  return pack_frames_any<float>((const float*)x, InNorm{nullptr, nullptr}, out_hi, out_lo, N, S, F, C, H, W, Cpad,
                                ADAMML_X2, stream);
}

int adamml_resize_frames_x2(const void* x, const float* mean, const float* stdv, void* out_hi, void* out_lo, int N,
                            int S, int F, int C, int H, int W, int OH, int OW, int fstep, int Cpad, int is_u8,
                            cudaStream_t stream) {
  ADAMML_REQUIRE(!is_u8 || (mean && stdv), "resize_frames_x2: uint8 input needs the per-channel mean and std");
  if (is_u8)
    return resize_frames_any<unsigned char>((const unsigned char*)x, InNorm{mean, stdv}, out_hi, out_lo, N, S, F, C, H,
                                            W, OH, OW, fstep, Cpad, ADAMML_X2, stream);
  return resize_frames_any<float>((const float*)x, InNorm{nullptr, nullptr}, out_hi, out_lo, N, S, F, C, H, W, OH, OW,
                                  fstep, Cpad, ADAMML_X2, stream);
}

int adamml_pack_frames_s2d_x2(const void* x, const float* mean, const float* stdv, void* out_hi, void* out_lo, int N,
                              int S, int F, int C, int H, int W, int Cs, int is_u8, cudaStream_t stream) {
  ADAMML_REQUIRE(out_lo, "pack_frames_s2d_x2: needs the lo plane");
  ADAMML_REQUIRE(!is_u8 || (mean && stdv), "pack_frames_s2d_x2: uint8 input needs the per-channel mean and std");
  if (is_u8)
    return pack_frames_s2d_any<unsigned char>((const unsigned char*)x, InNorm{mean, stdv}, out_hi, out_lo, N, S, F, C,
                                              H, W, Cs, stream);
  return pack_frames_s2d_any<float>((const float*)x, InNorm{nullptr, nullptr}, out_hi, out_lo, N, S, F, C, H, W, Cs,
                                    stream);
}

/* nn.Conv2d.weight OIHW fp32 -> the FOUR OHWI weight planes of the x2 tensor-core path, plane-major in one buffer
 * w4 [4][Cout][R][S][CinPad] of 2-byte elements: b1 = bf16(w), b2 = bf16(w - b1), b3 = bf16(w - b1 - b2) (bf16 cascade,
 * exact to 24 bits) and f = fp16(w).  stem = 1: the space-to-depth first-conv operand (adamml_pack_weight_stem layout,
 * Cin = C, CinPad = Cs, S = R). */
int adamml_pack_weight_x2(const float* w_oihw, void* w4, int Cout, int Cin, int R, int S, int CinPad, int stem,
                          cudaStream_t stream) {
  ADAMML_REQUIRE(Cout > 0 && Cin > 0 && R > 0 && S > 0 && w4, "pack_weight_x2: bad arguments");
  if (stem) {
    ADAMML_REQUIRE(CinPad >= 4 * Cin && (R == 7 || R == 3), "pack_weight_x2: bad stem dims");
    const int T = (R + 1) / 2;
    const long long n = (long long)Cout * T * T * CinPad;
    pack_weight_stem_kernel<W4Ptr><<<ew_blocks(n), 256, 0, stream>>>(w_oihw, W4Ptr{(bf16*)w4, n}, Cout, Cin, CinPad, R, T);
  } else {
    ADAMML_REQUIRE(CinPad >= Cin, "pack_weight_x2: CinPad < Cin");
    const long long n = (long long)Cout * R * S * CinPad;
    pack_weight_kernel<W4Ptr><<<ew_blocks(n), 256, 0, stream>>>(w_oihw, W4Ptr{(bf16*)w4, n}, Cout, Cin, R, S, CinPad);
  }
  return adamml_check_launch("pack_weight_x2");
}

int adamml_nhwc_to_s2d(const void* x, void* out, long long IMGS, int C, int H, int W, int Cs, int padl, int padr,
                       cudaStream_t stream) {
  ADAMML_REQUIRE(IMGS > 0 && C > 0 && H > 0 && W > 0 && padl >= 0 && padr >= 0, "nhwc_to_s2d: bad dims");
  ADAMML_REQUIRE(H % 2 == 0 && W % 2 == 0 && Cs >= 4 * C, "nhwc_to_s2d: needs even H, W and Cs >= 4C");
  long long total = IMGS * (H / 2) * (W / 2 + padl + padr);
  nhwc_to_s2d_kernel<<<ew_blocks(total), 256, 0, stream>>>((const bf16*)x, (bf16*)out, IMGS, C, H, W, Cs, padl, padr,
                                                           adamml_live_limit(IMGS));
  return adamml_check_launch("nhwc_to_s2d");
}

int adamml_pack_weight_stem(const float* w_oihw, void* w_packed, int Cout, int C, int Cs, int R,
                            cudaStream_t stream) {
  ADAMML_REQUIRE(Cout > 0 && C > 0 && Cs >= 4 * C && (R == 7 || R == 3), "pack_weight_stem: bad dims");
  const int T = (R + 1) / 2;
  pack_weight_stem_kernel<bf16*><<<ew_blocks((long long)Cout * T * T * Cs), 256, 0, stream>>>(
      w_oihw, (bf16*)w_packed, Cout, C, Cs, R, T);
  return adamml_check_launch("pack_weight_stem");
}

int adamml_unpack_wgrad_stem(const float* dw_packed, float* dw_oihw, int Cout, int C, int Cs, int R,
                             cudaStream_t stream) {
  ADAMML_REQUIRE(Cout > 0 && C > 0 && Cs >= 4 * C && (R == 7 || R == 3), "unpack_wgrad_stem: bad dims");
  const int T = (R + 1) / 2;
  unpack_wgrad_stem_kernel<<<ew_blocks((long long)Cout * C * R * R), 256, 0, stream>>>(dw_packed, dw_oihw, Cout, C, Cs,
                                                                                        R, T);
  return adamml_check_launch("unpack_wgrad_stem");
}

int adamml_pack_chunk(void) { return PACK_CHUNK; }

int adamml_pack_weights_multi(const long long* jobs, const int* chunk_job, const long long* chunk_start, int n_jobs,
                              int n_chunks, cudaStream_t stream) {
  ADAMML_REQUIRE(jobs && chunk_job && chunk_start && n_jobs > 0 && n_chunks > 0, "pack_weights_multi: bad arguments");
  pack_weights_multi_kernel<<<n_chunks, 256, 0, stream>>>(jobs, chunk_job, chunk_start);
  return adamml_check_launch("pack_weights_multi");
}

int adamml_pack_weight_dw(const float* w_c33, float* w_9c, int C, cudaStream_t stream) {
  ADAMML_REQUIRE(C > 0, "pack_weight_dw: bad dims");
  transpose_dw_kernel<<<ceil_div(C * 9, 256), 256, 0, stream>>>(w_c33, w_9c, C, 1);
  return adamml_check_launch("pack_weight_dw");
}

int adamml_unpack_wgrad_dw(const float* dw_9c, float* dw_c33, int C, cudaStream_t stream) {
  ADAMML_REQUIRE(C > 0, "unpack_wgrad_dw: bad dims");
  transpose_dw_kernel<<<ceil_div(C * 9, 256), 256, 0, stream>>>(dw_9c, dw_c33, C, 0);
  return adamml_check_launch("unpack_wgrad_dw");
}

// dtype codes for src/dst
int adamml_cast(const void* src, void* dst, long long total, int src_dtype, int dst_dtype, cudaStream_t stream) {
  ADAMML_REQUIRE(total >= 0, "cast: negative size");
  if (total == 0) return ADAMML_OK;
  if (src_dtype == ADAMML_F32 && dst_dtype == ADAMML_BF16)
    cast_kernel<float, bf16><<<ew_blocks(total), 256, 0, stream>>>((const float*)src, (bf16*)dst, total);
  else if (src_dtype == ADAMML_BF16 && dst_dtype == ADAMML_F32)
    cast_kernel<bf16, float><<<ew_blocks(total), 256, 0, stream>>>((const bf16*)src, (float*)dst, total);
  else if (src_dtype == ADAMML_F32 && dst_dtype == ADAMML_F32)
    cast_kernel<float, float><<<ew_blocks(total), 256, 0, stream>>>((const float*)src, (float*)dst, total);
  else {
    adamml_set_error("cast: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
    return ADAMML_ERR_ARG;
  }
  return adamml_check_launch("cast");
}

}  // extern "C"
