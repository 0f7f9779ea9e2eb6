// CUDA-core kernels of the UNet path:
//   * fold_pack_kernel      eval-mode BatchNorm folded into [tap][Cin][Cout] fp32 weights + bias
//   * conv1_kernel          U1: Conv2d(1->16,5x5,s2,p2)+BN+LeakyReLU straight from the caller's patch
//                           view (any strides, ragged frame counts) into the NHWC concat buffer
//   * conv_direct_kernel    generic 5x5 stride-2 conv / transposed conv, fp32 FMA.  This is the
//                           exact-arithmetic SVS_PRECISION_FP32 path and the on-device cross-check
//                           for the tcgen05 kernels; it is not a fallback for them.
//   * deconv6_kernel        D6 + sigmoid + (1-m) + mask x mixture fused, writes the caller's view
//   * read_activation_kernel  NHWC slice -> NCHW fp32 (parity hook)
//
// Reference: model.py:47-109 (layer definitions), model.py:176-200 (forward), inference.py:102,107.
#include "unet_internal.cuh"

namespace svs {

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_LEAKY) return v > 0.0f ? v : 0.2f * v;      // LeakyReLU(0.2), model.py:50
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

// 8 consecutive channels -> fp32 registers
__device__ __forceinline__ void load8(const float* p, float (&x)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&x)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[2 * i] = __uint_as_float(u[i] << 16);
    x[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

__device__ __forceinline__ void store16v(float* p, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void store16v(__nv_bfloat16* p, const float (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// ---------------------------------------------------------------------------------------------
// BatchNorm fold + repack.  conv: w[co][ci][kh][kw], deconv: w[ci][co][kh][kw]  ->  [tap][ci][co]
// w' = w * gamma/sqrt(var+eps) along Cout, b' = (b - mean) * gamma/sqrt(var+eps) + beta.
__global__ void fold_pack_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                 const float* __restrict__ g, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ var,
                                 int cin, int cout, int transposed, float* __restrict__ w_out,
                                 float* __restrict__ b_out) {
  const int total = 25 * cin * cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cout;
    const int ci = (i / cout) % cin;
    const int tap = i / (cout * cin);
    const float s = g ? g[co] / sqrtf(var[co] + 1e-5f) : 1.0f;
    const size_t src = transposed ? (static_cast<size_t>(ci) * cout + co) * 25 + tap
                                  : (static_cast<size_t>(co) * cin + ci) * 25 + tap;
    w_out[i] = w[src] * s;
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < cout; co += gridDim.x * blockDim.x) {
    if (g) {
      const float s = g[co] / sqrtf(var[co] + 1e-5f);
      b_out[co] = (b[co] - mean[co]) * s + beta[co];
    } else {
      b_out[co] = b[co];
    }
  }
}

int launch_fold_pack(const svs_conv_params& p, int cin, int cout, bool transposed, float* w_out,
                     float* b_out, cudaStream_t st) {
  const int total = 25 * cin * cout;
  const int blocks = (total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256;
  fold_pack_kernel<<<blocks, 256, 0, st>>>(p.weight, p.bias, p.bn_weight, p.bn_bias, p.bn_mean, p.bn_var,
                                           cin, cout, transposed ? 1 : 0, w_out, b_out);
  SVS_CHECK_LAUNCH("fold_pack_kernel");
  return SVS_OK;
}

// ---------------------------------------------------------------------------------------------
// U1.  CTA = 8 output rows x all 64 output columns of one patch; the 19 x 128 input window is staged
// once in shared memory (zero filled for the conv padding and for frames >= the patch's valid count),
// each thread then computes 1 row x 4 columns x 16 channels from registers (1600 FMA per 20 vector
// shared loads of input + 100 broadcast loads of weights).
constexpr int kC1Rows = 8;                       // output rows per CTA
constexpr int kC1Threads = 16 * kC1Rows;         // one thread = 1 row x 4 columns x 16 channels
constexpr int kC1InRows = 2 * kC1Rows + 3;       // 35 input rows
constexpr int kC1Pitch = 136;                    // 4 left pad + 128 + 4 right pad floats

template <typename T>
__global__ void __launch_bounds__(kC1Threads)
conv1_kernel(const float* __restrict__ in, const int64_t* __restrict__ patch_off, int64_t stride_b,
             int64_t stride_f, int64_t stride_t, const int32_t* __restrict__ in_frames,
             const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
             int out_pitch, int out_coff, int act) {
  __shared__ __align__(16) float sw[25 * 16];
  __shared__ __align__(16) float sb[16];
  __shared__ __align__(16) float tile[kC1InRows * kC1Pitch];
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < 400; i += kC1Threads) sw[i] = w[i];
  if (threadIdx.x < 16) sb[threadIdx.x] = bias[threadIdx.x];
  const int b = blockIdx.y;
  const int oh0 = blockIdx.x * kC1Rows;
  const float* __restrict__ src = in + (patch_off ? patch_off[b] : b * stride_b);
  const int nf = in_frames ? in_frames[b] : SVS_PATCH_FRAMES;
  const int f0 = 2 * oh0 - 2;
  // tile[r][c] = x[f0 + r][c - 4]
  const bool vec_ok = stride_t == 1 && (stride_f & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  if (vec_ok) {                                  // dense patches: 16-byte loads, all in flight at once
    for (int i = threadIdx.x; i < kC1InRows * (kC1Pitch / 4); i += kC1Threads) {
      const int r = i / (kC1Pitch / 4), q = i - r * (kC1Pitch / 4);
      const int f = f0 + r, t = 4 * (q - 1);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f >= 0 && f < SVS_PATCH_BINS && t >= 0 && t < SVS_PATCH_FRAMES) {
        v = __ldg(reinterpret_cast<const float4*>(src + f * stride_f + t));
        if (t + 0 >= nf) v.x = 0.f;
        if (t + 1 >= nf) v.y = 0.f;
        if (t + 2 >= nf) v.z = 0.f;
        if (t + 3 >= nf) v.w = 0.f;
      }
      *reinterpret_cast<float4*>(&tile[r * kC1Pitch + 4 * q]) = v;
    }
  } else if (stride_t == 1) {
    for (int i = threadIdx.x; i < kC1InRows * kC1Pitch; i += kC1Threads) {
      const int r = i / kC1Pitch, c = i - r * kC1Pitch;
      const int f = f0 + r, t = c - 4;
      tile[i] = (f >= 0 && f < SVS_PATCH_BINS && t >= 0 && t < nf) ? __ldg(src + f * stride_f + t) : 0.0f;
    }
  } else {                                       // frequency-contiguous views: walk r fastest
    for (int i = threadIdx.x; i < kC1InRows * kC1Pitch; i += kC1Threads) {
      const int c = i / kC1InRows, r = i - c * kC1InRows;
      const int f = f0 + r, t = c - 4;
      tile[r * kC1Pitch + c] =
          (f >= 0 && f < SVS_PATCH_BINS && t >= 0 && t < nf) ? __ldg(src + f * stride_f + t * stride_t) : 0.0f;
    }
  }
  __syncthreads();
  const int g = threadIdx.x & 15;                // column group: output columns 4g .. 4g+3
  const int ly = threadIdx.x >> 4;               // output row within the CTA
  float acc[4][16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[p][c] = sb[c];
#pragma unroll
  for (int kh = 0; kh < 5; ++kh) {
    // input columns t = 8g-2 .. 8g+8  ->  tile columns 8g+2 .. 8g+12 ; load tile columns 8g .. 8g+12
    const float* row = tile + (2 * ly + kh) * kC1Pitch + 8 * g;
    float x[13];
    const float4 a0 = *reinterpret_cast<const float4*>(row);
    const float4 a1 = *reinterpret_cast<const float4*>(row + 4);
    const float4 a2 = *reinterpret_cast<const float4*>(row + 8);
    x[0] = a0.x; x[1] = a0.y; x[2] = a0.z; x[3] = a0.w; x[4] = a1.x; x[5] = a1.y; x[6] = a1.z; x[7] = a1.w;
    x[8] = a2.x; x[9] = a2.y; x[10] = a2.z; x[11] = a2.w; x[12] = row[12];
#pragma unroll
    for (int kw = 0; kw < 5; ++kw) {
      const float4* wt = reinterpret_cast<const float4*>(sw + (kh * 5 + kw) * 16);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 wv = wt[c4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float xv = x[2 * p + kw + 2];     // tile column 8g + 2p + kw + 2  <->  t = 2(4g+p) + kw - 2
          acc[p][4 * c4 + 0] = fmaf(xv, wv.x, acc[p][4 * c4 + 0]);
          acc[p][4 * c4 + 1] = fmaf(xv, wv.y, acc[p][4 * c4 + 1]);
          acc[p][4 * c4 + 2] = fmaf(xv, wv.z, acc[p][4 * c4 + 2]);
          acc[p][4 * c4 + 3] = fmaf(xv, wv.w, acc[p][4 * c4 + 3]);
        }
      }
    }
  }
  const int oh = oh0 + ly;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    T* dst = out + ((static_cast<size_t>(b) * 256 + oh) * 64 + 4 * g + p) * out_pitch + out_coff;
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = apply_act(acc[p][c], act);
    store16v(dst, v);                            // whole 32-byte sectors: no partial-sector writes in L2
  }
}

// ---------------------------------------------------------------------------------------------
// Generic direct 5x5 stride-2 convolution / transposed convolution over NHWC buffers.
// One thread = one output position x 4 output channels x kNB batch samples (b, b + B/kNB, ...); fp32 accumulate.
// The samples share the tap validity and the weights, so each weight float4 is loaded once per kNB x 4 FMAs: the
// kernel is bound by those loads, not by the FMAs.  Per output the accumulation order does not depend on kNB.
template <typename T, bool kTransposed, int kNB>
__global__ void __launch_bounds__(256)
conv_direct_kernel(const T* __restrict__ in, int in_pitch, int in_coff, int hin, int win, int cin,
                   const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
                   int out_pitch, int out_coff, int hout, int wout, int cout, int act, int batch,
                   int accumulate) {
  const int cg_n = cout >> 2;
  const int bgroups = batch / kNB;
  const size_t total = static_cast<size_t>(bgroups) * hout * wout * cg_n;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % cg_n);
  size_t pix = idx / cg_n;
  const int ow = static_cast<int>(pix % wout); pix /= wout;
  const int oh = static_cast<int>(pix % hout);
  const int b0 = static_cast<int>(pix / hout);
  const int co = cg * 4;
  const size_t in_bstride = static_cast<size_t>(bgroups) * hin * win * in_pitch;
  float acc[kNB][4];
#pragma unroll
  for (int n = 0; n < kNB; ++n)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[n][u] = bias ? bias[co + u] : 0.f;
  for (int kh = 0; kh < 5; ++kh) {
    int ih;
    if (kTransposed) {                     // oh = 2 ih - 2 + kh
      const int th = oh + 2 - kh;
      if (th & 1) continue;
      ih = th >> 1;
    } else {
      ih = 2 * oh + kh - 2;
    }
    if (ih < 0 || ih >= hin) continue;
    for (int kw = 0; kw < 5; ++kw) {
      int iw;
      if (kTransposed) {
        const int tw = ow + 2 - kw;
        if (tw & 1) continue;
        iw = tw >> 1;
      } else {
        iw = 2 * ow + kw - 2;
      }
      if (iw < 0 || iw >= win) continue;
      const T* __restrict__ px = in + ((static_cast<size_t>(b0) * hin + ih) * win + iw) * in_pitch + in_coff;
      const float* __restrict__ wt = w + static_cast<size_t>(kh * 5 + kw) * cin * cout + co;
      for (int ci = 0; ci < cin; ci += 8) {
        float x[kNB][8];
#pragma unroll
        for (int n = 0; n < kNB; ++n) load8(px + n * in_bstride + ci, x[n]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wt + static_cast<size_t>(ci + u) * cout));
#pragma unroll
          for (int n = 0; n < kNB; ++n) {
            acc[n][0] = fmaf(x[n][u], wv.x, acc[n][0]);
            acc[n][1] = fmaf(x[n][u], wv.y, acc[n][1]);
            acc[n][2] = fmaf(x[n][u], wv.z, acc[n][2]);
            acc[n][3] = fmaf(x[n][u], wv.w, acc[n][3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int n = 0; n < kNB; ++n) {
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[n][u] = apply_act(acc[n][u], act);
    T* dst = out + ((static_cast<size_t>(b0 + n * bgroups) * hout + oh) * wout + ow) * out_pitch + out_coff + co;
    if (accumulate) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[n][u] += to_float(dst[u]);
    }
    store4(dst, acc[n]);
  }
}

template <typename T, bool kTransposed>
static void launch_conv_direct_any(const T* in, int in_pitch, int in_coff, int hin, int win, int cin, const float* w,
                                   const float* bias, T* out, int out_pitch, int out_coff, int hout, int wout,
                                   int cout, int act, int batch, int accumulate, cudaStream_t st) {
  const int nb = batch % 4 == 0 ? 4 : (batch % 2 == 0 ? 2 : 1);
  const size_t total = static_cast<size_t>(batch / nb) * hout * wout * (cout / 4);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
#define SVS_CD(NB)                                                                                                   \
  conv_direct_kernel<T, kTransposed, NB><<<blocks, 256, 0, st>>>(in, in_pitch, in_coff, hin, win, cin, w, bias, out, \
                                                                 out_pitch, out_coff, hout, wout, cout, act, batch,   \
                                                                 accumulate)
  if (nb == 4) SVS_CD(4);
  else if (nb == 2) SVS_CD(2);
  else SVS_CD(1);
#undef SVS_CD
}

// ---------------------------------------------------------------------------------------------
// D6 + sigmoid + mask application.  One thread = one output (bin, frame) element.
template <typename T>
__global__ void __launch_bounds__(256)
deconv6_kernel(const T* __restrict__ in /*cat1 [B][256][64][32]*/, const float* __restrict__ w /*[25][32]*/,
               const float* __restrict__ bias, const float* __restrict__ mix, const int64_t* __restrict__ mix_off,
               int64_t mix_sb, int64_t mix_sf, int64_t mix_st, float* __restrict__ out,
               const int64_t* __restrict__ out_off, int64_t out_sb, int64_t out_sf, int64_t out_st,
               const int32_t* __restrict__ in_frames, int flags) {
  __shared__ float sw[25 * 32];
  for (int i = threadIdx.x; i < 800; i += 256) sw[i] = w[i];
  __syncthreads();
  const int b = blockIdx.z;
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  int oh, ow;
  if (out_st == 1) { ow = blockIdx.x * 16 + lx; oh = blockIdx.y * 16 + ly; }
  else             { oh = blockIdx.y * 16 + lx; ow = blockIdx.x * 16 + ly; }
  const int nf = in_frames ? in_frames[b] : SVS_PATCH_FRAMES;
  if (ow >= nf) return;                                       // cropped padding, inference.py:113-114
  float acc = bias[0];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int kh = (oh & 1) + 2 * a;                          // kh = oh (mod 2)
    if (kh > 4) continue;
    const int ih = (oh + 2 - kh) >> 1;
    if (ih < 0 || ih >= 256) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int kw = (ow & 1) + 2 * c;
      if (kw > 4) continue;
      const int iw = (ow + 2 - kw) >> 1;
      if (iw < 0 || iw >= 64) continue;
      const T* __restrict__ px = in + ((static_cast<size_t>(b) * 256 + ih) * 64 + iw) * 32;
      const float* wt = sw + (kh * 5 + kw) * 32;
#pragma unroll
      for (int ci = 0; ci < 32; ci += 8) {
        float x[8];
        load8(px + ci, x);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = fmaf(x[u], wt[ci + u], acc);
      }
    }
  }
  float m = 1.0f / (1.0f + expf(-acc));                       // torch.sigmoid, model.py:200
  if (flags & SVS_FLAG_INVERT) m = 1.0f - m;                  // inference.py:102
  if (flags & SVS_FLAG_APPLY_MASK) {
    const float x = __ldg(mix + (mix_off ? mix_off[b] : b * mix_sb) + oh * mix_sf + ow * mix_st);
    m = x * m;                                                // inference.py:107
  }
  out[(out_off ? out_off[b] : b * out_sb) + oh * out_sf + ow * out_st] = m;
}

template <typename T>
__global__ void read_activation_kernel(const T* __restrict__ in, int pitch, int coff, int h, int w, int c,
                                       int batch, float* __restrict__ out) {
  const size_t total = static_cast<size_t>(batch) * c * h * w;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w);
    const int y = static_cast<int>((i / w) % h);
    const int ch = static_cast<int>((i / (static_cast<size_t>(w) * h)) % c);
    const int b = static_cast<int>(i / (static_cast<size_t>(w) * h * c));
    out[i] = to_float(in[((static_cast<size_t>(b) * h + y) * w + x) * pitch + coff + ch]);
  }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static int launch_layer_direct_t(const svs_unet_plan* plan, int li, const Workspace& ws,
                                 const svs_patch_view* in, const svs_patch_view* out,
                                 const int32_t* in_frames, int batch, int flags, cudaStream_t st) {
  const LayerGeom& g = kLayers[li];
  if (li == 0) {
    dim3 grid(256 / kC1Rows, batch);
    conv1_kernel<T><<<grid, kC1Threads, 0, st>>>(in->base, in->patch_off, in->stride_b, in->stride_f, in->stride_t,
                                         in_frames, plan->w_fold[0], plan->b_fold[0],
                                         reinterpret_cast<T*>(ws.buf[g.out_buf]), kBufGeom[g.out_buf].c,
                                         g.out_coff, ACT_LEAKY);
    SVS_CHECK_LAUNCH("conv1_kernel");
    return SVS_OK;
  }
  if (li == 11) {
    dim3 grid(128 / 16, 512 / 16, batch);
    deconv6_kernel<T><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(ws.buf[BUF_CAT1]), plan->w_fold[11],
                                           plan->b_fold[11], in->base, in->patch_off, in->stride_b,
                                           in->stride_f, in->stride_t, out->base, out->patch_off,
                                           out->stride_b, out->stride_f, out->stride_t, in_frames, flags);
    SVS_CHECK_LAUNCH("deconv6_kernel");
    return SVS_OK;
  }
  const T* src = reinterpret_cast<const T*>(ws.buf[g.in_buf]);
  T* dst = reinterpret_cast<T*>(ws.buf[g.out_buf]);
  if (g.transposed)
    launch_conv_direct_any<T, true>(src, kBufGeom[g.in_buf].c, g.in_coff, g.hin, g.win, g.cin, plan->w_fold[li],
                                    plan->b_fold[li], dst, kBufGeom[g.out_buf].c, g.out_coff, g.hout, g.wout, g.cout,
                                    g.act, batch, 0, st);
  else
    launch_conv_direct_any<T, false>(src, kBufGeom[g.in_buf].c, g.in_coff, g.hin, g.win, g.cin, plan->w_fold[li],
                                     plan->b_fold[li], dst, kBufGeom[g.out_buf].c, g.out_coff, g.hout, g.wout, g.cout,
                                     g.act, batch, 0, st);
  SVS_CHECK_LAUNCH("conv_direct_kernel");
  return SVS_OK;
}

// fp32 launchers used by the training step (train.cu)
int launch_conv_direct_f32(const float* in, int in_pitch, int in_coff, int hin, int win, int cin, const float* w,
                           const float* bias, float* out, int out_pitch, int out_coff, int hout, int wout,
                           int cout, int act, bool transposed, int batch, bool accumulate, cudaStream_t st) {
  if (transposed)
    launch_conv_direct_any<float, true>(in, in_pitch, in_coff, hin, win, cin, w, bias, out, out_pitch, out_coff, hout,
                                        wout, cout, act, batch, accumulate ? 1 : 0, st);
  else
    launch_conv_direct_any<float, false>(in, in_pitch, in_coff, hin, win, cin, w, bias, out, out_pitch, out_coff, hout,
                                         wout, cout, act, batch, accumulate ? 1 : 0, st);
  SVS_CHECK_LAUNCH("conv_direct_kernel");
  return SVS_OK;
}

int launch_conv1_f32(const float* mix, const float* w, const float* bias, float* out, int batch, cudaStream_t st) {
  dim3 grid(256 / kC1Rows, batch);
  conv1_kernel<float><<<grid, kC1Threads, 0, st>>>(mix, nullptr, 512 * 128, 128, 1, nullptr, w, bias, out, 16, 0,
                                                  ACT_NONE);
  SVS_CHECK_LAUNCH("conv1_kernel");
  return SVS_OK;
}

int launch_deconv6_f32(const float* cat1, const float* w, const float* bias, float* mask, int batch,
                       cudaStream_t st) {
  dim3 grid(128 / 16, 512 / 16, batch);
  deconv6_kernel<float><<<grid, 256, 0, st>>>(cat1, w, bias, mask, nullptr, 512 * 128, 128, 1, mask, nullptr,
                                             512 * 128, 128, 1, nullptr, 0);
  SVS_CHECK_LAUNCH("deconv6_kernel");
  return SVS_OK;
}

int launch_layer_direct(const svs_unet_plan* plan, int li, const Workspace& ws, const svs_patch_view* in,
                        const svs_patch_view* out, const int32_t* in_frames, int batch, int flags,
                        cudaStream_t st) {
  if (plan->elem_size == 2)
    return launch_layer_direct_t<__nv_bfloat16>(plan, li, ws, in, out, in_frames, batch, flags, st);
  return launch_layer_direct_t<float>(plan, li, ws, in, out, in_frames, batch, flags, st);
}

int launch_read_activation(const svs_unet_plan* plan, int layer, int batch, const Workspace& ws,
                           float* out_nchw, cudaStream_t st) {
  const LayerGeom& g = kLayers[layer];
  const size_t total = static_cast<size_t>(batch) * g.cout * g.hout * g.wout;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256 > 4736 ? 4736 : (total + 255) / 256);
  if (plan->elem_size == 2) {
    read_activation_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(ws.buf[g.out_buf]), kBufGeom[g.out_buf].c, g.out_coff, g.hout,
        g.wout, g.cout, batch, out_nchw);
  } else {
    read_activation_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(ws.buf[g.out_buf]),
                                                         kBufGeom[g.out_buf].c, g.out_coff, g.hout, g.wout,
                                                         g.cout, batch, out_nchw);
  }
  SVS_CHECK_LAUNCH("read_activation_kernel");
  return SVS_OK;
}

}  // namespace svs
