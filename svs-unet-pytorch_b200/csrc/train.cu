// T1: training step of the UNet — train-mode forward, masked-L1 loss, backward (svs_b200.h).
//
// Replaces the autograd graph of reference train.py:274-299 / model.py:203-220.  Activations are
// fp32 NHWC in the same concat-buffer layout as the inference path; every reduction (BatchNorm
// statistics, bias / BatchNorm gradients, weight gradients, the loss) is two-stage with a fixed order,
// so a step is bit-reproducible (no atomics).
//
// Two arithmetic modes, selected by the `plan` argument of the C ABI:
//   plan != NULL (svs_unet_train_plan_create): TF32 on tcgen05.  Forward and data-gradient convolutions are the
//     implicit-GEMM kernels of conv_tc.cu / conv_tc_cluster.cu (kind::tf32, fp32 accumulate; dgrad of a stride-2 conv
//     is the transposed-conv kernel with channel-transposed weights and vice versa), the weight gradient is the
//     MN-major pixel-reduction GEMM of wgrad_tc.cu.  Operands are rounded to TF32 where they are produced
//     (BatchNorm apply / BatchNorm backward); statistics, normalisation, loss and all reductions stay fp32.
//   plan == NULL: exact fp32 on the CUDA-core kernels of conv_direct.cu (parity mode).
// conv1 (Cin = 1) and deconv6 (Cout = 1) are memory-bound edge layers with CUDA-core kernels in both modes.
#include "unet_internal.cuh"
#include <cstdlib>
#include "tc_ptx.cuh"

#include <new>

namespace svs {

int launch_conv_direct_f32(const float* in, int in_pitch, int in_coff, int hin, int win, int cin, const float* w,
                           const float* bias, float* out, int out_pitch, int out_coff, int hout, int wout,
                           int cout, int act, bool transposed, int batch, bool accumulate, cudaStream_t st);
int launch_conv1_f32(const float* mix, const float* w, const float* bias, float* out, int batch, cudaStream_t st);
int launch_deconv6_f32(const float* cat1, const float* w, const float* bias, float* mask, int batch,
                       cudaStream_t st);

constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;
constexpr int kRedSplits = 64;                   // stage-1 partial sums per channel (fp32 parity mode)
constexpr int kRedBlocksMax = 1184;              // stage-1 blocks of the row-streaming reduction (8 per SM)

// conv_tc.cu / wgrad_tc.cu
int tc_plan_one(TcLayer& t, const ConvDesc& g, const float* w_fold, bool tf32, cudaStream_t st);
int tc_pack_one(TcLayer& t, const float* w_fold, bool tf32, cudaStream_t st);
void tc_free_one(TcLayer& t);
int tc_launch(const TcLayer& t, const TcIo& io, int batch, bool tf32, cudaStream_t st);
size_t tc_splitk_bytes_one(const TcLayer& t, int batch);
int d6_pack(const float* w_fold, bool tf32, void* d6_weights, cudaStream_t st);
int d6_make_weight_map(void* d6_weights, bool tf32, CUtensorMap* out);
int d6_launch_raw(bool tf32, const CUtensorMap& tmap_w, const float* bias, const void* cat1, const svs_patch_view* in,
                  const svs_patch_view* out, const int32_t* in_frames, int batch, int flags, cudaStream_t st);
bool wgrad_tc_supported(int gh, int gw, int s_c, int l_c, int s_pitch, int l_pitch, int s_coff, int l_coff);
size_t wgrad_tc_partial_floats(int gh, int gw, int batch, int s_c, int l_c);
int wgrad_tc_launch(const float* S, int s_pitch, int s_coff, int s_c, const float* L, int l_pitch, int l_coff, int l_c,
                    int gh, int gw, int batch, float* partial, size_t partial_floats, float* grad_w, cudaStream_t st);

// torch layout -> [tap][ci][co]  (conv: (co,ci,kh,kw); deconv: (ci,co,kh,kw)) and its channel transpose
__global__ void train_pack_kernel(const float* __restrict__ w, int cin, int cout, int transposed,
                                  float* __restrict__ w_fwd /*[tap][ci][co]*/, float* __restrict__ w_t /*[tap][co][ci]*/) {
  const int total = 25 * cin * cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cout, ci = (i / cout) % cin, tap = i / (cout * cin);
    const size_t src = transposed ? (static_cast<size_t>(ci) * cout + co) * 25 + tap
                                  : (static_cast<size_t>(co) * cin + ci) * 25 + tap;
    const float v = w[src];
    w_fwd[i] = v;
    w_t[(static_cast<size_t>(tap) * cout + co) * cin + ci] = v;
  }
}

// ---- per-channel reductions over an [N][C] fp32 slab (pitch / channel offset aware) ------------
// mode 0: (sum z, sum z^2)                      BatchNorm statistics
// mode 1: (sum g, sum g * xhat)  g = dy_eff     BatchNorm backward; dy_eff applies dropout and act'
struct RedArgs {
  const float* z;          // [N][C] pre-BN conv output
  const float* y;          // activation buffer (post act/dropout), pitch/coff
  const float* dy;         // gradient w.r.t. y, same pitch/coff
  const uint8_t* keep;     // [B][C] or null
  const float* mean; const float* invstd;
  int y_pitch, y_coff, C, act, pix_per_sample;
  size_t N;
};

__device__ __forceinline__ float dy_effective(const RedArgs& a, size_t n, int c) {
  const float yv = a.y[n * a.y_pitch + a.y_coff + c];
  float g = a.dy[n * a.y_pitch + a.y_coff + c];
  if (a.keep) {
    const int b = static_cast<int>(n / a.pix_per_sample);
    g = a.keep[b * a.C + c] ? 2.0f * g : 0.0f;                 // Dropout2d(0.5): scale 1/(1-p)
  }
  if (a.act == ACT_RELU) g = yv > 0.0f ? g : 0.0f;
  else if (a.act == ACT_LEAKY) g = yv > 0.0f ? g : 0.2f * g;
  return g;
}

template <int kMode>
__global__ void __launch_bounds__(256)
channel_reduce_kernel(RedArgs a, float* __restrict__ partial /*[splits][2][C]*/) {
  __shared__ float s0[8][32], s1[8][32];
  const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const size_t per = (a.N + gridDim.y - 1) / gridDim.y;
  const size_t n0 = blockIdx.y * per, n1 = min(a.N, n0 + per);
  float acc0 = 0.f, acc1 = 0.f;
  if (c < a.C) {
    for (size_t n = n0 + ty; n < n1; n += 8) {
      if (kMode == 0) {
        const float v = a.z[n * a.C + c];
        acc0 += v; acc1 += v * v;
      } else {
        const float g = dy_effective(a, n, c);
        const float xh = (a.z[n * a.C + c] - a.mean[c]) * a.invstd[c];
        acc0 += g; acc1 += g * xh;
      }
    }
  }
  s0[ty][cx] = acc0; s1[ty][cx] = acc1;
  __syncthreads();
  if (ty == 0 && c < a.C) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += s0[i][cx]; t1 += s1[i][cx]; }
    partial[(static_cast<size_t>(blockIdx.y) * 2 + 0) * a.C + c] = t0;
    partial[(static_cast<size_t>(blockIdx.y) * 2 + 1) * a.C + c] = t1;
  }
}

// Row-streaming form of the same two reductions: a CTA walks a contiguous run of pixels, thread = (4-channel group,
// row lane), 16-byte loads, so a warp reads whole rows and the pass runs at HBM speed (the column-strided kernel
// above keeps 64 CTAs busy and costs 0.2 - 0.6 ms per layer at batch 64).  Row lanes are added in fixed order and
// every CTA writes one partial -> deterministic.  Requires C % 4 == 0, C <= 1024, pitches / offsets % 4 == 0.
template <int kMode>
__global__ void __launch_bounds__(256)
channel_reduce_rows_kernel(RedArgs a, int pix_log2, float* __restrict__ partial /*[gridDim.x][2][C]*/) {
  __shared__ float4 s0[256], s1[256];
  const int cgs = a.C >> 2;                                   // 4-channel groups per row (<= 256)
  const int rows = 256 / cgs;                                 // rows per iteration
  const int cg = threadIdx.x % cgs, rl = threadIdx.x / cgs;
  size_t per = (a.N + gridDim.x - 1) / gridDim.x;
  per = (per + rows - 1) / rows * rows;
  const size_t n0 = blockIdx.x * per, n1 = min(a.N, n0 + per);
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
  float4 mean4 = acc0, istd4 = acc0;
  if (kMode == 1 && rl < rows) {
    mean4 = *reinterpret_cast<const float4*>(a.mean + 4 * cg);
    istd4 = *reinterpret_cast<const float4*>(a.invstd + 4 * cg);
  }
  if (rl < rows) {
    for (size_t n = n0 + rl; n < n1; n += rows) {
      const float4 z = *reinterpret_cast<const float4*>(a.z + n * a.C + 4 * cg);
      if (kMode == 0) {
        acc0.x += z.x; acc0.y += z.y; acc0.z += z.z; acc0.w += z.w;
        acc1.x += z.x * z.x; acc1.y += z.y * z.y; acc1.z += z.z * z.z; acc1.w += z.w * z.w;
      } else {
        const size_t o = n * a.y_pitch + a.y_coff + 4 * cg;
        const float4 yv = *reinterpret_cast<const float4*>(a.y + o);
        float4 g = *reinterpret_cast<const float4*>(a.dy + o);
        if (a.keep) {
          const uchar4 k = *reinterpret_cast<const uchar4*>(a.keep + (n >> pix_log2) * a.C + 4 * cg);
          g.x = k.x ? 2.0f * g.x : 0.0f; g.y = k.y ? 2.0f * g.y : 0.0f;
          g.z = k.z ? 2.0f * g.z : 0.0f; g.w = k.w ? 2.0f * g.w : 0.0f;
        }
        const float neg = a.act == ACT_LEAKY ? 0.2f : 0.0f;   // ACT_RELU: 0
        g.x = yv.x > 0.f ? g.x : neg * g.x; g.y = yv.y > 0.f ? g.y : neg * g.y;
        g.z = yv.z > 0.f ? g.z : neg * g.z; g.w = yv.w > 0.f ? g.w : neg * g.w;
        acc0.x += g.x; acc0.y += g.y; acc0.z += g.z; acc0.w += g.w;
        acc1.x += g.x * ((z.x - mean4.x) * istd4.x); acc1.y += g.y * ((z.y - mean4.y) * istd4.y);
        acc1.z += g.z * ((z.z - mean4.z) * istd4.z); acc1.w += g.w * ((z.w - mean4.w) * istd4.w);
      }
    }
  }
  s0[threadIdx.x] = acc0; s1[threadIdx.x] = acc1;
  __syncthreads();
  if (threadIdx.x < cgs) {
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
    for (int r = 0; r < rows; ++r) {                          // fixed order
      const float4 u = s0[r * cgs + cg], v = s1[r * cgs + cg];
      t0.x += u.x; t0.y += u.y; t0.z += u.z; t0.w += u.w;
      t1.x += v.x; t1.y += v.y; t1.z += v.z; t1.w += v.w;
    }
    *reinterpret_cast<float4*>(partial + (static_cast<size_t>(blockIdx.x) * 2 + 0) * a.C + 4 * cg) = t0;
    *reinterpret_cast<float4*>(partial + (static_cast<size_t>(blockIdx.x) * 2 + 1) * a.C + 4 * cg) = t1;
  }
}

// fixed-order sum of partial[i * stride] over i < splits by one warp: lanes stride over the partials in double,
// then a shuffle tree (the same tree every run -> deterministic)
__device__ __forceinline__ double warp_sum_partials(const float* __restrict__ partial, int splits, size_t stride, int lane) {
  double s = 0.0;
  for (int i = lane; i < splits; i += 32) s += static_cast<double>(partial[i * stride]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// one warp per channel
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int splits, int C, double n,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int update) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  const double s = warp_sum_partials(partial + c, splits, static_cast<size_t>(2) * C, lane);
  const double ss = warp_sum_partials(partial + C + c, splits, static_cast<size_t>(2) * C, lane);
  if (lane != 0) return;
  const double m = s / n;
  double var = ss / n - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = static_cast<float>(m);
  invstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kBnEps)));
  if (update && running_mean) {                                  // torch: momentum 0.1, unbiased variance
    running_mean[c] = (1.0f - kBnMomentum) * running_mean[c] + kBnMomentum * static_cast<float>(m);
    running_var[c] = (1.0f - kBnMomentum) * running_var[c] + kBnMomentum * static_cast<float>(var * n / (n - 1.0));
  }
}

// y = dropout(act(gamma * (z - mean) * invstd + beta)) -> concat buffer
__global__ void bn_apply_kernel(const float* __restrict__ z, size_t N, int C, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const uint8_t* __restrict__ keep,
                                int pix_per_sample, int act, float* __restrict__ y, int y_pitch, int y_coff,
                                int round_to_tf32) {
  const size_t total = N * C;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t n = i / C;
    float v = (z[i] - mean[c]) * invstd[c] * gamma[c] + beta[c];
    if (act == ACT_LEAKY) v = v > 0.f ? v : 0.2f * v;
    else if (act == ACT_RELU) v = fmaxf(v, 0.f);
    if (keep) v = keep[(n / pix_per_sample) * C + c] ? 2.0f * v : 0.0f;
    y[n * y_pitch + y_coff + c] = round_to_tf32 ? round_tf32(v) : v;         // operand of the next layer's MMAs
  }
}

// 16-byte forms of the two elementwise BatchNorm passes (C a power of two >= 4, pitches / offsets % 4 == 0, pixels per
// sample a power of two): index math by shifts instead of 64-bit divisions, four channels per thread
__global__ void __launch_bounds__(256)
bn_apply_vec_kernel(const float4* __restrict__ z, size_t n4, int cg_log2, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const uint8_t* __restrict__ keep, int pix_log2, int act, float* __restrict__ y, int y_pitch,
                    int y_coff, int round_to_tf32) {
  const int cg_mask = (1 << cg_log2) - 1;
  const float neg = act == ACT_LEAKY ? 0.2f : (act == ACT_RELU ? 0.0f : 1.0f);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i) & cg_mask;
    const size_t n = i >> cg_log2;
    const float4 v = z[i];
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + cg), s = __ldg(reinterpret_cast<const float4*>(invstd) + cg);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + cg), b = __ldg(reinterpret_cast<const float4*>(beta) + cg);
    float o[4] = {(v.x - m.x) * s.x * g.x + b.x, (v.y - m.y) * s.y * g.y + b.y, (v.z - m.z) * s.z * g.z + b.z,
                  (v.w - m.w) * s.w * g.w + b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = o[k] > 0.f ? o[k] : neg * o[k];
    if (keep) {
      const uchar4 kp = *reinterpret_cast<const uchar4*>(keep + (((n >> pix_log2) << (cg_log2 + 2)) + 4 * cg));
      o[0] = kp.x ? 2.0f * o[0] : 0.0f; o[1] = kp.y ? 2.0f * o[1] : 0.0f;
      o[2] = kp.z ? 2.0f * o[2] : 0.0f; o[3] = kp.w ? 2.0f * o[3] : 0.0f;
    }
    if (round_to_tf32) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = round_tf32(o[k]);
    }
    *reinterpret_cast<float4*>(y + n * y_pitch + y_coff + 4 * cg) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_vec_kernel(RedArgs a, int cg_log2, int pix_log2, const float* __restrict__ gamma,
                        const float* __restrict__ sum_g, const float* __restrict__ sum_gx, float4* __restrict__ z_inout,
                        int round_to_tf32) {
  const int cg_mask = (1 << cg_log2) - 1;
  const size_t n4 = (a.N << cg_log2);
  const float inv_n = 1.0f / static_cast<float>(a.N);
  const float neg = a.act == ACT_LEAKY ? 0.2f : 0.0f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i) & cg_mask;
    const size_t n = i >> cg_log2;
    const size_t o = n * a.y_pitch + a.y_coff + 4 * cg;
    const float4 yv = *reinterpret_cast<const float4*>(a.y + o);
    const float4 dy = *reinterpret_cast<const float4*>(a.dy + o);
    float g[4] = {dy.x, dy.y, dy.z, dy.w};
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
    if (a.keep) {
      const uchar4 kp = *reinterpret_cast<const uchar4*>(a.keep + (((n >> pix_log2) << (cg_log2 + 2)) + 4 * cg));
      g[0] = kp.x ? 2.0f * g[0] : 0.0f; g[1] = kp.y ? 2.0f * g[1] : 0.0f;
      g[2] = kp.z ? 2.0f * g[2] : 0.0f; g[3] = kp.w ? 2.0f * g[3] : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = yy[k] > 0.f ? g[k] : neg * g[k];
    const float4 zz = z_inout[i];
    const float zv[4] = {zz.x, zz.y, zz.z, zz.w};
    const float4 m = __ldg(reinterpret_cast<const float4*>(a.mean) + cg), s = __ldg(reinterpret_cast<const float4*>(a.invstd) + cg);
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + cg);
    const float4 sg = __ldg(reinterpret_cast<const float4*>(sum_g) + cg), sx = __ldg(reinterpret_cast<const float4*>(sum_gx) + cg);
    const float mm[4] = {m.x, m.y, m.z, m.w}, ss[4] = {s.x, s.y, s.z, s.w}, gg[4] = {gm.x, gm.y, gm.z, gm.w};
    const float s1[4] = {sg.x, sg.y, sg.z, sg.w}, s2[4] = {sx.x, sx.y, sx.z, sx.w};
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (zv[k] - mm[k]) * ss[k];
      const float dz = gg[k] * ss[k] * (g[k] - s1[k] * inv_n - xh * s2[k] * inv_n);
      r[k] = round_to_tf32 ? round_tf32(dz) : dz;
    }
    z_inout[i] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int splits, int C,
                                       float* __restrict__ sum_g, float* __restrict__ sum_gx,
                                       float* __restrict__ grad_gamma, float* __restrict__ grad_beta,
                                       float* __restrict__ grad_conv_bias) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  const double s = warp_sum_partials(partial + c, splits, static_cast<size_t>(2) * C, lane);
  const double sx = warp_sum_partials(partial + C + c, splits, static_cast<size_t>(2) * C, lane);
  if (lane != 0) return;
  sum_g[c] = static_cast<float>(s);
  sum_gx[c] = static_cast<float>(sx);
  grad_beta[c] = static_cast<float>(s);
  grad_gamma[c] = static_cast<float>(sx);
  grad_conv_bias[c] = 0.0f;      // d/d(bias) of a conv feeding a batch-stat BatchNorm is identically zero
}

// dz = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)), written in place over z
__global__ void bn_bwd_apply_kernel(RedArgs a, const float* __restrict__ gamma, const float* __restrict__ sum_g,
                                    const float* __restrict__ sum_gx, float* __restrict__ z_inout, int round_to_tf32) {
  const size_t total = a.N * a.C;
  const float inv_n = 1.0f / static_cast<float>(a.N);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % a.C);
    const size_t n = i / a.C;
    const float g = dy_effective(a, n, c);
    const float xh = (z_inout[i] - a.mean[c]) * a.invstd[c];
    const float dz = gamma[c] * a.invstd[c] * (g - sum_g[c] * inv_n - xh * sum_gx[c] * inv_n);
    z_inout[i] = round_to_tf32 ? round_tf32(dz) : dz;                        // operand of the dgrad / wgrad MMAs
  }
}

// ---- weight gradient: dW[tap][ci][co] = sum over pixels of the small grid -----------------------
// S = tensor living on the small grid (conv: dY, deconv: X), L = tensor on the 2x grid (conv: X,
// deconv: dY); pixel (gy, gx) of S pairs with (2gy + kh - 2, 2gx + kw - 2) of L.
struct WgradArgs {
  const float* S; int s_pitch, s_coff, s_c;
  const float* L; int l_pitch, l_coff, l_c;
  int gh, gw, batch;          // small grid
  int s_is_cout;              // 1: conv (S = dY -> co, L = X -> ci); 0: deconv (S = X -> ci, L = dY -> co)
  int cin, cout;
};

__global__ void __launch_bounds__(64)
wgrad_kernel(WgradArgs a, int gw_log2, int gh_log2, float* __restrict__ partial /*[splits][25][cin][cout]*/) {
  __shared__ __align__(16) float ss[16][32], sl[16][32];
  const int tap = blockIdx.x;
  const int kh = tap / 5, kw = tap % 5;
  const int co_tiles = (a.cout + 31) / 32;
  const int ci0 = (blockIdx.y / co_tiles) * 32, co0 = (blockIdx.y % co_tiles) * 32;
  const int s0 = a.s_is_cout ? co0 : ci0, l0 = a.s_is_cout ? ci0 : co0;
  const size_t npix = static_cast<size_t>(a.batch) * a.gh * a.gw;
  const size_t per = (npix + gridDim.z - 1) / gridDim.z;
  const size_t p_begin = blockIdx.z * per, p_end = min(npix, p_begin + per);
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;      // thread tile: 4 (ci) x 4 (co)
  float acc[4][4] = {};
  const int lh = 2 * a.gh, lw = 2 * a.gw;
  for (size_t p0 = p_begin; p0 < p_end; p0 += 16) {
    // stage 16 pixels x 32 channels of S and of L (zero where out of range)
    for (int i = threadIdx.x; i < 16 * 32; i += 64) {
      const int r = i >> 5, c = i & 31;
      const size_t p = p0 + r;
      float vs = 0.f, vl = 0.f;
      if (p < p_end) {
        // every grid extent of this net is a power of two: shifts, not 64-bit divisions
        const int gx = static_cast<int>(p) & (a.gw - 1);
        const int gy = static_cast<int>(p >> gw_log2) & (a.gh - 1);
        const int b = static_cast<int>(p >> (gw_log2 + gh_log2));
        if (s0 + c < a.s_c) vs = a.S[p * a.s_pitch + a.s_coff + s0 + c];
        const int ly = 2 * gy + kh - 2, lx = 2 * gx + kw - 2;
        if (ly >= 0 && ly < lh && lx >= 0 && lx < lw && l0 + c < a.l_c)
          vl = a.L[((static_cast<size_t>(b) * lh + ly) * lw + lx) * a.l_pitch + a.l_coff + l0 + c];
      }
      ss[r][c] = vs; sl[r][c] = vl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float4 vci = *reinterpret_cast<const float4*>(a.s_is_cout ? &sl[r][4 * ty] : &ss[r][4 * ty]);
      const float4 vco = *reinterpret_cast<const float4*>(a.s_is_cout ? &ss[r][4 * tx] : &sl[r][4 * tx]);
      const float xi[4] = {vci.x, vci.y, vci.z, vci.w}, yo[4] = {vco.x, vco.y, vco.z, vco.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xi[i], yo[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = partial + (static_cast<size_t>(blockIdx.z) * 25 + tap) * a.cin * a.cout;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + 4 * ty + i, co = co0 + 4 * tx + j;
      if (ci < a.cin && co < a.cout) dst[static_cast<size_t>(ci) * a.cout + co] = acc[i][j];
    }
}

// Layers with cin * cout <= 512 (conv1 1x16, conv2 16x32, deconv6 32x1): a 32 x 32 channel tile would be mostly
// padding and the reduction runs over up to a million pixels.  64 pixels are staged per step with shift-decoded
// coordinates (every grid dimension of this net is a power of two).  Thread = (pair, slice): `pairs_pad` (ci, co)
// pairs (two per thread when there are 512) x 256 / pairs_pad pixel slices of the staged chunk; the slices are summed
// in fixed order at the end, so the result is deterministic.
__global__ void __launch_bounds__(256)
wgrad_small_kernel(WgradArgs a, int gw_log2, int gh_log2, int sc_log2, int lc_log2, int pairs_log2,
                   float* __restrict__ partial /*[splits][25][cin][cout]*/) {
  __shared__ float ss[64 * 32], sl[64 * 32];
  const int tap = blockIdx.x;
  const int kh = tap / 5, kw = tap % 5;
  const size_t npix = static_cast<size_t>(a.batch) << (gw_log2 + gh_log2);
  const size_t per = (npix + gridDim.y - 1) / gridDim.y;
  const size_t p_begin = blockIdx.y * per, p_end = min(npix, p_begin + per);
  const int n_pairs = a.cin * a.cout;
  const int pairs_pad = 1 << pairs_log2;                 // >= n_pairs, <= 512
  const int per_thread = pairs_pad > 256 ? 2 : 1;
  const int lanes_log2 = pairs_pad > 256 ? 8 : pairs_log2;
  const int slices = 256 >> lanes_log2;                  // pixel slices of a chunk
  const int lane_pair = threadIdx.x & ((1 << lanes_log2) - 1), slice = threadIdx.x >> lanes_log2;
  // pair -> (S channel, L channel): conv S = dY (co), L = X (ci); deconv S = X (ci), L = dY (co)
  int cs[2] = {0, 0}, cl[2] = {0, 0};
  bool on[2] = {false, false};
  for (int k = 0; k < per_thread; ++k) {
    const int pr = lane_pair + 256 * k;
    on[k] = pr < n_pairs;
    const int ci = on[k] ? pr / a.cout : 0, co = on[k] ? pr % a.cout : 0;
    cs[k] = a.s_is_cout ? co : ci;
    cl[k] = a.s_is_cout ? ci : co;
  }
  float acc[2] = {0.f, 0.f};
  const int lh = 2 << gh_log2, lw = 2 << gw_log2;
  const int gw_mask = (1 << gw_log2) - 1, gh_mask = (1 << gh_log2) - 1;
  for (size_t p0 = p_begin; p0 < p_end; p0 += 64) {
    for (int e = threadIdx.x; e < (64 << sc_log2); e += 256) {
      const int r = e >> sc_log2, c = e & (a.s_c - 1);
      const size_t p = p0 + r;
      ss[e] = p < p_end ? a.S[p * a.s_pitch + a.s_coff + c] : 0.f;
    }
    for (int e = threadIdx.x; e < (64 << lc_log2); e += 256) {
      const int r = e >> lc_log2, c = e & (a.l_c - 1);
      const size_t p = p0 + r;
      float v = 0.f;
      if (p < p_end) {
        const int gx = static_cast<int>(p) & gw_mask, gy = static_cast<int>(p >> gw_log2) & gh_mask;
        const int b = static_cast<int>(p >> (gw_log2 + gh_log2));
        const int ly = 2 * gy + kh - 2, lx = 2 * gx + kw - 2;
        if (ly >= 0 && ly < lh && lx >= 0 && lx < lw)
          v = a.L[((static_cast<size_t>(b) * lh + ly) * lw + lx) * a.l_pitch + a.l_coff + c];
      }
      sl[e] = v;
    }
    __syncthreads();
    if (on[0]) {
#pragma unroll 4
      for (int r = slice; r < 64; r += slices) {     // ascending pixel order within a slice
        acc[0] = fmaf(ss[(r << sc_log2) + cs[0]], sl[(r << lc_log2) + cl[0]], acc[0]);
        if (per_thread == 2) acc[1] = fmaf(ss[(r << sc_log2) + cs[1]], sl[(r << lc_log2) + cl[1]], acc[1]);
      }
    }
    __syncthreads();
  }
  float* dst = partial + (static_cast<size_t>(blockIdx.y) * 25 + tap) * n_pairs;
  if (slices == 1) {
    if (on[0]) dst[lane_pair] = acc[0];                    // pair index = ci * cout + co
    if (per_thread == 2 && on[1]) dst[lane_pair + 256] = acc[1];
  } else {
    ss[threadIdx.x] = acc[0];                              // [slice][pair]
    __syncthreads();
    if (slice == 0 && on[0]) {
      float t = 0.f;
      for (int k = 0; k < slices; ++k) t += ss[(k << lanes_log2) + lane_pair];   // fixed order
      dst[lane_pair] = t;
    }
  }
}

// ---- weight gradient of the two single-channel edge layers -----------------------------------------
// conv1  : dW[co][tap] = sum dz[b,y,x,co] * mix[b, 2y+kh-2, 2x+kw-2]     (S = dz, 16 channels; L = the mixture)
// deconv6: dW[ci][tap] = sum x[b,y,x,ci]  * dz6[b, 2y+kh-2, 2x+kw-2]     (S = cat1, 32 channels; L = dz6)
// One CTA = one image x 8 rows of the 256 x 64 small grid: the (19 x 131)-sample window of L sits in shared memory,
// S streams through it one row at a time; thread = (channel, tap group), all lanes of a warp read the same L sample
// (broadcast).  Every CTA writes one partial [25][C]; wgrad_edge_finalize_kernel adds them in fixed order.  Memory
// bound (S is read once: 67 / 134 MB per 64 patches) instead of the 25 x re-read of wgrad_small_kernel.
constexpr int kEdgeRows = 8;
// Thread = (channel c, kernel row kh): warp kh of the 160-thread CTA holds the five kw taps of its row in registers
// and slides a five-sample window of L along x (two new samples per step, broadcast to all lanes); S is read straight
// from global memory, one coalesced 128-byte line per warp and step.  1 LDG + 2 LDS + 5 FFMA per step and no barrier
// inside the row loop (the round-2a form staged every S row through shared memory behind two barriers and spent 9
// instructions per 4 FFMA: 182 / 115 us per step for deconv6 / conv1).  C = 16: the two half-warps take the two halves
// of a row and are combined by one shuffle at the end.
template <int C>
__global__ void __launch_bounds__(160)
wgrad_edge_kernel(const float* __restrict__ S /*[B][256][64][C]*/, const float* __restrict__ L /*[B][512][128]*/,
                  float* __restrict__ partial /*[gridDim.y * gridDim.x][25][C]*/) {
  constexpr int kHalves = 32 / C;                        // lane groups along x: 1 (C = 32) or 2 (C = 16)
  constexpr int kXs = 64 / kHalves;                      // steps per row and lane
  constexpr int kWinW = 132, kWinH = 2 * kEdgeRows + 3;
  __shared__ float win[kWinH * kWinW];
  const int b = blockIdx.y, row0 = blockIdx.x * kEdgeRows;
  const float* Lb = L + static_cast<size_t>(b) * 512 * 128;
  for (int i = threadIdx.x; i < kWinH * kWinW; i += 160) {
    const int r = i / kWinW, c = i - r * kWinW;
    const int ly = 2 * row0 - 2 + r, lx = c - 2;
    win[i] = (ly >= 0 && ly < 512 && lx >= 0 && lx < 128 && c < 131) ? __ldg(Lb + ly * 128 + lx) : 0.0f;
  }
  __syncthreads();
  const int kh = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane % C, xh = lane / C;
  const int x0 = xh * kXs;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const float* Sb = S + (static_cast<size_t>(b) * 256 + row0) * 64 * C + c;
  for (int r = 0; r < kEdgeRows; ++r) {
    const float* wrow = win + (2 * r + kh) * kWinW + 2 * x0;       // sample 2x + kw - 2 sits at window column 2x + kw
    const float* srow = Sb + (static_cast<size_t>(r) * 64 + x0) * C;
    float l0 = wrow[0], l1 = wrow[1], l2 = wrow[2];
#pragma unroll 8
    for (int x = 0; x < kXs; ++x) {
      const float s = __ldg(srow + x * C);
      const float l3 = wrow[2 * x + 3], l4 = wrow[2 * x + 4];
      acc[0] = fmaf(s, l0, acc[0]); acc[1] = fmaf(s, l1, acc[1]); acc[2] = fmaf(s, l2, acc[2]);
      acc[3] = fmaf(s, l3, acc[3]); acc[4] = fmaf(s, l4, acc[4]);
      l0 = l2; l1 = l3; l2 = l4;
    }
  }
  if (kHalves == 2) {
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
  }
  float* dst = partial + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 25 * C;
  if (xh == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) dst[(kh * 5 + k) * C + c] = acc[k];
  }
}

// grad_w[c * 25 + tap] = sum over CTAs (warp per output, fixed order) of partial[cta][tap][c]
__global__ void wgrad_edge_finalize_kernel(const float* __restrict__ partial, int n_partials, int C,
                                           float* __restrict__ grad_w) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (o >= 25 * C) return;
  const double s = warp_sum_partials(partial + o, n_partials, static_cast<size_t>(25) * C, lane);
  if (lane == 0) grad_w[(o % C) * 25 + o / C] = static_cast<float>(s);
}

// sum the split partials (fixed order) and scatter to the torch layout
__global__ void wgrad_finalize_kernel(const float* __restrict__ partial, int splits, int cin, int cout,
                                      int transposed, float* __restrict__ grad_w) {
  const int total = 25 * cin * cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[static_cast<size_t>(k) * total + i];
    const int co = i % cout, ci = (i / cout) % cin, tap = i / (cout * cin);
    const size_t dst = transposed ? (static_cast<size_t>(ci) * cout + co) * 25 + tap
                                  : (static_cast<size_t>(co) * cin + ci) * 25 + tap;
    grad_w[dst] = s;
  }
}

// ---- deconv6 / loss ------------------------------------------------------------------------------
// dz6 = grad_mask * m * (1 - m)   (sigmoid backward, in place over the grad buffer)
__global__ void sigmoid_bwd_kernel(const float* __restrict__ mask, const float* __restrict__ grad_mask, size_t n,
                                   float* __restrict__ dz) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float m = mask[i];
    dz[i] = grad_mask[i] * m * (1.0f - m);
  }
}

// dX[b,ih,iw,ci] = sum_taps dz6[b, 2ih-2+kh, 2iw-2+kw] * w6[ci][tap]   (dgrad of ConvTranspose2d(32 -> 1))
// One CTA = 4 input rows x 64 columns x 32 channels of one image: the 11 x 128 window of dz it needs sits in shared
// memory with a 2-column zero border (the padding of the convolution), a thread owns 4 pixels (iw = pxg + 16 k) x 8
// channels and walks the 25 taps in (kh, kw) order: 4 LDS + 2 broadcast LDS.128 + 32 FFMA per tap, ~30 instructions
// per output against ~70 for the thread-per-(pixel, 4 channels) form with its redundant, branchy global loads
// (142 -> ~45 us per 64-patch step).
// The same kernel is conv1's training forward (1 -> 16 channels, + bias): C = 16 gives 8 rows per CTA.
constexpr int kD6dPitch = 128 + 4;                    // 2 zero columns each side
template <int C, bool kBias>
__global__ void __launch_bounds__(256)
edge_conv_kernel(const float* __restrict__ dz /*[B][512][128]*/, const float* __restrict__ w /*[25][C]*/,
                 const float* __restrict__ bias /*[C] or null*/, float* __restrict__ dx /*[B][256][64][C]*/, int batch) {
  constexpr int kGroups = C / 8;                      // 8-channel groups
  constexpr int kRows = 256 / (16 * kGroups);         // small-grid rows per CTA: 4 (C = 32) or 8 (C = 16)
  constexpr int kWin = 2 * kRows + 3;                 // image rows 2 ih0 - 2 .. 2 ih0 + 2 kRows
  __shared__ __align__(16) float sw[25 * C];
  __shared__ float sdz[kWin * kD6dPitch];
  const int b = blockIdx.y, ih0 = blockIdx.x * kRows;
  for (int i = threadIdx.x; i < 25 * C; i += 256) sw[i] = w[i];
  for (int i = threadIdx.x; i < kWin * kD6dPitch; i += 256) {
    const int r = i / kD6dPitch, c = i - r * kD6dPitch - 2;
    const int oh = 2 * ih0 - 2 + r;
    sdz[i] = (oh >= 0 && oh < 512 && c >= 0 && c < 128) ? dz[(static_cast<size_t>(b) * 512 + oh) * 128 + c] : 0.0f;
  }
  __syncthreads();
  const int cg = threadIdx.x / (16 * kRows);          // 8-channel group: uniform within a warp (weights broadcast)
  const int row = (threadIdx.x >> 4) % kRows, pxg = threadIdx.x & 15;
  float acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[k][u] = kBias ? bias[8 * cg + u] : 0.0f;
#pragma unroll
  for (int kh = 0; kh < 5; ++kh) {
    const float* drow = sdz + (2 * row + kh) * kD6dPitch + 2 * pxg;   // column 2 iw - 2 + kw + 2 (border) = 2 iw + kw
#pragma unroll
    for (int kw = 0; kw < 5; ++kw) {
      const float4 w0 = *reinterpret_cast<const float4*>(sw + (kh * 5 + kw) * C + 8 * cg);
      const float4 w1 = *reinterpret_cast<const float4*>(sw + (kh * 5 + kw) * C + 8 * cg + 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float g = drow[32 * k + kw];             // pixel iw = pxg + 16 k
        acc[k][0] = fmaf(g, w0.x, acc[k][0]); acc[k][1] = fmaf(g, w0.y, acc[k][1]);
        acc[k][2] = fmaf(g, w0.z, acc[k][2]); acc[k][3] = fmaf(g, w0.w, acc[k][3]);
        acc[k][4] = fmaf(g, w1.x, acc[k][4]); acc[k][5] = fmaf(g, w1.y, acc[k][5]);
        acc[k][6] = fmaf(g, w1.z, acc[k][6]); acc[k][7] = fmaf(g, w1.w, acc[k][7]);
      }
    }
  }
  const size_t pix0 = (static_cast<size_t>(b) * 256 + ih0 + row) * 64;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4* dst = reinterpret_cast<float4*>(dx + (pix0 + pxg + 16 * k) * C + 8 * cg);
    dst[0] = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
    dst[1] = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
  }
}

// generic fixed-order sum of an array -> one float (used for deconv6's bias gradient and the loss terms)
__global__ void __launch_bounds__(256) sum_stage1_kernel(const float* __restrict__ x, size_t n, float* __restrict__ partial) {
  __shared__ float s[256];
  const size_t per = (n + gridDim.x - 1) / gridDim.x;
  const size_t a = blockIdx.x * per, b = min(n, a + per);
  float acc = 0.f;
  for (size_t i = a + threadIdx.x; i < b; i += 256) acc += x[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}
// Fixed-order sum of a warp: lane l adds elements l, l + 32, ... in double, then a shuffle tree (deterministic; one
// thread walking all n elements serialised ~2,000 dependent L2 loads = 30-55 us per scalar).
__device__ __forceinline__ double warp_sum_double(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(v), o);
    const int hi = __shfl_xor_sync(0xffffffffu, __double2hiint(v), o);
    v += __hiloint2double(hi, lo);
  }
  return v;
}
__global__ void __launch_bounds__(32)
sum_stage2_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  s = warp_sum_double(s);
  if (threadIdx.x == 0) *out = static_cast<float>(s * scale);
}

// L = mean|m x - v| + mean|(1-m) x - max(x - v, 0)| ; grad wrt m
__global__ void __launch_bounds__(256)
l1_loss_kernel(const float* __restrict__ mask, const float* __restrict__ mix, const float* __restrict__ voc,
               size_t n, int two_term, float gscale, float* __restrict__ partial /*[grid][2]*/,
               float* __restrict__ grad_mask) {
  __shared__ float sv[256], sa[256];
  const size_t per = (n + gridDim.x - 1) / gridDim.x;
  const size_t a = blockIdx.x * per, b = min(n, a + per);
  float lv = 0.f, la = 0.f;
  for (size_t i = a + threadIdx.x; i < b; i += 256) {
    const float m = mask[i], x = mix[i], v = voc[i];
    const float dv = m * x - v;
    lv += fabsf(dv);
    float g = (dv > 0.f ? 1.f : (dv < 0.f ? -1.f : 0.f)) * x;         // torch: sign(0) = 0
    if (two_term) {
      const float da = (1.0f - m) * x - fmaxf(x - v, 0.0f);
      la += fabsf(da);
      g -= (da > 0.f ? 1.f : (da < 0.f ? -1.f : 0.f)) * x;
    }
    if (grad_mask) grad_mask[i] = g * gscale;
  }
  sv[threadIdx.x] = lv; sa[threadIdx.x] = la;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sv[threadIdx.x] += sv[threadIdx.x + o]; sa[threadIdx.x] += sa[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sv[0]; partial[2 * blockIdx.x + 1] = sa[0]; }
}
__global__ void __launch_bounds__(32)
l1_loss_finalize_kernel(const float* __restrict__ partial, int blocks, double n, float* __restrict__ out) {
  double v = 0.0, a = 0.0;
  for (int i = threadIdx.x; i < blocks; i += 32) { v += partial[2 * i]; a += partial[2 * i + 1]; }
  v = warp_sum_double(v);
  a = warp_sum_double(a);
  if (threadIdx.x == 0) {
    out[1] = static_cast<float>(v / n);
    out[2] = static_cast<float>(a / n);
    out[0] = static_cast<float>(v / n + a / n);
  }
}

// ---------------------------------------------------------------------------------------------
// workspace
struct TrainWs {
  float* cat[BUF_COUNT];      // activations (post BN / act / dropout), concat layout
  float* dcat[BUF_COUNT];     // gradients w.r.t. the same buffers
  float* z[12];               // pre-BN conv outputs (z[11] unused: the mask lives in the caller's tensor)
  float* dz6;                 // [B][512][128]
  float* w_fwd[12]; float* w_t[12];
  float* mean[12]; float* invstd[12]; float* sum_g[12]; float* sum_gx[12];
  float* red_partial;         // [kRedBlocksMax][2][512]
  float* wgrad_partial;
  size_t wgrad_partial_floats;
  float* scalar_partial;      // [1024 * 2]
  float* zero_bias;           // [512] zeros (data-gradient convolutions have no bias)
  float* splitk;              // split-K scratch of the tcgen05 conv kernels (non-cluster fallback)
  size_t splitk_bytes;
  size_t total;
};

static int ilog2_exact(int v) {
  for (int k = 0; k < 31; ++k) if ((1 << k) == v) return k;
  return -1;
}
// cin * cout <= 512 and power-of-two extents: wgrad_small_kernel
static bool wgrad_is_small(int li) {
  const LayerGeom& g = kLayers[li];
  const int gh = g.transposed ? g.hin : g.hout, gw = g.transposed ? g.win : g.wout;
  return g.cin * g.cout <= 512 && g.cin <= 32 && g.cout <= 32 && ilog2_exact(g.cin) >= 0 && ilog2_exact(g.cout) >= 0 &&
         ilog2_exact(gh) >= 0 && ilog2_exact(gw) >= 0;
}

static int wgrad_splits(int li, int batch) {
  const LayerGeom& g = kLayers[li];
  const int gh = g.transposed ? g.hin : g.hout, gw = g.transposed ? g.win : g.wout;
  const size_t npix = static_cast<size_t>(batch) * gh * gw;
  if (wgrad_is_small(li)) {                                      // 25 taps x splits CTAs of 256 threads
    size_t s = 48;
    const size_t max_small = (npix + 1023) / 1024;               // at least 1024 pixels per split
    return static_cast<int>(s > max_small ? (max_small ? max_small : 1) : s);
  }
  const int tiles = 25 * ((g.cin + 31) / 32) * ((g.cout + 31) / 32);
  int s = (148 * 16 + tiles - 1) / tiles;                        // ~16 CTAs of 64 threads per SM
  const size_t max_s = (npix + 255) / 256;                       // at least 256 pixels per split
  if (static_cast<size_t>(s) > max_s) s = static_cast<int>(max_s);
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return s;
}

// the S / L operands of layer li's weight gradient (see wgrad_tc.cu) in the training workspace layout
struct WgOperands { int gh, gw, s_c, l_c, s_pitch, l_pitch, s_coff, l_coff; };
static WgOperands wg_operands(int li) {
  const LayerGeom& g = kLayers[li];
  WgOperands o{};
  if (!g.transposed) {          // conv: S = dz (cout), L = x (cin)
    o.gh = g.hout; o.gw = g.wout; o.s_c = g.cout; o.s_pitch = g.cout; o.s_coff = 0;
    o.l_c = g.cin; o.l_pitch = g.in_buf >= 0 ? kBufGeom[g.in_buf].c : 1; o.l_coff = g.in_coff;
  } else {                      // deconv: S = x (cin), L = dz (cout)
    o.gh = g.hin; o.gw = g.win; o.s_c = g.cin; o.s_pitch = kBufGeom[g.in_buf].c; o.s_coff = g.in_coff;
    o.l_c = g.cout; o.l_pitch = g.cout; o.l_coff = 0;
  }
  return o;
}
static bool wgrad_on_tc(int li) {
  if (li < 1 || li > 10) return false;
  const WgOperands o = wg_operands(li);
  return wgrad_tc_supported(o.gh, o.gw, o.s_c, o.l_c, o.s_pitch, o.l_pitch, o.s_coff, o.l_coff);
}

static TrainWs carve_train(char* base, int batch) {
  TrainWs w{};
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += (floats * sizeof(float) + 1023) / 1024 * 1024;     // 1024-byte aligned: TMA bases, 16-byte row accesses
    return p;
  };
  for (int i = 0; i < BUF_COUNT; ++i) {
    // batch padded like the inference workspace: the deep-layer M tiles span 8 patches
    const size_t n = static_cast<size_t>(padded_batch(batch)) * kBufGeom[i].h * kBufGeom[i].w * kBufGeom[i].c;
    w.cat[i] = take(n);
    w.dcat[i] = take(n);
  }
  size_t wg = 0;
  for (int li = 0; li < 12; ++li) {
    const LayerGeom& g = kLayers[li];
    w.z[li] = li < 11 ? take(static_cast<size_t>(padded_batch(batch)) * g.hout * g.wout * g.cout) : nullptr;
    w.w_fwd[li] = take(static_cast<size_t>(25) * g.cin * g.cout);
    w.w_t[li] = take(static_cast<size_t>(25) * g.cin * g.cout);
    w.mean[li] = take(g.cout); w.invstd[li] = take(g.cout); w.sum_g[li] = take(g.cout); w.sum_gx[li] = take(g.cout);
    size_t need = static_cast<size_t>(wgrad_splits(li, batch)) * 25 * g.cin * g.cout;
    if (wgrad_on_tc(li)) {
      const WgOperands o = wg_operands(li);
      const size_t tc = wgrad_tc_partial_floats(o.gh, o.gw, batch, o.s_c, o.l_c);
      need = tc > need ? tc : need;
    }
    wg = need > wg ? need : wg;
  }
  {
    const size_t edge = static_cast<size_t>(batch) * (256 / kEdgeRows) * 25 * 32;     // wgrad_edge_kernel partials
    wg = edge > wg ? edge : wg;
  }
  w.dz6 = take(static_cast<size_t>(batch) * 512 * 128);
  w.red_partial = take(static_cast<size_t>(kRedBlocksMax) * 2 * 512);
  w.wgrad_partial = take(wg);
  w.wgrad_partial_floats = wg;
  w.scalar_partial = take(2048);
  w.zero_bias = take(512);
  // split-K scratch: upper bound over the forward / dgrad problems (M tiles x 128 x Cout x split <= 8)
  w.splitk_bytes = static_cast<size_t>(64) << 20;
  w.splitk = take(w.splitk_bytes / sizeof(float));
  w.total = off;
  return w;
}

static unsigned grid_for(size_t n) {
  size_t b = (n + 255) / 256;
  return static_cast<unsigned>(b > 148 * 32 ? 148 * 32 : (b ? b : 1));
}

// data-gradient problem of layer li: conv -> transposed conv (Cout -> Cin) and vice versa, on dz (dense, pitch Cout)
// writing into the gradient concat buffer of the layer's input
static ConvDesc dgrad_desc(int li) {
  const LayerGeom& g = kLayers[li];
  ConvDesc d;
  d.transposed = !g.transposed;
  d.cin = g.cout; d.cout = g.cin;
  d.hin = g.hout; d.win = g.wout; d.hout = g.hin; d.wout = g.win;
  d.in_pitch = g.cout; d.in_coff = 0;
  d.out_pitch = kBufGeom[g.in_buf].c; d.out_coff = g.in_coff;
  d.act = ACT_NONE;
  return d;
}
// train-mode forward problem of layer li: raw convolution + bias into the dense pre-BatchNorm buffer z
static ConvDesc train_fwd_desc(int li) {
  ConvDesc d = desc_of_layer(li);
  d.out_pitch = d.cout; d.out_coff = 0;
  d.act = ACT_NONE;
  return d;
}

}  // namespace svs

namespace svs {
int zc_plan_layer(svs_unet_plan* plan, int li, cudaStream_t st);
int zc_repack_layer(svs_unet_plan* plan, int li, const float* w_fold, cudaStream_t st);
int zc_launch_layer_io(const svs_unet_plan* plan, int li, const ZcIo& io, int batch, cudaStream_t st);
void zc_free_layers(svs_unet_plan* plan);
}  // namespace svs

struct svs_train_plan {
  int device = 0;
  // The zero-copy implicit-GEMM kernels of the inference path (zc_conv.cu, TF32 form) for the FORWARD of conv2-4 and
  // deconv3-5: `zcp` is an inference-plan shell that only carries their ZcLayer state (weights repacked per step).
  svs_unet_plan zcp;
  svs::TcLayer fwd[12];       // layers 1..10
  svs::TcLayer dgr[12];       // layers 1..10
  void* d6_weights = nullptr; // deconv6 as taps-as-N GEMM (deconv6_tc.cu), TF32
  CUtensorMap d6_tmap;
  // Weight gradients run on a side stream forked from / joined back into the caller's stream inside every backward
  // call (also under graph capture): they only read dz and the saved activations, so they overlap the data-gradient
  // chain (BatchNorm passes are HBM bound, the wgrad kernel is bound by operand bytes into shared memory).
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork[12] = {};
  cudaEvent_t ev_join = nullptr, ev_start = nullptr, ev_pack = nullptr;
};

using namespace svs;

extern "C" int svs_unet_train_plan_create(void* stream, svs_train_plan** plan_out) {
  SVS_REQUIRE(plan_out, "svs_unet_train_plan_create: null pointer");
  int dev = 0;
  SVS_CUDA_TRY(cudaGetDevice(&dev));
  int rc = svs_device_check(dev);
  if (rc != SVS_OK) return rc;
  svs_train_plan* plan = new (std::nothrow) svs_train_plan();
  if (!plan) return fail(SVS_ERR_CUDA, "svs_unet_train_plan_create: out of host memory");
  plan->device = dev;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int li = 1; li <= 10; ++li) {
    rc = tc_plan_one(plan->fwd[li], train_fwd_desc(li), nullptr, true, st);
    if (rc == SVS_OK) rc = tc_plan_one(plan->dgr[li], dgrad_desc(li), nullptr, true, st);
    if (rc == SVS_OK && !(plan->fwd[li].enabled && plan->dgr[li].enabled))
      rc = fail(SVS_ERR_NOT_IMPLEMENTED, "svs_unet_train_plan_create: no tcgen05 mapping for layer " + std::to_string(li));
    if (rc != SVS_OK) { svs_unet_train_plan_destroy(plan); return rc; }
  }
  plan->zcp.precision = SVS_PRECISION_TF32;
  plan->zcp.elem_size = 4;
  plan->zcp.device = dev;
  {
    static const bool zc_on = [] { const char* e = std::getenv("SVS_TRAIN_ZC"); return !(e && e[0] == '0'); }();
    for (int li : {1, 2, 3, 8, 9, 10}) {
      if (!zc_on) break;
      rc = zc_plan_layer(&plan->zcp, li, st);          // w_fold == nullptr: geometry, buffers and tensor maps only
      if (rc != SVS_OK) { svs_unet_train_plan_destroy(plan); return rc; }
    }
  }
  if (cudaMalloc(&plan->d6_weights, 32 * 32 * sizeof(float)) != cudaSuccess) {
    svs_unet_train_plan_destroy(plan);
    return fail(SVS_ERR_CUDA, "svs_unet_train_plan_create: cudaMalloc failed");
  }
  rc = d6_make_weight_map(plan->d6_weights, true, &plan->d6_tmap);
  if (rc != SVS_OK) { svs_unet_train_plan_destroy(plan); return rc; }
  bool ok = cudaStreamCreateWithFlags(&plan->side, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&plan->ev_join, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&plan->ev_start, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&plan->ev_pack, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < 12; ++i) ok = cudaEventCreateWithFlags(&plan->ev_fork[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    svs_unet_train_plan_destroy(plan);
    return fail(SVS_ERR_CUDA, "svs_unet_train_plan_create: stream / event creation failed");
  }
  *plan_out = plan;
  return SVS_OK;
}

extern "C" int svs_unet_train_plan_destroy(svs_train_plan* plan) {
  if (!plan) return SVS_OK;
  for (int li = 0; li < 12; ++li) { tc_free_one(plan->fwd[li]); tc_free_one(plan->dgr[li]); }
  zc_free_layers(&plan->zcp);
  if (plan->d6_weights) cudaFree(plan->d6_weights);
  for (int i = 0; i < 12; ++i) if (plan->ev_fork[i]) cudaEventDestroy(plan->ev_fork[i]);
  if (plan->ev_join) cudaEventDestroy(plan->ev_join);
  if (plan->ev_start) cudaEventDestroy(plan->ev_start);
  if (plan->ev_pack) cudaEventDestroy(plan->ev_pack);
  if (plan->side) cudaStreamDestroy(plan->side);
  delete plan;
  return SVS_OK;
}

extern "C" size_t svs_unet_train_workspace_bytes(int batch) {
  if (batch <= 0) return 0;
  return carve_train(nullptr, batch).total;
}

static int check_train_args(const svs_train_layer layers[12], const void* mix, int batch, void* workspace,
                            size_t workspace_bytes) {
  SVS_REQUIRE(layers && mix && workspace, "svs_unet_train: null pointer");
  SVS_REQUIRE(batch >= 1, "svs_unet_train: batch must be positive");   // B = 1 is fine: N = B*H*W >= 16 per channel
  SVS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "svs_unet_train: workspace must be 1024-byte aligned");
  for (int i = 0; i < 12; ++i) {
    SVS_REQUIRE(layers[i].weight && layers[i].bias, "svs_unet_train: weight/bias missing");
    SVS_REQUIRE((layers[i].bn_weight != nullptr) == (i != 11), "svs_unet_train: BatchNorm on every block but deconv6");
  }
  if (workspace_bytes < carve_train(nullptr, batch).total)
    return fail(SVS_ERR_WORKSPACE, "svs_unet_train: workspace too small");
  return SVS_OK;
}

namespace svs {
// per-channel two-stage reduction of layer li (mode 0: BatchNorm statistics, mode 1: BatchNorm backward sums);
// returns the number of stage-1 partials through *splits
template <int kMode>
static int run_channel_reduce(const RedArgs& a, float* partial, int* splits, bool fast, cudaStream_t st) {
  const int pix_log2 = ilog2_exact(a.pix_per_sample);
  if (fast && a.C % 4 == 0 && a.C <= 1024 && 256 % (a.C / 4) == 0 && a.y_pitch % 4 == 0 && a.y_coff % 4 == 0 &&
      (kMode == 0 || pix_log2 >= 0)) {
    const int rows = 256 / (a.C / 4);
    size_t blocks = (a.N + static_cast<size_t>(rows) * 8 - 1) / (static_cast<size_t>(rows) * 8);   // >= 8 iterations per CTA
    if (blocks > static_cast<size_t>(kRedBlocksMax)) blocks = kRedBlocksMax;
    if (blocks < 1) blocks = 1;
    channel_reduce_rows_kernel<kMode><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a, pix_log2 < 0 ? 0 : pix_log2, partial);
    SVS_CHECK_LAUNCH("channel_reduce_rows_kernel");
    *splits = static_cast<int>(blocks);
    return SVS_OK;
  }
  dim3 rgrid((a.C + 31) / 32, kRedSplits);
  channel_reduce_kernel<kMode><<<rgrid, 256, 0, st>>>(a, partial);
  SVS_CHECK_LAUNCH("channel_reduce_kernel");
  *splits = kRedSplits;
  return SVS_OK;
}
}  // namespace svs

extern "C" int svs_unet_train_forward(const svs_train_plan* plan, const svs_train_layer layers[12], const float* mix,
                                      int batch, int update_running_stats, float* mask_out, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  int rc = check_train_args(layers, mix, batch, workspace, workspace_bytes);
  if (rc != SVS_OK) return rc;
  SVS_REQUIRE(mask_out, "svs_unet_train_forward: null mask_out");
  int dev = 0;
  SVS_CUDA_TRY(cudaGetDevice(&dev));
  rc = svs_device_check(dev);
  if (rc != SVS_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TrainWs w = carve_train(static_cast<char*>(workspace), batch);
  const bool tc = plan != nullptr;
  if (tc) SVS_CUDA_TRY(cudaMemsetAsync(w.zero_bias, 0, 512 * sizeof(float), st));
  for (int li = 0; li < 12; ++li) {
    const LayerGeom& g = kLayers[li];
    const svs_train_layer& L = layers[li];
    train_pack_kernel<<<grid_for(25 * g.cin * g.cout), 256, 0, st>>>(L.weight, g.cin, g.cout, g.transposed ? 1 : 0,
                                                                    w.w_fwd[li], w.w_t[li]);
    SVS_CHECK_LAUNCH("train_pack_kernel");
    if (li == 11) {                                                // deconv6 + sigmoid -> mask
      if (tc) {
        rc = d6_pack(w.w_fwd[11], true, plan->d6_weights, st);
        if (rc != SVS_OK) return rc;
        svs_patch_view iv{const_cast<float*>(mix), nullptr, 512 * 128, 128, 1};
        svs_patch_view ov{mask_out, nullptr, 512 * 128, 128, 1};
        rc = d6_launch_raw(true, plan->d6_tmap, L.bias, w.cat[BUF_CAT1], &iv, &ov, nullptr, batch, 0, st);
      } else {
        rc = launch_deconv6_f32(w.cat[BUF_CAT1], w.w_fwd[11], L.bias, mask_out, batch, st);
      }
      if (rc != SVS_OK) return rc;
      // the backward needs the mask (sigmoid') and gets only the workspace: keep a copy
      SVS_CUDA_TRY(cudaMemcpyAsync(w.dz6, mask_out, sizeof(float) * static_cast<size_t>(batch) * 512 * 128,
                                   cudaMemcpyDeviceToDevice, st));
      break;
    }
    if (li == 0) {
      edge_conv_kernel<16, true><<<dim3(256 / 8, batch), 256, 0, st>>>(mix, w.w_fwd[0], L.bias, w.z[0], batch);
      SVS_CHECK_LAUNCH("edge_conv_kernel");
      rc = SVS_OK;
    } else if (tc && plan->zcp.zc[li].enabled) {
      svs_unet_plan* zp = const_cast<svs_unet_plan*>(&plan->zcp);  // zero-copy kernel: halo slab landed once per tile
      rc = zc_repack_layer(zp, li, w.w_fwd[li], st);
      if (rc != SVS_OK) return rc;
      ZcIo io;
      io.in = w.cat[g.in_buf]; io.out = w.z[li]; io.out_pitch = g.cout; io.out_coff = 0;
      io.bias = L.bias; io.act = ACT_NONE; io.keep_fp32 = 1; io.wait_first = 1;
      rc = zc_launch_layer_io(zp, li, io, batch, st);
    } else if (tc) {
      TcLayer& t = const_cast<TcLayer&>(plan->fwd[li]);            // the plan owns only packed weights: rewritten per step
      rc = tc_pack_one(t, w.w_fwd[li], true, st);
      if (rc != SVS_OK) return rc;
      TcIo io;
      io.in = w.cat[g.in_buf]; io.out = w.z[li]; io.bias = L.bias;
      io.splitk = w.splitk; io.splitk_bytes = w.splitk_bytes;
      io.out_flags = OUT_KEEP_FP32;
      rc = tc_launch(t, io, batch, true, st);
    } else {
      rc = launch_conv_direct_f32(w.cat[g.in_buf], kBufGeom[g.in_buf].c, g.in_coff, g.hin, g.win, g.cin,
                                  w.w_fwd[li], L.bias, w.z[li], g.cout, 0, g.hout, g.wout, g.cout, ACT_NONE,
                                  g.transposed, batch, false, st);
    }
    if (rc != SVS_OK) return rc;
    const size_t N = static_cast<size_t>(batch) * g.hout * g.wout;
    RedArgs a{};
    a.z = w.z[li]; a.C = g.cout; a.N = N; a.pix_per_sample = g.hout * g.wout;
    int splits = 0;
    rc = run_channel_reduce<0>(a, w.red_partial, &splits, true, st);
    if (rc != SVS_OK) return rc;
    bn_finalize_kernel<<<(g.cout + 3) / 4, 128, 0, st>>>(w.red_partial, splits, g.cout, static_cast<double>(N),
                                                            w.mean[li], w.invstd[li], L.bn_running_mean,
                                                            L.bn_running_var, update_running_stats);
    SVS_CHECK_LAUNCH("bn_finalize_kernel");
    {
      const int cg_log2 = ilog2_exact(g.cout / 4), pix_log2 = ilog2_exact(g.hout * g.wout);
      if (cg_log2 >= 0 && pix_log2 >= 0)
        bn_apply_vec_kernel<<<grid_for(N * g.cout / 4), 256, 0, st>>>(
            reinterpret_cast<const float4*>(w.z[li]), N * (g.cout / 4), cg_log2, w.mean[li], w.invstd[li], L.bn_weight,
            L.bn_bias, L.dropout_keep, pix_log2, g.act, w.cat[g.out_buf], kBufGeom[g.out_buf].c, g.out_coff, tc ? 1 : 0);
      else
        bn_apply_kernel<<<grid_for(N * g.cout), 256, 0, st>>>(w.z[li], N, g.cout, w.mean[li], w.invstd[li], L.bn_weight,
                                                             L.bn_bias, L.dropout_keep, g.hout * g.wout, g.act,
                                                             w.cat[g.out_buf], kBufGeom[g.out_buf].c, g.out_coff,
                                                             tc ? 1 : 0);
    }
    SVS_CHECK_LAUNCH("bn_apply_kernel");
  }
  return SVS_OK;
}

namespace svs {
static int run_wgrad(const TrainWs& w, const svs_train_layer& L, int li, int batch, const float* dY, int dy_pitch,
                     int dy_coff, const float* X, int x_pitch, int x_coff, bool tc, cudaStream_t st) {
  const LayerGeom& g = kLayers[li];
  if (li == 0 || li == 11) {       // single-channel edge layers: one streaming pass (both arithmetic modes)
    const float* S = li == 0 ? dY : X;
    const float* Lt = li == 0 ? X : dY;
    dim3 grid(256 / kEdgeRows, batch);
    const int n_partials = static_cast<int>(grid.x * grid.y);
    const int C = li == 0 ? 16 : 32;
    if (w.wgrad_partial_floats < static_cast<size_t>(n_partials) * 25 * C)
      return fail(SVS_ERR_WORKSPACE, "run_wgrad: partial buffer too small");
    if (li == 0) wgrad_edge_kernel<16><<<grid, 160, 0, st>>>(S, Lt, w.wgrad_partial);
    else wgrad_edge_kernel<32><<<grid, 160, 0, st>>>(S, Lt, w.wgrad_partial);
    SVS_CHECK_LAUNCH("wgrad_edge_kernel");
    wgrad_edge_finalize_kernel<<<(25 * C * 32 + 127) / 128, 128, 0, st>>>(w.wgrad_partial, n_partials, C, L.grad_weight);
    SVS_CHECK_LAUNCH("wgrad_edge_finalize_kernel");
    return SVS_OK;
  }
  if (tc && wgrad_on_tc(li)) {
    const WgOperands o = wg_operands(li);
    const float* S = g.transposed ? X : dY;
    const float* Lt = g.transposed ? dY : X;
    return wgrad_tc_launch(S, o.s_pitch, o.s_coff, o.s_c, Lt, o.l_pitch, o.l_coff, o.l_c, o.gh, o.gw, batch,
                           w.wgrad_partial, w.wgrad_partial_floats, L.grad_weight, st);
  }
  WgradArgs a{};
  a.cin = g.cin; a.cout = g.cout; a.batch = batch;
  if (!g.transposed) {             // conv: small grid = output (dY), large = input (X)
    a.S = dY; a.s_pitch = dy_pitch; a.s_coff = dy_coff; a.s_c = g.cout;
    a.L = X; a.l_pitch = x_pitch; a.l_coff = x_coff; a.l_c = g.cin;
    a.gh = g.hout; a.gw = g.wout; a.s_is_cout = 1;
  } else {                         // deconv: small grid = input (X), large = output (dY)
    a.S = X; a.s_pitch = x_pitch; a.s_coff = x_coff; a.s_c = g.cin;
    a.L = dY; a.l_pitch = dy_pitch; a.l_coff = dy_coff; a.l_c = g.cout;
    a.gh = g.hin; a.gw = g.win; a.s_is_cout = 0;
  }
  const int splits = wgrad_splits(li, batch);
  if (wgrad_is_small(li)) {
    wgrad_small_kernel<<<dim3(25, splits), 256, 0, st>>>(a, ilog2_exact(a.gw), ilog2_exact(a.gh), ilog2_exact(a.s_c),
                                                        ilog2_exact(a.l_c), ilog2_exact(g.cin * g.cout), w.wgrad_partial);
    SVS_CHECK_LAUNCH("wgrad_small_kernel");
  } else {
    dim3 grid(25, ((g.cin + 31) / 32) * ((g.cout + 31) / 32), splits);
    if (ilog2_exact(a.gw) < 0 || ilog2_exact(a.gh) < 0) return fail(SVS_ERR_INVALID_ARG, "run_wgrad: grid extents must be powers of two");
    wgrad_kernel<<<grid, 64, 0, st>>>(a, ilog2_exact(a.gw), ilog2_exact(a.gh), w.wgrad_partial);
    SVS_CHECK_LAUNCH("wgrad_kernel");
  }
  wgrad_finalize_kernel<<<grid_for(25 * g.cin * g.cout), 256, 0, st>>>(w.wgrad_partial, splits, g.cin, g.cout,
                                                                      g.transposed ? 1 : 0, L.grad_weight);
  SVS_CHECK_LAUNCH("wgrad_finalize_kernel");
  return SVS_OK;
}
}  // namespace svs

extern "C" int svs_unet_train_backward(const svs_train_plan* plan, const svs_train_layer layers[12], const float* mix,
                                       const float* grad_mask, int batch, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  return svs_unet_train_backward_layers(plan, layers, mix, grad_mask, batch, workspace, workspace_bytes, 0, 11, stream);
}

extern "C" int svs_unet_train_backward_layers(const svs_train_plan* plan, const svs_train_layer layers[12],
                                              const float* mix, const float* grad_mask, int batch, void* workspace,
                                              size_t workspace_bytes, int first_layer, int last_layer, void* stream) {
  int rc = check_train_args(layers, mix, batch, workspace, workspace_bytes);
  if (rc != SVS_OK) return rc;
  SVS_REQUIRE(first_layer >= 0 && last_layer <= 11 && first_layer <= last_layer,
              "svs_unet_train_backward_layers: bad layer range");
  SVS_REQUIRE(grad_mask, "svs_unet_train_backward: null grad_mask");
  for (int i = 0; i < 12; ++i) {
    SVS_REQUIRE(layers[i].grad_weight && layers[i].grad_bias, "svs_unet_train_backward: grad buffers missing");
    if (i != 11) SVS_REQUIRE(layers[i].grad_bn_weight && layers[i].grad_bn_bias, "svs_unet_train_backward: BN grad buffers missing");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TrainWs w = carve_train(static_cast<char*>(workspace), batch);
  const bool tc = plan != nullptr;
  static const bool fork_on = [] { const char* e = std::getenv("SVS_TRAIN_FORK"); return !(e && e[0] == '0'); }();
  cudaStream_t wst = (tc && fork_on && plan->side) ? plan->side : st;      // stream of the weight gradients
  bool forked = false;
  auto fork_wgrad = [&](int li) -> int {               // dz of layer li is complete on st: the side stream may read it
    if (wst == st) return SVS_OK;
    SVS_CUDA_TRY(cudaEventRecord(plan->ev_fork[li], st));
    SVS_CUDA_TRY(cudaStreamWaitEvent(wst, plan->ev_fork[li], 0));
    forked = true;
    return SVS_OK;
  };
  // The data-gradient kernels take channel-transposed weights repacked every step: with a side stream all of this
  // call's layers are packed there first, under the first BatchNorm passes, instead of one by one in front of each dgrad.
  bool packed_ahead = false;
  if (wst != st) {
    SVS_CUDA_TRY(cudaEventRecord(plan->ev_start, st));
    SVS_CUDA_TRY(cudaStreamWaitEvent(wst, plan->ev_start, 0));
    for (int li = last_layer < 10 ? last_layer : 10; li >= (first_layer > 1 ? first_layer : 1); --li) {
      rc = tc_pack_one(const_cast<TcLayer&>(plan->dgr[li]), w.w_t[li], true, wst);
      if (rc != SVS_OK) return rc;
    }
    SVS_CUDA_TRY(cudaEventRecord(plan->ev_pack, wst));
    forked = true;
    packed_ahead = true;
  }
  bool pack_waited = false;
  // ---- deconv6: sigmoid backward, bias / weight gradient, data gradient into dcat1 (all 32 channels) ----
  // dz6 holds the forward's mask; dz6 <- grad * m (1 - m) in place
  if (last_layer == 11) {
    const size_t n = static_cast<size_t>(batch) * 512 * 128;
    sigmoid_bwd_kernel<<<grid_for(n), 256, 0, st>>>(w.dz6, grad_mask, n, w.dz6);
    SVS_CHECK_LAUNCH("sigmoid_bwd_kernel");
    sum_stage1_kernel<<<1024, 256, 0, st>>>(w.dz6, n, w.scalar_partial);
    SVS_CHECK_LAUNCH("sum_stage1_kernel");
    sum_stage2_kernel<<<1, 32, 0, st>>>(w.scalar_partial, 1024, 1.0f, layers[11].grad_bias);
    SVS_CHECK_LAUNCH("sum_stage2_kernel");
    rc = fork_wgrad(11);
    if (rc != SVS_OK) return rc;
    rc = run_wgrad(w, layers[11], 11, batch, w.dz6, 1, 0, w.cat[BUF_CAT1], 32, 0, tc, wst);
    if (rc != SVS_OK) return rc;
    edge_conv_kernel<32, false><<<dim3(256 / 4, batch), 256, 0, st>>>(w.dz6, w.w_fwd[11], nullptr, w.dcat[BUF_CAT1], batch);
    SVS_CHECK_LAUNCH("edge_conv_kernel");
  }
  // ---- deconv5 .. deconv1, conv6 .. conv1 ----
  for (int li = last_layer < 10 ? last_layer : 10; li >= first_layer; --li) {
    const LayerGeom& g = kLayers[li];
    const svs_train_layer& L = layers[li];
    const size_t N = static_cast<size_t>(batch) * g.hout * g.wout;
    RedArgs a{};
    a.z = w.z[li]; a.C = g.cout; a.N = N;
    a.y = w.cat[g.out_buf]; a.dy = w.dcat[g.out_buf];
    a.y_pitch = kBufGeom[g.out_buf].c; a.y_coff = g.out_coff;
    a.keep = L.dropout_keep; a.mean = w.mean[li]; a.invstd = w.invstd[li];
    a.act = g.act; a.pix_per_sample = g.hout * g.wout;
    int splits = 0;
    rc = run_channel_reduce<1>(a, w.red_partial, &splits, true, st);
    if (rc != SVS_OK) return rc;
    bn_bwd_finalize_kernel<<<(g.cout + 3) / 4, 128, 0, st>>>(w.red_partial, splits, g.cout, w.sum_g[li],
                                                                w.sum_gx[li], L.grad_bn_weight, L.grad_bn_bias,
                                                                L.grad_bias);
    SVS_CHECK_LAUNCH("bn_bwd_finalize_kernel");
    {
      const int cg_log2 = ilog2_exact(g.cout / 4), pix_log2 = ilog2_exact(g.hout * g.wout);
      if (cg_log2 >= 0 && pix_log2 >= 0)
        bn_bwd_apply_vec_kernel<<<grid_for(N * g.cout / 4), 256, 0, st>>>(a, cg_log2, pix_log2, L.bn_weight, w.sum_g[li],
                                                                         w.sum_gx[li], reinterpret_cast<float4*>(w.z[li]),
                                                                         tc ? 1 : 0);
      else
        bn_bwd_apply_kernel<<<grid_for(N * g.cout), 256, 0, st>>>(a, L.bn_weight, w.sum_g[li], w.sum_gx[li], w.z[li],
                                                                 tc ? 1 : 0);
    }
    SVS_CHECK_LAUNCH("bn_bwd_apply_kernel");
    // now z[li] holds dz (gradient w.r.t. the conv output)
    const float* X = li == 0 ? mix : w.cat[g.in_buf];
    const int x_pitch = li == 0 ? 1 : kBufGeom[g.in_buf].c;
    rc = fork_wgrad(li);
    if (rc != SVS_OK) return rc;
    rc = run_wgrad(w, L, li, batch, w.z[li], g.cout, 0, X, x_pitch, g.in_coff, tc, wst);
    if (rc != SVS_OK) return rc;
    if (li == 0) break;                                            // the mixture needs no gradient
    // data gradient into dcat[in_buf][in_coff .. in_coff + cin): conv dgrad = transposed kernel, deconv dgrad =
    // conv kernel, both with channel-transposed weights.  Encoder layers ACCUMULATE into the skip half that the
    // decoder consumer has already written (decoder layers run first in this loop).
    const bool accumulate = !g.transposed;
    if (tc) {
      TcLayer& t = const_cast<TcLayer&>(plan->dgr[li]);
      if (!packed_ahead) {
        rc = tc_pack_one(t, w.w_t[li], true, st);
        if (rc != SVS_OK) return rc;
      } else if (!pack_waited) {
        SVS_CUDA_TRY(cudaStreamWaitEvent(st, plan->ev_pack, 0));
        pack_waited = true;
      }
      TcIo io;
      io.in = w.z[li]; io.out = w.dcat[g.in_buf]; io.bias = w.zero_bias;
      io.splitk = w.splitk; io.splitk_bytes = w.splitk_bytes;
      io.out_flags = OUT_KEEP_FP32 | (accumulate ? OUT_ACCUMULATE : 0);
      rc = tc_launch(t, io, batch, true, st);
    } else {
      rc = launch_conv_direct_f32(w.z[li], g.cout, 0, g.hout, g.wout, g.cout, w.w_t[li], nullptr, w.dcat[g.in_buf],
                                  kBufGeom[g.in_buf].c, g.in_coff, g.hin, g.win, g.cin, ACT_NONE, !g.transposed, batch,
                                  accumulate, st);
    }
    if (rc != SVS_OK) return rc;
  }
  if (forked) {                                        // join: the caller's stream continues after the last wgrad
    SVS_CUDA_TRY(cudaEventRecord(plan->ev_join, wst));
    SVS_CUDA_TRY(cudaStreamWaitEvent(st, plan->ev_join, 0));
  }
  return SVS_OK;
}

extern "C" int svs_l1_masked_loss(const float* mask, const float* mix, const float* voc, int64_t n, int two_term,
                                  float grad_scale, float* loss_out, float* grad_mask_out, float* scratch,
                                  void* stream) {
  SVS_REQUIRE(mask && mix && voc && loss_out && scratch && n > 0, "svs_l1_masked_loss: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  l1_loss_kernel<<<1024, 256, 0, st>>>(mask, mix, voc, static_cast<size_t>(n), two_term, grad_scale / static_cast<float>(n),
                                       scratch, grad_mask_out);
  SVS_CHECK_LAUNCH("l1_loss_kernel");
  l1_loss_finalize_kernel<<<1, 32, 0, st>>>(scratch, 1024, static_cast<double>(n), loss_out);
  SVS_CHECK_LAUNCH("l1_loss_finalize_kernel");
  return SVS_OK;
}
