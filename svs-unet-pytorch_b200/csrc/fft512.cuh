// 512-point complex FFT shared by the STFT (K1) and iSTFT (K2) kernels.
//
// A 1024-point real transform is one 512-point complex transform of the packed sequence
// z[n] = x[2n] + i x[2n+1] plus an O(N) split step.  The complex transform is computed by a GROUP
// of 64 threads (2 warps) holding 8 points each: three radix-8 passes (512 = 8*8*8) with two
// shared-memory exchanges.  All exchange layouts are bank-conflict free for 32-bit accesses:
//
//   exchange 1 (buffer X, SoA re/im):   value (k1, n2)   at  k1*72 + n2          (pitch 72 = 8 mod 32)
//   exchange 2 (buffer Y, SoA re/im):   value (k1, c, b) at  (k1*8 + c)*9 + b    (pitch 9)
//   exchange 3 (buffer X, linear Z[k]): value k          at  k + 4*(k >> 5)      (pitch 36 per 32)
//
// Groups synchronise with named barriers (bar.sync id, 64) so the four groups of a 256-thread CTA
// run their frames independently.
#pragma once
#include <cuda_runtime.h>

namespace svs {

constexpr int kFftScratchFloats = 576;            // per array (re or im), per buffer
constexpr int kFftGroupFloats = 4 * kFftScratchFloats;   // X.re X.im Y.re Y.im

__device__ __forceinline__ void group_bar(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// Forward 8-point DFT (e^{-2 pi i nk/8}), natural-order output, in place.
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  const float h = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  // b_k *= W_8^k
  b1 = make_float2(h * (b1.x + b1.y), h * (b1.y - b1.x));      // * (1 - i)/sqrt2
  b2 = mul_mi(b2);                                             // * (-i)
  b3 = make_float2(h * (b3.y - b3.x), -h * (b3.x + b3.y));     // * (-1 - i)/sqrt2
  // DFT4 of a -> X[0], X[2], X[4], X[6]
  {
    float2 s0 = cadd(a0, a2), s1 = csub(a0, a2), s2 = cadd(a1, a3), s3 = mul_mi(csub(a1, a3));
    v[0] = cadd(s0, s2); v[4] = csub(s0, s2); v[2] = cadd(s1, s3); v[6] = csub(s1, s3);
  }
  // DFT4 of b -> X[1], X[3], X[5], X[7]
  {
    float2 s0 = cadd(b0, b2), s1 = csub(b0, b2), s2 = cadd(b1, b3), s3 = mul_mi(csub(b1, b3));
    v[1] = cadd(s0, s2); v[5] = csub(s0, s2); v[3] = cadd(s1, s3); v[7] = csub(s1, s3);
  }
}

// Per-thread FFT twiddles live in a shared-memory table built once per CTA (they are the same for the
// four frame groups): keeping them in registers cost 28 registers per thread and capped occupancy.
//   a[k-1][j] = W_512^{j k}      (pass A, k = 1..7)
//   b[c-1][j] = W_64^{(j&7) c}   (pass B, c = 1..7)
constexpr int kFftTwiddleFloat2 = 2 * 7 * 64;

struct FftTwiddles {                 // shared-memory table (iSTFT: frees 28 registers per thread)
  const float2* a;
  const float2* b;
  int j;
  __device__ __forceinline__ float2 ta(int k) const { return a[(k - 1) * 64 + j]; }
  __device__ __forceinline__ float2 tb(int c) const { return b[(c - 1) * 64 + j]; }
};

struct FftTwiddlesReg {              // per-thread registers (STFT: shared-memory bandwidth is its limiter)
  float2 a[7], b[7];
  __device__ __forceinline__ float2 ta(int k) const { return a[k - 1]; }
  __device__ __forceinline__ float2 tb(int c) const { return b[c - 1]; }
};

__device__ __forceinline__ void load_fft_twiddles(FftTwiddlesReg& tw, const float2* __restrict__ tw1024, int j) {
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    tw.a[k - 1] = __ldg(&tw1024[(2 * j * k) & 1023]);
    tw.b[k - 1] = __ldg(&tw1024[(16 * (j & 7) * k) & 1023]);
  }
}

// all threads of the CTA cooperate; caller synchronises the CTA afterwards
__device__ __forceinline__ FftTwiddles build_fft_twiddles(float2* table, const float2* __restrict__ tw1024,
                                                          int tid, int nthreads, int j) {
  for (int i = tid; i < 7 * 64; i += nthreads) {
    const int k = i / 64 + 1, j = i % 64;
    table[i] = __ldg(&tw1024[(2 * j * k) & 1023]);
    table[7 * 64 + i] = __ldg(&tw1024[(16 * (j & 7) * k) & 1023]);
  }
  FftTwiddles tw;
  tw.a = table;
  tw.b = table + 7 * 64;
  tw.j = j;
  return tw;
}

__device__ __forceinline__ int z_addr(int k) { return k + 4 * (k >> 5); }

// Forward FFT of the 64-thread group.  in: v[n1] = z[j + 64 n1].  out: v[d] = Z[jj + 64 d] with
// jj = (j >> 3) + 8 (j & 7).  `xbuf` / `ybuf` = this group's exchange buffers X and Y (2 * kFftScratchFloats floats
// each).  On return every thread has finished with X; Y may be overwritten only after a further group_bar.
template <typename TW>
__device__ __forceinline__ void fft512_group(float2 (&v)[8], const TW& tw, float* xbuf, float* ybuf, int j, int bar) {
  float* xre = xbuf;
  float* xim = xbuf + kFftScratchFloats;
  float* yre = ybuf;
  float* yim = ybuf + kFftScratchFloats;
  // pass A: radix-8 over n1, twiddle W_512^{j k1}
  dft8(v);
#pragma unroll
  for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw.ta(k));
#pragma unroll
  for (int k = 0; k < 8; ++k) { xre[k * 72 + j] = v[k].x; xim[k * 72 + j] = v[k].y; }
  group_bar(bar);
  // pass B: thread (k1 = j>>3, b = j&7); radix-8 over a, twiddle W_64^{b c}
  const int k1 = j >> 3, b = j & 7;
#pragma unroll
  for (int a = 0; a < 8; ++a) { v[a].x = xre[k1 * 72 + 8 * a + b]; v[a].y = xim[k1 * 72 + 8 * a + b]; }
  dft8(v);
#pragma unroll
  for (int c = 1; c < 8; ++c) v[c] = cmul(v[c], tw.tb(c));
#pragma unroll
  for (int c = 0; c < 8; ++c) { yre[(k1 * 8 + c) * 9 + b] = v[c].x; yim[(k1 * 8 + c) * 9 + b] = v[c].y; }
  group_bar(bar);
  // pass C: thread (k1 = j>>3, c = j&7) reads its 8 b-values (row j of the pitch-9 layout)
#pragma unroll
  for (int bb = 0; bb < 8; ++bb) { v[bb].x = yre[j * 9 + bb]; v[bb].y = yim[j * 9 + bb]; }
  dft8(v);
}

}  // namespace svs
