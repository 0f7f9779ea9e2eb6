// Internal description of the UNet layers, the activation workspace and the plan object.
#pragma once
#include "svs_common.cuh"
#include <cuda.h>
#include <vector>

namespace svs {

enum Act { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2 };

// Activation buffers inside the workspace.  Every decoder input is a "concat buffer"
// [B][H][W][2C] whose first C channels are written by the previous decoder layer and whose last C
// channels are written by the matching encoder layer (torch.cat([deconv_out, conv_skip], 1) of
// reference model.py:186-198 realised as channel-offset stores).
enum Buf { BUF_CAT1 = 0, BUF_CAT2, BUF_CAT3, BUF_CAT4, BUF_CAT5, BUF_X6, BUF_COUNT };

struct BufGeom { int h, w, c; };
// H = frequency bins, W = frames, C = channel pitch.
static const BufGeom kBufGeom[BUF_COUNT] = {
    {256, 64, 32}, {128, 32, 64}, {64, 16, 128}, {32, 8, 256}, {16, 4, 512}, {8, 2, 512}};

struct LayerGeom {
  bool transposed;
  int cin, cout;
  int hin, win, hout, wout;
  int in_buf, in_coff;     // in_buf < 0: external patch view (conv1)
  int out_buf, out_coff;   // out_buf < 0: external patch view (deconv6)
  int act;
};
// layers[0..5] = conv1..conv6, layers[6..11] = deconv1..deconv6
static const LayerGeom kLayers[12] = {
    {false, 1, 16, 512, 128, 256, 64, -1, 0, BUF_CAT1, 16, ACT_LEAKY},
    {false, 16, 32, 256, 64, 128, 32, BUF_CAT1, 16, BUF_CAT2, 32, ACT_LEAKY},
    {false, 32, 64, 128, 32, 64, 16, BUF_CAT2, 32, BUF_CAT3, 64, ACT_LEAKY},
    {false, 64, 128, 64, 16, 32, 8, BUF_CAT3, 64, BUF_CAT4, 128, ACT_LEAKY},
    {false, 128, 256, 32, 8, 16, 4, BUF_CAT4, 128, BUF_CAT5, 256, ACT_LEAKY},
    {false, 256, 512, 16, 4, 8, 2, BUF_CAT5, 256, BUF_X6, 0, ACT_LEAKY},
    {true, 512, 256, 8, 2, 16, 4, BUF_X6, 0, BUF_CAT5, 0, ACT_RELU},
    {true, 512, 128, 16, 4, 32, 8, BUF_CAT5, 0, BUF_CAT4, 0, ACT_RELU},
    {true, 256, 64, 32, 8, 64, 16, BUF_CAT4, 0, BUF_CAT3, 0, ACT_RELU},
    {true, 128, 32, 64, 16, 128, 32, BUF_CAT3, 0, BUF_CAT2, 0, ACT_RELU},
    {true, 64, 16, 128, 32, 256, 64, BUF_CAT2, 0, BUF_CAT1, 0, ACT_RELU},
    {true, 32, 1, 256, 64, 512, 128, BUF_CAT1, 0, -1, 0, ACT_NONE},
};

// One convolution / transposed-convolution PROBLEM on NHWC buffers: an inference layer (desc_of_layer) or a pass of
// the training step (train-mode forward, data gradient).  The tcgen05 planners below take this, not a layer index.
struct ConvDesc {
  bool transposed = false;
  int cin = 0, cout = 0;
  int hin = 0, win = 0, hout = 0, wout = 0;
  int in_pitch = 0, in_coff = 0;     // channels per pixel of the input buffer / first channel read
  int out_pitch = 0, out_coff = 0;   // channels per pixel of the output buffer / first channel written
  int act = ACT_NONE;
};
inline ConvDesc desc_of_layer(int li) {
  const LayerGeom& g = kLayers[li];
  ConvDesc d;
  d.transposed = g.transposed; d.cin = g.cin; d.cout = g.cout;
  d.hin = g.hin; d.win = g.win; d.hout = g.hout; d.wout = g.wout;
  d.in_pitch = g.in_buf >= 0 ? kBufGeom[g.in_buf].c : 1; d.in_coff = g.in_coff;
  d.out_pitch = g.out_buf >= 0 ? kBufGeom[g.out_buf].c : 1; d.out_coff = g.out_coff;
  d.act = g.act;
  return d;
}
// epilogue modifiers of the tcgen05 conv kernels
enum OutFlags { OUT_ACCUMULATE = 1,   // out += result (data gradients meeting at a skip connection)
                OUT_KEEP_FP32 = 2 };  // fp32 outputs are NOT rounded to TF32 (pre-BatchNorm z, gradients)
struct TcIo {                         // buffers of one launch
  const void* in = nullptr;
  void* out = nullptr;
  const float* bias = nullptr;
  float* splitk = nullptr;
  size_t splitk_bytes = 0;
  int out_flags = 0;
  long long* dbg = nullptr;
};

constexpr int kBatchPad = 8;      // workspace batch is padded to a multiple of 8 (deep-layer M tiles)
inline int padded_batch(int b) { return (b + kBatchPad - 1) / kBatchPad * kBatchPad; }

struct Workspace {
  char* buf[BUF_COUNT];
  float* splitk;            // fp32 partial accumulators for split-K layers
  size_t splitk_bytes;
  size_t total_bytes;
};
// Carves `base` (may be nullptr to only compute sizes) for `batch` patches of element size `es`.
Workspace carve_workspace(char* base, int batch, int es, size_t splitk_bytes);

// ---- tensor-core (tcgen05) layer plan --------------------------------------------------------
struct TcChunk {            // one K-chunk of the implicit GEMM: which input slab feeds the MMA
  int c_inner;              // coordinate in the innermost (channel / parity-merged channel) dimension
  int dw;                   // offset added to the tile's w coordinate
  int ph;                   // coordinate in the row-parity dimension (conv) or 0 (deconv)
  int dh;                   // offset added to the tile's h coordinate
};

struct TcPhase {
  int n_chunks;
  int chunk_begin;          // into TcLayer::chunks
  int py, px;               // output sub-pixel phase (deconv) or 0
  int64_t b_elem_off;       // element offset of this phase's [Cout][K] weight matrix
};

struct TcLayer {
  bool enabled = false;
  int layer = -1;
  ConvDesc d;                        // the problem this plan was made for
  int2* d_src = nullptr;             // device: (tap, first channel) of every K chunk (kept for re-packing)
  int bw = 0, bh = 0, nb = 0;        // M tile = bw x bh x nb = 128 pixels
  int gw = 0, gh = 0;                // pixel grid the M index runs over (conv: output, deconv: input)
  int block_n = 0;                   // N tile
  int swz = 128;                     // bytes of K per smem row == TMA/UMMA swizzle span (32/64/128)
  int block_k = 0;                   // K elements per chunk = swz / elem_size
  int n_phases = 1;
  bool merged = false;               // deconv: all 4 sub-pixel phases stacked along N (zero weights where a
                                     // phase does not use a tap) so each input slab is loaded once, not 25/9 times
  TcPhase phases[4];
  std::vector<TcChunk> chunks;       // host copy
  TcChunk* d_chunks = nullptr;       // device copy
  void* d_weights = nullptr;         // packed [phase][Cout][K] K-major (bf16 or fp32/tf32)
  int k_total[4] = {0, 0, 0, 0};
  CUtensorMap tmap_b[4];             // per phase, 2-D [Cout][K]
  bool has_wide = false;             // deep layers also carry a 256-row box: large batches use 128 x 256 tiles
  CUtensorMap tmap_b_wide[4];
  // A-operand tensor map depends on (workspace, batch): a small round-robin cache (one workspace per stream and
  // batch size is in flight when several streams share a plan)
  static constexpr int kTmapCache = 8;
  mutable CUtensorMap tmap_a[kTmapCache];
  mutable const void* tmap_a_base[kTmapCache] = {};
  mutable int tmap_a_batch[kTmapCache] = {};
  mutable int tmap_a_next = 0;
};

// ---- zero-copy tcgen05 layers (zc_conv.cu) -------------------------------------------------------
constexpr int kZcMaxTaps = 80, kZcMaxSlabs = 8;
struct ZcSchedule {
  int n_slabs, n_taps;
  int slab_c[kZcMaxSlabs], slab_ph[kZcMaxSlabs];   // TMA coordinates (channel window, row parity) of each halo slab
  // taps are ordered by slab: (slab, dy, dx) and the mask of non-zero 32-byte k-steps
  signed char tap_slab[kZcMaxTaps], tap_dy[kZcMaxTaps], tap_dx[kZcMaxTaps], tap_kmask[kZcMaxTaps];
};
struct ZcLayer {
  bool enabled = false;
  bool resident = false;
  ZcSchedule sch{};
  int n_total = 0, row_elems = 0, row_bytes = 128;
  void* d_weights = nullptr;
  CUtensorMap tmap_b;
  CUtensorMap tmap_b_half;           // box of n_total / 2 rows: one CTA's share of a multicast weight chunk
  CUtensorMap tmap_b_quarter;        // box of n_total / 4 rows: one sub-pixel phase block of a merged deconv chunk
};

// buffers / epilogue of one zero-copy launch (zc_launch_layer_io)
struct ZcIo {
  const void* in = nullptr;
  void* out = nullptr;
  int out_pitch = 0, out_coff = 0;
  const float* bias = nullptr;
  int act = ACT_NONE;
  int keep_fp32 = 0;                 // fp32 outputs unrounded
  int wait_first = 0;                // weights are repacked in-stream: wait for the previous grid before preloading them
};

}  // namespace svs

struct svs_unet_plan {
  int precision = SVS_PRECISION_FP32;
  int elem_size = 4;
  int device = 0;
  float* w_fold[12] = {};            // folded fp32 weights [25][Cin][Cout]
  float* b_fold[12] = {};            // folded fp32 bias [Cout]
  svs::TcLayer tc[12];
  svs::ZcLayer zc[12];
  // conv1 as a thread-built im2col GEMM (conv1_tc.cu)
  bool c1_enabled = false;
  void* c1_weights = nullptr;
  CUtensorMap c1_tmap_w;
  // deconv6 as a taps-as-N GEMM + col2im gather (deconv6_tc.cu)
  bool c1z_enabled = false;          // conv1 with the image as the A operand (conv1_zc.cu); dense inputs
  float* c1z_weights = nullptr;
  CUtensorMap c1z_tmap_w;
  bool d6_enabled = false;
  void* d6_weights = nullptr;
  CUtensorMap d6_tmap_w;
};
