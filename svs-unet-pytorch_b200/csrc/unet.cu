// UNet inference plan + forward orchestration (C ABI: svs_unet_*).
//
// Replaces UNet.forward of reference model.py:169-201 in eval mode and the mask application of
// inference.py:102,107.  The plan folds eval-mode BatchNorm into the weights once; the forward
// enqueues 12 layer launches (+ split-K reductions) on the caller's stream and touches no host
// state, so it can be captured into a CUDA graph.
#include "unet_internal.cuh"

#include <cstdlib>
#include <new>

namespace svs {

int launch_fold_pack(const svs_conv_params& p, int cin, int cout, bool transposed, float* w_out,
                     float* b_out, cudaStream_t st);
int launch_layer_direct(const svs_unet_plan* plan, int li, const Workspace& ws, const svs_patch_view* in,
                        const svs_patch_view* out, const int32_t* in_frames, int batch, int flags,
                        cudaStream_t st);
int launch_read_activation(const svs_unet_plan* plan, int layer, int batch, const Workspace& ws,
                           float* out_nchw, cudaStream_t st);
// conv_tc.cu
int tc_plan_layers(svs_unet_plan* plan, cudaStream_t st);
void tc_free_layers(svs_unet_plan* plan);
size_t tc_splitk_bytes(const svs_unet_plan* plan, int batch);
int tc_launch_layer(const svs_unet_plan* plan, int li, const Workspace& ws, int batch, cudaStream_t st);
int tc_launch_count(const svs_unet_plan* plan, int li, int batch);
// conv1_tc.cu
int c1t_plan(svs_unet_plan* plan, cudaStream_t st);
int c1z_plan(svs_unet_plan* plan, cudaStream_t st);
void c1z_free(svs_unet_plan* plan);
int c1z_launch(const svs_unet_plan* plan, const Workspace& ws, const svs_patch_view* in, int batch, cudaStream_t st);
void c1t_free(svs_unet_plan* plan);
bool c1t_applicable(const svs_patch_view* in, const int32_t* in_frames);
int c1t_launch(const svs_unet_plan* plan, const Workspace& ws, const svs_patch_view* in, int batch, cudaStream_t st);
// zc_conv.cu
int zc_plan_layer(svs_unet_plan* plan, int li, cudaStream_t st);
void zc_free_layers(svs_unet_plan* plan);
int zc_launch_layer(const svs_unet_plan* plan, int li, const Workspace& ws, int batch, cudaStream_t st);
// deconv6_tc.cu
int d6_plan(svs_unet_plan* plan, cudaStream_t st);
void d6_free(svs_unet_plan* plan);
int d6_launch(const svs_unet_plan* plan, const Workspace& ws, const svs_patch_view* in, const svs_patch_view* out,
              const int32_t* in_frames, int batch, int flags, cudaStream_t st);

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Workspace carve_workspace(char* base, int batch, int es, size_t splitk_bytes) {
  Workspace ws{};
  const int bp = padded_batch(batch);
  size_t off = 0;
  for (int i = 0; i < BUF_COUNT; ++i) {
    ws.buf[i] = base ? base + off : nullptr;
    off += align_up(static_cast<size_t>(bp) * kBufGeom[i].h * kBufGeom[i].w * kBufGeom[i].c * es, 1024);
  }
  ws.splitk = base ? reinterpret_cast<float*>(base + off) : nullptr;
  ws.splitk_bytes = splitk_bytes;
  off += align_up(splitk_bytes, 1024);
  ws.total_bytes = off;
  return ws;
}

}  // namespace svs

using namespace svs;

extern "C" int svs_unet_plan_create(const svs_conv_params layers[12], int precision, void* stream,
                                    svs_unet_plan** plan_out) {
  SVS_REQUIRE(layers && plan_out, "svs_unet_plan_create: null pointer");
  SVS_REQUIRE(precision == SVS_PRECISION_FP32 || precision == SVS_PRECISION_BF16 ||
                  precision == SVS_PRECISION_TF32,
              "svs_unet_plan_create: unknown precision");
  int dev = 0;
  SVS_CUDA_TRY(cudaGetDevice(&dev));
  int rc = svs_device_check(dev);
  if (rc != SVS_OK) return rc;
  for (int i = 0; i < 12; ++i) {
    SVS_REQUIRE(layers[i].weight && layers[i].bias, "svs_unet_plan_create: layer weight/bias missing");
    const bool has_bn = layers[i].bn_weight != nullptr;
    SVS_REQUIRE(has_bn == (i != 11), "svs_unet_plan_create: BatchNorm expected on every block but deconv6");
    if (has_bn)
      SVS_REQUIRE(layers[i].bn_bias && layers[i].bn_mean && layers[i].bn_var,
                  "svs_unet_plan_create: incomplete BatchNorm parameters");
  }
  svs_unet_plan* plan = new (std::nothrow) svs_unet_plan();
  if (!plan) return fail(SVS_ERR_CUDA, "svs_unet_plan_create: out of host memory");
  plan->precision = precision;
  plan->elem_size = precision == SVS_PRECISION_BF16 ? 2 : 4;
  plan->device = dev;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < 12; ++i) {
    const LayerGeom& g = kLayers[i];
    cudaError_t e1 = cudaMalloc(&plan->w_fold[i], sizeof(float) * 25 * g.cin * g.cout);
    cudaError_t e2 = cudaMalloc(&plan->b_fold[i], sizeof(float) * g.cout);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      svs_unet_plan_destroy(plan);
      return fail(SVS_ERR_CUDA, "svs_unet_plan_create: cudaMalloc of folded weights failed");
    }
    rc = launch_fold_pack(layers[i], g.cin, g.cout, g.transposed, plan->w_fold[i], plan->b_fold[i], st);
    if (rc != SVS_OK) { svs_unet_plan_destroy(plan); return rc; }
  }
  if (precision != SVS_PRECISION_FP32) {
    rc = tc_plan_layers(plan, st);
    if (rc != SVS_OK) { svs_unet_plan_destroy(plan); return rc; }
    for (int li = 1; li <= 10; ++li) {
      if (!plan->tc[li].enabled) continue;
      rc = zc_plan_layer(plan, li, st);
      if (rc != SVS_OK) { svs_unet_plan_destroy(plan); return rc; }
    }
    const char* dis = std::getenv("SVS_TC_DISABLE_MASK");
    if (!(dis && (std::strtoul(dis, nullptr, 0) & 1u))) {
      rc = c1t_plan(plan, st);
      if (rc != SVS_OK) { svs_unet_plan_destroy(plan); return rc; }
      rc = c1z_plan(plan, st);
      if (rc != SVS_OK) { svs_unet_plan_destroy(plan); return rc; }
    }
    if (!(dis && ((std::strtoul(dis, nullptr, 0) >> 11) & 1u))) {
      rc = d6_plan(plan, st);
      if (rc != SVS_OK) { svs_unet_plan_destroy(plan); return rc; }
    }
  }
  *plan_out = plan;
  return SVS_OK;
}

extern "C" int svs_unet_plan_destroy(svs_unet_plan* plan) {
  if (!plan) return SVS_OK;
  tc_free_layers(plan);
  zc_free_layers(plan);
  c1t_free(plan);
  c1z_free(plan);
  d6_free(plan);
  for (int i = 0; i < 12; ++i) {
    if (plan->w_fold[i]) cudaFree(plan->w_fold[i]);
    if (plan->b_fold[i]) cudaFree(plan->b_fold[i]);
  }
  delete plan;
  return SVS_OK;
}

extern "C" int svs_unet_plan_precision(const svs_unet_plan* plan) {
  return plan ? plan->precision : SVS_ERR_INVALID_ARG;
}

extern "C" size_t svs_unet_workspace_bytes(const svs_unet_plan* plan, int batch) {
  if (!plan || batch <= 0) return 0;
  return carve_workspace(nullptr, batch, plan->elem_size, tc_splitk_bytes(plan, batch)).total_bytes;
}

extern "C" int svs_unet_forward(const svs_unet_plan* plan, const svs_patch_view* in, const svs_patch_view* out,
                                const int32_t* in_frames, int batch, int flags, void* workspace,
                                size_t workspace_bytes, void* stream) {
  return svs_unet_forward_layers(plan, in, out, in_frames, batch, flags, workspace, workspace_bytes, 0, 11, stream);
}

extern "C" int svs_unet_forward_layers(const svs_unet_plan* plan, const svs_patch_view* in,
                                       const svs_patch_view* out, const int32_t* in_frames, int batch, int flags,
                                       void* workspace, size_t workspace_bytes, int first_layer, int last_layer,
                                       void* stream) {
  SVS_REQUIRE(plan && in && out && in->base && out->base && workspace, "svs_unet_forward: null pointer");
  SVS_REQUIRE(first_layer >= 0 && last_layer <= 11 && first_layer <= last_layer, "svs_unet_forward_layers: bad layer range");
  SVS_REQUIRE(batch > 0, "svs_unet_forward: batch must be positive");
  SVS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0,
              "svs_unet_forward: workspace must be 1024-byte aligned");
  Workspace ws = carve_workspace(static_cast<char*>(workspace), batch, plan->elem_size,
                                 tc_splitk_bytes(plan, batch));
  if (workspace_bytes < ws.total_bytes)
    return fail(SVS_ERR_WORKSPACE, "svs_unet_forward: workspace too small (need " +
                                       std::to_string(ws.total_bytes) + " bytes)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int li = first_layer; li <= last_layer; ++li) {
    int rc;
    if (li == 0 && plan->c1z_enabled && c1t_applicable(in, in_frames)) rc = c1z_launch(plan, ws, in, batch, st);
    else if (li == 0 && plan->c1_enabled && c1t_applicable(in, in_frames)) rc = c1t_launch(plan, ws, in, batch, st);
    else if (li == 11 && plan->d6_enabled) rc = d6_launch(plan, ws, in, out, in_frames, batch, flags, st);
    else if (plan->tc[li].enabled)
      rc = plan->zc[li].enabled ? zc_launch_layer(plan, li, ws, batch, st) : tc_launch_layer(plan, li, ws, batch, st);
    else rc = launch_layer_direct(plan, li, ws, in, out, in_frames, batch, flags, st);
    if (rc != SVS_OK) return rc;
  }
  return SVS_OK;
}

extern "C" int svs_unet_read_activation(const svs_unet_plan* plan, int layer, int batch, const void* workspace,
                                        float* out_nchw, void* stream) {
  SVS_REQUIRE(plan && workspace && out_nchw, "svs_unet_read_activation: null pointer");
  SVS_REQUIRE(layer >= 0 && layer <= 10, "svs_unet_read_activation: layer must be in [0, 10]");
  Workspace ws = carve_workspace(const_cast<char*>(static_cast<const char*>(workspace)), batch,
                                 plan->elem_size, tc_splitk_bytes(plan, batch));
  return launch_read_activation(plan, layer, batch, ws, out_nchw, static_cast<cudaStream_t>(stream));
}

extern "C" int svs_unet_launch_count(const svs_unet_plan* plan, int batch) {
  if (!plan || batch <= 0) return 0;
  int n = 0;
  for (int li = 0; li < 12; ++li)
    n += (plan->tc[li].enabled && !plan->zc[li].enabled) ? tc_launch_count(plan, li, batch) : 1;
  return n;
}
