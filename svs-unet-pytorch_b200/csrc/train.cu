// T1: training step (train-mode forward, masked-L1 loss, backward) — see svs_b200.h.
#include "svs_common.cuh"

extern "C" size_t svs_unet_train_workspace_bytes(int batch) {
  (void)batch;
  return 0;
}

extern "C" int svs_unet_train_step(const svs_train_layer layers[12], const float* mix, const float* voc,
                                   int batch, int two_term, int update_running_stats, float* loss_out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  (void)layers; (void)mix; (void)voc; (void)batch; (void)two_term; (void)update_running_stats;
  (void)loss_out; (void)workspace; (void)workspace_bytes; (void)stream;
  return svs::fail(SVS_ERR_NOT_IMPLEMENTED, "svs_unet_train_step: not built yet (no fallback is provided)");
}
