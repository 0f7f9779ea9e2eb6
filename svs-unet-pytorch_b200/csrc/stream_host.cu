// Host-buffer batches through the UNet: the native loop behind pipeline.PatchStreamer.
//
// Reference inference.py:97-110 moves ONE patch to the device, runs the model and copies the mask back with a
// synchronous .to(device) / .cpu() pair.  Here batch i+1's upload and batch i-1's download run on their own streams
// (one DMA engine per direction) under the kernels of batch i, n_slots-deep device staging, ordered by events.  The
// loop used to live in Python: ~20 torch calls = 0.30 ms of host time per 64-patch step against 0.34 ms of PCIe time,
// so any slower or busier host CPU made the interpreter, not the link, the bottleneck (measured 0.39 -> 0.60 ms per
// step between two boxes).  In C the enqueue cost is the 12 launches + 2 copies of a step.
#include "svs_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <utility>
#include <vector>

struct svs_patch_stream {
  int n_slots = 0;
  std::vector<cudaEvent_t> ev_in, ev_cmp, ev_out;   // slot uploaded / slot consumed + result ready / result downloaded
  // Odd steps copy on two streams of the object's own (forked from / joined into the caller's copy streams): with one
  // stream per direction a copy is only handed to the DMA engine once its predecessor in the stream has retired, and
  // that hand-over (plus the cross-stream event in front of every download) left the engines idle ~10 % of the time.
  cudaStream_t alt_in = nullptr, alt_out = nullptr;
  cudaEvent_t ev_fork_in = nullptr, ev_fork_out = nullptr, ev_join_in = nullptr, ev_join_out = nullptr;
};

extern "C" int svs_patch_stream_create(int n_slots, svs_patch_stream** out) {
  using namespace svs;
  SVS_REQUIRE(out != nullptr, "svs_patch_stream_create: null pointer");
  SVS_REQUIRE(n_slots >= 2 && n_slots <= 64, "svs_patch_stream_create: n_slots must be in [2, 64]");
  auto* ps = new svs_patch_stream;
  ps->n_slots = n_slots;
  for (std::vector<cudaEvent_t>* v : {&ps->ev_in, &ps->ev_cmp, &ps->ev_out}) {
    v->resize(n_slots, nullptr);
    for (int i = 0; i < n_slots; ++i) {
      const cudaError_t e = cudaEventCreateWithFlags(&(*v)[i], cudaEventDisableTiming);
      if (e != cudaSuccess) {
        svs_patch_stream_destroy(ps);
        return fail(SVS_ERR_CUDA, std::string("svs_patch_stream_create: ") + cudaGetErrorString(e));
      }
    }
  }
  static const bool dual = [] { const char* e = std::getenv("SVS_STREAM_DUAL"); return !(e && e[0] == '0'); }();
  if (dual) {
    bool ok = cudaStreamCreateWithFlags(&ps->alt_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ps->alt_out, cudaStreamNonBlocking) == cudaSuccess;
    for (cudaEvent_t* e : {&ps->ev_fork_in, &ps->ev_fork_out, &ps->ev_join_in, &ps->ev_join_out})
      ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      svs_patch_stream_destroy(ps);
      return fail(SVS_ERR_CUDA, "svs_patch_stream_create: stream / event creation failed");
    }
  }
  *out = ps;
  return SVS_OK;
}

extern "C" int svs_patch_stream_destroy(svs_patch_stream* ps) {
  if (!ps) return SVS_OK;
  for (std::vector<cudaEvent_t>* v : {&ps->ev_in, &ps->ev_cmp, &ps->ev_out})
    for (cudaEvent_t e : *v)
      if (e) cudaEventDestroy(e);                    // deferred by the runtime until the event has completed
  for (cudaEvent_t e : {ps->ev_fork_in, ps->ev_fork_out, ps->ev_join_in, ps->ev_join_out})
    if (e) cudaEventDestroy(e);
  if (ps->alt_in) cudaStreamDestroy(ps->alt_in);
  if (ps->alt_out) cudaStreamDestroy(ps->alt_out);
  delete ps;
  return SVS_OK;
}

extern "C" int svs_patch_stream_run(svs_patch_stream* ps, const svs_unet_plan* plan, const float* const* host_in,
                                    float* const* host_out, int n_steps, int batch, int flags, float* const* dev_in,
                                    float* const* dev_out, void* workspace, size_t workspace_bytes, void* stream_h2d,
                                    void* stream_compute, void* stream_d2h) {
  using namespace svs;
  SVS_REQUIRE(ps && plan && host_in && host_out && dev_in && dev_out && workspace, "svs_patch_stream_run: null pointer");
  SVS_REQUIRE(n_steps >= 0 && batch > 0, "svs_patch_stream_run: bad sizes");
  SVS_REQUIRE(stream_h2d != stream_compute && stream_compute != stream_d2h && stream_h2d != stream_d2h,
              "svs_patch_stream_run: the three streams must be distinct");
  cudaStream_t s_in = static_cast<cudaStream_t>(stream_h2d), s_cmp = static_cast<cudaStream_t>(stream_compute),
               s_out = static_cast<cudaStream_t>(stream_d2h);
  const size_t bytes = static_cast<size_t>(batch) * SVS_PATCH_BINS * SVS_PATCH_FRAMES * sizeof(float);
  const int64_t stride_b = static_cast<int64_t>(SVS_PATCH_BINS) * SVS_PATCH_FRAMES;
  cudaStream_t s_in0 = s_in, s_out0 = s_out;
  bool dual = ps->alt_in != nullptr && n_steps > 1;
  if (dual) {
    // Two downloads into the SAME host buffer must keep their order: fine when they share a stream (even distance),
    // otherwise fall back to one stream per direction.
    std::vector<std::pair<const void*, int>> outs(n_steps);
    for (int i = 0; i < n_steps; ++i) outs[i] = {host_out[i], i & 1};
    std::sort(outs.begin(), outs.end());
    for (int i = 1; i < n_steps && dual; ++i)
      if (outs[i].first == outs[i - 1].first && outs[i].second != outs[i - 1].second) dual = false;
  }
  if (dual) {                                        // the object's streams start after whatever the caller's hold
    SVS_CUDA_TRY(cudaEventRecord(ps->ev_fork_in, s_in0));
    SVS_CUDA_TRY(cudaStreamWaitEvent(ps->alt_in, ps->ev_fork_in, 0));
    SVS_CUDA_TRY(cudaEventRecord(ps->ev_fork_out, s_out0));
    SVS_CUDA_TRY(cudaStreamWaitEvent(ps->alt_out, ps->ev_fork_out, 0));
  }
  for (int i = 0; i < n_steps; ++i) {
    const int k = i % ps->n_slots;
    if (dual) {
      s_in = (i & 1) ? ps->alt_in : s_in0;
      s_out = (i & 1) ? ps->alt_out : s_out0;
    }
    SVS_REQUIRE(host_in[i] && host_out[i] && dev_in[k] && dev_out[k], "svs_patch_stream_run: null buffer");
    // upload into slot k once the forward that last read it has finished (a never-recorded event does not block)
    SVS_CUDA_TRY(cudaStreamWaitEvent(s_in, ps->ev_cmp[k], 0));
    SVS_CUDA_TRY(cudaMemcpyAsync(dev_in[k], host_in[i], bytes, cudaMemcpyHostToDevice, s_in));
    SVS_CUDA_TRY(cudaEventRecord(ps->ev_in[k], s_in));
    // forward once the upload has landed and result slot k has been drained
    SVS_CUDA_TRY(cudaStreamWaitEvent(s_cmp, ps->ev_in[k], 0));
    SVS_CUDA_TRY(cudaStreamWaitEvent(s_cmp, ps->ev_out[k], 0));
    svs_patch_view iv{dev_in[k], nullptr, stride_b, SVS_PATCH_FRAMES, 1};
    svs_patch_view ov{dev_out[k], nullptr, stride_b, SVS_PATCH_FRAMES, 1};
    const int rc = svs_unet_forward(plan, &iv, &ov, nullptr, batch, flags, workspace, workspace_bytes, s_cmp);
    if (rc != SVS_OK) return rc;
    SVS_CUDA_TRY(cudaEventRecord(ps->ev_cmp[k], s_cmp));
    // download
    SVS_CUDA_TRY(cudaStreamWaitEvent(s_out, ps->ev_cmp[k], 0));
    SVS_CUDA_TRY(cudaMemcpyAsync(host_out[i], dev_out[k], bytes, cudaMemcpyDeviceToHost, s_out));
    SVS_CUDA_TRY(cudaEventRecord(ps->ev_out[k], s_out));
  }
  if (dual) {                                        // join: the caller's copy streams end after the last alternate copy
    SVS_CUDA_TRY(cudaEventRecord(ps->ev_join_in, ps->alt_in));
    SVS_CUDA_TRY(cudaStreamWaitEvent(s_in0, ps->ev_join_in, 0));
    SVS_CUDA_TRY(cudaEventRecord(ps->ev_join_out, ps->alt_out));
    SVS_CUDA_TRY(cudaStreamWaitEvent(s_out0, ps->ev_join_out, 0));
  }
  return SVS_OK;
}
