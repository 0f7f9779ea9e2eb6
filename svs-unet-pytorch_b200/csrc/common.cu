// Error reporting, device checks and the per-device constant tables of libsvs_b200.so.
#include "svs_common.cuh"

#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

namespace svs {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int num_sms() {
  static int cached = 0;     // one process drives one GPU model; all B200s report 148
  if (cached == 0) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached = n;
  }
  return cached;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = std::getenv("SVS_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}

namespace {
struct DeviceTables {
  float2* tw = nullptr;
  float* hann = nullptr;
  float* env_both = nullptr;
  float* env_single = nullptr;
};
std::mutex g_tab_mutex;
std::map<int, DeviceTables> g_tables;
}  // namespace

int get_spectral_tables(SpectralTables* out) {
  int dev = 0;
  SVS_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  auto it = g_tables.find(dev);
  if (it == g_tables.end()) {
    const double pi = 3.14159265358979323846264338327950288;
    std::vector<float2> tw(1024);
    std::vector<float> hann(1024), env_both(256), env_single(768);
    std::vector<double> wsq(1024);
    for (int m = 0; m < 1024; ++m) {
      const double a = -2.0 * pi * m / 1024.0;
      tw[m] = make_float2(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
      // periodic Hann (scipy get_window('hann', 1024, fftbins=True)): 0.5 - 0.5 cos(2 pi n / N)
      const double w = 0.5 - 0.5 * std::cos(2.0 * pi * m / 1024.0);
      hann[m] = static_cast<float>(w);
      wsq[m] = w * w;
    }
    // librosa window_sumsquare accumulates float64 w^2 into a float32 buffer, earlier frame first
    // the kernel multiplies by the RECIPROCAL of librosa's float32 envelope (division only where
    // env > tiny(float32), else the sample is left as is -> reciprocal 1)
    auto recip = [](float env) { return env > 1.17549435e-38f ? 1.0f / env : 1.0f; };
    for (int r = 0; r < 768; ++r) env_single[r] = recip(static_cast<float>(wsq[r]));
    for (int r = 0; r < 256; ++r) {
      const float first = static_cast<float>(wsq[r + 768]);
      env_both[r] = recip(static_cast<float>(static_cast<double>(first) + wsq[r]));
    }
    DeviceTables t;
    SVS_CUDA_TRY(cudaMalloc(&t.tw, sizeof(float2) * 1024));
    SVS_CUDA_TRY(cudaMalloc(&t.hann, sizeof(float) * 1024));
    SVS_CUDA_TRY(cudaMalloc(&t.env_both, sizeof(float) * 256));
    SVS_CUDA_TRY(cudaMalloc(&t.env_single, sizeof(float) * 768));
    SVS_CUDA_TRY(cudaMemcpy(t.tw, tw.data(), sizeof(float2) * 1024, cudaMemcpyHostToDevice));
    SVS_CUDA_TRY(cudaMemcpy(t.hann, hann.data(), sizeof(float) * 1024, cudaMemcpyHostToDevice));
    SVS_CUDA_TRY(cudaMemcpy(t.env_both, env_both.data(), sizeof(float) * 256, cudaMemcpyHostToDevice));
    SVS_CUDA_TRY(cudaMemcpy(t.env_single, env_single.data(), sizeof(float) * 768, cudaMemcpyHostToDevice));
    it = g_tables.emplace(dev, t).first;
  }
  out->tw1024 = it->second.tw;
  out->hann = it->second.hann;
  out->env_both = it->second.env_both;
  out->env_single = it->second.env_single;
  return SVS_OK;
}

}  // namespace svs

extern "C" int svs_version(void) { return SVS_ABI_VERSION; }

extern "C" const char* svs_last_error(void) { return svs::g_last_error.c_str(); }

extern "C" int svs_device_check(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return svs::fail(SVS_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major != 10) {
    return svs::fail(SVS_ERR_UNSUPPORTED_ARCH,
                     std::string("libsvs_b200 is built for sm_100a only; device '") + prop.name +
                         "' is compute capability " + std::to_string(prop.major) + "." +
                         std::to_string(prop.minor) + " (no fallback path exists)");
  }
  return SVS_OK;
}
