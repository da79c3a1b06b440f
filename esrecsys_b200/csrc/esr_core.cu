// Library plumbing of libesr: version, error strings, device info.
#include <string.h>

#include "esr_common.cuh"

namespace esr {

static thread_local char g_cuda_err[512] = "";

void set_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
  const char* base = strrchr(file, '/');
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), what,
           base ? base + 1 : file, line);
  cudaGetLastError();  // clear the sticky-less error so the next launch check starts clean
}

int sm_count() {
  static int cached[64] = {};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

}  // namespace esr

extern "C" int esr_version(void) { return ESR_VERSION; }

extern "C" const char* esr_strerror(int rc) {
  switch (rc) {
    case ESR_OK: return "ok";
    case ESR_EINVAL: return "invalid argument (shape, alignment, null pointer or unknown enum)";
    case ESR_EWORKSPACE: return "workspace too small (see esr_*_workspace_bytes)";
    case ESR_ECUDA: return "CUDA error (see esr_last_cuda_error)";
    case ESR_ENOTSUP: return "not supported by this build";
    case ESR_ENOMEM: return "host allocation failed";
    default: return "unknown error code";
  }
}

extern "C" const char* esr_last_cuda_error(void) { return esr::g_cuda_err; }

extern "C" int esr_device_info(int* sm, int* cc_major, int* cc_minor) {
  int dev = 0;
  ESR_CUDA(cudaGetDevice(&dev));
  int v = 0;
  if (sm) {
    ESR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    *sm = v;
  }
  if (cc_major) {
    ESR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    *cc_major = v;
  }
  if (cc_minor) {
    ESR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    *cc_minor = v;
  }
  return ESR_OK;
}
