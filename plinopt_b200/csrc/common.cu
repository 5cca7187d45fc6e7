// common.cu -- error plumbing, device queries, version.
#include <stdarg.h>

#include "plo_device.cuh"

namespace plo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device visible (%s): this engine has no CPU fallback", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    return PLO_E_NODEVICE;
  }
  return PLO_OK;
}

int sm_count() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

}  // namespace plo

extern "C" {

int plo_version(void) { return 100; }

int plo_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int plo_set_device(int device) {
  int rc = plo::check_device();
  if (rc) return rc;
  PLO_CUDA(cudaSetDevice(device));
  return PLO_OK;
}

const char* plo_last_error(void) { return plo::g_err; }

}  // extern "C"
