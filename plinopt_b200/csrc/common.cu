// common.cu -- error plumbing, device queries, version.
#include <stdarg.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

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

// ---- device workspace pool (see plo_device.cuh) ----
namespace {
struct Pool {
  std::mutex mu;
  std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks;  // (device, class bytes) -> cached blocks
  std::unordered_map<void*, std::pair<int, size_t>> owner;           // live and cached blocks
  size_t cached_bytes = 0;
};
Pool& pool() { static Pool p; return p; }
size_t size_class(size_t bytes) {
  if (bytes < 512) return 512;
  if (bytes > (64u << 20)) return (bytes + (16u << 20) - 1) / (16u << 20) * (16u << 20);
  size_t c = 512;
  while (c < bytes) c <<= 1;
  return c;
}
constexpr size_t kPoolCap = 8ull << 30;  // bytes kept cached per process
}  // namespace

cudaError_t pool_alloc_bytes(void** p, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const size_t cls = size_class(bytes);
  Pool& P = pool();
  {
    std::lock_guard<std::mutex> g(P.mu);
    auto it = P.free_blocks.find({dev, cls});
    if (it != P.free_blocks.end() && !it->second.empty()) {
      *p = it->second.back();
      it->second.pop_back();
      P.cached_bytes -= cls;
      return cudaSuccess;
    }
  }
  e = cudaMalloc(p, cls);
  if (e != cudaSuccess) {  // give the cache back and retry once
    cudaGetLastError();
    plo_release_workspace();
    e = cudaMalloc(p, cls);
    if (e != cudaSuccess) return e;
  }
  std::lock_guard<std::mutex> g(P.mu);
  P.owner[*p] = {dev, cls};
  return cudaSuccess;
}

void pool_free(void* p) {
  if (!p) return;
  cudaDeviceSynchronize();  // nothing in flight may still use the block when it is handed out again
  Pool& P = pool();
  std::lock_guard<std::mutex> g(P.mu);
  auto it = P.owner.find(p);
  if (it == P.owner.end()) { cudaFree(p); return; }
  if (P.cached_bytes + it->second.second > kPoolCap) { cudaFree(p); P.owner.erase(it); return; }
  P.free_blocks[it->second].push_back(p);
  P.cached_bytes += it->second.second;
}

}  // namespace plo

extern "C" {

void plo_release_workspace(void) {
  plo::quad_release_all();
  plo::Pool& P = plo::pool();
  std::lock_guard<std::mutex> g(P.mu);
  for (auto& kv : P.free_blocks)
    for (void* b : kv.second) { cudaFree(b); P.owner.erase(b); }
  P.free_blocks.clear();
  P.cached_bytes = 0;
}

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
