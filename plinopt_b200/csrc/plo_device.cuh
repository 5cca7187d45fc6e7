// plo_device.cuh -- shared device/host helpers of the sm_100a kernels:
// error plumbing, Philox4x32-10, the (primary,index) reduction key.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/plinopt_b200.h"

namespace plo {

// ---- error plumbing --------------------------------------------------------
void set_error(const char* fmt, ...);
int check_device();  // PLO_OK or PLO_E_NODEVICE (there is no CPU fallback)

#define PLO_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      plo::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return PLO_E_CUDA;                                                            \
    }                                                                               \
  } while (0)

int sm_count();

// ---- device workspace pool ---------------------------------------------------
// cudaMalloc / cudaFree cost milliseconds each and a sparsifier run creates and destroys one search plan per
// localSparsifier step: device blocks are therefore cached per device in size classes and reused
// (plo_release_workspace() returns them to the driver).  pool_free() synchronises the device first, like cudaFree.
cudaError_t pool_alloc_bytes(void** p, size_t bytes);
void pool_free(void* p);
void quad_release_all();  // lincomb_quad.cu: per-device staging buffers and stream (called by plo_release_workspace)
template <class T>
inline cudaError_t pool_alloc(T** p, size_t bytes) { return pool_alloc_bytes(reinterpret_cast<void**>(p), bytes); }

// ---- Philox4x32-10 (Salmon et al. 2011); key = seed, counter = (index, block) ----
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---- reduction key: lexicographic minimum of (primary, index) ---------------
struct Key {
  unsigned long long primary;  // measure 0: nnz<<32 | nno ; measure 3: bits of the (positive) double
  unsigned long long index;
};
__host__ __device__ __forceinline__ bool key_less(const Key& a, const Key& b) {
  return a.primary < b.primary || (a.primary == b.primary && a.index < b.index);
}

#ifdef __CUDACC__
__device__ __forceinline__ Key warp_min(Key k) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    Key o;
    o.primary = __shfl_xor_sync(0xffffffffu, k.primary, d);
    o.index = __shfl_xor_sync(0xffffffffu, k.index, d);
    if (key_less(o, k)) k = o;
  }
  return k;
}
// Block-wide minimum; result valid in thread 0.  `sh` holds >= 32 Keys.
__device__ __forceinline__ Key block_min(Key k, Key* sh) {
  k = warp_min(k);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) sh[wid] = k;
  __syncthreads();
  if (wid == 0) {
    Key v;
    v.primary = ~0ull; v.index = ~0ull;
    if (lane < nw) v = sh[lane];
    k = warp_min(v);
  }
  return k;
}
#endif

}  // namespace plo
