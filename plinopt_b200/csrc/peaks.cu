// peaks.cu -- roofline denominators MEASURED_PEAKS.json does not hold, and the
// dense growth-factor kernel.
//
// plo_measure_peaks: register-resident, fully unrolled dependent-chain-free
// loops on every SM (CUDA events, best of `reps`): IMAD (fma pipe), DFMA (fp64
// pipe) and an ISETP+IADD pair (alu pipe; the inner operation of the sparsifier
// kernel).  Same methodology as MEASURED_PEAKS.json ("best of N, CUDA events").
//
// plo_growth_G2: src/growthfactor.cpp:117-125 on explicit dense triples.
#include <vector>

#include "plo_device.cuh"

namespace plo {

constexpr int kPeakThreads = 256;
constexpr int kPeakIlp = 8;
constexpr int kPeakIters = 4096;

__global__ void __launch_bounds__(kPeakThreads) peak_imad_kernel(int* out, int a, int b) {
  int x[kPeakIlp];
#pragma unroll
  for (int i = 0; i < kPeakIlp; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < kPeakIlp; ++i) x[i] = x[i] * a + b;
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < kPeakIlp; ++i) s += x[i];
  if (s == 0x7fffffff) out[0] = s;
}

__global__ void __launch_bounds__(kPeakThreads) peak_dfma_kernel(double* out, double a, double b) {
  double x[kPeakIlp];
#pragma unroll
  for (int i = 0; i < kPeakIlp; ++i) x[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < kPeakIlp; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < kPeakIlp; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
}

// compare + conditional increment: two alu-pipe instructions per element
__global__ void __launch_bounds__(kPeakThreads) peak_ialu_kernel(int* out, unsigned int a, unsigned int b) {
  unsigned int x[kPeakIlp], cnt[kPeakIlp];
#pragma unroll
  for (int i = 0; i < kPeakIlp; ++i) { x[i] = threadIdx.x * 7u + i; cnt[i] = 0; }
  for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < kPeakIlp; ++i) {
        cnt[i] += (x[i] == a + u);   // ISETP + IADD
        x[i] ^= cnt[i] + b;          // keeps the chain data dependent (LOP3 + IADD)
      }
  }
  unsigned int s = 0;
#pragma unroll
  for (int i = 0; i < kPeakIlp; ++i) s += cnt[i] + x[i];
  if (s == 0x7fffffffu) out[0] = (int)s;
}

// Issue-rate peak: independent IMAD (fma-heavy pipe) and LOP3 (alu pipe) chains interleaved one to one; each pipe takes a warp
// instruction every other cycle, so together they fill the scheduler's one instruction per cycle.
__global__ void __launch_bounds__(kPeakThreads) peak_issue_kernel(int* out, int a, int b) {
  int x[kPeakIlp];
  unsigned int y[kPeakIlp];
#pragma unroll
  for (int i = 0; i < kPeakIlp; ++i) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3u + i; }
  for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < kPeakIlp; ++i) {
        x[i] = x[i] * a + b;                             // IMAD
        y[i] = (y[i] & (unsigned)a) ^ (unsigned)(b + u);  // LOP3
      }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < kPeakIlp; ++i) s += x[i] + (int)y[i];
  if (s == 0x7fffffff) out[0] = s;
}

// One block per triple, one thread per row i; s = sum_i |L_i| |R_i| |Pt_i| (row order kept: serial final sum)
__global__ void growth_g2_kernel(int r, int a, int b, int c, const double* __restrict__ L, const double* __restrict__ R,
                                 const double* __restrict__ P, double* __restrict__ out) {
  extern __shared__ double terms[];
  const int t = blockIdx.x;
  const double* Lt = L + (size_t)t * r * a;
  const double* Rt = R + (size_t)t * r * b;
  const double* Pt = P + (size_t)t * c * r;
  for (int i = threadIdx.x; i < r; i += blockDim.x) {
    double sl = 0, sr = 0, sp = 0;
    for (int j = 0; j < a; ++j) { const double x = Lt[(size_t)i * a + j]; sl = __dadd_rn(sl, __dmul_rn(x, x)); }
    for (int j = 0; j < b; ++j) { const double x = Rt[(size_t)i * b + j]; sr = __dadd_rn(sr, __dmul_rn(x, x)); }
    for (int j = 0; j < c; ++j) { const double x = Pt[(size_t)j * r + i]; sp = __dadd_rn(sp, __dmul_rn(x, x)); }
    terms[i] = __dmul_rn(__dmul_rn(sqrt(sl), sqrt(sr)), sqrt(sp));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < r; ++i) s = __dadd_rn(s, terms[i]);
    out[t] = s;
  }
}

}  // namespace plo

using namespace plo;

extern "C" {

int plo_measure_peaks(int reps, double* imad_per_s, double* dfma_per_s, double* ialu_per_s) {
  int rc = check_device();
  if (rc) return rc;
  if (reps < 1) reps = 1;
  const int grid = sm_count() * 8;
  int* d_i = nullptr;
  double* d_d = nullptr;
  PLO_CUDA(cudaMalloc(&d_i, 64));
  PLO_CUDA(cudaMalloc(&d_d, 64));
  cudaEvent_t e0, e1;
  PLO_CUDA(cudaEventCreate(&e0));
  PLO_CUDA(cudaEventCreate(&e1));
  const double ops = (double)grid * kPeakThreads * (double)kPeakIters * 8.0 * kPeakIlp;
  double best[3] = {0, 0, 0};
  for (int which = 0; which < 3; ++which) {
    for (int rep = 0; rep < reps + 1; ++rep) {  // first repetition is the warm-up
      cudaEventRecord(e0);
      if (which == 0) peak_imad_kernel<<<grid, kPeakThreads>>>(d_i, 3, 1);
      else if (which == 1) peak_dfma_kernel<<<grid, kPeakThreads>>>(d_d, 1.0000001, 1e-9);
      else peak_ialu_kernel<<<grid, kPeakThreads>>>(d_i, 12345u, 1u);
      cudaEventRecord(e1);
      cudaError_t e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) { set_error("peak kernel: %s", cudaGetErrorString(e)); cudaFree(d_i); cudaFree(d_d); return PLO_E_CUDA; }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double rate = ops / (ms * 1e-3);
      if (rep > 0 && rate > best[which]) best[which] = rate;
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_i); cudaFree(d_d);
  if (imad_per_s) *imad_per_s = best[0];
  if (dfma_per_s) *dfma_per_s = best[1];
  if (ialu_per_s) *ialu_per_s = best[2];  // compare+increment PAIRS per second
  return PLO_OK;
}

int plo_measure_issue_peak(int reps, double* inst_per_s) {
  int rc = check_device();
  if (rc) return rc;
  if (!inst_per_s) { set_error("plo_measure_issue_peak: null output"); return PLO_E_ARG; }
  if (reps < 1) reps = 1;
  const int grid = sm_count() * 8;
  int* d_i = nullptr;
  PLO_CUDA(cudaMalloc(&d_i, 64));
  cudaEvent_t e0, e1;
  PLO_CUDA(cudaEventCreate(&e0));
  PLO_CUDA(cudaEventCreate(&e1));
  const double inst = (double)grid * kPeakThreads * (double)kPeakIters * 4.0 * kPeakIlp * 2.0;  // thread instructions
  double best = 0;
  for (int rep = 0; rep < reps + 1; ++rep) {  // first repetition is the warm-up
    cudaEventRecord(e0);
    peak_issue_kernel<<<grid, kPeakThreads>>>(d_i, 0x7ffffff3, 1);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { set_error("peak kernel: %s", cudaGetErrorString(e)); cudaFree(d_i); return PLO_E_CUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rate = inst / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_i);
  *inst_per_s = best;
  return PLO_OK;
}

int plo_growth_G2(int batch, int r, int a, int b, int c, const double* L, const double* R, const double* P, double* out) {
  if (batch < 1 || r < 1 || a < 1 || b < 1 || c < 1 || !L || !R || !P || !out) { set_error("plo_growth_G2: bad argument"); return PLO_E_ARG; }
  int rc = check_device();
  if (rc) return rc;
  double *dL = nullptr, *dR = nullptr, *dP = nullptr, *dO = nullptr;
  const size_t nl = (size_t)batch * r * a, nr = (size_t)batch * r * b, np = (size_t)batch * c * r;
  auto fail = [&](const char* what) { set_error("plo_growth_G2: %s: %s", what, cudaGetErrorString(cudaGetLastError())); cudaFree(dL); cudaFree(dR); cudaFree(dP); cudaFree(dO); return PLO_E_CUDA; };
  if (cudaMalloc(&dL, nl * 8) != cudaSuccess || cudaMalloc(&dR, nr * 8) != cudaSuccess || cudaMalloc(&dP, np * 8) != cudaSuccess || cudaMalloc(&dO, (size_t)batch * 8) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMemcpy(dL, L, nl * 8, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dR, R, nr * 8, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dP, P, np * 8, cudaMemcpyHostToDevice) != cudaSuccess) return fail("H2D copy");
  const int threads = r < 256 ? ((r + 31) / 32) * 32 : 256;
  growth_g2_kernel<<<batch, threads, (size_t)r * 8>>>(r, a, b, c, dL, dR, dP, dO);
  if (cudaMemcpy(out, dO, (size_t)batch * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return fail("kernel / D2H copy");
  cudaFree(dL); cudaFree(dR); cudaFree(dP); cudaFree(dO);
  return PLO_OK;
}

}  // extern "C"
