// mmcheck.cu -- batched probabilistic matrix-multiplication check mod p on sm_100a.
//
// Replaces PLinOpt::MMchecker  include/plinopt_library.inl:472-558  for a batch
// of independent random evaluations (the reference does one per call):
//   va = L.ua, vb = R.ub (:504-505), vc = va o vb (:507), wc = P.vc (:509),
//   compare with the direct product reshape(ua).reshape(ub)  (:513-528).
// Layout: sample-minor vectors X[j][b] so that the 32 lanes of a warp (32
// samples of one matrix row) read one coalesced 128 B line per CSR entry while
// (col,val) are warp-uniform.  Products accumulate exactly in 96 bits
// (mad.lo.cc / madc.hi.cc / addc) and are reduced mod p once per output.
#include <vector>

#include "plo_device.cuh"

namespace plo {

constexpr int kMmThreads = 128;

struct Acc96 {
  unsigned int a0, a1, a2;
};
__device__ __forceinline__ void mac96(Acc96& a, unsigned int x, unsigned int y) {
  asm volatile(
      "mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(a.a0), "+r"(a.a1), "+r"(a.a2)
      : "r"(x), "r"(y));
}
__device__ __forceinline__ unsigned int reduce96(const Acc96& a, unsigned int p) {
  unsigned long long r = a.a2 % p;
  r = ((r << 32) | a.a1) % p;
  r = ((r << 32) | a.a0) % p;
  return (unsigned int)r;
}

// ua[j][b], ub[j][b] = Philox words mod p; counter = (sample, j / 4, which), key = seed.
__global__ void mm_gen_kernel(unsigned int p, unsigned long long seed, unsigned long long first_sample, int batch, int len_a,
                              int len_b, unsigned int* __restrict__ ua, unsigned int* __restrict__ ub) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const unsigned long long s = first_sample + (unsigned long long)b;
  for (int which = 0; which < 2; ++which) {
    const int len = which ? len_b : len_a;
    unsigned int* dst = which ? ub : ua;
    for (int j0 = 0; j0 < len; j0 += 4) {
      uint32_t w[4];
      philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)(j0 >> 2), (uint32_t)which, (uint32_t)seed, (uint32_t)(seed >> 32), w);
      for (int t = 0; t < 4 && j0 + t < len; ++t) dst[(size_t)(j0 + t) * batch + b] = w[t] % p;
    }
  }
}

// [batch][len] (caller layout) -> [len][batch], reduced mod p
__global__ void mm_transpose_kernel(unsigned int p, int batch, int len, const unsigned int* __restrict__ src, unsigned int* __restrict__ dst) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)batch * len) return;
  const int b = (int)(e % batch), j = (int)(e / batch);
  dst[e] = src[(size_t)b * len + j] % p;
}

// Partial sparse product: out[part][row][b] = sum over the part-th slice of row's entries of val * X[col][b]  (mod p)
__global__ void __launch_bounds__(kMmThreads) mm_spmm_kernel(unsigned int p, int rows, int batch, int parts, const long long* __restrict__ ptr,
                                                             const int* __restrict__ col, const unsigned int* __restrict__ val,
                                                             const unsigned int* __restrict__ X, unsigned int* __restrict__ out) {
  const int groups = (batch + 31) >> 5;
  const long long warp = ((long long)blockIdx.x * kMmThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)rows * parts * groups;
  if (warp >= total) return;
  const int g = (int)(warp % groups);
  const long long rp = warp / groups;
  const int part = (int)(rp % parts);
  const int row = (int)(rp / parts);
  const int b = g * 32 + lane;
  const long long beg = ptr[row], end = ptr[row + 1];
  const long long len = end - beg, chunk = (len + parts - 1) / parts;
  const long long lo = beg + chunk * part, hi = (lo + chunk < end) ? lo + chunk : end;
  Acc96 acc;
  acc.a0 = acc.a1 = acc.a2 = 0;
  if (b < batch) {
    long long t = lo;
    for (; t + 4 <= hi; t += 4) {  // 4 independent loads in flight
      const int c0 = col[t], c1 = col[t + 1], c2 = col[t + 2], c3 = col[t + 3];
      const unsigned int v0 = val[t], v1 = val[t + 1], v2 = val[t + 2], v3 = val[t + 3];
      const unsigned int x0 = X[(size_t)c0 * batch + b], x1 = X[(size_t)c1 * batch + b], x2 = X[(size_t)c2 * batch + b], x3 = X[(size_t)c3 * batch + b];
      mac96(acc, v0, x0); mac96(acc, v1, x1); mac96(acc, v2, x2); mac96(acc, v3, x3);
    }
    for (; t < hi; ++t) mac96(acc, val[t], X[(size_t)col[t] * batch + b]);
    out[((size_t)part * rows + row) * batch + b] = reduce96(acc, p);
  }
}

// vc[i][b] = (sum_parts va) * (sum_parts vb) mod p
__global__ void mm_hadamard_kernel(unsigned int p, int r, int batch, int parts, const unsigned int* __restrict__ va,
                                   const unsigned int* __restrict__ vb, unsigned int* __restrict__ vc) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)r * batch) return;
  unsigned long long a = 0, b = 0;
  for (int q = 0; q < parts; ++q) { a += va[(size_t)q * r * batch + e]; b += vb[(size_t)q * r * batch + e]; }
  vc[e] = (unsigned int)(((a % p) * (b % p)) % p);
}

// ok[b] &= (sum_parts wc[o][b] == sum_t ua[i*k+t][b] * ub[t*n+j][b])  for o = i*n + j
__global__ void mm_verify_kernel(unsigned int p, int m, int k, int n, int batch, int parts, const unsigned int* __restrict__ wc,
                                 const unsigned int* __restrict__ ua, const unsigned int* __restrict__ ub, unsigned int* __restrict__ bad) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)m * n * batch) return;
  const int b = (int)(e % batch);
  const int o = (int)(e / batch), i = o / n, j = o % n;
  unsigned long long w = 0;
  for (int q = 0; q < parts; ++q) w += wc[(size_t)q * m * n * batch + e];
  Acc96 acc;
  acc.a0 = acc.a1 = acc.a2 = 0;
  for (int t = 0; t < k; ++t) mac96(acc, ua[(size_t)(i * k + t) * batch + b], ub[(size_t)(t * n + j) * batch + b]);
  if ((unsigned int)(w % p) != reduce96(acc, p)) atomicOr(bad + b, 1u);
}

}  // namespace plo

using namespace plo;

struct DevCsr {
  int rows, cols;
  long long nnz;
  long long* ptr;
  int* col;
  unsigned int* val;
};

struct plo_mmcheck_plan {
  uint32_t p;
  int m, k, n, r, batch;
  DevCsr L, R, P;
  int partsLR, partsP;
  unsigned int *ua, *ub, *va, *vb, *vc, *wc, *bad, *stage;
};

static int upload_csr(const plo_csr* h, DevCsr* d) {
  d->rows = h->rows; d->cols = h->cols; d->nnz = h->ptr[h->rows];
  d->ptr = nullptr; d->col = nullptr; d->val = nullptr;
  PLO_CUDA(cudaMalloc(&d->ptr, sizeof(long long) * (h->rows + 1)));
  PLO_CUDA(cudaMalloc(&d->col, sizeof(int) * (d->nnz ? d->nnz : 1)));
  PLO_CUDA(cudaMalloc(&d->val, sizeof(unsigned int) * (d->nnz ? d->nnz : 1)));
  PLO_CUDA(cudaMemcpy(d->ptr, h->ptr, sizeof(long long) * (h->rows + 1), cudaMemcpyHostToDevice));
  PLO_CUDA(cudaMemcpy(d->col, h->col, sizeof(int) * d->nnz, cudaMemcpyHostToDevice));
  PLO_CUDA(cudaMemcpy(d->val, h->val, sizeof(unsigned int) * d->nnz, cudaMemcpyHostToDevice));
  return PLO_OK;
}
static void free_csr(DevCsr* d) { cudaFree(d->ptr); cudaFree(d->col); cudaFree(d->val); }

static bool csr_valid(const plo_csr* c, uint32_t p) {
  if (!c || !c->ptr || c->rows < 1 || c->cols < 1 || c->ptr[0] != 0) return false;
  for (int i = 0; i < c->rows; ++i) if (c->ptr[i + 1] < c->ptr[i]) return false;
  const long long nnz = c->ptr[c->rows];
  if (nnz && (!c->col || !c->val)) return false;
  for (long long t = 0; t < nnz; ++t) if (c->col[t] < 0 || c->col[t] >= c->cols || c->val[t] >= p) return false;
  return true;
}

static int pick_parts(const DevCsr& A, int batch) {
  // enough warps to fill the machine: rows * parts * ceil(batch/32) >= ~8 warps per SM-quadrant slot
  const long long groups = (batch + 31) / 32;
  const long long want = (long long)sm_count() * 32;
  long long parts = 1;
  const long long avg = A.rows ? A.nnz / A.rows : 0;
  while (A.rows * parts * groups < want && avg / (parts * 2) >= 32 && parts < 64) parts *= 2;
  return (int)parts;
}

extern "C" {

void plo_mmcheck_plan_destroy(plo_mmcheck_plan* pl) {
  if (!pl) return;
  free_csr(&pl->L); free_csr(&pl->R); free_csr(&pl->P);
  cudaFree(pl->ua); cudaFree(pl->ub); cudaFree(pl->va); cudaFree(pl->vb); cudaFree(pl->vc); cudaFree(pl->wc); cudaFree(pl->bad); cudaFree(pl->stage);
  delete pl;
}

int plo_mmcheck_plan_create(plo_mmcheck_plan** plan, uint32_t p, int m, int k, int n, int r, const plo_csr* L,
                            const plo_csr* R, const plo_csr* P, int batch) {
  if (!plan || p < 2 || m < 1 || k < 1 || n < 1 || r < 1 || batch < 1 || !csr_valid(L, p) || !csr_valid(R, p) || !csr_valid(P, p)) {
    set_error("plo_mmcheck_plan_create: bad argument (null, p < 2, malformed CSR or residue >= p)");
    return PLO_E_ARG;
  }
  if (L->rows != r || R->rows != r || P->cols != r) { set_error("mmcheck: inner dimension mismatch"); return 2; }   // MMchecker.cpp:65-71
  if (L->cols != m * k || R->cols != k * n || P->rows != m * n) { set_error("mmcheck: outer dimension mismatch"); return 3; }  // library.inl:487-495
  int rc = check_device();
  if (rc) return rc;
  plo_mmcheck_plan* pl = new plo_mmcheck_plan();
  memset(pl, 0, sizeof(*pl));
  pl->p = p; pl->m = m; pl->k = k; pl->n = n; pl->r = r; pl->batch = batch;
  rc = upload_csr(L, &pl->L);
  if (!rc) rc = upload_csr(R, &pl->R);
  if (!rc) rc = upload_csr(P, &pl->P);
  if (rc) { plo_mmcheck_plan_destroy(pl); return rc; }
  pl->partsLR = pick_parts(pl->L, batch);
  { int q = pick_parts(pl->R, batch); if (q > pl->partsLR) pl->partsLR = q; }
  pl->partsP = pick_parts(pl->P, batch);
  const size_t B = (size_t)batch;
  const size_t stage = B * (size_t)(m * k > k * n ? m * k : k * n);
  bool ok = cudaMalloc(&pl->ua, 4 * B * m * k) == cudaSuccess && cudaMalloc(&pl->ub, 4 * B * k * n) == cudaSuccess &&
            cudaMalloc(&pl->va, 4 * B * r * pl->partsLR) == cudaSuccess && cudaMalloc(&pl->vb, 4 * B * r * pl->partsLR) == cudaSuccess &&
            cudaMalloc(&pl->vc, 4 * B * r) == cudaSuccess && cudaMalloc(&pl->wc, 4 * B * m * n * pl->partsP) == cudaSuccess &&
            cudaMalloc(&pl->bad, 4 * B) == cudaSuccess && cudaMalloc(&pl->stage, 4 * stage) == cudaSuccess;
  if (!ok) { set_error("mmcheck: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())); plo_mmcheck_plan_destroy(pl); return PLO_E_CUDA; }
  *plan = pl;
  return PLO_OK;
}

static int mm_pipeline(plo_mmcheck_plan* pl, cudaStream_t st) {
  const int B = pl->batch, groups = (B + 31) / 32;
  auto blocks = [](long long warps) { return (unsigned)((warps * 32 + kMmThreads - 1) / kMmThreads); };
  PLO_CUDA(cudaMemsetAsync(pl->bad, 0, 4 * (size_t)B, st));
  mm_spmm_kernel<<<blocks((long long)pl->r * pl->partsLR * groups), kMmThreads, 0, st>>>(pl->p, pl->r, B, pl->partsLR, pl->L.ptr, pl->L.col, pl->L.val, pl->ua, pl->va);
  mm_spmm_kernel<<<blocks((long long)pl->r * pl->partsLR * groups), kMmThreads, 0, st>>>(pl->p, pl->r, B, pl->partsLR, pl->R.ptr, pl->R.col, pl->R.val, pl->ub, pl->vb);
  mm_hadamard_kernel<<<(unsigned)(((size_t)pl->r * B + 255) / 256), 256, 0, st>>>(pl->p, pl->r, B, pl->partsLR, pl->va, pl->vb, pl->vc);
  const int mn = pl->m * pl->n;
  mm_spmm_kernel<<<blocks((long long)mn * pl->partsP * groups), kMmThreads, 0, st>>>(pl->p, mn, B, pl->partsP, pl->P.ptr, pl->P.col, pl->P.val, pl->vc, pl->wc);
  mm_verify_kernel<<<(unsigned)(((size_t)mn * B + 255) / 256), 256, 0, st>>>(pl->p, pl->m, pl->k, pl->n, B, pl->partsP, pl->wc, pl->ua, pl->ub, pl->bad);
  PLO_CUDA(cudaGetLastError());
  return PLO_OK;
}

int plo_mmcheck_plan_run(plo_mmcheck_plan* pl, uint64_t seed, uint64_t first_sample, void* stream) {
  if (!pl) { set_error("plo_mmcheck_plan_run: null plan"); return PLO_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  mm_gen_kernel<<<(pl->batch + 127) / 128, 128, 0, st>>>(pl->p, seed, first_sample, pl->batch, pl->m * pl->k, pl->k * pl->n, pl->ua, pl->ub);
  return mm_pipeline(pl, st);
}

int plo_mmcheck_plan_launches(const plo_mmcheck_plan*) { return 6; }

// Run on caller-provided sample vectors ua (batch x mk), ub (batch x kn), host pointers.
static int mm_run_given(plo_mmcheck_plan* pl, const uint32_t* ua, const uint32_t* ub, cudaStream_t st) {
  const int B = pl->batch, la = pl->m * pl->k, lb = pl->k * pl->n;
  PLO_CUDA(cudaMemcpyAsync(pl->stage, ua, 4 * (size_t)B * la, cudaMemcpyHostToDevice, st));
  mm_transpose_kernel<<<(unsigned)(((size_t)B * la + 255) / 256), 256, 0, st>>>(pl->p, B, la, pl->stage, pl->ua);
  PLO_CUDA(cudaMemcpyAsync(pl->stage, ub, 4 * (size_t)B * lb, cudaMemcpyHostToDevice, st));
  mm_transpose_kernel<<<(unsigned)(((size_t)B * lb + 255) / 256), 256, 0, st>>>(pl->p, B, lb, pl->stage, pl->ub);
  return mm_pipeline(pl, st);
}

int plo_mmcheck_plan_result(plo_mmcheck_plan* pl, void* stream, uint8_t* ok, int* verdict) {
  if (!pl) { set_error("plo_mmcheck_plan_result: null plan"); return PLO_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<unsigned int> bad(pl->batch);
  PLO_CUDA(cudaMemcpyAsync(bad.data(), pl->bad, 4 * (size_t)pl->batch, cudaMemcpyDeviceToHost, st));
  PLO_CUDA(cudaStreamSynchronize(st));
  int v = 0;
  for (int b = 0; b < pl->batch; ++b) { if (ok) ok[b] = bad[b] ? 0 : 1; if (bad[b]) v = 1; }
  if (verdict) *verdict = v;
  return PLO_OK;
}

int plo_mmcheck_batch(uint32_t p, int m, int k, int n, int r, const plo_csr* L, const plo_csr* R, const plo_csr* P,
                      uint64_t seed, int batch, const uint32_t* ua, const uint32_t* ub, uint8_t* ok) {
  plo_mmcheck_plan* pl = nullptr;
  int rc = plo_mmcheck_plan_create(&pl, p, m, k, n, r, L, R, P, batch);
  if (rc) return rc;
  if ((ua == nullptr) != (ub == nullptr)) { set_error("mmcheck: ua and ub must both be given or both be NULL"); plo_mmcheck_plan_destroy(pl); return PLO_E_ARG; }
  rc = ua ? mm_run_given(pl, ua, ub, nullptr) : plo_mmcheck_plan_run(pl, seed, 0, nullptr);
  int verdict = 0;
  if (!rc) rc = plo_mmcheck_plan_result(pl, nullptr, ok, &verdict);
  plo_mmcheck_plan_destroy(pl);
  return rc ? rc : verdict;
}

}  // extern "C"
