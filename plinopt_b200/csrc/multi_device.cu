// multi_device.cu -- the sweeps sharded over the first `ndev` CUDA devices of ONE process (no torch, no NCCL needed: the merge of
// the per-device winners is a handful of words).  Same single-thread scheme as plo_orbit_sweep_devices: every device gets its plan
// and its launches asynchronously, then the results are collected in device order.  Reference analogues: the `omp parallel for`
// + `omp critical` of src/orbiter.cpp:272,298 and include/plinopt_sparsify.inl:962,968 -- here the "threads" are GPUs.
//   plo_lincomb_search_devices   one sparsifier search, prefix range (i*c+j)*c+k split over the devices
//   plo_mmcheck_batch_devices    a batch of MMchecker samples split over the devices
//   plo_factor_sweep_devices     Factorizer random restarts, index range split over the devices
// Multi-PROCESS runs (one rank per GPU) use the plan entry points + one all-reduce instead (plinopt_b200/sharding.py, bench.py).
#include <algorithm>
#include <vector>

#include "plo_device.cuh"

using namespace plo;

namespace {
struct DeviceScope {  // restores the caller's current device
  int prev = 0;
  DeviceScope() { cudaGetDevice(&prev); }
  ~DeviceScope() { cudaSetDevice(prev); }
};
int clamp_devices(int ndev) {
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) { cudaGetLastError(); return 0; }
  return std::max(1, std::min(ndev, have));
}
void split(uint64_t lo, uint64_t hi, int parts, std::vector<uint64_t>& cut) {  // contiguous ascending shards
  const uint64_t total = hi - lo, base = total / (uint64_t)parts, rem = total % (uint64_t)parts;
  cut.assign((size_t)parts + 1, lo);
  for (int d = 0; d < parts; ++d) cut[(size_t)d + 1] = cut[(size_t)d] + base + ((uint64_t)d < rem ? 1 : 0);
}
}  // namespace

extern "C" {

int plo_lincomb_search_devices(int ndev, uint32_t p, int nbatch, int n, int m, const int64_t* TM, int off, int c, const int64_t* coeffs,
                               int nprev, const int64_t* prev_rows, const int* init_rl, const int* init_cl, int* best_rl, int* best_cl,
                               uint64_t* best_index) {
  if (ndev < 1 || nbatch < 1 || !best_rl || !best_cl || !best_index) { set_error("plo_lincomb_search_devices: bad argument"); return PLO_E_ARG; }
  ndev = clamp_devices(ndev);
  if (!ndev) return check_device();
  DeviceScope scope;
  std::vector<uint64_t> cut;
  split(0, (uint64_t)c * c * c, ndev, cut);
  std::vector<plo_lincomb_plan*> plans((size_t)ndev, nullptr);
  int rc = PLO_OK;
  for (int d = 0; d < ndev && !rc; ++d) {
    if (cudaSetDevice(d) != cudaSuccess) { set_error("plo_lincomb_search_devices: cudaSetDevice(%d) failed", d); rc = PLO_E_CUDA; break; }
    rc = plo_lincomb_plan_create(&plans[(size_t)d], p, nbatch, n, m, TM, off, c, coeffs, nprev, prev_rows, init_rl, init_cl);
    if (!rc) rc = plo_lincomb_plan_run_range(plans[(size_t)d], cut[(size_t)d], cut[(size_t)d + 1], nullptr);
  }
  std::vector<int> rl((size_t)nbatch), cl((size_t)nbatch);
  std::vector<uint64_t> idx((size_t)nbatch);
  bool first = true;
  for (int d = 0; d < ndev; ++d) {
    if (!plans[(size_t)d]) continue;
    cudaSetDevice(d);
    if (!rc) rc = plo_lincomb_plan_result(plans[(size_t)d], nullptr, rl.data(), cl.data(), idx.data());
    if (!rc)
      for (int b = 0; b < nbatch; ++b) {
        // maximum of (rl, cl, -index); a shard without a candidate above the seed reports the seed weight with PLO_NO_INDEX
        const bool has = idx[(size_t)b] != PLO_NO_INDEX;
        bool better = first;
        if (!first && has) {
          const bool cur = best_index[b] != PLO_NO_INDEX;
          better = !cur || rl[(size_t)b] > best_rl[b] || (rl[(size_t)b] == best_rl[b] && (cl[(size_t)b] > best_cl[b] || (cl[(size_t)b] == best_cl[b] && idx[(size_t)b] < best_index[b])));
        }
        if (better) { best_rl[b] = rl[(size_t)b]; best_cl[b] = cl[(size_t)b]; best_index[b] = idx[(size_t)b]; }
      }
    first = false;
    plo_lincomb_plan_destroy(plans[(size_t)d]);
  }
  return rc;
}

int plo_mmcheck_batch_devices(int ndev, uint32_t p, int m, int k, int n, int r, const plo_csr* L, const plo_csr* R, const plo_csr* P, uint64_t seed,
                              int batch, uint8_t* ok) {
  if (ndev < 1 || batch < 1 || !ok) { set_error("plo_mmcheck_batch_devices: bad argument"); return PLO_E_ARG; }
  ndev = clamp_devices(ndev);
  if (!ndev) return check_device();
  if (ndev > batch) ndev = batch;
  DeviceScope scope;
  std::vector<uint64_t> cut;
  split(0, (uint64_t)batch, ndev, cut);
  std::vector<plo_mmcheck_plan*> plans((size_t)ndev, nullptr);
  int rc = PLO_OK;
  for (int d = 0; d < ndev && !rc; ++d) {
    if (cudaSetDevice(d) != cudaSuccess) { set_error("plo_mmcheck_batch_devices: cudaSetDevice(%d) failed", d); rc = PLO_E_CUDA; break; }
    rc = plo_mmcheck_plan_create(&plans[(size_t)d], p, m, k, n, r, L, R, P, (int)(cut[(size_t)d + 1] - cut[(size_t)d]));
    if (rc > 0) break;  // 3: outer dimension mismatch -- the same on every device
    if (!rc) rc = plo_mmcheck_plan_run(plans[(size_t)d], seed, cut[(size_t)d], nullptr);
  }
  int verdict = 0;
  for (int d = 0; d < ndev; ++d) {
    if (!plans[(size_t)d]) continue;
    cudaSetDevice(d);
    int v = 0;
    if (!rc) rc = plo_mmcheck_plan_result(plans[(size_t)d], nullptr, ok + cut[(size_t)d], &v);
    if (!rc && v > verdict) verdict = v;
    plo_mmcheck_plan_destroy(plans[(size_t)d]);
  }
  return rc ? rc : verdict;
}

int plo_factor_sweep_devices(int ndev, uint32_t p, int r, int n, int k, const uint32_t* M, uint64_t seed, uint64_t lo, uint64_t hi,
                             plo_factor_best* best) {
  if (ndev < 1 || !best || hi < lo) { set_error("plo_factor_sweep_devices: bad argument"); return PLO_E_ARG; }
  ndev = clamp_devices(ndev);
  if (!ndev) return check_device();
  DeviceScope scope;
  std::vector<uint64_t> cut;
  split(lo, hi, ndev, cut);
  std::vector<plo_factor_plan*> plans((size_t)ndev, nullptr);
  int rc = PLO_OK;
  for (int d = 0; d < ndev && !rc; ++d) {
    if (cudaSetDevice(d) != cudaSuccess) { set_error("plo_factor_sweep_devices: cudaSetDevice(%d) failed", d); rc = PLO_E_CUDA; break; }
    rc = plo_factor_plan_create(&plans[(size_t)d], p, r, n, k, M, seed);
    if (!rc) rc = plo_factor_plan_run(plans[(size_t)d], cut[(size_t)d], cut[(size_t)d + 1], nullptr);
  }
  plo_factor_best win;
  win.nnz_alt = win.nno_alt = win.nnz_cob = 0xFFFFFFFFu; win.pad_ = 0; win.index = PLO_NO_INDEX;
  for (int d = 0; d < ndev; ++d) {
    if (!plans[(size_t)d]) continue;
    cudaSetDevice(d);
    plo_factor_best b;
    if (!rc) rc = plo_factor_plan_result(plans[(size_t)d], nullptr, &b);
    if (!rc && b.index != PLO_NO_INDEX) {
      const bool better = win.index == PLO_NO_INDEX || b.nnz_alt < win.nnz_alt ||
                          (b.nnz_alt == win.nnz_alt && (b.nno_alt < win.nno_alt || (b.nno_alt == win.nno_alt && b.nnz_cob < win.nnz_cob)));  // ascending shards keep ties
      if (better) win = b;
    }
    plo_factor_plan_destroy(plans[(size_t)d]);
  }
  if (!rc) *best = win;
  return rc;
}

}  // extern "C"
