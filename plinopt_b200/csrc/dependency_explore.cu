// dependency_explore.cu -- the `Explore` enumeration of src/dependency.cpp:73-100 on sm_100a
// (SURVEY.md section 8 row f3).
//
// Reference: for every start row i (coefficient 1), add up to level-1 further rows
// q_1 < q_2 < ... (all > i), each with a coefficient from the list C, and report every
// combination W = M[i] + sum_t C[v_t].M[q_t] that is the zero vector (a linear dependency
// between the outputs, :84) or has exactly one non-zero coordinate (an input expressed by
// the outputs, :86-90).  Hits are printed in depth-first order.
//
// Formulation (exact, no multiplications on the device, same idea as lincomb_search.cu):
// the host tabulates prod[q][v][.] = C[v].M[q][.] once.  The combinations with d added rows
// are enumerated as (prefix, last term): a thread owns a prefix (i, q_1..q_{d-1}, v_1..v_{d-1}),
// keeps nb_j = -(M[i] + sum_{t<d} prod[q_t][v_t])_j in registers and, for every last term
// (q_d, v_d), counts the coordinates with prod[q_d][v_d][j] == nb_j  (one compare per
// coordinate per candidate).  Prefixes are grouped by their last row l, so that all threads
// of a block sweep the same tile prod[l+1..r-1][.][.], staged once in shared memory and read
// warp-uniformly (LDS.128 broadcasts).  Candidates with >= n-1 matching coordinates are rare:
// they are compacted through one atomic counter into a hit list, which the host sorts into
// the reference's depth-first order.
#include <algorithm>
#include <vector>

#include "plo_device.cuh"

namespace plo {

constexpr int kDepThreads = 128;
constexpr int kDepMaxDepth = 4;           // added rows per combination (level <= 5)
constexpr int kDepPrefixPerBlock = 2048;  // prefixes swept per block against one staged tile

struct DepBlock {
  unsigned long long first;  // first prefix id of the block
  unsigned int count;        // prefixes in the block
  int last;                  // last row of every prefix of the block
};
struct DepRawHit {
  unsigned long long prefix;
  unsigned int x;  // (q_d - last - 1) * c + v_d
  int pos;         // -1: zero vector, else the one non-zero coordinate
};
template <typename T>
struct DepParams {
  int r, n, c, d;
  unsigned int p;
  unsigned long long cpow;  // c^(d-1)
  const T* base;            // [r][NPAD]
  const T* prod;            // [r][c][NPAD]
  const unsigned char* combos;  // [ncombo][4]: i, q_1, .., q_{d-1} (d <= 4)
  const DepBlock* blocks;
  unsigned long long* counter;
  DepRawHit* hits;
  unsigned long long max_hits;
};

template <typename T, int NPAD, bool MODP>
__global__ void __launch_bounds__(kDepThreads) dependency_kernel(const DepParams<T> prm) {
  extern __shared__ __align__(16) unsigned char dep_smem[];
  T* tile = reinterpret_cast<T*>(dep_smem);
  constexpr int VEC = 16 / sizeof(T);
  const DepBlock blk = prm.blocks[blockIdx.x];
  const int c = prm.c, n = prm.n;
  const int nlast = (prm.r - blk.last - 1) * c;  // last terms (q_d, v_d) of this group
  {
    const uint4* src = reinterpret_cast<const uint4*>(prm.prod + (size_t)(blk.last + 1) * c * NPAD);
    uint4* dst = reinterpret_cast<uint4*>(tile);
    const int nvec = nlast * NPAD / VEC;
    for (int e = threadIdx.x; e < nvec; e += kDepThreads) dst[e] = src[e];
  }
  __syncthreads();
  const T SENT = (T)(~(T)0) >> (MODP ? 0 : 1);  // never a residue / beyond the integer bound
  for (unsigned int pi = threadIdx.x; pi < blk.count; pi += kDepThreads) {
    const unsigned long long id = blk.first + pi;
    const unsigned long long combo = id / prm.cpow;
    unsigned long long vc = id - combo * prm.cpow;
    const uchar4 rows = reinterpret_cast<const uchar4*>(prm.combos)[combo];
    const int q[4] = {rows.x, rows.y, rows.z, rows.w};
    T nb[NPAD];
#pragma unroll
    for (int e = 0; e < NPAD; ++e) nb[e] = prm.base[(size_t)q[0] * NPAD + e];
#pragma unroll
    for (int t = kDepMaxDepth - 1; t >= 1; --t) {  // v_{d-1} is the fastest digit of the prefix code
      if (t >= prm.d) continue;
      const int v = (int)(vc % (unsigned)c);
      vc /= (unsigned)c;
      const T* pr = prm.prod + ((size_t)q[t] * c + v) * NPAD;
#pragma unroll
      for (int e = 0; e < NPAD; ++e) {
        if (MODP) { const unsigned long long s = (unsigned long long)nb[e] + pr[e]; nb[e] = (T)(s >= prm.p ? s - prm.p : s); }
        else nb[e] += pr[e];
      }
    }
#pragma unroll
    for (int e = 0; e < NPAD; ++e) {
      if (MODP) nb[e] = nb[e] ? (T)(prm.p - nb[e]) : (T)0;
      else nb[e] = (T)0 - nb[e];
      if (e >= n) nb[e] = SENT;
    }
    for (int x = 0; x < nlast; ++x) {
      const T* row = tile + (size_t)x * NPAD;
      int z = 0;
#pragma unroll
      for (int e4 = 0; e4 < NPAD / VEC; ++e4) {
        const uint4 u = reinterpret_cast<const uint4*>(row)[e4];
        if (sizeof(T) == 4) {
          z += (nb[e4 * 4 + 0] == (T)u.x);
          z += (nb[e4 * 4 + 1] == (T)u.y);
          z += (nb[e4 * 4 + 2] == (T)u.z);
          z += (nb[e4 * 4 + 3] == (T)u.w);
        } else {
          z += (nb[e4 * 2 + 0] == (T)(((unsigned long long)u.y << 32) | u.x));
          z += (nb[e4 * 2 + 1] == (T)(((unsigned long long)u.w << 32) | u.z));
        }
      }
      if (z >= n - 1) {  // rare: zero vector or exactly one non-zero coordinate
        int pos = -1;
#pragma unroll
        for (int e = 0; e < NPAD; ++e)  // compile-time indices keep nb[] in registers
          if (e < n && !(row[e] == nb[e])) pos = e;
        const unsigned long long slot = atomicAdd(prm.counter, 1ull);
        if (slot < prm.max_hits) {
          DepRawHit h;
          h.prefix = id; h.x = (unsigned)x; h.pos = pos;
          prm.hits[slot] = h;
        }
      }
    }
  }
}

template <typename T, bool MODP>
static cudaError_t dep_launch(int npad, int grid, size_t smem, cudaStream_t st, const DepParams<T>& prm) {
#define PLO_DEP_CASE(NP)                                                                                    \
  case NP: {                                                                                                \
    auto kern = dependency_kernel<T, NP, MODP>;                                                             \
    if (smem > 48 * 1024) {                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
      if (e != cudaSuccess) return e;                                                                       \
    }                                                                                                       \
    kern<<<grid, kDepThreads, smem, st>>>(prm);                                                             \
    break;                                                                                                  \
  }
  switch (npad) {
    PLO_DEP_CASE(4) PLO_DEP_CASE(8) PLO_DEP_CASE(12) PLO_DEP_CASE(16) PLO_DEP_CASE(20) PLO_DEP_CASE(24) PLO_DEP_CASE(28) PLO_DEP_CASE(32)
    PLO_DEP_CASE(40) PLO_DEP_CASE(48) PLO_DEP_CASE(64)
    default: return cudaErrorInvalidValue;
  }
#undef PLO_DEP_CASE
  return cudaGetLastError();
}

static int dep_npad(int n) {
  static const int sizes[] = {4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 64};
  for (int s : sizes) if (n <= s) return s;
  return 0;
}

template <typename T, bool MODP>
static int dep_run(uint32_t p, int r, int n, int c, int level, const int64_t* base, const int64_t* prod, uint64_t max_hits,
                   plo_dep_hit* out, uint64_t* nhits, uint64_t* ncand) {
  const int npad = dep_npad(n);
  std::vector<T> hbase((size_t)r * npad, (T)0), hprod((size_t)r * c * npad, (T)0);
  for (int i = 0; i < r; ++i) for (int j = 0; j < n; ++j) hbase[(size_t)i * npad + j] = (T)base[(size_t)i * n + j];
  for (size_t qv = 0; qv < (size_t)r * c; ++qv) for (int j = 0; j < n; ++j) hprod[qv * npad + j] = (T)prod[qv * n + j];
  T *d_base = nullptr, *d_prod = nullptr;
  unsigned long long* d_counter = nullptr;
  DepRawHit* d_hits = nullptr;
  unsigned char* d_combos = nullptr;
  DepBlock* d_blocks = nullptr;
  const uint64_t cap = max_hits ? max_hits : 1;
  auto cleanup = [&]() { pool_free(d_base); pool_free(d_prod); pool_free(d_counter); pool_free(d_hits); pool_free(d_combos); pool_free(d_blocks); };
  if (pool_alloc(&d_base, hbase.size() * sizeof(T)) != cudaSuccess || pool_alloc(&d_prod, hprod.size() * sizeof(T)) != cudaSuccess ||
      pool_alloc(&d_counter, 8) != cudaSuccess || pool_alloc(&d_hits, cap * sizeof(DepRawHit)) != cudaSuccess ||
      cudaMemcpy(d_base, hbase.data(), hbase.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_prod, hprod.data(), hprod.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("plo_dependency_explore: %s", cudaGetErrorString(cudaGetLastError()));
    cleanup();
    return PLO_E_CUDA;
  }
  uint64_t total_hits = 0, total_cand = 0, stored = 0;
  for (int d = 1; d <= level - 1; ++d) {
    // prefixes = combinations i < q_1 < .. < q_{d-1}, grouped by their last row
    std::vector<std::vector<unsigned char>> by_last((size_t)r);
    {
      std::vector<int> idx(d);
      for (int t = 0; t < d; ++t) idx[t] = t;
      while (d <= r) {
        unsigned char rec[4] = {255, 255, 255, 255};
        for (int t = 0; t < d; ++t) rec[t] = (unsigned char)idx[t];
        by_last[(size_t)idx[d - 1]].insert(by_last[(size_t)idx[d - 1]].end(), rec, rec + 4);
        int t = d - 1;
        while (t >= 0 && idx[t] == r - d + t) --t;
        if (t < 0) break;
        ++idx[t];
        for (int u = t + 1; u < d; ++u) idx[u] = idx[u - 1] + 1;
      }
    }
    unsigned long long cpow = 1;
    for (int t = 1; t < d; ++t) cpow *= (unsigned)c;
    std::vector<unsigned char> combos;
    std::vector<DepBlock> blocks;
    for (int l = 0; l + 1 < r; ++l) {  // l = r-1 has no last term left
      const size_t ncombo = by_last[(size_t)l].size() / 4;
      if (!ncombo) continue;
      const unsigned long long first_combo = combos.size() / 4;
      combos.insert(combos.end(), by_last[(size_t)l].begin(), by_last[(size_t)l].end());
      const unsigned long long nprefix = ncombo * cpow;
      total_cand += nprefix * (unsigned long long)(r - l - 1) * (unsigned)c;
      for (unsigned long long o = 0; o < nprefix; o += kDepPrefixPerBlock) {
        DepBlock b;
        b.first = first_combo * cpow + o;
        b.count = (unsigned)std::min<unsigned long long>(kDepPrefixPerBlock, nprefix - o);
        b.last = l;
        blocks.push_back(b);
      }
    }
    if (blocks.empty()) continue;
    pool_free(d_combos); pool_free(d_blocks);
    d_combos = nullptr; d_blocks = nullptr;
    if (pool_alloc(&d_combos, combos.size()) != cudaSuccess || pool_alloc(&d_blocks, blocks.size() * sizeof(DepBlock)) != cudaSuccess ||
        cudaMemcpy(d_combos, combos.data(), combos.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_blocks, blocks.data(), blocks.size() * sizeof(DepBlock), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(d_counter, 0, 8) != cudaSuccess) {
      set_error("plo_dependency_explore: %s", cudaGetErrorString(cudaGetLastError()));
      cleanup();
      return PLO_E_CUDA;
    }
    DepParams<T> prm;
    prm.r = r; prm.n = n; prm.c = c; prm.d = d; prm.p = p; prm.cpow = cpow; prm.base = d_base; prm.prod = d_prod;
    prm.combos = d_combos; prm.blocks = d_blocks; prm.counter = d_counter; prm.hits = d_hits; prm.max_hits = cap;
    const size_t smem = (size_t)(r - 1) * c * npad * sizeof(T);
    cudaError_t e = dep_launch<T, MODP>(npad, (int)blocks.size(), smem, nullptr, prm);
    unsigned long long found = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&found, d_counter, 8, cudaMemcpyDeviceToHost);
    std::vector<DepRawHit> raw((size_t)std::min<unsigned long long>(found, cap));
    if (e == cudaSuccess && !raw.empty()) e = cudaMemcpy(raw.data(), d_hits, raw.size() * sizeof(DepRawHit), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("plo_dependency_explore: %s", cudaGetErrorString(e)); cleanup(); return PLO_E_CUDA; }
    total_hits += found;
    for (const DepRawHit& h : raw) {
      if (stored >= max_hits) break;
      plo_dep_hit o;
      memset(&o, 0, sizeof(o));
      o.depth = d; o.pos = h.pos;
      const unsigned long long combo = h.prefix / cpow;
      unsigned long long vc = h.prefix - combo * cpow;
      for (int t = 0; t < 5; ++t) { o.rows[t] = -1; o.coefs[t] = -1; }
      int last = 0;
      for (int t = 0; t < d; ++t) { o.rows[t] = combos[combo * 4 + t]; last = o.rows[t]; }
      for (int t = d - 1; t >= 1; --t) { o.coefs[t] = (int)(vc % (unsigned)c); vc /= (unsigned)c; }
      o.rows[d] = last + 1 + (int)(h.x / (unsigned)c);
      o.coefs[d] = (int)(h.x % (unsigned)c);
      out[stored++] = o;
    }
  }
  cleanup();
  // the reference's depth-first order: (i, q_1, v_1, q_2, v_2, ..), a combination before its extensions
  std::sort(out, out + stored, [](const plo_dep_hit& a, const plo_dep_hit& b) {
    if (a.rows[0] != b.rows[0]) return a.rows[0] < b.rows[0];
    for (int t = 1; t < 5; ++t) {
      if (a.rows[t] != b.rows[t]) return a.rows[t] < b.rows[t];
      if (a.coefs[t] != b.coefs[t]) return a.coefs[t] < b.coefs[t];
    }
    return false;
  });
  if (nhits) *nhits = total_hits;
  if (ncand) *ncand = total_cand;
  if (total_hits > max_hits) {
    // The stored subset was filled depth by depth in arrival order: it is NOT the first max_hits lines of the reference's
    // depth-first output.  Same protocol as plo_orbit_plan_survivors: report the total and let the caller come back with room.
    set_error("plo_dependency_explore: %llu hits, room for %llu", total_hits, (unsigned long long)max_hits);
    return PLO_E_RANGE;
  }
  return PLO_OK;
}

}  // namespace plo

using namespace plo;

extern "C" int plo_dependency_explore(uint32_t p, int r, int n, int c, int level, const int64_t* base, const int64_t* prod,
                                      uint64_t max_hits, plo_dep_hit* hits, uint64_t* nhits, uint64_t* ncand) {
  if (!base || !prod || (max_hits && !hits) || r < 1 || r > 255 || n < 1 || c < 1 || level < 1) {
    set_error("plo_dependency_explore: bad argument (need 1 <= r <= 255, n >= 1, c >= 1, level >= 1)");
    return PLO_E_ARG;
  }
  if (level - 1 > kDepMaxDepth) { set_error("plo_dependency_explore: level %d not supported (at most %d)", level, kDepMaxDepth + 1); return PLO_E_SHAPE; }
  const int npad = dep_npad(n);
  if (!npad) { set_error("plo_dependency_explore: more than 64 columns are not supported"); return PLO_E_SHAPE; }
  const size_t esize = p ? 4 : 8;
  if ((size_t)(r - 1) * c * npad * esize > 200 * 1024) { set_error("plo_dependency_explore: the product table (%d x %d x %d) does not fit in shared memory", r, c, n); return PLO_E_SHAPE; }
  if (p) {
    for (size_t e = 0; e < (size_t)r * n; ++e) if (base[e] < 0 || base[e] >= (int64_t)p) { set_error("plo_dependency_explore: residue out of range"); return PLO_E_ARG; }
    for (size_t e = 0; e < (size_t)r * c * n; ++e) if (prod[e] < 0 || prod[e] >= (int64_t)p) { set_error("plo_dependency_explore: residue out of range"); return PLO_E_ARG; }
  } else {
    const int64_t lim = INT64_MAX / 8;  // sums of at most 5 terms stay exact and below the sentinel
    for (size_t e = 0; e < (size_t)r * n; ++e) if (base[e] > lim || base[e] < -lim) { set_error("plo_dependency_explore: entry too large"); return PLO_E_RANGE; }
    for (size_t e = 0; e < (size_t)r * c * n; ++e) if (prod[e] > lim || prod[e] < -lim) { set_error("plo_dependency_explore: entry too large"); return PLO_E_RANGE; }
  }
  int rc = check_device();
  if (rc) return rc;
  if (p) return dep_run<unsigned int, true>(p, r, n, c, level, base, prod, max_hits, hits, nhits, ncand);
  return dep_run<unsigned long long, false>(0, r, n, c, level, base, prod, max_hits, hits, nhits, ncand);  // two's complement: equality is sign-agnostic
}
