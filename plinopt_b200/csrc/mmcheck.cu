// mmcheck.cu -- batched probabilistic matrix-multiplication check mod p on sm_100a.
//
// Replaces PLinOpt::MMchecker  include/plinopt_library.inl:472-558  for a batch
// of independent random evaluations (the reference does one per call):
//   va = L.ua, vb = R.ub (:504-505), vc = va o vb (:507), wc = P.vc (:509),
//   compare with the direct product reshape(ua).reshape(ub)  (:513-528).
//
// Design (B200): the three sparse products are one warp-specialised kernel,
// mm_slab_spmm_kernel.  Samples are grouped by 32 (lane = sample) and every
// vector is stored group-major  V[g][j][lane], so the slice of X that a sparse
// product needs for one sample group and one slab of 1024 columns is ONE
// contiguous 128 KB block: it is brought into shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier) and stays there while the CTA streams matrix rows
// against it.  The matrix itself is pre-cut on the host into self-contained
// "chunk blobs" (row offsets + (column byte offset, value) pairs of a run of
// rows inside one slab, ~equal non-zero counts); a producer warp streams the
// blobs through a 4-stage shared-memory ring with bulk copies, 16 consumer
// warps take rows from the current blob through a shared counter (rows are
// 16..1024 entries long) and do, per entry, one broadcast LDS.128 per two
// entries + one LDS + one IMAD.WIDE.U32 with carry-out (+ half an IADD3.X):
// exact 96-bit accumulation, one Barrett reduction mod p per output.
// Bound: shared-memory bandwidth (one 128 B wavefront per multiply-add per
// warp) -- DESIGN.md section 5.3.
#include <algorithm>
#include <vector>

#include "plo_device.cuh"

namespace plo {

constexpr int kSlabCols = 1024;                         // columns of X resident per CTA
constexpr int kSlabBytes = kSlabCols * 32 * 4;          // 128 KB
constexpr int kStages = 4;                              // blob ring depth
constexpr int kChunkEnt = 2560;                         // max padded entries per blob
constexpr int kChunkRows = 255;                         // max rows per blob
constexpr int kHdrBytes = (kChunkRows + 1) * 4;         // 1 KB of row offsets
constexpr int kStageBytes = kHdrBytes + kChunkEnt * 8;  // 21.5 KB
constexpr int kConsumerWarps = 24;
constexpr int kSpThreads = (kConsumerWarps + 1) * 32;   // + one producer warp
constexpr int kZeroColBytes = 128;                      // slab column 1024: 32 zero words (padding target of the grouped format)
constexpr int kRingBase = kSlabBytes + kZeroColBytes;
constexpr int kSpSmem = kRingBase + kStages * kStageBytes;
constexpr int kGroupedMinAvg = 6;                       // value-grouped format when a (row, value) group averages >= this many entries

struct ChunkDesc {
  int slab, row0, nrows, bytes;  // blob = [nrows+1 offsets, padded to 16 B][stream]; stream = (colbyte, val) pairs (plain format)
                                 // or, per (row, value) group, {val, nwords} + nwords x 4 columns, two per 32-bit word (grouped format)
  unsigned long long off;        // byte offset of the blob
  unsigned long long pad_;
};

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
__device__ __forceinline__ void add96(Acc96& a, const Acc96& b) {
  asm volatile(
      "add.cc.u32 %0, %0, %3;\n\t"
      "addc.cc.u32 %1, %1, %4;\n\t"
      "addc.u32 %2, %2, %5;"
      : "+r"(a.a0), "+r"(a.a1), "+r"(a.a2)
      : "r"(b.a0), "r"(b.a1), "r"(b.a2));
}
// x mod p for any 64-bit x: Barrett with M = floor((2^64-1)/p); q is at most 2 short.
__device__ __forceinline__ unsigned int barrett64(unsigned long long x, unsigned int p, unsigned long long M) {
  const unsigned long long q = __umul64hi(x, M);
  unsigned long long r = x - q * p;
  if (r >= p) r -= p;
  if (r >= p) r -= p;
  return (unsigned int)r;
}
__device__ __forceinline__ unsigned int reduce96(const Acc96& a, unsigned int p, unsigned long long M) {
  const unsigned int t = barrett64(((unsigned long long)a.a2 << 32) | a.a1, p, M);
  return barrett64(((unsigned long long)t << 32) | a.a0, p, M);
}

// ---- mbarrier / TMA bulk copy (PTX ISA 8.x, sm_90+) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ unsigned int lds_u32(uint32_t addr) {
  unsigned int v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ua[g][j][lane], ub[g][j][lane] = Philox words mod p; counter = (sample, j / 4, which), key = seed.
// One thread per (sample, quad of coordinates); lanes = samples, so the stores are coalesced.
__global__ void mm_gen_kernel(unsigned int p, unsigned long long seed, unsigned long long first_sample, int batch, int len_a,
                              int len_b, unsigned int* __restrict__ ua, unsigned int* __restrict__ ub) {
  const int qa = (len_a + 3) >> 2, qb = (len_b + 3) >> 2;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)batch * (qa + qb)) return;
  const int b = (int)(e % batch);
  int q = (int)(e / batch);
  const int which = q >= qa;
  if (which) q -= qa;
  const int len = which ? len_b : len_a;
  const unsigned long long s = first_sample + (unsigned long long)b;
  uint32_t w[4];
  philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)q, (uint32_t)which, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  unsigned int* dst = (which ? ub : ua) + ((size_t)(b >> 5) * len + (size_t)q * 4) * 32 + (b & 31);
  for (int t = 0; t < 4 && q * 4 + t < len; ++t) dst[(size_t)t * 32] = w[t] % p;
}

// [batch][len] (caller layout) -> [g][len][lane], reduced mod p
__global__ void mm_transpose_kernel(unsigned int p, int batch, int len, const unsigned int* __restrict__ src, unsigned int* __restrict__ dst) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)batch * len) return;
  const int j = (int)(e % len), b = (int)(e / len);
  dst[((size_t)(b >> 5) * len + j) * 32 + (b & 31)] = src[e] % p;
}

struct SpmmArgs {
  unsigned int p;
  unsigned long long M;  // floor((2^64-1)/p)
  int rows, xlen, groups, nchunks;
  const ChunkDesc* chunk;
  const unsigned char* blob;
  const unsigned int* X;    // [groups][xlen][32]
  unsigned int* out;        // [nslabs][groups][rows][32]
  const unsigned int* mul;  // optional [groups][rows][32]: out = (A X) o mul  (fused Hadamard step, single-slab matrices)
};

__device__ __forceinline__ void addwide(unsigned long long& s, unsigned int x, unsigned int one) {
  asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s) : "r"(x), "r"(one));  // 64-bit accumulate on the fma pipe
}

// Grouped format: a 32-bit word packs two slab columns, one in bits [10:0], one in bits [31:21].
// Shared-memory address of this lane's X word = column * 128 + xl:
//   low field : (w & 0x7ff) * 128 + xl            (LOP3 + IMAD)
//   high field: hi32(w * 2^18) + xl = (w >> 14) + xl, clean because bits [20:11] of w are zero  (one IMAD.HI)
__device__ __forceinline__ uint32_t col_lo_addr(unsigned int w, uint32_t xl) {
  uint32_t t, a;
  asm volatile("and.b32 %0, %1, 0x7ff;" : "=r"(t) : "r"(w));
  asm volatile("mad.lo.u32 %0, %1, 128, %2;" : "=r"(a) : "r"(t), "r"(xl));
  return a;
}
__device__ __forceinline__ uint32_t col_hi_addr(unsigned int w, uint32_t xl) {
  uint32_t a;
  asm volatile("mad.hi.u32 %0, %1, 262144, %2;" : "=r"(a) : "r"(w), "r"(xl));
  return a;
}

template <bool GROUPED>
__global__ void __launch_bounds__(kSpThreads, 1) mm_slab_spmm_kernel(const SpmmArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kStages], empty[kStages], slabbar;
  __shared__ int cnt[kStages];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long T = (long long)a.groups * a.nchunks;
  const long long i0 = (long long)blockIdx.x * T / gridDim.x, i1 = (long long)(blockIdx.x + 1) * T / gridDim.x;
  const int nitems = (int)(i1 - i0);
  if (tid < 32) reinterpret_cast<unsigned int*>(smem + kSlabBytes)[tid] = 0u;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    mbar_init(&slabbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int pg = -1, ps = -1;
  if (warp == kConsumerWarps) {
    // ---- producer: one thread feeds the slab and the blob ring ----
    if (lane != 0) return;
    for (int k = 0; k < nitems; ++k) {
      const long long i = i0 + k;
      const int g = (int)(i / a.nchunks), c = (int)(i - (long long)g * a.nchunks);
      const ChunkDesc ch = a.chunk[c];
      const int stage = k & (kStages - 1);
      if (g != pg || ch.slab != ps) {
        // every consumer must be done with the old slab: drain the ring
        for (int j = k > kStages ? k - kStages : 0; j < k; ++j) mbar_wait(&empty[j & (kStages - 1)], (j / kStages) & 1);
        const int ncols = min(kSlabCols, a.xlen - ch.slab * kSlabCols);
        const unsigned quarter = (unsigned)ncols * 32u;
        const unsigned char* src = (const unsigned char*)(a.X + ((size_t)g * a.xlen + (size_t)ch.slab * kSlabCols) * 32);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&slabbar, quarter * 4u);
        for (int q = 0; q < 4; ++q) bulk_g2s(smem + q * quarter, src + (size_t)q * quarter, quarter, &slabbar);
        pg = g; ps = ch.slab;
      } else if (k >= kStages) {
        mbar_wait(&empty[stage], ((k / kStages) - 1) & 1);
      }
      cnt[stage] = 0;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&full[stage], (unsigned)ch.bytes);
      bulk_g2s(smem + kRingBase + stage * kStageBytes, a.blob + ch.off, (unsigned)ch.bytes, &full[stage]);
    }
    return;
  }
  // ---- consumers: lane = sample of the group, warp takes rows of the current blob ----
  const uint32_t xl = smem_u32(smem) + lane * 4;  // this lane's word of slab column 0
  unsigned int one;
  asm volatile("mov.u32 %0, 1;" : "=r"(one));
  unsigned slabphase = 0;
  for (int k = 0; k < nitems; ++k) {
    const long long i = i0 + k;
    const int g = (int)(i / a.nchunks), c = (int)(i - (long long)g * a.nchunks);
    const int4 ch = *reinterpret_cast<const int4*>(a.chunk + c);  // slab, row0, nrows, bytes
    const int stage = k & (kStages - 1);
    if (g != pg || ch.x != ps) { mbar_wait(&slabbar, slabphase); slabphase ^= 1; pg = g; ps = ch.x; }
    mbar_wait(&full[stage], (k / kStages) & 1);
    const unsigned char* sb = smem + kRingBase + stage * kStageBytes;
    const unsigned int* hdr = reinterpret_cast<const unsigned int*>(sb);
    const uint4* ent = reinterpret_cast<const uint4*>(sb + (((ch.z + 1) * 4 + 15) & ~15));
    const size_t obase = (((size_t)ch.x * a.groups + g) * a.rows + ch.y) * 32 + lane;
    unsigned int* outp = a.out + obase;
    const unsigned int* mulp = a.mul ? a.mul + obase : nullptr;
    int cur = 0;
    if (lane == 0) cur = atomicAdd(&cnt[stage], 1);
    cur = __shfl_sync(0xffffffffu, cur, 0);
    while (cur < ch.z) {
      int nxt = 0;
      if (lane == 0) nxt = atomicAdd(&cnt[stage], 1);
      const unsigned o0 = hdr[cur], o1 = hdr[cur + 1];
      Acc96 A, B;
      A.a0 = A.a1 = A.a2 = 0;
      B.a0 = B.a1 = B.a2 = 0;
      if (GROUPED) {
        // stream of 8-byte words: {val, nwords} then nwords words of 4 columns; out += val * sum_cols X[col]
        const uint2* st = reinterpret_cast<const uint2*>(ent);
        unsigned w = o0;
        while (w < o1) {
          const uint2 h = st[w];
          const uint2* cw = st + w + 1;
          unsigned long long s0 = 0, s1 = 0;
#pragma unroll 2
          for (unsigned j = 0; j < h.y; ++j) {
            const uint2 c = cw[j];
            const unsigned x0 = lds_u32(col_lo_addr(c.x, xl));
            const unsigned x1 = lds_u32(col_hi_addr(c.x, xl));
            const unsigned x2 = lds_u32(col_lo_addr(c.y, xl));
            const unsigned x3 = lds_u32(col_hi_addr(c.y, xl));
            addwide(s0, x0, one);
            addwide(s1, x1, one);
            addwide(s0, x2, one);
            addwide(s1, x3, one);
          }
          s0 += s1;  // < 2^45: at most 2^13 terms below 2^32
          mac96(A, h.x, (unsigned)s0);
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(B.a1), "+r"(B.a2) : "r"(h.x), "r"((unsigned)(s0 >> 32)));
          w += h.y + 1;
        }
        B.a0 = 0;
      } else {
        const uint4* e = ent + (o0 >> 1);
        const int n = (int)((o1 - o0) >> 1);  // uint4 = 2 entries; rows are padded to 4 entries
#pragma unroll 2
        for (int t = 0; t < n; t += 2) {
          const uint4 q0 = e[t], q1 = e[t + 1];
          const unsigned x0 = lds_u32(xl + q0.x);
          const unsigned x1 = lds_u32(xl + q0.z);
          const unsigned x2 = lds_u32(xl + q1.x);
          const unsigned x3 = lds_u32(xl + q1.z);
          mac96(A, q0.y, x0);
          mac96(B, q0.w, x1);
          mac96(A, q1.y, x2);
          mac96(B, q1.w, x3);
        }
      }
      add96(A, B);
      unsigned int res = reduce96(A, a.p, a.M);
      if (mulp) res = barrett64((unsigned long long)res * mulp[(size_t)cur * 32], a.p, a.M);
      outp[(size_t)cur * 32] = res;
      cur = __shfl_sync(0xffffffffu, nxt, 0);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }
}

// vc[e] = (sum_parts va) * (sum_parts vb) mod p   (all in the group-major layout)
__global__ void mm_hadamard_kernel(unsigned int p, unsigned long long M, size_t count, int partsA, int partsB, const unsigned int* __restrict__ va,
                                   const unsigned int* __restrict__ vb, unsigned int* __restrict__ vc) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  unsigned long long x = 0, y = 0;
  for (int q = 0; q < partsA; ++q) x += va[(size_t)q * count + e];
  for (int q = 0; q < partsB; ++q) y += vb[(size_t)q * count + e];
  vc[e] = barrett64((unsigned long long)barrett64(x, p, M) * barrett64(y, p, M), p, M);
}

// bad[b] |= (sum_parts wc[g][o][lane] != sum_t ua[g][i*k+t][lane] * ub[g][t*n+j][lane])  for o = i*n + j, b = 32 g + lane
__global__ void mm_verify_kernel(unsigned int p, unsigned long long M, int m, int k, int n, int batch, int groups, int parts,
                                 const unsigned int* __restrict__ wc, const unsigned int* __restrict__ ua, const unsigned int* __restrict__ ub,
                                 unsigned int* __restrict__ bad) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int mn = m * n;
  if (e >= (size_t)groups * mn * 32) return;
  const int lane = (int)(e & 31);
  const int o = (int)((e >> 5) % mn), g = (int)((e >> 5) / mn), i = o / n, j = o % n;
  const int b = g * 32 + lane;
  if (b >= batch) return;
  unsigned long long w = 0;
  for (int q = 0; q < parts; ++q) w += wc[(size_t)q * groups * mn * 32 + e];
  Acc96 acc;
  acc.a0 = acc.a1 = acc.a2 = 0;
  const unsigned int* pa = ua + ((size_t)g * m * k + (size_t)i * k) * 32 + lane;
  const unsigned int* pb = ub + ((size_t)g * k * n + j) * 32 + lane;
  for (int t = 0; t < k; ++t) mac96(acc, pa[(size_t)t * 32], pb[(size_t)t * n * 32]);
  if (barrett64(w, p, M) != reduce96(acc, p, M)) atomicOr(bad + b, 1u);
}

}  // namespace plo

using namespace plo;

// Device form of one sparse matrix: chunk blobs + chunk table (see the header comment).
struct DevSlabCsr {
  int rows, cols, nslabs, nchunks;
  bool grouped;
  long long nnz;
  ChunkDesc* chunk;
  unsigned char* blob;
};

struct plo_mmcheck_plan {
  uint32_t p;
  unsigned long long M;
  int m, k, n, r, batch, groups, grid_cap;
  DevSlabCsr L, R, P;
  unsigned int *ua, *ub, *va, *vb, *vc, *wc, *bad, *stage;
};

static bool csr_valid(const plo_csr* c, uint32_t p) {
  if (!c || !c->ptr || c->rows < 1 || c->cols < 1 || c->ptr[0] != 0) return false;
  for (int i = 0; i < c->rows; ++i) if (c->ptr[i + 1] < c->ptr[i]) return false;
  const long long nnz = c->ptr[c->rows];
  if (nnz && (!c->col || !c->val)) return false;
  for (long long t = 0; t < nnz; ++t) if (c->col[t] < 0 || c->col[t] >= c->cols || c->val[t] >= p) return false;
  return true;
}

// Cuts the CSR into slabs of kSlabCols columns and, inside a slab, into runs of rows of about equal
// cost; serialises every run as a blob and uploads blobs + table.  Format: value-grouped when the
// matrix has few distinct values per row (HM matrices do: 25-31 distinct values in 32x32x32_15096),
// plain (column, value) pairs otherwise.
static int build_slab_csr(const plo_csr* h, int groups, DevSlabCsr* d) {
  d->rows = h->rows; d->cols = h->cols; d->nnz = h->ptr[h->rows];
  d->nslabs = (h->cols + kSlabCols - 1) / kSlabCols;
  d->chunk = nullptr; d->blob = nullptr; d->nchunks = 0; d->grouped = false;
  const int rows = h->rows, nslabs = d->nslabs;
  // entries bucketed by (slab, row), each bucket sorted by (value, column)
  std::vector<long long> start((size_t)nslabs * rows + 1, 0);
  for (int i = 0; i < rows; ++i)
    for (long long t = h->ptr[i]; t < h->ptr[i + 1]; ++t) ++start[(size_t)(h->col[t] / kSlabCols) * rows + i + 1];
  for (size_t q = 1; q < start.size(); ++q) start[q] += start[q - 1];
  std::vector<uint2> sorted((size_t)d->nnz);  // (local column, value)
  {
    std::vector<long long> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < rows; ++i)
      for (long long t = h->ptr[i]; t < h->ptr[i + 1]; ++t) {
        const int s = h->col[t] / kSlabCols;
        sorted[(size_t)fill[(size_t)s * rows + i]++] = make_uint2((unsigned)(h->col[t] - s * kSlabCols), h->val[t]);
      }
  }
  long long ngroups = 0;
  for (size_t q = 0; q + 1 < start.size(); ++q) {
    std::sort(sorted.begin() + start[q], sorted.begin() + start[q + 1], [](const uint2& x, const uint2& y) { return x.y != y.y ? x.y < y.y : x.x < y.x; });
    for (long long t = start[q]; t < start[q + 1]; ++t) ngroups += (t == start[q] || sorted[(size_t)t].y != sorted[(size_t)t - 1].y);
  }
  const bool grouped = ngroups > 0 && d->nnz >= (long long)kGroupedMinAvg * ngroups;
  d->grouped = grouped;
  // per bucket: stream length in 8-byte words and cost in multiply-add slots
  std::vector<unsigned> words(start.size() - 1), cost(start.size() - 1);
  long long total_cost = 0;
  for (size_t q = 0; q + 1 < start.size(); ++q) {
    long long w = 0, c = 0;
    if (grouped) {
      long long t = start[q];
      while (t < start[q + 1]) {
        long long u = t;
        while (u < start[q + 1] && sorted[(size_t)u].y == sorted[(size_t)t].y) ++u;
        const long long nw = (u - t + 3) / 4;
        w += 1 + nw; c += 4 * nw + 4;
        t = u;
      }
    } else {
      w = (start[q + 1] - start[q] + 3) & ~3ll; c = w;
    }
    if (w > kChunkEnt - 1) { set_error("mmcheck: a row has too many entries inside one %d-column slab (duplicate columns?)", kSlabCols); return PLO_E_ARG; }
    words[q] = (unsigned)w; cost[q] = (unsigned)c; total_cost += c;
  }
  long long target = total_cost * groups / ((long long)sm_count() * 16);
  if (target < 64) target = 64;
  if (target > 4 * kChunkEnt) target = 4 * kChunkEnt;
  std::vector<ChunkDesc> table;
  std::vector<unsigned char> blob;
  for (int s = 0; s < nslabs; ++s) {
    int row = 0;
    while (row < rows) {
      int nrows = 0;
      long long w = 0, c = 0;
      while (row + nrows < rows && nrows < kChunkRows) {
        const size_t q = (size_t)s * rows + row + nrows;
        if (nrows > 0 && (c + cost[q] > target || w + words[q] > kChunkEnt - 1)) break;
        w += words[q]; c += cost[q]; ++nrows;
      }
      const size_t hdr = (((size_t)nrows + 1) * 4 + 15) & ~(size_t)15;
      ChunkDesc ch;
      ch.slab = s; ch.row0 = row; ch.nrows = nrows; ch.bytes = (int)((hdr + (size_t)w * 8 + 15) & ~(size_t)15);  // bulk copies move multiples of 16 B
      ch.off = blob.size(); ch.pad_ = 0;
      blob.resize(blob.size() + (size_t)ch.bytes, 0);
      unsigned int* ho = reinterpret_cast<unsigned int*>(blob.data() + ch.off);
      uint2* eo = reinterpret_cast<uint2*>(blob.data() + ch.off + hdr);
      unsigned o = 0;
      for (int qq = 0; qq < nrows; ++qq) {
        ho[qq] = o;
        const size_t q = (size_t)s * rows + row + qq;
        if (grouped) {
          long long t = start[q];
          while (t < start[q + 1]) {
            long long u = t;
            while (u < start[q + 1] && sorted[(size_t)u].y == sorted[(size_t)t].y) ++u;
            const unsigned nw = (unsigned)((u - t + 3) / 4);
            eo[o++] = make_uint2(sorted[(size_t)t].y, nw);
            for (unsigned j = 0; j < nw; ++j) {
              unsigned c4[4];
              for (int z = 0; z < 4; ++z) c4[z] = t + 4 * j + z < u ? sorted[(size_t)(t + 4 * j + z)].x : (unsigned)kSlabCols;  // padding -> zero column
              eo[o++] = make_uint2(c4[0] | (c4[1] << 21), c4[2] | (c4[3] << 21));
            }
            t = u;
          }
        } else {
          for (long long t = start[q]; t < start[q + 1]; ++t) eo[o++] = make_uint2(sorted[(size_t)t].x * 128u, sorted[(size_t)t].y);
          while (o & 3) eo[o++] = make_uint2(0u, 0u);  // padding: 0 * X[slab column 0]
        }
      }
      ho[nrows] = o;
      table.push_back(ch);
      row += nrows;
    }
  }
  d->nchunks = (int)table.size();
  PLO_CUDA(pool_alloc(&d->chunk, sizeof(ChunkDesc) * table.size()));
  PLO_CUDA(pool_alloc(&d->blob, blob.size() ? blob.size() : 16));
  PLO_CUDA(cudaMemcpy(d->chunk, table.data(), sizeof(ChunkDesc) * table.size(), cudaMemcpyHostToDevice));
  PLO_CUDA(cudaMemcpy(d->blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  return PLO_OK;
}
static void free_slab_csr(DevSlabCsr* d) { pool_free(d->chunk); pool_free(d->blob); }

extern "C" {

void plo_mmcheck_plan_destroy(plo_mmcheck_plan* pl) {
  if (!pl) return;
  free_slab_csr(&pl->L); free_slab_csr(&pl->R); free_slab_csr(&pl->P);
  pool_free(pl->ua); pool_free(pl->ub); pool_free(pl->va); pool_free(pl->vb); pool_free(pl->vc); pool_free(pl->wc); pool_free(pl->bad); pool_free(pl->stage);
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
  PLO_CUDA(cudaFuncSetAttribute(mm_slab_spmm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem));
  PLO_CUDA(cudaFuncSetAttribute(mm_slab_spmm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem));
  plo_mmcheck_plan* pl = new plo_mmcheck_plan();
  memset(pl, 0, sizeof(*pl));
  pl->p = p; pl->M = ~0ull / p; pl->m = m; pl->k = k; pl->n = n; pl->r = r; pl->batch = batch;
  pl->groups = (batch + 31) / 32;
  pl->grid_cap = sm_count();
  rc = build_slab_csr(L, pl->groups, &pl->L);
  if (!rc) rc = build_slab_csr(R, pl->groups, &pl->R);
  if (!rc) rc = build_slab_csr(P, pl->groups, &pl->P);
  if (rc) { plo_mmcheck_plan_destroy(pl); return rc; }
  const size_t G32 = (size_t)pl->groups * 32;
  const size_t stage = (size_t)batch * (size_t)(m * k > k * n ? m * k : k * n);
  auto zalloc = [](unsigned int** ptr, size_t words) { return pool_alloc(ptr, 4 * words) == cudaSuccess && cudaMemset(*ptr, 0, 4 * words) == cudaSuccess; };
  const bool ok = zalloc(&pl->ua, G32 * m * k) && zalloc(&pl->ub, G32 * k * n) && zalloc(&pl->va, G32 * r * pl->L.nslabs) &&
                  zalloc(&pl->vb, G32 * r * pl->R.nslabs) && zalloc(&pl->vc, G32 * r) && zalloc(&pl->wc, G32 * m * n * pl->P.nslabs) &&
                  zalloc(&pl->bad, (size_t)batch) && zalloc(&pl->stage, stage);
  if (!ok) { set_error("mmcheck: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())); plo_mmcheck_plan_destroy(pl); return PLO_E_CUDA; }
  *plan = pl;
  return PLO_OK;
}

static void launch_spmm(const plo_mmcheck_plan* pl, const DevSlabCsr& A, const unsigned int* X, unsigned int* out, const unsigned int* mul, cudaStream_t st) {
  SpmmArgs a;
  a.mul = mul;
  a.p = pl->p; a.M = pl->M; a.rows = A.rows; a.xlen = A.cols; a.groups = pl->groups; a.nchunks = A.nchunks;
  a.chunk = A.chunk; a.blob = A.blob; a.X = X; a.out = out;
  const long long T = (long long)pl->groups * A.nchunks;
  const int grid = (int)std::min<long long>(T, pl->grid_cap);
  if (A.grouped) mm_slab_spmm_kernel<true><<<grid, kSpThreads, kSpSmem, st>>>(a);
  else mm_slab_spmm_kernel<false><<<grid, kSpThreads, kSpSmem, st>>>(a);
}

static int mm_pipeline(plo_mmcheck_plan* pl, cudaStream_t st) {
  const int B = pl->batch;
  const size_t G32 = (size_t)pl->groups * 32;
  PLO_CUDA(cudaMemsetAsync(pl->bad, 0, 4 * (size_t)B, st));
  launch_spmm(pl, pl->L, pl->ua, pl->va, nullptr, st);
  if (pl->L.nslabs == 1 && pl->R.nslabs == 1) {
    launch_spmm(pl, pl->R, pl->ub, pl->vc, pl->va, st);  // vc = (R ub) o (L ua) in the epilogue
  } else {
    launch_spmm(pl, pl->R, pl->ub, pl->vb, nullptr, st);
    const size_t cnt = G32 * pl->r;
    mm_hadamard_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(pl->p, pl->M, cnt, pl->L.nslabs, pl->R.nslabs, pl->va, pl->vb, pl->vc);
  }
  launch_spmm(pl, pl->P, pl->vc, pl->wc, nullptr, st);
  const size_t vcnt = G32 * pl->m * pl->n;
  mm_verify_kernel<<<(unsigned)((vcnt + 255) / 256), 256, 0, st>>>(pl->p, pl->M, pl->m, pl->k, pl->n, B, pl->groups, pl->P.nslabs, pl->wc, pl->ua, pl->ub, pl->bad);
  PLO_CUDA(cudaGetLastError());
  return PLO_OK;
}

int plo_mmcheck_plan_run(plo_mmcheck_plan* pl, uint64_t seed, uint64_t first_sample, void* stream) {
  if (!pl) { set_error("plo_mmcheck_plan_run: null plan"); return PLO_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t gthreads = (size_t)pl->batch * ((pl->m * pl->k + 3) / 4 + (pl->k * pl->n + 3) / 4);
  mm_gen_kernel<<<(unsigned)((gthreads + 255) / 256), 256, 0, st>>>(pl->p, seed, first_sample, pl->batch, pl->m * pl->k, pl->k * pl->n, pl->ua, pl->ub);
  return mm_pipeline(pl, st);
}

int plo_mmcheck_plan_launches(const plo_mmcheck_plan* pl) { return pl && pl->L.nslabs == 1 && pl->R.nslabs == 1 ? 5 : 6; }

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
