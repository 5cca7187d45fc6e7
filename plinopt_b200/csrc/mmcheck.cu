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
// against it.  The matrix is pre-cut on the host into self-contained "chunk
// blobs" (row descriptors + row streams of a run of rows inside one slab,
// about equal cost); a producer warp streams the blobs through a 4-stage
// shared-memory ring with bulk copies, consumer warps take rows from the
// current blob through a shared counter.  Lane = sample, so a load of X is one
// conflict-free 128 B wavefront; exact 96-bit accumulation, one Barrett
// reduction mod p per output.
//
// What makes it fast on Hopcroft-Musinski matrices is done by the ENCODER
// (host, once per plan; results identical to the plain CSR product, checked by
// plo_mmcheck_encode_check on the CPU and by the parity tests on the device):
//  * column block sums: next to the 1024 columns the slab holds the sums of X
//    over aligned blocks of 4 and of 16 columns (computed by the CTA when the
//    slab lands).  A row whose block mostly carries one value v takes v times
//    the block sum and corrects the other columns of the block; L and R of
//    32x32x32_15096 shrink from 83 to about 10 loads per row.
//  * row block sums (the transposed trick, for P): entries that most rows of a
//    block of 4 or 16 rows {base + t * stride} share become entries of a
//    VIRTUAL row, computed once; mm_verify adds the virtual rows back.
//  * what is left of a row: plain (column, value) pairs, or -- for a value
//    that still occurs often in the row -- an add-only group (exact 64-bit
//    sum of X words, one multiply-add per group).
// Bound: shared-memory bandwidth (one wavefront per X word per warp, plus the
// stream at 8 B per wavefront) and issue slots -- DESIGN.md section 5.3.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "plo_device.cuh"

namespace plo {

constexpr int kSlabCols = 1024;                         // columns of X resident per CTA
constexpr int kBlk4 = kSlabCols / 4, kBlk16 = kSlabCols / 16;
constexpr int kVCols = kSlabCols + kBlk4 + kBlk16;      // + block sums: 1344 virtual columns
constexpr int kSlabBytes = kVCols * 32 * 4;             // 168 KB
constexpr int kStages = 4;                              // blob ring depth
constexpr int kChunkEnt = 1536;                         // max 8-byte stream words per blob
constexpr int kChunkRows = 192;                         // max rows per blob
constexpr int kHdrBytes = kChunkRows * 8;               // row descriptors
constexpr int kStageBytes = kHdrBytes + kChunkEnt * 8;  // 13.5 KB
#ifndef PLO_MM_WARPS
#define PLO_MM_WARPS 28
#endif
constexpr int kConsumerWarps = PLO_MM_WARPS;
constexpr int kSpThreads = (kConsumerWarps + 1) * 32;   // + one producer warp
constexpr int kRingBase = kSlabBytes;
constexpr int kSpSmem = kRingBase + kStages * kStageBytes;
constexpr int kMinGroup = 8;                            // a (row, value) group with fewer entries is stored as plain (column, value) pairs

struct ChunkDesc {
  int slab, row0, nrows, bytes;  // blob = [nrows row descriptors, padded to 16 B][row streams] (format: see mm_slab_spmm_kernel)
  unsigned long long off;        // byte offset of the blob
  unsigned long long pad_;
};

// Block sums: a block of `lv` (4 or 16) columns of a slab = {hi * lv * cs + t * cs + lo : t < lv} for a column stride cs (a power
// of two, lv * cs <= 1024); block number hi * cs + lo.  Row blocks (virtual rows) use the same numbering on the whole row index.
__host__ __device__ __forceinline__ int blk_id(int c, int lv, int cs) { return (c / (lv * cs)) * cs + c % cs; }
__host__ __device__ __forceinline__ int blk_member(int b, int t, int lv, int cs) { return (b / cs) * (lv * cs) + t * cs + b % cs; }

struct Acc96 {
  unsigned long long lo;  // bits 0..63 (one aligned register pair: the multiply-add below becomes a single IMAD.WIDE with carry-out)
  unsigned int hi;        // bits 64..95
};
__device__ __forceinline__ void mac96(Acc96& a, unsigned int x, unsigned int y) {
  asm volatile(
      "{\n\t"
      ".reg .u32 l, h;\n\t"
      "mov.b64 {l, h}, %0;\n\t"
      "mad.lo.cc.u32 l, %2, %3, l;\n\t"
      "madc.hi.cc.u32 h, %2, %3, h;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "mov.b64 %0, {l, h};\n\t"
      "}"
      : "+l"(a.lo), "+r"(a.hi)
      : "r"(x), "r"(y));
}
__device__ __forceinline__ void add96(Acc96& a, const Acc96& b) {
  asm volatile(
      "{\n\t"
      ".reg .u32 l, h, m, n;\n\t"
      "mov.b64 {l, h}, %0;\n\t"
      "mov.b64 {m, n}, %2;\n\t"
      "add.cc.u32 l, l, m;\n\t"
      "addc.cc.u32 h, h, n;\n\t"
      "addc.u32 %1, %1, %3;\n\t"
      "mov.b64 %0, {l, h};\n\t"
      "}"
      : "+l"(a.lo), "+r"(a.hi)
      : "l"(b.lo), "r"(b.hi));
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
  const unsigned int t = barrett64(((unsigned long long)a.hi << 32) | (a.lo >> 32), p, M);
  return barrett64(((unsigned long long)t << 32) | (unsigned int)a.lo, p, M);
}

// Montgomery reduction of the 96-bit accumulator, p odd, pinv = -p^-1 mod 2^32:  A 2^-64 mod p  (two steps of 32 bits; A < 2^80).
// The encoder stores the matrix values times 2^64 mod p (times 2^96 where the product is to come out in Montgomery form), so this
// IS the reduction of the row sum -- 10 instructions where the Barrett route takes 24.
__device__ __forceinline__ unsigned int mont_reduce96(const Acc96& a, unsigned int p, unsigned int pinv) {
  const unsigned int m1 = (unsigned int)a.lo * pinv;
  const unsigned long long mp1 = (unsigned long long)m1 * p;
  const unsigned long long s1 = a.lo + mp1;                                              // low 32 bits vanish
  const unsigned long long t = (s1 >> 32) + ((unsigned long long)(a.hi + (s1 < mp1 ? 1u : 0u)) << 32);  // (A + m1 p) / 2^32 < 2^49
  const unsigned int m2 = (unsigned int)t * pinv;
  const unsigned long long mp2 = (unsigned long long)m2 * p;
  const unsigned long long s2 = t + mp2;
  const unsigned long long u = (s2 >> 32) + ((unsigned long long)(s2 < mp2 ? 1u : 0u) << 32);            // < p (1 + 2^-17): one subtraction
  return (unsigned int)(u >= p ? u - p : u);
}
// a b 2^-32 mod p for a, b < p (any odd p < 2^32)
__device__ __forceinline__ unsigned int mont_mul32(unsigned int a, unsigned int b, unsigned int p, unsigned int pinv) {
  const unsigned long long x = (unsigned long long)a * b;
  const unsigned int m = (unsigned int)x * pinv;
  const unsigned long long mp = (unsigned long long)m * p;
  const unsigned long long s = x + mp;
  const unsigned long long u = (s >> 32) + ((unsigned long long)(s < mp ? 1u : 0u) << 32);  // < 2p
  return (unsigned int)(u >= p ? u - p : u);
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

// ua[g][j][lane], ub[g][j][lane] = (Philox words & mask) mod p; counter = (sample, j / 4, which), key = seed.
// One thread per (sample, quad of coordinates); lanes = samples, so the stores are coalesced.
__global__ void mm_gen_kernel(unsigned int p, unsigned int mask, unsigned long long seed, unsigned long long first_sample, int batch, int len_a,
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
  for (int t = 0; t < 4 && q * 4 + t < len; ++t) dst[(size_t)t * 32] = (w[t] & mask) % p;  // the integer point is w & mask: the same for every modulus
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
  unsigned long long M;   // floor((2^64-1)/p)
  unsigned int pinv;      // -p^-1 mod 2^32 (odd p: the stored values carry 2^64, the row sum is reduced by mont_reduce96)
  int rows, xlen, groups, nchunks;  // rows: rows of the output per sample group (folded: the stored tasks of all slabs)
  int folded;                       // outputs numbered by task over all slabs (P: mm_verify adds them up through its fold lists), else by (slab, row)
  int cstride;                      // column stride of the block sums, 0: the matrix uses none
  const ChunkDesc* chunk;
  const unsigned char* blob;
  const unsigned int* X;    // [groups][xlen][32]
  unsigned int* out;        // [nslabs][groups][rows][32]; folded: [groups][tasks][32]
  const unsigned int* mul;  // [groups][rows][32] when HAD: out = (A X) o mul  (fused Hadamard step, single-slab matrices; odd p: mul holds
                            // Montgomery forms, i.e. the L product was made with values times 2^96)
};

// 64-bit sum += x: written as a wide multiply-add by an opaque 1; ptxas fuses two of them into one three-input add with two
// carry-outs (IADD3 + IADD3.X on the integer pipe): one instruction per entry, exact for any 32-bit residues.
__device__ __forceinline__ void addwide(unsigned long long& s, unsigned int x, unsigned int one) {
  asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(s) : "r"(x), "r"(one));
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, unsigned int v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory"); }

// the value groups of a row: pa = first group header (the units follow the headers); adds value x (exact 64-bit sum of X words) per group
__device__ __forceinline__ void spmm_groups(Acc96& A, uint32_t pa, unsigned ng, uint32_t xl, unsigned int one) {
  uint32_t un = pa + ((ng + 1) & ~1u) * 8;
#pragma unroll 1
  for (const uint32_t ge = pa + ng * 8; pa != ge; pa += 8) {
    const uint2 h = lds_u64(pa);
    unsigned long long s0 = 0, s1 = 0;
#pragma unroll 1
    for (const uint32_t ue = un + h.y * 16; un != ue; un += 16) {
      const uint4 cw = lds_u128(un);
      const unsigned x0 = lds_u32(xl + cw.x);
      const unsigned x1 = lds_u32(xl + cw.y);
      const unsigned x2 = lds_u32(xl + cw.z);
      const unsigned x3 = lds_u32(xl + cw.w);
      addwide(s0, x0, one);
      addwide(s1, x1, one);
      addwide(s0, x2, one);
      addwide(s1, x3, one);
    }
    const unsigned long long s = s0 + s1;  // < 2^45
    mac96(A, h.x, (unsigned)s);
    // + value * (s >> 32) * 2^32: lands in bits 32..95
    const unsigned long long t = (unsigned long long)h.x * (unsigned)(s >> 32) + (A.lo >> 32) + ((unsigned long long)A.hi << 32);
    A.lo = (A.lo & 0xffffffffull) | (t << 32);
    A.hi = (unsigned)(t >> 32);
  }
}
template <bool MONT>
__device__ __forceinline__ unsigned int spmm_reduce(const Acc96& A, const SpmmArgs& a) {
  return MONT ? mont_reduce96(A, a.p, a.pinv) : reduce96(A, a.p, a.M);
}

// Row descriptor (8 B): {start | np << 16, ng | (row - first row of the blob) << 16}; rows without entries are not stored (their
// outputs stay at the zeros the plan allocated); row stream, offsets in 8-byte words from `start`, every area 16-byte aligned:
//   [np plain pairs (byte offset of the virtual column = 128 * column, value); np even, padding = (0, 0)]
//   [ng group headers (value, nunits), padded to an even count]
//   [units: 4 byte offsets each, all groups in order; a group is padded with the plain pairs' help, never inside a unit]
// Virtual columns of a slab: 0..1023 the columns, 1024 + b the sum over block b of 4 columns, 1280 + b over block b of 16.
template <bool MONT, bool HAD>
__global__ void __launch_bounds__(kSpThreads, 1) mm_slab_spmm_kernel(const SpmmArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kStages], empty[kStages], slabbar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long T = (long long)a.groups * a.nchunks;
  const long long i0 = (long long)blockIdx.x * T / gridDim.x, i1 = (long long)(blockIdx.x + 1) * T / gridDim.x;
  const int nitems = (int)(i1 - i0);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    mbar_init(&slabbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int pg = -1, ps = -1;
  int g = (int)(i0 / a.nchunks), c = (int)(i0 - (long long)g * a.nchunks);
  if (warp == kConsumerWarps) {
    // ---- producer: one thread feeds the slab and the blob ring ----
    if (lane != 0) return;
    for (int k = 0; k < nitems; ++k) {
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
      mbar_expect_tx(&full[stage], (unsigned)ch.bytes);
      bulk_g2s(smem + kRingBase + stage * kStageBytes, a.blob + ch.off, (unsigned)ch.bytes, &full[stage]);
      if (++c == a.nchunks) { c = 0; ++g; }
    }
    return;
  }
  // ---- consumers: lane = sample of the group, warp takes rows of the current blob ----
  // loop invariants, made opaque: under the register cap ptxas would otherwise recompute them (S2R + shifts) inside the row loops
  uint32_t xl, ring;
  asm volatile("mov.u32 %0, %1;" : "=r"(xl) : "r"(smem_u32(smem) + lane * 4));  // this lane's word of slab column 0
  asm volatile("mov.u32 %0, %1;" : "=r"(ring) : "r"(smem_u32(smem) + kRingBase));
  unsigned int one;
  asm volatile("mov.u32 %0, 1;" : "=r"(one));
  unsigned slabphase = 0;
  int4 chn = nitems > 0 ? __ldg(reinterpret_cast<const int4*>(a.chunk + c)) : make_int4(0, 0, 0, 0);  // slab, row0, nrows, bytes
  for (int k = 0; k < nitems; ++k) {
#ifdef PLO_MM_NOPREFETCH
    chn = __ldg(reinterpret_cast<const int4*>(a.chunk + c));
#endif
    const int4 ch = chn;
    {  // descriptor of the next item: loaded a whole blob ahead (an L2 round trip otherwise stalls every warp at every blob)
      const int cn = c + 1 == a.nchunks ? 0 : c + 1;
#ifndef PLO_MM_NOPREFETCH
      if (k + 1 < nitems) chn = __ldg(reinterpret_cast<const int4*>(a.chunk + cn));
#endif
    }
    const int stage = k & (kStages - 1);
    if (g != pg || ch.x != ps) {
      mbar_wait(&slabbar, slabphase);
      slabphase ^= 1; pg = g; ps = ch.x;
      if (a.cstride) {
        // block sums of the new slab (mod p): X4[b] = sum of the 4 columns of block b, X16[b] = sum of the 4 blocks of 4 that make up
        // block b of 16.  The stride is a power of two: block (hi, lo) of level lv has the members ((hi * lv + t) << sh) | lo.
        const int sh = 31 - __clz(a.cstride), lomask = a.cstride - 1;
        for (int b = warp; b < kBlk4; b += kConsumerWarps) {
          const int hi = b >> sh, lo = b & lomask;
          unsigned long long s = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) s += lds_u32(xl + ((((hi * 4 + t) << sh) | lo) << 7));
          sts_u32(xl + (kSlabCols + b) * 128, barrett64(s, a.p, a.M));
        }
        consumer_sync();
        for (int b = warp; b < kBlk16; b += kConsumerWarps) {
          const int hi = b >> sh, lo = b & lomask;
          unsigned long long s = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) s += lds_u32(xl + ((kSlabCols + (((hi * 4 + u) << sh) | lo)) << 7));  // block of 4 number (hi * 4 + u, lo)
          sts_u32(xl + (kSlabCols + kBlk4 + b) * 128, barrett64(s, a.p, a.M));
        }
        consumer_sync();
      }
    }
    mbar_wait(&full[stage], (k / kStages) & 1);
    const uint32_t hdr = ring + stage * kStageBytes;
    const uint32_t st = hdr + ((ch.z * 8 + 15) & ~15);
    unsigned int* outp = a.out + (((size_t)(a.folded ? 0 : ch.x) * a.groups + g) * a.rows + ch.y) * 32 + lane;  // ch.y: first row / first task of the blob
    const unsigned int* mulp = HAD ? a.mul + ((size_t)g * a.rows + ch.y) * 32 + lane : nullptr;
    // rows of the blob are dealt round-robin: after the block sums they are short and of similar cost, and four blobs are in flight
    for (int cur = warp; cur < ch.z; cur += kConsumerWarps) {
      const uint2 rd = lds_u64(hdr + cur * 8);
      const unsigned np = rd.x >> 16, ng = rd.y & 0xffffu, row = rd.y >> 16;  // row: relative to the first row of the blob
      unsigned int mulv = 0;
      if (HAD) mulv = __ldg(mulp + (size_t)row * 32);  // issued now, needed in the epilogue
      uint32_t pa = st + (rd.x & 0xffffu) * 8;
      Acc96 A, B;
      A.lo = 0; A.hi = 0;
      B.lo = 0; B.hi = 0;
#pragma unroll 1
      for (const uint32_t pe = pa + np * 8; pa != pe; pa += 16) {
        const uint4 q = lds_u128(pa);
        const unsigned x0 = lds_u32(xl + q.x);
        const unsigned x1 = lds_u32(xl + q.z);
        mac96(A, q.y, x0);
        mac96(B, q.w, x1);
      }
      add96(A, B);
      if (ng) spmm_groups(A, pa, ng, xl, one);
      unsigned int res = spmm_reduce<MONT>(A, a);
      if (HAD) res = MONT ? mont_mul32(res, mulv, a.p, a.pinv) : barrett64((unsigned long long)res * mulv, a.p, a.M);
      outp[(size_t)(a.folded ? cur : row) * 32] = res;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (++c == a.nchunks) { c = 0; ++g; }
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

// Virtual rows of a matrix with `rows` real rows and row stride rs (0: none): rows + b for the blocks of 4 rows, rows + n4 + b
// for the blocks of 16 (only blocks that lie inside the matrix exist).
struct RowBlocks {
  int rows, rs, n4, n16;
  __host__ __device__ int total() const { return rows + n4 + n16; }
};
__host__ __device__ __forceinline__ RowBlocks row_blocks(int rows, int rs) {
  RowBlocks rb;
  rb.rows = rows; rb.rs = rs;
  rb.n4 = rs ? (rows / (4 * rs)) * rs : 0;
  rb.n16 = rs ? (rows / (16 * rs)) * rs : 0;
  return rb;
}

// bad[b] |= (sum of the partial products of output o != sum_t ua[g][i*k+t][lane] * ub[g][t*n+j][lane])  for o = i*n + j, b = 32 g + lane.
// The partial products of o -- its row in every slab of P plus the virtual rows of the row blocks it belongs to -- are the tasks
// fold_idx[fold_ptr[o] .. fold_ptr[o+1]) of wc[g][task][lane]  (the encoder lists only tasks that exist).
__global__ void mm_verify_kernel(unsigned int p, unsigned long long M, int m, int k, int n, int batch, int groups, int ntasks,
                                 const int* __restrict__ fold_ptr, const int* __restrict__ fold_idx,
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
  const unsigned int* base = wc + (size_t)g * ntasks * 32 + lane;
  for (int q = fold_ptr[o]; q < fold_ptr[o + 1]; ++q) w += base[(size_t)fold_idx[q] * 32];
  Acc96 acc;
  acc.lo = 0; acc.hi = 0;
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
  int cstride;   // column stride of the block sums the rows refer to (0: none)
  RowBlocks rb;  // virtual rows (rb.rs = 0: none)
  long long nnz, loads, blob_bytes;  // entries of the CSR / X loads per sample after encoding / bytes of the encoded matrix
  int folded, ntasks;                // folded: outputs per task, added up through fold_ptr / fold_idx (rows + 1 / entries)
  ChunkDesc* chunk;
  unsigned char* blob;
  int *fold_ptr, *fold_idx;
};

struct plo_mmcheck_plan {
  uint32_t p;
  unsigned long long M;
  int m, k, n, r, batch, groups, grid_cap;
  int mont;             // odd p: the encoded values carry 2^64 (L: 2^96 when the Hadamard step is fused), rows are reduced by mont_reduce96
  uint32_t input_mask;  // the random coordinates are Philox words & input_mask (plo_mmcheck_plan_input_bits)
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

// 2^e mod p, and -p^-1 mod 2^32 for odd p (Newton iteration)
static unsigned pow2_mod(int e, unsigned p) { unsigned long long r = 1 % p; for (int i = 0; i < e; ++i) r = (r * 2) % p; return (unsigned)r; }
static unsigned neg_inv32(unsigned p) { unsigned x = p; for (int i = 0; i < 5; ++i) x *= 2u - p * x; return 0u - x; }

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e && *e ? atoi(e) : dflt; }

namespace {
struct Ent { unsigned col, val; };  // slab-local (virtual) column, residue
typedef std::vector<Ent> Row;
inline bool by_col(const Ent& x, const Ent& y) { return x.col < y.col; }
inline bool by_val_col(const Ent& x, const Ent& y) { return x.val != y.val ? x.val < y.val : x.col < y.col; }
inline unsigned sub_mod(unsigned a, unsigned v, unsigned p) { return a >= v ? a - v : a + (p - v); }

// The value most members of a block carry and what replacing them by one block entry saves:
// gain = (members with that value) - 1 (the block entry) - (members without an entry, which get a correction).
struct Mode { unsigned val; int gain; };
inline Mode block_mode(const unsigned* val, const bool* has, int lv) {
  Mode best{0, -1000};
  int present = 0;
  for (int t = 0; t < lv; ++t) present += has[t];
  for (int t = 0; t < lv; ++t) {
    if (!has[t]) continue;
    int cnt = 0;
    for (int u = 0; u < lv; ++u) cnt += has[u] && val[u] == val[t];
    const int gain = cnt - 1 - (lv - present);
    if (gain > best.gain) { best.gain = gain; best.val = val[t]; }
  }
  return best;
}

// Column block sums of one row inside one slab (levels 16, then 4; stride cs): dense scratch over the 1024 columns.
struct ColScratch {
  std::vector<unsigned> val;
  std::vector<unsigned char> has;
  ColScratch() : val(kSlabCols), has(kSlabCols) {}
};
void col_blocks(Row& row, int ncols, int cs, unsigned p, ColScratch& w, bool apply, long long* gain_out) {
  std::fill(w.has.begin(), w.has.end(), 0);
  Row extra;
  for (const Ent& e : row) {
    if (e.col < (unsigned)kSlabCols) { w.val[e.col] = e.val; w.has[e.col] = 1; }
    else extra.push_back(e);
  }
  long long gain = 0;
  const int levels[2] = {16, 4};
  for (int li = 0; li < 2; ++li) {
    const int lv = levels[li], nb = kSlabCols / lv;
    for (int b = 0; b < nb; ++b) {
      unsigned v[16];
      bool h[16];
      int any = 0, last = 0;
      for (int t = 0; t < lv; ++t) { const int c = blk_member(b, t, lv, cs); v[t] = w.val[c]; h[t] = w.has[c]; any += h[t]; last = c; }
      if (any < 2 || last >= ncols) continue;  // the last member is the largest column: the block must lie inside the slab's columns
      const Mode m = block_mode(v, h, lv);
      if (m.gain < 1) continue;
      gain += m.gain;
      if (!apply) continue;
      extra.push_back(Ent{(unsigned)(kSlabCols + (lv == 4 ? 0 : kBlk4) + b), m.val});
      for (int t = 0; t < lv; ++t) {
        const int c = blk_member(b, t, lv, cs);
        const unsigned nv = sub_mod(h[t] ? v[t] : 0u, m.val, p);
        w.val[c] = nv; w.has[c] = nv != 0;
      }
    }
  }
  if (gain_out) *gain_out += gain;
  if (!apply) return;
  row.clear();
  for (int c = 0; c < kSlabCols; ++c) if (w.has[c]) row.push_back(Ent{(unsigned)c, w.val[c]});
  row.insert(row.end(), extra.begin(), extra.end());
}

// Row block sums inside one slab: the rows {blk_member(b, t, lv, rs)} of block b; entries most of them share move to `vrow`.
void row_block(std::vector<Row>& rows, int b, int lv, int rs, unsigned p, Row* vrow, bool apply, long long* gain_out) {
  struct T { unsigned col, val; int t; };
  std::vector<T> all;
  for (int t = 0; t < lv; ++t) for (const Ent& e : rows[(size_t)blk_member(b, t, lv, rs)]) all.push_back(T{e.col, e.val, t});
  std::sort(all.begin(), all.end(), [](const T& x, const T& y) { return x.col != y.col ? x.col < y.col : x.t < y.t; });
  std::vector<Row> fresh(apply ? (size_t)lv : 0);
  long long gain = 0;
  for (size_t i = 0; i < all.size();) {
    size_t j = i;
    unsigned v[16] = {};
    bool h[16] = {};
    while (j < all.size() && all[j].col == all[i].col) { v[all[j].t] = all[j].val; h[all[j].t] = true; ++j; }
    const Mode m = j - i >= 2 ? block_mode(v, h, lv) : Mode{0, -1};
    if (m.gain >= 1) {
      gain += m.gain;
      if (apply) {
        vrow->push_back(Ent{all[i].col, m.val});
        for (int t = 0; t < lv; ++t) { const unsigned nv = sub_mod(h[t] ? v[t] : 0u, m.val, p); if (nv) fresh[(size_t)t].push_back(Ent{all[i].col, nv}); }
      }
    } else if (apply) {
      for (int t = 0; t < lv; ++t) if (h[t]) fresh[(size_t)t].push_back(Ent{all[i].col, v[t]});
    }
    i = j;
  }
  if (gain_out) *gain_out += gain;
  if (apply) for (int t = 0; t < lv; ++t) rows[(size_t)blk_member(b, t, lv, rs)].swap(fresh[(size_t)t]);
}

// Serialised form of one row (see the kernel).
struct RowStream {
  std::vector<uint2> plain, heads;
  std::vector<uint4> units;
  unsigned words() const { return (unsigned)(((plain.size() + 1) & ~(size_t)1) + ((heads.size() + 1) & ~(size_t)1) + 2 * units.size()); }
  unsigned cost() const { return (unsigned)(24 + 3 * plain.size() + 6 * heads.size() + 5 * units.size()); }
  unsigned loads() const { return (unsigned)(plain.size() + 4 * units.size()); }
};
void encode_row(Row& row, int ming, unsigned long long scale, unsigned p, RowStream* out) {
  auto sc = [&](unsigned v) { return (unsigned)(scale == 1 ? v : (unsigned long long)v * scale % p); };  // values are stored times `scale`
  std::sort(row.begin(), row.end(), by_val_col);
  for (size_t t = 0; t < row.size();) {
    size_t u = t;
    while (u < row.size() && row[u].val == row[t].val) ++u;
    size_t n = u - t;
    if ((long long)n >= ming) {
      const size_t nu = n / 4;  // whole units; the up to 3 entries left over become plain pairs
      out->heads.push_back(make_uint2(sc(row[t].val), (unsigned)nu));
      for (size_t z = 0; z < nu; ++z) out->units.push_back(make_uint4(row[t + 4 * z].col * 128u, row[t + 4 * z + 1].col * 128u, row[t + 4 * z + 2].col * 128u, row[t + 4 * z + 3].col * 128u));
      t += 4 * nu;
    }
    for (; t < u; ++t) out->plain.push_back(make_uint2(row[t].col * 128u, sc(row[t].val)));
  }
}
unsigned write_row(const RowStream& rs, uint2* stream, unsigned o, unsigned rel_row, uint2* desc) {
  const unsigned np = (unsigned)((rs.plain.size() + 1) & ~(size_t)1), ng = (unsigned)rs.heads.size();
  uint2* w = stream + o;
  for (size_t z = 0; z < rs.plain.size(); ++z) w[z] = rs.plain[z];  // padding stays (0, 0): slab column 0 times zero
  w += np;
  for (size_t z = 0; z < rs.heads.size(); ++z) w[z] = rs.heads[z];
  w += (ng + 1) & ~1u;
  for (size_t z = 0; z < rs.units.size(); ++z) reinterpret_cast<uint4*>(w)[z] = rs.units[z];
  *desc = make_uint2(o | (np << 16), ng | (rel_row << 16));
  return rs.words();
}

struct Encoded {
  std::vector<ChunkDesc> table;
  std::vector<unsigned char> blob;
  int cstride = 0;
  RowBlocks rb{};
  long long loads = 0, plain = 0, units = 0, heads = 0, tasks = 0;
  std::vector<int> fold_ptr, fold_idx;  // folded matrices: tasks that add up to every real row
};
}  // namespace

// Cuts the CSR into slabs of kSlabCols columns, rewrites the rows of every slab with row and column block sums where that saves
// loads (see the header comment), cuts every slab into runs of rows of about equal cost and serialises every run as a blob.
static int encode_slab_csr(const plo_csr* h, unsigned p, int groups, int nsm, bool folded, unsigned scale, Encoded* out) {
  const int rows = h->rows, nslabs = (h->cols + kSlabCols - 1) / kSlabCols;
  static const int ming = std::max(2, env_int("PLO_MM_MINGROUP", kMinGroup));
  static const int env_cb = env_int("PLO_MM_COLBLOCKS", -1);  // 0: no column block sums; 2^k: force that stride; default: choose
  static const int env_rb = env_int("PLO_MM_ROWBLOCKS", -1);  // same for the row block sums
  // rows of every slab, sorted by column
  std::vector<std::vector<Row>> slab((size_t)nslabs, std::vector<Row>((size_t)rows));
  for (int i = 0; i < rows; ++i)
    for (long long t = h->ptr[i]; t < h->ptr[i + 1]; ++t) {
      if (h->val[t] == 0) continue;
      slab[(size_t)(h->col[t] / kSlabCols)][(size_t)i].push_back(Ent{(unsigned)(h->col[t] % kSlabCols), h->val[t]});
    }
  long long nnz = 0;
  for (auto& s : slab)
    for (Row& r : s) {
      std::sort(r.begin(), r.end(), by_col);
      // duplicates of a column add up (a CSR may hold them): merge them here
      size_t o = 0;
      for (size_t t = 0; t < r.size(); ++t) {
        if (o && r[o - 1].col == r[t].col) r[o - 1].val = (unsigned)(((unsigned long long)r[o - 1].val + r[t].val) % p);
        else r[o++] = r[t];
      }
      r.resize(o);
      r.erase(std::remove_if(r.begin(), r.end(), [](const Ent& e) { return e.val == 0; }), r.end());
      nnz += (long long)r.size();
    }
  // ---- row block sums: stride chosen on a sample of blocks of 4 ----
  int rs = 0;
  if (folded && env_rb != 0 && rows >= 8) {  // virtual rows need a consumer that folds
    double best = 0;
    for (int cand = 1; 4 * cand <= rows; cand <<= 1) {
      if (env_rb > 0 && cand != env_rb) continue;
      const int nb = (rows / (4 * cand)) * cand, step = std::max(1, nb / 64);
      long long gain = 0, ent = 0;
      for (int s = 0; s < nslabs; s += std::max(1, nslabs / 4))
        for (int b = 0; b < nb; b += step) {
          for (int t = 0; t < 4; ++t) ent += (long long)slab[(size_t)s][(size_t)blk_member(b, t, 4, cand)].size();
          row_block(slab[(size_t)s], b, 4, cand, p, nullptr, false, &gain);
        }
      const double frac = ent ? (double)gain / (double)ent : 0.0;
      if (frac > best) { best = frac; rs = cand; }
    }
    if (env_rb < 0 && best < 0.2) rs = 0;  // worth it from 20 % of the entries on
  }
  const RowBlocks rb = row_blocks(rows, rs);
  out->rb = rb;
  if (rs)
    for (auto& s : slab) {
      s.resize((size_t)rb.total());
      for (int b = 0; b < rb.n16; ++b) row_block(s, b, 16, rs, p, &s[(size_t)(rows + rb.n4 + b)], true, nullptr);
      for (int b = 0; b < rb.n4; ++b) row_block(s, b, 4, rs, p, &s[(size_t)(rows + b)], true, nullptr);
    }
  // ---- column block sums: stride chosen on a sample of rows ----
  ColScratch scratch;
  int cs = 0;
  if (env_cb != 0) {
    long long best = 0, seen = 0;
    const int step = std::max(1, rb.total() / 128);
    for (int s = 0; s < nslabs; s += std::max(1, nslabs / 4)) for (int i = 0; i < rb.total(); i += step) seen += (long long)slab[(size_t)s][(size_t)i].size();
    for (int cand = 1; 16 * cand <= kSlabCols; cand <<= 1) {
      if (env_cb > 0 && cand != env_cb) continue;
      long long gain = 0;
      for (int s = 0; s < nslabs; s += std::max(1, nslabs / 4)) {
        const int ncols = std::min(kSlabCols, h->cols - s * kSlabCols);
        for (int i = 0; i < rb.total(); i += step) col_blocks(slab[(size_t)s][(size_t)i], ncols, cand, p, scratch, false, &gain);
      }
      if (gain > best) { best = gain; cs = cand; }
    }
    if (env_cb < 0 && 10 * best < seen) cs = 0;  // worth the block sums from 10 % of the entries on
  }
  out->cstride = cs;
  // ---- row streams, chunks ----
  std::vector<RowStream> streams((size_t)rb.total());
  std::vector<int> task_of(folded ? (size_t)nslabs * rb.total() : 0, -1);
  int ntask = 0;
  for (int s = 0; s < nslabs; ++s) {
    const int ncols = std::min(kSlabCols, h->cols - s * kSlabCols);
    long long total_cost = 0;
    for (int i = 0; i < rb.total(); ++i) {
      Row& r = slab[(size_t)s][(size_t)i];
      if (cs) col_blocks(r, ncols, cs, p, scratch, true, nullptr);
      RowStream& st = streams[(size_t)i];
      st = RowStream();
      encode_row(r, ming, scale, p, &st);
      if (st.words() > (unsigned)kChunkEnt) { set_error("mmcheck: a row has too many entries inside one %d-column slab", kSlabCols); return PLO_E_ARG; }
      total_cost += st.cost();
      out->tasks += st.words() != 0;
      out->loads += st.loads(); out->plain += (long long)st.plain.size(); out->units += (long long)st.units.size(); out->heads += (long long)st.heads.size();
    }
    long long target = total_cost * nslabs * groups / ((long long)nsm * 16);
    if (target < 512) target = 512;
    int row = 0;
    while (row < rb.total()) {
      while (row < rb.total() && streams[(size_t)row].words() == 0) ++row;  // rows without entries: no task, the output stays zero
      if (row == rb.total()) break;
      int n = 0, span = 0;
      long long w = 0, c = 0;
      while (row + span < rb.total() && n < kChunkRows && span < 65536) {
        const RowStream& st = streams[(size_t)(row + span)];
        if (st.words() == 0) { ++span; continue; }
        if (n > 0 && (c + st.cost() > target || w + st.words() > (unsigned)kChunkEnt)) break;
        w += st.words(); c += st.cost(); ++n; ++span;
      }
      const size_t hdr = ((size_t)n * 8 + 15) & ~(size_t)15;
      ChunkDesc ch;
      ch.slab = s; ch.row0 = folded ? ntask : row; ch.nrows = n; ch.bytes = (int)(hdr + (size_t)w * 8);  // every area is a multiple of 16 B (bulk copies)
      ch.off = out->blob.size(); ch.pad_ = 0;
      out->blob.resize(out->blob.size() + (size_t)ch.bytes, 0);
      uint2* ho = reinterpret_cast<uint2*>(out->blob.data() + ch.off);
      uint2* eo = reinterpret_cast<uint2*>(out->blob.data() + ch.off + hdr);
      unsigned o = 0;
      int z = 0;
      for (int q = 0; q < span; ++q) {
        const RowStream& st = streams[(size_t)(row + q)];
        if (st.words() == 0) continue;
        o += write_row(st, eo, o, (unsigned)q, ho + z);
        if (folded) task_of[(size_t)s * rb.total() + row + q] = ntask++;
        ++z;
      }
      out->table.push_back(ch);
      row += span;
    }
  }
  (void)nnz;
  if (folded) {
    out->fold_ptr.assign(1, 0);
    for (int o = 0; o < rows; ++o) {
      const int v4 = rb.rs && o / (4 * rb.rs) * rb.rs < rb.n4 ? rb.rows + blk_id(o, 4, rb.rs) : -1;
      const int v16 = rb.rs && o / (16 * rb.rs) * rb.rs < rb.n16 ? rb.rows + rb.n4 + blk_id(o, 16, rb.rs) : -1;
      for (int s = 0; s < nslabs; ++s)
        for (int e : {o, v4, v16})
          if (e >= 0 && task_of[(size_t)s * rb.total() + e] >= 0) out->fold_idx.push_back(task_of[(size_t)s * rb.total() + e]);
      out->fold_ptr.push_back((int)out->fold_idx.size());
    }
  }
  return PLO_OK;
}

static int build_slab_csr(const plo_csr* h, unsigned p, int groups, bool folded, unsigned scale, DevSlabCsr* d) {
  d->rows = h->rows; d->cols = h->cols; d->nnz = h->ptr[h->rows];
  d->nslabs = (h->cols + kSlabCols - 1) / kSlabCols;
  d->chunk = nullptr; d->blob = nullptr; d->nchunks = 0; d->fold_ptr = nullptr; d->fold_idx = nullptr;
  Encoded enc;
  const int rc = encode_slab_csr(h, p, groups, sm_count(), folded, scale, &enc);
  if (rc) return rc;
  d->folded = folded ? 1 : 0; d->ntasks = (int)enc.tasks;
  if (folded) {
    PLO_CUDA(pool_alloc(&d->fold_ptr, sizeof(int) * enc.fold_ptr.size()));
    PLO_CUDA(pool_alloc(&d->fold_idx, sizeof(int) * std::max<size_t>(enc.fold_idx.size(), 1)));
    PLO_CUDA(cudaMemcpy(d->fold_ptr, enc.fold_ptr.data(), sizeof(int) * enc.fold_ptr.size(), cudaMemcpyHostToDevice));
    PLO_CUDA(cudaMemcpy(d->fold_idx, enc.fold_idx.data(), sizeof(int) * enc.fold_idx.size(), cudaMemcpyHostToDevice));
  }
  d->cstride = enc.cstride; d->rb = enc.rb; d->loads = enc.loads; d->blob_bytes = (long long)enc.blob.size();
  d->nchunks = (int)enc.table.size();
  PLO_CUDA(pool_alloc(&d->chunk, sizeof(ChunkDesc) * enc.table.size()));
  PLO_CUDA(pool_alloc(&d->blob, enc.blob.size() ? enc.blob.size() : 16));
  PLO_CUDA(cudaMemcpy(d->chunk, enc.table.data(), sizeof(ChunkDesc) * enc.table.size(), cudaMemcpyHostToDevice));
  PLO_CUDA(cudaMemcpy(d->blob, enc.blob.data(), enc.blob.size(), cudaMemcpyHostToDevice));
  return PLO_OK;
}

// Host twin of the consumer side of mm_slab_spmm_kernel + the fold of mm_verify_kernel for ONE sample: y = A x mod p from the
// encoded blobs.  Used by the CPU tests to check the encoder without a device; `stats` = {row stride, column stride, chunks,
// blob bytes, plain entries, units, value groups, X loads per sample, stored (row, slab) tasks}.
static int decode_check(const plo_csr* h, uint32_t p, int groups, int folded, const uint32_t* x, uint32_t* y, long long* stats) {
  Encoded enc;
  const bool mont = (p & 1u) != 0;  // as in the plans: values times 2^64, rows reduced by a Montgomery step that divides by 2^64
  const int rc = encode_slab_csr(h, p, groups, 148, folded != 0, mont ? pow2_mod(64, p) : 1u, &enc);
  if (rc) return rc;
  unsigned long long inv64 = 1;  // 2^-64 mod p = ((p + 1) / 2)^64
  if (mont) for (int i = 0; i < 64; ++i) inv64 = inv64 * ((p + 1ull) / 2) % p;
  const RowBlocks rb = enc.rb;
  const int nslabs = (h->cols + kSlabCols - 1) / kSlabCols;
  std::vector<unsigned long long> ext(folded ? (size_t)enc.tasks : (size_t)nslabs * rb.total(), 0);  // the kernel's output array for one sample
  std::vector<unsigned long long> xs((size_t)kVCols);
  int cur_slab = -1;
  for (const ChunkDesc& ch : enc.table) {
    if ((size_t)ch.nrows * 8 > (size_t)kHdrBytes || ch.bytes > kStageBytes || (ch.bytes & 15)) { set_error("mmcheck encoder: malformed blob"); return PLO_E_ARG; }
    if (ch.slab != cur_slab) {
      cur_slab = ch.slab;
      const int ncols = std::min(kSlabCols, h->cols - ch.slab * kSlabCols);
      for (int c = 0; c < kSlabCols; ++c) xs[(size_t)c] = c < ncols ? x[(size_t)ch.slab * kSlabCols + c] : 0xdeadbeefull;  // never referenced
      if (enc.cstride) {
        for (int b = 0; b < kBlk4; ++b) { unsigned long long s = 0; for (int t = 0; t < 4; ++t) s += xs[(size_t)blk_member(b, t, 4, enc.cstride)]; xs[(size_t)(kSlabCols + b)] = s % p; }
        for (int b = 0; b < kBlk16; ++b) { unsigned long long s = 0; for (int u = 0; u < 4; ++u) s += xs[(size_t)(kSlabCols + blk_id(blk_member(b, 4 * u, 16, enc.cstride), 4, enc.cstride))]; xs[(size_t)(kSlabCols + kBlk4 + b)] = s % p; }
      }
    }
    const uint2* hdr = reinterpret_cast<const uint2*>(enc.blob.data() + ch.off);
    const uint2* st = reinterpret_cast<const uint2*>(enc.blob.data() + ch.off + (((size_t)ch.nrows * 8 + 15) & ~(size_t)15));
    for (int t = 0; t < ch.nrows; ++t) {
      const unsigned start = hdr[t].x & 0xffffu, np = hdr[t].x >> 16, ng = hdr[t].y & 0xffffu, rel = hdr[t].y >> 16;
      unsigned __int128 A = 0;
      auto ld = [&](unsigned off) -> unsigned long long { return (off & 127u) || off / 128u >= (unsigned)kVCols ? ~0ull : xs[off / 128u]; };
      for (unsigned z = 0; z < np; ++z) A += (unsigned __int128)st[start + z].y * ld(st[start + z].x);
      const uint2* gh = st + start + np;
      const uint4* un = reinterpret_cast<const uint4*>(gh + ((ng + 1) & ~1u));
      for (unsigned gi = 0; gi < ng; ++gi) {
        unsigned long long s = 0;
        for (unsigned j = 0; j < gh[gi].y; ++j) s += ld(un[j].x) + ld(un[j].y) + ld(un[j].z) + ld(un[j].w);
        un += gh[gi].y;
        A += (unsigned __int128)gh[gi].x * s;
      }
      const size_t oi = folded ? (size_t)ch.row0 + t : (size_t)ch.slab * rb.total() + ch.row0 + rel;
      if (oi >= ext.size() || (!folded && ch.row0 + (int)rel >= rb.total())) { set_error("mmcheck encoder: output out of range"); return PLO_E_ARG; }
      ext[oi] = (unsigned long long)(A % p) * inv64 % p;
    }
  }
  for (int o = 0; o < h->rows; ++o) {
    unsigned long long w = 0;
    if (folded) {  // what mm_verify_kernel does
      for (int q = enc.fold_ptr[(size_t)o]; q < enc.fold_ptr[(size_t)o + 1]; ++q) w += ext[(size_t)enc.fold_idx[(size_t)q]];
    } else {       // what mm_hadamard_kernel does with the slabs of a wide L or R (no virtual rows without a folding consumer)
      for (int s = 0; s < nslabs; ++s) w += ext[(size_t)s * rb.total() + o];
    }
    y[o] = (uint32_t)(w % p);
  }
  if (stats) {
    const long long st_[9] = {rb.rs, enc.cstride, (long long)enc.table.size(), (long long)enc.blob.size(), enc.plain, enc.units, enc.heads, enc.loads, enc.tasks};
    for (int z = 0; z < 9; ++z) stats[z] = st_[z];
  }
  return PLO_OK;
}

static void free_slab_csr(DevSlabCsr* d) { pool_free(d->chunk); pool_free(d->blob); pool_free(d->fold_ptr); pool_free(d->fold_idx); }

extern "C" {

int plo_mmcheck_encode_check(uint32_t p, const plo_csr* A, int groups, int row_blocks, const uint32_t* x, uint32_t* y, long long* stats) {
  if (p < 2 || !csr_valid(A, p) || groups < 1 || !x || !y) { set_error("plo_mmcheck_encode_check: bad argument"); return PLO_E_ARG; }
  return decode_check(A, p, groups, row_blocks, x, y, stats);
}

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
  PLO_CUDA(cudaFuncSetAttribute(mm_slab_spmm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem));
  PLO_CUDA(cudaFuncSetAttribute(mm_slab_spmm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem));
  PLO_CUDA(cudaFuncSetAttribute(mm_slab_spmm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem));
  PLO_CUDA(cudaFuncSetAttribute(mm_slab_spmm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem));
  plo_mmcheck_plan* pl = new plo_mmcheck_plan();
  memset(pl, 0, sizeof(*pl));
  pl->p = p; pl->M = ~0ull / p; pl->m = m; pl->k = k; pl->n = n; pl->r = r; pl->batch = batch;
  pl->groups = (batch + 31) / 32;
  pl->input_mask = 0xffffffffu;
  pl->grid_cap = sm_count();
  pl->mont = (p & 1u) && !getenv("PLO_MM_NOMONT") ? 1 : 0;  // PLO_MM_NOMONT=1: Barrett reductions as for p = 2 (A/B timing)
  const bool fused_hadamard = L->cols <= kSlabCols && R->cols <= kSlabCols;  // then (L ua) must come out in Montgomery form for the multiplication in R's epilogue
  const unsigned k64 = pl->mont ? pow2_mod(64, p) : 1u, k96 = pl->mont ? pow2_mod(96, p) : 1u;
  rc = build_slab_csr(L, p, pl->groups, false, fused_hadamard ? k96 : k64, &pl->L);
  if (!rc) rc = build_slab_csr(R, p, pl->groups, false, k64, &pl->R);
  if (!rc) rc = build_slab_csr(P, p, pl->groups, true, k64, &pl->P);  // only mm_verify folds virtual rows
  if (rc) { plo_mmcheck_plan_destroy(pl); return rc; }
  const size_t G32 = (size_t)pl->groups * 32;
  const size_t stage = (size_t)batch * (size_t)(m * k > k * n ? m * k : k * n);
  auto zalloc = [](unsigned int** ptr, size_t words) { return pool_alloc(ptr, 4 * words) == cudaSuccess && cudaMemset(*ptr, 0, 4 * words) == cudaSuccess; };
  const bool ok = zalloc(&pl->ua, G32 * m * k) && zalloc(&pl->ub, G32 * k * n) && zalloc(&pl->va, G32 * r * pl->L.nslabs) &&
                  zalloc(&pl->vb, G32 * r * pl->R.nslabs) && zalloc(&pl->vc, G32 * r) && zalloc(&pl->wc, G32 * std::max(pl->P.ntasks, 1)) &&
                  zalloc(&pl->bad, (size_t)batch) && zalloc(&pl->stage, stage);
  if (!ok) { set_error("mmcheck: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())); plo_mmcheck_plan_destroy(pl); return PLO_E_CUDA; }
  *plan = pl;
  return PLO_OK;
}

static void launch_spmm(const plo_mmcheck_plan* pl, const DevSlabCsr& A, const unsigned int* X, unsigned int* out, const unsigned int* mul, cudaStream_t st) {
  SpmmArgs a;
  a.mul = mul;
  a.p = pl->p; a.M = pl->M; a.pinv = pl->mont ? neg_inv32(pl->p) : 0u;
  a.rows = A.folded ? A.ntasks : A.rb.total(); a.folded = A.folded; a.xlen = A.cols; a.groups = pl->groups; a.nchunks = A.nchunks; a.cstride = A.cstride;
  a.chunk = A.chunk; a.blob = A.blob; a.X = X; a.out = out;
  const long long T = (long long)pl->groups * A.nchunks;
  if (T == 0) return;  // a matrix without entries: its product is the zero vector the plan allocated
  const int grid = (int)std::min<long long>(T, pl->grid_cap);
  if (pl->mont && mul) mm_slab_spmm_kernel<true, true><<<grid, kSpThreads, kSpSmem, st>>>(a);
  else if (pl->mont) mm_slab_spmm_kernel<true, false><<<grid, kSpThreads, kSpSmem, st>>>(a);
  else if (mul) mm_slab_spmm_kernel<false, true><<<grid, kSpThreads, kSpSmem, st>>>(a);
  else mm_slab_spmm_kernel<false, false><<<grid, kSpThreads, kSpSmem, st>>>(a);
}

static int mm_pipeline(plo_mmcheck_plan* pl, cudaStream_t st) {
  const int B = pl->batch;
  const size_t G32 = (size_t)pl->groups * 32;
  static const bool timing = getenv("PLO_TIMING") != nullptr;  // per-kernel times on stderr (development aid; serialises the stream)
  cudaEvent_t ev[6] = {};
  int nev = 0;
  auto mark = [&]() { if (timing) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], st); ++nev; } };
  PLO_CUDA(cudaMemsetAsync(pl->bad, 0, 4 * (size_t)B, st));
  mark();
  launch_spmm(pl, pl->L, pl->ua, pl->va, nullptr, st);
  mark();
  if (pl->L.nslabs == 1 && pl->R.nslabs == 1) {
    launch_spmm(pl, pl->R, pl->ub, pl->vc, pl->va, st);  // vc = (R ub) o (L ua) in the epilogue
  } else {
    launch_spmm(pl, pl->R, pl->ub, pl->vb, nullptr, st);
    const size_t cnt = G32 * pl->r;
    mm_hadamard_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(pl->p, pl->M, cnt, pl->L.nslabs, pl->R.nslabs, pl->va, pl->vb, pl->vc);
  }
  mark();
  launch_spmm(pl, pl->P, pl->vc, pl->wc, nullptr, st);
  mark();
  const size_t vcnt = G32 * pl->m * pl->n;
  mm_verify_kernel<<<(unsigned)((vcnt + 255) / 256), 256, 0, st>>>(pl->p, pl->M, pl->m, pl->k, pl->n, B, pl->groups, pl->P.ntasks, pl->P.fold_ptr, pl->P.fold_idx, pl->wc, pl->ua, pl->ub, pl->bad);
  mark();
  PLO_CUDA(cudaGetLastError());
  if (timing) {
    cudaEventSynchronize(ev[nev - 1]);
    float t[4];
    for (int i = 0; i + 1 < nev; ++i) { cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]); }
    fprintf(stderr, "# mmcheck batch %d: L %.1f us, R(+Hadamard) %.1f us, P %.1f us, verify %.1f us\n", B, 1e3 * t[0], 1e3 * t[1], 1e3 * t[2], 1e3 * t[3]);
    for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
  }
  return PLO_OK;
}

int plo_mmcheck_plan_run(plo_mmcheck_plan* pl, uint64_t seed, uint64_t first_sample, void* stream) {
  if (!pl) { set_error("plo_mmcheck_plan_run: null plan"); return PLO_E_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t gthreads = (size_t)pl->batch * ((pl->m * pl->k + 3) / 4 + (pl->k * pl->n + 3) / 4);
  mm_gen_kernel<<<(unsigned)((gthreads + 255) / 256), 256, 0, st>>>(pl->p, pl->input_mask, seed, first_sample, pl->batch, pl->m * pl->k, pl->k * pl->n, pl->ua, pl->ub);
  return mm_pipeline(pl, st);
}

int plo_mmcheck_plan_input_bits(plo_mmcheck_plan* pl, int bits) {
  if (!pl || bits < 1 || bits > 32) { set_error("plo_mmcheck_plan_input_bits: bits must be 1..32"); return PLO_E_ARG; }
  pl->input_mask = bits == 32 ? 0xffffffffu : ((1u << bits) - 1u);
  return PLO_OK;
}

int plo_mmcheck_plan_encoding(const plo_mmcheck_plan* pl, int64_t* loads, int64_t* blob_bytes, int* strides) {
  if (!pl) { set_error("plo_mmcheck_plan_encoding: null plan"); return PLO_E_ARG; }
  const DevSlabCsr* A[3] = {&pl->L, &pl->R, &pl->P};
  for (int z = 0; z < 3; ++z) {
    if (loads) loads[z] = A[z]->loads;
    if (blob_bytes) blob_bytes[z] = A[z]->blob_bytes;
    if (strides) { strides[2 * z] = A[z]->cstride; strides[2 * z + 1] = A[z]->rb.rs; }
  }
  return PLO_OK;
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
