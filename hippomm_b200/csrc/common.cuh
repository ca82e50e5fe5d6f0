// Shared host/device helpers for libhippo_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/hippo_b200.h"

namespace hippo {

// ---------------------------------------------------------------- errors ----
void set_error(const char* fmt, ...);
hippo_status cuda_fail(cudaError_t e, const char* what);
hippo_status check_arch();          // HIPPO_OK iff current device is sm_100
int sm_count();

#define HIPPO_CUDA(expr)                                         \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return ::hippo::cuda_fail(_e, #expr); \
  } while (0)

#define HIPPO_REQUIRE(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::hippo::set_error(__VA_ARGS__);    \
      return HIPPO_E_BADARG;              \
    }                                     \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Carver {
  char* base; size_t size; size_t off;
  Carver(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T> T* take(size_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base ? base + off : nullptr);
    off += count * sizeof(T);
    return r;
  }
  size_t used() const { return align_up(off, 256); }
  bool ok() const { return base != nullptr && used() <= size && ((uintptr_t)base & 255) == 0; }
};

// ------------------------------------------------------- order-key packing ----
// A search hit is ordered by (score descending, row ascending); NaN above all
// numbers (np.argsort puts NaN last, vo:185 reverses).  Both are folded into
// one uint64 so that "better" is a plain integer '>':
//   hi32 = monotone map of the fp32 score (NaN -> 0xffffffff), lo32 = ~row.
// key 0 never occurs for a real hit and marks an empty slot.
__host__ __device__ __forceinline__ uint32_t score_to_ord(float s) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(s);
#else
  uint32_t b; memcpy(&b, &s, 4);
#endif
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;           // NaN
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_score(uint32_t o) {
  uint32_t b;
  if (o == 0xffffffffu) b = 0x7fc00000u;
  else b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f; memcpy(&f, &b, 4); return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t pack_key(float s, uint32_t row) {
  return ((uint64_t)score_to_ord(s) << 32) | (uint64_t)(~row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~(uint32_t)k; }
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return ord_to_score((uint32_t)(k >> 32)); }

#ifdef __CUDACC__
// Sorted (descending) insert of `key` into list[0..k); list[k-1] is the worst.
// Caller guarantees key > list[k-1].
__device__ __forceinline__ void topk_insert(uint64_t* list, int k, uint64_t key) {
  int j = k - 1;
  while (j > 0 && list[j - 1] < key) { list[j] = list[j - 1]; --j; }
  list[j] = key;
}

// ------------------------------------------------- warp-level k-way selection ----
// Every lane keeps the best k (<= 16) of the keys it has seen in a sorted register list (compare-swap chain with
// static indices: no local memory); warp_select_best then pops the global maximum k times.  Keys are unique per
// row, 0 = empty.
constexpr int kLaneList = HIPPO_TOPK_MAX;
__device__ __forceinline__ void lane_list_clear(uint64_t (&L)[kLaneList]) {
#pragma unroll
  for (int i = 0; i < kLaneList; ++i) L[i] = 0;
}
__device__ __forceinline__ void lane_list_insert(uint64_t (&L)[kLaneList], int k, uint64_t key) {
  uint64_t v = key;
#pragma unroll
  for (int i = 0; i < kLaneList; ++i) {
    if (i < k) {
      const bool gt = v > L[i];
      const uint64_t lo = gt ? L[i] : v;
      L[i] = gt ? v : L[i];
      v = lo;
    }
  }
}
// Returns, in lane r < k, the r-th best key over all lanes' lists (0 when fewer than r + 1 exist).
__device__ __forceinline__ uint64_t warp_select_best(uint64_t (&L)[kLaneList], int k, int lane) {
  uint64_t mine = 0;
  for (int r = 0; r < k; ++r) {
    uint64_t best = L[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == r) mine = best;
    if (best == 0) break;                       // uniform: nothing left anywhere
    if (L[0] == best) {                         // unique keys: exactly one lane pops
#pragma unroll
      for (int i = 0; i + 1 < kLaneList; ++i) L[i] = L[i + 1];
      L[kLaneList - 1] = 0;
    }
  }
  return mine;
}

// ------------------------------------------------------------ warp helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// "SSIM of this pair not delivered yet" (pattern.cu fills the array with it, frames.cu overwrites it with one 8-byte
// store per pair, segment.cu polls it): a NaN payload no arithmetic produces (0 / 0 gives the canonical 0x7ff8.. / 0xfff8..)
constexpr unsigned long long kSsimPending = 0xfff8b200c0de0001ull;
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// r[j] for a run-time j without local memory: a 5-level select tree over the 32 registers.
__device__ __forceinline__ uint32_t select32(const uint32_t (&r)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? r[2 * i + 1] : r[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return (j & 16) ? d[1] : d[0];
}

// ------------------------------------------------- mbarrier / TMA / tcgen05 ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 2-D tiled TMA load global -> shared, completion on an mbarrier (tx bytes).
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every tcgen05.mma issued so far has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster drive one 256-row MMA ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
// Same without the release fence: for arrivals that publish no ordinary memory writes (the data they
// stand for is tracked by complete_tx bytes or ordered by tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const void* tmap, uint32_t bar_cluster_addr,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// Same with an L2 cache-policy operand (createpolicy encodings as CUTLASS's TMA::CacheHintSm90).
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_pair_hint(void* dst, const void* tmap, uint32_t bar_cluster_addr,
                                                      int32_t c0, int32_t c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
// D[TMEM of both CTAs] (+)= A[256 rows, 128 per CTA] * B[N rows, N/2 per CTA]; issued by the leader CTA
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Once every MMA issued so far has retired, arrive on the mbarrier at this offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (sm_100 "version 1"):
// rows of 64 bf16 (128 B), 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                   // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace hippo
