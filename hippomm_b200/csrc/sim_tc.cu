// tcgen05 similarity contraction  S = A . B^T  (bf16 x bf16 -> fp32 in TMEM) with the
// consumer fused into the TMEM epilogue, so the fp32 similarity matrix never exists:
//
//   EPI_TOPK : batched detailed-recall search.  A = queries, B = bank rows.  Reference:
//              nq sequential top_k_cosine_similarity calls (vo:151-188, callers hm:3153,3304).
//              An epilogue thread owns one query row x 128 bank columns of a tile, scales the dots
//              by 1/|b| and only values passing a running k-th-best threshold reach the exact
//              IEEE evaluation dot/(|b||a|) and the thread's sorted top-k list.
//   EPI_MASK : memory consolidation.  A = B = feature rows.  Reference: the N x N
//              `np.dot(Fn, Fn.T)` of hm:952 and the `< threshold` test of hm:960.  Only the
//              lower triangle is contracted; the epilogue emits 1 bit per pair plus the
//              list of pairs too close to gamma to trust bf16 inputs.
//
// Structure (persistent, one CTA per SM, CTAs paired into 2-CTA clusters): the pair computes a
// 256 x 256 tile with tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A and HALF of
// the B tile (128 rows), the tensor core of each SM reads the other half from the peer's shared
// memory, and each CTA's TMEM receives its 128 rows of the result.
//   warp 0     TMA producer (128B-swizzled K-major tiles, 6-stage mbarrier ring).  Only the pair
//              LEADER arrives on a stage's full barrier (arrive.expect_tx for the bytes of both CTAs);
//              the peer's TMA credits its bytes to the leader's barrier directly, so no cross-CTA
//              arrive sits on the per-stage critical path (that alone was worth 2x, profiles/).
//   warp 1     MMA issuer (leader CTA only, one thread; M256 N256 K16; fp32 accumulators double-
//              buffered across the 512 TMEM columns; tcgen05.commit multicast to both CTAs).
//   warps 2-9  epilogue: two warps per TMEM lane quarter, each taking 128 of the tile's 256 columns
//              (tcgen05.ld 32x32b), next tile's column norms / thresholds prefetched under the math.
#include "sim_tc.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace hippo {

constexpr int EPI_TOPK = 0;
constexpr int EPI_MASK = 1;

constexpr uint32_t kABytes = kTcBM * kTcBK * 2;           // 16 KiB: this CTA's 128 rows of A
constexpr uint32_t kBBytes = (kTcBN / 2) * kTcBK * 2;     // 16 KiB: this CTA's half of the B tile
constexpr uint32_t kStageBytes = kABytes + kBBytes;       // 32 KiB per CTA and stage
constexpr uint32_t kTmemCols = 512;                       // 2 accumulators x 256 columns
constexpr int kEpiThreads = kTcThreads - 64;              // 256
constexpr int kEpiWarps = kEpiThreads / 32;               // 8
constexpr size_t kSmemTiles = (size_t)kTcStages * kStageBytes;
constexpr size_t kSmemColF = 2 * 2 * kTcBN * sizeof(float);           // [buf][norm|inv][256]
constexpr size_t kSmemSpec = 2 * 8 * sizeof(uint32_t);                // [buf][8] special-column masks
constexpr size_t kSmemCols = kSmemColF + kSmemSpec + 2 * 8 * 2 * sizeof(float);   // + [buf][8][min|max] chunk norm range
constexpr size_t kSmemBytes = 1024 /*align slack*/ + kSmemTiles + kSmemCols + 256 /*barriers*/;

struct TcParams {
  // common
  int64_t n;          // rows of B
  int kblocks;        // d / 64
  const float* bnorm; // norms of B rows
  int units;
  int debug;          // profiling knobs (HIPPO_TC_DEBUG): 1 = no epilogue at all, 2 = TMEM loads but no math,
                      // 4 = no norm staging / barrier, 8 = never take the slow path, 16 = no TMA loads, 32 = no MMAs
  // top-k
  const float* qnorm;
  int nq, k, m_pairs /* 256-query blocks */, n_tiles, splits, tiles_per_split;
  int64_t row_base;
  const uint64_t* after_key;
  uint64_t* part;     // [2 * splits, nq, k]
  uint32_t* thr_ord;  // [nq]      global lower bound on the k-th best score (monotone, atomicMax)
  uint32_t* pool;     // [nq, k]   global pool of the best scores seen by anyone (see pool_update)
  uint32_t* prog;     // [units]   tiles whose loads each unit's leader has started (flow control, see the producer)
  int fc_window;      // a pair may run at most this many tiles ahead of the slowest pair of its cohort (0 = off)
  unsigned long long* counters;  // profiling (debug & 64): [0] slow-chunk calls, [1] insertions, [2] cycles in the
                                 // slow path (per warp), [3] cycles epilogue warps wait for accumulators, [4] epilogue tiles x warps,
                                 // [5] cycles the MMA thread waits for TMEM, [6] cycles it waits for operands
  // mask (consolidation).  Two shapes, both with A = rows [a_row0, a_row0 + na) of the bf16 image X of the caller's
  // rows (one band) and bits written band-locally (row = A row - a_row0):
  //   mask_rect == 0  TRIANGLE: B = the same rows; column tiles t <= row block; bit (i, j) kept for j < i
  //   mask_rect == 1  RECTANGLE: B = the first *dyn_k rows of Y (the rows kept so far, compacted); every column tile,
  //                   bits of columns >= *dyn_k cleared (those rows of Y are stale)
  int mask_rect;
  const int32_t* dyn_k;   // rectangle: device count of kept rows = valid rows of B
  int na;                 // rows of A (the band)
  int64_t a_row0, b_row0; // first row of A / B inside their tensor maps
  const float* anorm;     // norms of the A rows (already offset: anorm[0] belongs to row a_row0)
  float gamma, band_exact, band_inexact;
  const int32_t* inexact;
  uint32_t* mask;         // triangle: bit matrix
  int64_t words_per_row;
  int32_t* rowhit;        // rectangle: [na] set to 1 for a band row that conflicts with a kept row for certain
  uint2* uncertain;
  int32_t* uncertain_count;
  int32_t uncertain_cap;
};

// One tile of a pair's static schedule.  A unit is (256-row block, range of bank tiles); the row block
// varies fastest across pairs so the pairs running concurrently stream the same bank tiles through L2.
struct Tile {
  int unit, m_base, t, t1, split;
  bool first;   // first tile of its unit
};

template <int EPI>
__device__ __forceinline__ void decode_unit(const TcParams& p, int unit, int& m_base, int& t0, int& t1,
                                            int& split) {
  if constexpr (EPI == EPI_TOPK) {
    split = unit / p.m_pairs;
    m_base = 2 * (unit - split * p.m_pairs);
    t0 = split * p.tiles_per_split;
    t1 = min(t0 + p.tiles_per_split, p.n_tiles);
  } else if (p.mask_rect) {
    // rectangle: h row blocks of the band against every column tile of the kept rows, column tile slowest so that
    // the pairs running concurrently share it
    const int h = (p.na + kTcBN - 1) / kTcBN;
    t0 = unit / h;
    m_base = 2 * (unit - t0 * h);
    t1 = t0 + 1;
    split = 0;
  } else {
    // triangle of the band's h row blocks (t <= I), column tile slowest
    const int h = (p.na + kTcBN - 1) / kTcBN;
    int r = unit, j = 0;
    while (r >= h - j) { r -= h - j; ++j; }
    t0 = j;
    m_base = 2 * (j + r);
    t1 = t0 + 1;
    split = 0;
  }
}
template <int EPI>
__device__ __forceinline__ bool first_tile(const TcParams& p, int pair, Tile& x) {
  if (pair >= p.units) return false;
  int t0;
  x.unit = pair;
  decode_unit<EPI>(p, x.unit, x.m_base, t0, x.t1, x.split);
  x.t = t0;
  x.first = true;
  return true;
}
template <int EPI>
__device__ __forceinline__ bool next_tile(const TcParams& p, int npairs, Tile& x) {
  if (x.t + 1 < x.t1) { ++x.t; x.first = false; return true; }
  if (x.unit + npairs >= p.units) return false;
  int t0;
  x.unit += npairs;
  decode_unit<EPI>(p, x.unit, x.m_base, t0, x.t1, x.split);
  x.t = t0;
  x.first = true;
  return true;
}

__device__ __forceinline__ float topk_filter_threshold(uint64_t kth_key, float an) {
  if (kth_key == 0) return -INFINITY;
  uint32_t ord = (uint32_t)(kth_key >> 32);
  if (ord == 0xffffffffu) return INFINITY;
  float lo = ord_to_score(ord) * an;
  if (isinf(lo)) return lo;
  return lo - fabsf(lo) * 9.5367431640625e-07f - 1e-37f;   // 2^-20 relative slack
}

// Global per-query score pool: k slots, a row goes to slot hash(row) % k, a slot keeps the best score-ord
// hashed to it (fire-and-forget RED.MAX, no dependent round trips).  The k slots hold the scores of k
// DISTINCT rows (different hash classes), so min(pool) is a valid lower bound on the final k-th best score,
// under any interleaving.  It is published through thr_ord (RED.MAX, monotone), which only gates the
// approximate filter (with slack): results do not depend on timing.
__device__ __forceinline__ void pool_push(uint32_t* pool_q, int k, uint32_t ord, uint32_t grow) {
  const int slot = (int)(((uint64_t)(grow * 2654435761u) * (uint32_t)k) >> 32);
  atomicMax(&pool_q[slot], ord);                       // result unused: a fire-and-forget RED.MAX
}
// min over the k slots = a valid lower bound on the final k-th best score; published through thr_ord.  One L2 round
// trip (the k loads are independent): done ONCE per chunk after all of the chunk's hits have been pushed -- a scene of
// near-duplicate rows puts dozens of hits into one chunk, and a round trip per hit held the warp (and with it the
// TMEM accumulator, and with that the MMA issuer) for ~700 cycles each.
__device__ __forceinline__ uint32_t pool_min(const uint32_t* pool_q, int k, uint32_t* thr_ord_q) {
  uint32_t m = 0xffffffffu;
  for (int j = 0; j < k; ++j) {
    const uint32_t g = __ldcg(&pool_q[j]);
    m = g < m ? g : m;
  }
  if (m != 0) atomicMax(thr_ord_q, m);
  return m;
}

// The per-thread top-k list lives in registers (kListRegs 64-bit entries, sorted descending; entries at
// and beyond k stay 0): with ~210 KB of shared memory carved out the L1 is tiny, and a list in local memory
// turned every insertion into a chain of L2 round trips (32k cycles per slow chunk, profiles/).
constexpr int kListRegs = HIPPO_TOPK_MAX;
static_assert(kListRegs == 16, "the register-resident top-k list is sized for HIPPO_TOPK_MAX == 16");

// Sorted insert by a compare-swap chain with static indices; the displaced tail value is dropped.
__device__ __forceinline__ void list_insert(uint64_t (&L)[kListRegs], int k, uint64_t key) {
  uint64_t v = key;
#pragma unroll
  for (int i = 0; i < kListRegs; ++i) {
    if (i < k) {
      const bool gt = v > L[i];
      const uint64_t lo = gt ? L[i] : v;
      L[i] = gt ? v : L[i];
      v = lo;
    }
  }
}
// list[k-1]: the list is sorted descending with zeros in unused slots, so that is the minimum of the first
// k entries (written as a reduction so the compiler cannot turn it into an indexed local-memory load).
__device__ __forceinline__ uint64_t list_kth(const uint64_t (&L)[kListRegs], int k) {
  uint64_t r = ~0ull;
#pragma unroll
  for (int i = 0; i < kListRegs; ++i) {
    const uint64_t v = (i < k) ? L[i] : ~0ull;
    r = v < r ? v : r;
  }
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3 on sm_100
  return r;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
sim_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const TcParams p_in) {
  extern __shared__ unsigned char smem_raw[];
  // the banded consolidation reads its extent from device memory (no host sync between bands); the search
  // kernel keeps reading its parameters straight from the constant bank
  TcParams pm;
  if constexpr (EPI == EPI_MASK) {
    pm = p_in;
    const int h = (p_in.na + kTcBN - 1) / kTcBN;
    if (p_in.mask_rect) {
      pm.n = *p_in.dyn_k;                                  // valid rows of B = rows kept so far
      pm.n_tiles = (int)((pm.n + kTcBN - 1) / kTcBN);
      pm.units = pm.n_tiles * h;
    } else {
      pm.n = p_in.na;
      pm.n_tiles = h;
      pm.units = h * (h + 1) / 2;
    }
  }
  const TcParams& p = [&]() -> const TcParams& {
    if constexpr (EPI == EPI_MASK) return pm; else return p_in;
  }();
  // identical carve-up in both CTAs of the pair: the MMA and the multicast commits address the peer's
  // shared memory by the same offsets (pointer arithmetic on the __shared__ array keeps LDS/STS)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* s_cols = reinterpret_cast<float*>(smem + kSmemTiles);                     // [buf][2][256]
  uint32_t* s_spec = reinterpret_cast<uint32_t*>(smem + kSmemTiles + kSmemColF);   // [buf][8]
  float* s_bmm = reinterpret_cast<float*>(smem + kSmemTiles + kSmemColF + kSmemSpec);   // [buf][8][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemTiles + kSmemCols);
  uint64_t* full = bars;                       // [stages] TMA (both CTAs) -> MMA; the leader's copy is used
  uint64_t* empty = bars + kTcStages;          // [stages] MMA -> TMA, one per CTA (multicast commit)
  uint64_t* tfull = bars + 2 * kTcStages;      // [2] MMA -> epilogue, one per CTA (multicast commit)
  uint64_t* tempty = bars + 2 * kTcStages + 2; // [2] epilogue warps of both CTAs -> MMA; the leader's copy
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();     // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kTcStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<kTmemCols>(s_tmem);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer ----
    // Lane 0 issues the loads.  Flow control (top-k only): the pairs whose units cover the same bank split in
    // the same round (a "cohort", up to m_pairs pairs) stream the same bank tiles; left alone they drift apart
    // (SMs far from an L2 slice run ~10% slower) until a tile is evicted before the last pair asks for it, and
    // the bank was read from HBM ~8x per step (profiles/).  Every leader therefore publishes how many tiles it
    // has started and does not run more than fc_window tiles ahead of the slowest member of its cohort.  The
    // counters only delay loads: results do not depend on them.
    int stage = 0;
    uint32_t phase = 0;
    Tile x;
    int tiu = 0, u_lo = 0, u_hi = 0;
    uint32_t known_min = 0;
    const bool fc = EPI == EPI_TOPK && p.fc_window > 0 && rank == 0 && p.prog != nullptr;
    for (bool have = first_tile<EPI>(p, pair, x); have; have = next_tile<EPI>(p, npairs, x)) {
      if (fc) {
        if (x.first) {
          tiu = 0;
          known_min = 0;
          const int round = x.unit / npairs;
          u_lo = max(x.split * p.m_pairs, round * npairs);
          u_hi = min(min((x.split + 1) * p.m_pairs, (round + 1) * npairs), p.units);
        }
        ++tiu;
        if (lane == 0) *(volatile uint32_t*)&p.prog[x.unit] = (uint32_t)tiu;
        const int need = tiu - p.fc_window;
        if (need > 0 && known_min < (uint32_t)need) {
          uint32_t m;
          for (;;) {
            m = 0xffffffffu;
            for (int u = u_lo + lane; u < u_hi; u += 32) {
              const uint32_t v = *(volatile const uint32_t*)&p.prog[u];
              m = v < m ? v : m;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const uint32_t other = __shfl_xor_sync(0xffffffffu, m, o);
              m = other < m ? other : m;
            }
            if (m >= (uint32_t)need) break;
            __nanosleep(256);
          }
          known_min = m;
        }
      }
      if (lane == 0) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = smem + (size_t)stage * kStageBytes;
          const uint32_t leader_full = map_to_cta(smem_u32(&full[stage]), 0);
          if (p.debug & 16) {
            if (rank == 0) mbar_arrive(&full[stage]);
          } else {
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * kStageBytes);   // bytes landing in both CTAs
            int32_t ra = (x.m_base + (int)rank) * kTcBM, rb = x.t * kTcBN + (int)rank * (kTcBN / 2);
            if constexpr (EPI == EPI_MASK) { ra += (int32_t)p.a_row0; rb += (int32_t)p.b_row0; }
            tma_load_2d_pair(sa, &tmA, leader_full, kb * kTcBK, ra);
            tma_load_2d_pair(sa + kABytes, &tmB, leader_full, kb * kTcBK, rb);
          }
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // -------------------------------------------------- MMA issuer (leader CTA) ----
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kTcBM, kTcBN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_count = 0;
      Tile x;
      for (bool have = first_tile<EPI>(p, pair, x); have; have = next_tile<EPI>(p, npairs, x), ++tile_count) {
        const uint32_t buf = tile_count & 1, use = tile_count >> 1;
        long long tw = (p.debug & 64) ? clock64() : 0;
        mbar_wait(&tempty[buf], (use & 1) ^ 1);
        if (p.debug & 64) atomicAdd(&p.counters[5], (unsigned long long)(clock64() - tw));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kTcBN;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tw = (p.debug & 64) ? clock64() : 0;
          mbar_wait(&full[stage], phase);
          if (p.debug & 64) atomicAdd(&p.counters[6], (unsigned long long)(clock64() - tw));
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          if (!(p.debug & 32)) {
#pragma unroll
            for (int k4 = 0; k4 < kTcBK / 16; ++k4) {
              umma_bf16_pair(d_tmem, umma_desc_sw128(sa + k4 * 32), umma_desc_sw128(sb + k4 * 32), idesc,
                             (uint32_t)((kb | k4) != 0));
            }
          }
          umma_commit_pair(&empty[stage], 3);  // frees the slot in both CTAs once these MMAs retire
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull[buf], 3);      // accumulators (both CTAs' TMEM) ready for the epilogues
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue ----
    const int ew = warp - 2;                      // 0..7
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) belong to this warp
    const int chalf = ew >> 2;                    // which 128 columns of the tile this warp takes
    const int row_in_tile = quarter * 32 + lane;
    const int et = threadIdx.x - 64;              // 0..255: the tile column whose norm this thread stages
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t tempty_addr[2];
    tempty_addr[0] = map_to_cta(smem_u32(&tempty[0]), 0);
    tempty_addr[1] = map_to_cta(smem_u32(&tempty[1]), 0);
    uint32_t tile_count = 0;
    uint64_t list[kListRegs];
    uint64_t kth_key = 0;                     // list[k-1], kept alongside so the hot path never indexes the list

    // per-(unit, row) state
    bool valid = false, force = false;
    float an = 1.f, g_i = 0.f, b_i = 0.f;
    uint64_t below = ~0ull;
    int64_t arow = 0;
    const float band = (EPI == EPI_MASK) ? ((p.inexact && *p.inexact) ? p.band_inexact : p.band_exact) : 0.f;

    Tile cur;
    bool have = first_tile<EPI>(p, pair, cur);
    // prefetched for the tile about to be processed: its column norm (this thread's column), the row
    // norm of the unit's row, and the global threshold
    float pf_bn = 1.f, pf_rown = 0.f;
    uint32_t pf_thr = 0;
    auto prefetch = [&](const Tile& x) {
      const int64_t col = (int64_t)x.t * kTcBN + et;
      pf_bn = col < p.n ? p.bnorm[col] : 1.f;
      const int64_t r = (int64_t)(x.m_base + (int)rank) * kTcBM + row_in_tile;
      if constexpr (EPI == EPI_TOPK) {
        if (x.first) pf_rown = r < p.nq ? p.qnorm[r] : 1.f;
        pf_thr = r < p.nq ? __ldcg(&p.thr_ord[r]) : 0u;
      } else {
        if (x.first) pf_rown = r < p.na ? p.anorm[r] : 0.f;
      }
    };
    if (have) prefetch(cur);

    while (have) {
      const uint32_t buf = tile_count & 1, use = tile_count >> 1;
      float* s_bn = s_cols + buf * 2 * kTcBN;
      float* s_inv = s_bn + kTcBN;
      const int t = cur.t;
      if (cur.first) {
        arow = (int64_t)(cur.m_base + (int)rank) * kTcBM + row_in_tile;
        if constexpr (EPI == EPI_TOPK) {
          valid = arow < p.nq;
          an = pf_rown;
          below = (valid && p.after_key) ? p.after_key[arow] : ~0ull;
          force = valid && !(an > 0.f && an < INFINITY);   // zero / non-finite query: every score is NaN
#pragma unroll
          for (int i = 0; i < kListRegs; ++i) list[i] = 0;
          kth_key = 0;
        } else {
          valid = arow < p.na;
          const float ni = valid ? pf_rown : 0.f;
          const bool ok = ni > 0.f && ni < INFINITY;
          g_i = ok ? p.gamma * ni : -INFINITY;   // zero / non-finite row: every sim is NaN -> bit set
          b_i = ok ? band * ni : -1.f;
        }
      }
      // stage this tile's column norms (buffer `buf`; the barrier below also separates the writes from
      // the reads of the tile two steps back).  Thread et covers column et, so a warp covers one
      // 32-column chunk and a ballot yields the chunk's mask of "special" columns: zero / non-finite
      // norms, whose similarity is NaN and must reach the exact path.
      if (!(p.debug & 4)) {
        const int64_t col = (int64_t)t * kTcBN + et;
        const bool in = col < p.n;
        const float bn = pf_bn;
        const bool special = in && !(bn > 0.f && bn < INFINITY);
        s_bn[et] = bn;
        s_inv[et] = in ? __frcp_rn(bn) : 0.f;
        const uint32_t m = __ballot_sync(0xffffffffu, special);
        if constexpr (EPI == EPI_TOPK) {
          // smallest / largest ordinary norm of the chunk: dot/|b| >= thr needs max(dot) >= thr * |b|_min
          // (thr > 0) or thr * |b|_max (thr <= 0), so the quick test below needs no per-element multiply
          float lo = (in && !special) ? bn : INFINITY, hi = (in && !special) ? bn : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
          }
          if (lane == 0) { s_bmm[(buf * 8 + ew) * 2] = lo; s_bmm[(buf * 8 + ew) * 2 + 1] = hi; }
        }
        if (lane == 0) s_spec[buf * 8 + ew] = m;
      }
      float thr = INFINITY;
      const uint32_t gthr = pf_thr;          // global bound on the k-th best score as of the last tile
      if constexpr (EPI == EPI_TOPK) {
        if (valid) {
          const uint64_t g = (uint64_t)pf_thr << 32;
          thr = topk_filter_threshold(g > kth_key ? g : kth_key, an);
        }
      }
      // next tile: issue its global loads now, they complete under this tile's math
      Tile nxt = cur;
      const bool have_next = next_tile<EPI>(p, npairs, nxt);
      if (have_next) prefetch(nxt);

      if (!(p.debug & 4)) epi_bar_sync();
      long long t_wait0 = 0, t_chunk = 0;
      if (p.debug & 64) t_wait0 = clock64();
      mbar_wait(&tfull[buf], use & 1);
      tc_fence_after();
      if (p.debug & 64) {
        t_chunk = clock64();
        if (lane == 0) { atomicAdd(&p.counters[3], (unsigned long long)(t_chunk - t_wait0)); atomicAdd(&p.counters[4], 1ull); }
      }

      const uint32_t acc_addr = lane_addr + buf * kTcBN + chalf * (kTcBN / 2);
      const int64_t col0 = (int64_t)t * kTcBN + chalf * (kTcBN / 2);
      uint32_t ra[32], rb[32];
      uint32_t words[4] = {0u, 0u, 0u, 0u};
      if (!(p.debug & 1)) {
        tmem_ld_32x32(acc_addr, ra);
        // two chunks per iteration: the TMEM load of one register set overlaps the arithmetic on the other
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int c = c2 * 2 + hh;            // chunk within this warp's 128 columns
            uint32_t(&cur_r)[32] = hh ? rb : ra;
            uint32_t(&nxt_r)[32] = hh ? ra : rb;
            tmem_ld_wait();
            if (c + 1 < 4) tmem_ld_32x32(acc_addr + (c + 1) * 32, nxt_r);
            const int cc = chalf * 4 + c;         // chunk within the tile
            if (p.debug & 2) continue;
            const float4* inv4 = reinterpret_cast<const float4*>(s_inv + cc * 32);
            const uint32_t spec = s_spec[buf * 8 + cc];
            if constexpr (EPI == EPI_TOPK) {
              // branch-free quick test: max over the chunk of the raw dots against thr * (extreme norm of the chunk)
              float m0 = __uint_as_float(cur_r[0]), m1 = __uint_as_float(cur_r[1]);
#pragma unroll
              for (int j = 2; j < 32; j += 4) {
                m0 = fmax3(m0, __uint_as_float(cur_r[j]), __uint_as_float(cur_r[j + 1]));
                if (j + 2 < 32) m1 = fmax3(m1, __uint_as_float(cur_r[j + 2]), __uint_as_float(cur_r[j + 3]));
              }
              const float2 bmm = *reinterpret_cast<const float2*>(&s_bmm[(buf * 8 + cc) * 2]);
              const float mx = fmaxf(m0, m1);
              const float bound = thr * (thr > 0.f ? bmm.x : bmm.y);
              if (valid && (!(mx < bound) || spec != 0u || force) && !(p.debug & 8)) {
                // rare path, all in registers: mask of the columns that pass, then one candidate at a time
                uint32_t hits = 0;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 iv = inv4[j4];
                  const float ivs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj) {
                    const float tv = __uint_as_float(cur_r[4 * j4 + jj]) * ivs[jj];
                    hits |= (!(tv < thr)) ? (1u << (4 * j4 + jj)) : 0u;   // NaN passes (vo:185 ranks NaN first)
                  }
                }
                if (force) hits = 0xffffffffu;
                uint32_t gord = gthr;
                uint32_t n_calls = 0;
                bool pushed = false;
                while (hits) {
                  const int j = __ffs(hits) - 1;
                  hits &= hits - 1;
                  const int64_t col = col0 + c * 32 + j;
                  if (col >= p.n) continue;
                  const float dot = __uint_as_float(select32(cur_r, j));
                  if (!force && (dot * s_inv[cc * 32 + j] < thr)) continue;     // threshold moved meanwhile
                  // the reference's operation order (vo:182): dot / (|b| * |a|), IEEE fp32
                  const float sc = __fdiv_rn(dot, __fmul_rn(s_bn[cc * 32 + j], an));
                  const uint32_t grow = (uint32_t)(p.row_base + col);
                  const uint64_t key = pack_key(sc, grow);
                  if (key < below && key > kth_key) {
                    list_insert(list, p.k, key);
                    kth_key = list_kth(list, p.k);
                    ++n_calls;
                    const uint32_t ord = (uint32_t)(key >> 32);
                    if (ord > gord && p.pool != nullptr) {
                      pool_push(p.pool + (size_t)arow * p.k, p.k, ord, grow);
                      pushed = true;
                    }
                    const uint64_t g = (uint64_t)gord << 32;
                    thr = topk_filter_threshold(g > kth_key ? g : kth_key, an);
                  }
                }
                if (pushed) {
                  const uint32_t m = pool_min(p.pool + (size_t)arow * p.k, p.k, &p.thr_ord[arow]);
                  gord = m > gord ? m : gord;
                  const uint64_t g = (uint64_t)gord << 32;
                  thr = topk_filter_threshold(g > kth_key ? g : kth_key, an);
                }
                if (p.debug & 64) { atomicAdd(&p.counters[0], 1ull); atomicAdd(&p.counters[1], (unsigned long long)n_calls); }
              }
              if (p.debug & 64) {   // cycles of chunks that took the slow path (a fast chunk is < 400 clk)
                __syncwarp();
                const long long t_now = clock64();
                if (lane == 0 && t_now - t_chunk > 400) atomicAdd(&p.counters[2], (unsigned long long)(t_now - t_chunk));
                t_chunk = t_now;
              }
            } else {
              // bit j = !(sim < gamma): sign bit of (dot/|b_j| - gamma |b_i|), shifted in from the top
              uint32_t w = 0;
              bool unc = false;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 iv = inv4[j4];
                const float ivs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  const float d = fmaf(__uint_as_float(cur_r[4 * j4 + jj]), ivs[jj], -g_i);
                  w = (w >> 1) | (~__float_as_uint(d) & 0x80000000u);
                  unc |= fabsf(d) <= b_i;
                }
              }
              w |= spec;                                  // NaN similarity: `NaN < gamma` is False (hm:960)
              // triangle: pairs j < i only; rectangle: columns below the number of kept rows only
              const int64_t jb = col0 + c * 32;
              const int64_t jend = p.mask_rect ? p.n : arow;
              uint32_t keep;
              if (!valid || jb >= jend) keep = 0u;
              else if (jb + 32 <= jend) keep = 0xffffffffu;
              else keep = (1u << (uint32_t)(jend - jb)) - 1u;
              uint32_t wk = w & keep;
              if (unc && keep) {
                // rare path: pairs inside the band go to the uncertain list (re-evaluated from fp32 rows, which then
                // sets their bit); a pair that does not fit the list keeps its tensor-core bit (and is reported)
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 iv = inv4[j4];
                  const float ivs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * j4 + jj;
                    const float d = fmaf(__uint_as_float(cur_r[j]), ivs[jj], -g_i);
                    if ((fabsf(d) <= b_i) && ((keep >> j) & 1u)) {
                      const int pos = atomicAdd(p.uncertain_count, 1);
                      if (pos < p.uncertain_cap) {
                        p.uncertain[pos] = make_uint2((uint32_t)arow, (uint32_t)(jb + j));
                        wk &= ~(1u << j);
                      }
                    }
                  }
                }
              }
              if (c2 == 0) { if (hh == 0) words[0] = wk; else words[1] = wk; }
              else { if (hh == 0) words[2] = wk; else words[3] = wk; }
            }
          }
        }
      }
      // TMEM buffer fully read by this warp: one arrival per warp on the LEADER's barrier
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(tempty_addr[buf]);
      if constexpr (EPI == EPI_MASK) {
        if (valid) {
          if (p.mask_rect) {
            // rectangle: every column is a row that was KEPT, so one conflict settles the band row; nobody needs the
            // individual bits (every writer stores the same 1)
            if ((words[0] | words[1] | words[2] | words[3]) != 0u) p.rowhit[arow] = 1;
          } else {
            // band-local bit matrix: row = row of the band, 8 words per column tile
            uint4* dst = reinterpret_cast<uint4*>(p.mask + arow * p.words_per_row + (int64_t)t * 8 + chalf * 4);
            *dst = make_uint4(words[0], words[1], words[2], words[3]);
          }
        }
      } else {
        const bool last = !have_next || nxt.first;
        if (last && valid) {
          uint64_t* dst = p.part + ((size_t)(cur.split * 2 + chalf) * p.nq + arow) * p.k;
#pragma unroll
          for (int i = 0; i < kListRegs; ++i)
            if (i < p.k) dst[i] = list[i];
        }
      }
      cur = nxt;
      have = have_next;
      ++tile_count;
    }
  }

  // neither CTA may exit (or free TMEM) while its peer can still touch its shared memory / barriers
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static hippo_status get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HIPPO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
      set_error("cuTensorMapEncodeTiled not available from the driver (query result %d)", (int)qres);
      return HIPPO_E_CUDA;
    }
    fn = (EncodeTiledFn)ptr;
  }
  *out = fn;
  return HIPPO_OK;
}

// 2-D bf16 row-major [rows, d] tensor, boxes of box_rows x 64 elements, 128-byte swizzle.
static hippo_status make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int d, int box_rows) {
  EncodeTiledFn enc;
  hippo_status st = get_encode_fn(&enc);
  if (st != HIPPO_OK) return st;
  if (((uintptr_t)base & 15) != 0) { set_error("TMA base pointer must be 16-byte aligned"); return HIPPO_E_BADARG; }
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld d=%d box=%d)", (int)r,
              (long long)rows, d, box_rows);
    return HIPPO_E_CUDA;
  }
  return HIPPO_OK;
}

// Flow-control window in tiles (HIPPO_TC_WINDOW overrides; 0 disables).
static int flow_window() {
  const char* e = getenv("HIPPO_TC_WINDOW");
  return e ? atoi(e) : 8;
}

static int debug_flags() {
  const char* e = getenv("HIPPO_TC_DEBUG");
  return e ? atoi(e) : 0;
}

template <int EPI>
static hippo_status launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t s,
                           int max_pairs = 0) {
  HIPPO_CUDA(cudaFuncSetAttribute(sim_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSmemBytes));
  int pairs = sm_count() / 2;                 // one CTA per SM, two SMs per 256-row tile
  if (max_pairs > 0 && pairs > max_pairs) pairs = max_pairs;
  if (pairs > p.units) pairs = p.units;
  if (pairs < 1) return HIPPO_OK;
  sim_tc_kernel<EPI><<<2 * pairs, kTcThreads, kSmemBytes, s>>>(tmA, tmB, p);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

int tc_topk_splits(int64_t n, int nq) {
  const int pairs = (sm_count() > 0 ? sm_count() : 148) / 2;
  const int m_blocks = (nq + 2 * kTcBM - 1) / (2 * kTcBM);    // 256-query blocks
  const int64_t n_tiles = (n + kTcBN - 1) / kTcBN;
  if (n_tiles <= 0 || m_blocks <= 0) return 1;
  // aim for units = m_blocks * splits ~ a multiple of the pair count, up to 8 waves
  int64_t best = 1;
  double best_cost = 1e300;
  for (int waves = 1; waves <= 8; ++waves) {
    int64_t splits = ((int64_t)pairs * waves + m_blocks - 1) / m_blocks;
    if (splits > n_tiles) splits = n_tiles;
    if (splits < 1) splits = 1;
    const int64_t tps = (n_tiles + splits - 1) / splits;
    const int64_t real_splits = (n_tiles + tps - 1) / tps;
    const int64_t units = real_splits * m_blocks;
    const int64_t rounds = (units + pairs - 1) / pairs;
    // cost = critical-path tiles + a small per-unit overhead
    const double cost = (double)rounds * (double)(tps + 2);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = real_splits; }
  }
  return (int)best;
}

hippo_status tc_topk_launch(const TcTopkArgs& a, cudaStream_t s) {
  CUtensorMap tmA, tmB;
  hippo_status st = make_tmap(&tmA, a.qbf16, a.nq_rows > a.nq ? a.nq_rows : a.nq, a.d, kTcBM);
  if (st != HIPPO_OK) return st;
  st = make_tmap(&tmB, a.bank, a.n, a.d, kTcBN / 2);
  if (st != HIPPO_OK) return st;
  TcParams p{};
  p.n = a.n;
  p.kblocks = a.d / kTcBK;
  p.bnorm = a.bnorm;
  p.qnorm = a.qnorm;
  p.nq = a.nq;
  p.k = a.k;
  p.m_pairs = (a.nq + 2 * kTcBM - 1) / (2 * kTcBM);
  p.n_tiles = (int)((a.n + kTcBN - 1) / kTcBN);
  p.tiles_per_split = (p.n_tiles + a.splits - 1) / a.splits;
  p.splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  if (p.splits != a.splits) { set_error("tc_topk_launch: inconsistent split count"); return HIPPO_E_BADARG; }
  p.units = p.m_pairs * p.splits;
  p.row_base = a.row_base;
  p.after_key = a.after_key;
  p.part = a.part;
  p.thr_ord = a.thr_ord;
  // a handful of queries: every pair scans its own part of the bank for the SAME few queries and the shared pool is a
  // handful of cache lines hammered by every SM (7.0 ms for 8 queries against 3.9 ms for 64) -- those batches take
  // topk_small.cu now; from 129 queries on the pool pays (same box: 129 queries 4.31 -> 3.87 ms, 192: 5.18 -> 4.84,
  // 256: 4.80 -> 4.63; HIPPO_TC_POOL_BLOCKS=2 restores the old rule)
  static const int pool_blocks = getenv("HIPPO_TC_POOL_BLOCKS") ? atoi(getenv("HIPPO_TC_POOL_BLOCKS")) : 1;
  p.pool = (p.m_pairs >= pool_blocks && (p.m_pairs >= 2 || a.nq > 128)) ? a.pool : nullptr;
  p.counters = a.counters;
  p.prog = a.progress;
  p.fc_window = flow_window();
  p.debug = debug_flags();
  return launch<EPI_TOPK>(tmA, tmB, p, s);
}

hippo_status tc_mask_launch(const TcMaskArgs& a, cudaStream_t s) {
  CUtensorMap tmA, tmB;
  hippo_status st = make_tmap(&tmA, a.a_rows, a.a_total, a.d, kTcBM);
  if (st != HIPPO_OK) return st;
  st = make_tmap(&tmB, a.b_rows, a.b_total, a.d, kTcBN / 2);
  if (st != HIPPO_OK) return st;
  TcParams p{};
  p.kblocks = a.d / kTcBK;
  p.mask_rect = a.rect ? 1 : 0;
  p.dyn_k = a.dyn_k;
  p.na = a.na;
  p.a_row0 = a.a_row0;
  p.b_row0 = a.b_row0;
  p.anorm = a.anorm;
  p.bnorm = a.bnorm;
  const int h = (a.na + kTcBN - 1) / kTcBN;
  // the device-side prologue recomputes n / n_tiles / units (the rectangle's extent lives in device memory); the
  // host values only bound the grid: the triangle's unit count, or the most the rectangle can have
  p.n = a.na;
  p.n_tiles = h;
  const int64_t bt = (a.b_total + kTcBN - 1) / kTcBN;
  const int64_t units = a.rect ? bt * h : (int64_t)h * (h + 1) / 2;
  if (units > 0x7fffffffll) { set_error("tc_mask_launch: too many tiles"); return HIPPO_E_BADARG; }
  p.units = (int)units;
  p.gamma = a.gamma;
  p.band_exact = a.band_exact;
  p.band_inexact = a.band_inexact;
  p.inexact = a.inexact;
  p.mask = a.mask;
  p.words_per_row = a.words_per_row;
  p.rowhit = a.rowhit;
  p.uncertain = a.uncertain;
  p.uncertain_count = a.uncertain_count;
  p.uncertain_cap = a.uncertain_cap;
  p.debug = debug_flags();
  return launch<EPI_MASK>(tmA, tmB, p, s, a.max_pairs);
}

}  // namespace hippo
