// topk_small: 3 .. 128 queries against the bank in ONE pass at (close to) HBM speed.
//
// Reference: a handful of sequential top_k_cosine_similarity calls (vo:151-188) -- each of them a full pass over the
// bank.  Between the single-query GEMV (topk_single.cu, HBM bound) and the 256-query tiles of sim_tc.cu (tensor
// bound) a small batch is HBM bound as well: 20.5 GB of bank per pass, whatever the number of queries.  Through
// sim_tc.cu such a batch costs 4.3 - 4.8 ms over 10M rows (3.1 ms is the HBM floor): its 256-row query tile is mostly
// padding and is re-fetched for every bank tile, so half of every pipeline stage -- half of the bytes in flight -- is
// not bank at all.  Here the operands swap roles:
//   * M = 128 BANK rows per tile (TMEM lanes), N = 32, 64 or 128 QUERIES (TMEM columns), tcgen05.mma cta_group::1;
//   * a stage holds 16 KB of bank and only 4 / 8 KB of queries, so 11 / 9 stages fit: 176 / 144 KB of bank in
//     flight per SM instead of 96 (128 queries: 16 KB, 6 stages, 96 KB -- as many as through sim_tc.cu, but every
//     tensor-core cycle works on real queries instead of a tile that is half padding);
//   * one CTA per SM, no clusters: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (one TMEM lane quarter
//     each); two accumulators, so the MMAs of tile t + 1 run under the epilogue of tile t.
// Epilogue: a thread owns one bank row and looks at its N dots: dot / |b| against the per-query threshold (k-th best
// score so far, in shared memory); the rare passes are scored exactly (IEEE fp32, the operation order of vo:182) and
// inserted into the CTA's per-query list under a per-query lock.  The CTAs exchange their k-th best scores through a
// global RED.MAX per query every few tiles -- a valid lower bound that only tightens the approximate filter, so the
// results do not depend on timing.
#include "common.cuh"

#include <cuda.h>
#include <cstdlib>

namespace hippo {

constexpr int kSmBM = 128;            // bank rows per tile
constexpr int kSmBK = 64;             // bf16 per 128-byte swizzle row
constexpr int kSmThreads = 192;       // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr uint32_t kSmABytes = kSmBM * kSmBK * 2;       // 16 KiB of bank per stage
constexpr int kSmShare = 8;           // tiles between two exchanges of the k-th best scores

template <int NP> struct SmallCfg {
  static constexpr uint32_t kBBytes = NP * kSmBK * 2;
  static constexpr uint32_t kStage = kSmABytes + kBBytes;
  static constexpr size_t kFixed = 1024 /*align*/ + 256 /*barriers*/ + NP * (HIPPO_TOPK_MAX * 8 + 4 + 4 + 8 + 4) /*lists, thr, an, below, lock*/;
  static constexpr int kStages = (int)((227 * 1024 - kFixed) / kStage);
  static constexpr size_t kSmem = kFixed + (size_t)kStages * kStage;
  static constexpr uint32_t kTmemCols = 2 * NP < 32 ? 32 : 2 * NP;
};

struct SmallParams {
  int64_t n;
  int kblocks;
  const float* bnorm;
  const float* qnorm;
  int nq, k;
  int64_t row_base;
  const uint64_t* after_key;
  uint64_t* part;        // [grid][nq][k]
  uint32_t* gthr;        // [nq] zeroed: max over CTAs of their k-th best score-ord
};

__device__ __forceinline__ float small_threshold(uint64_t kth_key, uint32_t gord, float an) {
  uint32_t ord = (uint32_t)(kth_key >> 32);
  if (gord > ord) ord = gord;
  if (ord == 0) return -INFINITY;                         // list not full, nothing known globally
  if (ord == 0xffffffffu) return INFINITY;                // k NaNs already
  const float lo = ord_to_score(ord) * an;
  if (isinf(lo) || isnan(lo)) return -INFINITY;
  return lo - fabsf(lo) * 9.5367431640625e-07f - 1e-37f;  // 2^-20 relative slack
}

template <int NP>
__global__ void __launch_bounds__(kSmThreads, 1)
topk_small_kernel(const __grid_constant__ CUtensorMap tmBank, const __grid_constant__ CUtensorMap tmQ, const SmallParams p) {
  using C = SmallCfg<NP>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* tiles = smem;
  unsigned char* tail = smem + (size_t)C::kStages * C::kStage;
  uint64_t* lists = reinterpret_cast<uint64_t*>(tail);                          // [NP][k]
  uint64_t* s_below = lists + NP * HIPPO_TOPK_MAX;                               // [NP]
  float* s_thr = reinterpret_cast<float*>(s_below + NP);                         // [NP]
  float* s_an = s_thr + NP;                                                      // [NP]
  uint32_t* s_lock = reinterpret_cast<uint32_t*>(s_an + NP);                     // [NP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lock + NP);
  bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 7) & ~(uintptr_t)7);
  uint64_t* full = bars;                         // [stages]
  uint64_t* empty = bars + C::kStages;           // [stages]
  uint64_t* tfull = bars + 2 * C::kStages;       // [2]
  uint64_t* tempty = tfull + 2;                  // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = p.k;
  const int64_t ntiles = (p.n + kSmBM - 1) / kSmBM;

  for (int i = threadIdx.x; i < NP * HIPPO_TOPK_MAX; i += kSmThreads) lists[i] = 0;
  for (int j = threadIdx.x; j < NP; j += kSmThreads) {
    const float an = j < p.nq ? p.qnorm[j] : 1.f;
    s_an[j] = an;
    s_below[j] = (p.after_key && j < p.nq) ? p.after_key[j] : ~0ull;
    s_thr[j] = j < p.nq ? -INFINITY : INFINITY;          // padding queries never pass
    s_lock[j] = 0;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmBank);
    tma_prefetch_desc(&tmQ);
    for (int i = 0; i < C::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer ----
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = tiles + (size_t)stage * C::kStage;
          mbar_expect_tx(&full[stage], C::kStage);
          tma_load_2d(sa, &tmBank, &full[stage], kb * kSmBK, (int32_t)(t * kSmBM));
          tma_load_2d(sa + kSmABytes, &tmQ, &full[stage], kb * kSmBK, 0);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer ----
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kSmBM, NP);
      int stage = 0;
      uint32_t phase = 0, tile_count = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++tile_count) {
        const uint32_t buf = tile_count & 1, use = tile_count >> 1;
        mbar_wait(&tempty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * NP;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * C::kStage);
          const uint32_t sb = sa + kSmABytes;
#pragma unroll
          for (int k4 = 0; k4 < kSmBK / 16; ++k4)
            umma_bf16(d_tmem, umma_desc_sw128(sa + k4 * 32), umma_desc_sw128(sb + k4 * 32), idesc, (uint32_t)((kb | k4) != 0));
          umma_commit(&empty[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue ----
    const int quarter = warp & 3;                         // TMEM lanes [32 * quarter, +32)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t tile_count = 0;
    float bn_next = 1.f;
    {
      const int64_t r = (int64_t)blockIdx.x * kSmBM + quarter * 32 + lane;
      if (blockIdx.x < ntiles && r < p.n) bn_next = p.bnorm[r];
    }
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++tile_count) {
      const uint32_t buf = tile_count & 1, use = tile_count >> 1;
      const int64_t row = t * kSmBM + quarter * 32 + lane;
      const float bn = bn_next;
      {
        const int64_t rn = row + (int64_t)gridDim.x * kSmBM;
        bn_next = (t + gridDim.x < ntiles && rn < p.n) ? p.bnorm[rn] : 1.f;
      }
      // every kSmShare tiles: publish this CTA's k-th best scores, adopt the best bound anyone has published
      if ((tile_count % kSmShare) == kSmShare - 1 && quarter == 0) {
        for (int j = lane; j < p.nq; j += 32) {
          const uint64_t kth = lists[j * HIPPO_TOPK_MAX + k - 1];
          const uint32_t ord = (uint32_t)(kth >> 32);
          if (kth != 0 && ord != 0xffffffffu) atomicMax(&p.gthr[j], ord);
          const uint32_t g = __ldcg(&p.gthr[j]);
          const float an = s_an[j];
          if (an > 0.f && an < INFINITY) {
            const float cand = small_threshold(kth, g, an);
            if (cand > s_thr[j]) s_thr[j] = cand;          // benign race with the locked updates: both only raise it
          }
        }
      }
      mbar_wait(&tfull[buf], use & 1);
      tc_fence_after();
      const bool in = row < p.n;
      const float inv = __frcp_rn(bn);                     // zero / non-finite norm: NaN or inf below, passes the filter
      // the thread's NP dots in groups of up to 64 columns (registers); the accumulator is handed back to the MMA
      // warp as soon as the last group is in registers
      constexpr int G = NP < 64 ? NP : 64;
#pragma unroll
      for (int g = 0; g < NP / G; ++g) {
        uint32_t v[G];
#pragma unroll
        for (int c = 0; c < G / 32; ++c) {
          uint32_t(&chunk)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32 * c]);
          tmem_ld_32x32(lane_addr + buf * NP + g * G + c * 32, chunk);
        }
        tmem_ld_wait();
        if (g == NP / G - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[buf]);        // the accumulator is in registers: the next MMAs may start
        }
        uint64_t hits = 0;
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const float tv = __uint_as_float(v[j]) * inv;
          hits |= (!(tv < s_thr[g * G + j])) ? (1ull << j) : 0ull;   // NaN passes (vo:185 ranks NaN first)
        }
        if (!in) hits = 0;
        while (hits) {
          const int jl = __ffsll((long long)hits) - 1;
          hits &= hits - 1;
          const int j = g * G + jl;
          if (j >= p.nq) continue;
          // v[jl] for a run-time jl without local memory: select tree over the registers
          uint32_t raw;
          if constexpr (G == 32) {
            raw = select32(*reinterpret_cast<const uint32_t(*)[32]>(&v[0]), jl);
          } else {
            const uint32_t lo = select32(*reinterpret_cast<const uint32_t(*)[32]>(&v[0]), jl & 31);
            const uint32_t hi = select32(*reinterpret_cast<const uint32_t(*)[32]>(&v[32]), jl & 31);
            raw = (jl & 32) ? hi : lo;
          }
          const float dot = __uint_as_float(raw);
          const float an = s_an[j];
          if (dot * inv < s_thr[j]) continue;                // the threshold moved meanwhile
          // the reference's operation order (vo:182): dot / (|b| * |a|), IEEE fp32
          const float sc = __fdiv_rn(dot, __fmul_rn(bn, an));
          const uint64_t key = pack_key(sc, (uint32_t)(p.row_base + row));
          if (!(key < s_below[j])) continue;
          uint64_t* L = lists + j * HIPPO_TOPK_MAX;
          bool done = false;
          while (!done) {
            if (atomicCAS(&s_lock[j], 0u, 1u) == 0u) {
              __threadfence_block();
              if (key > L[k - 1]) {
                int i = k - 1;
                while (i > 0 && L[i - 1] < key) { L[i] = L[i - 1]; --i; }
                L[i] = key;
                if (an > 0.f && an < INFINITY) {
                  const float cand = small_threshold(L[k - 1], 0u, an);
                  if (cand > s_thr[j]) s_thr[j] = cand;
                }
              }
              __threadfence_block();
              atomicExch(&s_lock[j], 0u);
              done = true;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
  // the CTA's lists -> part[blockIdx.x][nq][k]
  for (int i = threadIdx.x; i < p.nq * k; i += kSmThreads) {
    const int j = i / k, r = i - j * k;
    p.part[((size_t)blockIdx.x * p.nq + j) * k + r] = lists[j * HIPPO_TOPK_MAX + r];
  }
}

// ------------------------------------------------------------------ host side ----
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static hippo_status small_tmap(CUtensorMap* tm, const void* base, int64_t rows, int d, int box_rows) {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HIPPO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
      set_error("cuTensorMapEncodeTiled not available from the driver (query result %d)", (int)qres);
      return HIPPO_E_CUDA;
    }
    fn = (EncodeTiledFn2)ptr;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {(cuuint32_t)kSmBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld d=%d box=%d)", (int)r, (long long)rows, d, box_rows);
    return HIPPO_E_CUDA;
  }
  return HIPPO_OK;
}

// 3 .. 128 queries of a dimension that is a multiple of 64.  HIPPO_SMALL_BATCH=0 sends them through sim_tc.cu instead
// (HIPPO_SMALL_BATCH=64: only up to 64 queries, the A/B switch for the 128-query tile).
bool topk_small_supported(int d, int nq) {
  static const int lim = getenv("HIPPO_SMALL_BATCH") ? atoi(getenv("HIPPO_SMALL_BATCH")) : 128;
  return lim > 0 && d % 64 == 0 && nq >= 3 && nq <= (lim < 128 ? lim : 128);
}
int topk_small_npad(int nq) { return nq <= 32 ? 32 : (nq <= 64 ? 64 : 128); }
int topk_small_grid(int64_t n) {
  const int64_t tiles = (n + kSmBM - 1) / kSmBM;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  return (int)(tiles < sms ? (tiles < 1 ? 1 : tiles) : sms);
}

// qbf: [npad, d] bf16 (rows >= nq zero), qnorm [nq], part [grid][nq][k], gthr [nq] zeroed.  *nparts_out = grid.
hippo_status topk_small_launch(const void* bank, const float* norm, int64_t n, int d, const void* qbf, const float* qnorm,
                               int nq, int k, int64_t row_base, const uint64_t* after_key, uint64_t* part, uint32_t* gthr,
                               int* nparts_out, cudaStream_t s) {
  const int npad = topk_small_npad(nq);
  CUtensorMap tmBank, tmQ;
  hippo_status st = small_tmap(&tmBank, bank, n, d, kSmBM);
  if (st != HIPPO_OK) return st;
  st = small_tmap(&tmQ, qbf, npad, d, npad);
  if (st != HIPPO_OK) return st;
  SmallParams p{};
  p.n = n;
  p.kblocks = d / kSmBK;
  p.bnorm = norm;
  p.qnorm = qnorm;
  p.nq = nq;
  p.k = k;
  p.row_base = row_base;
  p.after_key = after_key;
  p.part = part;
  p.gthr = gthr;
  const int grid = topk_small_grid(n);
  if (npad == 32) {
    HIPPO_CUDA(cudaFuncSetAttribute(topk_small_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmallCfg<32>::kSmem));
    topk_small_kernel<32><<<grid, kSmThreads, SmallCfg<32>::kSmem, s>>>(tmBank, tmQ, p);
  } else if (npad == 64) {
    HIPPO_CUDA(cudaFuncSetAttribute(topk_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmallCfg<64>::kSmem));
    topk_small_kernel<64><<<grid, kSmThreads, SmallCfg<64>::kSmem, s>>>(tmBank, tmQ, p);
  } else {
    HIPPO_CUDA(cudaFuncSetAttribute(topk_small_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmallCfg<128>::kSmem));
    topk_small_kernel<128><<<grid, kSmThreads, SmallCfg<128>::kSmem, s>>>(tmBank, tmQ, p);
  }
  HIPPO_CUDA(cudaGetLastError());
  *nparts_out = grid;
  return HIPPO_OK;
}

}  // namespace hippo
