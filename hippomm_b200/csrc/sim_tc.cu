// tcgen05 similarity contraction  S = A . B^T  (bf16 x bf16 -> fp32 in TMEM) with the
// consumer fused into the TMEM epilogue, so the fp32 similarity matrix never exists:
//
//   EPI_TOPK : batched detailed-recall search.  A = queries, B = bank rows.  Reference:
//              nq sequential top_k_cosine_similarity calls (vo:151-188, callers hm:3153,3304).
//              Each epilogue thread owns one query (one TMEM lane), scales the 256 dots of a
//              tile by 1/|b|, and only values passing a running k-th-best threshold reach
//              the exact IEEE evaluation dot/(|b||a|) and the thread's sorted top-k list.
//   EPI_MASK : memory consolidation.  A = B = feature rows.  Reference: the N x N
//              `np.dot(Fn, Fn.T)` of hm:952 and the `< threshold` test of hm:960.  Only the
//              lower triangle is contracted; the epilogue emits 1 bit per pair plus the
//              list of pairs too close to gamma to trust bf16 inputs.
//
// Structure (one CTA per SM, persistent): warp 0 = TMA producer (128B-swizzled K-major
// tiles, 4-stage mbarrier ring), warp 1 = single-thread tcgen05.mma issuer (M128 N256 K16,
// fp32 accumulators double-buffered across the 512 TMEM columns), warps 2..5 = epilogue
// (tcgen05.ld 32x32b, one TMEM lane quarter per warp).
#include "sim_tc.cuh"

#include <cuda.h>

namespace hippo {

constexpr int EPI_TOPK = 0;
constexpr int EPI_MASK = 1;

constexpr uint32_t kABytes = kTcBM * kTcBK * 2;           // 16 KiB
constexpr uint32_t kBBytes = kTcBN * kTcBK * 2;           // 32 KiB
constexpr uint32_t kStageBytes = kABytes + kBBytes;       // 48 KiB
constexpr uint32_t kTmemCols = 512;                       // 2 accumulators x 256 columns
constexpr size_t kSmemTiles = (size_t)kTcStages * kStageBytes;
constexpr size_t kSmemCols = 2 * 2 * kTcBN * sizeof(float);  // [buf][norm|inv][256]
constexpr size_t kSmemBytes = 1024 /*align slack*/ + kSmemTiles + kSmemCols + 256 /*barriers*/;

struct TcParams {
  // common
  int64_t n;          // rows of B
  int kblocks;        // d / 64
  const float* bnorm; // norms of B rows
  int units;
  // top-k
  const float* qnorm;
  int nq, k, m_blocks, n_tiles, splits, tiles_per_split;
  int64_t row_base;
  const uint64_t* after_key;
  uint64_t* part;
  uint32_t* thr_ord;
  // mask
  float gamma, band_exact, band_inexact;
  const int32_t* inexact;
  uint32_t* mask;
  int64_t words_per_row;
  uint2* uncertain;
  int32_t* uncertain_count;
  int32_t uncertain_cap;
};

// A unit is what one CTA takes from the static schedule: `m_count` row blocks starting at
// `m_first`, each against bank tiles [t0, t1).
template <int EPI>
__device__ __forceinline__ void decode_unit(const TcParams& p, int unit, int& m_first, int& m_count,
                                            int& t0, int& t1, int& split) {
  if constexpr (EPI == EPI_TOPK) {
    // query block varies fastest so the CTAs running concurrently share bank tiles in L2
    split = unit / p.m_blocks;
    m_first = unit - split * p.m_blocks;
    m_count = 1;
    t0 = split * p.tiles_per_split;
    t1 = min(t0 + p.tiles_per_split, p.n_tiles);
  } else {
    // lower triangle of 256 x 256 super-blocks: unit = I(I+1)/2 + t, t <= I
    int I = (int)((sqrtf(8.f * (float)unit + 1.f) - 1.f) * 0.5f);
    while ((int64_t)(I + 1) * (I + 2) / 2 <= unit) ++I;
    while ((int64_t)I * (I + 1) / 2 > unit) --I;
    t0 = unit - (int)((int64_t)I * (I + 1) / 2);
    t1 = t0 + 1;
    m_first = 2 * I;
    m_count = ((int64_t)(2 * I + 1) * kTcBM < p.n) ? 2 : 1;
    split = 0;
  }
}

__device__ __forceinline__ float topk_filter_threshold(uint64_t kth_key, float an) {
  if (kth_key == 0) return -INFINITY;
  uint32_t ord = (uint32_t)(kth_key >> 32);
  if (ord == 0xffffffffu) return INFINITY;
  float lo = ord_to_score(ord) * an;
  if (isinf(lo)) return lo;
  return lo - fabsf(lo) * 9.5367431640625e-07f - 1e-37f;
}

// Exact evaluation + list insertion of one candidate (rare path, kept out of line).
__device__ __noinline__ void topk_consider(uint64_t* list, int k, float dot, float bn, float an,
                                           uint32_t grow, uint64_t below, uint32_t* thr_ord_q) {
  // the reference's operation order (vo:182): dot / (|b| * |a|), IEEE fp32
  float s = __fdiv_rn(dot, __fmul_rn(bn, an));
  uint64_t key = pack_key(s, grow);
  if (key < below && key > list[k - 1]) {
    topk_insert(list, k, key);
    if (list[k - 1] != 0) atomicMax(thr_ord_q, (uint32_t)(list[k - 1] >> 32));
  }
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int EPI>
__global__ void __launch_bounds__(kTcThreads, 1)
sim_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* s_cols = reinterpret_cast<float*>(smem + kSmemTiles);      // [buf][2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemTiles + kSmemCols);
  uint64_t* full = bars;                 // [stages]  TMA -> MMA
  uint64_t* empty = bars + kTcStages;    // [stages]  MMA -> TMA
  uint64_t* tfull = bars + 2 * kTcStages;      // [2] MMA -> epilogue
  uint64_t* tempty = bars + 2 * kTcStages + 2; // [2] epilogue -> MMA
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kTcStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer ----
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        int m_first, m_count, t0, t1, split;
        decode_unit<EPI>(p, unit, m_first, m_count, t0, t1, split);
        for (int mi = 0; mi < m_count; ++mi) {
          for (int t = t0; t < t1; ++t) {
            for (int kb = 0; kb < p.kblocks; ++kb) {
              mbar_wait(&empty[stage], phase ^ 1);
              unsigned char* sa = smem + (size_t)stage * kStageBytes;
              mbar_expect_tx(&full[stage], kStageBytes);
              tma_load_2d(sa, &tmA, &full[stage], kb * kTcBK, (m_first + mi) * kTcBM);
              tma_load_2d(sa + kABytes, &tmB, &full[stage], kb * kTcBK, t * kTcBN);
              if (++stage == kTcStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer ----
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTcBM, kTcBN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_count = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        int m_first, m_count, t0, t1, split;
        decode_unit<EPI>(p, unit, m_first, m_count, t0, t1, split);
        for (int mi = 0; mi < m_count; ++mi) {
          for (int t = t0; t < t1; ++t, ++tile_count) {
            const uint32_t buf = tile_count & 1, use = tile_count >> 1;
            mbar_wait(&tempty[buf], (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * kTcBN;
            for (int kb = 0; kb < p.kblocks; ++kb) {
              mbar_wait(&full[stage], phase);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + (size_t)stage * kStageBytes);
              const uint32_t sb = sa + kABytes;
#pragma unroll
              for (int k4 = 0; k4 < kTcBK / 16; ++k4) {
                umma_bf16(d_tmem, umma_desc_sw128(sa + k4 * 32), umma_desc_sw128(sb + k4 * 32), idesc,
                          (uint32_t)((kb | k4) != 0));
              }
              umma_commit(&empty[stage]);  // frees the smem slot once these MMAs retire
              if (++stage == kTcStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull[buf]);      // accumulator ready for the epilogue
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue ----
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) belong to this warp
    const int row_in_tile = quarter * 32 + lane;
    const int et = threadIdx.x - 64;              // 0..127
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t tile_count = 0;
    uint64_t list[HIPPO_TOPK_MAX];

    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      int m_first, m_count, t0, t1, split;
      decode_unit<EPI>(p, unit, m_first, m_count, t0, t1, split);
      for (int mi = 0; mi < m_count; ++mi) {
        const int64_t arow = (int64_t)(m_first + mi) * kTcBM + row_in_tile;  // query / row i
        // ---- per-(unit, row) state
        bool valid;
        float an = 1.f;
        uint64_t below = ~0ull;
        float g_i = 0.f, b_i = 0.f;
        if constexpr (EPI == EPI_TOPK) {
          valid = arow < p.nq;
          if (valid) {
            an = p.qnorm[arow];
            if (p.after_key) below = p.after_key[arow];
          }
#pragma unroll
          for (int i = 0; i < HIPPO_TOPK_MAX; ++i) list[i] = 0;
        } else {
          valid = arow < p.n;
          const float band = (p.inexact && *p.inexact) ? p.band_inexact : p.band_exact;
          float ni = valid ? p.bnorm[arow] : 0.f;
          const bool ok = ni > 0.f && !isinf(ni);
          g_i = ok ? p.gamma * ni : -INFINITY;   // zero / non-finite row: every sim is NaN -> bit set
          b_i = ok ? band * ni : -1.f;
        }

        for (int t = t0; t < t1; ++t, ++tile_count) {
          const uint32_t buf = tile_count & 1, use = tile_count >> 1;
          float* s_bn = s_cols + buf * 2 * kTcBN;
          float* s_inv = s_bn + kTcBN;
          // stage this tile's column norms (writes buffer `buf`; the barrier below also
          // separates them from the reads of the tile two steps back)
          for (int c = et; c < kTcBN; c += 128) {
            const int64_t col = (int64_t)t * kTcBN + c;
            const float bn = col < p.n ? p.bnorm[col] : 0.f;
            s_bn[c] = bn;
            s_inv[c] = col < p.n ? __frcp_rn(bn) : 0.f;
          }
          float thr = INFINITY;
          if constexpr (EPI == EPI_TOPK) {
            if (valid) {
              uint64_t kth = list[p.k - 1];
              const uint64_t g = (uint64_t)__ldcg(&p.thr_ord[arow]) << 32;
              kth = g > kth ? g : kth;
              thr = topk_filter_threshold(kth, an);
            }
          }
          epi_bar_sync();
          mbar_wait(&tfull[buf], use & 1);
          tc_fence_after();

          const uint32_t acc_addr = lane_addr + buf * kTcBN;
          const int64_t col0 = (int64_t)t * kTcBN;
          uint32_t ra[32], rb[32];
          uint32_t words[8];
          tmem_ld_32x32(acc_addr, ra);
#pragma unroll
          for (int c = 0; c < kTcBN / 32; ++c) {
            uint32_t(&cur)[32] = (c & 1) ? rb : ra;
            uint32_t(&nxt)[32] = (c & 1) ? ra : rb;
            tmem_ld_wait();
            if (c + 1 < kTcBN / 32) tmem_ld_32x32(acc_addr + (c + 1) * 32, nxt);
            const float* inv = s_inv + c * 32;
            if constexpr (EPI == EPI_TOPK) {
              bool hit = false;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float tv = __uint_as_float(cur[j]) * inv[j];
                hit |= !(tv < thr);
              }
              if (hit && valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float dot = __uint_as_float(cur[j]);
                  const float tv = dot * inv[j];
                  if (!(tv < thr)) {
                    const int64_t col = col0 + c * 32 + j;
                    if (col < p.n) {
                      topk_consider(list, p.k, dot, s_bn[c * 32 + j], an,
                                    (uint32_t)(p.row_base + col), below, &p.thr_ord[arow]);
                      uint64_t kth = list[p.k - 1];
                      const uint64_t g = (uint64_t)__ldcg(&p.thr_ord[arow]) << 32;
                      kth = g > kth ? g : kth;
                      thr = topk_filter_threshold(kth, an);
                    }
                  }
                }
              }
            } else {
              uint32_t w = 0;
              bool unc = false;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float tv = __uint_as_float(cur[j]) * inv[j];
                w |= (!(tv < g_i)) ? (1u << j) : 0u;
                unc |= fabsf(tv - g_i) <= b_i;
              }
              // keep pairs j < i only
              const int64_t jb = col0 + c * 32;
              uint32_t keep;
              if (!valid || jb >= arow) keep = 0u;
              else if (jb + 32 <= arow) keep = 0xffffffffu;
              else keep = (1u << (uint32_t)(arow - jb)) - 1u;
              words[c] = w & keep;
              if (unc && keep) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float tv = __uint_as_float(cur[j]) * inv[j];
                  if ((fabsf(tv - g_i) <= b_i) && ((keep >> j) & 1u)) {
                    const int pos = atomicAdd(p.uncertain_count, 1);
                    if (pos < p.uncertain_cap)
                      p.uncertain[pos] = make_uint2((uint32_t)arow, (uint32_t)(jb + j));
                  }
                }
              }
            }
          }
          // TMEM buffer fully read: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tempty[buf]);
          if constexpr (EPI == EPI_MASK) {
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(p.mask + arow * p.words_per_row + (int64_t)t * 8);
              dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
              dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
            }
          }
        }
        if constexpr (EPI == EPI_TOPK) {
          if (valid) {
            uint64_t* dst = p.part + ((size_t)split * p.nq + arow) * p.k;
            for (int i = 0; i < p.k; ++i) dst[i] = list[i];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
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

template <int EPI>
static hippo_status launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t s) {
  HIPPO_CUDA(cudaFuncSetAttribute(sim_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kSmemBytes));
  int grid = sm_count();
  if (grid > p.units) grid = p.units;
  if (grid < 1) return HIPPO_OK;
  sim_tc_kernel<EPI><<<grid, kTcThreads, kSmemBytes, s>>>(tmA, tmB, p);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

int tc_topk_splits(int64_t n, int nq) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const int m_blocks = (nq + kTcBM - 1) / kTcBM;
  const int64_t n_tiles = (n + kTcBN - 1) / kTcBN;
  if (n_tiles <= 0 || m_blocks <= 0) return 1;
  // aim for units = m_blocks * splits ~ a multiple of the SM count, up to 8 waves,
  // while keeping at least 8 tiles per split (the per-unit list start-up is not free)
  int64_t best = 1;
  double best_cost = 1e300;
  for (int waves = 1; waves <= 8; ++waves) {
    int64_t splits = ((int64_t)sms * waves + m_blocks - 1) / m_blocks;
    if (splits > n_tiles) splits = n_tiles;
    if (splits < 1) splits = 1;
    const int64_t tps = (n_tiles + splits - 1) / splits;
    const int64_t real_splits = (n_tiles + tps - 1) / tps;
    const int64_t units = real_splits * m_blocks;
    const int64_t rounds = (units + sms - 1) / sms;
    // cost = critical-path tiles + a small per-unit overhead
    const double cost = (double)rounds * (double)(tps + 2);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = real_splits; }
  }
  return (int)best;
}

hippo_status tc_topk_launch(const TcTopkArgs& a, cudaStream_t s) {
  CUtensorMap tmA, tmB;
  hippo_status st = make_tmap(&tmA, a.qbf16, a.nq, a.d, kTcBM);
  if (st != HIPPO_OK) return st;
  st = make_tmap(&tmB, a.bank, a.n, a.d, kTcBN);
  if (st != HIPPO_OK) return st;
  TcParams p{};
  p.n = a.n;
  p.kblocks = a.d / kTcBK;
  p.bnorm = a.bnorm;
  p.qnorm = a.qnorm;
  p.nq = a.nq;
  p.k = a.k;
  p.m_blocks = (a.nq + kTcBM - 1) / kTcBM;
  p.n_tiles = (int)((a.n + kTcBN - 1) / kTcBN);
  p.tiles_per_split = (p.n_tiles + a.splits - 1) / a.splits;
  p.splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  if (p.splits != a.splits) { set_error("tc_topk_launch: inconsistent split count"); return HIPPO_E_BADARG; }
  p.units = p.m_blocks * p.splits;
  p.row_base = a.row_base;
  p.after_key = a.after_key;
  p.part = a.part;
  p.thr_ord = a.thr_ord;
  return launch<EPI_TOPK>(tmA, tmB, p, s);
}

hippo_status tc_mask_launch(const TcMaskArgs& a, cudaStream_t s) {
  CUtensorMap tmA, tmB;
  hippo_status st = make_tmap(&tmA, a.feats_bf16, a.n, a.d, kTcBM);
  if (st != HIPPO_OK) return st;
  st = make_tmap(&tmB, a.feats_bf16, a.n, a.d, kTcBN);
  if (st != HIPPO_OK) return st;
  TcParams p{};
  p.n = a.n;
  p.kblocks = a.d / kTcBK;
  p.bnorm = a.norm;
  const int64_t nI = (a.n + kTcBN - 1) / kTcBN;
  const int64_t units = nI * (nI + 1) / 2;
  if (units > 0x7fffffffll) { set_error("tc_mask_launch: n too large"); return HIPPO_E_BADARG; }
  p.units = (int)units;
  p.gamma = a.gamma;
  p.band_exact = a.band_exact;
  p.band_inexact = a.band_inexact;
  p.inexact = a.inexact;
  p.mask = a.mask;
  p.words_per_row = a.words_per_row;
  p.uncertain = a.uncertain;
  p.uncertain_count = a.uncertain_count;
  p.uncertain_cap = a.uncertain_cap;
  return launch<EPI_MASK>(tmA, tmB, p, s);
}

}  // namespace hippo
