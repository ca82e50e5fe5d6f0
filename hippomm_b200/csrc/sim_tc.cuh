// Interface of the tcgen05 similarity contraction (sim_tc.cu).
#pragma once
#include "common.cuh"

namespace hippo {

constexpr int kTcBM = 128;      // query / row-i block per CTA (TMEM lanes); a CTA pair covers 256
constexpr int kTcBN = 256;      // bank / row-j block (TMEM columns); each CTA of the pair stages half
constexpr int kTcBK = 64;       // bf16 elements per 128-byte swizzle row
constexpr int kTcStages = 6;
constexpr int kTcThreads = 320; // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue

struct TcTopkArgs {
  const void* bank;          // [n, d] bf16
  const float* bnorm;        // [n]
  int64_t n;
  int d;
  const void* qbf16;         // [nq_rows, d] bf16
  const float* qnorm;        // [nq]
  int nq;
  int nq_rows;               // rows of qbf16 that exist (>= nq, zero rows beyond nq): extent of the TMA tensor map
  int k;
  int64_t row_base;
  const uint64_t* after_key; // [nq] or null
  uint64_t* part;            // [2 * splits, nq, k] (one list per bank split and column half)
  uint32_t* thr_ord;         // [nq], zero-initialised by the caller
  uint32_t* pool;            // [nq, k], zero-initialised by the caller
  unsigned long long* counters;  // [8] profiling counters (only touched with HIPPO_TC_DEBUG & 64)
  uint32_t* progress;        // [256-query blocks * splits], zero-initialised by the caller (flow control)
  int splits;                // bank splits (units = 256-query blocks * splits)
};
// number of bank splits the launch will use for (n, nq) on this device
int tc_topk_splits(int64_t n, int nq);
hippo_status tc_topk_launch(const TcTopkArgs& a, cudaStream_t s);

// Similarity decisions of one band of rows (consolidate.cu).  A = rows [a_row0, a_row0 + na) of a_rows; a pair's bit
// is !(sim < gamma); pairs too close to gamma for bf16 inputs go to the `uncertain` list instead.
//   rect == false: B = the same band (b_rows = a_rows, b_row0 = a_row0): lower triangle, bit (i, j) for j < i into
//                  the band-local bit matrix `mask`
//   rect == true : B = rows [b_row0, b_row0 + *dyn_k) of b_rows (the rows kept so far), all columns below *dyn_k;
//                  only the OR over a band row's bits is kept (`rowhit`)
struct TcMaskArgs {
  const void* a_rows;        // [a_total, d] bf16
  int64_t a_total;
  const void* b_rows;        // [b_total, d] bf16
  int64_t b_total;
  int d;
  bool rect;
  int64_t a_row0, b_row0;
  int na;                    // rows of the band
  const float* anorm;        // [na] norms of the band's rows
  const float* bnorm;        // norms of the B rows, bnorm[0] belongs to row b_row0
  const int32_t* dyn_k;      // rectangle: device count of valid B rows
  float gamma;
  float band_exact, band_inexact;
  const int32_t* inexact;    // device flag from hippo_bank_build
  uint32_t* mask;            // triangle: band-local bit matrix (rows with 8 words per 256-column tile)
  int64_t words_per_row;
  int32_t* rowhit;           // rectangle: [na], zeroed by the caller; 1 = the band row conflicts with a kept row
  uint2* uncertain;          // (band row, column) pairs within the band of gamma
  int32_t* uncertain_count;  // zero-initialised by the caller
  int32_t uncertain_cap;
  int max_pairs;             // > 0: use at most this many CTA pairs (leave SMs to kernels running beside this one)
};
hippo_status tc_mask_launch(const TcMaskArgs& a, cudaStream_t s);

}  // namespace hippo
