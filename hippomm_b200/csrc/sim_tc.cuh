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

struct TcMaskArgs {
  const void* feats_bf16;    // [n, d] bf16
  const float* norm;         // [n]
  int64_t n;
  int d;
  float gamma;
  float band_exact, band_inexact;
  const int32_t* inexact;    // device flag from hippo_bank_build
  uint32_t* mask;            // [n, words_per_row] bit j of row i: !(sim(i,j) < gamma), j < i
  int64_t words_per_row;
  uint2* uncertain;          // (i, j) pairs within the band
  int32_t* uncertain_count;  // zero-initialised by the caller
  int32_t uncertain_cap;
  const int32_t* dyn_k;      // banded mode (consolidate.cu): device count of final rows; null = whole triangle
  int band_rows;             // rows of this band (A rows [*dyn_k, *dyn_k + band_rows))
};
hippo_status tc_mask_launch(const TcMaskArgs& a, cudaStream_t s);

}  // namespace hippo
