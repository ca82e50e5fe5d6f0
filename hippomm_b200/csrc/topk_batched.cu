// hippo_topk_batched: nq queries against the bank in one tcgen05 pass (see sim_tc.cu).
// Reference: nq sequential calls of top_k_cosine_similarity (vo:151-188).
#include "common.cuh"
#include "exchange.cuh"
#include "sim_tc.cuh"

namespace hippo {

// topk_single.cu: the GEMV with two accumulators per row, for a batch too small for the tensor cores
bool topk_few_supported(int d, int nq);
size_t topk_few_part_elems(int nq, int k);
hippo_status topk_few_launch(const void* bank, const float* norm, int64_t n, const float* q, int nq, int k,
                             int64_t row_base, const uint64_t* after_key, uint64_t* part, int* nparts_out,
                             cudaStream_t s);

// topk_small.cu: 3 .. 64 queries, bank rows on the M side of the MMA, deeper bank pipeline
bool topk_small_supported(int d, int nq);
int topk_small_grid(int64_t n);
hippo_status topk_small_launch(const void* bank, const float* norm, int64_t n, int d, const void* qbf, const float* qnorm,
                               int nq, int k, int64_t row_base, const uint64_t* after_key, uint64_t* part, uint32_t* gthr,
                               int* nparts_out, cudaStream_t s);

struct BatchedLayout {
  __nv_bfloat16* qbf;
  float* qnorm;
  uint32_t* thr_ord;   // [nq] followed by the score pool [nq, k] and the per-unit progress counters
  uint32_t* pool;      // (one memset clears all three)
  uint32_t* progress;
  size_t clear_words;
  unsigned long long* counters;   // [8], first thing in the workspace so a profiling script can find it
  uint64_t* part;
  int splits;
  int nq_pad;
  size_t bytes;
};

static BatchedLayout batched_layout(void* ws, size_t ws_bytes, int64_t n, int d, int nq, int k) {
  Carver c(ws, ws_bytes);
  BatchedLayout L{};
  L.counters = c.take<unsigned long long>(8);
  // padded with zero rows to whole 256-query blocks: TMA boxes that are partly or wholly out of bounds load at a
  // fraction of the normal rate (8 queries took 0.75 ms where 256 took 0.41 ms on a 1M-row bank)
  L.nq_pad = (nq + 2 * kTcBM - 1) / (2 * kTcBM) * (2 * kTcBM);
  L.qbf = c.take<__nv_bfloat16>((size_t)L.nq_pad * d);
  L.qnorm = c.take<float>((size_t)nq);
  L.splits = tc_topk_splits(n, nq);
  const size_t units = (size_t)((nq + 2 * kTcBM - 1) / (2 * kTcBM)) * (size_t)L.splits;
  L.clear_words = (size_t)nq * (1 + (size_t)k) + units;
  L.thr_ord = c.take<uint32_t>(L.clear_words);
  L.pool = L.thr_ord ? L.thr_ord + nq : nullptr;
  L.progress = L.thr_ord ? L.thr_ord + (size_t)nq * (1 + (size_t)k) : nullptr;
  size_t part_elems = (size_t)2 * L.splits * nq * k;
  if (topk_few_supported(d, nq) && topk_few_part_elems(nq, k) > part_elems) part_elems = topk_few_part_elems(nq, k);
  if (topk_small_supported(d, nq)) {
    const size_t need = (size_t)topk_small_grid(n) * nq * k;
    if (need > part_elems) part_elems = need;
  }
  L.part = c.take<uint64_t>(part_elems);
  L.bytes = c.used();
  return L;
}

// The local search up to (not including) the merge of the per-split lists: *part_out = [nparts][nq][k] order keys.
static hippo_status batched_parts(const void* bank, const float* norm, int64_t n, int32_t d, const float* q, int32_t nq,
                                  int32_t k, int64_t row_base, const uint64_t* after_key, void* ws, size_t ws_bytes,
                                  void* stream, uint64_t** part_out, int* nparts_out, const char* who) {
  HIPPO_REQUIRE(n >= 0 && d > 0 && d % 64 == 0, "%s: need d %% 64 == 0 (d=%d)", who, d);
  HIPPO_REQUIRE(nq >= 1, "%s: nq < 1", who);
  HIPPO_REQUIRE(k >= 1 && k <= HIPPO_TOPK_MAX, "%s: k=%d outside 1..%d", who, k, HIPPO_TOPK_MAX);
  HIPPO_REQUIRE(row_base >= 0 && row_base + n < 0xffffffffll, "%s: global row numbers must stay below 2^32-1", who);
  HIPPO_REQUIRE(q != nullptr && (n == 0 || (bank && norm)), "%s: null pointer", who);
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  cudaStream_t s = (cudaStream_t)stream;
  BatchedLayout L = batched_layout(ws, ws_bytes, n, d, nq, k);
  if (ws == nullptr || ((uintptr_t)ws & 255) || L.bytes > ws_bytes) {
    set_error("%s: workspace of %zu bytes needed (256-byte aligned), got %zu", who, L.bytes, ws_bytes);
    return HIPPO_E_WORKSPACE;
  }
  *part_out = L.part;
  if (n == 0) {
    HIPPO_CUDA(cudaMemsetAsync(L.part, 0, (size_t)nq * k * 8, s));
    *nparts_out = 1;
    return HIPPO_OK;
  }
  if (topk_few_supported(d, nq))   // two queries: one GEMV pass with two accumulators per row (topk_single.cu)
    return topk_few_launch(bank, norm, n, q, nq, k, row_base, after_key, L.part, nparts_out, s);
  HIPPO_CUDA(cudaMemsetAsync(L.thr_ord, 0, L.clear_words * 4, s));
  // queries -> bf16 + |a| (same pass the bank went through)
  st = hippo_bank_build(q, HIPPO_F32, nq, d, d, L.qbf, L.qnorm, nullptr, stream);
  if (st != HIPPO_OK) return st;
  if (L.nq_pad > nq) HIPPO_CUDA(cudaMemsetAsync(L.qbf + (size_t)nq * d, 0, (size_t)(L.nq_pad - nq) * d * 2, s));
  if (topk_small_supported(d, nq))   // 3 .. 64 queries: HBM bound, bank rows on the M side (topk_small.cu)
    return topk_small_launch(bank, norm, n, d, L.qbf, L.qnorm, nq, k, row_base, after_key, L.part, L.thr_ord, nparts_out, s);
  TcTopkArgs a{};
  a.bank = bank;
  a.bnorm = norm;
  a.n = n;
  a.d = d;
  a.qbf16 = L.qbf;
  a.qnorm = L.qnorm;
  a.nq = nq;
  a.nq_rows = L.nq_pad;
  a.k = k;
  a.row_base = row_base;
  a.after_key = after_key;
  a.part = L.part;
  a.thr_ord = L.thr_ord;
  a.pool = L.pool;
  a.counters = L.counters;
  a.progress = L.progress;
  a.splits = L.splits;
  *nparts_out = 2 * L.splits;
  return tc_topk_launch(a, s);
}

}  // namespace hippo

extern "C" {

size_t hippo_topk_batched_workspace_bytes(int64_t n, int32_t d, int32_t nq, int32_t k) {
  if (n < 0 || d <= 0 || nq <= 0 || k <= 0) return 256;
  return hippo::batched_layout(nullptr, 0, n, d, nq, k).bytes;
}

hippo_status hippo_topk_batched(const void* bank, const float* norm, int64_t n, int32_t d, const float* q,
                                int32_t nq, int32_t k, int64_t row_base, const uint64_t* after_key,
                                int64_t* out_idx, float* out_score, uint64_t* out_key, void* ws,
                                size_t ws_bytes, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(nq >= 0, "hippo_topk_batched: nq < 0");
  if (nq == 0) return HIPPO_OK;
  uint64_t* part = nullptr;
  int nparts = 0;
  hippo_status st = batched_parts(bank, norm, n, d, q, nq, k, row_base, after_key, ws, ws_bytes, stream, &part, &nparts,
                                  "hippo_topk_batched");
  if (st != HIPPO_OK) return st;
  return hippo_topk_merge(part, nparts, nq, k, k, out_idx, out_score, out_key, stream);
}

hippo_status hippo_topk_batched_sharded(const void* bank, const float* norm, int64_t n, int32_t d, const float* q,
                                        int32_t nq, int32_t k, int64_t row_base, const uint64_t* after_key,
                                        void* const* peer_bases, size_t buf_bytes, int32_t rank, int32_t world,
                                        uint32_t epoch, int64_t* out_idx, float* out_score, uint64_t* out_key,
                                        void* ws, size_t ws_bytes, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(nq >= 0, "hippo_topk_batched_sharded: nq < 0");
  if (nq == 0) return HIPPO_OK;
  uint64_t* part = nullptr;
  int nparts = 0;
  hippo_status st = batched_parts(bank, norm, n, d, q, nq, k, row_base, after_key, ws, ws_bytes, stream, &part, &nparts,
                                  "hippo_topk_batched_sharded");
  if (st != HIPPO_OK) return st;
  // the per-split lists are merged inside the exchange kernel, just before its push phase: one launch for
  // local merge + NVLink push + global merge
  return exchange_launch(part, nparts, nq, k, k, peer_bases, buf_bytes, rank, world, epoch, out_idx, out_score, out_key,
                         (cudaStream_t)stream);
}

}  // extern "C"
