// hippo_topk_rows / hippo_rescore: cosine top-k straight from the caller's fp32 / fp64 rows.
//
// Reference: top_k_cosine_similarity (vo:151-188) receives the (N, D) feature array itself on every call
// and computes `np.dot(b, a) / (np.linalg.norm(b, axis=1) * np.linalg.norm(a))` in the arrays' own
// precision (fp32 for fresh features, fp64 for ThetaEvents reloaded from JSON, hm:391).  The bf16 bank of
// hippo_bank_build is the right layout for a store that is searched many times; for a feature array that
// arrives with the call (the drop-in signature) a bank would cost a read of n*d*4 bytes, a write of n*d*2
// and a second read of n*d*2 -- and round the rows.  hippo_topk_rows is ONE streaming pass over the rows
// as they are: dot product and sum of squares of a row from the same registers, accumulated in fp64
// (products of fp32 values are exact in fp64), rounded once to the arrays' precision, then the
// reference's operation order dot / (|b| * |a|).  No rounding of the inputs anywhere, so the scores agree
// with NumPy's to the last few ulp of its own BLAS summation order.
//
// hippo_rescore evaluates the same expression for a short list of candidate rows per query: the exact
// second stage behind the bf16 tensor-core search (MemoryBank(keep_rows=True).search(exact=True)).
#include "common.cuh"

namespace hippo {

constexpr int kRowsThreads = 256;
constexpr int kRowsWarps = kRowsThreads / 32;

__device__ __forceinline__ void ld4(const float* p, double (&v)[4]) {
  const uint4 a = ldg_stream(p);
  v[0] = (double)__uint_as_float(a.x); v[1] = (double)__uint_as_float(a.y);
  v[2] = (double)__uint_as_float(a.z); v[3] = (double)__uint_as_float(a.w);
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
  const uint4 a = ldg_stream(p), b = ldg_stream(p + 2);
  v[0] = __hiloint2double((int)a.y, (int)a.x); v[1] = __hiloint2double((int)a.w, (int)a.z);
  v[2] = __hiloint2double((int)b.y, (int)b.x); v[3] = __hiloint2double((int)b.w, (int)b.z);
}

// |x| in the precision NumPy would have produced it in: fp32 sqrt of the fp32-rounded sum for fp32 data
template <typename T> __device__ __forceinline__ double norm_as(double ss) {
  if constexpr (sizeof(T) == 4) return (double)__fsqrt_rn((float)ss);
  else return sqrt(ss);
}
// dot / (|b| * |a|): fp32 IEEE chain when both operands are fp32 (vo:182), fp64 otherwise (NumPy promotes)
template <typename T, typename Q>
__device__ __forceinline__ double cosine_as(double dot, double bn, double an) {
  if constexpr (sizeof(T) == 4 && sizeof(Q) == 4) return (double)__fdiv_rn((float)dot, __fmul_rn((float)bn, (float)an));
  else return dot / (bn * an);
}

// dot(row, q) and sum of squares of the row, all lanes of the warp return both
template <typename T>
__device__ __forceinline__ void row_dot(const T* __restrict__ row, const double* __restrict__ sq, int d, bool vec,
                                        int lane, double& dot, double& ss) {
  double a0 = 0.0, a1 = 0.0, s0 = 0.0, s1 = 0.0;
  if (vec) {
    int c = lane * 4;
    for (; c + 128 < d; c += 256) {          // two independent 4-element groups per step
      double x[4], y[4];
      ld4(row + c, x); ld4(row + c + 128, y);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a0 = fma(x[j], sq[c + j], a0); s0 = fma(x[j], x[j], s0);
        a1 = fma(y[j], sq[c + 128 + j], a1); s1 = fma(y[j], y[j], s1);
      }
    }
    if (c < d) {
      double x[4];
      ld4(row + c, x);
#pragma unroll
      for (int j = 0; j < 4; ++j) { a0 = fma(x[j], sq[c + j], a0); s0 = fma(x[j], x[j], s0); }
    }
  } else {
    for (int c = lane; c < d; c += 32) {
      const double x = (double)row[c];
      a0 = fma(x, sq[c], a0); s0 = fma(x, x, s0);
    }
  }
  dot = warp_sum(a0 + a1);
  ss = warp_sum(s0 + s1);
}

// stage the query as fp64 in shared memory and return |a| (in the query's own precision)
template <typename Q>
__device__ __forceinline__ double stage_query(const Q* __restrict__ q, int d, double* sq, double* s_red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double ss = 0.0;
  for (int e = threadIdx.x; e < d; e += blockDim.x) {
    const double v = (double)q[e];
    sq[e] = v;
    ss = fma(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) s_red[wid] = ss;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
  return norm_as<Q>(t);
}

template <typename T, typename Q>
__global__ void __launch_bounds__(kRowsThreads) topk_rows_kernel(
    const T* __restrict__ rows, int64_t n, int d, int64_t ld, const Q* __restrict__ q, int k, int64_t row_base,
    const uint64_t* __restrict__ after_key, uint64_t* __restrict__ part /*[grid][k]*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sq = reinterpret_cast<double*>(smem_raw);                                 // [d]
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + (size_t)d * 8);          // [warps][k]
  __shared__ double s_red[kRowsWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kRowsWarps * k; i += kRowsThreads) lists[i] = 0;
  const double an = stage_query<Q>(q, d, sq, s_red);
  __syncthreads();
  const uint64_t below = after_key ? after_key[0] : ~0ull;
  uint64_t* mylist = lists + wid * k;
  const bool vec = (d % 4 == 0) && (ld % 4 == 0) && ((((uintptr_t)rows) & 15) == 0);

  const int64_t gwarp = (int64_t)blockIdx.x * kRowsWarps + wid;
  const int64_t gstride = (int64_t)gridDim.x * kRowsWarps;
  for (int64_t r = gwarp; r < n; r += gstride) {
    double dot, ss;
    row_dot<T>(rows + r * ld, sq, d, vec, lane, dot, ss);
    if (lane == 0) {
      const float sc = (float)cosine_as<T, Q>(dot, norm_as<T>(ss), an);
      const uint64_t key = pack_key(sc, (uint32_t)(row_base + r));
      if (key < below && key > mylist[k - 1]) topk_insert(mylist, k, key);
    }
  }
  __syncthreads();
  if (wid == 0) {      // block merge: the k best of kRowsWarps * k candidates by repeated selection
    uint64_t prev = ~0ull;
    for (int r = 0; r < k; ++r) {
      uint64_t best = 0;
      if (prev != 0) {
        for (int i = lane; i < kRowsWarps * k; i += 32) {
          const uint64_t c = lists[i];
          if (c < prev && c > best) best = c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
          best = other > best ? other : best;
        }
      }
      if (lane == 0) part[(size_t)blockIdx.x * k + r] = best;
      prev = best;
    }
  }
}

// One warp per (query, candidate): exact cosine of candidate row cand[qi][j] with query qi.
template <typename T, typename Q>
__global__ void __launch_bounds__(kRowsThreads) rescore_kernel(
    const T* __restrict__ rows, int64_t n, int d, int64_t ld, int64_t row_base, const Q* __restrict__ q, int64_t q_ld,
    int nq, const int64_t* __restrict__ cand, int kc, uint64_t* __restrict__ out_key, double* __restrict__ out_score) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * kRowsWarps + (threadIdx.x >> 5);
  if (w >= (int64_t)nq * kc) return;
  const int qi = (int)(w / kc);
  const int64_t grow = cand[w];
  const int64_t r = grow - row_base;
  if (grow < 0 || r < 0 || r >= n) {
    if (lane == 0) { if (out_key) out_key[w] = 0; if (out_score) out_score[w] = 0.0; }
    return;
  }
  const T* row = rows + r * ld;
  const Q* qv = q + (size_t)qi * q_ld;
  double a = 0.0, s = 0.0, t = 0.0;
  for (int c = lane; c < d; c += 32) {
    const double x = (double)row[c], y = (double)qv[c];
    a = fma(x, y, a); s = fma(x, x, s); t = fma(y, y, t);
  }
  a = warp_sum(a); s = warp_sum(s); t = warp_sum(t);
  if (lane == 0) {
    const double sc = cosine_as<T, Q>(a, norm_as<T>(s), norm_as<Q>(t));
    if (out_key) out_key[w] = pack_key((float)sc, (uint32_t)grow);
    if (out_score) out_score[w] = sc;
  }
}

static int rows_grid(int64_t n) {
  int64_t want = (n + kRowsWarps - 1) / kRowsWarps;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace hippo

extern "C" {

size_t hippo_topk_rows_workspace_bytes(int64_t n, int32_t d, int32_t k) {
  (void)n; (void)d;
  int sms = hippo::sm_count();
  if (sms <= 0) sms = 148;
  return hippo::align_up((size_t)sms * 8 * (size_t)(k > 0 ? k : 1) * 8, 256);
}

hippo_status hippo_topk_rows(const void* rows, int32_t dtype, int64_t n, int32_t d, int64_t ld, const void* q,
                             int32_t q_dtype, int32_t k, int64_t row_base, const uint64_t* after_key,
                             int64_t* out_idx, float* out_score, uint64_t* out_key, void* ws, size_t ws_bytes,
                             void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(n >= 0 && d > 0 && ld >= d, "hippo_topk_rows: bad shape (n=%lld d=%d ld=%lld)", (long long)n, d, (long long)ld);
  HIPPO_REQUIRE((dtype == HIPPO_F32 || dtype == HIPPO_F64) && (q_dtype == HIPPO_F32 || q_dtype == HIPPO_F64),
                "hippo_topk_rows: rows and query must be fp32 or fp64");
  HIPPO_REQUIRE(k >= 1 && k <= HIPPO_TOPK_MAX, "hippo_topk_rows: k=%d outside 1..%d", k, HIPPO_TOPK_MAX);
  HIPPO_REQUIRE(row_base >= 0 && row_base + n < 0xffffffffll, "hippo_topk_rows: global row numbers must stay below 2^32-1");
  HIPPO_REQUIRE(q != nullptr && (n == 0 || rows != nullptr), "hippo_topk_rows: null pointer");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = rows_grid(n);
  if (ws == nullptr || ws_bytes < (size_t)grid * k * 8 || ((uintptr_t)ws & 255)) {
    set_error("hippo_topk_rows: workspace of %zu bytes needed (256-byte aligned)", hippo_topk_rows_workspace_bytes(n, d, k));
    return HIPPO_E_WORKSPACE;
  }
  uint64_t* part = (uint64_t*)ws;
  if (n == 0) {
    HIPPO_CUDA(cudaMemsetAsync(part, 0, (size_t)k * 8, s));
    return hippo_topk_merge(part, 1, 1, k, k, out_idx, out_score, out_key, stream);
  }
  const size_t smem = (size_t)d * 8 + (size_t)kRowsWarps * k * 8;
  HIPPO_REQUIRE(smem <= 200 * 1024, "hippo_topk_rows: d=%d too large", d);
#define HIPPO_ROWS_LAUNCH(T, Q)                                                                                       \
  do {                                                                                                                \
    HIPPO_CUDA(cudaFuncSetAttribute(topk_rows_kernel<T, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    topk_rows_kernel<T, Q><<<grid, kRowsThreads, smem, s>>>((const T*)rows, n, d, ld, (const Q*)q, k, row_base,       \
                                                            after_key, part);                                         \
  } while (0)
  if (dtype == HIPPO_F32 && q_dtype == HIPPO_F32) HIPPO_ROWS_LAUNCH(float, float);
  else if (dtype == HIPPO_F32) HIPPO_ROWS_LAUNCH(float, double);
  else if (q_dtype == HIPPO_F32) HIPPO_ROWS_LAUNCH(double, float);
  else HIPPO_ROWS_LAUNCH(double, double);
#undef HIPPO_ROWS_LAUNCH
  HIPPO_CUDA(cudaGetLastError());
  return hippo_topk_merge(part, grid, 1, k, k, out_idx, out_score, out_key, stream);
}

hippo_status hippo_rescore(const void* rows, int32_t dtype, int64_t n, int32_t d, int64_t ld, int64_t row_base,
                           const void* q, int32_t q_dtype, int64_t q_ld, int32_t nq, const int64_t* cand_idx,
                           int32_t kc, uint64_t* out_key, double* out_score, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(n >= 0 && d > 0 && ld >= d && nq >= 0 && kc >= 1 && (q_ld == 0 || q_ld >= d), "hippo_rescore: bad sizes");
  HIPPO_REQUIRE((dtype == HIPPO_F32 || dtype == HIPPO_F64) && (q_dtype == HIPPO_F32 || q_dtype == HIPPO_F64),
                "hippo_rescore: rows and queries must be fp32 or fp64");
  if (nq == 0) return HIPPO_OK;
  HIPPO_REQUIRE(q && cand_idx && (n == 0 || rows), "hippo_rescore: null pointer");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t warps = (int64_t)nq * kc;
  const unsigned grid = (unsigned)((warps + kRowsWarps - 1) / kRowsWarps);
#define HIPPO_RESCORE_LAUNCH(T, Q)                                                                                \
  rescore_kernel<T, Q><<<grid, kRowsThreads, 0, s>>>((const T*)rows, n, d, ld, row_base, (const Q*)q, q_ld, nq, \
                                                     cand_idx, kc, out_key, out_score)
  if (dtype == HIPPO_F32 && q_dtype == HIPPO_F32) HIPPO_RESCORE_LAUNCH(float, float);
  else if (dtype == HIPPO_F32) HIPPO_RESCORE_LAUNCH(float, double);
  else if (q_dtype == HIPPO_F32) HIPPO_RESCORE_LAUNCH(double, float);
  else HIPPO_RESCORE_LAUNCH(double, double);
#undef HIPPO_RESCORE_LAUNCH
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

}  // extern "C"
