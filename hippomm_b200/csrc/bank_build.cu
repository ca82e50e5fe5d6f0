// hippo_bank_build: caller rows -> device bank (bf16 rows + fp32 row norms).
//
// Replaces the per-call row norms of the reference (vo:179 `np.linalg.norm(b, axis=1)`,
// recomputed for the whole bank on every query; hm:951 row normalisation) with a
// one-off streaming pass: one warp per row, 8 elements per lane per step,
// 128-bit loads and stores, squares accumulated in fp64 and reduced by shuffles.
#include "common.cuh"

namespace hippo {

template <typename T> struct Load8;
template <> struct Load8<float> {
  __device__ static void load(const float* p, float (&v)[8]) {
    uint4 a = ldg_stream(p), b = ldg_stream(p + 4);
    v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
    v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
  }
};

// rows [n, d] T (stride ld) -> bank [n, d] bf16, norm [n] fp32
template <typename T>
__global__ void __launch_bounds__(256) bank_build_kernel(const T* __restrict__ rows, int64_t n, int d,
                                                         int64_t ld, __nv_bfloat16* __restrict__ bank,
                                                         float* __restrict__ norm,
                                                         int32_t* __restrict__ inexact) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  bool changed = false;
  for (int64_t r = warp; r < n; r += nwarps) {
    const T* src = rows + r * ld;
    __nv_bfloat16* dst = bank + r * (int64_t)d;
    double ss = 0.0;
    for (int c = lane * 8; c < d; c += 256) {
      float v[8];
      double sq = 0.0;
      if constexpr (sizeof(T) == 4) {
        if ((((uintptr_t)(src + c)) & 15) == 0) {
          Load8<float>::load((const float*)(src + c), v);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (float)src[c + j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sq += (double)v[j] * (double)v[j];
      } else if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double x = (double)src[c + j];
          sq += x * x;
          v[j] = (float)x;
          // a value that is not fp32-representable is certainly not bf16-exact
          changed |= ((double)v[j] != x);
        }
      } else {  // bf16 source
        uint4 w = *reinterpret_cast<const uint4*>(src + c);
        v[0] = bf16lo(w.x); v[1] = bf16hi(w.x); v[2] = bf16lo(w.y); v[3] = bf16hi(w.y);
        v[4] = bf16lo(w.z); v[5] = bf16hi(w.z); v[6] = bf16lo(w.w); v[7] = bf16hi(w.w);
#pragma unroll
        for (int j = 0; j < 8; ++j) sq += (double)v[j] * (double)v[j];
      }
      ss += sq;
      uint32_t packed[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat16 lo = __float2bfloat16_rn(v[2 * j]);
        __nv_bfloat16 hi = __float2bfloat16_rn(v[2 * j + 1]);
        changed |= (__bfloat162float(lo) != v[2 * j]) | (__bfloat162float(hi) != v[2 * j + 1]);
        packed[j] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      }
      *reinterpret_cast<uint4*>(dst + c) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
    ss = warp_sum(ss);
    if (lane == 0) norm[r] = __fsqrt_rn((float)ss);
  }
  if (inexact != nullptr && __any_sync(0xffffffffu, changed)) {
    if (lane == 0) atomicOr(inexact, 1);
  }
}

}  // namespace hippo

extern "C" hippo_status hippo_bank_build(const void* rows, int32_t dtype, int64_t n, int32_t d,
                                         int64_t ld, void* bank, float* norm, int32_t* inexact,
                                         void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(n >= 0 && d > 0 && d % 64 == 0, "hippo_bank_build: need n >= 0 and d %% 64 == 0 (n=%lld d=%d)",
                (long long)n, d);
  HIPPO_REQUIRE(ld >= d, "hippo_bank_build: row stride %lld < d %d", (long long)ld, d);
  if (n == 0) return HIPPO_OK;
  HIPPO_REQUIRE(rows && bank && norm, "hippo_bank_build: null pointer");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks64 = (n + 7) / 8;
  int blocks = (int)(blocks64 < (int64_t)sm_count() * 16 ? blocks64 : (int64_t)sm_count() * 16);
  switch (dtype) {
    case HIPPO_F32:
      bank_build_kernel<float><<<blocks, 256, 0, s>>>((const float*)rows, n, d, ld,
                                                      (__nv_bfloat16*)bank, norm, inexact);
      break;
    case HIPPO_F64:
      bank_build_kernel<double><<<blocks, 256, 0, s>>>((const double*)rows, n, d, ld,
                                                       (__nv_bfloat16*)bank, norm, inexact);
      break;
    case HIPPO_BF16:
      HIPPO_REQUIRE(ld % 8 == 0 && (((uintptr_t)rows) & 15) == 0,
                    "hippo_bank_build: bf16 rows need 16-byte aligned rows");
      bank_build_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)rows, n, d, ld,
                                                              (__nv_bfloat16*)bank, norm, inexact);
      break;
    default:
      HIPPO_REQUIRE(false, "hippo_bank_build: unsupported dtype %d", dtype);
  }
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}
