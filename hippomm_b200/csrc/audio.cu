// hippo_audio_energy / hippo_audio_levels: RMS / silence detection over long audio streams.
//
// Reference: _compute_audio_level (hm:993-1000), called on 0.5 s windows whose positions
// depend on the previous segment boundary (hm:1061-1077), so windows cannot be pre-binned
// on a fixed grid.  One streaming pass builds a two-level pyramid of fp64 sums of squares
// (16- and 512-sample blocks); any window is then a short sum of pyramid entries plus at
// most 30 edge samples.
#include "audio.cuh"

namespace hippo {

// Mono fast path: one warp per 512-sample block, every lane issues 128-bit coalesced loads.
// VEC = samples per 16-byte load; 16/VEC neighbouring lanes share one 16-sample block.
template <typename T>
__global__ void __launch_bounds__(256) audio_energy_mono_kernel(const T* __restrict__ pcm, int64_t ns,
                                                                double* __restrict__ e16,
                                                                double* __restrict__ e512) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int LOADS = 512 / (32 * VEC);
  constexpr int LANES_PER_16 = 16 / VEC;
  const int lane = threadIdx.x & 31;
  const int64_t nblk512 = (ns + 511) / 512;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t blk = warp; blk < nblk512; blk += nwarps) {
    const int64_t base = blk * 512;
    double part[LOADS];
    if (base + 512 <= ns) {
      uint4 raw[LOADS];
#pragma unroll
      for (int j = 0; j < LOADS; ++j)
        raw[j] = ldg_stream(reinterpret_cast<const uint4*>(pcm + base) + j * 32 + lane);
#pragma unroll
      for (int j = 0; j < LOADS; ++j) {
        double s = 0.0;
        if constexpr (sizeof(T) == 4) {
          const float v[4] = {__uint_as_float(raw[j].x), __uint_as_float(raw[j].y),
                              __uint_as_float(raw[j].z), __uint_as_float(raw[j].w)};
#pragma unroll
          for (int i = 0; i < 4; ++i) { const double x = (double)v[i]; s += x * x; }
        } else if constexpr (sizeof(T) == 2) {
          const uint32_t wv[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double a = (double)(int16_t)(wv[i] & 0xffffu) * (1.0 / 32768.0);
            const double b = (double)(int16_t)(wv[i] >> 16) * (1.0 / 32768.0);
            s += a * a; s += b * b;
          }
        } else {
          const double a = __hiloint2double((int)raw[j].y, (int)raw[j].x);
          const double b = __hiloint2double((int)raw[j].w, (int)raw[j].z);
          s = a * a + b * b;
        }
        part[j] = s;
      }
    } else {
#pragma unroll
      for (int j = 0; j < LOADS; ++j) {
        double s = 0.0;
        const int64_t i0 = base + (int64_t)(j * 32 + lane) * VEC;
        for (int i = 0; i < VEC; ++i)
          if (i0 + i < ns) { const double x = pcm_raw(pcm, sizeof(T) == 4 ? HIPPO_F32 : sizeof(T) == 2 ? HIPPO_I16 : HIPPO_F64, i0 + i); s += x * x; }
        part[j] = s;
      }
    }
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < LOADS; ++j) {
      double s = part[j];
#pragma unroll
      for (int o = 1; o < LANES_PER_16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const int64_t b16 = blk * 32 + (j * 32 + lane) / LANES_PER_16;
      if ((lane % LANES_PER_16) == 0 && b16 * 16 < ns) e16[b16] = s;
      tot += part[j];
    }
    tot = warp_sum(tot);
    if (lane == 0) e512[blk] = tot;
  }
}

// Generic path (any channel count): lane l of a warp owns 16-block l of the warp's 512-block.
__global__ void __launch_bounds__(256) audio_energy_generic_kernel(const void* __restrict__ pcm, int dtype,
                                                                   int64_t ns, int nch,
                                                                   double* __restrict__ e16,
                                                                   double* __restrict__ e512) {
  const int lane = threadIdx.x & 31;
  const int64_t nblk512 = (ns + 511) / 512;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t blk = warp; blk < nblk512; blk += nwarps) {
    const int64_t i0 = blk * 512 + lane * 16;
    double s = 0.0;
    for (int i = 0; i < 16; ++i)
      if (i0 + i < ns) { const double x = pcm_mono(pcm, dtype, nch, i0 + i); s += x * x; }
    if (i0 < ns) e16[blk * 32 + lane] = s;
    s = warp_sum(s);
    if (lane == 0) e512[blk] = s;
  }
}

// One warp per window.  With a pyramid lane 0 walks it; without one the lanes stride the samples.
__global__ void __launch_bounds__(128) audio_levels_kernel(const void* __restrict__ pcm, int dtype, int64_t ns,
                                                           int nch, const double* __restrict__ e16,
                                                           const double* __restrict__ e512,
                                                           const int64_t* __restrict__ win_start,
                                                           const int64_t* __restrict__ win_len, int nwin,
                                                           double* __restrict__ out_db) {
  const int lane = threadIdx.x & 31;
  const int wi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wi >= nwin) return;
  // NumPy slicing clips the window to the array (hm:1071)
  int64_t s = win_start[wi], e = s + win_len[wi];
  if (s < 0) s = 0;
  if (e > ns) e = ns;
  double acc = 0.0;
  if (e16 != nullptr && e512 != nullptr) {
    if (lane == 0 && e > s) acc = window_sumsq_pyramid(pcm, dtype, nch, e16, e512, s, e);
  } else {
    for (int64_t i = s + lane; i < e; i += 32) { const double x = pcm_mono(pcm, dtype, nch, i); acc += x * x; }
    acc = warp_sum(acc);
  }
  if (lane == 0) out_db[wi] = level_db(acc, e - s);
}

}  // namespace hippo

extern "C" {

hippo_status hippo_audio_energy(const void* pcm, int32_t dtype, int64_t ns, int32_t nch, double* out_e16,
                                double* out_e512, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(ns >= 0 && nch >= 1, "hippo_audio_energy: bad sizes");
  HIPPO_REQUIRE(dtype == HIPPO_I16 || dtype == HIPPO_F32 || dtype == HIPPO_F64, "hippo_audio_energy: bad dtype %d", dtype);
  if (ns == 0) return HIPPO_OK;
  HIPPO_REQUIRE(pcm && out_e16 && out_e512, "hippo_audio_energy: null pointer");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nblk = (ns + 511) / 512;
  int64_t want = (nblk + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  const int grid = (int)(want < cap ? want : cap);
  const bool aligned = (((uintptr_t)pcm) & 15) == 0;
  if (nch == 1 && aligned) {
    if (dtype == HIPPO_F32) audio_energy_mono_kernel<float><<<grid, 256, 0, s>>>((const float*)pcm, ns, out_e16, out_e512);
    else if (dtype == HIPPO_I16) audio_energy_mono_kernel<int16_t><<<grid, 256, 0, s>>>((const int16_t*)pcm, ns, out_e16, out_e512);
    else audio_energy_mono_kernel<double><<<grid, 256, 0, s>>>((const double*)pcm, ns, out_e16, out_e512);
  } else {
    audio_energy_generic_kernel<<<grid, 256, 0, s>>>(pcm, dtype, ns, nch, out_e16, out_e512);
  }
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

hippo_status hippo_audio_levels(const void* pcm, int32_t dtype, int64_t ns, int32_t nch, const double* e16,
                                const double* e512, const int64_t* win_start, const int64_t* win_len,
                                int32_t nwin, double* out_db, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(ns >= 0 && nch >= 1 && nwin >= 0, "hippo_audio_levels: bad sizes");
  HIPPO_REQUIRE(dtype == HIPPO_I16 || dtype == HIPPO_F32 || dtype == HIPPO_F64, "hippo_audio_levels: bad dtype %d", dtype);
  if (nwin == 0) return HIPPO_OK;
  HIPPO_REQUIRE((pcm || ns == 0) && win_start && win_len && out_db, "hippo_audio_levels: null pointer");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  audio_levels_kernel<<<(nwin + 3) / 4, 128, 0, (cudaStream_t)stream>>>(pcm, dtype, ns, nch, e16, e512,
                                                                      win_start, win_len, nwin, out_db);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

}  // extern "C"
