// Device helpers shared by audio.cu and segment.cu.
#pragma once
#include "common.cuh"

namespace hippo {

// One mono sample in fp64.  Multi-channel input is averaged like `audio_data.mean(axis=1)`
// (hm:995-996): sequential fp64 sum divided by the channel count.  int16 PCM is k/32768,
// what soundfile hands the reference for pcm_s16le (bp:285, bp:331).
__device__ __forceinline__ double pcm_raw(const void* pcm, int dtype, int64_t idx) {
  switch (dtype) {
    case HIPPO_I16: return (double)reinterpret_cast<const int16_t*>(pcm)[idx] * (1.0 / 32768.0);
    case HIPPO_F32: return (double)reinterpret_cast<const float*>(pcm)[idx];
    default:        return reinterpret_cast<const double*>(pcm)[idx];
  }
}
__device__ __forceinline__ double pcm_mono(const void* pcm, int dtype, int nch, int64_t i) {
  if (nch == 1) return pcm_raw(pcm, dtype, i);
  double s = 0.0;
  for (int c = 0; c < nch; ++c) s += pcm_raw(pcm, dtype, i * nch + c);
  return s / (double)nch;
}

// Sum of squares of samples [s, e) by ONE thread from the energy pyramid: head samples up
// to a 16-boundary, 16-blocks up to a 512-boundary, 512-blocks, then back down.  No
// prefix differences, so nothing cancels; exact for int16-origin PCM.
__device__ __forceinline__ double window_sumsq_pyramid(const void* pcm, int dtype, int nch,
                                                       const double* __restrict__ e16,
                                                       const double* __restrict__ e512, int64_t s, int64_t e) {
  double acc = 0.0;
  int64_t i = s;
  while (i < e && (i & 15)) { const double x = pcm_mono(pcm, dtype, nch, i); acc += x * x; ++i; }
  while (i + 16 <= e && (i & 511)) { acc += e16[i >> 4]; i += 16; }
  while (i + 512 <= e) { acc += e512[i >> 9]; i += 512; }
  while (i + 16 <= e) { acc += e16[i >> 4]; i += 16; }
  while (i < e) { const double x = pcm_mono(pcm, dtype, nch, i); acc += x * x; ++i; }
  return acc;
}

// hm:998-999: rms = sqrt(mean(x^2)); db = 20*log10(rms) if rms > 0 else -100.
// An empty window gives mean = NaN in NumPy, `NaN > 0` is False, hence -100 as well.
__device__ __forceinline__ double level_db(double sumsq, int64_t len) {
  if (len <= 0) return -100.0;
  const double rms = sqrt(sumsq / (double)len);
  return rms > 0.0 ? 20.0 * log10(rms) : -100.0;
}

}  // namespace hippo
