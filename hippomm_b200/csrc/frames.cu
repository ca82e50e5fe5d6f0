// hippo_frame_pairs: visual change scoring for temporal pattern separation.
//
// Reference: _compute_frame_similarity (hm:980-991): cv2 BGR2GRAY of both frames, then
// skimage structural_similarity(gray1, gray2, data_range = gray1.max() - gray1.min());
// compute_frame_difference (bp:32-71): the same on frames scaled to [0,1] with
// data_range = 1, MSE of the scaled gray frames as the fallback.
//
// scikit-image is not part of the reference tree; its published algorithm is restated:
// 7x7 uniform window, K1 = .01, K2 = .03, sample covariance (x 49/48), border of 3 cropped,
// fp64 mean of the per-window S.  All five window moments are exact integers here (gray is
// uint8), the variance / covariance numerators 49*Sxx - Sx^2 are formed in int32 with no
// cancellation error, and only the final ratio is evaluated in fp32.
//
//   gray_minmax_kernel : BGR -> gray (cv2's fixed point: (3735 B + 19235 G + 9798 R + 16384) >> 15),
//                        per-frame min / max (the data range of hm:990 is computed in uint8)
//   ssim_pair_kernel   : one CTA per (pair, band of rows); threads own columns, march down the
//                        rows with sliding 7-row column sums, exchange them through shared memory
//                        for the 7-column horizontal sum
//   ssim_finalize_kernel: ordered sum of the band partials -> mean SSIM, MSE
#include "common.cuh"

namespace hippo {

__device__ __forceinline__ uint32_t bgr2gray(uint32_t b, uint32_t g, uint32_t r) {
  return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// grid (blocks_per_frame, nf).  3-channel fast path: a warp takes 32 groups of 16 pixels per step = 1536
// contiguous BGR bytes, loaded as three fully coalesced 512-byte rows into the warp's shared-memory slot, then
// every lane picks up its own 48 bytes (conflict-free: 12-word stride) and stores 16 gray bytes, coalesced.
__global__ void __launch_bounds__(256) gray_minmax_kernel(const uint8_t* __restrict__ frames, int64_t npix,
                                                          int ch, uint8_t* __restrict__ gray,
                                                          int2* __restrict__ minmax) {
  __shared__ uint4 s_stage[8][96];
  __shared__ uint32_t s_lo[8], s_hi[8];
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint8_t* src = frames + (int64_t)f * npix * ch;
  uint8_t* dst = gray + (int64_t)f * npix;
  uint32_t lo = 255, hi = 0;
  const bool vec = (npix % 16 == 0) && ((((uintptr_t)src) & 15) == 0) && ((((uintptr_t)dst) & 15) == 0);
  if (ch == 3 && vec) {
    const int64_t ngroups = npix / 16;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t g0 = ((int64_t)blockIdx.x * 8 + warp) * 32; g0 < ngroups; g0 += nwarps * 32) {
      const int64_t left = ngroups - g0;                       // groups this warp still has: >= 1
      const int nvec = left >= 32 ? 96 : (int)left * 3;        // 16-byte vectors to stage
      const uint4* p = reinterpret_cast<const uint4*>(src + g0 * 48);
      uint4 v[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) if (lane + 32 * u < nvec) v[u] = ldg_stream(p + lane + 32 * u);
#pragma unroll
      for (int u = 0; u < 3; ++u) if (lane + 32 * u < nvec) s_stage[warp][lane + 32 * u] = v[u];
      __syncwarp();
      if (lane < left) {
        const uint4 a = s_stage[warp][3 * lane], b = s_stage[warp][3 * lane + 1], c = s_stage[warp][3 * lane + 2];
        const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t out[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          uint32_t packed = 0;
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const int byte0 = (o * 4 + px) * 3;
            uint32_t c3[3];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
              const int bi = byte0 + cc;
              c3[cc] = (w[bi >> 2] >> ((bi & 3) * 8)) & 0xffu;
            }
            const uint32_t y = bgr2gray(c3[0], c3[1], c3[2]);
            lo = min(lo, y); hi = max(hi, y);
            packed |= y << (px * 8);
          }
          out[o] = packed;
        }
        *reinterpret_cast<uint4*>(dst + (g0 + lane) * 16) = make_uint4(out[0], out[1], out[2], out[3]);
      }
      __syncwarp();
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
         i += (int64_t)gridDim.x * blockDim.x) {
      uint32_t y;
      if (ch == 3) y = bgr2gray(src[i * 3], src[i * 3 + 1], src[i * 3 + 2]);
      else y = src[i];
      dst[i] = (uint8_t)y;
      lo = min(lo, y); hi = max(hi, y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { lo = min(lo, s_lo[i]); hi = max(hi, s_hi[i]); }
    atomicMin(&minmax[f].x, (int)lo);
    atomicMax(&minmax[f].y, (int)hi);
  }
}

__global__ void minmax_init_kernel(int2* minmax, int nf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nf) minmax[i] = make_int2(255, 0);
}

constexpr int kSsimThreads = 256;
constexpr int kSsimChunk = kSsimThreads - 6;   // output columns per column chunk

// grid (nbands, npairs).  Band k produces SSIM rows [k*bh, min((k+1)*bh, h-6)) (window top rows)
// and the squared-error sum of image rows [k*bh, ...) (last band: through h).
__global__ void __launch_bounds__(kSsimThreads) ssim_pair_kernel(
    const uint8_t* __restrict__ gray, int h, int w, const int32_t* __restrict__ pair_a,
    const int32_t* __restrict__ pair_b, const int2* __restrict__ minmax, int range_mode, int bh,
    int nbands, double* __restrict__ part_ssim, unsigned long long* __restrict__ part_sse) {
  __shared__ int4 s_cs[2][kSsimThreads];
  __shared__ double s_red[kSsimThreads / 32];
  __shared__ unsigned long long s_red2[kSsimThreads / 32];

  const int band = blockIdx.x, p = blockIdx.y;
  const int fa = pair_a ? pair_a[p] : p + 1;
  const int fb = pair_b ? pair_b[p] : p;
  const int64_t npix = (int64_t)h * w;
  const uint8_t* ga = gray + (int64_t)fa * npix;
  const uint8_t* gb = gray + (int64_t)fb * npix;

  // data range: hm:990 takes max - min of the FIRST frame in uint8; bp:61 fixes it at 1.0 on
  // the /255 scale, i.e. 255 on the integer scale
  double R = 255.0;
  if (range_mode == 0) { const int2 mm = minmax[fa]; R = (double)(mm.y - mm.x); }
  const double C1 = (0.01 * R) * (0.01 * R), C2 = (0.03 * R) * (0.03 * R);
  const float c1s = (float)(C1 * 2401.0);   // both S factors are scaled by 49^2 resp. 48*49
  const float c2s = (float)(C2 * 2352.0);

  const int out_rows = h - 6, out_cols = w - 6;
  const int y0 = band * bh;
  const int y1 = min(y0 + bh, out_rows);          // window-top rows [y0, y1)
  const int t = threadIdx.x;
  double acc = 0.0;
  unsigned long long sse = 0;
  const int sse_r1 = (band == nbands - 1) ? h : min(y0 + bh, h);

  for (int cb = 0; cb < max(out_cols, 1); cb += kSsimChunk) {
    const int x = cb + t;                         // input column of this thread
    const bool col_ok = x < w;
    // every image column is counted for the squared error by exactly one chunk (chunks overlap
    // by 6 columns): a chunk owns its first 250 columns, the last chunk owns all of its columns
    const bool sse_col = col_ok && (t < kSsimChunk || cb + kSsimChunk >= out_cols);
    int sxy_p = 0, sxx = 0, syy = 0, sxy = 0;     // sliding 7-row column sums (sx | sy << 16 packed)
    const int rows_in = (y1 > y0) ? (y1 - y0 + 6) : 0;
    const int rend = max(y0 + rows_in, sse_r1);
    for (int r = y0; r < rend; ++r) {
      int xa = 0, xb = 0;
      if (col_ok && r < h) { xa = __ldg(ga + (int64_t)r * w + x); xb = __ldg(gb + (int64_t)r * w + x); }
      if (sse_col && r < sse_r1) {
        const int dlt = xa - xb;
        sse += (unsigned long long)(dlt * dlt);
      }
      if (r >= y0 + rows_in) continue;            // rows only needed for the squared error (uniform)
      sxy_p += xa | (xb << 16);
      sxx += xa * xa; syy += xb * xb; sxy += xa * xb;
      if (r >= y0 + 7) {
        int oa = 0, ob = 0;
        if (col_ok) { oa = __ldg(ga + (int64_t)(r - 7) * w + x); ob = __ldg(gb + (int64_t)(r - 7) * w + x); }
        sxy_p -= oa | (ob << 16);
        sxx -= oa * oa; syy -= ob * ob; sxy -= oa * ob;
      }
      if (r >= y0 + 6) {
        const int buf = r & 1;
        s_cs[buf][t] = make_int4(sxy_p, sxx, syy, sxy);
        __syncthreads();
        // thread t produces the window whose left column is cb + t
        if (t < kSsimChunk && cb + t < out_cols) {
          int4 s = s_cs[buf][t];
#pragma unroll
          for (int dx = 1; dx < 7; ++dx) {
            const int4 v = s_cs[buf][t + dx];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
          }
          const int sx = s.x & 0xffff, sy = (s.x >> 16) & 0xffff;
          const int pxy = sx * sy;
          const int vx = 49 * s.y - sx * sx;        // 48*49 * var_x, exact
          const int vy = 49 * s.z - sy * sy;
          const int vxy = 49 * s.w - pxy;           // 48*49 * cov_xy, exact
          const float a1 = 2.f * (float)pxy + c1s;
          const float a2 = 2.f * (float)vxy + c2s;
          const float b1 = (float)(sx * sx + sy * sy) + c1s;
          const float b2 = (float)(vx + vy) + c2s;
          acc += (double)__fdiv_rn(a1 * a2, b1 * b2);
        }
      }
    }
    __syncthreads();
  }

  acc = warp_sum(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
  if ((t & 31) == 0) { s_red[t >> 5] = acc; s_red2[t >> 5] = sse; }
  __syncthreads();
  if (t == 0) {
    double a = 0.0; unsigned long long e = 0;
    for (int i = 0; i < kSsimThreads / 32; ++i) { a += s_red[i]; e += s_red2[i]; }
    part_ssim[(int64_t)p * nbands + band] = a;
    part_sse[(int64_t)p * nbands + band] = e;
  }
}

__global__ void ssim_finalize_kernel(const double* __restrict__ part_ssim,
                                     const unsigned long long* __restrict__ part_sse, int npairs, int nbands,
                                     int h, int w, double* __restrict__ out_ssim, double* __restrict__ out_mse) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  double a = 0.0; unsigned long long e = 0;
  for (int b = 0; b < nbands; ++b) { a += part_ssim[(int64_t)p * nbands + b]; e += part_sse[(int64_t)p * nbands + b]; }
  if (out_ssim) {
    if (h >= 7 && w >= 7) out_ssim[p] = a / ((double)(h - 6) * (double)(w - 6));
    else out_ssim[p] = __longlong_as_double(0x7ff8000000000000ll);
  }
  if (out_mse) out_mse[p] = (double)e / (65025.0 * (double)h * (double)w);
}

struct FrameLayout {
  uint8_t* gray; int2* minmax; double* part_ssim; unsigned long long* part_sse;
  int bh, nbands; size_t bytes;
};
static FrameLayout frame_layout(void* ws, size_t ws_bytes, int nf, int h, int w, int npairs) {
  Carver c(ws, ws_bytes);
  FrameLayout L{};
  L.gray = c.take<uint8_t>((size_t)nf * h * w);
  L.minmax = c.take<int2>((size_t)nf);
  const int out_rows = h >= 7 ? h - 6 : 0;
  L.bh = 56;
  L.nbands = out_rows > 0 ? (out_rows + L.bh - 1) / L.bh : 1;
  L.part_ssim = c.take<double>((size_t)npairs * L.nbands);
  L.part_sse = c.take<unsigned long long>((size_t)npairs * L.nbands);
  L.bytes = c.used();
  return L;
}

}  // namespace hippo

extern "C" {

size_t hippo_frame_pairs_workspace_bytes(int32_t nf, int32_t h, int32_t w, int32_t npairs) {
  if (nf <= 0 || h <= 0 || w <= 0 || npairs < 0) return 256;
  return hippo::frame_layout(nullptr, 0, nf, h, w, npairs).bytes;
}

hippo_status hippo_frame_pairs(const uint8_t* frames, int32_t nf, int32_t h, int32_t w, int32_t ch,
                               const int32_t* pair_a, const int32_t* pair_b, int32_t npairs,
                               int32_t range_mode, double* out_ssim, double* out_mse, void* ws,
                               size_t ws_bytes, void* stream) {
  using namespace hippo;
  HIPPO_REQUIRE(nf >= 1 && h >= 1 && w >= 1 && (ch == 1 || ch == 3), "hippo_frame_pairs: bad frame shape");
  HIPPO_REQUIRE(w <= 65535 && h <= 65535, "hippo_frame_pairs: frame too large");
  HIPPO_REQUIRE(npairs >= 0 && (range_mode == 0 || range_mode == 1), "hippo_frame_pairs: bad arguments");
  HIPPO_REQUIRE((pair_a == nullptr) == (pair_b == nullptr), "hippo_frame_pairs: pair_a/pair_b must both be given");
  HIPPO_REQUIRE(pair_a != nullptr || npairs == nf - 1, "hippo_frame_pairs: adjacent mode needs npairs == nf-1");
  if (npairs == 0) return HIPPO_OK;
  HIPPO_REQUIRE(frames != nullptr, "hippo_frame_pairs: null frames");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  cudaStream_t s = (cudaStream_t)stream;
  FrameLayout L = frame_layout(ws, ws_bytes, nf, h, w, npairs);
  if (ws == nullptr || ((uintptr_t)ws & 255) || L.bytes > ws_bytes) {
    set_error("hippo_frame_pairs: workspace of %zu bytes needed (256-byte aligned), got %zu", L.bytes, ws_bytes);
    return HIPPO_E_WORKSPACE;
  }
  const int64_t npix = (int64_t)h * w;
  minmax_init_kernel<<<(nf + 255) / 256, 256, 0, s>>>(L.minmax, nf);
  int bpf = (int)((npix / 16 + 255) / 256);
  if (bpf < 1) bpf = 1;
  if (bpf > 64) bpf = 64;
  HIPPO_REQUIRE(nf <= 65535, "hippo_frame_pairs: at most 65535 frames per call");
  gray_minmax_kernel<<<dim3(bpf, nf), 256, 0, s>>>(frames, npix, ch, L.gray, L.minmax);
  HIPPO_CUDA(cudaGetLastError());
  HIPPO_REQUIRE(npairs <= 65535, "hippo_frame_pairs: at most 65535 pairs per call");
  ssim_pair_kernel<<<dim3(L.nbands, npairs), kSsimThreads, 0, s>>>(L.gray, h, w, pair_a, pair_b, L.minmax,
                                                                   range_mode, L.bh, L.nbands, L.part_ssim,
                                                                   L.part_sse);
  HIPPO_CUDA(cudaGetLastError());
  ssim_finalize_kernel<<<(npairs + 127) / 128, 128, 0, s>>>(L.part_ssim, L.part_sse, npairs, L.nbands, h, w,
                                                           out_ssim, out_mse);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

}  // extern "C"
