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
//   gray_minmax_*kernel : BGR -> gray (cv2's fixed point: (3735 B + 19235 G + 9798 R + 16384) >> 15), per-frame
//                         min / max (the data range of hm:990 is computed in uint8).  Two row layouts, one per lane
//                         mapping of the SSIM kernel: rows padded with zeros to a multiple of 4 bytes, or groups of 7
//                         pixels in 8 bytes (byte 7 = 0); a vectorised kernel for each (persistent warps, the next
//                         item's loads in flight under the conversion) and a byte-wise one for odd sizes
//   ssim_pair*_kernel   : one WARP per (pair, band of rows, chunk of columns), no block-level synchronisation.  A
//                         lane owns 4 (or 7) adjacent columns, marches down the rows with sliding 7-row column sums
//                         (x^2 + y^2 and 2xy by dp2a on permuted bytes) and gets the columns to its right from its
//                         neighbours by shuffles.  Issue-bound (integer ALU), not HBM-bound: see DESIGN.md 4.4.
//   ssim_finalize_kernel: ordered sum of the warp partials -> mean SSIM, MSE
#include "common.cuh"
#include <cstdlib>
#include <type_traits>

namespace hippo {

constexpr uint32_t kGrayW_BG = 3735u | (19235u << 16);   // 16-bit weights of bytes 0, 1 (B, G)
constexpr uint32_t kGrayW_R = 9798u;                      // 16-bit weights of bytes 2, 3 (R, unused)
__device__ __forceinline__ uint32_t bgr2gray(uint32_t b, uint32_t g, uint32_t r) {
  return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// 3-channel fast path (frames whose pixel count is a multiple of 16, no row padding): persistent warps loop over
// (frame, 512-pixel slice) items.  A warp takes 32 groups of 16 pixels per item = 1536 contiguous BGR bytes,
// loaded as three fully coalesced 512-byte rows into the warp's shared-memory slot, then every lane picks up
// its own 48 bytes (conflict-free: 12-word stride) and stores 16 gray bytes, coalesced.
// One lane's 16 pixels of a staged warp item (48 BGR bytes at s[3 * lane ..]) -> 16 gray bytes, running min / max.
__device__ __forceinline__ uint4 gray16(const uint4* s, int lane, uint32_t& lo, uint32_t& hi) {
  const uint4 a = s[3 * lane], b = s[3 * lane + 1], c = s[3 * lane + 2];
  const uint32_t w[13] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, 0u};
  uint32_t out[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    uint32_t packed = 0;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      // the pixel's B, G, R bytes into one register (one PRMT across two words), then two 16 x 8-bit dot
      // products: 3735 B + 19235 G + 16384, + 9798 R
      const int byte0 = (o * 4 + px) * 3;
      const uint32_t bgr = __byte_perm(w[byte0 >> 2], w[(byte0 >> 2) + 1], 0x3210 + 0x1111 * (byte0 & 3));
      const uint32_t y = (uint32_t)__dp2a_hi(kGrayW_R, bgr, __dp2a_lo(kGrayW_BG, bgr, 16384u)) >> 15;
      lo = min(lo, y); hi = max(hi, y);
      packed |= y << (px * 8);
    }
    out[o] = packed;
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

// `gstride`: bytes from one gray frame to the next (a multiple of 128: no cache line holds bytes of two frames).
__global__ void __launch_bounds__(256) gray_minmax_vec_kernel(const uint8_t* __restrict__ frames, int64_t npix,
                                                              int nf, uint8_t* __restrict__ gray, int64_t gstride,
                                                              int2* __restrict__ minmax) {
  __shared__ uint4 s_stage[8][96];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ngroups = npix / 16;
  const int64_t wpf = (ngroups + 31) / 32;                   // warp items per frame
  const int64_t nitems = wpf * nf;
  // Every warp takes a CONTIGUOUS run of items (no division per item, and the frame's min / max is reduced across the
  // warp once per frame instead of once per item); the loads of its NEXT item are in flight while it converts the
  // current one (one item at a time left the warps waiting on HBM for half of their cycles).
  const int64_t nwarps = (int64_t)gridDim.x * 8, wid = (int64_t)blockIdx.x * 8 + warp;
  const int64_t per = (nitems + nwarps - 1) / nwarps;
  int64_t item = wid * per;
  const int64_t item_end = item + per < nitems ? item + per : nitems;
  if (item >= item_end) return;
  int64_t f = item / wpf;
  int64_t j = item - f * wpf;                                // item of the frame
  uint4 v[3];
  auto issue = [&](int64_t fi, int64_t ji) {
    const int64_t g0 = ji * 32, left = ngroups - g0;
    const int nvec = left >= 32 ? 96 : (int)left * 3;        // 16-byte vectors to stage
    const uint4* p = reinterpret_cast<const uint4*>(frames + fi * npix * 3 + g0 * 48);
#pragma unroll
    for (int u = 0; u < 3; ++u) if (lane + 32 * u < nvec) v[u] = ldg_stream(p + lane + 32 * u);
  };
  issue(f, j);
  uint32_t lo = 255, hi = 0;
  for (; item < item_end; ++item) {
    const int64_t g0 = j * 32;
    uint8_t* dst = gray + f * gstride;
    const int64_t left = ngroups - g0;                       // groups this warp still has: >= 1
    const int nvec = left >= 32 ? 96 : (int)left * 3;
#pragma unroll
    for (int u = 0; u < 3; ++u) if (lane + 32 * u < nvec) s_stage[warp][lane + 32 * u] = v[u];
    __syncwarp();
    const bool frame_ends = j + 1 == wpf;
    const int64_t fn = frame_ends ? f + 1 : f, jn = frame_ends ? 0 : j + 1;
    if (item + 1 < item_end) issue(fn, jn);
    if (lane < left) *reinterpret_cast<uint4*>(dst + (g0 + lane) * 16) = gray16(s_stage[warp], lane, lo, hi);
    __syncwarp();
    if (frame_ends || item + 1 == item_end) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      if (lane == 0) {
        atomicMin(&minmax[f].x, (int)lo);
        atomicMax(&minmax[f].y, (int)hi);
      }
      lo = 255; hi = 0;
    }
    f = fn; j = jn;
  }
}

// General path (1 channel, odd sizes, padded rows): grid (blocks_per_frame, nf), one byte per thread and step.
// cpl 7: a row is stored as groups of 7 pixels in 8 bytes (byte 7 of a group and the pixels beyond the row are 0).
__global__ void __launch_bounds__(256) gray_minmax_kernel(const uint8_t* __restrict__ frames, int64_t npix,
                                                          int w, int pitch, int ch, int cpl, uint8_t* __restrict__ gray,
                                                          int64_t gstride, int2* __restrict__ minmax) {
  __shared__ uint32_t s_lo[8], s_hi[8];
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint8_t* src = frames + (int64_t)f * npix * ch;
  const int64_t gpix = npix / w * pitch;      // bytes of one padded gray frame
  uint8_t* dst = gray + (int64_t)f * gstride;
  uint32_t lo = 255, hi = 0;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < gpix;
       o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = o / pitch;
    const int xb = (int)(o - r * pitch);
    const int x = cpl == 7 ? (xb >> 3) * 7 + (xb & 7) : xb;
    if (x >= w || (cpl == 7 && (xb & 7) == 7)) { dst[o] = 0; continue; }   // padding: zero in both frames of a pair
    const int64_t i = r * w + x;
    uint32_t y;
    if (ch == 3) y = bgr2gray(src[i * 3], src[i * 3 + 1], src[i * 3 + 2]);
    else y = src[i];
    dst[o] = (uint8_t)y;
    lo = min(lo, y); hi = max(hi, y);
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

// 3-channel fast path of the 7-in-8 layout (frame width a multiple of 7, whole warp items): a warp takes 32 runs of 14
// pixels = 1,344 contiguous BGR bytes, staged like above; a lane's 42 bytes start on a 2-byte boundary for odd lanes,
// so it reads 12 words and realigns them with funnel shifts, then stores two groups (16 bytes) at once.
__global__ void __launch_bounds__(256) gray_minmax_vec7_kernel(const uint8_t* __restrict__ frames, int64_t npix,
                                                               int nf, uint8_t* __restrict__ gray, int64_t gstride,
                                                               int2* __restrict__ minmax) {
  __shared__ uint4 s_stage[8][85];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wpf = npix / (14 * 32);                      // warp items per frame
  const int64_t nitems = wpf * nf;
  // contiguous run of items per warp, next item's loads in flight under the conversion, min / max reduced once per
  // frame (see gray_minmax_vec_kernel)
  const int64_t nwarps = (int64_t)gridDim.x * 8, wid = (int64_t)blockIdx.x * 8 + warp;
  const int64_t per = (nitems + nwarps - 1) / nwarps;
  int64_t item = wid * per;
  const int64_t item_end = item + per < nitems ? item + per : nitems;
  if (item >= item_end) return;
  int64_t f = item / wpf;
  int64_t j = item - f * wpf;
  uint4 v[3];
  auto issue = [&](int64_t fi, int64_t ji) {
    const uint4* p = reinterpret_cast<const uint4*>(frames + fi * npix * 3 + ji * (32 * 42));
#pragma unroll
    for (int u = 0; u < 3; ++u) if (lane + 32 * u < 84) v[u] = ldg_stream(p + lane + 32 * u);
  };
  issue(f, j);
  uint32_t lo = 255, hi = 0;
  for (; item < item_end; ++item) {
    const int64_t g0 = j * 32;                                               // first 14-pixel run of the item
#pragma unroll
    for (int u = 0; u < 3; ++u) if (lane + 32 * u < 84) s_stage[warp][lane + 32 * u] = v[u];
    __syncwarp();
    const bool frame_ends = j + 1 == wpf;
    const int64_t fn = frame_ends ? f + 1 : f, jn = frame_ends ? 0 : j + 1;
    if (item + 1 < item_end) issue(fn, jn);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s_stage[warp]) + (42 * lane >> 2);
    const uint32_t sh = (lane & 1) * 16;
    uint32_t raw[12], a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) raw[i] = sw[i];             // the last word of lane 31 is the slot's spare vector
#pragma unroll
    for (int i = 0; i < 11; ++i) a[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
    a[11] = 0u;
    uint32_t out[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      uint32_t packed = 0;
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        if ((o & 1) && px == 3) continue;                    // byte 7 of a group stays 0
        const int q = (o >> 1) * 7 + (o & 1) * 4 + px;       // pixel of the run
        const int byte0 = q * 3;
        const uint32_t bgr = __byte_perm(a[byte0 >> 2], a[(byte0 >> 2) + 1], 0x3210 + 0x1111 * (byte0 & 3));
        const uint32_t y = (uint32_t)__dp2a_hi(kGrayW_R, bgr, __dp2a_lo(kGrayW_BG, bgr, 16384u)) >> 15;
        lo = min(lo, y); hi = max(hi, y);
        packed |= y << (px * 8);
      }
      out[o] = packed;
    }
    *reinterpret_cast<uint4*>(gray + f * gstride + (g0 + lane) * 16) = make_uint4(out[0], out[1], out[2], out[3]);
    __syncwarp();
    if (frame_ends || item + 1 == item_end) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      if (lane == 0) {
        atomicMin(&minmax[f].x, (int)lo);
        atomicMax(&minmax[f].y, (int)hi);
      }
      lo = 255; hi = 0;
    }
    f = fn; j = jn;
  }
}

__global__ void minmax_init_kernel(int2* minmax, int nf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nf) minmax[i] = make_int2(255, 0);
}

constexpr int kSsimThreads = 256;
constexpr int kSsimWarps = kSsimThreads / 32;
constexpr int kSsimLiveSmem = 8 * 1024;        // see frames_adjacent_live_launch
constexpr int kSsimBand = 56;                  // window rows per band (6 halo rows per band: 11%; 112-row bands measured
                                               // slower on one stream-hour: fewer, longer warp items, longer tail)

// Packed fp32 pairs (sm_100 f32x2 arithmetic): two IEEE round-to-nearest results per instruction.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// One warp per (pair, band, chunk).  Band k produces SSIM window-top rows [k*bh, min((k+1)*bh, h-6)) and the
// squared error of image rows [k*bh, ...) (last band: through h).  Two gray layouts / lane mappings (`cpl`, columns
// per lane; frame_layout picks the one that wastes fewer lane slots for the frame width):
//   cpl 4   rows as they are (pitch = w rounded up to 4): a lane owns one 32-bit word per row, a chunk is 30 lanes =
//           120 window columns, lanes 30 / 31 only feed their column sums to the windows of lanes 28 / 29;
//   cpl 7   rows stored as groups of 7 pixels in 8 bytes (byte 7 of a group is 0): a lane owns one group = one 8-byte
//           load per row, a window spans its own and the next lane's columns only, so a chunk is 31 lanes = 217 window
//           columns -- and the last lane's first window lies in its own group, which makes a 224-pixel row (218
//           windows, 32 groups) exactly ONE warp: 97% of the lane slots carry a window, against 85% with two
//           120-column chunks.
struct SsimArgs {
  const uint8_t* gray; int64_t gstride; int h, w, pitch;
  const int32_t* pair_a; const int32_t* pair_b;
  const int2* minmax; int range_mode, bh, nbands, nchunks;
  double* part_ssim; unsigned long long* part_sse;   // partials of the launch's items, [nitems]
  // live finalisation (pattern.cu, null otherwise): the warp that delivers the LAST partial of a pair adds the pair's
  // partials in the fixed order of ssim_finalize_kernel and stores the pair's SSIM / MSE -- one 8-byte store that a
  // boundary chain running beside this kernel polls (segment.cu, follow mode); pair_done[p] counts the partials of
  // pair p (zeroed by the caller), out_* are indexed by the pair's number in the stream
  unsigned int* pair_done; double* out_ssim; double* out_mse;
};

constexpr int ssim_chunk_cols(int cpl) { return cpl == 7 ? 217 : 120; }
// chunks per row: cpl 7 lets the last chunk own one window more (the last lane's first)
__host__ __device__ inline int ssim_nchunks(int out_cols, int cpl) {
  if (out_cols <= 0) return 1;
  const int n = cpl == 7 ? (out_cols - 1 + 216) / 217 : (out_cols + 119) / 120;
  return n < 1 ? 1 : n;
}

// One (pair, band, chunk) item, by one warp; `item` numbers the items of ALL pairs (pair-major), `slot` is where its
// partial sums go.
//
// Instruction budget of a row step (the kernel is issue / latency bound): the row words come through
// four running pointers (no per-row index arithmetic, no bounds test: lanes beyond the row read the row's first word, the
// last step reads one row past the band -- the workspace carries a spare row -- and nothing such a value feeds is owned);
// per column ONE byte permute yields (x, y, y, x) and a second its zero-extended halves x | y << 16, so that
// x^2 + y^2 and 2 x y are one dp2a each and sum x | sum y << 16 advances by the difference of two permutes; the squared
// error is accumulated by every lane and dropped at the end by the lanes that do not own their word.
template <int CPL>
__device__ __forceinline__ void ssim_item(const SsimArgs& A, int64_t item, int64_t slot, int lane) {
  static_assert(CPL == 4 || CPL == 7, "columns per lane");
  constexpr int W = CPL == 7 ? 2 : 1;                        // 32-bit words a lane loads per row and frame
  constexpr int kOwnLanes = CPL == 7 ? 31 : 30;              // lanes whose windows belong to the chunk
  const uint8_t* __restrict__ gray = A.gray;
  const int h = A.h, w = A.w, pitch = A.pitch, bh = A.bh, nbands = A.nbands, nchunks = A.nchunks;
  const int32_t* __restrict__ pair_a = A.pair_a;
  const int32_t* __restrict__ pair_b = A.pair_b;
  const int2* __restrict__ minmax = A.minmax;
  const int range_mode = A.range_mode;
  const int chunk = (int)(item % nchunks);
  const int band = (int)((item / nchunks) % nbands);
  const int p = (int)(item / ((int64_t)nchunks * nbands));
  const int fa = pair_a ? pair_a[p] : p + 1;
  const int fb = pair_b ? pair_b[p] : p;
  const int64_t gpix = A.gstride;
  const int wordx = (chunk * kOwnLanes + lane) * 4 * W;      // byte offset of this lane's word(s) in a row
  const bool col_ok = wordx < pitch;
  const int wordx_ld = col_ok ? wordx : 0;                   // what a lane beyond the row reads instead (never owned)
  const uint8_t* ga = gray + (int64_t)fa * gpix + wordx_ld;
  const uint8_t* gb = gray + (int64_t)fb * gpix + wordx_ld;

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
  const int rows_in = (y1 > y0) ? (y1 - y0 + 6) : 0;
  const int sse_r1 = (band == nbands - 1) ? h : min(y0 + bh, h);
  // every image word is counted for the squared error by exactly one warp (chunks overlap by the feeding lanes); which
  // of this lane's windows exist and belong to this chunk
  const bool last_chunk = chunk == nchunks - 1;
  const bool sse_own = col_ok && (lane < kOwnLanes || last_chunk);
  const int col0 = (chunk * kOwnLanes + lane) * CPL;          // image column of this lane's first pixel
  bool own[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    if constexpr (CPL == 7) own[k] = (lane < kOwnLanes || (k == 0 && last_chunk)) && col0 + k < out_cols;
    else own[k] = lane < kOwnLanes && col0 + k < out_cols;
  }

  // sliding 7-row column sums: sum x | sum y << 16, sum (x^2 + y^2) (the two variances only ever appear added), 2 sum x y
  uint32_t sp[CPL], sq[CPL], sxy[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) sp[k] = sq[k] = sxy[k] = 0;
  double acc = 0.0;
  unsigned long long sse = 0;
  uint32_t sqe = 0, cr = 0;                       // sum a^2 + b^2, sum a b of the rows owned for the squared error
  const uint64_t c1s2 = pack_f32x2(c1s, c1s), c2s2 = pack_f32x2(c2s, c2s);

  // One row step.  kFull: the row completes a 7-row window (horizontal sums + ratio) and the row that leaves the
  // window is fetched for the next step; kSse: this warp owns the row for the squared error.  The phases below
  // call it with compile-time flags, so the steady state carries no per-row predicates.
  static_assert(kSsimBand + 6 <= 2048, "sqe / cr hold rows x 8 bytes x 2 x 255^2 < 2^32 without a flush: at most 4,128 rows");
  using Row = std::conditional_t<W == 2, uint2, uint32_t>;
  const int pitchw = pitch / (4 * W);                        // row stride in loads (the pitch is a multiple of 4 W)
  const Row* ga4 = reinterpret_cast<const Row*>(ga);
  const Row* gb4 = reinterpret_cast<const Row*>(gb);
  const Row* pna = ga4 + (int64_t)(y0 + 1) * pitchw;         // the next row of either frame
  const Row* pnb = gb4 + (int64_t)(y0 + 1) * pitchw;
  const Row* poa = ga4 + (int64_t)y0 * pitchw;               // the next row to leave the window
  const Row* pob = gb4 + (int64_t)y0 * pitchw;
  // four independent running pointers, one 64-bit multiply-add each per row (left to itself the compiler folds them
  // into two running offsets and re-adds the bases for every load)
  asm("" : "+l"(pna)); asm("" : "+l"(pnb)); asm("" : "+l"(poa)); asm("" : "+l"(pob));
  auto ld = [](const Row* q, uint32_t (&o)[W]) {
    const Row v = __ldg(q);
    if constexpr (W == 2) { o[0] = v.x; o[1] = v.y; } else { o[0] = v; }
  };
  uint32_t wa[W], wb[W], oa[W], ob[W];
#pragma unroll
  for (int i = 0; i < W; ++i) wa[i] = wb[i] = oa[i] = ob[i] = 0;
  auto row_step = [&](auto full_tag, auto sse_tag) {
    constexpr bool kFull = decltype(full_tag)::value, kSse = decltype(sse_tag)::value;
    // next row's words (and the row leaving the window): issued now, used in the next step
    uint32_t nwa[W], nwb[W], noa[W], nob[W];
    ld(pna, nwa); ld(pnb, nwb);
    pna += pitchw; pnb += pitchw;
#pragma unroll
    for (int i = 0; i < W; ++i) noa[i] = nob[i] = 0;
    if constexpr (kFull) {
      ld(poa, noa); ld(pob, nob);
      poa += pitchw; pob += pitchw;
    }
    if constexpr (kSse) {
#pragma unroll
      for (int i = 0; i < W; ++i) { sqe = __dp4a(wa[i], wa[i], sqe); sqe = __dp4a(wb[i], wb[i], sqe); cr = __dp4a(wa[i], wb[i], cr); }
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int kk = k & 3, wi = k >> 2;
      const uint32_t sel = (uint32_t)(kk | ((4 + kk) << 4) | ((4 + kk) << 8) | (kk << 12));
      const uint32_t bn = __byte_perm(wa[wi], wb[wi], sel), hn = __byte_perm(bn, 0u, 0x4140);   // bytes (x, y, y, x); halves x | y << 16
      sq[k] = __dp2a_lo(hn, bn, sq[k]);            // + x^2 + y^2
      sxy[k] = __dp2a_hi(hn, bn, sxy[k]);          // + 2 x y
      if constexpr (kFull) {
        const uint32_t bo = __byte_perm(oa[wi], ob[wi], sel), ho = __byte_perm(bo, 0u, 0x4140);
        sq[k] -= __dp2a_lo(ho, bo, 0u);
        sxy[k] -= __dp2a_hi(ho, bo, 0u);
        sp[k] += hn - ho;                          // the running sums never go negative
      } else {
        sp[k] += hn;
      }
    }
    if constexpr (kFull) {
      // o[k] = sum of columns k .. k+6 of a quantity: a sliding chain, one 3-input add per window
      int o_sp[CPL], o_sq[CPL], o_xy[CPL];
      auto horiz = [&](const uint32_t (&cu)[CPL], int (&o)[CPL]) {
        if constexpr (CPL == 4) {
          // the rest of this lane's columns, then the neighbour's prefix, then one or two columns of the lane after it
          const int c0 = (int)cu[0], c1 = (int)cu[1], c2 = (int)cu[2], c3 = (int)cu[3];
          const int C = c0 + c1 + c2, D = C + c3;
          const int C1n = __shfl_down_sync(0xffffffffu, C, 1), n3 = __shfl_down_sync(0xffffffffu, c3, 1);
          const int m0 = __shfl_down_sync(0xffffffffu, c0, 2), m1 = __shfl_down_sync(0xffffffffu, c1, 2);
          o[0] = D + C1n;
          o[1] = o[0] + n3 - c0;
          o[2] = o[1] + m0 - c1;
          o[3] = o[2] + m1 - c2;
        } else {
          // window 0 is this lane's own group; window k trades columns 0 .. k-1 for the same columns of the next lane
          int c[7];
#pragma unroll
          for (int k = 0; k < 7; ++k) c[k] = (int)cu[k];
          o[0] = ((c[0] + c[1] + c[2]) + (c[3] + c[4] + c[5])) + c[6];
#pragma unroll
          for (int k = 1; k < 7; ++k) o[k] = o[k - 1] + __shfl_down_sync(0xffffffffu, c[k - 1], 1) - c[k - 1];
        }
      };
      horiz(sp, o_sp); horiz(sq, o_sq); horiz(sxy, o_xy);
      float fu[CPL + 1], fvs[CPL + 1], fpxy[CPL + 1], fvxy[CPL + 1];
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int sx = o_sp[k] & 0xffff, sy = (int)((uint32_t)o_sp[k] >> 16);
        const int pxy2 = sx * (sy + sy);                    // 2 * 49^2 mu_x mu_y
        const int u = sy * sy + sx * sx;                    // 49^2 (mu_x^2 + mu_y^2)
        fu[k] = (float)u;
        fvs[k] = (float)(49 * o_sq[k] - u);                 // 48*49 (var_x + var_y), exact
        fpxy[k] = (float)pxy2;
        fvxy[k] = (float)(49 * o_xy[k] - pxy2);             // 2 * 48*49 cov_xy, exact
      }
      fu[CPL] = fu[CPL - 1]; fvs[CPL] = fvs[CPL - 1]; fpxy[CPL] = fpxy[CPL - 1]; fvxy[CPL] = fvxy[CPL - 1];   // odd CPL: pad the last pair
      // the fp32 ratio, two windows per instruction (f32x2: same IEEE results as the scalar forms; the doubled
      // integers convert to exactly twice the fp32 values, so `x2 + c` rounds like fma(2, x, c))
      float s4 = 0.f;
#pragma unroll
      for (int k = 0; k < CPL; k += 2) {
        const uint64_t a1 = add_f32x2(pack_f32x2(fpxy[k], fpxy[k + 1]), c1s2);
        const uint64_t a2 = add_f32x2(pack_f32x2(fvxy[k], fvxy[k + 1]), c2s2);
        const uint64_t b1 = add_f32x2(pack_f32x2(fu[k], fu[k + 1]), c1s2);
        const uint64_t b2 = add_f32x2(pack_f32x2(fvs[k], fvs[k + 1]), c2s2);
        float d0, d1, r0, r1;
        unpack_f32x2(mul_f32x2(b1, b2), d0, d1);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));   // 1 ulp; 0 * inf (constant frames, R = 0) stays NaN like 0/0
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
        float v0, v1;
        unpack_f32x2(mul_f32x2(mul_f32x2(a1, a2), pack_f32x2(r0, r1)), v0, v1);
        if (k == 0) s4 = own[0] ? v0 : 0.f; else s4 += own[k] ? v0 : 0.f;
        if (k + 1 < CPL) s4 += own[k + 1 < CPL ? k + 1 : k] ? v1 : 0.f;
      }
      acc += (double)s4;
    }
#pragma unroll
    for (int i = 0; i < W; ++i) { wa[i] = nwa[i]; wb[i] = nwb[i]; oa[i] = noa[i]; ob[i] = nob[i]; }
  };
  using T = std::true_type;
  using F = std::false_type;
  if (rows_in > 0) {
    // rows [y0, y0+6) fill the window; [y0+6, sse_end) are complete rows this warp also owns for the squared
    // error; [sse_end, y0+rows_in) are the six rows that belong to the next band's squared error
    const int win_end = y0 + rows_in, sse_end = min(sse_r1, win_end);
    ld(ga4 + (int64_t)y0 * pitchw, wa);
    ld(gb4 + (int64_t)y0 * pitchw, wb);
    int r = y0;
#pragma unroll 1
    for (; r < y0 + 6; ++r) row_step(F{}, T{});
#pragma unroll 1
    for (; r < sse_end; ++r) row_step(T{}, T{});
#pragma unroll 1
    for (; r < win_end; ++r) row_step(T{}, F{});
  } else {
    // a band without windows (frames lower than 7 rows): squared error only
    for (int r = y0; r < sse_r1; ++r) {
      uint32_t xa[W], xb[W];
      ld(ga4 + (int64_t)r * pitchw, xa);
      ld(gb4 + (int64_t)r * pitchw, xb);
#pragma unroll
      for (int i = 0; i < W; ++i) { sqe = __dp4a(xa[i], xa[i], sqe); sqe = __dp4a(xb[i], xb[i], sqe); cr = __dp4a(xa[i], xb[i], cr); }
    }
  }
  if (sse_own) sse += (unsigned long long)sqe - 2ull * cr;

  acc = warp_sum(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
  if (lane == 0) {
    A.part_ssim[slot] = acc;
    A.part_sse[slot] = sse;
    if (A.pair_done != nullptr) {
      const int nparts = nbands * nchunks;
      __threadfence();                                    // the partial is visible before the count
      if (atomicAdd(&A.pair_done[p], 1u) == (unsigned)(nparts - 1)) {
        __threadfence();
        const int64_t slot0 = slot - (item - (int64_t)p * nparts);
        double a = 0.0; unsigned long long e = 0;
        for (int b = 0; b < nparts; ++b) { a += __ldcg(A.part_ssim + slot0 + b); e += __ldcg(A.part_sse + slot0 + b); }
        if (A.out_mse) A.out_mse[p] = (double)e / (65025.0 * (double)h * (double)w);
        double v = __longlong_as_double(0x7ff8000000000000ll);
        if (h >= 7 && w >= 7) v = a / ((double)(h - 6) * (double)(w - 6));
        if (A.out_ssim) __stcg(A.out_ssim + p, v);
      }
    }
  }
}

// One warp per item: warp i of the grid takes item item0 + i (explicit pair lists: pre-filter chain, QA de-dup).
// cpl 4: four CTAs of 256 threads per SM (64 registers, no spill); cpl 7: 21 column sums per lane, five CTAs of 128
// threads at 96 registers (six at 80 registers spill and measured 2% slower; 28-row bands: SSIM 1% slower, a batch of
// streams 2.5% slower, a single stream 3% faster -- its boundary chain trails a shorter last wave of pairs).
__global__ void __launch_bounds__(kSsimThreads, 4) ssim_pair_kernel(const SsimArgs A, int64_t nitems, int64_t item0) {
  const int64_t i = (int64_t)blockIdx.x * kSsimWarps + (threadIdx.x >> 5);
  if (i >= nitems) return;
  ssim_item<4>(A, item0 + i, i, threadIdx.x & 31);
}
constexpr int kSsim7Threads = 128;
constexpr int kSsim7Warps = kSsim7Threads / 32;
__global__ void __launch_bounds__(kSsim7Threads, 5) ssim_pair7_kernel(const SsimArgs A, int64_t nitems, int64_t item0) {
  const int64_t i = (int64_t)blockIdx.x * kSsim7Warps + (threadIdx.x >> 5);
  if (i >= nitems) return;
  ssim_item<7>(A, item0 + i, i, threadIdx.x & 31);
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
  int cpl, pitch, bh, nbands, nchunks, nparts; int64_t gstride; size_t bytes;
};
// columns per lane of the SSIM kernel = layout of the gray rows: the mapping that issues fewer row steps x instructions
// for this width (173 instructions per row step of a 120-column chunk, 290 per 217-column chunk) -- unless that would
// trade the vectorised gray conversion for the byte-wise one (7-in-8 rows are only written by the fast kernel when the
// width is a multiple of 7 and the frame is whole warp items).  HIPPO_SSIM_CPL overrides (4 / 7, A/B and tests).
static int ssim_pick_cpl(int h, int w) {
  static const int forced = getenv("HIPPO_SSIM_CPL") ? atoi(getenv("HIPPO_SSIM_CPL")) : 0;
  if (forced == 4 || forced == 7) return forced;
  if (w < 7) return 4;
  const int out_cols = w - 6;
  if (ssim_nchunks(out_cols, 7) * 290 >= ssim_nchunks(out_cols, 4) * 173) return 4;
  const int64_t npix = (int64_t)h * w;
  const bool vec7 = w % 7 == 0 && npix % (14 * 32) == 0 && npix % 16 == 0;
  const bool vec4 = w % 4 == 0 && npix % 16 == 0;
  return (vec7 || !vec4) ? 7 : 4;
}
static FrameLayout frame_layout(void* ws, size_t ws_bytes, int nf, int h, int w, int npairs) {
  Carver c(ws, ws_bytes);
  FrameLayout L{};
  L.cpl = ssim_pick_cpl(h, w);
  L.pitch = L.cpl == 7 ? (w + 6) / 7 * 8 : (w + 3) & ~3;
  L.gstride = (int64_t)align_up((size_t)h * L.pitch, 128);
  // + one spare row: the SSIM warps prefetch the row below their band without a bounds test (the value is never used)
  L.gray = c.take<uint8_t>((size_t)nf * L.gstride + align_up((size_t)L.pitch, 128));
  L.minmax = c.take<int2>((size_t)nf);
  const int out_rows = h >= 7 ? h - 6 : 0, out_cols = w >= 7 ? w - 6 : 0;
  L.bh = kSsimBand;
  L.nbands = out_rows > 0 ? (out_rows + L.bh - 1) / L.bh : 1;
  L.nchunks = ssim_nchunks(out_cols, L.cpl);
  L.nparts = L.nbands * L.nchunks;
  L.part_ssim = c.take<double>((size_t)npairs * L.nparts);
  L.part_sse = c.take<unsigned long long>((size_t)npairs * L.nparts);
  L.bytes = c.used();
  return L;
}

// gray conversion of all frames (+ min / max), stream s
// `reserved_sms`: SMs another resident kernel keeps to itself (the boundary chain in follow mode)
static void gray_launch(const uint8_t* frames, int nf, int h, int w, int ch, const FrameLayout& L, cudaStream_t s,
                        int reserved_sms = 0) {
  const int64_t npix = (int64_t)h * w;
  minmax_init_kernel<<<(nf + 255) / 256, 256, 0, s>>>(L.minmax, nf);
  int bpf = (int)((npix / 16 + 255) / 256);
  if (bpf < 1) bpf = 1;
  if (bpf > 64) bpf = 64;
  const bool aligned = ch == 3 && npix % 16 == 0 && (((uintptr_t)frames) & 15) == 0;   // L.gray is 256-byte aligned
  // persistent grids: exactly the CTAs that are resident at once (a second, thin wave would run on a mostly idle GPU)
  auto resident = [reserved_sms](const void* fn, int& cached) -> int64_t {
    if (cached == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached, fn, 256, 0) != cudaSuccess || cached < 1)) cached = 4;
    const int sms = sm_count() - reserved_sms;
    return (int64_t)(sms > 1 ? sms : 1) * cached;
  };
  static int occ4 = 0, occ7 = 0;
  if (aligned && L.cpl == 4 && L.pitch == w) {
    const int64_t items = (npix / 16 + 31) / 32 * nf;
    const int64_t want = (items + 7) / 8, cap = resident((const void*)gray_minmax_vec_kernel, occ4);
    gray_minmax_vec_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(frames, npix, nf, L.gray, L.gstride, L.minmax);
  } else if (aligned && L.cpl == 7 && w % 7 == 0 && npix % (14 * 32) == 0) {
    // rows are whole groups, so the frame is one run of groups: 14 pixels (two groups, 16 gray bytes) per lane
    const int64_t items = npix / (14 * 32) * nf;
    const int64_t want = (items + 7) / 8, cap = resident((const void*)gray_minmax_vec7_kernel, occ7);
    gray_minmax_vec7_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(frames, npix, nf, L.gray, L.gstride, L.minmax);
  } else {
    gray_minmax_kernel<<<dim3(bpf, nf), 256, 0, s>>>(frames, npix, w, L.pitch, ch, L.cpl, L.gray, L.gstride, L.minmax);
  }
}

// the SSIM launch of `nitems` (pair, band, chunk) items for layout L
static void ssim_launch(const FrameLayout& L, const SsimArgs& A, int64_t nitems, size_t smem, cudaStream_t s) {
  if (L.cpl == 7)
    ssim_pair7_kernel<<<(unsigned)((nitems + kSsim7Warps - 1) / kSsim7Warps), kSsim7Threads, smem, s>>>(A, nitems, 0);
  else
    ssim_pair_kernel<<<(unsigned)((nitems + kSsimWarps - 1) / kSsimWarps), kSsimThreads, smem, s>>>(A, nitems, 0);
}

// Adjacent pairs [0, nf - 1) of a stream with LIVE finalisation (pattern.cu): gray conversion, then one SSIM launch
// whose warps store every pair's result as they complete it (`pair_done`: nf - 1 zeroed counters, see SsimArgs) -- no
// finalize launch.  The SSIM CTAs ask for kSsimLiveSmem bytes of (untouched) dynamic shared memory, the gray CTAs hold
// 12 KB of static shared memory: neither fits on the SM the boundary chain reserved for itself (segment.cu, follow mode).
hippo_status frames_adjacent_live_launch(const uint8_t* frames, int nf, int h, int w, int ch, void* ws, size_t ws_bytes,
                                         unsigned int* pair_done, double* out_ssim, double* out_mse, cudaStream_t s) {
  FrameLayout L = frame_layout(ws, ws_bytes, nf, h, w, nf - 1);
  if (ws == nullptr || ((uintptr_t)ws & 255) || L.bytes > ws_bytes) {
    set_error("hippo_pattern_separation: frame workspace of %zu bytes needed (256-byte aligned), got %zu", L.bytes, ws_bytes);
    return HIPPO_E_WORKSPACE;
  }
  const int npairs = nf - 1;
  if (npairs <= 0) return HIPPO_OK;
  gray_launch(frames, nf, h, w, ch, L, s, pair_done != nullptr ? 1 : 0);
  HIPPO_CUDA(cudaGetLastError());
  const int64_t nitems = (int64_t)npairs * L.nparts;
  const SsimArgs A{L.gray, L.gstride, h, w, L.pitch, nullptr, nullptr, L.minmax, 0, L.bh, L.nbands, L.nchunks,
                   L.part_ssim, L.part_sse, pair_done, out_ssim, out_mse};
  ssim_launch(L, A, nitems, kSsimLiveSmem, s);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
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
  HIPPO_REQUIRE(nf <= 65535 && npairs <= 65535, "hippo_frame_pairs: at most 65535 frames / pairs per call");
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  cudaStream_t s = (cudaStream_t)stream;
  FrameLayout L = frame_layout(ws, ws_bytes, nf, h, w, npairs);
  if (ws == nullptr || ((uintptr_t)ws & 255) || L.bytes > ws_bytes) {
    set_error("hippo_frame_pairs: workspace of %zu bytes needed (256-byte aligned), got %zu", L.bytes, ws_bytes);
    return HIPPO_E_WORKSPACE;
  }
  gray_launch(frames, nf, h, w, ch, L, s);
  HIPPO_CUDA(cudaGetLastError());
  const int64_t nitems = (int64_t)npairs * L.nparts;
  const SsimArgs A{L.gray, L.gstride, h, w, L.pitch, pair_a, pair_b, L.minmax, range_mode, L.bh, L.nbands, L.nchunks,
                   L.part_ssim, L.part_sse, nullptr, nullptr, nullptr};
  ssim_launch(L, A, nitems, 0, s);
  HIPPO_CUDA(cudaGetLastError());
  ssim_finalize_kernel<<<(npairs + 127) / 128, 128, 0, s>>>(L.part_ssim, L.part_sse, npairs, L.nparts, h, w,
                                                           out_ssim, out_mse);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

}  // extern "C"
