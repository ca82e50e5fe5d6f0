/*
 * hippo_b200.h — C ABI of libhippo_b200.so
 *
 * B200 (sm_100a) implementation of HippoMM's data-parallel memory hot path.
 * Every entry point replaces (part of) one plain-Python call site of the
 * reference; the file:line each one stands in for is cited on the declaration
 * (paths relative to the reference checkout, `hm` = hippomm/core/
 * hippocampal_memory.py, `bp` = hippomm/core/batch_process.py, `vo` =
 * hippomm/utils/vector_ops.py).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch / C++ types.
 *   - Every pointer is a DEVICE pointer unless the name ends in `_host`.
 *     The caller (PyTorch) owns every buffer; nothing is allocated inside a
 *     hot call.  Scratch comes in through (ws, ws_bytes); ask the matching
 *     *_workspace_bytes() for the size.  Workspaces need 256-byte alignment.
 *   - Every call takes the cudaStream_t to launch on (as void*), is
 *     asynchronous with respect to the host and is reentrant across streams
 *     as long as workspaces are not shared.
 *   - Every call returns a hippo_status (0 = OK, <0 = error); the message of
 *     the last error on the calling thread is hippo_last_error().
 *   - There is no CPU fallback: on a device that is not sm_100 the compute
 *     calls return HIPPO_E_ARCH.
 */
#ifndef HIPPO_B200_H_
#define HIPPO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIPPO_ABI_VERSION 2

typedef int32_t hippo_status;
#define HIPPO_OK           0
#define HIPPO_E_BADARG    (-1)  /* null pointer, bad size, bad enum            */
#define HIPPO_E_ARCH      (-2)  /* current device is not sm_100                */
#define HIPPO_E_CUDA      (-3)  /* a CUDA runtime / driver call failed         */
#define HIPPO_E_NCCL      (-4)  /* reserved (collectives live on the torch side)*/
#define HIPPO_E_WORKSPACE (-5)  /* workspace missing / too small / misaligned  */

/* element types of caller-provided rows / samples */
#define HIPPO_F32  0
#define HIPPO_F64  1
#define HIPPO_BF16 2
#define HIPPO_I16  3

/* Largest k one search launch keeps per query (larger k: the host wrapper
 * calls again with the `after_*` cursor, see hippo_topk_*).                 */
#define HIPPO_TOPK_MAX 16

/* ---- library ------------------------------------------------------------ */
int32_t      hippo_abi_version(void);
const char*  hippo_last_error(void);
/* HIPPO_OK iff the CURRENT device can run the kernels (compute capability 10.x). */
hippo_status hippo_device_check(void);
/* Number of SMs of the current device (grid sizing on the host side). */
int32_t      hippo_sm_count(void);

/* ---- memory bank -------------------------------------------------------- */
/*
 * Build the device-resident bank from caller rows: bf16 copy of the rows plus
 * the fp32 L2 norm of every ORIGINAL row.  Replaces the per-query
 * `np.linalg.norm(b, axis=1)` of vo:179 (97 % of the reference's search time)
 * and the row normalisation of hm:951.
 *   rows      [n, d] of `dtype` (HIPPO_F32 / HIPPO_F64 / HIPPO_BF16), row stride `ld` elements
 *   bank      [n, d] bf16 out (row stride d), d % 64 == 0 (host zero-pads)
 *   norm      [n] fp32 out, ||row||_2 (sqrt of the fp32-rounded sum of squares)
 *   inexact   optional int32*: OR-ed with 1 if any element changed when rounded to bf16
 */
hippo_status hippo_bank_build(const void* rows, int32_t dtype, int64_t n, int32_t d, int64_t ld,
                              void* bank, float* norm, int32_t* inexact, void* stream);

/* ---- detailed-recall feature search (vo:151-188, callers hm:3153, hm:3304) */
/*
 * Search results are (score, row) pairs ordered by score descending; ties
 * are broken by LOWER row first (the reference's argsort leaves ties
 * unordered); NaN scores (zero-norm rows) sort above every number exactly as
 * `np.argsort` puts them last (vo:185).  score = dot / (norm_b * norm_a) in
 * fp32 with IEEE division, the operation order of vo:182.
 *
 *   bank/norm   from hippo_bank_build, n rows of d
 *   q           query vector(s) in fp32, [nq, d], row stride d
 *   k           1..HIPPO_TOPK_MAX
 *   row_base    added to local row numbers in out_idx (shard offset)
 *   after_key   optional uint64[nq] cursor: only rows ordered strictly AFTER
 *               this packed (score,row) key are considered (NULL = all rows).
 *               Pass out_key of the previous call to page through k > HIPPO_TOPK_MAX.
 *   out_idx     [nq, k] int64, -1 where fewer than k rows qualify
 *   out_score   [nq, k] fp32
 *   out_key     optional [nq, k] uint64 packed order keys (0 = empty slot);
 *               this is also the payload ranks exchange for the sharded merge
 */
size_t       hippo_topk_single_workspace_bytes(int64_t n, int32_t d, int32_t k);
/* One query: bandwidth-bound GEMV + warp-level top-k (north_star "single-query search"). */
hippo_status hippo_topk_single(const void* bank, const float* norm, int64_t n, int32_t d,
                               const float* q, int32_t k, int64_t row_base,
                               const uint64_t* after_key,
                               int64_t* out_idx, float* out_score, uint64_t* out_key,
                               void* ws, size_t ws_bytes, void* stream);

size_t       hippo_topk_batched_workspace_bytes(int64_t n, int32_t d, int32_t nq, int32_t k);
/* nq queries at once: tcgen05 similarity contraction with the top-k fused into the
 * TMEM epilogue (north_star "batched search").  The reference issues nq
 * sequential vo:151 calls. */
hippo_status hippo_topk_batched(const void* bank, const float* norm, int64_t n, int32_t d,
                                const float* q, int32_t nq, int32_t k, int64_t row_base,
                                const uint64_t* after_key,
                                int64_t* out_idx, float* out_score, uint64_t* out_key,
                                void* ws, size_t ws_bytes, void* stream);

/*
 * Search straight over the caller's rows -- the drop-in signature of vo:151-188 hands the (N, D)
 * feature array in with every call, in fp32 (fresh features) or fp64 (ThetaEvents reloaded from
 * JSON, hm:391).  One streaming pass, no bank: dot product and row norm from the same registers,
 * accumulated in fp64, rounded once to the arrays' own precision, then dot / (|b| * |a|) in the
 * operation order of vo:182 (fp32 IEEE chain when rows and query are both fp32, fp64 otherwise,
 * as NumPy promotes).  Nothing is rounded to bf16 on this path.
 *   rows      [n, d] HIPPO_F32 / HIPPO_F64, row stride `ld` elements (any d >= 1)
 *   q         [d] HIPPO_F32 / HIPPO_F64
 *   out_*     as hippo_topk_single (out_score is the fp32 rounding of the score; hippo_rescore
 *             returns the winners' scores in fp64)
 */
size_t       hippo_topk_rows_workspace_bytes(int64_t n, int32_t d, int32_t k);
hippo_status hippo_topk_rows(const void* rows, int32_t dtype, int64_t n, int32_t d, int64_t ld,
                             const void* q, int32_t q_dtype, int32_t k, int64_t row_base,
                             const uint64_t* after_key,
                             int64_t* out_idx, float* out_score, uint64_t* out_key,
                             void* ws, size_t ws_bytes, void* stream);
/*
 * Exact second stage behind the bf16 searches: the same expression for `kc` candidate rows per
 * query, from the ORIGINAL fp32 / fp64 rows (one warp per candidate).  cand_idx holds GLOBAL row
 * numbers (rows [row_base, row_base + n) are local; -1 or a foreign row yields key 0).
 *   q, q_ld    queries [nq, d], row stride q_ld elements; q_ld = 0: ONE query shared by all nq
 *              candidate lists (the per-event lists of hippo_topk_segmented)
 *   out_key    optional [nq, kc] uint64 order keys of the exact scores (feed hippo_topk_merge
 *              with nparts = 1 to re-rank)
 *   out_score  optional [nq, kc] fp64 exact scores
 */
hippo_status hippo_rescore(const void* rows, int32_t dtype, int64_t n, int32_t d, int64_t ld,
                           int64_t row_base, const void* q, int32_t q_dtype, int64_t q_ld, int32_t nq,
                           const int64_t* cand_idx, int32_t kc,
                           uint64_t* out_key, double* out_score, void* stream);

/*
 * Merge per-shard results: keys [nparts, nq, k_in] packed order keys (as written to
 * out_key, rows already global because every shard passed its row_base) ->
 * the k best per query.  Runs after the all-gather of the sharded search
 * (SURVEY §8e); also used inside the single-GPU kernels' final stage.
 */
hippo_status hippo_topk_merge(const uint64_t* keys, int32_t nparts, int32_t nq, int32_t k_in,
                              int32_t k, int64_t* out_idx, float* out_score, uint64_t* out_key,
                              void* stream);

/*
 * Sharded search, collective fused into the merge: every rank's kernel pushes its
 * out_key block [nq, k_in] into slot [rank] of EVERY rank's gather buffer through peer-mapped
 * pointers (NVLink stores, no NCCL call), publishes an epoch flag, waits for the flags of all
 * ranks in its own buffer and merges world x k_in candidates per query.  All ranks obtain the
 * identical result, equal to hippo_topk_merge over an all-gather of the same keys.
 *   peer_bases  DEVICE array [world] of pointers: base of rank r's exchange buffer as mapped into
 *               THIS process (one symmetric allocation of hippo_topk_exchange_bytes() bytes per rank,
 *               zero-filled once before the first call; e.g. torch symmetric memory's buffer_ptrs_dev)
 *   epoch       1, 2, 3, ... incremented by every rank on every call (selects one of two gather buffers)
 * Every rank of the group must make the matching call; the kernel spins until its peers' have run.
 */
size_t       hippo_topk_exchange_bytes(int32_t world, int32_t nq, int32_t k);
hippo_status hippo_topk_exchange_merge(const uint64_t* local_keys, int32_t nq, int32_t k_in, int32_t k,
                                       void* const* peer_bases, size_t buf_bytes, int32_t rank, int32_t world,
                                       uint32_t epoch, int64_t* out_idx, float* out_score, uint64_t* out_key,
                                       void* stream);

/*
 * The sharded searches with NOTHING between the local pass and the exchange: one call = this rank's whole
 * share of a sharded query (batch).
 *   hippo_topk_batched_sharded  hippo_topk_batched, but the per-split lists of the tcgen05 pass go straight
 *                               into the exchange kernel, which merges them per query just before its push
 *                               phase (local merge + NVLink push + global merge: one launch instead of two).
 *   hippo_topk_single_sharded   hippo_topk_single, but the LAST CTA of the GEMV to finish (ticket counter)
 *                               merges the per-block lists, pushes the k keys to every peer, exchanges flags
 *                               and merges the world x k candidates: a sharded query is ONE launch per rank.
 * peer_bases / buf_bytes / rank / world / epoch as for hippo_topk_exchange_merge (the same symmetric buffers
 * and the same epoch sequence serve all three); world == 1 needs no buffers.  Results are identical on every
 * rank and equal to the single-GPU answer over the whole bank.
 */
hippo_status hippo_topk_batched_sharded(const void* bank, const float* norm, int64_t n, int32_t d,
                                        const float* q, int32_t nq, int32_t k, int64_t row_base,
                                        const uint64_t* after_key,
                                        void* const* peer_bases, size_t buf_bytes, int32_t rank,
                                        int32_t world, uint32_t epoch,
                                        int64_t* out_idx, float* out_score, uint64_t* out_key,
                                        void* ws, size_t ws_bytes, void* stream);
hippo_status hippo_topk_single_sharded(const void* bank, const float* norm, int64_t n, int32_t d,
                                       const float* q, int32_t k, int64_t row_base,
                                       const uint64_t* after_key,
                                       void* const* peer_bases, size_t buf_bytes, int32_t rank,
                                       int32_t world, uint32_t epoch,
                                       int64_t* out_idx, float* out_score, uint64_t* out_key,
                                       void* ws, size_t ws_bytes, void* stream);

/* ---- detailed recall over all events at once ------ */
/*
 * The reference searches event by event: `for event in long_term_store:
 * top_k_cosine_similarity(query, event.features[modality], k=5)` (hm:3143-3153,
 * hm:3295-3304).  Here every event's rows sit in ONE bank and `seg_offsets`
 * [nseg+1] (int64, ascending, seg_offsets[0] = 0, seg_offsets[nseg] = n) marks
 * the events, so one streaming pass serves the whole store.
 *
 * hippo_scores_single: out_score[row] = dot / (norm_b * norm_a) for every row
 * (fp32, IEEE division, the operation order of vo:182).
 */
hippo_status hippo_scores_single(const void* bank, const float* norm, int64_t n, int32_t d,
                                 const float* q, float* out_score, void* stream);
/*
 * hippo_topk_segmented: the pass above into the workspace, then the k best rows of
 * every event (score descending, NaN first, lower row first on ties).
 *   out_idx    [nseg, k] int64 event-LOCAL row numbers, -1 where the event has fewer than k rows
 *   out_score  [nseg, k] fp32
 *   out_max    optional [nseg] fp32: max of the event's top-k as np.max sees it
 *              (hm:3156 / hm:3307 compare it with 0.4 to route an event to the LLM path)
 */
size_t       hippo_topk_segmented_workspace_bytes(int64_t n);
hippo_status hippo_topk_segmented(const void* bank, const float* norm, int64_t n, int32_t d,
                                  const float* q, const int64_t* seg_offsets, int32_t nseg, int32_t k,
                                  int64_t* out_idx, float* out_score, float* out_max,
                                  void* ws, size_t ws_bytes, void* stream);
/*
 * Tail of the recall loop (hm:3258-3277, hm:3366-3381): candidate (event e, rank r)
 * is live iff enabled[e] (NULL = all), idx >= 0 and idx < time_offsets[e+1]-time_offsets[e]
 * (the `idx < len(event.frame_times)` guard of hm:3262, which applies all-frame indices to
 * the key-frame time table); the m best live candidates by similarity (stable: earlier
 * event, then better rank, first -- Python's sort with reverse=True) come back with their
 * windows [max(0, t - pad), t + pad].
 *   times        fp64, all events' time tables concatenated; time_offsets [nseg+1] int64
 *   out_event    [m] int32 (-1 = unused slot), out_idx [m] int64, out_score [m] fp32,
 *   out_window   [m, 2] fp64 (start, end), out_count int32* (live entries written)
 */
hippo_status hippo_recall_windows(const int64_t* seg_idx, const float* seg_score, int32_t nseg, int32_t k,
                                  const int64_t* time_offsets, const double* times,
                                  const uint8_t* enabled, double pad, int32_t m,
                                  int32_t* out_event, int64_t* out_idx, float* out_score,
                                  double* out_window, int32_t* out_count, void* stream);

/* ---- memory consolidation (hm:944-967 _select_key_frames) --------------- */
/*
 * Greedy redundancy filter: row 0 is kept; row i is kept iff
 * cos(row i, row j) < gamma for every kept j < i (hm:958-961).  gamma is
 * compared in fp32 like NumPy does for an fp32 matrix and a Python float.
 * The rule only consults similarities against KEPT rows, so the rows are
 * processed in bands: a band is contracted on tcgen05 against the rows kept so
 * far (compacted) and against itself (strict lower triangle), leaving bits
 * (never the fp32 matrix); pairs whose tensor-core similarity is within `band`
 * of gamma are re-evaluated from the fp32 rows before the band's greedy scan.
 * The whole call is asynchronous: how many rows are final is read from device
 * memory by the next band's kernels.  Decisions equal the full-matrix rule.
 *   feats      [n, d] fp32 rows (time-ordered), row stride d, d % 64 == 0
 *   out_keep   [n] int64 kept row numbers ascending (first *out_count valid)
 *   out_count  int32* (device)
 *   out_stats  optional int32[4] (device): {pairs re-evaluated, recheck overflow,
 *              bf16-inexact flag, reserved}
 */
size_t       hippo_consolidate_workspace_bytes(int64_t n, int32_t d);
hippo_status hippo_consolidate(const float* feats, int64_t n, int32_t d, float gamma,
                               float band_exact, float band_inexact,
                               int64_t* out_keep, int32_t* out_count, int32_t* out_stats,
                               void* ws, size_t ws_bytes, void* stream);
/*
 * The same with explicit sizing: `band_rows` rows per band (0 = default 8,192; rounded to 512) and
 * room for `uncertain_cap` near-threshold pairs per band (0 = default 64 * (band + 1024) + 2^20).
 * out_stats[1] != 0 reports that a band produced more near-threshold pairs than fit; the pairs
 * beyond the capacity kept their tensor-core decision, so the caller must call again with a larger
 * capacity (the Python wrapper does) before trusting the result.
 */
size_t       hippo_consolidate_ex_workspace_bytes(int64_t n, int32_t d, int32_t band_rows,
                                                  int64_t uncertain_cap);
hippo_status hippo_consolidate_ex(const float* feats, int64_t n, int32_t d, float gamma,
                                  float band_exact, float band_inexact,
                                  int32_t band_rows, int64_t uncertain_cap,
                                  int64_t* out_keep, int32_t* out_count, int32_t* out_stats,
                                  void* ws, size_t ws_bytes, void* stream);

/*
 * Profiling aid (never needed in production): when the environment variable HIPPO_CONS_TIMING is set,
 * hippo_consolidate[_ex] brackets every launch with CUDA events, synchronises at the end and keeps the
 * time per stage of the LAST call on the calling thread; this returns them in milliseconds (HOST array):
 * {similarity bits on tcgen05, fp32 re-evaluation, greedy scan, bank build + band staging / compaction}.
 */
void         hippo_debug_consolidate_timing(double* out4_host);

/* ---- temporal pattern separation ---------------------------------------- */
/*
 * Frame-pair scoring: BGR -> gray (cv2's fixed-point BGR2GRAY) -> mean SSIM
 * (7x7 uniform window, K1=.01, K2=.03, sample covariance: the scikit-image
 * defaults the reference relies on) and the MSE of the gray frames.
 * Replaces hm:980-991 (_compute_frame_similarity; range_mode 0: data_range =
 * max-min of the FIRST frame of the pair, in uint8) and bp:32-71
 * (compute_frame_difference; range_mode 1: frames scaled to [0,1],
 * data_range 1.0).
 *   frames     [nf, h, w, ch] uint8, ch = 3 (BGR) or 1 (gray), densely packed
 *   pair_a/b   [np] int32 frame numbers; pair p scores (frames[a], frames[b]);
 *              NULL = adjacent pairs (p+1, p) for p in [0, nf-1) as hm:1052-1056 scans them
 *   out_ssim   [np] fp64 mean SSIM (NaN where the reference yields NaN)
 *   out_mse    [np] fp64 mean squared gray difference on the [0,1] scale (bp:67)
 */
size_t       hippo_frame_pairs_workspace_bytes(int32_t nf, int32_t h, int32_t w, int32_t npairs);
hippo_status hippo_frame_pairs(const uint8_t* frames, int32_t nf, int32_t h, int32_t w, int32_t ch,
                               const int32_t* pair_a, const int32_t* pair_b, int32_t npairs,
                               int32_t range_mode, double* out_ssim, double* out_mse,
                               void* ws, size_t ws_bytes, void* stream);

/*
 * Audio energy pyramid: per-sample mono x^2 summed in fp64 over blocks of 16
 * and of 512 samples, so that any window's sum of squares (hm:993-1000) is a
 * short exact sum instead of a fresh pass over the samples.
 *   pcm      [ns, nch] of `dtype` (HIPPO_I16 -> x = k/32768; HIPPO_F32; HIPPO_F64)
 *   out_e16  [ceil(ns/16)] fp64, out_e512 [ceil(ns/512)] fp64
 */
hippo_status hippo_audio_energy(const void* pcm, int32_t dtype, int64_t ns, int32_t nch,
                                double* out_e16, double* out_e512, void* stream);

/*
 * RMS level in dB of arbitrary windows [start, start+len) (hm:993-1000), from the pyramid.
 *   out_db [nwin] fp64: 20*log10(sqrt(mean x^2)), -100 when the window is all zero
 */
hippo_status hippo_audio_levels(const void* pcm, int32_t dtype, int64_t ns, int32_t nch,
                                const double* e16, const double* e512,
                                const int64_t* win_start, const int64_t* win_len, int32_t nwin,
                                double* out_db, void* stream);

/*
 * The boundary state machine of hm:1034-1111, operation for operation in
 * fp64, for `nstreams` independent streams (one warp each).
 * Per stream s (all arrays indexed by the stream's offsets):
 *   ssim        adjacent-pair SSIM, pair p = (frame p+1, frame p); NULL = no video
 *   frame_times fp64 [nframes], non-decreasing
 *   pcm/e16/e512 as above; NULL pcm = no audio
 *   out_bounds  [max_segments, 2] fp64 (start, end); out_count int32
 * Streams are described by a device array of hippo_stream_desc.
 */
typedef struct hippo_stream_desc {
  const double*  ssim;         /* [nframes-1] or NULL */
  const double*  frame_times;  /* [nframes]   or NULL */
  int64_t        nframes;
  const void*    pcm;          /* [ns, nch]   or NULL */
  const double*  e16;
  const double*  e512;
  int64_t        ns;
  int32_t        nch;
  int32_t        pcm_dtype;
  double         sample_rate;  /* audio_sample_rate as the reference receives it */
  double*        out_bounds;   /* [max_segments, 2] */
  int32_t*       out_count;    /* segments written; -needed if max_segments was too small */
  int32_t        max_segments;
  int32_t        reserved;
} hippo_stream_desc;

hippo_status hippo_segment_boundaries(const hippo_stream_desc* streams, int32_t nstreams,
                                      double max_segment_duration, double min_segment_duration,
                                      double frame_similarity_threshold,
                                      double audio_silence_threshold, void* stream);

/*
 * The same state machine in resumable form, so that the boundary chain of a stream (sequential by
 * nature: hm:1036 loops on current_start) can run UNDER the SSIM kernels of the stream's later
 * frames.  `states[s]` (device memory, zero-filled before the first call) carries (current_start,
 * frame hint, segments written) from launch to launch.  A launch with final_pass == 0 takes only the
 * segments whose window [current_start, current_start + max] is covered by the first `frames_ready`
 * frames of every stream (pairs p < frames_ready - 1 of `ssim` are final), then suspends; the launch
 * with final_pass != 0 runs the chain to the end.  The segments written are identical to one
 * hippo_segment_boundaries call on the complete inputs.  Audio (pcm, pyramid) must be complete from
 * the first launch on.
 */
typedef struct hippo_segment_state {
  double  current_start;
  int64_t hint;
  int32_t count;
  int32_t done;
} hippo_segment_state;

hippo_status hippo_segment_boundaries_resume(const hippo_stream_desc* streams, int32_t nstreams,
                                             hippo_segment_state* states, int64_t frames_ready,
                                             int32_t final_pass,
                                             double max_segment_duration, double min_segment_duration,
                                             double frame_similarity_threshold,
                                             double audio_silence_threshold, void* stream);

/*
 * One stream, all stages, overlapped (hm:1002-1114 with hm:980-1000 underneath).  The boundary state machine is
 * launched FIRST on side_streams_host[2] (give it a high-priority stream) in follow mode: one CTA on an SM of its
 * own that polls each adjacent pair's SSIM as the SSIM warps deliver it; the audio pyramid runs on
 * side_streams_host[1], gray conversion + one SSIM launch on side_streams_host[0]; a final resumable pass on [2]
 * behind everything completes the chain if the follower gave up waiting (it never waits longer than 4 ms, env
 * HIPPO_FOLLOW_US: serialising profilers and launch-blocking debug runs stay correct).  `stream` is joined at both
 * ends (events), so the call behaves like any other asynchronous call on `stream`.  The results equal
 * hippo_frame_pairs + hippo_audio_energy + hippo_segment_boundaries run one after the other.
 * `chunk_pairs` is ignored (kept for ABI v2: the first version took the frames in chunks).
 *   frames / frame_times  NULL = no video;  pcm NULL = no audio (then out_e16 / out_e512 may be NULL)
 *   out_ssim, out_mse     [nf - 1] fp64;  out_e16 [ceil(ns/16)], out_e512 [ceil(ns/512)] fp64
 *   side_streams_host     HOST array of three cudaStream_t distinct from `stream` and from each other
 * The events used for the ordering are created once per host thread and kept (the only allocation
 * any entry point of this library performs).
 * With audio the call sets aside up to 8 MB of L2 for persisting accesses (cudaLimitPersistingL2CacheSize, once per
 * host thread) and gives side_streams_host[1] and [2] an access-policy window over out_e512: the chain reads those sums
 * while the frame kernels stream ~0.9 GB through L2, and without the window every read goes to DRAM behind that
 * traffic (a segment then takes ~3 us instead of 1.4); the window is taken off the two streams again before the call
 * returns.  HIPPO_PATTERN_L2PIN=0 switches it off.
 * A segment whose span holds a window KNOWN to be silent is settled by the audio boundary (it overwrites the video
 * boundary, hm:1061-1077), so the chain does not wait for SSIM values of that span that are still pending.
 */
size_t       hippo_pattern_separation_workspace_bytes(int32_t nf, int32_t h, int32_t w, int32_t chunk_pairs);
hippo_status hippo_pattern_separation(const uint8_t* frames, int32_t nf, int32_t h, int32_t w, int32_t ch,
                                      const double* frame_times,
                                      const void* pcm, int32_t pcm_dtype, int64_t ns, int32_t nch,
                                      double sample_rate,
                                      double max_segment_duration, double min_segment_duration,
                                      double frame_similarity_threshold, double audio_silence_threshold,
                                      int32_t chunk_pairs,
                                      double* out_ssim, double* out_mse, double* out_e16, double* out_e512,
                                      double* out_bounds, int32_t* out_count, int32_t max_segments,
                                      void* ws, size_t ws_bytes, void* const* side_streams_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIPPO_B200_H_ */
