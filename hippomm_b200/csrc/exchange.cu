// hippo_topk_exchange_merge: the collective of the sharded search fused into ONE kernel over peer memory.
//
// After the local search every rank holds its k best (score,row) order keys per query; the global answer
// is the k best of the union.  Instead of an NCCL all-gather followed by a merge kernel, every
// rank's kernel
//   0. (when handed the UNMERGED per-split lists of the local search) merges them per query, one warp per query,
//   1. PUSHES its keys straight into slot [rank] of every peer's gather buffer with plain stores through
//      the NVLink-mapped peer pointers (the buffers are one symmetric allocation, mapped by the host),
//   2. publishes a per-(peer, parity) epoch flag with a system-scope release once its last CTA has pushed,
//   3. waits for the flags of all ranks in its OWN buffer (system-scope acquire) and merges the world x k
//      candidates per query from local memory (lower row wins ties, so every rank computes the same list).
// Two gather buffers alternate by epoch parity: a rank can start call e+1 only after every peer has
// published flag e, i.e. has finished READING buffer (e-1) % 2 ... so nobody overwrites a buffer that is
// still being merged (see the argument in DESIGN.md §5).
#include "exchange.cuh"

namespace hippo {

__global__ void __launch_bounds__(256)
exchange_merge_kernel(const uint64_t* __restrict__ local_keys /*[nparts][nq][k_in]*/, int nparts, int nq, int k_in, int k,
                      unsigned char* const* peer_bases, size_t slot_stride /* uint64 elements per (parity, rank) slot */,
                      int rank, int world, uint32_t epoch, int64_t* __restrict__ out_idx,
                      float* __restrict__ out_score, uint64_t* __restrict__ out_key) {
  const uint32_t par = epoch & 1u;
  unsigned char* own = peer_bases[rank];
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  // keys per query in the gather slots: the merged k when this kernel merges the local lists itself
  const int k_slot = nparts > 1 ? k : k_in;

  // ---- 1. push: slot [par][rank] of every rank's buffer (own included) <- this rank's keys ----
  if (nparts == 1) {
    const size_t n64 = (size_t)nq * k_in;
    for (int p = 0; p < world; ++p) {
      uint64_t* dst = xchg_slot(peer_bases[p], par, world, rank, slot_stride);
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n64; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = local_keys[i];
    }
  } else {
    for (int qi = blockIdx.x * warps + (threadIdx.x >> 5); qi < nq; qi += gridDim.x * warps) {
      uint64_t L[kLaneList];
      lane_list_clear(L);
      const int total = nparts * k_in;
      for (int i = lane; i < total; i += 32) {
        const int p = i / k_in, j = i - p * k_in;
        lane_list_insert(L, k, local_keys[((size_t)p * nq + qi) * k_in + j]);
      }
      const uint64_t mine = warp_select_best(L, k, lane);
      if (lane < k)
        for (int p = 0; p < world; ++p) xchg_slot(peer_bases[p], par, world, rank, slot_stride)[(size_t)qi * k + lane] = mine;
    }
  }
  __threadfence_system();
  __syncthreads();

  // ---- 2. the last CTA to finish pushing publishes the epoch to every rank ----
  __shared__ uint32_t s_last;
  uint32_t* counters = reinterpret_cast<uint32_t*>(own + 256);
  if (threadIdx.x == 0) {
    const uint32_t old = atomicAdd(&counters[par], 1u);
    s_last = (old == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) counters[par] = 0;                 // ready for epoch + 2
    __threadfence_system();
    if ((int)threadIdx.x < world) {
      uint32_t* flags = reinterpret_cast<uint32_t*>(peer_bases[threadIdx.x]);
      st_release_sys(&flags[par * kXchgMaxWorld + rank], epoch);
    }
  }

  // ---- 3. wait for every rank's keys in the own buffer, then merge ----
  if ((int)threadIdx.x < world) {
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(own);
    while (ld_acquire_sys(&flags[par * kXchgMaxWorld + threadIdx.x]) != epoch) __nanosleep(64);
  }
  __syncthreads();
  __threadfence_system();

  const uint64_t* gathered = xchg_slot(own, par, world, 0, slot_stride);
  for (int qi = blockIdx.x * warps + (threadIdx.x >> 5); qi < nq; qi += gridDim.x * warps) {
    uint64_t mine;
    if (k <= kLaneList) {
      uint64_t L[kLaneList];
      lane_list_clear(L);
      for (int i = lane; i < world * k_slot; i += 32) {
        const int p = i / k_slot, j = i - p * k_slot;
        lane_list_insert(L, k, __ldcg(&gathered[(size_t)p * slot_stride + (size_t)qi * k_slot + j]));
      }
      mine = warp_select_best(L, k, lane);
      if (lane < k) {
        const size_t o = (size_t)qi * k + lane;
        if (out_idx) out_idx[o] = mine ? (int64_t)key_row(mine) : -1;
        if (out_score) out_score[o] = mine ? key_score(mine) : 0.f;
        if (out_key) out_key[o] = mine;
      }
      continue;
    }
    uint64_t prev = ~0ull;
    for (int r = 0; r < k; ++r) {
      uint64_t best = 0;
      if (prev != 0) {
        for (int i = lane; i < world * k_slot; i += 32) {
          const int p = i / k_slot, j = i - p * k_slot;
          const uint64_t c = __ldcg(&gathered[(size_t)p * slot_stride + (size_t)qi * k_slot + j]);
          if (c < prev && c > best) best = c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
          best = other > best ? other : best;
        }
      }
      if (lane == 0) {
        const size_t o = (size_t)qi * k + r;
        if (out_idx) out_idx[o] = best ? (int64_t)key_row(best) : -1;
        if (out_score) out_score[o] = best ? key_score(best) : 0.f;
        if (out_key) out_key[o] = best;
      }
      prev = best;
    }
  }
}

hippo_status exchange_launch(const uint64_t* local_keys, int nparts, int nq, int k_in, int k, void* const* peer_bases,
                             size_t buf_bytes, int rank, int world, uint32_t epoch, int64_t* out_idx, float* out_score,
                             uint64_t* out_key, cudaStream_t s) {
  HIPPO_REQUIRE(world >= 1 && world <= kXchgMaxWorld && rank >= 0 && rank < world, "hippo_topk_exchange_merge: bad rank / world");
  HIPPO_REQUIRE(nq >= 0 && k_in >= 1 && k >= 1 && nparts >= 1, "hippo_topk_exchange_merge: bad sizes");
  HIPPO_REQUIRE(nparts == 1 || k <= kLaneList, "hippo_topk_exchange_merge: merging local lists needs k <= %d", kLaneList);
  HIPPO_REQUIRE(epoch != 0, "hippo_topk_exchange_merge: epoch 0 is the cleared state, start at 1");
  if (nq == 0) return HIPPO_OK;
  HIPPO_REQUIRE(local_keys && peer_bases, "hippo_topk_exchange_merge: null pointer");
  const int k_slot = nparts > 1 ? k : k_in;
  HIPPO_REQUIRE(buf_bytes >= hippo_topk_exchange_bytes(world, nq, k_slot),
                "hippo_topk_exchange_merge: symmetric buffer of %zu bytes needed, got %zu",
                hippo_topk_exchange_bytes(world, nq, k_slot), buf_bytes);
  hippo_status st = check_arch();
  if (st != HIPPO_OK) return st;
  const size_t slot_stride = (buf_bytes - kXchgHeader) / ((size_t)2 * world * sizeof(uint64_t));
  // every CTA spins on the peers' flags, so the whole grid must be resident: at most one CTA per SM
  int grid = (nq + 7) / 8;
  const int sms = sm_count();
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  exchange_merge_kernel<<<grid, 256, 0, s>>>(local_keys, nparts, nq, k_in, k, (unsigned char* const*)peer_bases,
                                             slot_stride, rank, world, epoch, out_idx, out_score, out_key);
  HIPPO_CUDA(cudaGetLastError());
  return HIPPO_OK;
}

}  // namespace hippo

extern "C" {

size_t hippo_topk_exchange_bytes(int32_t world, int32_t nq, int32_t k) {
  if (world < 1 || nq < 1 || k < 1) return hippo::kXchgHeader;
  return hippo::align_up(hippo::kXchgHeader + (size_t)2 * world * (size_t)nq * k * sizeof(uint64_t), 256);
}

hippo_status hippo_topk_exchange_merge(const uint64_t* local_keys, int32_t nq, int32_t k_in, int32_t k,
                                       void* const* peer_bases, size_t buf_bytes, int32_t rank, int32_t world,
                                       uint32_t epoch, int64_t* out_idx, float* out_score, uint64_t* out_key,
                                       void* stream) {
  return hippo::exchange_launch(local_keys, 1, nq, k_in, k, peer_bases, buf_bytes, rank, world, epoch, out_idx, out_score,
                                out_key, (cudaStream_t)stream);
}

}  // extern "C"
