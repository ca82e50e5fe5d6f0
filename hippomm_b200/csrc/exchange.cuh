// Layout and primitives of the peer-memory gather buffers shared by exchange.cu (batched search) and the fused
// tail of the single-query GEMV (topk_single.cu).
#pragma once
#include "common.cuh"

namespace hippo {

constexpr size_t kXchgHeader = 1024;   // [0,256) flags uint32[2][32]; [256,264) CTA counters uint32[2]
constexpr int kXchgMaxWorld = 32;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// gather slot [par][rank] of the buffer at `base`
__device__ __forceinline__ uint64_t* xchg_slot(unsigned char* base, uint32_t par, int world, int rank, size_t slot_stride) {
  return reinterpret_cast<uint64_t*>(base + kXchgHeader) + ((size_t)par * world + rank) * slot_stride;
}

// launches exchange_merge_kernel; nparts > 1: local_keys holds the UNMERGED per-split lists [nparts][nq][k_in]
hippo_status exchange_launch(const uint64_t* local_keys, int nparts, int nq, int k_in, int k, void* const* peer_bases,
                             size_t buf_bytes, int rank, int world, uint32_t epoch, int64_t* out_idx, float* out_score,
                             uint64_t* out_key, cudaStream_t s);

}  // namespace hippo
