// tcgen05 / TMA / mbarrier building blocks shared by the tensor-core kernels (corr_volume.cu, local_corr_tc.cu).
#pragma once
#include <cuda.h>

#include <cstdint>
#include <mutex>

#include "common.cuh"

namespace eem {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// L2 cache policies: the feature-map operands (a few MB per sample, re-read by every tile) are kept
// with evict_last; the volume being written (hundreds of MB, not re-read by this kernel) is marked
// evict_first so that it does not flush the operands out of the 126 MB L2.
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_evict_first(float* p, float v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
// Predicated form: a single @p STG, so the unrolled epilogue has no branches.
__device__ __forceinline__ void st_evict_first_if(float* p, float v, uint64_t pol, bool pred) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %3, 0;\n"
      "@q st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;\n"
      "}\n" ::"l"(p), "f"(v), "l"(pol), "r"((int)pred)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "l"(pol)
      : "memory");
}
// Multicast variant: the box lands at the same shared-memory offset in every CTA of `mask` and
// completes bytes on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar,
                                               uint16_t mask, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5, %6;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "h"(mask), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// One lane of a CONVERGED warp.  The producer / MMA warps run their loops with all 32 lanes and
// predicate only the issuing instructions with this, so the compiler can keep descriptors and
// addresses in uniform registers; under `if (lane == 0)` it wraps every UTCHMMA / UTMALDG in an
// ELECT + R2UR.BROADCAST + BRA.U.ANY loop (~70 cycles per MMA issue, measured with the clock64 probe).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive (once the MMAs issued so far have completed) on the mbarrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor for an MN-major fp32/tf32 operand.  For MN-major tf32 the only
// swizzled layout tcgen05 accepts is "128B swizzle with 32B atomicity" (layout type 1; CuTe's
// Layout_MN_SW128_32B_Atom, Swizzle<2,5,2>): an atom is 32 positions (128 B) x 4 channel rows and
// the four 32-byte chunks of a row are XOR-permuted with (row % 4).  TMA writes exactly that with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: a box is [BK channel rows][32 positions = 128 B], i.e. BK/4
// atoms stacked along K.  Canonical form ((8,n),(4,k)) : ((1,LBO),(8,SBO)) in uint128 units:
//   leading-byte offset = distance between 32-position chunks (one box, 4 KiB here)
//   stride-byte offset  = distance between 4-channel groups   (512 B)
// One kind::tf32 MMA consumes K = 8 channels = two such groups; the next MMA starts 1 KiB further.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lo_fields, uint32_t hi_fields) {
  return ((uint64_t)hi_fields << 32) | (uint64_t)(lo_fields | ((smem_addr >> 4) & 0x3fff));
}

// Host-side encoding of the constant descriptor fields.
inline void desc_fields(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type, uint32_t* lo, uint32_t* hi) {
  *lo = ((lbo_bytes >> 4) & 0x3fff) << 16;                               // bits [16,30): leading byte offset
  *hi = ((sbo_bytes >> 4) & 0x3fff) | (1u << 14) | (layout_type << 29);  // [32,46) SBO, [46,48) version 1, [61,64) layout
}

// ---- host side: cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency) ----

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static std::mutex mu;
  static EncodeTiledFn fn = nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

}  // namespace tc
}  // namespace eem
