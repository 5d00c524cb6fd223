// K3b: batched TF32 GEMM on tcgen05 (sm_100a) for the BACKWARD of the all-pairs pyramid.
//
// The reference gets the gradient of `torch.matmul(fmap1^T, fmap2) / sqrt(D)` (model/corr.py:52-60) from autograd
// inside its training loop (train_mvsec.py:251-258).  With level_l[b,i,j] = s * sum_d f1[b,d,i] * pool^l(f2)[b,d,j]:
//   d f1[b,d,i]         = s * sum_l sum_j pool^l(f2)[b,d,j] * dV_l[b,i,j]     C = A . B^T  (A = f2_l [D,P_l], B = dV_l [P,P_l])
//   d pool^l(f2)[b,d,j] = s * sum_i f1[b,d,i]          * dV_l[b,i,j]     C = A . B    (A = f1 [D,P],   B = dV_l [P,P_l])
// Same contract as eem_batched_gemm_f32 (backward.cu, exact FFMA):
//   C[b] (M x N, row pitch ldc) = alpha * A[b] (M x K row-major) * op(B[b])  (+ C[b] when accumulate != 0)
// with the products in TF32 (10-bit mantissa operands, fp32 accumulate) -- the precision the TF32 forward has.
//
// Mapping.  M = D <= 256 is small and N is large, so the kernel computes C^T tiles: the 128 TMEM lanes hold 128
// consecutive columns n of C (so a warp stores 32 consecutive floats of one C row: 128-byte coalesced) and the TMEM
// columns hold the M rows.  Hence
//   tcgen05 "A" operand = API B:  b_transposed (N x K, k contiguous) -> K-major, one TMA box of 128 rows x 32 k
//                                 plain        (K x N, n contiguous) -> MN-major, four boxes of 32 k x 32 n (the
//                                 128B_ATOM_32B layout of the forward kernels, corr_volume.cu)
//   tcgen05 "B" operand = API A:  (M x K, k contiguous) -> K-major, one TMA box of M rows x 32 k
// K-major tiles are the canonical SWIZZLE_128B layout: 128-byte rows (32 tf32), 8-row groups of 1 KiB (stride byte
// offset 1024), one K = 8 MMA step advances the start address by 32 bytes inside the swizzle atom.
// All tensor maps are 3-D (inner, rows, batch): rows / columns beyond the matrix are zero-filled PER SAMPLE, so ragged
// K (P = 1584 is 49.5 stages), ragged N tiles and tiny levels (P_l = 20) need no special cases.
//
// Roles (320 threads, one persistent CTA per SM, tiles dealt round-robin): warps 0-7 epilogue (TMEM lane quarter w % 4,
// 32-column chunks of parity w / 4), warp 8 TMA producer (4-stage ring of 48 KiB), warp 9 MMA issuer.  Two 256-column
// accumulators in TMEM: the epilogue of tile t overlaps the MMAs of tile t + 1.
#include <cstdlib>

#include "tc_common.cuh"

namespace eem {
namespace {
using namespace tc;

constexpr int kTileN = 128;                      // C columns per tile -> TMEM lanes
constexpr int kMaxM = 256;                       // C rows -> TMEM columns of one accumulator
constexpr int kBK = 32;                          // k per pipeline stage (one 128-byte row)
constexpr int kUK = 8;                           // k per tcgen05.mma kind::tf32
constexpr int kABytes = kTileN * kBK * 4;        // 16 KiB
constexpr int kBBytesMax = kMaxM * kBK * 4;      // 32 KiB
constexpr int kStages = 4;
constexpr int kEpiWarps = 8;
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kTmemCols = 2 * kMaxM;
constexpr int kMaxSeg = 6;                       // K segments summed into one product (the pyramid's levels)

struct GemmTcParams {
  CUtensorMap map_a[kMaxSeg];        // tcgen05 A operand (API B), one map per K segment
  CUtensorMap map_b[kMaxSeg];        // tcgen05 B operand (API A)
  float* C;
  int M, N;
  int K[kMaxSeg];
  int n_seg;
  int64_t ldc, strideC;
  int n_tiles;                       // ceil(N / 128)
  int total_tiles;                   // batch * n_tiles
  float alpha;
  int accumulate;
  uint32_t idesc;
  uint32_t a_lo, a_hi, b_lo, b_hi;   // constant shared-memory descriptor fields (tc::desc_fields)
  uint32_t a_kstep;                  // start-address advance of the A operand per K = 8 step (bytes)
};

struct __align__(1024) GemmTcSmem {
  uint8_t a[kStages][kABytes];
  uint8_t b[kStages][kBBytesMax];
  uint64_t full[kStages], empty[kStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar,
                                            uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "l"(pol)
      : "memory");
}

template <bool kAMn>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ GemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  GemmTcSmem& s = *reinterpret_cast<GemmTcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s.acc_full[i], 1); mbar_init(&s.acc_empty[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;

  if (warp == kProducerWarp) {
    // ===== TMA producer (whole warp converged, one elected lane issues) =====
    const uint64_t pol_a = policy_evict_first();   // dV_l: hundreds of MB, read once
    const uint64_t pol_b = policy_evict_last();    // the feature map of the sample: re-read by every tile
    const uint32_t b_bytes = (uint32_t)p.M * kBK * 4;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / p.n_tiles;
      const int n0 = (tile - b * p.n_tiles) * kTileN;
      for (int seg = 0; seg < p.n_seg; ++seg) {
        const int KB = (p.K[seg] + kBK - 1) / kBK;
        const CUtensorMap* map_a = &p.map_a[seg];
        const CUtensorMap* map_b = &p.map_b[seg];
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&s.empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&s.full[stage], kABytes + b_bytes);
            const int k0 = kb * kBK;
            if constexpr (kAMn) {
#pragma unroll
              for (int ch = 0; ch < kTileN / 32; ++ch)
                tma_load_3d(s.a[stage] + ch * (32 * kBK * 4), map_a, n0 + ch * 32, k0, b, &s.full[stage], pol_a);
            } else {
              tma_load_3d(s.a[stage], map_a, k0, n0, b, &s.full[stage], pol_a);
            }
            tma_load_3d(s.b[stage], map_b, k0, 0, b, &s.full[stage], pol_b);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer (whole warp converged, one elected lane issues) =====
    int stage = 0;
    uint32_t phase = 0;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
      const uint32_t acc = t & 1;
      mbar_wait(&s.acc_empty[acc], ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + acc * kMaxM;
      uint32_t issued = 0;                       // 0 only for the first MMA of the tile: it overwrites the accumulator
      for (int seg = 0; seg < p.n_seg; ++seg) {
        const int KB = (p.K[seg] + kBK - 1) / kBK;
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&s.full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(s.a[stage]);
          const uint32_t b_addr = smem_u32(s.b[stage]);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < kBK / kUK; ++ks)
              tc_mma_tf32(d_tmem, make_desc(a_addr + ks * p.a_kstep, p.a_lo, p.a_hi),
                          make_desc(b_addr + ks * (kUK * 4), p.b_lo, p.b_hi), p.idesc, issued | (uint32_t)ks);
            tc_commit(&s.empty[stage]);
          }
          __syncwarp();
          issued = 1;
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (elect_one()) tc_commit(&s.acc_full[acc]);
      __syncwarp();
    }
  } else {
    // ===== epilogue warps 0-7: TMEM lanes [32 * (w % 4), +32) = 32 consecutive columns n of C =====
    const int quarter = warp & 3, half = warp >> 2;
    const int chunks = p.M / 32;
    const float alpha = p.alpha;
    const bool accumulate = p.accumulate != 0;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
      const uint32_t acc = t & 1;
      const int b = tile / p.n_tiles;
      const int n = (tile - b * p.n_tiles) * kTileN + quarter * 32 + lane;
      const bool n_ok = n < p.N;
      float* cbase = p.C + (int64_t)b * p.strideC + n;
      mbar_wait(&s.acc_full[acc], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + acc * kMaxM;
#pragma unroll 1
      for (int cc = half; cc < chunks; cc += 2) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr + cc * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (n_ok) {
          // lanes = 32 consecutive n of C row m: one 128-byte store (and load, when accumulating) per row per warp
          float* o = cbase + (int64_t)(cc * 32) * p.ldc;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            float val = __uint_as_float(v[r]) * alpha;
            if (accumulate) val += *o;
            *o = val;
            o += p.ldc;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.acc_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
}

// 3-D fp32 tensor (inner, rows, batch); elements outside any dimension are zero-filled.
int encode_3d(CUtensorMap* map, const float* base, int64_t inner, int64_t rows, int64_t batch, int64_t row_pitch,
              int64_t batch_stride, int box_inner, int box_rows, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(EEM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)row_pitch * sizeof(float), (cuuint64_t)batch_stride * sizeof(float)};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EEM_ERR_CUDA, "cuTensorMapEncodeTiled (batched GEMM operand) failed with CUresult %d", (int)r);
  return EEM_OK;
}

bool shape_ok(int batch, int M, int N, int K, int64_t lda, int64_t ldb, int64_t strideA, int64_t strideB, int b_transposed) {
  if (batch <= 0 || M <= 0 || N <= 0 || K <= 0) return false;
  if (M % 32 != 0 || M > kMaxM) return false;                          // whole 32-column TMEM chunks, one accumulator
  if (lda % 4 != 0 || ldb % 4 != 0) return false;                      // TMA: 16-byte row pitch
  if (batch > 1 && (strideA % 4 != 0 || strideB % 4 != 0)) return false;
  if (lda < K || ldb < (b_transposed ? K : N)) return false;
  if ((int64_t)batch * ceil_div(N, kTileN) >= (int64_t)1 << 30) return false;
  return true;
}

}  // namespace
}  // namespace eem

using namespace eem;

extern "C" int eem_batched_gemm_tf32_supported(int batch, int M, int N, int K, int64_t lda, int64_t ldb, int64_t strideA,
                                               int64_t strideB, int b_transposed) {
  return shape_ok(batch, M, N, K, lda, ldb, strideA, strideB, b_transposed) ? 1 : 0;
}

extern "C" int eem_batched_gemm_tf32_multi(const float* const* A, const float* const* B, float* C, int n_seg, int batch, int M,
                                           int N, const int* K, const int64_t* lda, const int64_t* ldb, int64_t ldc,
                                           const int64_t* strideA, const int64_t* strideB, int64_t strideC, int b_transposed,
                                           float alpha, int accumulate, eem_stream_t stream_) {
  EEM_CHECK_ARG(A && B && C && K && lda && ldb && strideA && strideB, "eem_batched_gemm_tf32_multi: NULL pointer");
  EEM_CHECK_ARG(n_seg > 0 && n_seg <= kMaxSeg, "eem_batched_gemm_tf32_multi: n_seg must be in [1,%d]", kMaxSeg);
  EEM_CHECK_ARG(batch > 0 && M > 0 && N > 0, "eem_batched_gemm_tf32_multi: sizes must be > 0");
  EEM_CHECK_ARG(ldc >= N, "eem_batched_gemm_tf32_multi: ldc < N");
  const bool a_mn = b_transposed == 0;   // API B given K x N (n contiguous): the tcgen05 A operand is MN-major

  GemmTcParams p{};
  p.C = C;
  p.M = M; p.N = N;
  p.n_seg = n_seg;
  p.ldc = ldc; p.strideC = strideC;
  p.n_tiles = (int)ceil_div(N, kTileN);
  p.total_tiles = batch * p.n_tiles;
  p.alpha = alpha;
  p.accumulate = accumulate;
  // kind::tf32, fp32 accumulate, A tf32 / B tf32, A major per layout (bit 15: 1 = MN-major), B K-major, N = M rows, M = 128
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | ((uint32_t)(M >> 3) << 17) |
            ((uint32_t)(kTileN >> 4) << 24);
  // K-major SWIZZLE_128B: 8-row groups of 1 KiB; the leading byte offset is not used by swizzled K-major layouts
  desc_fields(16, 1024, 2, &p.b_lo, &p.b_hi);
  if (a_mn) {
    desc_fields(32u * kBK * 4u, 512, 1, &p.a_lo, &p.a_hi);   // as in corr_volume.cu: boxes 4 KiB apart, 4-k groups 512 B apart
    p.a_kstep = 1024;                                        // 8 k rows of 128 B
  } else {
    p.a_lo = p.b_lo;
    p.a_hi = p.b_hi;
    p.a_kstep = kUK * 4;
  }
  for (int sgm = 0; sgm < n_seg; ++sgm) {
    EEM_CHECK_ARG(A[sgm] && B[sgm] && K[sgm] > 0, "eem_batched_gemm_tf32_multi: segment %d: NULL pointer or K <= 0", sgm);
    if (!shape_ok(batch, M, N, K[sgm], lda[sgm], ldb[sgm], strideA[sgm], strideB[sgm], b_transposed))
      return fail(EEM_ERR_UNSUPPORTED,
                  "eem_batched_gemm_tf32: needs M %% 32 == 0, M <= %d and row pitches / batch strides that are multiples of 4 "
                  "elements (got M=%d, lda=%lld, ldb=%lld); use eem_batched_gemm_f32",
                  kMaxM, M, (long long)lda[sgm], (long long)ldb[sgm]);
    EEM_CHECK_ALIGNED(A[sgm], 16);
    EEM_CHECK_ALIGNED(B[sgm], 16);
    p.K[sgm] = K[sgm];
    int rc;
    if (a_mn)
      rc = encode_3d(&p.map_a[sgm], B[sgm], N, K[sgm], batch, ldb[sgm], strideB[sgm], 32, kBK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    else
      rc = encode_3d(&p.map_a[sgm], B[sgm], K[sgm], N, batch, ldb[sgm], strideB[sgm], kBK, kTileN, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != EEM_OK) return rc;
    rc = encode_3d(&p.map_b[sgm], A[sgm], K[sgm], M, batch, lda[sgm], strideA[sgm], kBK, M, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != EEM_OK) return rc;
  }

  const int sms = sm_count();
  if (sms <= 0) return fail(EEM_ERR_CUDA, "eem_batched_gemm_tf32: cannot query SM count");
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  constexpr size_t kSmem = sizeof(GemmTcSmem) + 1024;
  cudaStream_t stream = as_stream(stream_);
  if (a_mn) {
    static DynSmemOptIn optin;
    EEM_CHECK_CUDA(optin.ensure(gemm_tf32_kernel<true>, kSmem));
    gemm_tf32_kernel<true><<<grid, kThreads, kSmem, stream>>>(p);
  } else {
    static DynSmemOptIn optin;
    EEM_CHECK_CUDA(optin.ensure(gemm_tf32_kernel<false>, kSmem));
    gemm_tf32_kernel<false><<<grid, kThreads, kSmem, stream>>>(p);
  }
  EEM_CHECK_LAUNCH("gemm_tf32_kernel");
  return EEM_OK;
}

extern "C" int eem_batched_gemm_tf32(const float* A, const float* B, float* C, int batch, int M, int N, int K, int64_t lda,
                                     int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB, int64_t strideC,
                                     int b_transposed, float alpha, int accumulate, eem_stream_t stream_) {
  EEM_CHECK_ARG(A && B && C, "eem_batched_gemm_tf32: NULL pointer");
  EEM_CHECK_ARG(batch > 0 && M > 0 && N > 0 && K > 0, "eem_batched_gemm_tf32: sizes must be > 0");
  return eem_batched_gemm_tf32_multi(&A, &B, C, 1, batch, M, N, &K, &lda, &ldb, ldc, &strideA, &strideB, strideC, b_transposed,
                                     alpha, accumulate, stream_);
}
