// K6t: the local 9x9 correlation as a BANDED GEMM on tcgen05 (kind::tf32), sm_100a.
//
// Same semantics as local_corr.cu (EEMFlow.py:14-23 / EEMFlow+.py:16-25: SpatialCorrelationSampler(1, 9, 1, 0, 1)
// followed by "/ c" and the fixed index_select), with the channel contraction done by the tensor cores in TF32
// (fp32 accumulate) instead of FFMA.  Opt-in: the products carry 10-bit mantissas, the reference's are fp32.
//
// Formulation.  For a block of 4 image rows x 32 pixels of f1 (M = 128 TMEM lanes) and ONE row of 64 pixels of f2
// starting 8 pixels left of the block (N = 64 TMEM columns),
//     D[(r, j)][n] = sum_c f1[c, y0 + r, x0 + j] * f2[c, y2, x0 - 8 + n]
// holds, on the diagonal band n = j + dx + 8 (dx = -4..4), the nine displacements of vertical offset dy = y2 - (y0 + r).
// (8, not 4: a TMA box whose first byte is not 32-byte aligned is delivered at half rate or less -- measured 33.8 us
// against 18.9 us for the 40 x 48 x 64-channel level -- and x0 - 8 keeps every f2 box sector-aligned.)
// A job is (sample, 32-pixel column, 4-row block k of f1, 4-row block m = k - 1 | k | k + 1 of f2): 4 f2 rows x 64
// columns = ONE 256-column operand (M = 128, N = 256 MMAs), accumulated over C in stages of 32 channels.  Every
// (dy, dx) of every pixel is produced by exactly one job, so jobs are independent and nothing is accumulated in global
// memory; f2 rows / columns outside the image are zero-filled by TMA, which is the reference's zero padding.
//
// Both operands are read "MN-major" straight from the NCHW fp32 maps: a TMA box is [4 rows][32 channels][32 pixels]
// (16 KiB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), i.e. four of the canonical 4-KiB operand groups corr_volume.cu uses.
// Per stage: 1 box of f1 + 2 boxes of f2 = 48 KiB, 4 MMAs of 128 x 256 x 8.
//
// Roles (320 threads, one persistent CTA per SM, a contiguous range of jobs each): warps 0-7 epilogue (warp w -> f1
// row w % 4 of the block = TMEM lane quarter, f2 rows 2 * (w / 4) and 2 * (w / 4) + 1 of the job), warp 8 TMA producer,
// warp 9 MMA issuer.  TMEM holds two 256-column accumulators so the epilogue of job i overlaps the MMAs of job i + 1.
// The epilogue pulls columns 0 .. 47 of a (row r, f2 row s) pair out of TMEM, stages the 40 live ones in shared memory
// (pitch 44: conflict-free 16-byte row writes and conflict-free diagonal reads) and each lane picks its own diagonal:
// lane j reads staged column j + d, so a warp stores 32 consecutive pixels of one output channel per instruction
// (128-byte coalesced).
#include <cstdlib>

#include "tc_common.cuh"

namespace eem {
namespace {
using namespace tc;

constexpr int kND = 9, kMD = 4;
constexpr int kBK = 32;                                   // channels per stage
constexpr int kBoxBytes = 32 * kBK * 4;                   // 4 KiB
constexpr int kABoxes = 4, kBBoxes = 8;                   // f1: 4 rows x 32 px; f2: 4 rows x 64 px
constexpr int kStageBytes = (kABoxes + kBBoxes) * kBoxBytes;
constexpr int kStages = 3;
constexpr int kAccCols = 256;                             // 4 f2 rows x 64 columns
constexpr int kTmemCols = 2 * kAccCols;
constexpr int kEpiWarps = 8;                             // two per TMEM lane quarter: f2 rows {0, 1} and {2, 3} of the job
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kPitch = 44;                                // staging row pitch (floats)
constexpr int kF2Shift = 8;                               // f2 boxes start 8 px (32 B: one sector) left of the f1 block
constexpr int kLive = 40;                                 // columns 4 .. 43 of a 64-column row are the ones the band touches
constexpr int kLd = 48;                                   // columns pulled out of TMEM per row (32 + 16)
// kind::tf32, fp32 accumulate, both operands MN-major, M = 128, N = 256 (the four f2 rows of a job are 8 consecutive
// boxes, i.e. ONE 256-column operand: f1 is read from shared memory once per k-step instead of once per f2 row)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

struct LcTcParams {
  CUtensorMap map1, map2;   // [B][C][H][W] fp32 addressed (x, channel, y, sample); box 32 px x 32 channels x 4 rows
  float* out;
  int B, C, H, W, n_out;
  int ntx, nby;             // 32-pixel columns, 4-row blocks
  int n_jobs;               // B * ntx * nby * 3
  float scale;
  uint32_t desc_lo, desc_hi;
  signed char slot[kND * kND];
  unsigned short row_mask[kND];   // per dy: which dx are selected
  uint32_t debug;                 // EEM_LC_DEBUG bits (timing experiments only): 1 skip MMA, 2 skip global stores, 4 skip TMA loads, 8 skip the band extraction
};

struct __align__(1024) LcTcSmem {
  uint8_t ring[kStages][kStageBytes];
  float stage_out[kEpiWarps][32 * kPitch];
  int slot_s[kND][12];                        // output channel of (dy, dx) or -1; rows padded to three 16-byte loads
  uint64_t full[kStages], empty[kStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int x, int y, int c, int b, uint64_t* bar,
                                            uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(b), "l"(pol)
      : "memory");
}

// columns 0 .. 47 of one f2 row (32 from the left 32-pixel group, 16 from the right one, 128 columns further) of this
// warp's 32 lanes -> 48 registers per lane (the wait is the caller's); the band lives in columns 4 .. 43
__device__ __forceinline__ void tmem_ld48(uint32_t (&v)[kLd], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47])
      : "r"(taddr + 128));
}

// single predicated STG: the unrolled band extraction has no branches
__device__ __forceinline__ void st_stream_if(float* ptr, float v, bool pred) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %2, 0;\n"
      "@q st.global.L1::no_allocate.f32 [%0], %1;\n"
      "}\n" ::"l"(ptr), "f"(v), "r"((int)pred)
      : "memory");
}

// Jobs are numbered ((sample * ntx + column) * nby + k) * 3 + pass; a CTA owns a contiguous range, decoded once and
// then advanced with carries (no division in the per-job loops of the three roles).
struct Job {
  int b, x0, k, pass;   // f1 rows 4k .. 4k+3, f2 rows 4(k + pass - 1) ..
  __device__ __forceinline__ void decode(const LcTcParams& p, int job) {
    pass = job % 3;
    int t = job / 3;
    k = t % p.nby;
    t /= p.nby;
    x0 = (t % p.ntx) * 32;
    b = t / p.ntx;
  }
  __device__ __forceinline__ void next(const LcTcParams& p) {
    if (++pass == 3) {
      pass = 0;
      if (++k == p.nby) {
        k = 0;
        x0 += 32;
        if (x0 >= p.ntx * 32) { x0 = 0; ++b; }
      }
    }
  }
};

__global__ void __launch_bounds__(kThreads, 1)
local_corr_tf32_kernel(const __grid_constant__ LcTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  LcTcSmem& s = *reinterpret_cast<LcTcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = (p.C + kBK - 1) / kBK;
  const int per_cta = (p.n_jobs + (int)gridDim.x - 1) / (int)gridDim.x;
  const int job_begin = min((int)blockIdx.x * per_cta, p.n_jobs), job_end = min(job_begin + per_cta, p.n_jobs);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s.acc_full[i], 1); mbar_init(&s.acc_empty[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kND * 12; i += kThreads) s.slot_s[i / 12][i % 12] = (i % 12 < kND) ? (int)p.slot[(i / 12) * kND + i % 12] : -1;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;

  if (warp == kProducerWarp) {
    // ===== TMA producer (whole warp converged, one elected lane issues) =====
    const uint64_t keep = policy_evict_last();   // every box is re-read by the neighbouring jobs
    int stage = 0;
    uint32_t phase = 0;
    Job j;
    j.decode(p, job_begin);
    for (int job = job_begin; job < job_end; ++job, j.next(p)) {
      const int y1 = 4 * j.k, y2 = 4 * (j.k + j.pass - 1);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&s.empty[stage], phase ^ 1);
        if (p.debug & 4) {
          if (elect_one()) mbar_arrive(&s.full[stage]);
        } else if (elect_one()) {
          // timing experiments: 16 skips the f1 box, 32 the f2 boxes, 64 loads the f2 boxes 128-byte aligned (wrong data)
          const uint32_t bytes = ((p.debug & 16) ? 0u : kABoxes * kBoxBytes) + ((p.debug & 32) ? 0u : kBBoxes * kBoxBytes);
          const int xs = (p.debug & 64) ? 0 : kF2Shift;
          mbar_expect_tx(&s.full[stage], bytes);
          uint8_t* dst = s.ring[stage];
          // one box = 32 px x 32 channels x 4 rows (16 KiB), written as [row][channel][32 px]: four 4-KiB operand groups
          if (!(p.debug & 16)) tma_load_4d(dst, &p.map1, j.x0, kb * kBK, y1, j.b, &s.full[stage], keep);
          if (!(p.debug & 32)) {
            tma_load_4d(dst + kABoxes * kBoxBytes, &p.map2, j.x0 - xs, kb * kBK, y2, j.b, &s.full[stage], keep);
            tma_load_4d(dst + (kABoxes + 4) * kBoxBytes, &p.map2, j.x0 - xs + 32, kb * kBK, y2, j.b, &s.full[stage], keep);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =====
    int stage = 0;
    uint32_t phase = 0, n = 0;
    for (int job = job_begin; job < job_end; ++job, ++n) {
      const uint32_t acc = n & 1;
      mbar_wait(&s.acc_empty[acc], ((n >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + acc * kAccCols;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&s.full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(s.ring[stage]);
        const uint32_t b_addr = a_addr + kABoxes * kBoxBytes;
        const bool leader = elect_one();
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < kBK / 8; ++ks) {
            if (p.debug & 1) break;
            tc_mma_tf32(d_tmem, make_desc(a_addr + ks * 1024, p.desc_lo, p.desc_hi), make_desc(b_addr + ks * 1024, p.desc_lo, p.desc_hi),
                        kIdesc, (kb | ks) != 0);
          }
          tc_commit(&s.empty[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(&s.acc_full[acc]);
      __syncwarp();
    }
  } else {
    // ===== epilogue: warp w -> f1 row r = w % 4 of the block (TMEM lanes 32r .. 32r+31) and the f2 rows 2*(w/4), 2*(w/4)+1;
    // lane = pixel.  Both rows are pulled out of TMEM first and the accumulator is handed back to the MMA warp before
    // the band is extracted, so the tensor pipe never waits for the shared-memory / global part.
    const uint32_t st_row = smem_u32(s.stage_out[warp]) + lane * kPitch * 4, slot_base = smem_u32(&s.slot_s[0][0]);
    const int r = warp & 3, half = warp >> 2;
    const int64_t plane = (int64_t)p.H * p.W;
    uint32_t n = 0;
    Job j;
    j.decode(p, job_begin);
    for (int job = job_begin; job < job_end; ++job, ++n, j.next(p)) {
      const uint32_t acc = n & 1;
      const int dy0 = 4 * (j.pass - 1) + 2 * half - r, dy1 = dy0 + 1;          // warp-uniform
      const unsigned mask0 = (dy0 >= -kMD && dy0 <= kMD && !(p.debug & 8)) ? p.row_mask[dy0 + kMD] : 0u;
      const unsigned mask1 = (dy1 >= -kMD && dy1 <= kMD && !(p.debug & 8)) ? p.row_mask[dy1 + kMD] : 0u;
      mbar_wait(&s.acc_full[acc], (n >> 1) & 1);
      tc_fence_after();
      // f2 row q of the job, pixel column n (0..63): TMEM column (n / 32) * 128 + q * 32 + n % 32
      const uint32_t taddr = tmem + ((uint32_t)(r * 32) << 16) + acc * kAccCols + half * 64;
      uint32_t v0[kLd], v1[kLd];
      if (mask0) tmem_ld48(v0, taddr);
      if (mask1) tmem_ld48(v1, taddr + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.acc_empty[acc]);

      const int y = 4 * j.k + r, x = j.x0 + lane;
      const bool px_ok = y < p.H && x < p.W && !(p.debug & 2);
      float* obase = p.out + (int64_t)j.b * p.n_out * plane + (int64_t)y * p.W + x;
      auto extract = [&](const uint32_t (&v)[kLd], int dy) {
        // staging column c = TMEM column c + 4: pixel j, displacement column d sits at TMEM column j + d + 4
#pragma unroll
        for (int q = 0; q < kLive / 4; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + 16 * q), "r"(v[4 * q + 4]), "r"(v[4 * q + 5]),
                       "r"(v[4 * q + 6]), "r"(v[4 * q + 7]) : "memory");
        __syncwarp();
        int sl[12];
#pragma unroll
        for (int q = 0; q < 3; ++q)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(sl[4 * q]), "=r"(sl[4 * q + 1]), "=r"(sl[4 * q + 2]), "=r"(sl[4 * q + 3])
                       : "r"(slot_base + (dy + kMD) * 48 + 16 * q));
        float val[kND];
#pragma unroll
        for (int d = 0; d < kND; ++d) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val[d]) : "r"(st_row + 4 * (lane + d)));   // column j + d of row j
#pragma unroll
        for (int d = 0; d < kND; ++d) st_stream_if(obase + (int64_t)sl[d] * plane, val[d] * p.scale, px_ok && sl[d] >= 0);
        __syncwarp();
      };
      if (mask0) extract(v0, dy0);
      if (mask1) extract(v1, dy1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
}

// 4-D fp32 tensor [B][C][H][W]; box = 32 px x kBK channels x 4 rows x 1 sample.  Out-of-range elements (negative
// coordinates, pixels beyond W / H, channels beyond C) are zero-filled.
int encode_nchw_map(CUtensorMap* map, const float* base, int B, int C, int H, int W) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(EEM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // dimension order (x, channel, y, sample): a box of 4 image rows lands in shared memory as [row][channel][32 px]
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)H * W * 4, (cuuint64_t)W * 4, (cuuint64_t)C * H * W * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)kBK, 4, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EEM_ERR_CUDA, "cuTensorMapEncodeTiled (NCHW map) failed with CUresult %d", (int)r);
  return EEM_OK;
}

}  // namespace
}  // namespace eem

using namespace eem;

extern "C" int eem_local_corr_tf32_supported(int B, int C, int H, int W, int max_disp) {
  return (max_disp == kMD && B > 0 && C > 0 && H > 0 && W > 0 && W % 4 == 0) ? 1 : 0;
}

extern "C" int eem_local_corr_tf32(const float* f1, const float* f2, int B, int C, int H, int W, int max_disp,
                                   const int* index, int n_out, float scale, float* out, eem_stream_t stream_) {
  EEM_CHECK_ARG(f1 && f2 && out, "eem_local_corr_tf32: NULL pointer");
  EEM_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "eem_local_corr_tf32: sizes must be > 0");
  if (max_disp != kMD)
    return fail(EEM_ERR_UNSUPPORTED, "eem_local_corr_tf32: only max_disp == %d (patch_size 9) is implemented, got %d", kMD, max_disp);
  if (W % 4 != 0 || ((reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2)) % 16) != 0)
    return fail(EEM_ERR_UNSUPPORTED, "eem_local_corr_tf32: needs W %% 4 == 0 and 16-byte aligned maps (TMA row pitch); W = %d", W);
  LcTcParams p{};
  p.out = out;
  p.B = B; p.C = C; p.H = H; p.W = W; p.n_out = n_out; p.scale = scale;
  if (index == nullptr) {
    EEM_CHECK_ARG(n_out == kND * kND, "eem_local_corr_tf32: n_out must be %d without an index list (got %d)", kND * kND, n_out);
    for (int ch = 0; ch < kND * kND; ++ch) p.slot[ch] = (signed char)ch;
  } else {
    EEM_CHECK_ARG(n_out > 0 && n_out <= kND * kND, "eem_local_corr_tf32: n_out must be in [1,%d]", kND * kND);
    for (int ch = 0; ch < kND * kND; ++ch) p.slot[ch] = -1;
    for (int k = 0; k < n_out; ++k) {
      EEM_CHECK_ARG(index[k] >= 0 && index[k] < kND * kND, "eem_local_corr_tf32: index[%d]=%d out of range", k, index[k]);
      if (p.slot[index[k]] != -1)
        return fail(EEM_ERR_UNSUPPORTED, "eem_local_corr_tf32: repeated channel %d in index list", index[k]);
      p.slot[index[k]] = (signed char)k;
    }
  }
  for (int r = 0; r < kND; ++r) {
    p.row_mask[r] = 0;
    for (int d = 0; d < kND; ++d)
      if (p.slot[r * kND + d] >= 0) p.row_mask[r] |= (unsigned short)(1u << d);
  }
  p.ntx = ceil_div(W, 32);
  p.nby = ceil_div(H, 4);
  const int64_t n_jobs = (int64_t)B * p.ntx * p.nby * 3;
  EEM_CHECK_ARG(n_jobs < (int64_t)1 << 30, "eem_local_corr_tf32: too many tiles (%lld)", (long long)n_jobs);
  p.n_jobs = (int)n_jobs;
  if (const char* dbg = getenv("EEM_LC_DEBUG")) p.debug = (uint32_t)atoi(dbg);
  desc_fields(kBoxBytes, 512, 1, &p.desc_lo, &p.desc_hi);
  int rc = encode_nchw_map(&p.map1, f1, B, C, H, W);
  if (rc != EEM_OK) return rc;
  rc = encode_nchw_map(&p.map2, f2, B, C, H, W);
  if (rc != EEM_OK) return rc;

  constexpr size_t kSmem = sizeof(LcTcSmem) + 1024;
  static DynSmemOptIn optin;
  const cudaError_t attr_err = optin.ensure(local_corr_tf32_kernel, kSmem);
  if (attr_err != cudaSuccess) return fail(EEM_ERR_CUDA, "local_corr_tf32_kernel attribute: %s", cudaGetErrorString(attr_err));
  const int grid = (int)std::min<int64_t>(n_jobs, sm_count());
  local_corr_tf32_kernel<<<grid, kThreads, kSmem, as_stream(stream_)>>>(p);
  EEM_CHECK_LAUNCH("local_corr_tf32_kernel");
  return EEM_OK;
}
