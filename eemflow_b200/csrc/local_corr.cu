// K6 local (2*md+1)^2 correlation with fused scale and fused output-channel selection, sm_100a.
//
// Reference semantics (EEMFlow.py:14-23 / EEMFlow+.py:16-25 around
// spatial_correlation_sampler.SpatialCorrelationSampler(kernel_size=1, patch_size=9, stride=1,
// padding=0, dilation=1); in-tree restatement model/IRRPWC/pwc_modules.py:42-63):
//   out[b, (dy+md)*(2md+1) + (dx+md), y, x] = sum_c f1[b,c,y,x] * f2[b,c,y+dy,x+dx]   (zero outside f2)
// followed by "/ c" and torch.index_select over a fixed channel list (49 or 53 of 81).
//
// Tiling: a CTA owns a 4 x 32 pixel tile of one sample.  Warp = vertical displacement dy (9
// warps), lane = a quad of 4 horizontally adjacent pixels; per channel a thread reads one f1
// float4 and three f2 float4 (12 consecutive columns) from shared memory and issues 36 FMAs, so
// the loop is FMA-bound rather than LDS-bound.  Channels stream through shared memory in chunks
// of 8 with the f2 halo (+-4) loaded once per chunk.
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "common.cuh"

namespace eem {
namespace {

constexpr int MD = 4;             // max displacement implemented (the reference only uses 4)
constexpr int ND = 2 * MD + 1;    // 9
constexpr int TW = 32, TH = 4;    // pixel tile
constexpr int CC = 8;             // channels per shared-memory chunk
constexpr int F2W = TW + 2 * MD;  // 40 (multiple of 4 -> float4-aligned rows)
constexpr int F2H = TH + 2 * MD;  // 12
constexpr int kThreads = ND * 32;

struct LocalCorrParams {
  const float* f1;
  const float* f2;
  float* out;
  int B, C, H, W, n_out;
  float scale;
  signed char slot[ND * ND];  // output channel of displacement channel ch, or -1 when not selected
  signed char cls[ND];        // per dy row: which of the kernel's compile-time dx masks applies (5 = all nine)
};

__global__ void __launch_bounds__(kThreads)
local_corr_generic_kernel(const __grid_constant__ LocalCorrParams p) {
  __shared__ __align__(16) float s1[CC][TH][TW];
  __shared__ __align__(16) float s2[CC][F2H][F2W];

  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qy = lane >> 3, qx = (lane & 7) * 4;  // quad position inside the tile
  const int dy = warp - MD;
  const int64_t plane = (int64_t)p.H * p.W;
  const float* f1 = p.f1 + (int64_t)b * p.C * plane;
  const float* f2 = p.f2 + (int64_t)b * p.C * plane;

  float acc[4][ND];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[i][d] = 0.f;

  for (int c0 = 0; c0 < p.C; c0 += CC) {
    const int nc = min(CC, p.C - c0);
    __syncthreads();
    // fixed trip counts, loads first and stores after, so the ~17 global loads of a thread overlap
    constexpr int N1 = (CC * TH * TW + kThreads - 1) / kThreads, N2 = (CC * F2H * F2W + kThreads - 1) / kThreads;
    float v1[N1], v2[N2];
#pragma unroll
    for (int k = 0; k < N1; ++k) {
      const int t = threadIdx.x + k * kThreads;
      const int c = t / (TH * TW), r = (t / TW) % TH, x = t % TW;
      const int gy = y0 + r, gx = x0 + x;
      v1[k] = (t < CC * TH * TW && c < nc && gy < p.H && gx < p.W) ? __ldg(f1 + (int64_t)(c0 + c) * plane + (int64_t)gy * p.W + gx) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < N2; ++k) {
      const int t = threadIdx.x + k * kThreads;
      const int c = t / (F2H * F2W), r = (t / F2W) % F2H, x = t % F2W;
      const int gy = y0 + r - MD, gx = x0 + x - MD;
      v2[k] = (t < CC * F2H * F2W && c < nc && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
                  ? __ldg(f2 + (int64_t)(c0 + c) * plane + (int64_t)gy * p.W + gx) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < N1; ++k) {
      const int t = threadIdx.x + k * kThreads;
      if (t < CC * TH * TW) (&s1[0][0][0])[t] = v1[k];
    }
#pragma unroll
    for (int k = 0; k < N2; ++k) {
      const int t = threadIdx.x + k * kThreads;
      if (t < CC * F2H * F2W) (&s2[0][0][0])[t] = v2[k];
    }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < CC; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(&s1[c][qy][qx]);
      const float* row = &s2[c][qy + MD + dy][qx];
      const float4 r0 = *reinterpret_cast<const float4*>(row);
      const float4 r1 = *reinterpret_cast<const float4*>(row + 4);
      const float4 r2 = *reinterpret_cast<const float4*>(row + 8);
      const float w[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[i][d] = fmaf(av[i], w[i + d], acc[i][d]);
    }
  }

  const int gy = y0 + qy, gx = x0 + qx;
  if (gy >= p.H || gx >= p.W) return;
  const bool vec = (p.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const int k = p.slot[(dy + MD) * ND + d];
    if (k < 0) continue;
    float* o = p.out + ((int64_t)b * p.n_out + k) * plane + (int64_t)gy * p.W + gx;
    const float4 v = make_float4(acc[0][d] * p.scale, acc[1][d] * p.scale, acc[2][d] * p.scale, acc[3][d] * p.scale);
    if (vec) {
      st_stream4(o, v);
    } else {
      o[0] = v.x;
      if (gx + 1 < p.W) o[1] = v.y;
      if (gx + 2 < p.W) o[2] = v.z;
      if (gx + 3 < p.W) o[3] = v.w;
    }
  }
}


// ---- fast path (W % 4 == 0, 16-byte aligned inputs) -----------------------------------------
// Tile 8 x 32 pixels.  Warp = dy (9 warps), lane = (row 0..7, octet 0..3): a thread owns 8
// horizontally adjacent pixels x 9 dx = 72 accumulators and per channel reads 2 float4 of f1 and
// 4 float4 of f2 (16 consecutive columns), i.e. 6 LDS.128 per 72 FMAs.  Channel chunks of 8 are
// staged with cp.async (16-byte copies, zero-filled outside the image) into a double buffer so the
// loads of chunk k+1 overlap the FMAs of chunk k.  Row pitches (36 / 44 floats) are chosen so the
// 8 lanes of an LDS.128 phase hit distinct banks.
constexpr int VW = 32, VH = 8;
constexpr int VC = 8;                       // channels per stage
constexpr int S1P = 36;                     // f1 smem row pitch (floats)
constexpr int S2W = VW + 2 * MD;            // 40 valid columns
constexpr int S2P = 44;                     // f2 smem row pitch (floats)
constexpr int S2H = VH + 2 * MD;            // 16
constexpr int kS1Floats = VC * VH * S1P;    // 2304
constexpr int kS2Floats = VC * S2H * S2P;   // 5632
constexpr int kStageFloats = kS1Floats + kS2Floats;
constexpr int kVec1 = VC * VH * (VW / 4);   // 512 16-byte copies of f1 per stage
constexpr int kVec2 = VC * S2H * (S2W / 4); // 1280 of f2

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;  // src-size 0: nothing is read, the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}

// MA..ME: the distinct per-row dx masks (bit d = displacement column d is kept) of a known channel-selection
// pattern.  The kernel is FFMA-issue-bound (72 FMAs per channel per thread for all nine columns) and EEMFlow
// keeps 53 or 49 of the 81 displacements, so every warp (= dy row) runs a copy of the WHOLE channel loop that
// is specialised for its row's mask: the FMAs and accumulators of dropped displacements do not exist in that
// copy (at most 56 FMAs per channel).  Rows whose mask is not among the five (p.cls == 5) compute all nine
// columns, so any selection stays correct.  All copies execute the same number of block barriers.
template <unsigned MA, unsigned MB, unsigned MC, unsigned MD_, unsigned ME>
__global__ void __launch_bounds__(kThreads, 2)
local_corr_vec_kernel(const __grid_constant__ LocalCorrParams p) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * VW, y0 = blockIdx.y * VH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = lane >> 2, oct = (lane & 3) * 8;
  const int dy = warp - MD;
  const int64_t plane = (int64_t)p.H * p.W;
  const float* f1 = p.f1 + (int64_t)b * p.C * plane;
  const float* f2 = p.f2 + (int64_t)b * p.C * plane;
  const int n_chunks = (p.C + VC - 1) / VC;

  // Per-thread copy descriptors, computed ONCE: which 16-byte vectors of a stage this thread moves
  // (shared-memory offset, global element offset relative to the chunk's first channel, in-image
  // flag).  Only the channel base changes from chunk to chunk, so issuing a stage is ~3
  // instructions per vector instead of re-deriving (channel, row, column) with div/mod every time.
  constexpr int kSlots = (kVec1 + kVec2 + kThreads - 1) / kThreads;   // 7
  int s_off[kSlots], g_off[kSlots];
  unsigned ch_of = 0, in_img = 0, is_f2 = 0;   // 4-bit channel index per slot / 1-bit flags per slot
#pragma unroll
  for (int k = 0; k < kSlots; ++k) {
    const int v = threadIdx.x + k * kThreads;
    int c = 0, so = 0, go = 0;
    bool ok = false, f2v = false;
    if (v < kVec1) {
      c = v >> 6;
      const int r = (v >> 3) & 7, q = v & 7;
      const int gy = y0 + r, gx = x0 + 4 * q;
      ok = gy < p.H && gx < p.W;
      so = (c * VH + r) * S1P + 4 * q;
      go = c * (int)plane + gy * p.W + gx;
    } else if (v < kVec1 + kVec2) {
      const int w = v - kVec1;
      c = w / (S2H * (S2W / 4));
      const int rem = w - c * (S2H * (S2W / 4));
      const int r = rem / (S2W / 4), q = rem - r * (S2W / 4);
      const int gy = y0 + r - MD, gx = x0 - MD + 4 * q;
      ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      so = kS1Floats + (c * S2H + r) * S2P + 4 * q;
      go = c * (int)plane + gy * p.W + gx;
      f2v = true;
    } else {
      so = -1;
    }
    s_off[k] = so;
    g_off[k] = ok ? go : 0;
    ch_of |= (unsigned)c << (4 * k);
    in_img |= (unsigned)ok << k;
    is_f2 |= (unsigned)f2v << k;
  }

  auto issue = [&](int chunk, int buf) {
    float* stage = smem + buf * kStageFloats;
    const int c0 = chunk * VC;
    const int64_t cbase = (int64_t)c0 * plane;
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
      if (s_off[k] < 0) continue;
      const int c = (ch_of >> (4 * k)) & 15;
      const bool ok = ((in_img >> k) & 1) && (c0 + c < p.C);
      const float* src = (((is_f2 >> k) & 1) ? f2 : f1) + (ok ? cbase + g_off[k] : 0);
      cp_async16(stage + s_off[k], src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  auto run = [&](auto mask_tag) {
    constexpr unsigned M = decltype(mask_tag)::value;
    float acc[8][ND];
  #pragma unroll
    for (int i = 0; i < 8; ++i)
  #pragma unroll
      for (int d = 0; d < ND; ++d) acc[i][d] = 0.f;

    issue(0, 0);
    for (int k = 0; k < n_chunks; ++k) {
      if (k + 1 < n_chunks) {
        issue(k + 1, (k + 1) & 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      const float* s1 = smem + (k & 1) * kStageFloats;
      const float* s2 = s1 + kS1Floats;
  #pragma unroll 1
      for (int c = 0; c < VC; ++c) {
        const float4* a4 = reinterpret_cast<const float4*>(s1 + (c * VH + row) * S1P + oct);
        const float4* w4 = reinterpret_cast<const float4*>(s2 + (c * S2H + row + MD + dy) * S2P + oct);
        const float4 a0 = a4[0], a1 = a4[1];
        const float4 w0 = w4[0], w1 = w4[1], w2 = w4[2], w3 = w4[3];
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
  #pragma unroll
        for (int i = 0; i < 8; ++i)
  #pragma unroll
          for (int d = 0; d < ND; ++d)
            if ((M >> d) & 1u) acc[i][d] = fmaf(av[i], wv[i + d], acc[i][d]);
      }
      __syncthreads();
    }

    const int gy = y0 + row, gx = x0 + oct;
    if (gy >= p.H || gx >= p.W) return;
    const bool second = gx + 4 < p.W;
  #pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (!((M >> d) & 1u)) continue;
      const int slot = p.slot[(dy + MD) * ND + d];
      if (slot < 0) continue;
      float* o = p.out + ((int64_t)b * p.n_out + slot) * plane + (int64_t)gy * p.W + gx;
      st_stream4(o, make_float4(acc[0][d] * p.scale, acc[1][d] * p.scale, acc[2][d] * p.scale, acc[3][d] * p.scale));
      if (second)
        st_stream4(o + 4, make_float4(acc[4][d] * p.scale, acc[5][d] * p.scale, acc[6][d] * p.scale, acc[7][d] * p.scale));
    }
  };
  switch (p.cls[warp]) {      // warp-uniform: one specialised copy of the loop per row class
    case 0: run(std::integral_constant<unsigned, MA>{}); break;
    case 1: run(std::integral_constant<unsigned, MB>{}); break;
    case 2: run(std::integral_constant<unsigned, MC>{}); break;
    case 3: run(std::integral_constant<unsigned, MD_>{}); break;
    case 4: run(std::integral_constant<unsigned, ME>{}); break;
    default: run(std::integral_constant<unsigned, 0x1FFu>{}); break;
  }
}

// ---- small maps (the coarsest pyramid levels: 5x6, 10x12, 12x20) -------------------------------------------------
// With a few hundred pixels per sample the tiled kernels above run 8 strictly serial load -> barrier -> FMA rounds per
// CTA on a fraction of the SMs (16-18 us for < 1 MB of data).  Here a CTA owns (sample, dy): it pulls ALL channels of
// f1 and of the f2 rows y + dy (zero rows outside the map, zero halo columns) into shared memory with every copy in
// flight at once, then thread = (pixel, channel quarter) accumulates the nine dx of its pixel over its channels and
// the quarters are summed in a fixed order through shared memory: one memory round trip instead of eight.
constexpr int kSmallMaxPix = 256;
constexpr int kSmallGroups = 4;

struct SmallPlan {
  bool ok;
  int pixp, threads;     // pixels rounded up to a warp multiple; pixp * kSmallGroups threads
  size_t smem;
};

inline SmallPlan small_plan(int C, int H, int W) {
  SmallPlan s{false, 0, 0, 0};
  const int hw = H * W;
  if (hw > kSmallMaxPix) return s;
  s.pixp = (hw + 31) / 32 * 32;
  s.threads = s.pixp * kSmallGroups;
  // f1 [C][hw] | f2 [C][H][W + 2 MD] | partial sums [groups - 1][ND][pixp]
  s.smem = ((size_t)C * hw + (size_t)C * H * (W + 2 * MD) + (size_t)(kSmallGroups - 1) * ND * s.pixp) * sizeof(float);
  s.ok = s.smem <= 160 * 1024;
  return s;
}

__global__ void __launch_bounds__(kSmallMaxPix * kSmallGroups)
local_corr_small_kernel(const __grid_constant__ LocalCorrParams p, int pixp) {
  extern __shared__ __align__(16) float smem[];
  const int dyi = blockIdx.x, dy = dyi - MD, b = blockIdx.y;
  unsigned row_mask = 0;
#pragma unroll
  for (int d = 0; d < ND; ++d) row_mask |= (unsigned)(p.slot[dyi * ND + d] >= 0) << d;
  if (row_mask == 0) return;                       // no selected displacement in this row (block-uniform)
  const int H = p.H, W = p.W, C = p.C, hw = H * W, W2 = W + 2 * MD;
  float* s1 = smem;
  float* s2 = s1 + C * hw;
  float* part = s2 + C * H * W2;
  const float* f1 = p.f1 + (int64_t)b * C * hw;
  const float* f2 = p.f2 + (int64_t)b * C * hw;
  const int nthr = blockDim.x;

  for (int i = threadIdx.x; i < C * hw; i += nthr) {
    const uint32_t d1 = (uint32_t)__cvta_generic_to_shared(s1 + i);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d1), "l"(f1 + i) : "memory");
    const int c = i / hw, r = i - c * hw, y = r / W, x = r - y * W;
    const int y2 = y + dy;
    const bool ok = y2 >= 0 && y2 < H;
    const uint32_t d2 = (uint32_t)__cvta_generic_to_shared(s2 + (c * H + y) * W2 + MD + x);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d2), "l"(f2 + (ok ? i + dy * W : 0)), "r"(ok ? 4 : 0) : "memory");
  }
  for (int i = threadIdx.x; i < C * H * 2 * MD; i += nthr) {      // halo columns (disjoint from the copies above)
    const int row = i / (2 * MD), k = i - row * (2 * MD);
    s2[row * W2 + (k < MD ? k : W + k)] = 0.f;
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int pix = threadIdx.x % pixp, g = threadIdx.x / pixp;
  const bool live = pix < hw;
  float acc[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) acc[d] = 0.f;
  if (live) {
    const int y = pix / W, x = pix - y * W;
    const int cpg = (C + kSmallGroups - 1) / kSmallGroups;
    const int c_end = min(C, (g + 1) * cpg);
    const float* a = s1 + pix;
    const float* w = s2 + y * W2 + x;
#pragma unroll 4
    for (int c = g * cpg; c < c_end; ++c) {
      const float av = a[c * hw];
      const float* row = w + c * H * W2;
#pragma unroll
      for (int d = 0; d < ND; ++d) acc[d] = fmaf(av, row[d], acc[d]);
    }
  }
  if (g > 0) {
#pragma unroll
    for (int d = 0; d < ND; ++d) part[((g - 1) * ND + d) * pixp + pix] = acc[d];
  }
  __syncthreads();
  if (g != 0 || !live) return;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const int slot = p.slot[dyi * ND + d];
    if (slot < 0) continue;
    float v = acc[d];
#pragma unroll
    for (int q = 0; q < kSmallGroups - 1; ++q) v += part[(q * ND + d) * pixp + pix];
    p.out[((int64_t)b * p.n_out + slot) * hw + pix] = v * p.scale;
  }
}

int launch_small(const LocalCorrParams& p, const SmallPlan& sp, cudaStream_t stream) {
  static DynSmemOptIn optin;
  if (sp.smem > 48 * 1024) {
    const cudaError_t attr_err = optin.ensure(local_corr_small_kernel, sp.smem);
    if (attr_err != cudaSuccess) return fail(EEM_ERR_CUDA, "local_corr_small_kernel attribute: %s", cudaGetErrorString(attr_err));
  }
  local_corr_small_kernel<<<dim3(ND, (unsigned)p.B), sp.threads, sp.smem, stream>>>(p, sp.pixp);
  EEM_CHECK_LAUNCH("local_corr_small_kernel");
  return EEM_OK;
}

template <unsigned MA, unsigned MB, unsigned MC, unsigned MD_, unsigned ME>
int launch_vec(const LocalCorrParams& p, int B, int H, int W, cudaStream_t stream) {
  const size_t smem = 2 * (size_t)kStageFloats * sizeof(float);
  static DynSmemOptIn optin;
  const cudaError_t attr_err = optin.ensure(local_corr_vec_kernel<MA, MB, MC, MD_, ME>, smem);
  if (attr_err != cudaSuccess) return fail(EEM_ERR_CUDA, "local_corr_vec_kernel attribute: %s", cudaGetErrorString(attr_err));
  dim3 grid((unsigned)ceil_div(W, VW), (unsigned)ceil_div(H, VH), (unsigned)B);
  local_corr_vec_kernel<MA, MB, MC, MD_, ME><<<grid, kThreads, smem, stream>>>(p);
  EEM_CHECK_LAUNCH("local_corr_vec_kernel");
  return EEM_OK;
}

}  // namespace
}  // namespace eem

using namespace eem;

extern "C" int eem_local_corr(const float* f1, const float* f2, int B, int C, int H, int W,
                              int max_disp, const int* index, int n_out, float scale, float* out,
                              eem_stream_t stream_) {
  EEM_CHECK_ARG(f1 && f2 && out, "eem_local_corr: NULL pointer");
  EEM_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "eem_local_corr: sizes must be > 0");
  EEM_CHECK_ARG(B <= 65535, "eem_local_corr: batch > 65535 not supported in one call");
  if (max_disp != MD)
    return fail(EEM_ERR_UNSUPPORTED, "eem_local_corr: only max_disp == %d (patch_size 9) is implemented, got %d", MD, max_disp);
  LocalCorrParams p{};
  p.f1 = f1; p.f2 = f2; p.out = out;
  p.B = B; p.C = C; p.H = H; p.W = W; p.n_out = n_out; p.scale = scale;
  if (index == nullptr) {
    EEM_CHECK_ARG(n_out == ND * ND, "eem_local_corr: n_out must be %d without an index list (got %d)", ND * ND, n_out);
    for (int ch = 0; ch < ND * ND; ++ch) p.slot[ch] = (signed char)ch;
  } else {
    EEM_CHECK_ARG(n_out > 0 && n_out <= ND * ND, "eem_local_corr: n_out must be in [1,%d]", ND * ND);
    for (int ch = 0; ch < ND * ND; ++ch) p.slot[ch] = -1;
    for (int k = 0; k < n_out; ++k) {
      EEM_CHECK_ARG(index[k] >= 0 && index[k] < ND * ND, "eem_local_corr: index[%d]=%d out of range", k, index[k]);
      // index_select allows repeats; the fused path writes each displacement once.
      if (p.slot[index[k]] != -1)
        return fail(EEM_ERR_UNSUPPORTED, "eem_local_corr: repeated channel %d in index list", index[k]);
      p.slot[index[k]] = (signed char)k;
    }
  }
  {
    // small maps: whole-map-per-CTA kernel (EEM_LC_SMALL=0 switches it off for comparisons)
    static const bool small_on = [] {
      const char* v = getenv("EEM_LC_SMALL");
      return !(v != nullptr && atoi(v) == 0);
    }();
    const SmallPlan sp = small_plan(C, H, W);
    if (small_on && sp.ok) return launch_small(p, sp, as_stream(stream_));
  }
  const bool vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2) |
                                        reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  if (vec_ok) {
    // per-row dx masks of this selection, matched against the two precompiled patterns (EEMFlow_cdc's 53 and
    // EEMFlow's 49 channels); rows that match no compile-time mask fall back to all nine columns
    unsigned row_mask[ND];
    for (int r = 0; r < ND; ++r) {
      row_mask[r] = 0;
      for (int d = 0; d < ND; ++d)
        if (p.slot[r * ND + d] >= 0) row_mask[r] |= 1u << d;
    }
    auto classify = [&](const unsigned (&set)[5]) {
      int hits = 0;
      for (int r = 0; r < ND; ++r) {
        p.cls[r] = 5;
        for (int k = 0; k < 5; ++k)
          if (row_mask[r] == set[k]) { p.cls[r] = (signed char)k; ++hits; break; }
      }
      return hits;
    };
    static const unsigned kCdc[5] = {0x155u, 0x0AAu, 0x17Du, 0x0FEu, 0x1FFu};
    static const unsigned kEem[5] = {0x0AAu, 0x155u, 0x0BAu, 0x17Du, 0x0FEu};
    const int hits_cdc = classify(kCdc);
    const bool use_cdc = hits_cdc >= classify(kEem);
    if (use_cdc) classify(kCdc);
    return use_cdc ? launch_vec<0x155u, 0x0AAu, 0x17Du, 0x0FEu, 0x1FFu>(p, B, H, W, as_stream(stream_))
                   : launch_vec<0x0AAu, 0x155u, 0x0BAu, 0x17Du, 0x0FEu>(p, B, H, W, as_stream(stream_));
  }
  dim3 grid((unsigned)ceil_div(W, TW), (unsigned)ceil_div(H, TH), (unsigned)B);
  local_corr_generic_kernel<<<grid, kThreads, 0, as_stream(stream_)>>>(p);
  EEM_CHECK_LAUNCH("local_corr_generic_kernel");
  return EEM_OK;
}
