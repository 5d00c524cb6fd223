// K5 multi-level correlation window lookup + K4 2x2 average pooling, sm_100a.
//
// Reference semantics (model/corr.py:29-50, model/model_utils.py:7-15):
//   for level l: centre = coords / 2^l; window offsets d in [-r, r]^2 (unit spacing);
//   out[b, l*(2r+1)^2 + a*(2r+1) + c, y, x] = bilinear(level_l[b*P + y*W + x], (cx + a - r, cy + c - r))
//   where the first window index moves x and the second moves y (RAFT's transposed meshgrid),
//   sampling goes through  g = 2*p/(S-1) - 1  and grid_sample(align_corners=True, zeros padding).
//
// A CTA owns 32 consecutive positions of one sample.  It gathers each position's (2r+2)^2 tap
// window of EVERY level into shared memory in one phase (every 32-byte sector of the volume is
// fetched once per position, all loads in flight together), then warp `a` / lane `position`
// produces the 2r+1 outputs of window column `a`, so every store instruction writes 32
// consecutive positions of one output channel (128 B).
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace eem {
namespace {

constexpr int kMaxLevels = 8;
constexpr int kPosPerBlock = 32;

struct LookupParams {
  const float* level[kMaxLevels];
  int h[kMaxLevels], w[kMaxLevels];
  int B, H, W, L;
  const float* coords;
  float* out;
};

// The reference's coordinate round trip: pixel -> [-1,1] (bilinear_sampler) -> pixel (grid_sample,
// align_corners=True), in fp32 with the same operation order.
__device__ __forceinline__ float roundtrip(float p, int size) {
  const float s1 = (float)(size - 1);
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, p), s1), 1.0f);   // 2*x/(S-1) - 1
  return __fmul_rn(__fadd_rn(g, 1.0f), s1 * 0.5f);                      // ATen CPU: (g+1) * ((S-1)/2)
}

// Shared-memory layout (dynamic): per level l
//   taps[l][pos][kStride]   (2r+2)^2 window taps of position pos, zero outside the map
//   fx[l][pos][K], fy[l][pos][K]   bilinear fractions per window column / row
//   org[l][pos][2]          integer window origin
// Phases: (1) geometry for every (position, level); (2) ONE gather phase that issues every tap
// load of every level before anything is consumed -- thread = (position, window column), walking
// down the rows, so consecutive lanes read consecutive addresses of a window row; (3) interpolation
// with warp = window column, lane = position, so each store instruction writes 32 consecutive
// positions of one output channel (128 B).  Two block barriers in total.
template <int R, int PB>
struct LookupSmem {
  static constexpr int K = 2 * R + 1, T = K + 1, TT = T * T;
  // tap stride per position: == T+1 (mod 32) so the T lanes of consecutive positions written by one
  // cp.async instruction land in disjoint banks, and odd so lane = position reads are conflict-free
  static constexpr int kStride = TT + ((T + 1 - TT % 32 + 32) % 32);
  static constexpr int kColsPerTask = 3, kGroups = (K + kColsPerTask - 1) / kColsPerTask;
  // threads: >= PB*T for the gather phase, and enough (32/PB)-position slots that a 4-level pyramid is one
  // interpolation task per slot
  static constexpr int kSlotsPerWarp = 32 / PB;
  static constexpr int kGatherWarps = (PB * T + 31) / 32;
  static constexpr int kTaskWarps = (4 * kGroups + kSlotsPerWarp - 1) / kSlotsPerWarp;
  static constexpr int kThreads = 32 * (kGatherWarps > kTaskWarps ? kGatherWarps : kTaskWarps);
  static constexpr int kPerLevelFloats = PB * kStride + 2 * PB * K;
  static constexpr int kPerLevelBytes = kPerLevelFloats * 4 + PB * 2 * 4;
};

template <int R, int PB>
__global__ void __launch_bounds__(LookupSmem<R, PB>::kThreads)
corr_lookup_kernel(const __grid_constant__ LookupParams p) {
  using S = LookupSmem<R, PB>;
  constexpr int kPosPerBlock = PB;   // shadows the file-level default inside this kernel
  constexpr int K = S::K, T = S::T, kStride = S::kStride;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int P = p.H * p.W;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * kPosPerBlock;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npos = min(kPosPerBlock, P - i0);
  const int L = p.L;

  auto taps_of = [&](int l) { return reinterpret_cast<float*>(smem_raw + (size_t)l * S::kPerLevelBytes); };
  auto fx_of = [&](int l) { return taps_of(l) + kPosPerBlock * kStride; };
  auto fy_of = [&](int l) { return fx_of(l) + kPosPerBlock * K; };
  auto org_of = [&](int l) { return reinterpret_cast<int*>(fy_of(l) + kPosPerBlock * K); };

  // 1) window geometry, one thread per (position, level)
  for (int t = threadIdx.x; t < kPosPerBlock * L; t += blockDim.x) {
    const int l = t / kPosPerBlock, pos = t % kPosPerBlock;
    const int hl = p.h[l], wl = p.w[l];
    float cx = 0.f, cy = 0.f;
    if (pos < npos) {
      cx = __ldg(p.coords + ((int64_t)b * 2 + 0) * P + i0 + pos);
      cy = __ldg(p.coords + ((int64_t)b * 2 + 1) * P + i0 + pos);
    }
    const float inv = 1.0f / (float)(1 << l);
    const float lx = cx * inv, ly = cy * inv;  // exact: power-of-two scaling
    // Clamp far-away / non-finite centres so the integer origin stays representable; every tap of
    // such a window is out of the map and reads as zero either way.
    const float ox = floorf(fminf(fmaxf(roundtrip(lx - (float)R, wl), -1.0e6f), 1.0e6f));
    const float oy = floorf(fminf(fmaxf(roundtrip(ly - (float)R, hl), -1.0e6f), 1.0e6f));
    org_of(l)[pos * 2 + 0] = (int)ox;
    org_of(l)[pos * 2 + 1] = (int)oy;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      // fraction relative to tap column a of the shared window; equals the reference's
      // (ix - floor(ix)) except on knife-edge roundings, where it extrapolates by <= 1 ulp.
      fx_of(l)[pos * K + a] = roundtrip(lx + (float)(a - R), wl) - (ox + (float)a);
      fy_of(l)[pos * K + a] = roundtrip(ly + (float)(a - R), hl) - (oy + (float)a);
    }
  }
  __syncthreads();

  // 2) gather: thread = (position, tap column); T row taps per level, all levels back to back.
  //    cp.async (4-byte, zero-filling when the tap lies outside the map) moves every tap straight
  //    from global to shared memory, so all (2r+2)*L loads of a thread are in flight at once with
  //    no register staging; one wait + barrier afterwards.
  {
    const int pos = threadIdx.x / T, col = threadIdx.x - pos * T;
    if (pos < npos) {  // (threads beyond 32*T have pos >= 32 and only take part in phase 3)
      for (int l = 0; l < L; ++l) {
        const int hl = p.h[l], wl = p.w[l];
        const int64_t plane = (int64_t)hl * wl;
        if (plane == 0) continue;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(taps_of(l) + pos * kStride + col);
        const int x = org_of(l)[pos * 2 + 0] + col, y0 = org_of(l)[pos * 2 + 1];
        const bool x_ok = (unsigned)x < (unsigned)wl;
        // 32-bit element offsets inside this position's map (h_l*w_l < 2^31); an out-of-map tap copies
        // 0 bytes from offset 0 (zero-fill), so every address handed to cp.async is valid.
        const float* base = p.level[l] + ((int64_t)b * P + i0 + pos) * plane;
        const int off0 = y0 * wl + x;
#pragma unroll
        for (int r = 0; r < T; ++r) {
          const bool ok = x_ok && (unsigned)(y0 + r) < (unsigned)hl;
          const int off = ok ? off0 + r * wl : 0;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + r * T * 4), "l"(base + off),
                       "r"(ok ? 4 : 0)
                       : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  // 3) interpolate.  A task = (level, group of G adjacent window columns); tasks are dealt to the
  //    warps, lane = position.  Per tap row a thread reads G+1 taps and forms G horizontal lerps
  //    (adjacent columns share a tap), then combines with the previous row: (G+1)*T tap loads per
  //    G*K outputs, and every store instruction writes 32 consecutive positions of one channel.
  constexpr int G = S::kColsPerTask, kGroups = S::kGroups;
  // slot = a group of PB lanes working on one task; with PB = 16 a warp runs two tasks side by side
  const int n_slots = (blockDim.x >> 5) * S::kSlotsPerWarp;
  const int slot = warp * S::kSlotsPerWarp + lane / PB;
  const int lane_pos = lane % PB;
  if (lane_pos < npos) {
    for (int task = slot; task < L * kGroups; task += n_slots) {
      const int l = task / kGroups, a0 = (task - l * kGroups) * G;
      float* o = p.out + ((int64_t)b * L * K * K + (int64_t)l * K * K + a0 * K) * P + i0 + lane_pos;
      if ((int64_t)p.h[l] * p.w[l] == 0) {  // level pooled away: an empty map contributes zeros
#pragma unroll
        for (int g = 0; g < G; ++g)
          if (a0 + g < K)
#pragma unroll
            for (int c = 0; c < K; ++c) st_stream(o + (int64_t)(g * K + c) * P, 0.f);
        continue;
      }
      const float* tp = taps_of(l) + lane_pos * kStride + a0;
      const float* fyp = fy_of(l) + lane_pos * K;
      float fx[G], prev[G];
#pragma unroll
      for (int g = 0; g < G; ++g) fx[g] = (a0 + g < K) ? fx_of(l)[lane_pos * K + a0 + g] : 0.f;
      {
        float t[G + 1];
#pragma unroll
        for (int g = 0; g <= G; ++g) t[g] = (a0 + g < T) ? tp[g] : 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) prev[g] = t[g] + fx[g] * (t[g + 1] - t[g]);
      }
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const float* row = tp + (c + 1) * T;
        float t[G + 1];
#pragma unroll
        for (int g = 0; g <= G; ++g) t[g] = (a0 + g < T) ? row[g] : 0.f;
        const float fy = fyp[c];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float cur = t[g] + fx[g] * (t[g + 1] - t[g]);
          if (a0 + g < K) st_stream(o + (int64_t)(g * K + c) * P, prev[g] + fy * (cur - prev[g]));
          prev[g] = cur;
        }
      }
    }
  }
}

template <int R, int PB>
int launch_lookup(const LookupParams& p, cudaStream_t stream) {
  using S = LookupSmem<R, PB>;
  const size_t smem = (size_t)p.L * S::kPerLevelBytes;
  dim3 grid((unsigned)ceil_div(p.H * p.W, PB), (unsigned)p.B);
  static DynSmemOptIn optin;
  EEM_CHECK_CUDA(optin.ensure(corr_lookup_kernel<R, PB>, smem));
  corr_lookup_kernel<R, PB><<<grid, S::kThreads, smem, stream>>>(p);
  return EEM_OK;
}

// Positions per CTA: 32.  EEM_LOOKUP_PB=16 selects a CTA with half the shared memory (~7 instead of 3 CTAs per
// SM, two interpolation tasks per warp) for timing comparisons: measured SLOWER on B200 (50.3 vs 40.0 us per
// launch at MVSEC B=32), so more phase-staggered CTAs is not what the kernel lacks (profiles/r01/README.md).
template <int R>
int launch_lookup_any(const LookupParams& p, cudaStream_t stream) {
  static const int pb = [] {
    const char* v = getenv("EEM_LOOKUP_PB");
    return (v != nullptr && atoi(v) == 16) ? 16 : 32;
  }();
  return pb == 32 ? launch_lookup<R, 32>(p, stream) : launch_lookup<R, 16>(p, stream);
}

__global__ void __launch_bounds__(256)
avg_pool2x2_kernel(const float* __restrict__ in, int64_t n_planes, int h, int w, float* __restrict__ out) {
  const int ho = h / 2, wo = w / 2;
  const int64_t total = n_planes * ho * wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo);
    const int64_t r = i / wo;
    const int y = (int)(r % ho);
    const int64_t pl = r / ho;
    const float* s = in + pl * h * w + (int64_t)(2 * y) * w + 2 * x;
    // same summation order as ATen's avg_pool2d inner loop (row-major over the 2x2 window)
    out[i] = (((s[0] + s[1]) + s[w]) + s[w + 1]) * 0.25f;
  }
}


// ---- backward of the lookup: d(level_l) from d(out) -------------------------------------------------
// out[b, l*K*K + a*K + c, i] = bilinear(level_l[b*P+i], window tap (a, c)), so the gradient of row
// (b, i) of level l is non-zero only inside that position's (K+1)^2 footprint.  A CTA owns 32
// consecutive positions: it stages their L*K*K output gradients in shared memory (coalesced, 128 B per
// channel), then writes every row of every level in full -- zeros outside the footprint, and inside it
// the <= 4 contributions per cell gathered in a fixed order.  No atomics, no separate memset, and the
// result is deterministic.  Coordinates carry no gradient (the callers detach them, model/eraft.py:141).
struct LookupBwdParams {
  float* dlevel[kMaxLevels];
  int h[kMaxLevels], w[kMaxLevels];
  int B, H, W, L;
  const float* coords;
  const float* gout;
};

template <int R>
__global__ void __launch_bounds__(256)
corr_lookup_backward_kernel(const __grid_constant__ LookupBwdParams p) {
  constexpr int K = 2 * R + 1, T = K + 1, KK = K * K;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int P = p.H * p.W, L = p.L;
  const int b = blockIdx.y, i0 = blockIdx.x * kPosPerBlock;
  const int npos = min(kPosPerBlock, P - i0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  float* gs = reinterpret_cast<float*>(smem_raw);                 // [L*KK][33]
  float* fxs = gs + (size_t)L * KK * 33;                           // [L][32][K]
  float* fys = fxs + (size_t)L * kPosPerBlock * K;                 // [L][32][K]
  int* org = reinterpret_cast<int*>(fys + (size_t)L * kPosPerBlock * K);   // [L][32][2]

  for (int ch = warp; ch < L * KK; ch += 8)
    gs[ch * 33 + lane] = lane < npos ? __ldg(p.gout + ((int64_t)b * L * KK + ch) * P + i0 + lane) : 0.f;
  for (int t = threadIdx.x; t < kPosPerBlock * L; t += blockDim.x) {   // same geometry as the forward kernel
    const int l = t / kPosPerBlock, pos = t % kPosPerBlock;
    const int hl = p.h[l], wl = p.w[l];
    float cx = 0.f, cy = 0.f;
    if (pos < npos) {
      cx = __ldg(p.coords + ((int64_t)b * 2 + 0) * P + i0 + pos);
      cy = __ldg(p.coords + ((int64_t)b * 2 + 1) * P + i0 + pos);
    }
    const float inv = 1.0f / (float)(1 << l);
    const float lx = cx * inv, ly = cy * inv;
    const float ox = floorf(fminf(fmaxf(roundtrip(lx - (float)R, wl), -1.0e6f), 1.0e6f));
    const float oy = floorf(fminf(fmaxf(roundtrip(ly - (float)R, hl), -1.0e6f), 1.0e6f));
    org[(l * kPosPerBlock + pos) * 2 + 0] = (int)ox;
    org[(l * kPosPerBlock + pos) * 2 + 1] = (int)oy;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      fxs[(l * kPosPerBlock + pos) * K + a] = roundtrip(lx + (float)(a - R), wl) - (ox + (float)a);
      fys[(l * kPosPerBlock + pos) * K + a] = roundtrip(ly + (float)(a - R), hl) - (oy + (float)a);
    }
  }
  __syncthreads();

  for (int l = 0; l < L; ++l) {
    const int hl = p.h[l], wl = p.w[l];
    const int plane = hl * wl;
    if (plane == 0) continue;
    const int step_y = 256 / wl, step_x = 256 - step_y * wl;
    for (int pos = 0; pos < npos; ++pos) {
      float* row = p.dlevel[l] + ((int64_t)b * P + i0 + pos) * plane;
      const int ox = org[(l * kPosPerBlock + pos) * 2 + 0], oy = org[(l * kPosPerBlock + pos) * 2 + 1];
      const float* fx = fxs + (l * kPosPerBlock + pos) * K;
      const float* fy = fys + (l * kPosPerBlock + pos) * K;
      const float* g = gs + (size_t)l * KK * 33 + pos;
      int y = (int)threadIdx.x / wl, x = (int)threadIdx.x - y * wl;
      for (int j = threadIdx.x; j < plane; j += 256) {
        const int u = x - ox, v = y - oy;
        float val = 0.f;
        if ((unsigned)u < (unsigned)T && (unsigned)v < (unsigned)T) {
#pragma unroll
          for (int da = 1; da >= 0; --da) {       // tap column a = u - da gives weight fx[a] (da=1) or 1-fx[a] (da=0)
            const int a = u - da;
            if ((unsigned)a >= (unsigned)K) continue;
            const float wx = da ? fx[a] : 1.0f - fx[a];
#pragma unroll
            for (int dc = 1; dc >= 0; --dc) {
              const int c = v - dc;
              if ((unsigned)c >= (unsigned)K) continue;
              const float wy = dc ? fy[c] : 1.0f - fy[c];
              val = fmaf(g[(a * K + c) * 33], wx * wy, val);
            }
          }
        }
        st_stream(row + j, val);
        x += step_x;
        y += step_y;
        if (x >= wl) { x -= wl; ++y; }
      }
    }
  }
}

template <int R>
int launch_lookup_backward(const LookupBwdParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int K = 2 * R + 1;
  const size_t smem = ((size_t)p.L * K * K * 33 + 2 * (size_t)p.L * kPosPerBlock * K) * 4 + (size_t)p.L * kPosPerBlock * 2 * 4;
  static DynSmemOptIn optin;
  EEM_CHECK_CUDA(optin.ensure(corr_lookup_backward_kernel<R>, smem));
  corr_lookup_backward_kernel<R><<<grid, 256, smem, stream>>>(p);
  return EEM_OK;
}

// gradient of avg_pool2d(2,2,floor): every fine cell under a coarse cell receives a quarter of it; the
// odd last row / column (cropped by the floor) receives nothing.
__global__ void __launch_bounds__(256)
avg_pool2x2_backward_kernel(const float* __restrict__ gout, int64_t n_planes, int h, int w, float* __restrict__ gin,
                            int accumulate) {
  const int ho = h / 2, wo = w / 2;
  const int64_t total = n_planes * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int64_t r = i / w;
    const int y = (int)(r % h);
    const int64_t pl = r / h;
    float v = 0.f;
    if ((y >> 1) < ho && (x >> 1) < wo) v = 0.25f * __ldg(gout + (pl * ho + (y >> 1)) * wo + (x >> 1));
    gin[i] = accumulate ? gin[i] + v : v;
  }
}

}  // namespace
}  // namespace eem

using namespace eem;

extern "C" {

int eem_corr_lookup(const float* const* levels, int B, int H, int W, int num_levels, int radius,
                    const float* coords, float* out, eem_stream_t stream_) {
  EEM_CHECK_ARG(levels && coords && out, "eem_corr_lookup: NULL pointer");
  EEM_CHECK_ARG(B > 0 && H > 0 && W > 0, "eem_corr_lookup: sizes must be > 0");
  EEM_CHECK_ARG(num_levels > 0 && num_levels <= kMaxLevels, "eem_corr_lookup: num_levels must be in [1,%d]", kMaxLevels);
  EEM_CHECK_ARG(B <= 65535, "eem_corr_lookup: batch > 65535 not supported in one call");
  LookupParams p{};
  int h = H, w = W;
  for (int l = 0; l < num_levels; ++l) {
    p.level[l] = levels[l];
    p.h[l] = h;
    p.w[l] = w;
    EEM_CHECK_ARG((int64_t)h * w == 0 || levels[l] != nullptr, "eem_corr_lookup: levels[%d] is NULL", l);
    h /= 2;
    w /= 2;
  }
  p.B = B; p.H = H; p.W = W; p.L = num_levels;
  p.coords = coords;
  p.out = out;
  cudaStream_t stream = as_stream(stream_);
  int rc = EEM_OK;
  switch (radius) {
    case 4: rc = launch_lookup_any<4>(p, stream); break;
    case 3: rc = launch_lookup_any<3>(p, stream); break;
    case 2: rc = launch_lookup_any<2>(p, stream); break;
    case 1: rc = launch_lookup_any<1>(p, stream); break;
    default:
      return fail(EEM_ERR_UNSUPPORTED, "eem_corr_lookup: radius %d not in {1,2,3,4}", radius);
  }
  if (rc != EEM_OK) return rc;
  EEM_CHECK_LAUNCH("corr_lookup_kernel");
  return EEM_OK;
}

int eem_avg_pool2x2(const float* in, int64_t n_planes, int h, int w, float* out, eem_stream_t stream_) {
  EEM_CHECK_ARG(in && out, "eem_avg_pool2x2: NULL pointer");
  EEM_CHECK_ARG(n_planes > 0 && h > 0 && w > 0, "eem_avg_pool2x2: sizes must be > 0");
  const int64_t total = n_planes * (h / 2) * (w / 2);
  if (total == 0) return EEM_OK;
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (cap > 0 && blocks > cap) blocks = cap;
  avg_pool2x2_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream_)>>>(in, n_planes, h, w, out);
  EEM_CHECK_LAUNCH("avg_pool2x2_kernel");
  return EEM_OK;
}

int eem_corr_lookup_backward(const float* grad_out, const float* coords, int B, int H, int W, int num_levels, int radius,
                             float* const* grad_levels, eem_stream_t stream_) {
  EEM_CHECK_ARG(grad_out && coords && grad_levels, "eem_corr_lookup_backward: NULL pointer");
  EEM_CHECK_ARG(B > 0 && H > 0 && W > 0, "eem_corr_lookup_backward: sizes must be > 0");
  EEM_CHECK_ARG(num_levels > 0 && num_levels <= kMaxLevels, "eem_corr_lookup_backward: num_levels must be in [1,%d]", kMaxLevels);
  EEM_CHECK_ARG(B <= 65535, "eem_corr_lookup_backward: batch > 65535 not supported in one call");
  LookupBwdParams p{};
  int h = H, w = W;
  for (int l = 0; l < num_levels; ++l) {
    p.dlevel[l] = grad_levels[l];
    p.h[l] = h;
    p.w[l] = w;
    EEM_CHECK_ARG((int64_t)h * w == 0 || grad_levels[l] != nullptr, "eem_corr_lookup_backward: grad_levels[%d] is NULL", l);
    h /= 2;
    w /= 2;
  }
  p.B = B; p.H = H; p.W = W; p.L = num_levels;
  p.coords = coords;
  p.gout = grad_out;
  dim3 grid((unsigned)ceil_div(H * W, kPosPerBlock), (unsigned)B);
  cudaStream_t stream = as_stream(stream_);
  int rc = EEM_OK;
  switch (radius) {
    case 4: rc = launch_lookup_backward<4>(p, grid, stream); break;
    case 3: rc = launch_lookup_backward<3>(p, grid, stream); break;
    case 2: rc = launch_lookup_backward<2>(p, grid, stream); break;
    case 1: rc = launch_lookup_backward<1>(p, grid, stream); break;
    default:
      return fail(EEM_ERR_UNSUPPORTED, "eem_corr_lookup_backward: radius %d not in {1,2,3,4}", radius);
  }
  if (rc != EEM_OK) return rc;
  EEM_CHECK_LAUNCH("corr_lookup_backward_kernel");
  return EEM_OK;
}

int eem_avg_pool2x2_backward(const float* grad_out, int64_t n_planes, int h, int w, float* grad_in, int accumulate,
                             eem_stream_t stream_) {
  EEM_CHECK_ARG(grad_in && (grad_out || (h / 2) * (w / 2) == 0), "eem_avg_pool2x2_backward: NULL pointer");
  EEM_CHECK_ARG(n_planes > 0 && h > 0 && w > 0, "eem_avg_pool2x2_backward: sizes must be > 0");
  const int64_t total = n_planes * h * w;
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (cap > 0 && blocks > cap) blocks = cap;
  avg_pool2x2_backward_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream_)>>>(grad_out, n_planes, h, w, grad_in, accumulate);
  EEM_CHECK_LAUNCH("avg_pool2x2_backward_kernel");
  return EEM_OK;
}

}  // extern "C"
