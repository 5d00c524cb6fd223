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

#include <cuda_fp16.h>

#include <cstdio>

#include "common.cuh"
#include "tc_common.cuh"
#include "packed_layout.cuh"

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

// ---- lookup on the packed fp16 working pyramid (packed_layout.cuh) ----------------------------------------------
// Same outputs as corr_lookup_kernel, different input: every level of a source position is stored as 4x4-pixel
// tiles of fp16 (32 B = one sector each) in ONE row per position.  A CTA owns 32 consecutive positions.  Phases:
// (1) window geometry per (position, level) -- identical arithmetic to the f32 kernel; (2) ONE gather phase: the
// window rows are copied out of the NT tiles per tile row that cover a position's (2r+2)^2 tap window with 8-byte
// cp.async (zero-filled when the tile lies outside the map; cells of a partial last tile are zeros in the volume
// itself), all levels back to back, every load in flight before anything is consumed; (3) interpolation in fp32 with
// warp = (level, 3 window columns), lane = position, so each store instruction writes 32 consecutive positions of
// one output channel (128 B).  Per position the kernel fetches ~10.6 sectors per full-size level instead of the
// 16 + that the f32 row-major volume costs, and half the bytes per tap.
struct PackedLookupParams {
  const uint16_t* packed;
  int h[kMaxLevels], w[kMaxLevels], tx[kMaxLevels], ty[kMaxLevels], off[kMaxLevels];
  float rcp_w1[kMaxLevels], rcp_h1[kMaxLevels];     // RN(1 / (w_l - 1)), RN(1 / (h_l - 1)) (inf for a 1-wide map)
  int row;
  int B, H, W, L;
  const float* coords;
  float* out;
  int debug;       // EEM_LOOKUP_DEBUG (timing experiments only): 1 no output stores, 2 no tile copies, 4 no interpolation
};

// roundtrip() with the IEEE division replaced by a multiply and two FMAs: q0 = RN(x * y), r = x - q0 * s1 (exact in
// one FMA), q = RN(q0 + r * y) with y = RN(1 / s1) is the correctly rounded quotient x / s1 (Markstein's theorem; the
// divisor is a small integer, results in the normal range), i.e. the same bits as __fdiv_rn without its ~10
// instruction sequence and slow-path branch -- 20 of these per (position, level) made up half of the geometry phase.
__device__ __forceinline__ float roundtrip_rcp(float p, float s1, float rcp_s1) {
  const float x = __fmul_rn(2.0f, p);
  const float q0 = __fmul_rn(x, rcp_s1);
  const float r = __fmaf_rn(-q0, s1, x);
  const float q = __fmaf_rn(r, rcp_s1, q0);
  const float g = __fsub_rn(q, 1.0f);
  return __fmul_rn(__fadd_rn(g, 1.0f), s1 * 0.5f);
}

template <int R>
struct PackedSmem {
  static constexpr int K = 2 * R + 1, T = K + 1;
  static constexpr int NT = (T + 6) / 4;                    // tiles per dimension covering T taps at any phase 0..3
  static constexpr int PB = 32;
  // Patch of one (level, position) in shared memory: T rows (the window's rows, row 0 = the window origin row) of
  // NT*4 fp16 (the NT tiles that cover the window's columns; column 0 = first column of the first tile), row-major.
  static constexpr int kRowBytes = NT * 8;
  static constexpr int kPatchBytes = T * kRowBytes;
  // + 8 B: consecutive positions start 2 banks apart (mod 32): the 4-byte tap reads of a warp conflict 2-way at most
  static constexpr int kStrideBytes = kPatchBytes + ((kPatchBytes / 4) % 32 == 2 ? 0 : 8);
  static constexpr int kColsPerTask = 3, kGroups = (K + kColsPerTask - 1) / kColsPerTask;
  static constexpr int kThreads = 32 * 4 * kGroups;          // one interpolation task per warp for a 4-level pyramid
  // per position: window origin (ox, oy), the NT tile columns of the window clamped into the map (one byte each) and
  // a bit mask of the columns that really lie inside it -- everything the gather needs besides the window row
  static constexpr int kPerLevelBytes = PB * kStrideBytes + 2 * PB * K * 4 + PB * 4 * 4;
};

// The three phases of a batch of PB positions, shared by the one-batch-per-CTA kernel and the pipelined persistent one.
template <int R, bool kDbg = false>
struct PackedPhases {
  using S = PackedSmem<R>;
  static constexpr int K = S::K, T = S::T, NT = S::NT, PB = S::PB;

  static __device__ __forceinline__ unsigned char* patch_of(unsigned char* st, int l) { return st + (size_t)l * S::kPerLevelBytes; }
  static __device__ __forceinline__ float* fx_of(unsigned char* st, int l) { return reinterpret_cast<float*>(patch_of(st, l) + PB * S::kStrideBytes); }
  static __device__ __forceinline__ float* fy_of(unsigned char* st, int l) { return fx_of(st, l) + PB * K; }
  static __device__ __forceinline__ int* org_of(unsigned char* st, int l) { return reinterpret_cast<int*>(fy_of(st, l) + PB * K); }

  // 1) window geometry, one thread per (position, level, axis): same arithmetic as corr_lookup_kernel.  The x and the
  //    y half of a (position, level) are independent (origin + K fractions each), so they run on different warps: the
  //    phase is a dependent chain of ~10 coordinate round trips per thread and sits on the CTA's critical path.
  //    (tid, nthr): the calling thread's index in the group of threads that shares the phase, and the group's size
  static __device__ __forceinline__ void geometry(const PackedLookupParams& p, unsigned char* st, int b, int i0, int npos, int l0, int nl,
                                                  int tid, int nthr) {
    const int P = p.H * p.W;
    const int per_axis = PB * nl;
    for (int t = tid; t < 2 * per_axis; t += nthr) {
      const int axis = t >= per_axis ? 1 : 0, r = t - axis * per_axis;      // warp-uniform: PB is a warp
      const int ls = r / PB, l = l0 + ls, pos = r % PB;                     // ls: level slot in shared memory
      const float s1 = (float)((axis ? p.h[l] : p.w[l]) - 1);
      const float rc = axis ? p.rcp_h1[l] : p.rcp_w1[l];
      float c = 0.f;
      if (pos < npos) c = __ldg(p.coords + ((int64_t)b * 2 + axis) * P + i0 + pos);
      const float lc = c * (1.0f / (float)(1 << l));
      const float o = floorf(fminf(fmaxf(roundtrip_rcp(lc - (float)R, s1, rc), -1.0e6f), 1.0e6f));
      int* org = org_of(st, ls) + pos * 4;
      float* frac = (axis ? fy_of(st, ls) : fx_of(st, ls)) + pos * K;
      if (axis == 0) {
        // record of (position, level) for the gather and the interpolation: .x = (ox & 3) | in-map mask of the NT tile
        // columns << 4, .y = oy (written by the y thread), .z / .w = byte offsets (tile column * 32, 0 when outside the
        // map) of the tile columns as four 16-bit fields
        const int oxi = (int)o, tx0 = oxi >> 2, txl = p.tx[l];      // >> : floor for negatives
        unsigned lo = 0, hi = 0, mask = 0;
#pragma unroll
        for (int sx = 0; sx < NT; ++sx) {
          const int tx = tx0 + sx;
          const bool ok = (unsigned)tx < (unsigned)txl;
          const unsigned field = (unsigned)(ok ? tx * 32 : 0) << (16 * (sx & 1));
          if (sx < 2) lo |= field; else hi |= field;
          mask |= (unsigned)ok << sx;
        }
        org[0] = (oxi & 3) | (int)(mask << 4);
        *reinterpret_cast<int2*>(org + 2) = make_int2((int)lo, (int)hi);
      } else {
        org[1] = (int)o;
      }
#pragma unroll
      for (int a = 0; a < K; ++a) frac[a] = roundtrip_rcp(lc + (float)(a - R), s1, rc) - (o + (float)a);
    }
  }

  // 2) gather: unit = (position, window row), all levels; the thread copies that row's 8-byte piece out of each of the
  //    NT tiles the window's columns touch (zero-filled when the tile lies outside the map; cells of a partial last tile
  //    are zeros in the volume itself), so the patch in shared memory is a plain row-major image whose row 0 is the
  //    window's first row.  Everything position-dependent comes out of the geometry record with a handful of integer
  //    instructions (32-bit byte offsets from the position's row, one 64-bit multiply-add per copy); the level loop is
  //    unrolled so the per-level constants are direct constant-bank operands.  Issues the copies and commits ONE
  //    cp.async group.
  static __device__ __forceinline__ void gather(const PackedLookupParams& p, unsigned char* st, int b, int i0, int npos, int l0, int nl,
                                                int tid, int nthr) {
    const int P = p.H * p.W;
    for (int u = tid; u < PB * T; u += nthr) {                     // (position, window row): decoded once, then all levels
      const int pos = u / T, rr = u - pos * T;
      if (pos >= npos) continue;
      const unsigned char* rowbase = reinterpret_cast<const unsigned char*>(p.packed + ((int64_t)b * P + i0 + pos) * p.row);
      const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(st + pos * S::kStrideBytes + rr * S::kRowBytes);
      const uint32_t org0 = (uint32_t)__cvta_generic_to_shared(org_of(st, 0) + pos * 4);
#pragma unroll
      for (int ls = 0; ls < kMaxLevels; ++ls) {
        if (ls >= nl) break;
        const int l = l0 + ls;
        const int txl = p.tx[l], tyl = p.ty[l];
        if (txl == 0 || tyl == 0) continue;                        // empty level
        int4 g;                                                    // (ox & 3) | mask << 4, oy, tile-column byte offsets
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(g.x), "=r"(g.y), "=r"(g.z), "=r"(g.w) : "r"(org0 + ls * S::kPerLevelBytes));
        const int py = g.y + rr;                                   // map row of this window row
        const int ty = py >> 2;                                    // >> : floor for negatives
        const bool row_ok = (unsigned)ty < (unsigned)tyl;
        const unsigned live = row_ok ? ((unsigned)g.x >> 4) : 0u;  // tile columns to fetch
        const unsigned roff = (unsigned)p.off[l] * 2u + (row_ok ? (unsigned)(ty * txl * 32 + (py & 3) * 8) : 0u);
        const uint32_t dst = dst0 + ls * S::kPerLevelBytes;
#pragma unroll
        for (int sx = 0; sx < NT; ++sx) {
          const unsigned field = ((sx < 2 ? (unsigned)g.z : (unsigned)g.w) >> (16 * (sx & 1))) & 0xffffu;
          const unsigned char* src;
          asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(src) : "r"(roff + field), "l"(rowbase));
          asm volatile(
              "{\n"
              ".reg .pred q;\n"
              "setp.eq.u32 q, %2, 0;\n"
              "cp.async.ca.shared.global [%0], [%1], 8, q;\n"       // q = ignore-src: the 8 bytes are zero-filled
              "}\n" ::"r"(dst + sx * 8), "l"(src), "r"((kDbg && (p.debug & 2)) ? 0u : (live & (1u << sx))) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // 3) interpolate: task = (level, group of G adjacent window columns), lane = position.  Per window row a thread
  //    reads three 4-byte words (6 fp16: the G + 1 = 4 taps it needs start at an even or odd element), shifts them
  //    into place with two funnel shifts, and forms G horizontal lerps + G vertical lerps; row addresses are
  //    compile-time offsets of one per-task base pointer.
  //    (warp, n_warps): the calling warp's index among the warps that share the interpolation, and their number
  static __device__ __forceinline__ void interp(const PackedLookupParams& p, unsigned char* st, int b, int i0, int npos, int l0, int nl,
                                                int warp, int n_warps) {
    constexpr int G = S::kColsPerTask, kGroups = S::kGroups;
    static_assert(G == 3, "load_row unpacks G + 1 = 4 taps");
    const int P = p.H * p.W, L = p.L;
    const int lane = threadIdx.x & 31;
    if (lane >= npos || (kDbg && (p.debug & 4))) return;
    // (L2 residency of the tiles across the successive lookups of a pyramid was tried: an evict_first policy on the output
    // stream changes nothing measurable -- 28.9 vs 28.4 us -- and cp.async with an evict_last cache policy assembles to an
    // instruction the hardware rejects in this kernel, cudaErrorIllegalInstruction with CUDA 12.9.)
    for (int task = warp; task < nl * kGroups; task += n_warps) {
      const int ls = task / kGroups, l = l0 + ls, a0 = (task - ls * kGroups) * G;
      float* o = p.out + ((int64_t)b * L * K * K + (int64_t)l * K * K + a0 * K) * P + i0 + lane;
      if (p.h[l] * p.w[l] == 0) {            // level pooled away: an empty map contributes zeros
#pragma unroll
        for (int g = 0; g < G; ++g)
          if (a0 + g < K)
#pragma unroll
            for (int c = 0; c < K; ++c) st_stream(o + (int64_t)(g * K + c) * P, 0.f);
        continue;
      }
      const int q0 = (org_of(st, ls)[lane * 4 + 0] & 3) + a0;    // first tap column of this task inside the patch row
      // 12 bytes per row from word q0/2 on.  For the last column group of a narrow window (r < 4) the third word may
      // lie one word past the row: it only feeds taps that are never used, and the bytes read are still inside this
      // stage's shared memory (the next row / the position's pad / the fraction arrays).
      const unsigned char* base = patch_of(st, ls) + lane * S::kStrideBytes + (q0 >> 1) * 4;
      const unsigned sh = (unsigned)(q0 & 1) * 16u;
      const float* fyp = fy_of(st, ls) + lane * K;
      float fx[G], prev[G];
#pragma unroll
      for (int g = 0; g < G; ++g) fx[g] = (a0 + g < K) ? fx_of(st, ls)[lane * K + a0 + g] : 0.f;
      auto load_row = [&](int rr, float (&t)[G + 1]) {
        const uint32_t* rowp = reinterpret_cast<const uint32_t*>(base + rr * S::kRowBytes);
        const uint32_t x0 = rowp[0], x1 = rowp[1], x2 = rowp[2];
        const uint32_t v_lo = __funnelshift_r(x0, x1, sh), v_hi = __funnelshift_r(x1, x2, sh);
        const float2 ab = __half22float2(*reinterpret_cast<const __half2*>(&v_lo));
        const float2 cd = __half22float2(*reinterpret_cast<const __half2*>(&v_hi));
        t[0] = ab.x; t[1] = ab.y; t[2] = cd.x; t[3] = cd.y;
      };
      {
        float t[G + 1];
        load_row(0, t);
#pragma unroll
        for (int g = 0; g < G; ++g) prev[g] = t[g] + fx[g] * (t[g + 1] - t[g]);
      }
#pragma unroll
      for (int c = 0; c < K; ++c) {
        float t[G + 1];
        load_row(c + 1, t);
        const float fy = fyp[c];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float cur = t[g] + fx[g] * (t[g + 1] - t[g]);
          // output channel (a0 + g) * K + c: one 32 x 32 -> 64-bit multiply-add per store address
          float* dst;
          asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(dst) : "r"(P), "r"((g * K + c) * 4), "l"(o));
          if (a0 + g < K && !(kDbg && (p.debug & 1))) st_stream(dst, prev[g] + fy * (cur - prev[g]));
          prev[g] = cur;
        }
      }
    }
  }
};

// One batch per CTA (the default).
template <int R, bool kDbg>
__global__ void __launch_bounds__(PackedSmem<R>::kThreads)
corr_lookup_packed_kernel(const __grid_constant__ PackedLookupParams p) {
  using Ph = PackedPhases<R, kDbg>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int P = p.H * p.W;
  const int b = blockIdx.y, i0 = blockIdx.x * Ph::PB;
  const int npos = min(Ph::PB, P - i0);
  const int tid = threadIdx.x, nthr = blockDim.x;
  Ph::geometry(p, smem_raw, b, i0, npos, 0, p.L, tid, nthr);
  __syncthreads();
  Ph::gather(p, smem_raw, b, i0, npos, 0, p.L, tid, nthr);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  Ph::interp(p, smem_raw, b, i0, npos, 0, p.L, tid >> 5, nthr >> 5);
}

// Persistent, warp-specialised form: the three phases of the one-batch kernel run on DIFFERENT warps of a CTA and are
// coupled by mbarriers over a ring of shared-memory stages, so the window geometry of batch k+2, the tile copies of batch
// k+1 (and their DRAM latency) and the interpolation + output stores of batch k overlap inside every SM -- in the
// one-batch kernel the phases are serial per CTA and only overlap by chance between the CTAs of an SM (measured there:
// geometry + copy issue 10.4 us, tile copies + 8 us, interpolation + 7 us, output stores + 9.4 us ~ the whole 34.7 us,
// scripts/lookup_ablation.py).
//   warps [0, n_cons)            consumers: wait full[s], interpolate stage s (one (level, column group) task per warp),
//                                store, arrive empty[s]
//   the next n_geo warps         geometry: wait empty[s], window geometry of the batch into stage s, arrive ready[s]
//   the last n_copy warps        copies: wait ready[s], issue the tile copies (one cp.async group per batch); full[s'] of
//                                the PREVIOUS batch is signalled once its group has landed (cp.async.wait_group 1), so the
//                                copy warps always run one batch ahead of the data they wait for.
// A CTA walks batches blockIdx.x, blockIdx.x + gridDim.x, ...
constexpr int kWsMaxProdWarps = 18;
// default configuration (0: one-batch kernel), see ws_config()
struct WsConfig {
  int stages, per_sm, geo_warps, copy_warps;
};

template <int R, bool kDbg>
__global__ void __launch_bounds__(PackedSmem<R>::kThreads + kWsMaxProdWarps * 32)
corr_lookup_packed_ws_kernel(const __grid_constant__ PackedLookupParams p, int batches_per_sample, int n_batches, int stage_bytes,
                             int n_stages, int n_geo_warps, int n_copy_warps) {
  using Ph = PackedPhases<R, kDbg>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)n_stages * stage_bytes);
  uint64_t* empty = full + n_stages;
  uint64_t* ready = empty + n_stages;
  const int P = p.H * p.W;
  const int n_cons_warps = (int)(blockDim.x >> 5) - n_geo_warps - n_copy_warps;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      tc::mbar_init(&full[i], n_copy_warps * 32);
      tc::mbar_init(&empty[i], n_cons_warps);
      tc::mbar_init(&ready[i], n_geo_warps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto decode = [&](int bid, int& b, int& i0, int& npos) {
    b = bid / batches_per_sample;
    i0 = (bid - b * batches_per_sample) * Ph::PB;
    npos = min(Ph::PB, P - i0);
  };
  const int first = blockIdx.x, step = gridDim.x;
  if (warp < n_cons_warps) {
    // ===== consumers =====
    int s = 0;
    uint32_t parity = 0;
    for (int bid = first; bid < n_batches; bid += step) {
      int b, i0, npos;
      decode(bid, b, i0, npos);
      tc::mbar_wait(&full[s], parity);
      Ph::interp(p, smem_raw + (size_t)s * stage_bytes, b, i0, npos, 0, p.L, warp, n_cons_warps);
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&empty[s]);
      if (++s == n_stages) { s = 0; parity ^= 1; }
    }
  } else if (warp < n_cons_warps + n_geo_warps) {
    // ===== geometry =====
    const int gtid = (int)threadIdx.x - n_cons_warps * 32, gn = n_geo_warps * 32;
    int s = 0;
    uint32_t parity = 1;                       // a fresh barrier passes a wait on the phase "before" its first one
    for (int bid = first; bid < n_batches; bid += step) {
      int b, i0, npos;
      decode(bid, b, i0, npos);
      tc::mbar_wait(&empty[s], parity);
      Ph::geometry(p, smem_raw + (size_t)s * stage_bytes, b, i0, npos, 0, p.L, gtid, gn);
      tc::mbar_arrive(&ready[s]);              // release: this thread's records / fractions are visible to the waiters
      if (++s == n_stages) { s = 0; parity ^= 1; }
    }
  } else {
    // ===== tile copies =====
    const int ctid = (int)threadIdx.x - (n_cons_warps + n_geo_warps) * 32, cn = n_copy_warps * 32;
    int s = 0, prev_s = -1;
    uint32_t parity = 0;
    for (int bid = first; bid < n_batches; bid += step) {
      int b, i0, npos;
      decode(bid, b, i0, npos);
      tc::mbar_wait(&ready[s], parity);
      Ph::gather(p, smem_raw + (size_t)s * stage_bytes, b, i0, npos, 0, p.L, ctid, cn);
      if (prev_s >= 0) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");                     // the previous batch's copies have landed
        __threadfence_block();
        tc::mbar_arrive(&full[prev_s]);
      }
      prev_s = s;
      if (++s == n_stages) { s = 0; parity ^= 1; }
    }
    if (prev_s >= 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __threadfence_block();
      tc::mbar_arrive(&full[prev_s]);
    }
  }
}

// EEM_LOOKUP_PACKED_WS = "<stages>x<CTAs per SM>x<geometry warps>x<copy warps>" selects a configuration of the persistent
// kernel, "0" the one-batch kernel; read per call (tests switch between the kernels inside one process).
inline WsConfig ws_config() {
  // default.  Measured on B200 (MVSEC B = 32, graph of 12 launches, us per launch): one-batch kernel 34.7; two roles
  // (geometry + copies on the same 4 / 8 / 12 warps) 37.1 / 32.3 / 31.6; three roles, 3 stages x 1 CTA per SM, geometry x
  // copy warps 8x10 28.4, 8x5 28.4, 4x5 31.6, 4x10 31.6; 4 stages 37.3 (213 KiB of shared memory leave the 8-byte
  // tile-row copies too little L1); 2 stages x 2 CTAs per SM 42.4.  At 28.4 us the kernel moves its 143 MB of DRAM
  // traffic at 5.0 TB/s (77 % of the measured peak).
  WsConfig c{3, 1, 8, 5};
  if (const char* v = getenv("EEM_LOOKUP_PACKED_WS")) {
    WsConfig e{0, 1, 8, 5};
    const int n = sscanf(v, "%dx%dx%dx%d", &e.stages, &e.per_sm, &e.geo_warps, &e.copy_warps);
    if (n >= 1 && e.stages == 0) return e;
    if (n >= 2 && e.stages >= 2 && e.stages <= 8 && e.per_sm >= 1 && e.per_sm <= 4 && e.geo_warps >= 1 && e.copy_warps >= 1 &&
        e.geo_warps + e.copy_warps <= kWsMaxProdWarps)
      return e;                                // >= 2 stages: full[k-1] is signalled after the copies of batch k are issued
  }
  return c;
}

template <int R>
int launch_lookup_packed(const PackedLookupParams& p, cudaStream_t stream) {
  using S = PackedSmem<R>;
  const size_t stage = (size_t)p.L * S::kPerLevelBytes;
  const int bps = (int)ceil_div(p.H * p.W, S::PB);
  const int64_t n_batches = (int64_t)bps * p.B;
  const WsConfig ws = ws_config();
  if (ws.stages != 0 && n_batches < (int64_t)0x7fffffff) {
    const size_t smem = (size_t)ws.stages * stage + 3 * (size_t)ws.stages * sizeof(uint64_t);
    const int sms = sm_count();
    if (smem <= 227 * 1024 && sms > 0) {
      int cons = p.L * S::kGroups;
      if (cons > S::kThreads / 32) cons = S::kThreads / 32;
      const int threads = (cons + ws.geo_warps + ws.copy_warps) * 32;
      int64_t grid = (int64_t)sms * ws.per_sm;
      if (const char* v = getenv("EEM_LOOKUP_WS_CTAS")) {      // timing experiments: leave SMs to concurrent kernels
        const int cap = atoi(v);
        if (cap >= 1 && cap < grid) grid = cap;
      }
      if (grid > n_batches) grid = n_batches;
      if (p.debug != 0) {                  // phase-ablation build (EEM_LOOKUP_DEBUG, timing experiments only)
        static DynSmemOptIn optin_dbg;
        EEM_CHECK_CUDA(optin_dbg.ensure(corr_lookup_packed_ws_kernel<R, true>, smem));
        corr_lookup_packed_ws_kernel<R, true><<<(unsigned)grid, threads, smem, stream>>>(p, bps, (int)n_batches, (int)stage, ws.stages,
                                                                                       ws.geo_warps, ws.copy_warps);
        return EEM_OK;
      }
      static DynSmemOptIn optin;
      EEM_CHECK_CUDA(optin.ensure(corr_lookup_packed_ws_kernel<R, false>, smem));
      corr_lookup_packed_ws_kernel<R, false><<<(unsigned)grid, threads, smem, stream>>>(p, bps, (int)n_batches, (int)stage, ws.stages,
                                                                                      ws.geo_warps, ws.copy_warps);
      return EEM_OK;
    }
  }
  // EEM_LOOKUP_PACKED_PAD_KB (timing experiments only): extra dynamic shared memory per CTA, i.e. fewer CTAs per SM
  static const size_t pad_bytes = [] {
    const char* v = getenv("EEM_LOOKUP_PACKED_PAD_KB");
    return v != nullptr ? (size_t)atoi(v) * 1024 : (size_t)0;
  }();
  const size_t smem = (size_t)p.L * S::kPerLevelBytes + pad_bytes;
  const int threads = 32 * p.L * S::kGroups;                    // one interpolation task per warp
  dim3 grid((unsigned)bps, (unsigned)p.B, 1);
  const int nthreads = threads < 64 ? 64 : (threads > S::kThreads ? S::kThreads : threads);
  if (p.debug != 0) {                      // phase-ablation build of the kernel (EEM_LOOKUP_DEBUG, timing experiments only)
    static DynSmemOptIn optin_dbg;
    if (smem > 48 * 1024) EEM_CHECK_CUDA(optin_dbg.ensure(corr_lookup_packed_kernel<R, true>, smem));
    corr_lookup_packed_kernel<R, true><<<grid, nthreads, smem, stream>>>(p);
    return EEM_OK;
  }
  static DynSmemOptIn optin;
  if (smem > 48 * 1024) EEM_CHECK_CUDA(optin.ensure(corr_lookup_packed_kernel<R, false>, smem));
  corr_lookup_packed_kernel<R, false><<<grid, nthreads, smem, stream>>>(p);
  return EEM_OK;
}

// packed fp16 level -> the reference's f32 [B*P, h_l*w_l] tensor (lazily materialised `corr_pyramid`)
__global__ void __launch_bounds__(256)
corr_unpack_level_kernel(const uint16_t* __restrict__ packed, int64_t rows, int row_elems, int off, int h, int w, int tiles_x,
                         float* __restrict__ out) {
  const int64_t plane = (int64_t)h * w, total = rows * plane;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / plane;
    const int cell = (int)(i - r * plane);
    const int y = cell / w, x = cell - y * w;
    const __half v = *reinterpret_cast<const __half*>(packed + r * row_elems + off + packed_cell(y, x, tiles_x));
    out[i] = __half2float(v);
  }
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

int eem_corr_lookup_packed(const void* packed, int B, int H, int W, int num_levels, int radius, const float* coords,
                           float* out, eem_stream_t stream_) {
  EEM_CHECK_ARG(packed && coords && out, "eem_corr_lookup_packed: NULL pointer");
  EEM_CHECK_ARG(B > 0 && H > 0 && W > 0, "eem_corr_lookup_packed: sizes must be > 0");
  EEM_CHECK_ARG(num_levels > 0 && num_levels <= kMaxLevels, "eem_corr_lookup_packed: num_levels must be in [1,%d]", kMaxLevels);
  EEM_CHECK_ARG(B <= 65535, "eem_corr_lookup_packed: batch > 65535 not supported in one call");
  EEM_CHECK_ALIGNED(packed, 32);
  const PackedLayout pl = packed_layout(H, W, num_levels);
  if (pl.tx[0] > 255)
    return fail(EEM_ERR_UNSUPPORTED, "eem_corr_lookup_packed: feature maps wider than 1020 are not supported (got %d)", W);
  PackedLookupParams p{};
  p.packed = static_cast<const uint16_t*>(packed);
  for (int l = 0; l < num_levels; ++l) {
    p.h[l] = pl.h[l]; p.w[l] = pl.w[l]; p.tx[l] = pl.tx[l]; p.ty[l] = pl.ty[l]; p.off[l] = pl.off[l];
    if (pl.h[l] * pl.w[l] == 0) p.tx[l] = p.ty[l] = 0;
    p.rcp_w1[l] = (float)(1.0 / (double)(pl.w[l] - 1));      // correctly rounded reciprocal (inf for a 1-wide map)
    p.rcp_h1[l] = (float)(1.0 / (double)(pl.h[l] - 1));
  }
  p.row = pl.row;
  p.B = B; p.H = H; p.W = W; p.L = num_levels;
  p.coords = coords;
  p.out = out;
  if (const char* dbg = getenv("EEM_LOOKUP_DEBUG")) p.debug = atoi(dbg);
  cudaStream_t stream = as_stream(stream_);
  int rc = EEM_OK;
  switch (radius) {
    case 4: rc = launch_lookup_packed<4>(p, stream); break;
    case 3: rc = launch_lookup_packed<3>(p, stream); break;
    case 2: rc = launch_lookup_packed<2>(p, stream); break;
    case 1: rc = launch_lookup_packed<1>(p, stream); break;
    default:
      return fail(EEM_ERR_UNSUPPORTED, "eem_corr_lookup_packed: radius %d not in {1,2,3,4}", radius);
  }
  if (rc != EEM_OK) return rc;
  EEM_CHECK_LAUNCH("corr_lookup_packed_kernel");
  return EEM_OK;
}

int eem_corr_pyramid_unpack(const void* packed, int B, int H, int W, int num_levels, float* const* levels,
                            eem_stream_t stream_) {
  EEM_CHECK_ARG(packed && levels, "eem_corr_pyramid_unpack: NULL pointer");
  EEM_CHECK_ARG(B > 0 && H > 0 && W > 0, "eem_corr_pyramid_unpack: sizes must be > 0");
  EEM_CHECK_ARG(num_levels > 0 && num_levels <= kMaxLevels, "eem_corr_pyramid_unpack: num_levels must be in [1,%d]", kMaxLevels);
  const PackedLayout pl = packed_layout(H, W, num_levels);
  const int64_t rows = (int64_t)B * H * W;
  for (int l = 0; l < num_levels; ++l) {
    const int64_t total = rows * pl.h[l] * pl.w[l];
    if (total == 0) continue;
    EEM_CHECK_ARG(levels[l] != nullptr, "eem_corr_pyramid_unpack: levels[%d] is NULL", l);
    int64_t blocks = ceil_div(total, 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (cap > 0 && blocks > cap) blocks = cap;
    corr_unpack_level_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream_)>>>(static_cast<const uint16_t*>(packed), rows, pl.row,
                                                                                 pl.off[l], pl.h[l], pl.w[l], pl.tx[l], levels[l]);
    EEM_CHECK_LAUNCH("corr_unpack_level_kernel");
  }
  return EEM_OK;
}

}  // extern "C"
