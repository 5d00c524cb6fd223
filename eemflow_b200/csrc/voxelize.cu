// K1 event voxelization + K2 voxel-grid normalisation for sm_100a.
//
// Reference semantics (utils/transformers.py:56-122, EventSequenceToVoxelGrid_Pytorch.__call__):
//   ts   = (nb-1) * (t - t_first) / dT          float64, multiply then divide, dT==0 -> 1.0
//   x,y  = trunc(ev[1]), trunc(ev[2])           int64
//   pol  = float(ev[3]); pol==0 -> -1
//   ti   = floor(ts); dt = float(ts - ti)
//   grid[x + y*W + ti*W*H]     += pol*(1-dt)    if 0 <= ti < nb          ("left" vote)
//   grid[x + y*W + (ti+1)*W*H] += pol*dt        if 0 <= ti, ti+1 < nb    ("right" vote)
//   optional: mean / unbiased std over non-zero voxels, v = (v-mean)/std on them.
//
// Data layout in HBM: events stay in the reference's [N,4] float64 row format (32 B/event); one
// 256-bit LDG per event row, so a warp reads 1 KiB contiguous per instruction.  The grid is
// zero-filled with a memset and updated with fire-and-forget RED.ADD.F32 resolved in L2 (the 55 MB
// HREM grid is L2-resident on B200), so HBM sees 32*N bytes in and 4*nb*H*W bytes out.
#include <cooperative_groups.h>
#include <cstdlib>
#include <mutex>

#include <algorithm>

#include "common.cuh"

namespace eem {
namespace {

constexpr int kVoteThreads = 256;
constexpr int kVoteEventsPerThread = 4;

struct EventRow {
  double t, x, y, p;
};

// Event sources.  Rows: the reference's [N,4] float64 layout, one 32-byte row per load (LDG.E.256 on
// sm_100a), streaming -- rows are read exactly once.  SoA: packed columns (t float64 or raw int64
// nanoseconds, x/y int16, p int8; 13 B/event), the layout of the HREM .npz files before
// get_compressed_events / EventSequence expand them (loader/loader_utils.py:26-37, 352-397).
struct RowSource {
  const double* ev;
  double mul;      // EventSequence's timestamp_multiplier (loader/loader_utils.py:367-368) applied on the fly; 1.0 = none (exact)
  __device__ __forceinline__ EventRow load(int64_t i) const {
    EventRow r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.t), "=d"(r.x), "=d"(r.y), "=d"(r.p)
                 : "l"(ev + 4 * i));
    r.t = __dmul_rn(r.t, mul);
    return r;
  }
  __device__ __forceinline__ double time(int64_t i) const { return __dmul_rn(__ldg(ev + 4 * i), mul); }
};

struct SoaSource {
  const void* t;
  const int16_t* x;
  const int16_t* y;
  const int8_t* p;
  int t_is_ns;
  // raw nanoseconds follow the reference's chain in float64: t*1e-9 (get_compressed_events), then
  // *1e6 (EventSequence timestamp_multiplier); the subtraction of the first stamp happens in the vote
  __device__ __forceinline__ double time(int64_t i) const {
    if (t_is_ns) return __dmul_rn(__dmul_rn((double)__ldg(static_cast<const long long*>(t) + i), 1e-9), 1e6);
    return __ldg(static_cast<const double*>(t) + i);
  }
  __device__ __forceinline__ EventRow load(int64_t i) const {
    EventRow r;
    r.t = time(i);
    r.x = (double)__ldg(x + i);
    r.y = (double)__ldg(y + i);
    r.p = (double)__ldg(p + i);
    return r;
  }
};

struct Vote {
  int64_t idx_left, idx_right;  // flat index inside the window's grid, or -1 when the vote is skipped
  float val_left, val_right;
  bool oob_left, oob_right;     // vote addressed a voxel outside the grid (reference: IndexError)
};

__device__ __forceinline__ Vote make_vote(const EventRow& e, double t_first, double dT, int nb,
                                          int64_t W, int64_t HW, int64_t total) {
  Vote v;
  // float64, multiply first, then divide -- the reference's evaluation order.
  const double ts = __ddiv_rn(__dmul_rn((double)(nb - 1), __dsub_rn(e.t, t_first)), dT);
  const int64_t xi = (int64_t)e.x;  // .long(): truncation toward zero
  const int64_t yi = (int64_t)e.y;
  float pol = (float)e.p;
  if (pol == 0.0f) pol = -1.0f;
  const double tf = floor(ts);
  const float dt = (float)__dsub_rn(ts, tf);
  v.val_left = __fmul_rn(pol, __fsub_rn(1.0f, dt));
  v.val_right = __fmul_rn(pol, dt);
  const bool nonneg = tf >= 0.0;
  const bool ok_left = nonneg && tf < (double)nb;
  const bool ok_right = nonneg && (tf + 1.0) < (double)nb;
  const int64_t ti = ok_left ? (int64_t)tf : 0;
  const int64_t base = xi + yi * W + ti * HW;
  v.idx_left = -1;
  v.idx_right = -1;
  v.oob_left = v.oob_right = false;
  if (ok_left) {
    if (base >= 0 && base < total) v.idx_left = base; else v.oob_left = true;
  }
  if (ok_right) {
    const int64_t r = base + HW;
    if (r >= 0 && r < total) v.idx_right = r; else v.oob_right = true;
  }
  return v;
}

struct WindowTimes {
  double t_first, dT;
};

template <class Src>
__device__ __forceinline__ WindowTimes window_times(const Src& ev, int64_t begin, int64_t end) {
  WindowTimes w;
  w.t_first = ev.time(begin);
  const double t_last = ev.time(end - 1);
  w.dT = __dsub_rn(t_last, w.t_first);
  if (w.dT == 0.0) w.dT = 1.0;
  return w;
}

// Events are time-sorted, so a contiguous chunk votes into only two bin planes.  CTAs are scheduled
// roughly in blockIdx order; mapping consecutive CTAs to chunks that are far apart in time spreads the
// concurrent REDs over all bin planes, which removes most of the same-sector serialisation in L2
// for spatially clustered event data (ncu: 397 us -> see profiles/).  Bijective on [0, n_chunks).
__device__ __forceinline__ int64_t interleaved_chunk(int64_t b, int64_t n_chunks, int lanes) {
  const int64_t per = (n_chunks + lanes - 1) / lanes;       // chunks per time lane
  const int64_t full_lanes = n_chunks - (per - 1) * lanes;   // lanes that own `per` chunks (rest own per-1)
  const int64_t lane = b % lanes, k = b / lanes;
  // lane l owns a contiguous run of chunks; runs of the first `full_lanes` lanes are one longer
  const int64_t start = lane < full_lanes ? lane * per : full_lanes * per + (lane - full_lanes) * (per - 1);
  const int64_t len = lane < full_lanes ? per : per - 1;
  return k < len ? start + k : -1;
}

// ---- atomic mode ------------------------------------------------------------------------------
// grid = (ceil(max_events / (threads*EPT)), n_windows).  Loads of a thread's EPT rows are issued
// back to back before any vote so each thread keeps EPT 32-byte requests in flight.
template <class Src>
__global__ void __launch_bounds__(kVoteThreads)
voxel_vote_atomic_kernel(const Src ev, const int64_t* __restrict__ offsets, int nb,
                         int H, int W, int lanes, float* __restrict__ grid, int64_t* __restrict__ dropped) {
  const int w = blockIdx.y;
  const int64_t begin = offsets[w], end = offsets[w + 1];
  const int64_t n = end - begin;
  const int64_t per_block = kVoteThreads * kVoteEventsPerThread;
  const int64_t chunk = interleaved_chunk(blockIdx.x, (n + per_block - 1) / per_block, lanes);
  if (chunk < 0) return;
  const int64_t first = chunk * per_block + threadIdx.x;
  const WindowTimes wt = window_times(ev, begin, end);
  const int64_t HW = (int64_t)H * W, total = HW * nb;
  float* g = grid + (int64_t)w * total;

  EventRow rows[kVoteEventsPerThread];
#pragma unroll
  for (int k = 0; k < kVoteEventsPerThread; ++k) {
    const int64_t i = first + (int64_t)k * kVoteThreads;
    if (i < n) rows[k] = ev.load(begin + i);
  }
  int ndrop = 0;
#pragma unroll
  for (int k = 0; k < kVoteEventsPerThread; ++k) {
    const int64_t i = first + (int64_t)k * kVoteThreads;
    if (i < n) {
      const Vote v = make_vote(rows[k], wt.t_first, wt.dT, nb, W, HW, total);
      if (v.idx_left >= 0) red_add_f32(g + v.idx_left, v.val_left);
      if (v.idx_right >= 0) red_add_f32(g + v.idx_right, v.val_right);
      ndrop += (int)v.oob_left + (int)v.oob_right;
    }
  }
  if (dropped != nullptr && ndrop != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(dropped), (unsigned long long)ndrop);
}

// ---- atomic mode, pair layout (event-dominated windows) -----------------------------------------
// L2 is the binding unit of the direct kernel above: every 4-byte RED costs a full 32-byte sector
// operation (ncu: lts throughput 75 %, 2 sector ops per event).  When a call has at least ~2 events
// per voxel the votes go to a scratch array S[window][bin k][pixel][2] instead, where slot
// 0 collects the "left" weights of events with floor(ts) == k (-> bin k) and slot 1 their "right"
// weights (-> bin k+1): ONE 8-byte vector RED (REDG.ADD.F32x2) per event, half the L2 sector ops.
// A streaming pass then forms grid[b] = S[b][.][0] + S[b-1][.][1], optionally accumulating the
// normalisation statistics on the fly so the separate stats pass disappears.
__device__ __forceinline__ void red_add_f32x2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

template <class Src>
__global__ void __launch_bounds__(kVoteThreads)
voxel_vote_pair_kernel(const Src ev, const int64_t* __restrict__ offsets, int nb,
                       int H, int W, int lanes, float* __restrict__ scratch, int64_t* __restrict__ dropped) {
  const int w = blockIdx.y;
  const int64_t begin = offsets[w], end = offsets[w + 1];
  const int64_t n = end - begin;
  const int64_t per_block = kVoteThreads * kVoteEventsPerThread;
  const int64_t chunk = interleaved_chunk(blockIdx.x, (n + per_block - 1) / per_block, lanes);
  if (chunk < 0) return;
  const int64_t first = chunk * per_block + threadIdx.x;
  const WindowTimes wt = window_times(ev, begin, end);
  const int64_t HW = (int64_t)H * W, total = HW * nb;
  float* S = scratch + (int64_t)w * total * 2;

  EventRow rows[kVoteEventsPerThread];
#pragma unroll
  for (int k = 0; k < kVoteEventsPerThread; ++k) {
    const int64_t i = first + (int64_t)k * kVoteThreads;
    if (i < n) rows[k] = ev.load(begin + i);
  }
  int ndrop = 0;
#pragma unroll
  for (int k = 0; k < kVoteEventsPerThread; ++k) {
    const int64_t i = first + (int64_t)k * kVoteThreads;
    if (i < n) {
      const Vote v = make_vote(rows[k], wt.t_first, wt.dT, nb, W, HW, total);
      // flat = pix + bin*HW.  Fast path: both votes address the same pixel of adjacent bins.
      if (v.idx_left >= 0 && (v.idx_right == v.idx_left + HW || v.idx_right < 0) && !v.oob_right) {
        red_add_f32x2(S + 2 * v.idx_left, v.val_left, v.idx_right >= 0 ? v.val_right : 0.0f);
      } else {  // a vote was dropped or wrapped differently: address the slots one by one
        if (v.idx_left >= 0) red_add_f32(S + 2 * v.idx_left, v.val_left);
        if (v.idx_right >= HW) red_add_f32(S + 2 * (v.idx_right - HW) + 1, v.val_right);
        else if (v.idx_right >= 0) red_add_f32(S + 2 * v.idx_right, v.val_right);  // no plane below: use slot 0
      }
      ndrop += (int)v.oob_left + (int)v.oob_right;
    }
  }
  if (dropped != nullptr && ndrop != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(dropped), (unsigned long long)ndrop);
}

// ---- deterministic mode -------------------------------------------------------------------------
// Votes are materialised as (global voxel key, value) pairs in the reference's accumulation order
// -- all left votes in event order, then all right votes in event order -- and sorted by key with
// a STABLE least-significant-digit radix sort (8-bit digits).  Each voxel's votes then sit
// contiguously in reference order and are summed sequentially in fp32 by one thread, which
// reproduces the CPU reference's two index_add_ passes bit for bit.
constexpr int kSortWarpsPerBlock = 8;
constexpr int kSortSegment = 2048;  // keys owned by one warp per pass (kept in order)
constexpr int kRadix = 256;

template <class Src>
__global__ void __launch_bounds__(kVoteThreads)
voxel_vote_pairs_kernel(const Src ev, const int64_t* __restrict__ offsets, int nb,
                        int H, int W, int64_t n_total, uint32_t invalid_key,
                        uint32_t* __restrict__ keys, float* __restrict__ vals,
                        int64_t* __restrict__ dropped) {
  const int w = blockIdx.y;
  const int64_t begin = offsets[w], end = offsets[w + 1];
  const int64_t n = end - begin;
  const int64_t i = (int64_t)blockIdx.x * kVoteThreads + threadIdx.x;
  if (i >= n) return;
  const WindowTimes wt = window_times(ev, begin, end);
  const int64_t HW = (int64_t)H * W, total = HW * nb;
  const EventRow e = ev.load(begin + i);
  const Vote v = make_vote(e, wt.t_first, wt.dT, nb, W, HW, total);
  const int64_t g = begin + i;  // global event rank: concatenation order == per-window event order
  const int64_t wbase = (int64_t)w * total;
  keys[g] = v.idx_left >= 0 ? (uint32_t)(wbase + v.idx_left) : invalid_key;
  vals[g] = v.val_left;
  keys[n_total + g] = v.idx_right >= 0 ? (uint32_t)(wbase + v.idx_right) : invalid_key;
  vals[n_total + g] = v.val_right;
  const int ndrop = (int)v.oob_left + (int)v.oob_right;
  if (dropped != nullptr && ndrop != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(dropped), (unsigned long long)ndrop);
}

// counts[d * n_segs + seg] = number of keys of segment `seg` whose current digit is d.
__global__ void __launch_bounds__(kSortWarpsPerBlock * 32)
radix_count_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int64_t n_segs,
                   uint32_t* __restrict__ counts) {
  __shared__ uint32_t hist[kSortWarpsPerBlock][kRadix];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t seg = (int64_t)blockIdx.x * kSortWarpsPerBlock + warp;
  for (int d = lane; d < kRadix; d += 32) hist[warp][d] = 0;
  __syncwarp();
  if (seg < n_segs) {
    const int64_t s0 = seg * kSortSegment;
#pragma unroll 4
    for (int c = 0; c < kSortSegment / 32; ++c) {
      const int64_t i = s0 + c * 32 + lane;
      if (i < n) atomicAdd(&hist[warp][(keys[i] >> shift) & (kRadix - 1)], 1u);
    }
    __syncwarp();
    for (int d = lane; d < kRadix; d += 32) counts[(int64_t)d * n_segs + seg] = hist[warp][d];
  }
}

// Exclusive scan of `counts` (m entries) in three steps: per-block sums, scan of block sums, apply.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;  // entries per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t ws = lane < (blockDim.x >> 5) ? smem_warp[lane] : 0;
    uint32_t winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem_warp[lane] = winc - ws;
    if (lane == 31) smem_warp[32] = winc;
  }
  __syncthreads();
  const uint32_t r = inc - v + smem_warp[warp];
  if (total) *total = smem_warp[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads)
scan_block_sums_kernel(const uint32_t* __restrict__ in, int64_t m, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t sw[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < m) s += in[base + k];
  uint32_t total;
  block_exclusive_scan(s, sw, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
scan_of_block_sums_kernel(uint32_t* __restrict__ block_sums, int64_t nblocks) {
  __shared__ uint32_t sw[33];
  uint32_t carry = 0;
  for (int64_t base = 0; base < nblocks; base += kScanThreads) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? block_sums[i] : 0;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, sw, &total);
    if (i < nblocks) block_sums[i] = ex + carry;
    carry += total;
  }
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(uint32_t* __restrict__ data, int64_t m, const uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t sw[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < m ? data[base + k] : 0;
    s += v[k];
  }
  uint32_t ex = block_exclusive_scan(s, sw, nullptr) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < m) data[base + k] = ex;
    ex += v[k];
  }
}

// Stable scatter: a warp walks its segment in order, 32 keys at a time.  Lanes holding the same
// digit are ranked by lane id (match.any), so equal digits keep their input order.
__global__ void __launch_bounds__(kSortWarpsPerBlock * 32)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const float* __restrict__ vals_in,
                     int64_t n, int shift, int64_t n_segs, const uint32_t* __restrict__ offsets,
                     uint32_t* __restrict__ keys_out, float* __restrict__ vals_out) {
  __shared__ uint32_t offs[kSortWarpsPerBlock][kRadix];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t seg = (int64_t)blockIdx.x * kSortWarpsPerBlock + warp;
  if (seg >= n_segs) return;
  for (int d = lane; d < kRadix; d += 32) offs[warp][d] = offsets[(int64_t)d * n_segs + seg];
  __syncwarp();
  const int64_t s0 = seg * kSortSegment;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int c = 0; c < kSortSegment / 32; ++c) {
    const int64_t i = s0 + c * 32 + lane;
    if (s0 + c * 32 >= n) break;
    const bool valid = i < n;
    const uint32_t key = valid ? keys_in[i] : 0u;
    const float val = valid ? vals_in[i] : 0.0f;
    // Lanes past the end get a private pseudo-digit so they never share a peer group.
    const uint32_t d = valid ? ((key >> shift) & (kRadix - 1)) : (uint32_t)(kRadix + lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank = __popc(peers & lt_mask);
    uint32_t base = 0;
    if (valid) base = offs[warp][d];
    __syncwarp();
    if (valid && rank == 0) offs[warp][d] = base + __popc(peers);
    __syncwarp();
    if (valid) {
      keys_out[base + rank] = key;
      vals_out[base + rank] = val;
    }
  }
}

// One thread per segment head; sequential fp32 accumulation in sorted (= reference) order.
__global__ void __launch_bounds__(256)
segmented_sum_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ vals, int64_t n,
                     uint32_t invalid_key, float* __restrict__ grid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = keys[i];
  if (k == invalid_key) return;
  if (i > 0 && keys[i - 1] == k) return;
  float acc = 0.0f;
  int64_t j = i;
  do {
    acc = __fadd_rn(acc, vals[j]);
    ++j;
  } while (j < n && keys[j] == k);
  grid[k] = acc;
}


// ---- atomic mode, bin-interleaved layout (many bins per pixel) -----------------------------------------
// The two votes of an event go to bins k and k+1 of the SAME pixel.  In a scratch array T[window][pixel][bin]
// (bins innermost, padded to an even count) they are neighbours, so whenever k is even ONE 8-byte vector RED
// (REDG.ADD.F32x2) carries both: 1.5 instead of 2 L2 sector operations per event on average, and L2 atomic
// throughput is what binds the vote pass.  The statistics pass reads T as it is (padding stays zero) and the
// normalisation pass transposes T into the reference's [bin][y][x] layout on its way out, so compared with the
// direct path no pass is added; T costs nbp/nb more bytes (HREM: 15 -> 16 bins).  (Experiment, off by default.)
template <class Src>
__global__ void __launch_bounds__(kVoteThreads)
voxel_vote_interleaved_kernel(const Src ev, const int64_t* __restrict__ offsets, int nb, int nbp, int H, int W, int lanes,
                              float* __restrict__ T, int64_t* __restrict__ dropped) {
  const int w = blockIdx.y;
  const int64_t begin = offsets[w], end = offsets[w + 1];
  const int64_t n = end - begin;
  const int64_t per_block = kVoteThreads * kVoteEventsPerThread;
  const int64_t chunk = interleaved_chunk(blockIdx.x, (n + per_block - 1) / per_block, lanes);
  if (chunk < 0) return;
  const int64_t first = chunk * per_block + threadIdx.x;
  const WindowTimes wt = window_times(ev, begin, end);
  const int64_t HW = (int64_t)H * W, total = HW * nb;
  const unsigned hw = (unsigned)HW;                      // total < 2^31 on this path (checked by the host)
  float* t = T + (int64_t)w * HW * nbp;

  EventRow rows[kVoteEventsPerThread];
#pragma unroll
  for (int k = 0; k < kVoteEventsPerThread; ++k) {
    const int64_t i = first + (int64_t)k * kVoteThreads;
    if (i < n) rows[k] = ev.load(begin + i);
  }
  int ndrop = 0;
#pragma unroll
  for (int k = 0; k < kVoteEventsPerThread; ++k) {
    const int64_t i = first + (int64_t)k * kVoteThreads;
    if (i < n) {
      const Vote v = make_vote(rows[k], wt.t_first, wt.dT, nb, W, HW, total);
      // flat reference index -> (bin, pixel); the reference's silent wrap of x >= W into the next row / bin is kept
      unsigned bl = 0, pl = 0, br = 0, pr = 0;
      if (v.idx_left >= 0) { bl = (unsigned)v.idx_left / hw; pl = (unsigned)v.idx_left - bl * hw; }
      if (v.idx_right >= 0) { br = (unsigned)v.idx_right / hw; pr = (unsigned)v.idx_right - br * hw; }
      if (v.idx_left >= 0 && v.idx_right >= 0 && pl == pr && br == bl + 1 && (bl & 1u) == 0) {
        red_add_f32x2(t + (int64_t)pl * nbp + bl, v.val_left, v.val_right);
      } else {
        if (v.idx_left >= 0) red_add_f32(t + (int64_t)pl * nbp + bl, v.val_left);
        if (v.idx_right >= 0) red_add_f32(t + (int64_t)pr * nbp + br, v.val_right);
      }
      ndrop += (int)v.oob_left + (int)v.oob_right;
    }
  }
  if (dropped != nullptr && ndrop != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(dropped), (unsigned long long)ndrop);
}

// ---- K2 normalisation -----------------------------------------------------------------------
constexpr int kStatThreads = 256;
constexpr int kStatBlocksPerWindowMax = 2048;

struct StatPartial {
  double count, sum, sumsq;
};

__device__ __forceinline__ void accum_stat(float v, double& c, double& s, double& q) {
  if (v != 0.0f) {
    c += 1.0;
    s += (double)v;
    q += (double)v * (double)v;
  }
}

// Block-reduce (count, sum, sum of squares), publish the block partial, and let the last block of the
// window combine all partials in block order (ticket counter) so mean/std do not depend on scheduling.
__device__ __forceinline__ void finish_stats_at(double c, double s, double q, int w, unsigned int part_idx, unsigned int part_cnt,
                                                StatPartial* __restrict__ partials, unsigned int* __restrict__ tickets,
                                                float* __restrict__ mean_std, double* __restrict__ stats_out);

__device__ __forceinline__ void finish_stats(double c, double s, double q, int w, StatPartial* __restrict__ partials,
                                             unsigned int* __restrict__ tickets, float* __restrict__ mean_std,
                                             double* __restrict__ stats_out) {
  finish_stats_at(c, s, q, w, blockIdx.x, gridDim.x, partials, tickets, mean_std, stats_out);
}

// part_idx / part_cnt: this block's slot among the window's `part_cnt` partials (every block of the window calls this)
__device__ __forceinline__ void finish_stats_at(double c, double s, double q, int w, unsigned int part_idx, unsigned int part_cnt,
                                                StatPartial* __restrict__ partials, unsigned int* __restrict__ tickets,
                                                float* __restrict__ mean_std, double* __restrict__ stats_out) {
  __shared__ double red[3][32];
  __shared__ bool is_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    red[0][warp] = c;
    red[1][warp] = s;
    red[2][warp] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bc = 0, bs = 0, bq = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      bc += red[0][k];
      bs += red[1][k];
      bq += red[2][k];
    }
    StatPartial p{bc, bs, bq};
    partials[(int64_t)w * part_cnt + part_idx] = p;
    __threadfence();
    const unsigned int t = atomicAdd(&tickets[w], 1u);
    is_last = (t == part_cnt - 1);
  }
  __syncthreads();
  if (is_last) {
    // The last block combines the partials: thread t sums partials t, t+T, ... in index order, then a
    // fixed shuffle/shared-memory tree -- parallel, and still independent of block scheduling.
    __threadfence();
    double pc = 0, ps = 0, pq = 0;
    for (unsigned int k = threadIdx.x; k < part_cnt; k += blockDim.x) {
      const StatPartial p = partials[(int64_t)w * part_cnt + k];
      pc += p.count;
      ps += p.sum;
      pq += p.sumsq;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      pc += __shfl_xor_sync(0xffffffffu, pc, o);
      ps += __shfl_xor_sync(0xffffffffu, ps, o);
      pq += __shfl_xor_sync(0xffffffffu, pq, o);
    }
    __syncthreads();   // red[] is free again
    if (lane == 0) {
      red[0][warp] = pc;
      red[1][warp] = ps;
      red[2][warp] = pq;
    }
    __syncthreads();
  }
  if (is_last && threadIdx.x == 0) {
    double tc = 0, ts = 0, tq = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      tc += red[0][k];
      ts += red[1][k];
      tq += red[2][k];
    }
    // mean and unbiased std, rounded to fp32 like the 0-dim fp32 tensors of the reference.
    float mean = 0.0f, sd = 0.0f;
    if (tc > 0) {
      const double m = ts / tc;
      mean = (float)m;
      if (tc > 1) {
        double var = (tq - ts * m) / (tc - 1.0);
        if (var < 0) var = 0;
        sd = (float)sqrt(var);
      } else {
        sd = __int_as_float(0x7fc00000);  // torch: std of one element is NaN -> "v - mean" branch
      }
    }
    mean_std[2 * w + 0] = mean;
    mean_std[2 * w + 1] = sd;
    if (stats_out) {
      stats_out[3 * w + 0] = tc;
      stats_out[3 * w + 1] = (double)mean;
      stats_out[3 * w + 2] = (double)sd;
    }
    tickets[w] = 0;  // leave the workspace reusable without a memset
  }
}

// =================================================================================================================
// Tile-binned voxelization (the default path of both modes): no atomics on global memory, every voxel written once.
//
//   pass 1  voxel_bin_kernel    a CTA takes a chunk of 8192 consecutive events of one window, turns each event into an
//                               8-byte record (pixel inside its tile, bin, polarity sign, which of the two votes exist,
//                               dt as fp32 -- the reference's float(ts - floor(ts))), and counting-sorts the chunk BY
//                               TILE (64 x 32 pixels) in shared memory.  The sort is STABLE (per-warp histograms, ranks
//                               from match.any in lane order, warps own consecutive event ranges), so inside a tile's
//                               run the records keep their event order.  The sorted chunk is written back contiguously
//                               (fully coalesced) next to a per-chunk table of tile offsets; HBM sees 32 B/event in and
//                               8 B/event out.
//   pass 2  voxel_tile_kernel   a CTA owns one tile and builds its bin planes one after the other in shared memory from
//                               the tile's runs of all chunks (a per-bin chunk range recorded by pass 1 keeps it to the
//                               chunks that can contain the bin: events are time-sorted, so that is 1/nb of them; for
//                               unsorted input the range simply grows and the result is still right), then writes each
//                               plane to the grid once with plain stores -- the grid needs no memset -- while
//                               accumulating the normalisation statistics of K2 on the fly.
//       order-free mode         every record is visited once: its left vote goes to the plane being finished, its
//                               right vote to the next one (two planes alternate); warps split the records and add
//                               with shared-memory atomics.
//       deterministic mode      plane b = all LEFT votes of the records with floor(ts) == b in event order, then all
//                               RIGHT votes of the records with floor(ts) == b-1 in event order -- exactly the order of
//                               the reference's two index_add_ passes (utils/transformers.py:98-110).  Every warp scans
//                               the records in order but applies only the pixels it owns (pixel rows interleaved over
//                               the warps); duplicates of a pixel inside a 32-record batch are applied in lane order.
// Polarity other than +-1 (after the reference's 0 -> -1) cannot be packed into the record's sign bit: such an event
// is flagged and its polarity is kept in a side array indexed like the records (never touched otherwise).
// =================================================================================================================
#ifndef EEM_TILE_W_SHIFT
#define EEM_TILE_W_SHIFT 6
#endif
constexpr int kTileWShift = EEM_TILE_W_SHIFT, kTileHShift = 5;  // 64 x 32 pixels
constexpr int kTileW = 1 << kTileWShift, kTileH = 1 << kTileHShift, kTilePix = kTileW * kTileH;
constexpr int kBinThreads = 256, kBinWarps = kBinThreads / 32;
constexpr int kBinChunk = 4096, kBinPerWarp = kBinChunk / kBinWarps, kBinIters = kBinPerWarp / 32;
constexpr int kMaxTiles = 2048;                                 // per window (8 tiles per thread in the chunk scan)
constexpr int kTilesPerThread = kMaxTiles / kBinThreads;
constexpr int kTileThreads = 256, kTileWarps = kTileThreads / 32;
constexpr int kMaxStatParts = 1 << 16;                          // tiles * bins per window the fused statistics can hold

// record meta: [0,P) pixel in tile (P = 11 bits for 64 x 32) | has-left (-> plane bin) | has-right (-> plane bin+1) |
// swap: a lone RIGHT vote that lands in plane `bin` itself (an off-sensor event whose left vote fell outside the grid;
// it is added in the right-vote pass of that plane, in event order) | polarity is negative | polarity in the side
// array | [P+5, 32) bin
constexpr int kPixBits = kTileWShift + kTileHShift, kBinShift = kPixBits + 5;
constexpr uint32_t kRecL = 1u << kPixBits, kRecR = 2u << kPixBits, kRecSwap = 4u << kPixBits, kRecNeg = 8u << kPixBits,
                   kRecSide = 16u << kPixBits;
constexpr int kMaxBinsTiled = 1 << (32 - kBinShift - 1);

struct TilePlan {
  int tiles_x, tiles_y, n_tiles;
  int chunks_max;            // chunks per window (table stride)
  int table_stride;          // uint16 entries per (window, chunk): n_tiles + 1
};

__host__ __device__ inline TilePlan tile_plan(int H, int W, int64_t max_events_per_window) {
  TilePlan tp;
  tp.tiles_x = (W + kTileW - 1) >> kTileWShift;
  tp.tiles_y = (H + kTileH - 1) >> kTileHShift;
  tp.n_tiles = tp.tiles_x * tp.tiles_y;
  tp.chunks_max = (int)((max_events_per_window + kBinChunk - 1) / kBinChunk);
  if (tp.chunks_max < 1) tp.chunks_max = 1;
  tp.table_stride = tp.n_tiles + 1;
  return tp;
}

struct TileWorkspace {
  size_t recs, side, table, range, swap_flags, total;
};

// layout for `n_total` events in `n_windows` windows; chunks_bound >= chunks per window
inline TileWorkspace tile_workspace(int64_t n_total, int n_windows, int num_bins, int n_tiles, int64_t chunks_bound) {
  TileWorkspace L{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t n = (size_t)(n_total > 0 ? n_total : 1);
  L.recs = take(n * 8);
  L.side = take(n * 4);
  L.table = take((size_t)n_windows * (size_t)chunks_bound * (size_t)(n_tiles + 1) * sizeof(uint16_t));
  L.range = take((size_t)n_windows * (size_t)num_bins * 2 * sizeof(int));
  L.swap_flags = take((size_t)n_windows * sizeof(int));
  L.total = off;
  return L;
}

// Correctly rounded double division a / b through b's reciprocal: q0 = RN(a*y), r = a - q0*b (exact, one FMA),
// q = RN(q0 + r*y) with y = RN(1/b) (Markstein).  Same bits as __ddiv_rn for the finite, normal-range stamps of an
// event window, at 3 DFMA-class instructions instead of the ~25 of the IEEE division sequence; the deterministic
// mode's bit-exactness tests (10 M / 40 M events against the CPU reference restatement) pin it.
__device__ __forceinline__ double div_by_rcp(double a, double b, double rcp_b) {
  const double q0 = __dmul_rn(a, rcp_b);
  const double r = __fma_rn(-q0, b, a);
  return __fma_rn(r, rcp_b, q0);
}

struct TileRec {
  uint32_t meta;
  float dt;
  int tile;        // < 0: the event casts no vote
  int ndrop;
};

// Events that do not lie on the sensor (x or y outside [0,W) x [0,H)): the reference votes whatever flat index
// x + y*W + bin*W*H comes out as long as it is inside the grid (e.g. its silent row wrap for x >= W) and raises
// otherwise.  Reproduce the former, count the latter.  Rare: kept out of line so the hot loop stays small.
__device__ __noinline__ TileRec off_sensor_rec(double ex, double ey, int ti, bool ok_right, uint32_t flags, float dt, int nb,
                                               int H, int W, int tiles_x) {
  TileRec rec;
  rec.meta = 0; rec.dt = dt; rec.tile = -1;
  const int64_t xi = (int64_t)ex, yi = (int64_t)ey;              // .long(): truncation toward zero
  const int64_t HW = (int64_t)H * W, total = HW * nb;
  const int64_t fl = xi + yi * (int64_t)W + (int64_t)ti * HW, fr = fl + HW;
  const bool in_l = fl >= 0 && fl < total, in_r = ok_right && fr >= 0 && fr < total;
  rec.ndrop = (int)!in_l + (int)(ok_right && !in_r);
  if (!in_l && !in_r) return rec;
  const int64_t f = in_l ? fl : fr;
  const int bin = (int)(f / HW);
  const int64_t pix = f - (int64_t)bin * HW;
  const int y = (int)(pix / W), x = (int)(pix - (int64_t)y * W);
  flags |= in_l ? (kRecL | (in_r ? kRecR : 0u)) : kRecSwap;
  rec.tile = (y >> kTileHShift) * tiles_x + (x >> kTileWShift);
  rec.meta = (uint32_t)(((y & (kTileH - 1)) << kTileWShift) | (x & (kTileW - 1))) | flags | ((uint32_t)bin << kBinShift);
  return rec;
}

__device__ __forceinline__ TileRec make_tile_rec(const EventRow& e, double t_first, double dT, double rcp_dT, int nb,
                                                  int H, int W, int tiles_x) {
  TileRec rec;
  rec.meta = 0; rec.dt = 0.f; rec.tile = -1; rec.ndrop = 0;
  const double ts = div_by_rcp(__dmul_rn((double)(nb - 1), __dsub_rn(e.t, t_first)), dT, rcp_dT);
  const double tf = floor(ts);
  if (!(tf >= 0.0 && tf < (double)nb)) return rec;              // neither vote exists (utils/transformers.py:90-91, 104-105)
  const int ti = (int)tf;
  const bool ok_right = ti + 1 < nb;
  rec.dt = (float)__dsub_rn(ts, tf);
  const float pol = (float)e.p;
  // +1 -> 0, -1 and 0 (the reference maps 0 to -1) -> kRecNeg, anything else -> side array
  const uint32_t flags = pol == 1.0f ? 0u : ((pol == -1.0f || pol == 0.0f) ? kRecNeg : kRecSide);
  const int x = __double2int_rz(e.x), y = __double2int_rz(e.y);  // saturating: far-out values fail the range test below
  if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) {  // on the sensor: both votes address the grid
    rec.tile = (y >> kTileHShift) * tiles_x + (x >> kTileWShift);
    rec.meta = (uint32_t)(((y & (kTileH - 1)) << kTileWShift) | (x & (kTileW - 1))) | flags | kRecL | (ok_right ? kRecR : 0u) |
               ((uint32_t)ti << kBinShift);
    return rec;
  }
  return off_sensor_rec(e.x, e.y, ti, ok_right, flags, rec.dt, nb, H, W, tiles_x);
}

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

template <class Src>
__global__ void __launch_bounds__(kBinThreads, 2)
voxel_bin_kernel(const Src ev, const int64_t* __restrict__ offsets, int nb, int H, int W, const TilePlan tp,
                 uint2* __restrict__ recs, float* __restrict__ side, uint16_t* __restrict__ table,
                 int* __restrict__ range, int* __restrict__ swap_flags, int64_t* __restrict__ dropped) {
  extern __shared__ __align__(16) unsigned char bin_smem[];
  const int w = blockIdx.y, chunk = blockIdx.x;
  const int64_t begin = offsets[w], end = offsets[w + 1];
  const int64_t n = end - begin;
  const int64_t first = (int64_t)chunk * kBinChunk;
  if (first >= n) return;
  const int64_t rel = begin - offsets[0];          // records are indexed relative to the call's first window
  const int n_tiles = tp.n_tiles;
  const int hist_stride = (n_tiles + 1) | 1;                         // odd: the warps' rows start in different banks
  uint2* stage = reinterpret_cast<uint2*>(bin_smem);                                   // [kBinChunk]
  uint16_t* hist = reinterpret_cast<uint16_t*>(bin_smem + (size_t)kBinChunk * 8);      // [kBinWarps][hist_stride]
  uint16_t* tile_off = hist + (size_t)kBinWarps * hist_stride;                         // [n_tiles + 1]
  __shared__ uint32_t scan_ws[kBinWarps + 1];
  __shared__ int s_tmin, s_tmax, s_side;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    uint32_t* h32 = reinterpret_cast<uint32_t*>(hist);               // hist starts 8-byte aligned (after the stage)
    for (int i = threadIdx.x; i < (kBinWarps * hist_stride + 1) / 2; i += kBinThreads) h32[i] = 0;
  }
  if (threadIdx.x == 0) { s_tmin = 0x7fffffff; s_tmax = -1; s_side = 0; }
  __syncthreads();

  const WindowTimes wt = window_times(ev, begin, end);
  const double rcp_dT = __drcp_rn(wt.dT);
  uint16_t* my_hist = hist + (size_t)warp * hist_stride;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int left = (int)min((int64_t)kBinChunk, n - first);          // events of this chunk

  // A) records + stable ranks.  Warp `warp` owns events [warp*512, +512) of the chunk, 32 consecutive ones per iteration.
  uint32_t r_meta[kBinIters];
  float r_dt[kBinIters];
  uint32_t r_key[kBinIters];          // tile << 16 | rank among the warp's earlier events of that tile; 0xffffffff: no record
  int ndrop = 0, tmin = 0x7fffffff, tmax = -1;
  const int wfirst = warp * kBinPerWarp;
  const int64_t ev0 = begin + first;
#pragma unroll
  for (int it0 = 0; it0 < kBinIters; it0 += 4) {
    EventRow rows[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = wfirst + (it0 + k) * 32 + lane;
      if (i < left) rows[k] = ev.load(ev0 + i);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int it = it0 + k;
      const int i = wfirst + it * 32 + lane;
      TileRec rec;
      rec.tile = -1; rec.meta = 0; rec.dt = 0.f; rec.ndrop = 0;
      if (i < left) rec = make_tile_rec(rows[k], wt.t_first, wt.dT, rcp_dT, nb, H, W, tp.tiles_x);
      ndrop += rec.ndrop;
      const bool has = rec.tile >= 0;
      // lanes without a record get a private pseudo-tile so they never share a peer group
      const uint32_t tkey = has ? (uint32_t)rec.tile : (uint32_t)(kMaxTiles + lane);
      const uint32_t peers = __match_any_sync(0xffffffffu, tkey);
      const uint32_t before = peers & lt_mask;
      uint32_t key = 0xffffffffu;
      if (has) key = (tkey << 16) | (my_hist[tkey] + __popc(before));
      __syncwarp();
      if (has && before == 0) my_hist[tkey] += (uint16_t)__popc(peers);
      __syncwarp();
      const int bin = has ? (int)(rec.meta >> kBinShift) : -1;     // pass 2 looks records up by THEIR bin (the right vote rides along)
      tmax = max(tmax, bin);
      tmin = min(tmin, has ? bin : 0x7fffffff);
      r_meta[it] = rec.meta;
      r_dt[it] = rec.dt;
      r_key[it] = key;
    }
  }
  {
    uint32_t any_side = 0;
#pragma unroll
    for (int it = 0; it < kBinIters; ++it) any_side |= (r_key[it] != 0xffffffffu) ? (r_meta[it] & (kRecSide | kRecSwap)) : 0u;
    const bool any_swap = __any_sync(0xffffffffu, (any_side & kRecSwap) != 0);
    any_side = __any_sync(0xffffffffu, (any_side & kRecSide) != 0);
    if (any_swap && lane == 0) atomicOr(&swap_flags[w], 1);     // rare: pass 2 widens the right-vote scan of this window
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tmin = min(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
      tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      ndrop += __shfl_xor_sync(0xffffffffu, ndrop, o);
    }
    if (lane == 0) {
      if (tmax >= 0) { atomicMin(&s_tmin, tmin); atomicMax(&s_tmax, tmax); }
      if (any_side) atomicOr(&s_side, 1);                  // rare: polarity other than +-1 -> side array (phase C)
      if (dropped != nullptr && ndrop != 0) atomicAdd(reinterpret_cast<unsigned long long*>(dropped), (unsigned long long)ndrop);
    }
  }
  __syncthreads();

  // B) per tile: exclusive prefix over the warps (in place), tile totals, exclusive scan over the tiles
  {
    const int t0 = threadIdx.x * kTilesPerThread;          // tiles [t0, t0 + 8): n_tiles <= 8 * kBinThreads
    uint32_t tot[kTilesPerThread];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kTilesPerThread; ++k) {
      tot[k] = 0;
      if (t0 + k < n_tiles) {
        uint32_t run = 0;
#pragma unroll
        for (int wp = 0; wp < kBinWarps; ++wp) {
          const uint32_t c = hist[(size_t)wp * hist_stride + t0 + k];
          hist[(size_t)wp * hist_stride + t0 + k] = (uint16_t)run;
          run += c;
        }
        tot[k] = run;
      }
      mine += tot[k];
    }
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) scan_ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t ws = lane < kBinWarps ? scan_ws[lane] : 0;
      uint32_t winc = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += v;
      }
      if (lane < kBinWarps) scan_ws[lane] = winc - ws;
      if (lane == kBinWarps - 1) scan_ws[kBinWarps] = winc;
    }
    __syncthreads();
    uint32_t ex = inc - mine + scan_ws[warp];
    uint16_t* trow = table + ((size_t)w * tp.chunks_max + chunk) * tp.table_stride;
#pragma unroll
    for (int k = 0; k < kTilesPerThread; ++k) {
      if (t0 + k < n_tiles) {
        tile_off[t0 + k] = (uint16_t)ex;
        trow[t0 + k] = (uint16_t)ex;
        ex += tot[k];
      }
    }
    if (threadIdx.x == 0) {
      const uint32_t total = scan_ws[kBinWarps];
      tile_off[n_tiles] = (uint16_t)total;       // total <= kBinChunk fits
      trow[n_tiles] = (uint16_t)total;
      // the bins this chunk holds records of: widen the chunk ranges of those bins
      if (s_tmax >= 0)
        for (int b = s_tmin; b <= s_tmax && b < nb; ++b) {
          atomicMin(&range[(size_t)w * nb + b], chunk);                                   // lo[w][b]
          atomicMax(&range[(size_t)gridDim.y * nb + (size_t)w * nb + b], chunk);          // hi[w][b]
        }
    }
  }
  __syncthreads();

  // C) scatter into the sorted order (shared memory), then stream the chunk out
#pragma unroll
  for (int it = 0; it < kBinIters; ++it) {
    const uint32_t key = r_key[it];
    if (key != 0xffffffffu) {
      const uint32_t tile = key >> 16;
      stage[(uint32_t)tile_off[tile] + (uint32_t)my_hist[tile] + (key & 0xffffu)] = make_uint2(r_meta[it], __float_as_uint(r_dt[it]));
    }
  }
  if (s_side != 0) {
    // rare path: re-read the polarity of the flagged events and park it next to the record's final position
    for (int it = 0; it < kBinIters; ++it) {
      const uint32_t key = r_key[it];
      if (key != 0xffffffffu && (r_meta[it] & kRecSide)) {
        const uint32_t tile = key >> 16;
        const uint32_t pos = (uint32_t)tile_off[tile] + (uint32_t)my_hist[tile] + (key & 0xffffu);
        side[rel + first + pos] = (float)ev.load(ev0 + wfirst + it * 32 + lane).p;
      }
    }
  }
  __syncthreads();
  const int total = (int)tile_off[n_tiles];
  uint2* out = recs + rel + first;
  // the records are read again by pass 2: ask L2 to keep them (the event rows above stream through once)
  const uint64_t keep = l2_policy_evict_last();
  if (((reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(stage);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    for (int i = threadIdx.x; i < total / 2; i += kBinThreads) {
      const uint4 v = s4[i];
      asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(o4 + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(keep) : "memory");
    }
    if ((total & 1) && threadIdx.x == 0) out[total - 1] = stage[total - 1];
  } else {
    for (int i = threadIdx.x; i < total; i += kBinThreads) out[i] = stage[i];
  }
}

// pass 2 ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rec_polarity(uint32_t meta, const float* __restrict__ side_w, int64_t idx) {
  if (meta & kRecSide) return side_w[idx];
  return (meta & kRecNeg) ? -1.0f : 1.0f;
}

// One CTA builds ONE bin plane of ONE tile: grid = (tiles, bins, windows).  Plane b of a tile is
//   all LEFT votes of the tile's records with floor(ts) == b, in event order, then
//   all RIGHT votes of its records with floor(ts) == b-1, in event order
// -- the order of the reference's two index_add_ passes (utils/transformers.py:98-110) -- so a record is read by two
// CTAs (out of L2), every CTA owns its plane outright, and the work splits into tiles x bins even-sized pieces (a
// hot tile of a clustered stream no longer serialises on one CTA).
// The tile's runs inside the bin's chunk range are flattened by a block prefix sum; 256 consecutive records are
// handled at a time, thread = record.
//   kOrdered   exact mode: threads claim their pixel with atomicMin(thread id) and the lowest claimant of every pixel
//              applies its vote, round after round, so duplicates of a pixel inside a batch are added in event order.
//              Batches, chunk rounds and the two phases follow each other in order: the sums are bit-identical to the
//              reference's.
//   !kOrdered  shared-memory float atomics (order-free).
template <bool kOrdered>
__global__ void __launch_bounds__(kTileThreads)
voxel_plane_kernel(const int64_t* __restrict__ offsets, int nb, int H, int W, const TilePlan tp,
                   const uint2* __restrict__ recs, const float* __restrict__ side, const uint16_t* __restrict__ table,
                   const int* __restrict__ range, const int* __restrict__ swap_flags, float* __restrict__ grid, int with_stats,
                   StatPartial* __restrict__ partials, unsigned int* __restrict__ tickets, float* __restrict__ mean_std,
                   double* __restrict__ stats_out) {
  __shared__ __align__(16) float plane[kTilePix];
  __shared__ uint32_t claim[kOrdered ? kTilePix : 1];
  __shared__ int run_ex[kTileThreads + 1];
  __shared__ int run_base[kTileThreads];
  __shared__ int warp_tot[kTileWarps + 1];
  const int tile = blockIdx.x, b = blockIdx.y, w = blockIdx.z;
  const int n_windows = gridDim.z;
  const int64_t rel = offsets[w] - offsets[0];
  const uint2* recs_w = recs + rel;
  const float* side_w = side + rel;
  const uint16_t* table_w = table + (size_t)w * tp.chunks_max * tp.table_stride + tile;
  const int* range_lo = range + (size_t)w * nb;
  const int* range_hi = range + (size_t)n_windows * nb + (size_t)w * nb;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < kTilePix; i += kTileThreads) {
    plane[i] = 0.0f;
    if (kOrdered) claim[i] = 0xffffffffu;
  }
  __syncthreads();

  const bool has_swap = swap_flags[w] != 0;
  for (int phase = 0; phase < 2; ++phase) {
    // phase 0: LEFT votes of the records with floor(ts) == b.  phase 1: RIGHT votes of the records with floor(ts) == b-1
    // plus, in windows that hold off-sensor events, the lone right votes filed under bin b itself ("swap" records) --
    // all in event order, because the chunks (and the runs inside them) are visited in order.
    int c_lo, c_hi;
    if (phase == 0) {
      c_lo = range_lo[b];
      c_hi = range_hi[b];
    } else {
      c_lo = 0x7fffffff;
      c_hi = -1;
      if (b > 0) { c_lo = range_lo[b - 1]; c_hi = range_hi[b - 1]; }
      if (has_swap) { c_lo = min(c_lo, range_lo[b]); c_hi = max(c_hi, range_hi[b]); }
    }
    for (int cc = c_lo; cc <= c_hi; cc += kTileThreads) {     // rounds of 256 chunks (one round for time-sorted input)
      const int c = cc + tid;
      int start = 0, len = 0;
      if (c <= c_hi) {
        const uint16_t* t2 = table_w + (size_t)c * tp.table_stride;
        start = t2[0];
        len = (int)t2[1] - start;
      }
      int inc = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (lane == 31) warp_tot[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        const int ws = lane < kTileWarps ? warp_tot[lane] : 0;
        int winc = ws;
#pragma unroll
        for (int o = 1; o < kTileWarps; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, winc, o);
          if (lane >= o) winc += v;
        }
        if (lane < kTileWarps) warp_tot[lane] = winc - ws;
        if (lane == kTileWarps - 1) warp_tot[kTileWarps] = winc;
      }
      __syncthreads();
      run_ex[tid] = inc - len + warp_tot[warp];
      run_base[tid] = c * kBinChunk + start;
      const int total = warp_tot[kTileWarps];
      if (tid == 0) run_ex[kTileThreads] = total;
      __syncthreads();

      for (int i0 = 0; i0 < total; i0 += kTileThreads) {
        const int i = i0 + tid;
        bool sel = false;
        int pix = 0;
        float v = 0.0f;
        if (i < total) {
          int r = 0;                                   // the last run whose exclusive prefix is <= i
#pragma unroll
          for (int step = kTileThreads / 2; step > 0; step >>= 1)
            if (run_ex[r + step] <= i) r += step;
          const int64_t idx = (int64_t)run_base[r] + (i - run_ex[r]);
          const uint2 rec = __ldg(recs_w + idx);
          const uint32_t meta = rec.x;
          const int rbin = (int)(meta >> kBinShift);
          sel = phase == 0 ? (rbin == b && (meta & kRecL) != 0)
                           : ((rbin == b - 1 && (meta & kRecR) != 0) || (rbin == b && (meta & kRecSwap) != 0));
          if (sel) {
            pix = (int)(meta & (kTilePix - 1));
            const float pol = rec_polarity(meta, side_w, idx);
            const float dt = __uint_as_float(rec.y);
            v = phase == 1 ? __fmul_rn(pol, dt) : __fmul_rn(pol, __fsub_rn(1.0f, dt));
          }
        }
        if (kOrdered) {
          bool pending = sel;
          while (true) {
            if (pending) atomicMin(&claim[pix], (uint32_t)tid);
            __syncthreads();
            // (a loser may read the claim word while the winner resets it: both accesses are volatile / atomic, and
            // either value it can see -- the winner's id or "free" -- differs from its own id)
            if (pending && *reinterpret_cast<volatile uint32_t*>(&claim[pix]) == (uint32_t)tid) {
              plane[pix] = __fadd_rn(plane[pix], v);
              atomicExch(&claim[pix], 0xffffffffu);
              pending = false;
            }
            if (!__syncthreads_or((int)pending)) break;
          }
        } else if (sel) {
          atomicAdd(&plane[pix], v);
        }
      }
      __syncthreads();        // run_ex / run_base are rewritten by the next round
    }
  }
  __syncthreads();

  // the plane is complete: write it to the grid (every voxel of the tile once), accumulate the statistics
  const int ty = tile / tp.tiles_x, tx = tile - ty * tp.tiles_x;
  const int x0 = tx << kTileWShift, y0 = ty << kTileHShift;
  const int64_t HW = (int64_t)H * W;
  float* gp = grid + ((int64_t)w * nb + b) * HW;
  double sc = 0, ss = 0, sq = 0;
  for (int i = tid; i < kTilePix; i += kTileThreads) {
    const int x = x0 + (i & (kTileW - 1)), y = y0 + (i >> kTileWShift);
    if (x < W && y < H) {
      const float v = plane[i];
      gp[(int64_t)y * W + x] = v;
      accum_stat(v, sc, ss, sq);
    }
  }
  if (with_stats)
    finish_stats_at(sc, ss, sq, w, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, partials, tickets, mean_std, stats_out);
}

// grid = (blocks_per_window, n_windows).  Partials are combined in block order by the last block
// to finish (ticket counter), so mean/std do not depend on scheduling.
__global__ void __launch_bounds__(kStatThreads)
voxel_stats_kernel(const float* __restrict__ grid, int64_t vox, StatPartial* __restrict__ partials,
                   unsigned int* __restrict__ tickets, float* __restrict__ mean_std,
                   double* __restrict__ stats_out) {
  const int w = blockIdx.y;
  const float* g = grid + (int64_t)w * vox;
  double c = 0, s = 0, q = 0;
  const int64_t stride = (int64_t)gridDim.x * kStatThreads;
  const int64_t tid = (int64_t)blockIdx.x * kStatThreads + threadIdx.x;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  const int64_t nvec = vec_ok ? vox / 4 : 0;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = tid; i < nvec; i += stride) {
    const float4 v = g4[i];
    accum_stat(v.x, c, s, q);
    accum_stat(v.y, c, s, q);
    accum_stat(v.z, c, s, q);
    accum_stat(v.w, c, s, q);
  }
  for (int64_t i = nvec * 4 + tid; i < vox; i += stride) accum_stat(g[i], c, s, q);
  finish_stats(c, s, q, w, partials, tickets, mean_std, stats_out);
}

// grid[b][pix] = S[b][pix][0] + S[b-1][pix][1]; thread = pixel, walking the bins so every scratch
// element is read exactly once (8-byte loads, coalesced across pixels).  With kStats the
// normalisation statistics are accumulated on the fly.  grid = (blocks, n_windows).
template <bool kStats>
__global__ void __launch_bounds__(kStatThreads)
voxel_combine_kernel(const float2* __restrict__ scratch, int nb, int64_t HW, float* __restrict__ grid,
                     StatPartial* __restrict__ partials, unsigned int* __restrict__ tickets,
                     float* __restrict__ mean_std, double* __restrict__ stats_out) {
  const int w = blockIdx.y;
  const float2* S = scratch + (int64_t)w * nb * HW;
  float* g = grid + (int64_t)w * nb * HW;
  double c = 0, s = 0, q = 0;
  for (int64_t pix = (int64_t)blockIdx.x * kStatThreads + threadIdx.x; pix < HW; pix += (int64_t)gridDim.x * kStatThreads) {
    float carry = 0.0f;
    for (int b = 0; b < nb; ++b) {
      float2 v;
      asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(S + (int64_t)b * HW + pix));
      const float out = __fadd_rn(v.x, carry);
      carry = v.y;
      g[(int64_t)b * HW + pix] = out;
      if (kStats) accum_stat(out, c, s, q);
    }
  }
  if (kStats) finish_stats(c, s, q, w, partials, tickets, mean_std, stats_out);
}

__device__ __forceinline__ float normalize_one(float v, float mean, float sd, bool divide) {
  if (v == 0.0f) return v;
  const float c = __fsub_rn(v, mean);
  return divide ? __fdiv_rn(c, sd) : c;
}

__global__ void __launch_bounds__(256)
voxel_apply_kernel(float* __restrict__ grid, int64_t vox, const float* __restrict__ mean_std) {
  const int w = blockIdx.y;
  float* g = grid + (int64_t)w * vox;
  const float mean = mean_std[2 * w + 0], sd = mean_std[2 * w + 1];
  const bool divide = sd > 0.0f;  // false for NaN, matching "if std > 0"
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  const int64_t nvec = vec_ok ? vox / 4 : 0;
  float4* g4 = reinterpret_cast<float4*>(g);
  for (int64_t i = tid; i < nvec; i += stride) {
    float4 v = g4[i];
    v.x = normalize_one(v.x, mean, sd, divide);
    v.y = normalize_one(v.y, mean, sd, divide);
    v.z = normalize_one(v.z, mean, sd, divide);
    v.w = normalize_one(v.w, mean, sd, divide);
    g4[i] = v;
  }
  for (int64_t i = nvec * 4 + tid; i < vox; i += stride) g[i] = normalize_one(g[i], mean, sd, divide);
}


// T[window][pixel][nbp] -> grid[window][bin][pixel], normalising on the way when kNorm.  A thread owns one pixel:
// it reads the pixel's nbp contiguous bins and writes one value per bin plane, so a warp reads 32*nbp*4 contiguous
// bytes and every store instruction writes 128 contiguous bytes of one plane.  grid = (blocks, n_windows).
template <bool kNorm>
__global__ void __launch_bounds__(256)
voxel_deinterleave_kernel(const float* __restrict__ T, int nb, int nbp, int64_t HW, float* __restrict__ grid,
                          const float* __restrict__ mean_std) {
  const int w = blockIdx.y;
  const float* t = T + (int64_t)w * HW * nbp;
  float* g = grid + (int64_t)w * HW * nb;
  float mean = 0.f, sd = 0.f;
  if (kNorm) {
    mean = mean_std[2 * w + 0];
    sd = mean_std[2 * w + 1];
  }
  const bool divide = sd > 0.0f;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += (int64_t)gridDim.x * blockDim.x) {
    const float2* src = reinterpret_cast<const float2*>(t + pix * nbp);      // nbp is even: 8-byte aligned
    for (int b = 0; b < nbp; b += 2) {
      float2 v;
      asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(src + (b >> 1)));
      st_stream(g + (int64_t)b * HW + pix, kNorm ? normalize_one(v.x, mean, sd, divide) : v.x);
      if (b + 1 < nb) st_stream(g + (int64_t)(b + 1) * HW + pix, kNorm ? normalize_one(v.y, mean, sd, divide) : v.y);
    }
  }
}

// ---- cluster-resident path (windows whose grid fits the shared memory of one 8-CTA cluster) -------------------
// MVSEC-sized windows (5 x 260 x 346 fp32 = 1.8 MB) fit the distributed shared memory of a cluster of 8 CTAs
// (8 x 225 KB).  One cluster then owns a window from the first vote to the normalised result: every CTA keeps
// one eighth of the grid in its shared memory, votes are atomics into the owning CTA's slice over DSMEM, the non-zero
// statistics are reduced inside the cluster, and the finished grid goes to HBM exactly once.  HBM traffic per window
// = 32 N bytes of events + one grid write, instead of memset + L2 atomics + a statistics read + a read-modify-write for
// the normalisation; five launches (and the 2 us floor each has) become one.
// Shared memory has no float RED on sm_100a (atomicAdd(float) is a compare-and-swap loop, over the cluster network a
// remote round trip each: measured 277 us against 156 us for the L2 path in round 1), but INTEGER adds are native.  The
// votes are therefore accumulated in fixed point, 2^-22 per unit (rounding error <= 1.2e-7 per vote, the size of the
// fp32 rounding of the vote itself; the order-free mode's gate is 1e-5).  An int32 cell overflows at |sum| = 512, so
// every add is issued with return and checked: a cell that ever reaches half the range (or a vote larger than 64)
// raises a flag, and the cluster then REDOES that window with the float compare-and-swap adds -- slow, correct, and
// only taken by windows with > 256 net votes on one voxel.
namespace cg = cooperative_groups;
constexpr int kClusterSize = 8;
constexpr int kClusterThreads = 1024;
constexpr int kClusterScratchBytes = 1024;   // [3][32] warp partials + 3 doubles CTA partial + mean/std

// Tail shared by the cluster-resident kernels: the CTA's statistics (count, sum, sum of squares of its slice's non-zero
// voxels) are reduced in the block, exchanged over distributed shared memory (rank-ordered sum: every CTA derives the
// same mean / std, independent of scheduling), and the slice is normalised out of shared memory into `g`.
__device__ __forceinline__ void cluster_stats_apply(cg::cluster_group& cluster, int rank, const float* mine, int valid, float* __restrict__ g,
                                                    double c, double s, double q, bool normalize, double* red, double* partial, float* ms,
                                                    double* __restrict__ stats_out, int w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = valid / 4;
  float mean = 0.f, sd = 0.f;
  if (normalize) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, o);
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      red[warp] = c;
      red[32 + warp] = s;
      red[64 + warp] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double bc = 0, bs = 0, bq = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
        bc += red[k];
        bs += red[32 + k];
        bq += red[64 + k];
      }
      partial[0] = bc;
      partial[1] = bs;
      partial[2] = bq;
    }
    cluster.sync();
    if (threadIdx.x == 0) {
      double tc = 0, ts = 0, tq = 0;
      for (int r = 0; r < kClusterSize; ++r) {       // rank order: every CTA derives identical statistics
        const double* pr = cluster.map_shared_rank(partial, r);
        tc += pr[0];
        ts += pr[1];
        tq += pr[2];
      }
      float m_ = 0.0f, sd_ = 0.0f;                    // same arithmetic as finish_stats_at
      if (tc > 0) {
        const double m = ts / tc;
        m_ = (float)m;
        if (tc > 1) {
          double var = (tq - ts * m) / (tc - 1.0);
          if (var < 0) var = 0;
          sd_ = (float)sqrt(var);
        } else {
          sd_ = __int_as_float(0x7fc00000);
        }
      }
      ms[0] = m_;
      ms[1] = sd_;
      if (stats_out != nullptr && rank == 0) {
        stats_out[3 * w + 0] = tc;
        stats_out[3 * w + 1] = (double)m_;
        stats_out[3 * w + 2] = (double)sd_;
      }
    }
    __syncthreads();
    mean = ms[0];
    sd = ms[1];
  }
  const bool divide = sd > 0.0f;
  // (v - mean) / sd as q0 = c * r, rem = c - q0 * sd (exact, one FMA), q = q0 + rem * r with r = RN(1 / sd): the
  // correctly rounded quotient (Markstein), i.e. the bits of __fdiv_rn in 3 instructions; huge / tiny operands (where the
  // intermediate products could leave the normal range) take the IEEE division
  const bool fast_div = divide && sd > 1.0e-15f && sd < 1.0e15f;
  const float rcp = fast_div ? __frcp_rn(sd) : 0.0f;
  auto norm = [&](float v) {
    if (!normalize || v == 0.0f) return v;
    const float cv = __fsub_rn(v, mean);
    if (!divide) return cv;
    if (!fast_div || !(fabsf(cv) < 1.0e15f && fabsf(cv) > 1.0e-15f)) return __fdiv_rn(cv, sd);
    const float q0 = __fmul_rn(cv, rcp);
    return __fmaf_rn(__fmaf_rn(-q0, sd, cv), rcp, q0);
  };
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  const int nv = vec_ok ? nvec : 0;
  #pragma unroll 4
  for (int i = threadIdx.x; i < nv; i += (int)blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(mine + 4 * i);
    v.x = norm(v.x);
    v.y = norm(v.y);
    v.z = norm(v.z);
    v.w = norm(v.w);
    st_stream4(g + 4 * i, v);
  }
  for (int i = nv * 4 + threadIdx.x; i < valid; i += (int)blockDim.x) g[i] = norm(mine[i]);
}

// statistics of a slice that already sits in shared memory (four voxels summed in fp32, running sums in fp64)
__device__ __forceinline__ void slice_stats(const float* mine, int valid, double& c, double& s, double& q) {
  const int nvec = valid / 4;
#pragma unroll 4
  for (int i = threadIdx.x; i < nvec; i += (int)blockDim.x) {
    const float4 f = *reinterpret_cast<const float4*>(mine + 4 * i);
    c += (double)((f.x != 0.0f) + (f.y != 0.0f) + (f.z != 0.0f) + (f.w != 0.0f));
    s += (double)((f.x + f.y) + (f.z + f.w));
    q += (double)fmaf(f.x, f.x, fmaf(f.y, f.y, fmaf(f.z, f.z, f.w * f.w)));
  }
  for (int i = nvec * 4 + threadIdx.x; i < valid; i += (int)blockDim.x) accum_stat(mine[i], c, s, q);
}

constexpr float kFixScale = 4194304.0f;          // 2^22 units per 1.0
constexpr int kFixGuard = 1 << 30;               // half of the int32 range

template <class Src>
__global__ void __launch_bounds__(kClusterThreads, 1)
voxel_cluster_kernel(const Src ev, const int64_t* __restrict__ offsets, int n_windows, int nb, int H, int W,
                     int slice, int normalize, float* __restrict__ grid, int64_t* __restrict__ dropped,
                     double* __restrict__ stats_out, int noret) {
  extern __shared__ __align__(16) float cl_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / kClusterSize, n_clusters = gridDim.x / kClusterSize;
  float* mine = cl_smem;
  double* red = reinterpret_cast<double*>(cl_smem + slice);     // slice is a multiple of 4 floats
  double* partial = red + 96;
  float* ms = reinterpret_cast<float*>(partial + 3);
  auto dsmem_addr = [](const void* local, int cta) {            // shared::cluster address of `local` in CTA `cta` of the cluster
    uint32_t a;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"((uint32_t)__cvta_generic_to_shared(local)), "r"(cta));
    return a;
  };
  int* ovf = reinterpret_cast<int*>(ms + 2);                    // this CTA saw a cell near the fixed-point range
  const int64_t HW = (int64_t)H * W, vox = HW * nb;

  for (int w = cluster_id; w < n_windows; w += n_clusters) {
    const int64_t begin = offsets[w], end = offsets[w + 1];
    const int64_t n = end - begin;
    bool use_float = false;
    for (int attempt = 0; attempt < 2; ++attempt) {             // cluster-uniform: 0 fixed point, 1 float redo
      for (int i = threadIdx.x * 4; i < slice; i += kClusterThreads * 4)
        *reinterpret_cast<float4*>(mine + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      cluster.sync();
      int ndrop = 0;
      bool bad = false;
      if (n > 0) {
        const WindowTimes wt = window_times(ev, begin, end);
        constexpr int kE = 2;                                     // (64 registers per thread at 1024 threads)
        const int64_t stride = (int64_t)kClusterSize * kClusterThreads;
        for (int64_t i0 = (int64_t)rank * kClusterThreads + threadIdx.x; i0 < n; i0 += stride * kE) {
          EventRow rows[kE];
#pragma unroll
          for (int k = 0; k < kE; ++k)
            if (i0 + k * stride < n) rows[k] = ev.load(begin + i0 + k * stride);
#pragma unroll
          for (int k = 0; k < kE; ++k) {
            if (i0 + k * stride < n) {
              const Vote v = make_vote(rows[k], wt.t_first, wt.dT, nb, W, HW, vox);
              auto add = [&](int64_t idx, float val) {          // vox < 2^31 on this path: 32-bit index arithmetic
                const int il = (int)idx, r = il / slice;
                float* cell = cluster.map_shared_rank(mine + (il - r * slice), r);
                if (use_float) {
                  atomicAdd(cell, val);
                } else {
                  if (noret) {            // timing experiment (EEM_VOXEL_CLUSTER_NORET=1): fire-and-forget adds, no range check
                    asm volatile("red.shared::cluster.add.s32 [%0], %1;" ::"r"(dsmem_addr(mine + (il - r * slice), r)), "r"(__float2int_rn(val * kFixScale)) : "memory");
                  } else {
                    const int old = atomicAdd(reinterpret_cast<int*>(cell), __float2int_rn(val * kFixScale));
                    bad = bad || !(fabsf(val) < 64.0f) || old >= kFixGuard || old <= -kFixGuard;
                  }
                }
              };
              if (v.idx_left >= 0) add(v.idx_left, v.val_left);
              if (v.idx_right >= 0) add(v.idx_right, v.val_right);
              ndrop += (int)v.oob_left + (int)v.oob_right;
            }
          }
        }
      }
      if (attempt == 0 && dropped != nullptr && ndrop != 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(dropped), (unsigned long long)ndrop);
      const int any_bad = __syncthreads_or(bad ? 1 : 0);
      if (threadIdx.x == 0) *ovf = any_bad;
      cluster.sync();                                           // all votes of the window have landed; flags are visible
      if (use_float) break;
      int redo = 0;
      for (int r = 0; r < kClusterSize; ++r) redo |= *cluster.map_shared_rank(ovf, r);
      if (!redo) break;
      use_float = true;
      cluster.sync();                                           // every CTA has read the flags before they are rewritten
    }

    const int64_t left = vox - (int64_t)rank * slice;             // cells of this slice inside the window
    const int valid = (int)(left < 0 ? 0 : (left < slice ? left : slice));
    if (!use_float) {                                             // fixed point -> fp32, in place (exact below |v| = 4)
      constexpr float kInv = 1.0f / kFixScale;
      for (int i = threadIdx.x * 4; i < slice; i += kClusterThreads * 4) {
        const int4 q = *reinterpret_cast<const int4*>(mine + i);
        *reinterpret_cast<float4*>(mine + i) = make_float4((float)q.x * kInv, (float)q.y * kInv, (float)q.z * kInv, (float)q.w * kInv);
      }
      __syncthreads();
    }
    double c = 0, s = 0, q = 0;
    if (normalize) slice_stats(mine, valid, c, s, q);
    float* out = grid + (int64_t)w * vox + (int64_t)rank * slice;
    cluster_stats_apply(cluster, rank, mine, valid, out, c, s, q, normalize != 0, red, partial, ms, stats_out, w);
    cluster.sync();   // the slice / `partial` / `ms` are rewritten for the next window only after every peer is done with them
  }
}

// ---- cluster-synchronised STREAMING normalisation (EEM_VOXEL_NORM=stream; measured, not the default) -----------------
// The kernel boundary between K2's statistics and apply passes exists only because the statistics are a whole-window
// reduction.  Here a window is owned by a cluster of 8 CTAs without shared-memory residency: pass 1 reads the CTA's
// eighth of the window and accumulates the non-zero statistics, a cluster barrier + distributed shared memory exchange
// gives every CTA the window's mean / std, pass 2 re-reads the slice (from L2) and writes the normalised values.  Every
// window proceeds on its own and all windows of a call are in flight at once -- but it moves 12 bytes per voxel like
// the two-kernel path, and that traffic is what bounds all forms (see norm_mode()).
constexpr int kStreamThreads = 1024;
constexpr int64_t kStreamNormMaxBytes = 8ll << 20;   // windows up to 8 MB: the second pass should find the first one's lines in L2

__global__ void __launch_bounds__(kStreamThreads)
voxel_normalize_stream_kernel(float* __restrict__ grid, int n_windows, int64_t vox, int slice, double* __restrict__ stats_out) {
  __shared__ double red[96];
  __shared__ double partial[3];
  __shared__ float ms[2];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / kClusterSize, n_clusters = gridDim.x / kClusterSize;
  for (int w = cluster_id; w < n_windows; w += n_clusters) {
    float* g = grid + (int64_t)w * vox + (int64_t)rank * slice;
    const int64_t left = vox - (int64_t)rank * slice;             // cells of this slice inside the window
    const int valid = (int)(left < 0 ? 0 : (left < slice ? left : slice));
    double c = 0, s = 0, q = 0;
    slice_stats(g, valid, c, s, q);
    cluster_stats_apply(cluster, rank, g, valid, g, c, s, q, true, red, partial, ms, stats_out, w);
    cluster.sync();   // `partial` / `ms` are rewritten for the next window only after every peer has read them
  }
}

// ---- cluster-resident NORMALISATION (windows whose grid fits the shared memory of one 8-CTA cluster) ------------------
// K2 as two kernels reads every voxel twice and writes it once, with a launch boundary in between because the
// statistics are a whole-window reduction.  For MVSEC-sized windows (5 x 260 x 346 fp32 = 1.8 MB) one 8-CTA cluster
// holds a whole window in its shared memory (8 x 225 KB): every CTA pulls its eighth in ONCE while accumulating the
// non-zero statistics, the partials are exchanged over distributed shared memory (cluster barrier, rank-ordered sum, so
// the result does not depend on scheduling), and the normalised values are written back from shared memory -- one
// launch, 8 bytes of traffic per voxel instead of 12.  Same arithmetic as voxel_stats_kernel / finish_stats_at /
// voxel_apply_kernel.  (The votes themselves stay L2 atomics: shared memory has no float RED, see cluster_plan().)
__global__ void __launch_bounds__(kClusterThreads, 1)
voxel_normalize_cluster_kernel(float* __restrict__ grid, int n_windows, int64_t vox, int slice, double* __restrict__ stats_out) {
  extern __shared__ __align__(16) float cl_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / kClusterSize, n_clusters = gridDim.x / kClusterSize;
  float* mine = cl_smem;
  double* red = reinterpret_cast<double*>(cl_smem + slice);     // slice is a multiple of 4 floats
  double* partial = red + 96;
  float* ms = reinterpret_cast<float*>(partial + 3);

  for (int w = cluster_id; w < n_windows; w += n_clusters) {
    float* g = grid + (int64_t)w * vox + (int64_t)rank * slice;
    const int64_t left = vox - (int64_t)rank * slice;             // cells of this slice inside the window
    const int valid = (int)(left < 0 ? 0 : (left < slice ? left : slice));
    const int nvec = valid / 4;                                   // the caller guarantees 16-byte aligned windows
    double c = 0, s = 0, q = 0;
    constexpr int kU = 7;                                         // loads in flight per thread (MVSEC: 2 rounds of 7)
    for (int i0 = threadIdx.x; i0 < nvec; i0 += kClusterThreads * kU) {
      float4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * kClusterThreads;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < nvec) v[u] = __ldcs(reinterpret_cast<const float4*>(g) + i);
      }
      int cnt = 0;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * kClusterThreads;
        if (i < nvec) *reinterpret_cast<float4*>(mine + 4 * i) = v[u];
        // zeros (also the padding ones) contribute nothing: four voxels are summed in fp32 (relative error 1e-7 of a
        // partial of at most four terms), the running sums stay fp64 -- a quarter of the fp64 instructions
        const float4 f = v[u];
        cnt += (f.x != 0.0f) + (f.y != 0.0f) + (f.z != 0.0f) + (f.w != 0.0f);
        s += (double)((f.x + f.y) + (f.z + f.w));
        q += (double)fmaf(f.x, f.x, fmaf(f.y, f.y, fmaf(f.z, f.z, f.w * f.w)));
      }
      c += (double)cnt;
    }
    for (int i = nvec * 4 + threadIdx.x; i < valid; i += kClusterThreads) {
      const float v = g[i];
      mine[i] = v;
      accum_stat(v, c, s, q);
    }
    cluster_stats_apply(cluster, rank, mine, valid, g, c, s, q, true, red, partial, ms, stats_out, w);
    cluster.sync();   // `partial` / `ms` are rewritten for the next window only after every peer has read them
  }
}

struct NormClusterPlan {
  bool ok;
  int slice;
  size_t smem;
  int max_clusters;
};

// Possible when a window fits the cluster's shared memory and its slices are 16-byte aligned; EEM_VOXEL_NORM=split
// forces the two-kernel path (comparisons).
NormClusterPlan norm_cluster_plan(const float* grid, int64_t vox) {
  NormClusterPlan pl{false, 0, 0, 0};
  if (vox >= (1ll << 31) || vox % 4 != 0 || (reinterpret_cast<uintptr_t>(grid) & 15) != 0) return pl;
  pl.slice = (int)align_up((size_t)ceil_div(vox, kClusterSize), 4);
  pl.smem = (size_t)pl.slice * sizeof(float) + kClusterScratchBytes;
  int dev = 0, optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return pl;
  if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return pl;
  if (pl.smem > (size_t)optin) return pl;
  static std::mutex mu;
  static size_t configured_dev[kMaxDevices] = {};       // per device: the attribute lives in the device's context
  static int clusters_dev[kMaxDevices] = {};
  std::lock_guard<std::mutex> lock(mu);
  if (pl.smem > configured_dev[dev] || clusters_dev[dev] == 0) {
    if (cudaFuncSetAttribute(voxel_normalize_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) {
      cudaGetLastError();
      return pl;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kClusterSize, 1, 1);
    cfg.blockDim = dim3(kClusterThreads, 1, 1);
    cfg.dynamicSmemBytes = pl.smem;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kClusterSize;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, voxel_normalize_cluster_kernel, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      return pl;
    }
    configured_dev[dev] = pl.smem;
    clusters_dev[dev] = n;
    if (getenv("EEM_VOXEL_DEBUG")) fprintf(stderr, "[eemflow_b200] cluster normalisation: %d co-resident clusters of %d CTAs (%zu B smem)\n", n, kClusterSize, pl.smem);
  }
  pl.max_clusters = clusters_dev[dev];
  pl.ok = true;
  return pl;
}

int bit_length(uint64_t v) {
  int b = 0;
  while (v) {
    ++b;
    v >>= 1;
  }
  return b;
}

struct DetLayout {
  size_t keys_a, keys_b, vals_a, vals_b, counts, block_sums, total;
  int64_t n_votes, n_segs, n_counts, n_scan_blocks;
};

DetLayout det_layout(int64_t n_total) {
  DetLayout L{};
  L.n_votes = 2 * n_total;
  L.n_segs = ceil_div(L.n_votes > 0 ? L.n_votes : 1, kSortSegment);
  L.n_counts = L.n_segs * kRadix;
  L.n_scan_blocks = ceil_div(L.n_counts, kScanTile);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t nv = (size_t)(L.n_votes > 0 ? L.n_votes : 1);
  L.keys_a = take(nv * 4);
  L.keys_b = take(nv * 4);
  L.vals_a = take(nv * 4);
  L.vals_b = take(nv * 4);
  L.counts = take((size_t)L.n_counts * 4);
  L.block_sums = take((size_t)L.n_scan_blocks * 4);
  L.total = off;
  return L;
}

}  // namespace
}  // namespace eem

using namespace eem;

namespace eem {
namespace {

int stat_blocks(int64_t vox, int n_windows) {
  // enough CTAs to fill the chip even for a single window (8 per SM), never more than the work
  int64_t want = ceil_div((int64_t)sm_count() > 0 ? (int64_t)sm_count() * 8 : 1184, n_windows > 0 ? n_windows : 1);
  if (want < 8) want = 8;
  int64_t b = ceil_div(vox, (int64_t)kStatThreads * 4);
  if (b > want) b = want;
  if (b < 1) b = 1;
  if (b > kStatBlocksPerWindowMax) b = kStatBlocksPerWindowMax;
  return (int)b;
}

struct StatLayout {
  size_t partials, tickets, mean_std, total;
};

StatLayout stat_layout(int n_windows, int64_t parts_per_window = kStatBlocksPerWindowMax) {
  StatLayout L{};
  size_t off = 0;
  L.partials = off;
  off = align_up(off + (size_t)n_windows * (size_t)parts_per_window * sizeof(StatPartial), 256);
  L.tickets = off;
  off = align_up(off + (size_t)n_windows * sizeof(unsigned int), 256);
  L.mean_std = off;
  off = align_up(off + (size_t)n_windows * 2 * sizeof(float), 256);
  L.total = off;
  return L;
}

// Pair layout pays 2x grid bytes of scratch traffic (memset + combine read) and wins when the events
// dominate: measured crossover on B200 at ~2 events per voxel (HREM dt4: 500 -> 380 us; dt1: 146 vs 163 us).
// Which implementation a call takes.
//   deterministic mode  the tile-binned path (bit-exact; HREM dt1, 10 M events: 308 us against 2.9 ms for the radix sort
//                       of round 1, which remains as the fallback for shapes the tiles do not cover)
//   order-free mode     the L2-atomic paths of round 1 (direct / pair layout): measured FASTER than the tile-binned
//                       path on B200 (HREM dt1 K1 144 us against 260 us; MVSEC x64 61 us against 154 us): both tile
//                       passes are instruction-latency-bound (pass 1 ~100 instructions per event at 29 % issue
//                       utilisation, pass 2 ~7 warp instructions per record), see DESIGN.md section 4.
// EEM_VOXEL_PATH (timing experiments and the forced-path parity tests) overrides: t(iled) for both modes -- the
// order-free mode then runs the exact kernel too, or the shared-memory-atomic variant with EEM_VOXEL_ORDERFREE=1 --,
// d(irect L2 atomics), p(air layout), c(luster-resident), i(nterleaved) for the order-free mode, r(adix sort) for the
// deterministic mode.
bool tiled_path_enabled(int mode) {
  if (const char* v = getenv("EEM_VOXEL_PATH")) return v[0] == 't';
  return mode == EEM_VOXEL_DETERMINISTIC;
}

bool tiled_path_fits(int num_bins, int height, int width, int64_t vox) {
  const int64_t tiles = (int64_t)((width + kTileW - 1) >> kTileWShift) * ((height + kTileH - 1) >> kTileHShift);
  return tiles <= kMaxTiles && num_bins <= kMaxBinsTiled && num_bins <= 65535 && tiles * num_bins <= kMaxStatParts && vox < (1ll << 31);
}

// partial-statistics slots per window: one per (tile, bin) CTA of pass 2, never fewer than the streaming kernels use
int64_t tiled_stat_parts(int num_bins, int height, int width) {
  const int64_t tiles = (int64_t)((width + kTileW - 1) >> kTileWShift) * ((height + kTileH - 1) >> kTileHShift);
  const int64_t parts = tiles * num_bins;
  return parts > kStatBlocksPerWindowMax ? parts : kStatBlocksPerWindowMax;
}

size_t bin_kernel_smem(int n_tiles) {
  const int hist_stride = (n_tiles + 1) | 1;
  return (size_t)kBinChunk * 8 + ((size_t)kBinWarps * hist_stride + (size_t)n_tiles + 1) * sizeof(uint16_t) + 16;
}

bool use_pair_path(int64_t n_total, int64_t total_vox) {
  if (const char* v = getenv("EEM_VOXEL_PATH")) {   // timing experiments only
    if (v[0] == 'p') return true;
    if (v[0] == 'd') return false;
  }
  return n_total >= 2 * total_vox;
}

// Bin-interleaved scratch layout.  MEASURED SLOWER than the direct path (HREM dt1, 15 bins, 4 x 10 M events: voxel
// family 0.735 vs 0.697 ms) although it issues 25 % fewer L2 atomic operations, so it is only taken when forced
// (EEM_VOXEL_PATH=interleaved); kept, with its parity test, as a measured experiment.
bool use_interleaved_path(int num_bins) {
  if (const char* v = getenv("EEM_VOXEL_PATH")) {   // timing experiments only
    if (v[0] == 'i') return true;
    if (v[0] == 'd' || v[0] == 'p' || v[0] == 'c') return false;
  }
  (void)num_bins;
  return false;
}

int64_t l2_group_bytes() {
  if (const char* v = getenv("EEM_VOXEL_GROUP_MB")) {   // timing experiments only; 0 disables the grouping
    const long mb = atol(v);
    if (mb == 0) return (int64_t)1 << 62;
    if (mb > 0) return (int64_t)mb << 20;
  }
  return (int64_t)60 << 20;
}

int time_lanes(int dflt) {
  if (const char* v = getenv("EEM_VOXEL_LANES")) {   // timing experiments only
    const int n = atoi(v);
    if (n >= 1 && n <= 1024) return n;
  }
  return dflt;
}

// Cluster-resident path: possible when one window's grid fits the shared memory of an 8-CTA cluster.  MEASURED SLOWER
// than the L2-atomic path in both of its forms and therefore only taken when forced (EEM_VOXEL_PATH=cluster; tested):
// round 1, float compare-and-swap adds over DSMEM: 277 us against 156 us for MVSEC x64 windows; round 2, fixed-point
// integer adds (native): voxelize family of the bench 0.194 ms (adds with return + range check) / 0.185 ms
// (fire-and-forget red.shared::cluster, EEM_VOXEL_CLUSTER_NORET=1) against 0.135 ms -- so the adds are not what costs:
// only 15 clusters are co-resident on a B200 (5 rounds for 64 windows) and inside a cluster the phases (zero, stamps,
// two batches of event loads, votes, conversion, statistics, exchange, apply) run strictly one after the other at
// ~37 us per window, where the L2 kernels keep every SM busy with thousands of independent CTAs.
struct ClusterPlan {
  bool ok;
  int slice;          // floats of the window grid held by each CTA
  size_t smem;
  int max_clusters;   // co-resident clusters on this device
};

template <class Src>
ClusterPlan cluster_plan(int64_t vox, int n_windows, int64_t n_total) {
  ClusterPlan pl{false, 0, 0, 0};
  const char* forced = getenv("EEM_VOXEL_PATH");
  if (forced == nullptr || forced[0] != 'c') return pl;         // opt-in
  if (vox >= (1ll << 31)) return pl;
  pl.slice = (int)align_up((size_t)ceil_div(vox, kClusterSize), 4);
  pl.smem = (size_t)pl.slice * sizeof(float) + kClusterScratchBytes;
  int dev = 0, optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return pl;
  if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return pl;
  if (pl.smem > (size_t)optin) return pl;
  static std::mutex mu;
  static size_t configured_dev[kMaxDevices] = {};       // per device: the attribute lives in the device's context
  static int cached_clusters_dev[kMaxDevices] = {};
  if (dev < 0 || dev >= kMaxDevices) return pl;
  std::lock_guard<std::mutex> lock(mu);
  size_t& configured = configured_dev[dev];
  int& cached_clusters = cached_clusters_dev[dev];
  if (pl.smem > configured || cached_clusters == 0) {
    if (cudaFuncSetAttribute(voxel_cluster_kernel<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) {
      cudaGetLastError();
      return pl;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kClusterSize, 1, 1);
    cfg.blockDim = dim3(kClusterThreads, 1, 1);
    cfg.dynamicSmemBytes = pl.smem;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kClusterSize;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, voxel_cluster_kernel<Src>, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      return pl;
    }
    configured = pl.smem;
    cached_clusters = n;
  }
  pl.max_clusters = cached_clusters;
  const int active = n_windows < pl.max_clusters ? n_windows : pl.max_clusters;
  (void)active;
  (void)n_total;
  pl.ok = true;
  return pl;
}

// EEM_VOXEL_NORM: "cluster" (shared-memory resident; the default where it applies), "stream" (cluster-synchronised, two
// passes from L2), "split" (two kernels).  Measured in the MVSEC step (voxelize family / whole step, ms): cluster 0.135 /
// 0.795, stream 0.158 / 0.807 (256-thread CTAs: 0.163 / 0.834), split 0.129 / 0.850 -- all three are within 1.7x of the
// two L2 passes over 115 MB that any normalisation needs; the resident form moves the fewest bytes and leaves 28 SMs to
// the other branches of the step.
inline int norm_mode() {
  const char* v = getenv("EEM_VOXEL_NORM");
  if (v == nullptr || v[0] == 'c') return 1;
  return (v[0] == 's' && v[1] == 't') ? 0 : 2;
}

// Windows of up to 8 x 2^22 voxels whose slices are 16-byte aligned; grid = min(windows, co-resident clusters) x 8 CTAs.
int launch_normalize_stream(float* grid, int n_windows, int64_t vox, double* stats_out, cudaStream_t stream, bool* done) {
  *done = false;
  if (vox % 4 != 0 || (reinterpret_cast<uintptr_t>(grid) & 15) != 0 || vox > ((int64_t)kClusterSize << 22)) return EEM_OK;
  const int slice = (int)align_up((size_t)ceil_div(vox, kClusterSize), 4);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return EEM_OK;
  static std::mutex mu;
  static int clusters_dev[kMaxDevices] = {};
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kStreamThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = kClusterSize;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (clusters_dev[dev] == 0) {
      cfg.gridDim = dim3(kClusterSize, 1, 1);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, voxel_normalize_stream_kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return EEM_OK;
      }
      clusters_dev[dev] = n;
    }
    max_clusters = clusters_dev[dev];
  }
  const int n_clusters = n_windows < max_clusters ? n_windows : max_clusters;
  cfg.gridDim = dim3((unsigned)(n_clusters * kClusterSize), 1, 1);
  EEM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, voxel_normalize_stream_kernel, grid, n_windows, vox, slice, stats_out));
  EEM_CHECK_LAUNCH("voxel_normalize_stream_kernel");
  *done = true;
  return EEM_OK;
}

int launch_normalize(float* grid, int n_windows, int64_t vox, double* stats_out, char* ws, cudaStream_t stream) {
  if (norm_mode() == 0 && vox * (int64_t)sizeof(float) <= kStreamNormMaxBytes) {
    bool done = false;
    const int rc = launch_normalize_stream(grid, n_windows, vox, stats_out, stream, &done);
    if (rc != EEM_OK || done) return rc;
  }
  if (norm_mode() == 1) {
    const NormClusterPlan pl = norm_cluster_plan(grid, vox);
    if (pl.ok) {
      const int n_clusters = n_windows < pl.max_clusters ? n_windows : pl.max_clusters;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(n_clusters * kClusterSize), 1, 1);
      cfg.blockDim = dim3(kClusterThreads, 1, 1);
      cfg.dynamicSmemBytes = pl.smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr{};
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = kClusterSize;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      EEM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, voxel_normalize_cluster_kernel, grid, n_windows, vox, pl.slice, stats_out));
      EEM_CHECK_LAUNCH("voxel_normalize_cluster_kernel");
      return EEM_OK;
    }
  }
  const StatLayout L = stat_layout(n_windows);
  StatPartial* partials = reinterpret_cast<StatPartial*>(ws + L.partials);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(ws + L.tickets);
  float* mean_std = reinterpret_cast<float*>(ws + L.mean_std);
  // Tickets are reset by the kernel itself after use; the memset makes a fresh workspace valid.
  EEM_CHECK_CUDA(cudaMemsetAsync(tickets, 0, (size_t)n_windows * sizeof(unsigned int), stream));
  dim3 g((unsigned)stat_blocks(vox, n_windows), (unsigned)n_windows);
  voxel_stats_kernel<<<g, kStatThreads, 0, stream>>>(grid, vox, partials, tickets, mean_std, stats_out);
  EEM_CHECK_LAUNCH("voxel_stats_kernel");
  voxel_apply_kernel<<<g, 256, 0, stream>>>(grid, vox, mean_std);
  EEM_CHECK_LAUNCH("voxel_apply_kernel");
  return EEM_OK;
}

}  // namespace
}  // namespace eem

extern "C" size_t eem_voxelize_workspace_bytes(int64_t n_total, int n_windows, int num_bins, int height, int width,
                                               int mode, int normalize);

template <class Src>
int voxelize_impl(const Src events, const int64_t* offsets, int n_windows, int64_t n_total,
                  int64_t max_events_per_window, int num_bins, int height, int width, int mode,
                  int normalize, float* grid, int64_t* dropped, double* stats_out, void* workspace,
                  size_t workspace_bytes, eem_stream_t stream_) {
  EEM_CHECK_ARG(n_windows > 0, "eem_voxelize: n_windows must be > 0 (got %d)", n_windows);
  EEM_CHECK_ARG(num_bins > 0, "eem_voxelize: num_bins must be > 0 (got %d)", num_bins);
  EEM_CHECK_ARG(height > 0 && width > 0, "eem_voxelize: height/width must be > 0 (got %dx%d)", height, width);
  EEM_CHECK_ARG(n_total >= 0 && max_events_per_window >= 0 && max_events_per_window <= n_total,
                "eem_voxelize: bad event counts (n_total=%lld, max_per_window=%lld)",
                (long long)n_total, (long long)max_events_per_window);
  EEM_CHECK_ARG(grid != nullptr && offsets != nullptr, "eem_voxelize: NULL grid/offsets");
  EEM_CHECK_ARG(mode == EEM_VOXEL_ATOMIC || mode == EEM_VOXEL_DETERMINISTIC,
                "eem_voxelize: unknown mode %d", mode);
  EEM_CHECK_ARG(n_windows <= 65535, "eem_voxelize: more than 65535 windows in one call");
  EEM_CHECK_ALIGNED(grid, 4);
  cudaStream_t stream = as_stream(stream_);
  const int64_t vox = (int64_t)num_bins * height * width;
  const int64_t total_vox = vox * n_windows;
  const size_t need = eem_voxelize_workspace_bytes(n_total, n_windows, num_bins, height, width, mode, normalize);
  if (need > 0) {
    if (workspace == nullptr || workspace_bytes < need)
      return fail(EEM_ERR_WORKSPACE, "eem_voxelize: workspace of %zu bytes required, got %zu", need, workspace_bytes);
    EEM_CHECK_ALIGNED(workspace, 256);
  }
  // Many small windows whose grids together exceed the L2: run them in groups whose grids fit (~half of the
  // 126 MB L2), so the memset, the L2-resolved votes, the statistics read and the normalisation read-modify-write
  // of a group all hit L2 and HBM only sees the events and one write-back of each grid.
  const bool tiled = tiled_path_enabled(mode) && tiled_path_fits(num_bins, height, width, vox);
  // Order-free mode, windows that fit a cluster's shared memory: one launch from the first vote to the normalised grid.
  // Checked before the L2 grouping below: the grids never live in L2 on this path.
  if (mode == EEM_VOXEL_ATOMIC && !tiled && n_total > 0 && max_events_per_window > 0 && !use_pair_path(n_total, total_vox) &&
      (reinterpret_cast<uintptr_t>(grid) & 15) == 0 && vox % 4 == 0) {
    const ClusterPlan pl = cluster_plan<Src>(vox, n_windows, n_total);
    if (pl.ok) {
      const int n_clusters = n_windows < pl.max_clusters ? n_windows : pl.max_clusters;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(n_clusters * kClusterSize), 1, 1);
      cfg.blockDim = dim3(kClusterThreads, 1, 1);
      cfg.dynamicSmemBytes = pl.smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr{};
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = kClusterSize;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      EEM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, voxel_cluster_kernel<Src>, events, offsets, n_windows, num_bins, height, width,
                                        pl.slice, normalize, grid, dropped, stats_out,
                                        getenv("EEM_VOXEL_CLUSTER_NORET") != nullptr ? 1 : 0));
      EEM_CHECK_LAUNCH("voxel_cluster_kernel");
      return EEM_OK;
    }
  }
  if ((mode == EEM_VOXEL_ATOMIC || tiled) && n_windows > 1 && total_vox * (int64_t)sizeof(float) > l2_group_bytes() &&
      vox * (int64_t)sizeof(float) <= l2_group_bytes()) {
    const int per_group = (int)(l2_group_bytes() / (vox * (int64_t)sizeof(float)));
    const int n_groups = (int)ceil_div(n_windows, per_group);
    const int even = (int)ceil_div(n_windows, n_groups);             // equal-sized groups
    // The shared-memory-resident normalisation (EEM_VOXEL_NORM=cluster) runs in rounds of as many windows as clusters are
    // co-resident (15 on a B200): one launch over ALL windows after the last group's votes needs ceil(64 / 15) = 5 rounds
    // where two launches over 32 windows need 2 x 3.  The streaming form (default) has every window in flight at once and
    // normalises each group right after its votes, while its grids are L2-hot.
    const bool norm_at_end = normalize && !tiled && norm_mode() == 1 && norm_cluster_plan(grid, vox).ok;
    for (int w0 = 0; w0 < n_windows; w0 += even) {
      const int nw = n_windows - w0 < even ? n_windows - w0 : even;
      // upper bound of the group's event count (sizes its scratch): never more than the whole call
      int64_t n_est = (int64_t)nw * max_events_per_window;
      if (n_est > n_total) n_est = n_total;
      const int rc = voxelize_impl<Src>(events, offsets + w0, nw, n_est, max_events_per_window, num_bins, height, width, mode,
                                        norm_at_end ? 0 : normalize, grid + (int64_t)w0 * vox, dropped,
                                        stats_out ? stats_out + 3 * w0 : nullptr, workspace, workspace_bytes, stream_);
      if (rc != EEM_OK) return rc;
    }
    if (norm_at_end) return launch_normalize(grid, n_windows, vox, stats_out, static_cast<char*>(workspace), stream);
    return EEM_OK;
  }
  char* ws = static_cast<char*>(workspace);
  char* ws_stats = ws;                                             // [stats | mode-specific]
  char* ws_mode = ws + (normalize ? stat_layout(n_windows, tiled ? tiled_stat_parts(num_bins, height, width) : kStatBlocksPerWindowMax).total : 0);

  const bool no_events = (n_total == 0 || max_events_per_window == 0);
  if (tiled && !no_events) {
    const TilePlan tp = tile_plan(height, width, max_events_per_window);
    const TileWorkspace TL = tile_workspace(n_total, n_windows, num_bins, tp.n_tiles, ceil_div(n_total, kBinChunk) + 1);
    uint2* recs = reinterpret_cast<uint2*>(ws_mode + TL.recs);
    float* side = reinterpret_cast<float*>(ws_mode + TL.side);
    uint16_t* table = reinterpret_cast<uint16_t*>(ws_mode + TL.table);
    int* range = reinterpret_cast<int*>(ws_mode + TL.range);
    int* swap_flags = reinterpret_cast<int*>(ws_mode + TL.swap_flags);
    EEM_CHECK_CUDA(cudaMemsetAsync(swap_flags, 0, (size_t)n_windows * sizeof(int), stream));
    const size_t range_half = (size_t)n_windows * num_bins * sizeof(int);
    EEM_CHECK_CUDA(cudaMemsetAsync(range, 0x7f, range_half, stream));                                   // lo = "no chunk yet"
    EEM_CHECK_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(range) + range_half, 0xff, range_half, stream));   // hi = -1
    {
      const size_t smem = bin_kernel_smem(tp.n_tiles);
      static DynSmemOptIn optin;
      EEM_CHECK_CUDA(optin.ensure(voxel_bin_kernel<Src>, smem));
      dim3 g((unsigned)tp.chunks_max, (unsigned)n_windows);
      voxel_bin_kernel<Src><<<g, kBinThreads, smem, stream>>>(events, offsets, num_bins, height, width, tp, recs, side, table, range, swap_flags, dropped);
      EEM_CHECK_LAUNCH("voxel_bin_kernel");
    }
    StatPartial* partials = nullptr;
    unsigned int* tickets = nullptr;
    float* mean_std = nullptr;
    if (normalize) {
      const StatLayout L = stat_layout(n_windows, tiled_stat_parts(num_bins, height, width));
      partials = reinterpret_cast<StatPartial*>(ws_stats + L.partials);
      tickets = reinterpret_cast<unsigned int*>(ws_stats + L.tickets);
      mean_std = reinterpret_cast<float*>(ws_stats + L.mean_std);
      EEM_CHECK_CUDA(cudaMemsetAsync(tickets, 0, (size_t)n_windows * sizeof(unsigned int), stream));
    }
    dim3 g2((unsigned)tp.n_tiles, (unsigned)num_bins, (unsigned)n_windows);
    // the exact (ordered) kernel also serves a forced order-free call unless EEM_VOXEL_ORDERFREE=1 asks for the
    // shared-memory-atomic variant (260 vs 308 us per 10 M events)
    static const bool order_free = [] {
      const char* v = getenv("EEM_VOXEL_ORDERFREE");
      return v != nullptr && atoi(v) == 1;
    }();
    if (mode == EEM_VOXEL_DETERMINISTIC || !order_free)
      voxel_plane_kernel<true><<<g2, kTileThreads, 0, stream>>>(offsets, num_bins, height, width, tp, recs, side, table, range, swap_flags, grid,
                                                                normalize, partials, tickets, mean_std, stats_out);
    else
      voxel_plane_kernel<false><<<g2, kTileThreads, 0, stream>>>(offsets, num_bins, height, width, tp, recs, side, table, range, swap_flags, grid,
                                                                 normalize, partials, tickets, mean_std, stats_out);
    EEM_CHECK_LAUNCH("voxel_plane_kernel");
    if (normalize) {
      dim3 ga((unsigned)stat_blocks(vox, n_windows), (unsigned)n_windows);
      voxel_apply_kernel<<<ga, 256, 0, stream>>>(grid, vox, mean_std);
      EEM_CHECK_LAUNCH("voxel_apply_kernel");
    }
    return EEM_OK;
  }
  const bool pair = !no_events && mode == EEM_VOXEL_ATOMIC && use_pair_path(n_total, total_vox);
  if (!no_events && mode == EEM_VOXEL_ATOMIC && !pair && use_interleaved_path(num_bins) && vox < (1ll << 31)) {
    const int nbp = (num_bins + 1) & ~1;
    const int64_t HW = (int64_t)height * width, vox_p = HW * nbp;
    float* T = reinterpret_cast<float*>(ws_mode);
    EEM_CHECK_CUDA(cudaMemsetAsync(T, 0, (size_t)n_windows * vox_p * sizeof(float), stream));
    const int64_t per_block = (int64_t)kVoteThreads * kVoteEventsPerThread;
    const int64_t chunks = ceil_div(max_events_per_window, per_block);
    const int lanes = (int)std::min<int64_t>(time_lanes(64), chunks);   // short windows: no CTAs without a chunk
    dim3 g((unsigned)(ceil_div(chunks, lanes) * lanes), (unsigned)n_windows);
    voxel_vote_interleaved_kernel<Src><<<g, kVoteThreads, 0, stream>>>(events, offsets, num_bins, nbp, height, width, lanes, T, dropped);
    EEM_CHECK_LAUNCH("voxel_vote_interleaved_kernel");
    dim3 gd((unsigned)stat_blocks(HW * 4, n_windows), (unsigned)n_windows);
    if (normalize) {
      const StatLayout L = stat_layout(n_windows);
      StatPartial* partials = reinterpret_cast<StatPartial*>(ws_stats + L.partials);
      unsigned int* tickets = reinterpret_cast<unsigned int*>(ws_stats + L.tickets);
      float* mean_std = reinterpret_cast<float*>(ws_stats + L.mean_std);
      EEM_CHECK_CUDA(cudaMemsetAsync(tickets, 0, (size_t)n_windows * sizeof(unsigned int), stream));
      dim3 gs((unsigned)stat_blocks(vox_p, n_windows), (unsigned)n_windows);
      voxel_stats_kernel<<<gs, kStatThreads, 0, stream>>>(T, vox_p, partials, tickets, mean_std, stats_out);
      EEM_CHECK_LAUNCH("voxel_stats_kernel");
      voxel_deinterleave_kernel<true><<<gd, 256, 0, stream>>>(T, num_bins, nbp, HW, grid, mean_std);
    } else {
      voxel_deinterleave_kernel<false><<<gd, 256, 0, stream>>>(T, num_bins, nbp, HW, grid, nullptr);
    }
    EEM_CHECK_LAUNCH("voxel_deinterleave_kernel");
    return EEM_OK;
  }
  if (!pair) EEM_CHECK_CUDA(cudaMemsetAsync(grid, 0, (size_t)total_vox * sizeof(float), stream));

  if (no_events) {
    // nothing to vote
  } else if (mode == EEM_VOXEL_ATOMIC && !pair) {
    const int64_t per_block = (int64_t)kVoteThreads * kVoteEventsPerThread;
    const int64_t chunks = ceil_div(max_events_per_window, per_block);
    const int lanes = (int)std::min<int64_t>(time_lanes(64), chunks);   // short windows: no CTAs without a chunk
    dim3 g((unsigned)(ceil_div(chunks, lanes) * lanes), (unsigned)n_windows);
    voxel_vote_atomic_kernel<Src><<<g, kVoteThreads, 0, stream>>>(events, offsets, num_bins, height,
                                                             width, lanes, grid, dropped);
    EEM_CHECK_LAUNCH("voxel_vote_atomic_kernel");
  } else if (pair) {
    float* scratch = reinterpret_cast<float*>(ws_mode);
    EEM_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)total_vox * 2 * sizeof(float), stream));
    const int64_t per_block = (int64_t)kVoteThreads * kVoteEventsPerThread;
    const int64_t chunks = ceil_div(max_events_per_window, per_block);
    // the scratch is 2x the grid: keep the set of concurrently active bin planes small enough for L2
    const int lanes = (int)std::min<int64_t>(time_lanes(4), chunks);
    dim3 g((unsigned)(ceil_div(chunks, lanes) * lanes), (unsigned)n_windows);
    voxel_vote_pair_kernel<Src><<<g, kVoteThreads, 0, stream>>>(events, offsets, num_bins, height, width, lanes, scratch, dropped);
    EEM_CHECK_LAUNCH("voxel_vote_pair_kernel");
    const int64_t HW = (int64_t)height * width;
    dim3 gc((unsigned)stat_blocks(HW * 4, n_windows), (unsigned)n_windows);
    if (normalize) {
      const StatLayout L = stat_layout(n_windows);
      StatPartial* partials = reinterpret_cast<StatPartial*>(ws_stats + L.partials);
      unsigned int* tickets = reinterpret_cast<unsigned int*>(ws_stats + L.tickets);
      float* mean_std = reinterpret_cast<float*>(ws_stats + L.mean_std);
      EEM_CHECK_CUDA(cudaMemsetAsync(tickets, 0, (size_t)n_windows * sizeof(unsigned int), stream));
      voxel_combine_kernel<true><<<gc, kStatThreads, 0, stream>>>(reinterpret_cast<const float2*>(scratch), num_bins, HW, grid,
                                                                  partials, tickets, mean_std, stats_out);
      EEM_CHECK_LAUNCH("voxel_combine_kernel");
      dim3 ga((unsigned)stat_blocks(vox, n_windows), (unsigned)n_windows);
      voxel_apply_kernel<<<ga, 256, 0, stream>>>(grid, vox, mean_std);
      EEM_CHECK_LAUNCH("voxel_apply_kernel");
      return EEM_OK;
    }
    voxel_combine_kernel<false><<<gc, kStatThreads, 0, stream>>>(reinterpret_cast<const float2*>(scratch), num_bins, HW, grid,
                                                                 nullptr, nullptr, nullptr, nullptr);
    EEM_CHECK_LAUNCH("voxel_combine_kernel");
    return EEM_OK;
  } else {
    // deterministic: vote pairs -> stable LSD radix sort by global voxel key -> sequential sums
    if (total_vox >= (int64_t)0xffffffffLL)
      return fail(EEM_ERR_UNSUPPORTED,
                  "eem_voxelize(deterministic): %lld voxels in one call exceed the 32-bit key space; split the batch",
                  (long long)total_vox);
    if (2 * n_total >= (int64_t)0x7fffffffLL)
      return fail(EEM_ERR_UNSUPPORTED, "eem_voxelize(deterministic): too many events in one call (%lld)", (long long)n_total);
    const DetLayout L = det_layout(n_total);
    uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(ws_mode + L.keys_a), reinterpret_cast<uint32_t*>(ws_mode + L.keys_b)};
    float* vals[2] = {reinterpret_cast<float*>(ws_mode + L.vals_a), reinterpret_cast<float*>(ws_mode + L.vals_b)};
    uint32_t* counts = reinterpret_cast<uint32_t*>(ws_mode + L.counts);
    uint32_t* block_sums = reinterpret_cast<uint32_t*>(ws_mode + L.block_sums);
    const uint32_t invalid_key = (uint32_t)total_vox;
    {
      dim3 g((unsigned)ceil_div(max_events_per_window, kVoteThreads), (unsigned)n_windows);
      voxel_vote_pairs_kernel<Src><<<g, kVoteThreads, 0, stream>>>(events, offsets, num_bins, height, width,
                                                              n_total, invalid_key, keys[0], vals[0], dropped);
      EEM_CHECK_LAUNCH("voxel_vote_pairs_kernel");
    }
    const int key_bits = bit_length((uint64_t)invalid_key);
    const int passes = (key_bits + 7) / 8;
    const unsigned sort_blocks = (unsigned)ceil_div(L.n_segs, kSortWarpsPerBlock);
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
      const int shift = 8 * p;
      radix_count_kernel<<<sort_blocks, kSortWarpsPerBlock * 32, 0, stream>>>(keys[cur], L.n_votes, shift, L.n_segs, counts);
      EEM_CHECK_LAUNCH("radix_count_kernel");
      scan_block_sums_kernel<<<(unsigned)L.n_scan_blocks, kScanThreads, 0, stream>>>(counts, L.n_counts, block_sums);
      EEM_CHECK_LAUNCH("scan_block_sums_kernel");
      scan_of_block_sums_kernel<<<1, kScanThreads, 0, stream>>>(block_sums, L.n_scan_blocks);
      EEM_CHECK_LAUNCH("scan_of_block_sums_kernel");
      scan_apply_kernel<<<(unsigned)L.n_scan_blocks, kScanThreads, 0, stream>>>(counts, L.n_counts, block_sums);
      EEM_CHECK_LAUNCH("scan_apply_kernel");
      radix_scatter_kernel<<<sort_blocks, kSortWarpsPerBlock * 32, 0, stream>>>(
          keys[cur], vals[cur], L.n_votes, shift, L.n_segs, counts, keys[cur ^ 1], vals[cur ^ 1]);
      EEM_CHECK_LAUNCH("radix_scatter_kernel");
      cur ^= 1;
    }
    segmented_sum_kernel<<<(unsigned)ceil_div(L.n_votes, 256), 256, 0, stream>>>(keys[cur], vals[cur], L.n_votes,
                                                                               invalid_key, grid);
    EEM_CHECK_LAUNCH("segmented_sum_kernel");
  }
  if (normalize) return launch_normalize(grid, n_windows, vox, stats_out, ws_stats, stream);
  return EEM_OK;
}


extern "C" {

size_t eem_voxelize_workspace_bytes(int64_t n_total, int n_windows, int num_bins, int height,
                                    int width, int mode, int normalize) {
  if (n_total < 0 || n_windows <= 0 || num_bins <= 0 || height <= 0 || width <= 0) return 0;
  const int64_t total_vox = (int64_t)n_windows * num_bins * height * width;
  size_t bytes = normalize ? stat_layout(n_windows).total : 0;
  const int64_t vox = (int64_t)num_bins * height * width;
  if (tiled_path_enabled(mode) && tiled_path_fits(num_bins, height, width, vox)) {
    bytes = normalize ? stat_layout(n_windows, tiled_stat_parts(num_bins, height, width)).total : 0;
    const TilePlan tp = tile_plan(height, width, n_total);
    return bytes + tile_workspace(n_total, n_windows, num_bins, tp.n_tiles, ceil_div(n_total, kBinChunk) + 1).total;
  }
  if (mode == EEM_VOXEL_DETERMINISTIC) bytes += det_layout(n_total).total;
  else if (use_pair_path(n_total, total_vox)) bytes += align_up((size_t)total_vox * 2 * sizeof(float), 256);
  else if (use_interleaved_path(num_bins))
    bytes += align_up((size_t)n_windows * height * width * ((num_bins + 1) & ~1) * sizeof(float), 256);
  return bytes;
}

int eem_voxelize(const double* events, const int64_t* offsets, int n_windows, int64_t n_total,
                 int64_t max_events_per_window, int num_bins, int height, int width, int mode,
                 int normalize, float* grid, int64_t* dropped, double* stats_out, void* workspace,
                 size_t workspace_bytes, eem_stream_t stream) {
  return eem_voxelize_scaled(events, 1.0, offsets, n_windows, n_total, max_events_per_window, num_bins, height, width, mode,
                             normalize, grid, dropped, stats_out, workspace, workspace_bytes, stream);
}

int eem_voxelize_scaled(const double* events, double timestamp_multiplier, const int64_t* offsets, int n_windows, int64_t n_total,
                        int64_t max_events_per_window, int num_bins, int height, int width, int mode,
                        int normalize, float* grid, int64_t* dropped, double* stats_out, void* workspace,
                        size_t workspace_bytes, eem_stream_t stream) {
  EEM_CHECK_ARG(n_total <= 0 || events != nullptr, "eem_voxelize: NULL events");
  EEM_CHECK_ARG(timestamp_multiplier > 0.0, "eem_voxelize_scaled: timestamp_multiplier must be > 0");
  EEM_CHECK_ALIGNED(events, 32);
  return voxelize_impl(RowSource{events, timestamp_multiplier}, offsets, n_windows, n_total, max_events_per_window, num_bins, height, width,
                       mode, normalize, grid, dropped, stats_out, workspace, workspace_bytes, stream);
}

int eem_voxelize_soa(const void* t, int t_is_ns, const int16_t* x, const int16_t* y, const int8_t* p,
                     const int64_t* offsets, int n_windows, int64_t n_total, int64_t max_events_per_window,
                     int num_bins, int height, int width, int mode, int normalize, float* grid, int64_t* dropped,
                     double* stats_out, void* workspace, size_t workspace_bytes, eem_stream_t stream) {
  EEM_CHECK_ARG(n_total <= 0 || (t && x && y && p), "eem_voxelize_soa: NULL event column");
  EEM_CHECK_ALIGNED(t, 8);
  EEM_CHECK_ALIGNED(x, 2);
  EEM_CHECK_ALIGNED(y, 2);
  return voxelize_impl(SoaSource{t, x, y, p, t_is_ns ? 1 : 0}, offsets, n_windows, n_total, max_events_per_window,
                       num_bins, height, width, mode, normalize, grid, dropped, stats_out, workspace, workspace_bytes, stream);
}

size_t eem_voxel_normalize_workspace_bytes(int n_windows, int64_t voxels_per_window) {
  if (n_windows <= 0 || voxels_per_window <= 0) return 0;
  return stat_layout(n_windows).total;
}

int eem_voxel_normalize(float* grid, int n_windows, int64_t voxels_per_window, double* stats_out,
                        void* workspace, size_t workspace_bytes, eem_stream_t stream_) {
  EEM_CHECK_ARG(grid != nullptr, "eem_voxel_normalize: NULL grid");
  EEM_CHECK_ARG(n_windows > 0 && voxels_per_window > 0, "eem_voxel_normalize: sizes must be > 0");
  EEM_CHECK_ARG(n_windows <= 65535, "eem_voxel_normalize: more than 65535 windows in one call");
  const size_t need = eem_voxel_normalize_workspace_bytes(n_windows, voxels_per_window);
  if (workspace == nullptr || workspace_bytes < need)
    return fail(EEM_ERR_WORKSPACE, "eem_voxel_normalize: workspace of %zu bytes required, got %zu", need, workspace_bytes);
  EEM_CHECK_ALIGNED(workspace, 256);
  return launch_normalize(grid, n_windows, voxels_per_window, stats_out, static_cast<char*>(workspace), as_stream(stream_));
}

}  // extern "C"
