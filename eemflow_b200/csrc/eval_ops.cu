// Loader-side and evaluation-side helpers that sit right next to the hot path:
//   * per-pixel event mask of a window          loader/MVSEC.py:133-142 (np.histogram2d > 0)
//   * sum of a voxel grid over its bins          loader/HREM.py:238-239  (np.sum(volume, axis=0))
//   * masked end-point-error statistics          test_mvsec.py:291-346   (Test.flow_error)
//   * dense flow -> 16x16 mesh flow              loader/HREM.py:41-101   (motion_propagate)
// so that an evaluation loop reads back 40 bytes per sample instead of the full-resolution flow maps, and the
// HREM loader's per-sample Python double loop becomes one launch per batch.
#include "common.cuh"

namespace eem {
namespace {

// ---- event mask: mask[w][y][x] = 1 iff some event of window w falls into pixel bin (x, y) -------------------
// np.histogram2d(bins=(W,H), range=[[0,W],[0,H]]): bin = floor(v) for v in [0, S), v == S joins the last bin,
// everything else (and NaN) is ignored.
__device__ __forceinline__ int hist_bin(double v, int size) {
  if (!(v >= 0.0 && v <= (double)size)) return -1;
  const int b = (int)v;
  return b == size ? size - 1 : b;
}

__global__ void __launch_bounds__(256)
event_mask_kernel(const double* __restrict__ events, const int64_t* __restrict__ offsets, int n_windows, int H, int W,
                  uint8_t* __restrict__ mask) {
  const int win = blockIdx.y;
  const int64_t lo = offsets[win], hi = offsets[win + 1];
  uint8_t* m = mask + (int64_t)win * H * W;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 tx = *reinterpret_cast<const double2*>(events + 4 * i);        // (t, x)
    const double2 yp = *reinterpret_cast<const double2*>(events + 4 * i + 2);    // (y, p)
    const double x = tx.y, y = yp.x;
    const int bx = hist_bin(x, W), by = hist_bin(y, H);
    if (bx >= 0 && by >= 0) m[(int64_t)by * W + bx] = 1;     // idempotent store: races are benign
  }
}

// ---- sum over bins, sequentially in fp32 like numpy's reduction over the outer axis ------------------------
__global__ void __launch_bounds__(256)
bin_sum_kernel(const float* __restrict__ grid, int64_t n, int nb, int64_t plane, float* __restrict__ out) {
  const int64_t total = n * plane;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = i / plane, px = i - w * plane;
    const float* g = grid + w * nb * plane + px;
    float acc = ld_stream(g);
    for (int b = 1; b < nb; ++b) acc = __fadd_rn(acc, ld_stream(g + (int64_t)b * plane));
    out[i] = acc;
  }
}

// ---- flow_error ---------------------------------------------------------------------------------------------
// stats[b] = { n_points, #(EE < 1), #(EE < 3 or EE < 0.1*|gt|), sum EE, sum |gt| }  (doubles)
__device__ __forceinline__ float norm2(float a, float b) {     // np.linalg.norm(axis=-1) of a float32 pair
  return __fsqrt_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
}

__global__ void __launch_bounds__(256)
flow_error_kernel(const float* __restrict__ gt, const float* __restrict__ pred, const float* __restrict__ event_img,
                  int H, int W, int max_row, double* __restrict__ stats) {
  const int b = blockIdx.y;
  const int64_t plane = (int64_t)H * W, rows = (int64_t)max_row * W;
  const float* gu = gt + (int64_t)b * 2 * plane;
  const float* gv = gu + plane;
  const float* pu = pred + (int64_t)b * 2 * plane;
  const float* pv = pu + plane;
  const float* ev = event_img ? event_img + (int64_t)b * plane : nullptr;
  double acc[5] = {0, 0, 0, 0, 0};
  auto one = [&](float u, float v, float qu, float qv, float e) {
    const float mag = norm2(u, v);
    const bool ok = !isinf(u) && !isinf(v) && mag > 0.f && e > 0.f;
    if (!ok) return;
    const float ee = norm2(__fsub_rn(u, qu), __fsub_rn(v, qv));
    acc[0] += 1.0;
    acc[1] += ee < 1.0f ? 1.0 : 0.0;
    acc[2] += (ee < 3.0f || ee < __fmul_rn(0.1f, mag)) ? 1.0 : 0.0;
    acc[3] += (double)ee;
    acc[4] += (double)mag;
  };
  // four pixels per iteration with 16-byte loads when every plane starts 16-byte aligned (the sums are order-free
  // in fp64 either way; counts are exact)
  const bool vec = ((plane & 3) == 0) && ((rows & 3) == 0) && (((uintptr_t)gt | (uintptr_t)pred | (uintptr_t)event_img) & 15) == 0;
  if (vec) {
    const float4 ones = make_float4(1.f, 1.f, 1.f, 1.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows / 4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(gu) + i), v = __ldg(reinterpret_cast<const float4*>(gv) + i);
      const float4 qu = __ldg(reinterpret_cast<const float4*>(pu) + i), qv = __ldg(reinterpret_cast<const float4*>(pv) + i);
      const float4 e = ev ? __ldg(reinterpret_cast<const float4*>(ev) + i) : ones;
      one(u.x, v.x, qu.x, qv.x, e.x);
      one(u.y, v.y, qu.y, qv.y, e.y);
      one(u.z, v.z, qu.z, qv.z, e.z);
      one(u.w, v.w, qu.w, qv.w, e.w);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x)
      one(gu[i], gv[i], pu[i], pv[i], ev ? ev[i] : 1.f);
  }
  __shared__ double red[5][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    if (v != 0.0) atomicAdd(stats + b * 5 + threadIdx.x, v);
  }
}

// ---- motion_propagate ---------------------------------------------------------------------------------------
// One CTA per sample, one thread per mesh vertex.  Stage 1: median (element n/2 of the sorted list) of the
// 4*radius samples around the vertex, indices clamped into the frame.  Stage 2: 5x5 median over the vertex
// mesh with replicated borders.  Pure selection, so results are exactly the reference's values.
constexpr int kMaxMesh = 32, kMaxSamples = 32;

__device__ __forceinline__ void insertion_sort(float* a, int n) {
  for (int i = 1; i < n; ++i) {
    const float v = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
    a[j + 1] = v;
  }
}

__global__ void __launch_bounds__(kMaxMesh * kMaxMesh)
motion_propagate_kernel(const float* __restrict__ fflow, int H, int W, int mesh, int radius, float* __restrict__ out) {
  __shared__ float s[2][kMaxMesh * kMaxMesh];
  const float* f = fflow + (int64_t)blockIdx.x * H * W * 2;
  float* o = out + (int64_t)blockIdx.x * 2 * mesh * mesh;
  const int t = threadIdx.x;
  const int i = t / mesh, j = t - i * mesh;
  const bool live = t < mesh * mesh;
  const int mesh_cols = W / mesh, mesh_rows = H / mesh;
  if (live) {
    float us[kMaxSamples], vs[kMaxSamples];
    int n = 0;
    for (int r = 0; r < radius; ++r) {
      const int off_x = r * mesh_rows / 2, off_y = r * mesh_cols / 2;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int pi = mesh_rows * i + ((q & 2) ? -off_x : off_x);
        int pj = mesh_cols * j + ((q & 1) ? -off_y : off_y);
        pi = min(max(pi, 0), H - 1);
        pj = min(max(pj, 0), W - 1);
        const float2 uv = *reinterpret_cast<const float2*>(f + ((int64_t)pi * W + pj) * 2);
        us[n] = uv.x;
        vs[n] = uv.y;
        ++n;
      }
    }
    insertion_sort(us, n);
    insertion_sort(vs, n);
    s[0][t] = n ? us[n / 2] : 0.f;
    s[1][t] = n ? vs[n / 2] : 0.f;
  }
  __syncthreads();
  if (live) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float win[25];
#pragma unroll
      for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx)
          win[(dy + 2) * 5 + dx + 2] = s[c][min(max(i + dy, 0), mesh - 1) * mesh + min(max(j + dx, 0), mesh - 1)];
      insertion_sort(win, 25);
      o[c * mesh * mesh + t] = win[12];
    }
  }
}

}  // namespace
}  // namespace eem

using namespace eem;

extern "C" {

int eem_event_mask(const double* events, const int64_t* offsets, int n_windows, int64_t max_events_per_window,
                   int height, int width, uint8_t* mask, eem_stream_t stream_) {
  EEM_CHECK_ARG(events && offsets && mask, "eem_event_mask: NULL pointer");
  EEM_CHECK_ARG(n_windows > 0 && n_windows <= 65535 && height > 0 && width > 0, "eem_event_mask: bad sizes");
  EEM_CHECK_ALIGNED(events, 16);
  cudaStream_t stream = as_stream(stream_);
  EEM_CHECK_CUDA(cudaMemsetAsync(mask, 0, (size_t)n_windows * height * width, stream));
  if (max_events_per_window <= 0) return EEM_OK;
  int64_t bx = ceil_div(max_events_per_window, 256 * 4);
  const int64_t cap = ceil_div((int64_t)sm_count() * 8, n_windows);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  event_mask_kernel<<<dim3((unsigned)bx, (unsigned)n_windows), 256, 0, stream>>>(events, offsets, n_windows, height, width, mask);
  EEM_CHECK_LAUNCH("event_mask_kernel");
  return EEM_OK;
}

int eem_voxel_bin_sum(const float* grid, int64_t n_windows, int num_bins, int height, int width, float* out,
                      eem_stream_t stream_) {
  EEM_CHECK_ARG(grid && out, "eem_voxel_bin_sum: NULL pointer");
  EEM_CHECK_ARG(n_windows > 0 && num_bins > 0 && height > 0 && width > 0, "eem_voxel_bin_sum: sizes must be > 0");
  const int64_t plane = (int64_t)height * width, total = n_windows * plane;
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  bin_sum_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream_)>>>(grid, n_windows, num_bins, plane, out);
  EEM_CHECK_LAUNCH("bin_sum_kernel");
  return EEM_OK;
}

int eem_flow_error(const float* flow_gt, const float* flow_pred, const float* event_img, int B, int height, int width,
                   int max_row, double* stats, eem_stream_t stream_) {
  EEM_CHECK_ARG(flow_gt && flow_pred && stats, "eem_flow_error: NULL pointer");
  EEM_CHECK_ARG(B > 0 && B <= 65535 && height > 0 && width > 0, "eem_flow_error: bad sizes");
  EEM_CHECK_ARG(max_row >= 0 && max_row <= height, "eem_flow_error: max_row must be in [0, height]");
  cudaStream_t stream = as_stream(stream_);
  EEM_CHECK_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * 5 * sizeof(double), stream));
  if (max_row == 0) return EEM_OK;
  int64_t bx = ceil_div((int64_t)max_row * width, 256 * 4);
  const int64_t cap = ceil_div((int64_t)sm_count() * 8, B);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  flow_error_kernel<<<dim3((unsigned)bx, (unsigned)B), 256, 0, stream>>>(flow_gt, flow_pred, event_img, height, width, max_row, stats);
  EEM_CHECK_LAUNCH("flow_error_kernel");
  return EEM_OK;
}

int eem_motion_propagate(const float* fflow, int B, int height, int width, int mesh_size, int radius, float* mesh,
                         eem_stream_t stream_) {
  EEM_CHECK_ARG(fflow && mesh, "eem_motion_propagate: NULL pointer");
  EEM_CHECK_ARG(B > 0 && height > 0 && width > 0, "eem_motion_propagate: sizes must be > 0");
  EEM_CHECK_ARG(mesh_size > 0 && mesh_size <= kMaxMesh, "eem_motion_propagate: mesh_size must be in [1,%d]", kMaxMesh);
  EEM_CHECK_ARG(radius >= 0 && 4 * radius <= kMaxSamples, "eem_motion_propagate: radius must be in [0,%d]", kMaxSamples / 4);
  EEM_CHECK_ALIGNED(fflow, 8);
  const int threads = (int)align_up((size_t)mesh_size * mesh_size, 32);
  motion_propagate_kernel<<<(unsigned)B, threads, 0, as_stream(stream_)>>>(fflow, height, width, mesh_size, radius, mesh);
  EEM_CHECK_LAUNCH("motion_propagate_kernel");
  return EEM_OK;
}

}  // extern "C"
