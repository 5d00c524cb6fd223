// K7 backward warp (+ validity mask, + fused CDC blend), K8 bilinear resize, K9 replicate pad.
// All are HBM-bound gathers: thread = output pixel with x fastest (128 B coalesced stores),
// texture-free bilinear interpolation, sample geometry computed once per pixel and reused over
// the channel loop.
//
// Coordinate conventions are the reference's, kept bug-for-bug (trap 5):
//   g  = 2*(px+u)/max(W-1,1) - 1                          every warp variant (tools.py:2289-2290)
//   EXACT   ix = (g+1)/2*(W-1)       grid_sample(align_corners=True)   EEMFlow+.py:145-148
//   HALFPIX ix = ((g+1)*W-1)/2       grid_sample default (False)       tools.py:2295, cdc_utils.py:71
// Bilinear weights and accumulation order follow ATen's grid_sampler_2d (nw, ne, sw, se).
#include <cstdlib>

#include "common.cuh"

namespace eem {
namespace {

struct Bilin {
  int x0, y0;             // north-west tap
  float nw, ne, sw, se;   // weights
  bool in_nw, in_ne, in_sw, in_se;
};

// Un-normalisation exactly as ATen's CPU grid sampler evaluates it (GridSamplerKernel.cpp,
// ComputeLocation): align_corners=True -> (g+1) * ((S-1)/2); align_corners=False ->
// fma(g+1, S/2, -0.5) (one rounding).  Matching the rounding matters: the reference thresholds the
// interpolated ones-tensor at 1.0 (cdc_utils.py:77), and for an interior sample that sum is
// 1 +- 1 ulp, so the 0/1 validity mask is decided by these roundings.
__device__ __forceinline__ float unnormalize(float g, int size, int convention) {
  const float g1 = __fadd_rn(g, 1.0f);
  if (convention == EEM_WARP_EXACT) return __fmul_rn(g1, (float)(size - 1) * 0.5f);
  return __fmaf_rn(g1, (float)size * 0.5f, -0.5f);
}

__device__ __forceinline__ void fill_bilin(Bilin& s, float ix, float iy, int H, int W) {
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  // keep the integer conversion defined for wild flows; such samples are fully out of bounds
  s.x0 = (int)fminf(fmaxf(fx0, -2.0e9f), 2.0e9f);
  s.y0 = (int)fminf(fmaxf(fy0, -2.0e9f), 2.0e9f);
  // ATen compute_interp_params: w = x - floor(x), e = 1 - w, n = y - floor(y), s = 1 - n
  const float w = __fsub_rn(ix, fx0), e = __fsub_rn(1.0f, w);
  const float n = __fsub_rn(iy, fy0), so = __fsub_rn(1.0f, n);
  s.nw = __fmul_rn(so, e);
  s.ne = __fmul_rn(so, w);
  s.sw = __fmul_rn(n, e);
  s.se = __fmul_rn(n, w);
  const bool xin0 = s.x0 >= 0 && s.x0 < W, xin1 = s.x0 + 1 >= 0 && s.x0 + 1 < W;
  const bool yin0 = s.y0 >= 0 && s.y0 < H, yin1 = s.y0 + 1 >= 0 && s.y0 + 1 < H;
  const bool finite = (fx0 == fx0) && (fy0 == fy0) && fabsf(fx0) < 1.0e9f && fabsf(fy0) < 1.0e9f;
  s.in_nw = finite && xin0 && yin0;
  s.in_ne = finite && xin1 && yin0;
  s.in_sw = finite && xin0 && yin1;
  s.in_se = finite && xin1 && yin1;
}

__device__ __forceinline__ Bilin make_bilin(float px, float py, int H, int W, int convention) {
  // vgrid = 2.0 * (grid + flo) / max(W-1, 1) - 1.0: multiply, true division, subtract (tools.py:2289-2290)
  const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, px), (float)max(W - 1, 1)), 1.0f);
  const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, py), (float)max(H - 1, 1)), 1.0f);
  Bilin s;
  fill_bilin(s, unnormalize(gx, W, convention), unnormalize(gy, H, convention), H, W);
  return s;
}

// nw*v_nw + ne*v_ne + sw*v_sw + se*v_se, out-of-bounds taps read as 0 (ATen masks the gather).
__device__ __forceinline__ float sample(const float* __restrict__ plane, const Bilin& s, int W) {
  const float* p = plane + (int64_t)s.y0 * W + s.x0;
  const float v_nw = s.in_nw ? __ldg(p) : 0.f;
  const float v_ne = s.in_ne ? __ldg(p + 1) : 0.f;
  const float v_sw = s.in_sw ? __ldg(p + W) : 0.f;
  const float v_se = s.in_se ? __ldg(p + W + 1) : 0.f;
  return fmaf(v_se, s.se, fmaf(v_sw, s.sw, fmaf(v_ne, s.ne, v_nw * s.nw)));
}

// grid_sample of an all-ones tensor: the weights of the in-bounds taps, summed nw, ne, sw, se.
__device__ __forceinline__ float ones_sample(const Bilin& s) {
  float acc = s.in_nw ? s.nw : 0.f;
  acc = __fadd_rn(acc, s.in_ne ? s.ne : 0.f);
  acc = __fadd_rn(acc, s.in_sw ? s.sw : 0.f);
  acc = __fadd_rn(acc, s.in_se ? s.se : 0.f);
  return acc;
}

constexpr int kWarpChunk = 8;  // channels per thread; grid.z = B * ceil(C / kWarpChunk)

__global__ void __launch_bounds__(256)
backwarp_kernel(const float* __restrict__ x, const float* __restrict__ flow, int B, int C, int H, int W,
                int convention, int mask_mode, float* __restrict__ out, float* __restrict__ mask_out) {
  const int px = blockIdx.x * 32 + (threadIdx.x & 31);
  const int py = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (px >= W || py >= H) return;
  const int chunks = (C + kWarpChunk - 1) / kWarpChunk;
  const int b = blockIdx.z / chunks, c0 = (blockIdx.z % chunks) * kWarpChunk;
  const int64_t plane = (int64_t)H * W, pix = (int64_t)py * W + px;
  const float u = flow[((int64_t)b * 2 + 0) * plane + pix];
  const float v = flow[((int64_t)b * 2 + 1) * plane + pix];
  const Bilin s = make_bilin((float)px + u, (float)py + v, H, W, convention);
  float m = 1.f;
  if (mask_mode != EEM_MASK_NONE) {
    const float ms = ones_sample(s);
    m = (mask_mode == EEM_MASK_GE1) ? (ms >= 1.0f ? 1.f : 0.f) : (ms < 0.9999f ? 0.f : 1.f);
    if (mask_out != nullptr && c0 == 0) mask_out[(int64_t)b * plane + pix] = m;
  }
  // fixed trip count: the 4 x kWarpChunk tap loads of a thread are issued back to back
  float r[kWarpChunk];
#pragma unroll
  for (int k = 0; k < kWarpChunk; ++k) {
    const int c = min(c0 + k, C - 1);
    r[k] = sample(x + ((int64_t)b * C + c) * plane, s, W);
  }
#pragma unroll
  for (int k = 0; k < kWarpChunk; ++k) {
    const int c = c0 + k;
    if (c < C) st_stream(out + ((int64_t)b * C + c) * plane + pix, mask_mode != EEM_MASK_NONE ? r[k] * m : r[k]);
  }
}

__global__ void __launch_bounds__(256)
warp_blend_kernel(const float* __restrict__ flow_init, const float* __restrict__ inter, const float* __restrict__ mk,
                  int B, int H, int W, float* __restrict__ out) {
  const int px = blockIdx.x * 32 + (threadIdx.x & 31);
  const int py = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (px >= W || py >= H) return;
  const int64_t plane = (int64_t)H * W, pix = (int64_t)py * W + px;
  const float u = inter[((int64_t)b * 2 + 0) * plane + pix];
  const float v = inter[((int64_t)b * 2 + 1) * plane + pix];
  const Bilin s = make_bilin((float)px + u, (float)py + v, H, W, EEM_WARP_HALFPIX);
  const float m = mk[(int64_t)b * plane + pix];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int64_t off = ((int64_t)b * 2 + c) * plane;
    const float w = sample(flow_init + off, s, W);
    const float f = flow_init[off + pix];
    out[off + pix] = w * (1.f - m) + f * m;
  }
}

// ATen area_pixel_compute_source_index (UpSample.h), fp32.
__device__ __forceinline__ void source_index(int dst, int in_size, int out_size, int align_corners,
                                             int& i0, int& i1, float& l0, float& l1) {
  float src;
  if (align_corners) {
    const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    src = scale * (float)dst;
  } else {
    const float scale = (float)in_size / (float)out_size;
    src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// One rounding sequence for every resize kernel (they are tested to agree bit for bit):
//   (ly0 * (la * a + lb * b) + ly1 * (la * c + lb * d)) * scale   with the products la * a, ly0 * top rounded first.
__device__ __forceinline__ float hlerp(float la, float a, float lb, float b) { return __fmaf_rn(lb, b, __fmul_rn(la, a)); }
__device__ __forceinline__ float vlerp(float ly0, float top, float ly1, float bot, float sc) {
  return __fmul_rn(__fmaf_rn(ly1, bot, __fmul_rn(ly0, top)), sc);
}

// Thread = V horizontally adjacent output pixels of one row (V = 4/2/1 by the divisibility of W and the
// alignment of `out`): the vertical taps/weights are shared, the V horizontal tap pairs are computed once, and
// the plane loop issues 4V independent (L1-resident) loads and one V-wide streaming store per plane.
template <int V>
__global__ void __launch_bounds__(256)
bilinear_resize_kernel(const float* __restrict__ in, int B, int C, int h, int w, float* __restrict__ out,
                       int H, int W, int align_corners, float scale0, float scale1, float scale_rest) {
  const int X = (blockIdx.x * 32 + (threadIdx.x & 31)) * V;
  const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (X >= W || Y >= H) return;
  int xa[V], xb[V], y0, y1;
  float la[V], lb[V], ly0, ly1;
#pragma unroll
  for (int i = 0; i < V; ++i) source_index(X + i, w, W, align_corners, xa[i], xb[i], la[i], lb[i]);
  source_index(Y, h, H, align_corners, y0, y1, ly0, ly1);
  const int64_t ip = (int64_t)h * w, op = (int64_t)H * W;
  // plane loop with running pointers and a running channel id (no div/mod, no 64-bit multiplies)
  const int gz = gridDim.z, BC = B * C;
  const float* r0 = in + (int64_t)blockIdx.z * ip + y0 * w;
  const float* r1 = in + (int64_t)blockIdx.z * ip + y1 * w;
  float* o = out + (int64_t)blockIdx.z * op + (int64_t)Y * W + X;
  const int64_t s_step = (int64_t)gz * ip, o_step = (int64_t)gz * op;
  int c = blockIdx.z % C;
  const int c_step = gz % C;
#pragma unroll 2
  for (int bc = blockIdx.z; bc < BC; bc += gz) {
    const float sc = c == 0 ? scale0 : (c == 1 ? scale1 : scale_rest);
    float v[V];
#pragma unroll
    for (int i = 0; i < V; ++i)
      v[i] = vlerp(ly0, hlerp(la[i], __ldg(r0 + xa[i]), lb[i], __ldg(r0 + xb[i])),
                   ly1, hlerp(la[i], __ldg(r1 + xa[i]), lb[i], __ldg(r1 + xb[i])), sc);
    if constexpr (V == 4) {
      st_stream4(o, make_float4(v[0], v[1], v[2], v[3]));
    } else if constexpr (V == 2) {
      asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(o), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
      st_stream(o, v[0]);
    }
    r0 += s_step;
    r1 += s_step;
    o += o_step;
    c += c_step;
    if (c >= C) c -= C;
  }
}

__global__ void __launch_bounds__(256)
scale_uv_kernel(float* __restrict__ flow, int B, int C, int64_t plane, float s0, float s1) {
  const int64_t total = (int64_t)B * 2 * plane;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / (2 * plane), r = i % (2 * plane);
    const int c = (int)(r / plane);
    float* p = flow + ((int64_t)b * C + c) * plane + (r % plane);
    *p = *p * (c == 0 ? s0 : s1);
  }
}

// ---- fused EEMFlow-level chains -------------------------------------------------------------------------------------
// Per pyramid level EEMFlow_cdc runs  upsample2d_flow_as -> WarpingLayer_no_div  (cdc_utils.py:156-160) and, after its
// estimator convolutions,  CDC blend -> EEMFlow_cdc.warp  (cdc_utils.py:173, EEMFlow+.py:181...): in both pairs the
// second op only needs the first op's per-pixel flow.  The fused kernels compute that flow per thread with the SAME
// expressions as bilinear_resize_kernel / warp_blend_kernel, write it out once (the channel-chunk-0 thread) and warp
// the feature map with it: one launch and one flow round trip less per pair.
struct FusedWarpParams {
  const float* x;            // [B,C,H,W] map to warp
  float* out;                // [B,C,H,W]
  float* flow_out;           // [B,2,H,W] the flow the warp used
  // kUpsample: coarse flow [B,2,h,w] resized with align_corners=True, channels scaled by s0 / s1; HALFPIX warp + >= 1.0 mask
  const float* coarse;
  int h, w;
  float s0, s1;
  // !kUpsample: flow = torch_warp(flow_init, inter) * (1 - m) + flow_init * m; EXACT warp, no mask
  const float* flow_init;
  const float* inter;
  const float* mk;
  int B, C, H, W;
};

template <bool kUpsample>
__global__ void __launch_bounds__(256)
fused_flow_warp_kernel(const __grid_constant__ FusedWarpParams p) {
  const int px = blockIdx.x * 32 + (threadIdx.x & 31);
  const int py = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int H = p.H, W = p.W, C = p.C;
  if (px >= W || py >= H) return;
  const int chunks = (C + kWarpChunk - 1) / kWarpChunk;
  const int b = blockIdx.z / chunks, c0 = (blockIdx.z % chunks) * kWarpChunk;
  const int64_t plane = (int64_t)H * W, pix = (int64_t)py * W + px;
  float u, v;
  if (kUpsample) {
    int xa, xb, y0, y1;
    float la, lb, ly0, ly1;
    source_index(px, p.w, W, 1, xa, xb, la, lb);
    source_index(py, p.h, H, 1, y0, y1, ly0, ly1);
    const int64_t ip = (int64_t)p.h * p.w;
    const float* r0 = p.coarse + (int64_t)b * 2 * ip + y0 * p.w;
    const float* r1 = p.coarse + (int64_t)b * 2 * ip + y1 * p.w;
    u = vlerp(ly0, hlerp(la, __ldg(r0 + xa), lb, __ldg(r0 + xb)), ly1, hlerp(la, __ldg(r1 + xa), lb, __ldg(r1 + xb)), p.s0);
    r0 += ip;
    r1 += ip;
    v = vlerp(ly0, hlerp(la, __ldg(r0 + xa), lb, __ldg(r0 + xb)), ly1, hlerp(la, __ldg(r1 + xa), lb, __ldg(r1 + xb)), p.s1);
  } else {
    const float iu = p.inter[((int64_t)b * 2 + 0) * plane + pix];
    const float iv = p.inter[((int64_t)b * 2 + 1) * plane + pix];
    const Bilin sb = make_bilin((float)px + iu, (float)py + iv, H, W, EEM_WARP_HALFPIX);
    const float m = p.mk[(int64_t)b * plane + pix];
    const int64_t off0 = ((int64_t)b * 2 + 0) * plane, off1 = off0 + plane;
    u = sample(p.flow_init + off0, sb, W) * (1.f - m) + p.flow_init[off0 + pix] * m;
    v = sample(p.flow_init + off1, sb, W) * (1.f - m) + p.flow_init[off1 + pix] * m;
  }
  if (c0 == 0) {
    p.flow_out[((int64_t)b * 2 + 0) * plane + pix] = u;
    p.flow_out[((int64_t)b * 2 + 1) * plane + pix] = v;
  }
  const Bilin s = make_bilin((float)px + u, (float)py + v, H, W, kUpsample ? EEM_WARP_HALFPIX : EEM_WARP_EXACT);
  float m = 1.f;
  if (kUpsample) m = ones_sample(s) >= 1.0f ? 1.f : 0.f;
  float r[kWarpChunk];
#pragma unroll
  for (int k = 0; k < kWarpChunk; ++k) {
    const int c = min(c0 + k, C - 1);
    r[k] = sample(p.x + ((int64_t)b * C + c) * plane, s, W);
  }
#pragma unroll
  for (int k = 0; k < kWarpChunk; ++k) {
    const int c = c0 + k;
    if (c < C) st_stream(p.out + ((int64_t)b * C + c) * plane + pix, kUpsample ? r[k] * m : r[k]);
  }
}

// ---- several flow maps to one size in ONE launch -------------------------------------------------------------
// The five flow predictions of EEMFlow_cdc are resized to the input size by five calls of upsample2d_flow_as
// (model/EEMFlow/EEMFlow+.py:231-232), each followed by that function's in-place scaling of its input
// (cdc_utils.py:85-86): ten launches of at most 23 MB here.  Same arithmetic per element as bilinear_resize_kernel /
// scale_uv_kernel; blockIdx.z = (map, plane group), blockIdx.y of the scaling kernel = map.
constexpr int kMaxResizeMaps = 8;
struct ResizeMultiParams {
  const float* in[kMaxResizeMaps];
  float* out[kMaxResizeMaps];
  int h[kMaxResizeMaps], w[kMaxResizeMaps];
  float s0[kMaxResizeMaps], s1[kMaxResizeMaps];
  int n, B, C, H, W, align_corners, gz;
  float scale_rest;
};

template <int V>
__global__ void __launch_bounds__(256)
bilinear_resize_multi_kernel(const __grid_constant__ ResizeMultiParams p) {
  const int X = (blockIdx.x * 32 + (threadIdx.x & 31)) * V;
  const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (X >= p.W || Y >= p.H) return;
  const int k = blockIdx.z / p.gz, z = blockIdx.z - k * p.gz;
  const int h = p.h[k], w = p.w[k], C = p.C;
  int xa[V], xb[V], y0, y1;
  float la[V], lb[V], ly0, ly1;
#pragma unroll
  for (int i = 0; i < V; ++i) source_index(X + i, w, p.W, p.align_corners, xa[i], xb[i], la[i], lb[i]);
  source_index(Y, h, p.H, p.align_corners, y0, y1, ly0, ly1);
  const int64_t ip = (int64_t)h * w, op = (int64_t)p.H * p.W;
  const int gz = p.gz, BC = p.B * C;
  const float* r0 = p.in[k] + (int64_t)z * ip + y0 * w;
  const float* r1 = p.in[k] + (int64_t)z * ip + y1 * w;
  float* o = p.out[k] + (int64_t)z * op + (int64_t)Y * p.W + X;
  const int64_t s_step = (int64_t)gz * ip, o_step = (int64_t)gz * op;
  const float scale0 = p.s0[k], scale1 = p.s1[k];
  int c = z % C;
  const int c_step = gz % C;
#pragma unroll 2
  for (int bc = z; bc < BC; bc += gz) {
    const float sc = c == 0 ? scale0 : (c == 1 ? scale1 : p.scale_rest);
    float v[V];
#pragma unroll
    for (int i = 0; i < V; ++i)
      v[i] = vlerp(ly0, hlerp(la[i], __ldg(r0 + xa[i]), lb[i], __ldg(r0 + xb[i])),
                   ly1, hlerp(la[i], __ldg(r1 + xa[i]), lb[i], __ldg(r1 + xb[i])), sc);
    if constexpr (V == 4) {
      st_stream4(o, make_float4(v[0], v[1], v[2], v[3]));
    } else if constexpr (V == 2) {
      asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(o), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
      st_stream(o, v[0]);
    }
    r0 += s_step;
    r1 += s_step;
    o += o_step;
    c += c_step;
    if (c >= C) c -= C;
  }
}

// Row-walking form of the same resize for UPSAMPLING targets (the case of every caller): a thread owns V adjacent output
// columns and walks down a band of output rows.  Consecutive output rows share their source rows (an 80-row map feeds
// 260 rows), so the thread keeps the two horizontally interpolated source rows it is between in registers and fetches
// a new one only when the source row advances: ~2 loads per 3.25 output rows instead of 4 per pixel, and the per-row
// vertical taps come from a small shared-memory table filled once per CTA.  Same rounding sequence as the kernels above.
template <int V>
__global__ void __launch_bounds__(256)
bilinear_resize_multi_rows_kernel(const __grid_constant__ ResizeMultiParams p, int band_rows) {
  extern __shared__ __align__(16) float4 ytab[];                  // per band row: y0, y1 (as int bits), ly0, ly1
  const int Y0 = blockIdx.y * band_rows, nrows = min(band_rows, p.H - Y0);
  const int k = blockIdx.z / p.gz, z = blockIdx.z - k * p.gz;
  const int h = p.h[k], w = p.w[k], C = p.C;
  for (int t = threadIdx.x; t < nrows; t += blockDim.x) {
    int y0, y1;
    float ly0, ly1;
    source_index(Y0 + t, h, p.H, p.align_corners, y0, y1, ly0, ly1);
    ytab[t] = make_float4(__int_as_float(y0), __int_as_float(y1), ly0, ly1);
  }
  __syncthreads();
  const int X = (blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (X >= p.W) return;
  int xa[V], xb[V];
  float la[V], lb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) source_index(X + i, w, p.W, p.align_corners, xa[i], xb[i], la[i], lb[i]);
  const int64_t ip = (int64_t)h * w, op = (int64_t)p.H * p.W;
  const int gz = p.gz, BC = p.B * C;
  const float scale0 = p.s0[k], scale1 = p.s1[k];
  int c = z % C;
  const int c_step = gz % C;
  for (int bc = z; bc < BC; bc += gz) {
    const float sc = c == 0 ? scale0 : (c == 1 ? scale1 : p.scale_rest);
    const float* src = p.in[k] + (int64_t)bc * ip;
    float* o = p.out[k] + (int64_t)bc * op + (int64_t)Y0 * p.W + X;
    auto hrow = [&](int y, float (&r)[V]) {
      const float* row = src + y * w;
#pragma unroll
      for (int i = 0; i < V; ++i) r[i] = hlerp(la[i], __ldg(row + xa[i]), lb[i], __ldg(row + xb[i]));
    };
    int ycur = -2;
    float top[V], bot[V];
    for (int t = 0; t < nrows; ++t) {
      const float4 yt = ytab[t];
      const int y0 = __float_as_int(yt.x), y1 = __float_as_int(yt.y);
      if (y0 != ycur) {                                            // CTA-uniform
        if (y0 == ycur + 1) {
#pragma unroll
          for (int i = 0; i < V; ++i) top[i] = bot[i];             // (ycur's y1 was ycur + 1: it was not the last row)
        } else {
          hrow(y0, top);
        }
        if (y1 != y0) {
          hrow(y1, bot);
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) bot[i] = top[i];
        }
        ycur = y0;
      }
      float v[V];
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = vlerp(yt.z, top[i], yt.w, bot[i], sc);
      if constexpr (V == 4) {
        st_stream4(o, make_float4(v[0], v[1], v[2], v[3]));
      } else if constexpr (V == 2) {
        asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(o), "f"(v[0]), "f"(v[1]) : "memory");
      } else {
        st_stream(o, v[0]);
      }
      o += p.W;
    }
    c += c_step;
    if (c >= C) c -= C;
  }
}

struct ScaleMultiParams {
  float* flow[kMaxResizeMaps];
  int64_t plane[kMaxResizeMaps];
  float s0[kMaxResizeMaps], s1[kMaxResizeMaps];
  int B, C;
};

__global__ void __launch_bounds__(256)
scale_uv_multi_kernel(const __grid_constant__ ScaleMultiParams p) {
  const int k = blockIdx.y;
  const int64_t plane = p.plane[k], total = (int64_t)p.B * 2 * plane;
  float* flow = p.flow[k];
  const float s0 = p.s0[k], s1 = p.s1[k];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / (2 * plane), r = i % (2 * plane);
    const int c = (int)(r / plane);
    float* q = flow + ((int64_t)b * p.C + c) * plane + (r % plane);
    *q = *q * (c == 0 ? s0 : s1);
  }
}

// Thread = V horizontally adjacent output pixels (one V-wide streaming store), looping over a strided set of
// planes so a few thousand fat CTAs cover the tensor instead of one tiny CTA per (tile, plane).
template <int V>
__global__ void __launch_bounds__(256)
replicate_pad_kernel(const float* __restrict__ in, int64_t n_planes, int H, int W, int left, int top,
                     int Ho, int Wo, float* __restrict__ out) {
  const int X = (blockIdx.x * 32 + (threadIdx.x & 31)) * V;
  const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (X >= Wo || Y >= Ho) return;
  const int sy = min(max(Y - top, 0), H - 1);
  int sx[V];
#pragma unroll
  for (int i = 0; i < V; ++i) sx[i] = min(max(X + i - left, 0), W - 1);
  const int64_t ip = (int64_t)H * W, op = (int64_t)Ho * Wo;
  const float* src = in + (int64_t)blockIdx.z * ip + (int64_t)sy * W;
  float* dst = out + (int64_t)blockIdx.z * op + (int64_t)Y * Wo + X;
  const int64_t s_step = (int64_t)gridDim.z * ip, d_step = (int64_t)gridDim.z * op;
#pragma unroll 4
  for (int64_t pl = blockIdx.z; pl < n_planes; pl += gridDim.z) {
    float v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = ld_stream(src + sx[i]);
    if constexpr (V == 4) {
      st_stream4(dst, make_float4(v[0], v[1], v[2], v[3]));
    } else if constexpr (V == 2) {
      asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(dst), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
      st_stream(dst, v[0]);
    }
    src += s_step;
    dst += d_step;
  }
}

// bilinear_sampler (model/model_utils.py:7-15): img [N,C,H,W], coords [N,Ho,Wo,2] in pixels, sampled
// through the reference's normalise / grid_sample(align_corners=True) round trip, zeros outside.
__global__ void __launch_bounds__(256)
bilinear_sample_kernel(const float* __restrict__ img, const float* __restrict__ coords, int N, int C, int H, int W,
                       int Ho, int Wo, float* __restrict__ out, float* __restrict__ mask_out) {
  const int64_t per = (int64_t)Ho * Wo, total = (int64_t)N * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / per, r = i - n * per;
    const float cx = coords[2 * i + 0], cy = coords[2 * i + 1];
    // xgrid = 2*xgrid/(W-1) - 1 (model_utils.py:11-12), then ATen's align_corners=True un-normalisation
    const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, cx), (float)(W - 1)), 1.f);
    const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, cy), (float)(H - 1)), 1.f);
    Bilin s;
    fill_bilin(s, unnormalize(gx, W, EEM_WARP_EXACT), unnormalize(gy, H, EEM_WARP_EXACT), H, W);
    for (int c = 0; c < C; ++c)
      out[((int64_t)n * C + c) * per + r] = sample(img + ((int64_t)n * C + c) * H * W, s, W);
    if (mask_out) mask_out[i] = (gx > -1.f && gy > -1.f && gx < 1.f && gy < 1.f) ? 1.f : 0.f;
  }
}

}  // namespace
}  // namespace eem

using namespace eem;

extern "C" {

int eem_backwarp(const float* x, const float* flow, int B, int C, int H, int W, int convention,
                 int mask_mode, float* out, float* mask_out, eem_stream_t stream_) {
  EEM_CHECK_ARG(x && flow && out, "eem_backwarp: NULL pointer");
  EEM_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "eem_backwarp: sizes must be > 0");
  EEM_CHECK_ARG(convention == EEM_WARP_EXACT || convention == EEM_WARP_HALFPIX, "eem_backwarp: unknown convention %d", convention);
  EEM_CHECK_ARG(mask_mode >= EEM_MASK_NONE && mask_mode <= EEM_MASK_9999, "eem_backwarp: unknown mask_mode %d", mask_mode);
  const int64_t gz = (int64_t)B * ceil_div(C, kWarpChunk);
  EEM_CHECK_ARG(gz <= 65535, "eem_backwarp: B*ceil(C/%d) = %lld exceeds 65535; split the batch", kWarpChunk, (long long)gz);
  dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 8), (unsigned)gz);
  backwarp_kernel<<<grid, 256, 0, as_stream(stream_)>>>(x, flow, B, C, H, W, convention, mask_mode, out, mask_out);
  EEM_CHECK_LAUNCH("backwarp_kernel");
  return EEM_OK;
}

int eem_upsample_flow_warp(const float* coarse_flow, int h, int w, float scale0, float scale1, const float* x, int B, int C,
                           int H, int W, float* flow_out, float* out, eem_stream_t stream_) {
  EEM_CHECK_ARG(coarse_flow && x && flow_out && out, "eem_upsample_flow_warp: NULL pointer");
  EEM_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && h > 0 && w > 0, "eem_upsample_flow_warp: sizes must be > 0");
  const int64_t gz = (int64_t)B * ceil_div(C, kWarpChunk);
  EEM_CHECK_ARG(gz <= 65535, "eem_upsample_flow_warp: B*ceil(C/%d) = %lld exceeds 65535; split the batch", kWarpChunk, (long long)gz);
  FusedWarpParams p{};
  p.x = x; p.out = out; p.flow_out = flow_out; p.coarse = coarse_flow; p.h = h; p.w = w; p.s0 = scale0; p.s1 = scale1;
  p.B = B; p.C = C; p.H = H; p.W = W;
  dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 8), (unsigned)gz);
  fused_flow_warp_kernel<true><<<grid, 256, 0, as_stream(stream_)>>>(p);
  EEM_CHECK_LAUNCH("fused_flow_warp_kernel<upsample>");
  return EEM_OK;
}

int eem_blend_flow_warp(const float* flow_init, const float* inter_flow, const float* m, const float* x, int B, int C, int H,
                        int W, float* flow_out, float* out, eem_stream_t stream_) {
  EEM_CHECK_ARG(flow_init && inter_flow && m && x && flow_out && out, "eem_blend_flow_warp: NULL pointer");
  EEM_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "eem_blend_flow_warp: sizes must be > 0");
  EEM_CHECK_ARG(flow_out != flow_init, "eem_blend_flow_warp: flow_out must not alias flow_init (it is gathered from)");
  const int64_t gz = (int64_t)B * ceil_div(C, kWarpChunk);
  EEM_CHECK_ARG(gz <= 65535, "eem_blend_flow_warp: B*ceil(C/%d) = %lld exceeds 65535; split the batch", kWarpChunk, (long long)gz);
  FusedWarpParams p{};
  p.x = x; p.out = out; p.flow_out = flow_out; p.flow_init = flow_init; p.inter = inter_flow; p.mk = m;
  p.B = B; p.C = C; p.H = H; p.W = W;
  dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 8), (unsigned)gz);
  fused_flow_warp_kernel<false><<<grid, 256, 0, as_stream(stream_)>>>(p);
  EEM_CHECK_LAUNCH("fused_flow_warp_kernel<blend>");
  return EEM_OK;
}

int eem_warp_blend(const float* flow_init, const float* inter_flow, const float* m, int B, int H,
                   int W, float* out, eem_stream_t stream_) {
  EEM_CHECK_ARG(flow_init && inter_flow && m && out, "eem_warp_blend: NULL pointer");
  EEM_CHECK_ARG(B > 0 && H > 0 && W > 0 && B <= 65535, "eem_warp_blend: bad sizes");
  EEM_CHECK_ARG(out != flow_init, "eem_warp_blend: out must not alias flow_init (it is gathered from)");
  dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 8), (unsigned)B);
  warp_blend_kernel<<<grid, 256, 0, as_stream(stream_)>>>(flow_init, inter_flow, m, B, H, W, out);
  EEM_CHECK_LAUNCH("warp_blend_kernel");
  return EEM_OK;
}

int eem_bilinear_resize(const float* in, int B, int C, int h, int w, float* out, int H, int W,
                        int align_corners, float scale0, float scale1, float scale_rest, eem_stream_t stream_) {
  EEM_CHECK_ARG(in && out, "eem_bilinear_resize: NULL pointer");
  EEM_CHECK_ARG(B > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "eem_bilinear_resize: sizes must be > 0");
  const int64_t bc = (int64_t)B * C;
  const int V = (W % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) ? 4
              : (W % 2 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0) ? 2 : 1;
  // enough CTAs to fill the chip (~8 per SM), the rest of the planes are looped inside the thread
  const int64_t blocks_xy = ceil_div(W, 32 * V) * ceil_div(H, 8);
  int64_t gz = ceil_div((int64_t)sm_count() * 8, blocks_xy);
  if (gz < 1) gz = 1;
  if (gz > bc) gz = bc;
  if (gz > 65535) gz = 65535;
  dim3 grid((unsigned)ceil_div(W, 32 * V), (unsigned)ceil_div(H, 8), (unsigned)gz);
  cudaStream_t stream = as_stream(stream_);
  const int ac = align_corners ? 1 : 0;
  if (V == 4) bilinear_resize_kernel<4><<<grid, 256, 0, stream>>>(in, B, C, h, w, out, H, W, ac, scale0, scale1, scale_rest);
  else if (V == 2) bilinear_resize_kernel<2><<<grid, 256, 0, stream>>>(in, B, C, h, w, out, H, W, ac, scale0, scale1, scale_rest);
  else bilinear_resize_kernel<1><<<grid, 256, 0, stream>>>(in, B, C, h, w, out, H, W, ac, scale0, scale1, scale_rest);
  EEM_CHECK_LAUNCH("bilinear_resize_kernel");
  return EEM_OK;
}

int eem_bilinear_resize_multi(const float* const* ins, const int* hs, const int* ws, int n_maps, int B, int C,
                              float* const* outs, int H, int W, int align_corners, const float* scale0, const float* scale1,
                              float scale_rest, eem_stream_t stream_) {
  EEM_CHECK_ARG(ins && hs && ws && outs && scale0 && scale1, "eem_bilinear_resize_multi: NULL pointer");
  EEM_CHECK_ARG(n_maps > 0 && n_maps <= kMaxResizeMaps, "eem_bilinear_resize_multi: n_maps must be in [1,%d]", kMaxResizeMaps);
  EEM_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "eem_bilinear_resize_multi: sizes must be > 0");
  ResizeMultiParams p{};
  int V = 4;
  for (int k = 0; k < n_maps; ++k) {
    EEM_CHECK_ARG(ins[k] && outs[k] && hs[k] > 0 && ws[k] > 0, "eem_bilinear_resize_multi: bad map %d", k);
    p.in[k] = ins[k]; p.out[k] = outs[k]; p.h[k] = hs[k]; p.w[k] = ws[k]; p.s0[k] = scale0[k]; p.s1[k] = scale1[k];
    const int v = (W % 4 == 0 && reinterpret_cast<uintptr_t>(outs[k]) % 16 == 0) ? 4
                : (W % 2 == 0 && reinterpret_cast<uintptr_t>(outs[k]) % 8 == 0) ? 2 : 1;
    if (v < V) V = v;
  }
  const int64_t bc = (int64_t)B * C;
  const int64_t blocks_xy = ceil_div(W, 32 * V) * ceil_div(H, 8);
  int64_t gz = ceil_div((int64_t)sm_count() * 8, blocks_xy * n_maps);
  if (gz < 1) gz = 1;
  if (gz > bc) gz = bc;
  if (gz * n_maps > 65535) gz = 65535 / n_maps;
  p.n = n_maps; p.B = B; p.C = C; p.H = H; p.W = W; p.align_corners = align_corners ? 1 : 0; p.gz = (int)gz; p.scale_rest = scale_rest;
  cudaStream_t stream = as_stream(stream_);
  {
    // upsampling in y for every map (a source row never skips: the row walk fetches each source row once per band):
    // the row-walking kernel; EEM_RESIZE_ROWS=0 keeps the per-pixel kernel (comparisons)
    static const bool rows_on = [] {
      const char* v = getenv("EEM_RESIZE_ROWS");
      return !(v != nullptr && atoi(v) == 0);
    }();
    bool up = rows_on;
    for (int k = 0; k < n_maps; ++k) up = up && hs[k] <= H;
    if (up) {
      constexpr int kBand = 32;
      const int cols = (int)ceil_div(W, V);
      const int threads = cols >= 256 ? 256 : (int)align_up((size_t)cols, 32);
      const int64_t bands = ceil_div(H, kBand), bx = ceil_div(cols, threads);
      int64_t gzr = ceil_div((int64_t)sm_count() * 8, bands * bx * n_maps);
      if (gzr < 1) gzr = 1;
      if (gzr > bc) gzr = bc;
      if (gzr * n_maps > 65535) gzr = 65535 / n_maps;
      p.gz = (int)gzr;
      dim3 grid((unsigned)bx, (unsigned)bands, (unsigned)(gzr * n_maps));
      const size_t smem = kBand * sizeof(float4);
      if (V == 4) bilinear_resize_multi_rows_kernel<4><<<grid, threads, smem, stream>>>(p, kBand);
      else if (V == 2) bilinear_resize_multi_rows_kernel<2><<<grid, threads, smem, stream>>>(p, kBand);
      else bilinear_resize_multi_rows_kernel<1><<<grid, threads, smem, stream>>>(p, kBand);
      EEM_CHECK_LAUNCH("bilinear_resize_multi_rows_kernel");
      return EEM_OK;
    }
  }
  dim3 grid((unsigned)ceil_div(W, 32 * V), (unsigned)ceil_div(H, 8), (unsigned)(gz * n_maps));
  if (V == 4) bilinear_resize_multi_kernel<4><<<grid, 256, 0, stream>>>(p);
  else if (V == 2) bilinear_resize_multi_kernel<2><<<grid, 256, 0, stream>>>(p);
  else bilinear_resize_multi_kernel<1><<<grid, 256, 0, stream>>>(p);
  EEM_CHECK_LAUNCH("bilinear_resize_multi_kernel");
  return EEM_OK;
}

int eem_scale_uv_inplace_multi(float* const* flows, const int* hs, const int* ws, int n_maps, int B, int C,
                               const float* scale0, const float* scale1, eem_stream_t stream_) {
  EEM_CHECK_ARG(flows && hs && ws && scale0 && scale1, "eem_scale_uv_inplace_multi: NULL pointer");
  EEM_CHECK_ARG(n_maps > 0 && n_maps <= kMaxResizeMaps, "eem_scale_uv_inplace_multi: n_maps must be in [1,%d]", kMaxResizeMaps);
  EEM_CHECK_ARG(B > 0 && C >= 2, "eem_scale_uv_inplace_multi: need C >= 2 and a positive batch");
  ScaleMultiParams p{};
  int64_t largest = 0;
  for (int k = 0; k < n_maps; ++k) {
    EEM_CHECK_ARG(flows[k] && hs[k] > 0 && ws[k] > 0, "eem_scale_uv_inplace_multi: bad map %d", k);
    p.flow[k] = flows[k]; p.plane[k] = (int64_t)hs[k] * ws[k]; p.s0[k] = scale0[k]; p.s1[k] = scale1[k];
    if (p.plane[k] > largest) largest = p.plane[k];
  }
  p.B = B; p.C = C;
  int64_t blocks = ceil_div((int64_t)B * 2 * largest, 256);
  if (blocks > 1024) blocks = 1024;
  scale_uv_multi_kernel<<<dim3((unsigned)blocks, (unsigned)n_maps), 256, 0, as_stream(stream_)>>>(p);
  EEM_CHECK_LAUNCH("scale_uv_multi_kernel");
  return EEM_OK;
}

int eem_scale_uv_inplace(float* flow, int B, int C, int h, int w, float scale0, float scale1,
                         eem_stream_t stream_) {
  EEM_CHECK_ARG(flow != nullptr, "eem_scale_uv_inplace: NULL pointer");
  EEM_CHECK_ARG(B > 0 && C >= 2 && h > 0 && w > 0, "eem_scale_uv_inplace: need C >= 2 and positive sizes");
  const int64_t total = (int64_t)B * 2 * h * w;
  int64_t blocks = ceil_div(total, 256);
  if (blocks > 4096) blocks = 4096;
  scale_uv_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream_)>>>(flow, B, C, (int64_t)h * w, scale0, scale1);
  EEM_CHECK_LAUNCH("scale_uv_kernel");
  return EEM_OK;
}

int eem_replicate_pad(const float* in, int B, int C, int H, int W, int left, int right, int top,
                      int bottom, float* out, eem_stream_t stream_) {
  EEM_CHECK_ARG(in && out, "eem_replicate_pad: NULL pointer");
  EEM_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "eem_replicate_pad: sizes must be > 0");
  EEM_CHECK_ARG(left >= 0 && right >= 0 && top >= 0 && bottom >= 0, "eem_replicate_pad: negative padding");
  const int Ho = H + top + bottom, Wo = W + left + right;
  const int64_t planes = (int64_t)B * C;
  const int V = (Wo % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) ? 4
              : (Wo % 2 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0) ? 2 : 1;
  const int64_t blocks_xy = ceil_div(Wo, 32 * V) * ceil_div(Ho, 8);
  int64_t gz = ceil_div((int64_t)sm_count() * 16, blocks_xy);      // ~16 CTAs per SM, planes looped inside
  if (gz < 1) gz = 1;
  if (gz > planes) gz = planes;
  if (gz > 65535) gz = 65535;
  dim3 grid((unsigned)ceil_div(Wo, 32 * V), (unsigned)ceil_div(Ho, 8), (unsigned)gz);
  cudaStream_t stream = as_stream(stream_);
  if (V == 4) replicate_pad_kernel<4><<<grid, 256, 0, stream>>>(in, planes, H, W, left, top, Ho, Wo, out);
  else if (V == 2) replicate_pad_kernel<2><<<grid, 256, 0, stream>>>(in, planes, H, W, left, top, Ho, Wo, out);
  else replicate_pad_kernel<1><<<grid, 256, 0, stream>>>(in, planes, H, W, left, top, Ho, Wo, out);
  EEM_CHECK_LAUNCH("replicate_pad_kernel");
  return EEM_OK;
}

int eem_bilinear_sample(const float* img, const float* coords, int N, int C, int H, int W, int Ho, int Wo,
                        float* out, float* mask_out, eem_stream_t stream_) {
  EEM_CHECK_ARG(img && coords && out, "eem_bilinear_sample: NULL pointer");
  EEM_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "eem_bilinear_sample: sizes must be > 0");
  const int64_t total = (int64_t)N * Ho * Wo;
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (cap > 0 && blocks > cap) blocks = cap;
  bilinear_sample_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream_)>>>(img, coords, N, C, H, W, Ho, Wo, out, mask_out);
  EEM_CHECK_LAUNCH("bilinear_sample_kernel");
  return EEM_OK;
}

}  // extern "C"
