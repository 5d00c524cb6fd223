// Backward (adjoint) kernels of the EEMFlow-side ops, for the training drop-in.
//
//   local correlation  out[b,k,y,x] = s * sum_c f1[b,c,y,x] * f2[b,c,y+dy_k,x+dx_k]
//       d f1[b,c,y,x]   = s * sum_k g[b,k,y,x]         * f2[b,c,y+dy_k,x+dx_k]
//       d f2[b,c,y,x]   = s * sum_k g[b,k,y-dy_k,x-dx_k] * f1[b,c,y-dy_k,x-dx_k]
//     (what the reference's dead extension computed in correlation_backward_input1/2,
//      model/IRRPWC/correlation_package/correlation_cuda_kernel.cu:117-298); both are gathers, so the
//      result is deterministic.
//   backward warp      out[b,c,p] = m(p) * sum_t w_t(p) * x[b,c,tap_t(p)]
//       d x   : scatter m*w_t*g with RED.ADD (like ATen's grid_sampler_2d_backward)
//       d flow: m * sum_c g * d(bilinear)/d(ix,iy) * d(ix,iy)/d(u,v); the 0/1 mask is a constant
//   bilinear resize    adjoint of the interpolation: scatter of the four weights with RED.ADD
#include "common.cuh"

namespace eem {
namespace {

constexpr int MD = 4, ND = 9;

struct LcbParams {
  const float* f1;
  const float* f2;
  const float* g;     // [B, n_out, H, W]
  float* df1;
  float* df2;
  int B, C, H, W, n_out;
  float scale;
  signed char slot[ND * ND];  // gradient plane of displacement channel ch, or -1 when not selected
};

// thread = (pixel, channel); grid.z = B * ceil(C / 8); 8 channels per thread share the g loads.
constexpr int kLcbChunk = 8;

__global__ void __launch_bounds__(256)
local_corr_backward_kernel(const __grid_constant__ LcbParams p) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= p.W || y >= p.H) return;
  const int chunks = (p.C + kLcbChunk - 1) / kLcbChunk;
  const int b = blockIdx.z / chunks, c0 = (blockIdx.z % chunks) * kLcbChunk;
  const int64_t plane = (int64_t)p.H * p.W;
  const float* g = p.g + (int64_t)b * p.n_out * plane;
  const float* f1 = p.f1 + ((int64_t)b * p.C + c0) * plane;
  const float* f2 = p.f2 + ((int64_t)b * p.C + c0) * plane;
  const int nc = min(kLcbChunk, p.C - c0);
  float a1[kLcbChunk], a2[kLcbChunk];
#pragma unroll
  for (int c = 0; c < kLcbChunk; ++c) a1[c] = a2[c] = 0.f;
  for (int ch = 0; ch < ND * ND; ++ch) {
    const int k = p.slot[ch];
    if (k < 0) continue;
    const int dy = ch / ND - MD, dx = ch % ND - MD;
    // d f1: g at this pixel times f2 at the displaced pixel
    const int y2 = y + dy, x2 = x + dx;
    if (y2 >= 0 && y2 < p.H && x2 >= 0 && x2 < p.W) {
      const float gv = __ldg(g + (int64_t)k * plane + (int64_t)y * p.W + x);
      const int64_t o = (int64_t)y2 * p.W + x2;
#pragma unroll
      for (int c = 0; c < kLcbChunk; ++c)
        if (c < nc) a1[c] = fmaf(gv, __ldg(f2 + c * plane + o), a1[c]);
    }
    // d f2: g and f1 at the pixel this one is the displaced partner of
    const int y1 = y - dy, x1 = x - dx;
    if (y1 >= 0 && y1 < p.H && x1 >= 0 && x1 < p.W) {
      const int64_t o = (int64_t)y1 * p.W + x1;
      const float gv = __ldg(g + (int64_t)k * plane + o);
#pragma unroll
      for (int c = 0; c < kLcbChunk; ++c)
        if (c < nc) a2[c] = fmaf(gv, __ldg(f1 + c * plane + o), a2[c]);
    }
  }
  const int64_t o = (int64_t)y * p.W + x;
#pragma unroll
  for (int c = 0; c < kLcbChunk; ++c) {
    if (c < nc) {
      if (p.df1) p.df1[((int64_t)b * p.C + c0 + c) * plane + o] = a1[c] * p.scale;
      if (p.df2) p.df2[((int64_t)b * p.C + c0 + c) * plane + o] = a2[c] * p.scale;
    }
  }
}

// ---- backward warp ------------------------------------------------------------------------------
struct Geo {
  int x0, y0;
  float w, e, n, s;  // fractional x, 1 - w, fractional y, 1 - n
  bool in_nw, in_ne, in_sw, in_se;
  float dix, diy;    // d ix / d u, d iy / d v
};

__device__ __forceinline__ Geo make_geo(float px, float py, int H, int W, int convention) {
  // same forward arithmetic as warp.cu::make_bilin
  const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, px), (float)max(W - 1, 1)), 1.0f);
  const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, py), (float)max(H - 1, 1)), 1.0f);
  float ix, iy;
  Geo q;
  if (convention == EEM_WARP_EXACT) {
    ix = __fmul_rn(__fadd_rn(gx, 1.0f), (float)(W - 1) * 0.5f);
    iy = __fmul_rn(__fadd_rn(gy, 1.0f), (float)(H - 1) * 0.5f);
    q.dix = (float)(W - 1) / (float)max(W - 1, 1);
    q.diy = (float)(H - 1) / (float)max(H - 1, 1);
  } else {
    ix = __fmaf_rn(__fadd_rn(gx, 1.0f), (float)W * 0.5f, -0.5f);
    iy = __fmaf_rn(__fadd_rn(gy, 1.0f), (float)H * 0.5f, -0.5f);
    q.dix = (float)W / (float)max(W - 1, 1);
    q.diy = (float)H / (float)max(H - 1, 1);
  }
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  q.x0 = (int)fminf(fmaxf(fx0, -2.0e9f), 2.0e9f);
  q.y0 = (int)fminf(fmaxf(fy0, -2.0e9f), 2.0e9f);
  q.w = ix - fx0; q.e = 1.0f - q.w;
  q.n = iy - fy0; q.s = 1.0f - q.n;
  const bool xin0 = q.x0 >= 0 && q.x0 < W, xin1 = q.x0 + 1 >= 0 && q.x0 + 1 < W;
  const bool yin0 = q.y0 >= 0 && q.y0 < H, yin1 = q.y0 + 1 >= 0 && q.y0 + 1 < H;
  const bool finite = (fx0 == fx0) && (fy0 == fy0) && fabsf(fx0) < 1.0e9f && fabsf(fy0) < 1.0e9f;
  q.in_nw = finite && xin0 && yin0;
  q.in_ne = finite && xin1 && yin0;
  q.in_sw = finite && xin0 && yin1;
  q.in_se = finite && xin1 && yin1;
  return q;
}

__global__ void __launch_bounds__(256)
backwarp_backward_kernel(const float* __restrict__ x, const float* __restrict__ flow, const float* __restrict__ gout,
                         int B, int C, int H, int W, int convention, int mask_mode, float* __restrict__ dx,
                         float* __restrict__ dflow) {
  const int px = blockIdx.x * 32 + (threadIdx.x & 31);
  const int py = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (px >= W || py >= H) return;
  const int64_t plane = (int64_t)H * W, pix = (int64_t)py * W + px;
  const float u = flow[((int64_t)b * 2 + 0) * plane + pix];
  const float v = flow[((int64_t)b * 2 + 1) * plane + pix];
  const Geo q = make_geo((float)px + u, (float)py + v, H, W, convention);
  const float nw = q.s * q.e, ne = q.s * q.w, sw = q.n * q.e, se = q.n * q.w;
  float m = 1.f;
  if (mask_mode != EEM_MASK_NONE) {
    float ms = q.in_nw ? nw : 0.f;
    ms = __fadd_rn(ms, q.in_ne ? ne : 0.f);
    ms = __fadd_rn(ms, q.in_sw ? sw : 0.f);
    ms = __fadd_rn(ms, q.in_se ? se : 0.f);
    m = (mask_mode == EEM_MASK_GE1) ? (ms >= 1.0f ? 1.f : 0.f) : (ms < 0.9999f ? 0.f : 1.f);
  }
  float gu = 0.f, gv = 0.f;
  const int64_t tap = (int64_t)q.y0 * W + q.x0;
  for (int c = 0; c < C; ++c) {
    const int64_t off = ((int64_t)b * C + c) * plane;
    const float g = gout[off + pix] * m;
    if (dx != nullptr && g != 0.f) {
      if (q.in_nw) red_add_f32(dx + off + tap, g * nw);
      if (q.in_ne) red_add_f32(dx + off + tap + 1, g * ne);
      if (q.in_sw) red_add_f32(dx + off + tap + W, g * sw);
      if (q.in_se) red_add_f32(dx + off + tap + W + 1, g * se);
    }
    if (dflow != nullptr) {
      const float v_nw = q.in_nw ? __ldg(x + off + tap) : 0.f;
      const float v_ne = q.in_ne ? __ldg(x + off + tap + 1) : 0.f;
      const float v_sw = q.in_sw ? __ldg(x + off + tap + W) : 0.f;
      const float v_se = q.in_se ? __ldg(x + off + tap + W + 1) : 0.f;
      gu += g * (q.s * (v_ne - v_nw) + q.n * (v_se - v_sw));
      gv += g * (q.e * (v_sw - v_nw) + q.w * (v_se - v_ne));
    }
  }
  if (dflow != nullptr) {
    dflow[((int64_t)b * 2 + 0) * plane + pix] = gu * q.dix;
    dflow[((int64_t)b * 2 + 1) * plane + pix] = gv * q.diy;
  }
}

// ---- bilinear_sampler backward (model/model_utils.py:7-21 under autograd) -----------------------------------
// out[n,c,r] = bilinear(img[n,c], coords[n,r]) with zeros outside: thread = one sampling location, loop over the
// channels; d img by RED scatter of the 4 taps, d coords analytically (the normalise / un-normalise round trip has
// slope 1).  The optional in-range mask of the forward is piecewise constant and gets no gradient.
__global__ void __launch_bounds__(256)
bilinear_sample_backward_kernel(const float* __restrict__ img, const float* __restrict__ coords, const float* __restrict__ gout,
                                int N, int C, int H, int W, int Ho, int Wo, float* __restrict__ dimg,
                                float* __restrict__ dcoords) {
  const int64_t per = (int64_t)Ho * Wo, total = (int64_t)N * per, plane = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / per, r = i - n * per;
    const Geo q = make_geo(coords[2 * i + 0], coords[2 * i + 1], H, W, EEM_WARP_EXACT);
    const float nw = q.s * q.e, ne = q.s * q.w, sw = q.n * q.e, se = q.n * q.w;
    const int64_t tap = (int64_t)q.y0 * W + q.x0;
    float gx = 0.f, gy = 0.f;
    for (int c = 0; c < C; ++c) {
      const int64_t off = (n * C + c) * plane;
      const float g = gout[(n * C + c) * per + r];
      if (dimg != nullptr && g != 0.f) {
        if (q.in_nw) red_add_f32(dimg + off + tap, g * nw);
        if (q.in_ne) red_add_f32(dimg + off + tap + 1, g * ne);
        if (q.in_sw) red_add_f32(dimg + off + tap + W, g * sw);
        if (q.in_se) red_add_f32(dimg + off + tap + W + 1, g * se);
      }
      if (dcoords != nullptr) {
        const float v_nw = q.in_nw ? __ldg(img + off + tap) : 0.f;
        const float v_ne = q.in_ne ? __ldg(img + off + tap + 1) : 0.f;
        const float v_sw = q.in_sw ? __ldg(img + off + tap + W) : 0.f;
        const float v_se = q.in_se ? __ldg(img + off + tap + W + 1) : 0.f;
        gx += g * (q.s * (v_ne - v_nw) + q.n * (v_se - v_sw));
        gy += g * (q.e * (v_sw - v_nw) + q.w * (v_se - v_ne));
      }
    }
    if (dcoords != nullptr) {
      dcoords[2 * i + 0] = gx * q.dix;
      dcoords[2 * i + 1] = gy * q.diy;
    }
  }
}

// ---- bilinear resize backward ---------------------------------------------------------------------
// ATen's source index with the scale hoisted (same value: it is formed by the same fp32 division).
__device__ __forceinline__ void source_index_s(int dst, int in_size, float scale, int align_corners, int& i0, int& i1,
                                               float& l0, float& l1) {
  float src;
  if (align_corners) {
    src = scale * (float)dst;
  } else {
    src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// WARP = one input cell of one plane (gather form): its lanes stride over the output pixels whose interpolation
// footprint contains the cell and combine with a fixed shuffle tree, so there are no atomics and the result is
// deterministic.  (One thread per cell serialised the 100 x 100-pixel footprints of the model's final 4x5 ->
// full-resolution upsamples: 2.9 ms of a 14 ms training step.)
__global__ void __launch_bounds__(256)
bilinear_resize_backward_kernel(const float* __restrict__ gout, int B, int C, int h, int w, int H, int W,
                                int align_corners, float scale0, float scale1, float scale_rest,
                                float* __restrict__ gin) {
  const int lane = threadIdx.x & 31;
  const int64_t cells = (int64_t)B * C * h * w;
  const float sy = align_corners ? (H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f) : (float)h / (float)H;
  const float sx = align_corners ? (W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f) : (float)w / (float)W;
  const float off = align_corners ? 0.f : 0.5f;
  const int64_t op = (int64_t)H * W;
  for (int64_t cell = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); cell < cells; cell += (int64_t)gridDim.x * 8) {
    const int x = (int)(cell % w);
    const int y = (int)((cell / w) % h);
    const int64_t bc = cell / ((int64_t)w * h);
    // conservative output ranges that can touch input row y / column x
    int Y0 = 0, Y1 = H - 1, X0 = 0, X1 = W - 1;
    if (sy > 0.f) {
      Y0 = max(0, (int)floorf(((float)(y - 1) + off) / sy - off) - 1);
      Y1 = min(H - 1, (int)ceilf(((float)(y + 1) + off) / sy - off) + 1);
    }
    if (sx > 0.f) {
      X0 = max(0, (int)floorf(((float)(x - 1) + off) / sx - off) - 1);
      X1 = min(W - 1, (int)ceilf(((float)(x + 1) + off) / sx - off) + 1);
    }
    const float* g = gout + bc * op;
    float acc = 0.f;
    for (int Y = Y0; Y <= Y1; ++Y) {
      int y0, y1;
      float ly0, ly1;
      source_index_s(Y, h, sy, align_corners, y0, y1, ly0, ly1);
      const float wy = (y0 == y ? ly0 : 0.f) + (y1 == y ? ly1 : 0.f);
      if (wy == 0.f) continue;                               // warp-uniform
      for (int X = X0 + lane; X <= X1; X += 32) {
        int x0, x1;
        float lx0, lx1;
        source_index_s(X, w, sx, align_corners, x0, x1, lx0, lx1);
        const float wx = (x0 == x ? lx0 : 0.f) + (x1 == x ? lx1 : 0.f);
        if (wx != 0.f) acc = fmaf(wy * wx, __ldg(g + (int64_t)Y * W + X), acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const int c = (int)(bc % C);
      const float sc = c == 0 ? scale0 : (c == 1 ? scale1 : scale_rest);
      gin[cell] = acc * sc;
    }
  }
}


// ---- batched fp32 GEMM for the backward of the all-pairs pyramid ------------------------------------------------
// level_l[b,i,j] = s * sum_d f1[b,d,i] * pool^l(f2)[b,d,j]  (model/corr.py:13-27, 52-60) gives
//   d f1[b,d,i]          = s * sum_l sum_j pool^l(f2)[b,d,j] * dV_l[b,i,j]     C = A . B^T   (A = f2_l [D,P_l], B = dV_l [P,P_l])
//   d pool^l(f2)[b,d,j]  = s * sum_i f1[b,d,i]           * dV_l[b,i,j]     C = A . B     (A = f1 [D,P],    B = dV_l [P,P_l])
// (the reference gets these from autograd through torch.matmul; round 1 used torch.bmm -> cuBLAS).  Exact fp32 FFMA,
// 64 x 64 x 16 shared-memory tiles, 4 x 4 outputs per thread, same inner loop as the forward's fp32 path.
//   C[b] (M x N, ldc) = alpha * A[b] (M x K, row-major, lda) * op(B[b]) (+ C[b] when accumulate)
//   op(B) = B (K x N, row-major, ldb) or B^T with B given N x K row-major (b_transposed)
constexpr int GT = 64, GK = 16;

template <bool kBT>
__global__ void __launch_bounds__(256)
batched_gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
                        int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB, int64_t strideC, float alpha,
                        int accumulate) {
  __shared__ __align__(16) float As[GK][GT + 4];
  __shared__ __align__(16) float Bs[GK][GT + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  A += (int64_t)b * strideA;
  B += (int64_t)b * strideB;
  C += (int64_t)b * strideC;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += GK) {
    __syncthreads();
    // A tile: rows m0..m0+63, columns k0..k0+15 (k contiguous in memory) -> As[k][m]
    for (int t = threadIdx.x; t < GT * GK; t += 256) {
      const int m = t / GK, k = t % GK;
      As[k][m] = (m0 + m < M && k0 + k < K) ? __ldg(A + (int64_t)(m0 + m) * lda + k0 + k) : 0.f;
    }
    if (kBT) {      // B is N x K: rows n0..n0+63, columns k0.. (k contiguous) -> Bs[k][n]
      for (int t = threadIdx.x; t < GT * GK; t += 256) {
        const int n = t / GK, k = t % GK;
        Bs[k][n] = (n0 + n < N && k0 + k < K) ? __ldg(B + (int64_t)(n0 + n) * ldb + k0 + k) : 0.f;
      }
    } else {        // B is K x N: rows k0.., columns n0..n0+63 (n contiguous) -> Bs[k][n]
      for (int t = threadIdx.x; t < GT * GK; t += 256) {
        const int k = t / GT, n = t % GT;
        Bs[k][n] = (n0 + n < N && k0 + k < K) ? __ldg(B + (int64_t)(k0 + k) * ldb + n0 + n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = m0 + ty * 4 + r;
    if (m >= M) continue;
    float* o = C + (int64_t)m * ldc + n0 + tx * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (n0 + tx * 4 + c < N) o[c] = accumulate ? fmaf(alpha, acc[r][c], o[c]) : alpha * acc[r][c];
  }
}

}  // namespace
}  // namespace eem

using namespace eem;

extern "C" {

int eem_local_corr_backward(const float* f1, const float* f2, const float* grad_out, int B, int C, int H, int W,
                            int max_disp, const int* index, int n_out, float scale, float* grad_f1, float* grad_f2,
                            eem_stream_t stream_) {
  EEM_CHECK_ARG(f1 && f2 && grad_out, "eem_local_corr_backward: NULL pointer");
  EEM_CHECK_ARG(grad_f1 || grad_f2, "eem_local_corr_backward: at least one gradient output is required");
  EEM_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "eem_local_corr_backward: sizes must be > 0");
  if (max_disp != MD)
    return fail(EEM_ERR_UNSUPPORTED, "eem_local_corr_backward: only max_disp == %d is implemented, got %d", MD, max_disp);
  LcbParams p{};
  p.f1 = f1; p.f2 = f2; p.g = grad_out; p.df1 = grad_f1; p.df2 = grad_f2;
  p.B = B; p.C = C; p.H = H; p.W = W; p.n_out = n_out; p.scale = scale;
  if (index == nullptr) {
    EEM_CHECK_ARG(n_out == ND * ND, "eem_local_corr_backward: n_out must be %d without an index list", ND * ND);
    for (int ch = 0; ch < ND * ND; ++ch) p.slot[ch] = (signed char)ch;
  } else {
    EEM_CHECK_ARG(n_out > 0 && n_out <= ND * ND, "eem_local_corr_backward: n_out must be in [1,%d]", ND * ND);
    for (int ch = 0; ch < ND * ND; ++ch) p.slot[ch] = -1;
    for (int k = 0; k < n_out; ++k) {
      EEM_CHECK_ARG(index[k] >= 0 && index[k] < ND * ND, "eem_local_corr_backward: index[%d]=%d out of range", k, index[k]);
      if (p.slot[index[k]] != -1)
        return fail(EEM_ERR_UNSUPPORTED, "eem_local_corr_backward: repeated channel %d in index list", index[k]);
      p.slot[index[k]] = (signed char)k;
    }
  }
  const int64_t gz = (int64_t)B * ceil_div(C, kLcbChunk);
  EEM_CHECK_ARG(gz <= 65535, "eem_local_corr_backward: B*ceil(C/%d) exceeds 65535; split the batch", kLcbChunk);
  dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 8), (unsigned)gz);
  local_corr_backward_kernel<<<grid, 256, 0, as_stream(stream_)>>>(p);
  EEM_CHECK_LAUNCH("local_corr_backward_kernel");
  return EEM_OK;
}

int eem_backwarp_backward(const float* x, const float* flow, const float* grad_out, int B, int C, int H, int W,
                          int convention, int mask_mode, float* grad_x, float* grad_flow, eem_stream_t stream_) {
  EEM_CHECK_ARG(x && flow && grad_out, "eem_backwarp_backward: NULL pointer");
  EEM_CHECK_ARG(grad_x || grad_flow, "eem_backwarp_backward: at least one gradient output is required");
  EEM_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "eem_backwarp_backward: bad sizes");
  EEM_CHECK_ARG(convention == EEM_WARP_EXACT || convention == EEM_WARP_HALFPIX, "eem_backwarp_backward: unknown convention %d", convention);
  EEM_CHECK_ARG(mask_mode >= EEM_MASK_NONE && mask_mode <= EEM_MASK_9999, "eem_backwarp_backward: unknown mask_mode %d", mask_mode);
  cudaStream_t stream = as_stream(stream_);
  if (grad_x) EEM_CHECK_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)B * C * H * W * sizeof(float), stream));
  dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 8), (unsigned)B);
  backwarp_backward_kernel<<<grid, 256, 0, stream>>>(x, flow, grad_out, B, C, H, W, convention, mask_mode, grad_x, grad_flow);
  EEM_CHECK_LAUNCH("backwarp_backward_kernel");
  return EEM_OK;
}

int eem_bilinear_sample_backward(const float* img, const float* coords, const float* grad_out, int N, int C, int H, int W,
                                 int Ho, int Wo, float* grad_img, float* grad_coords, eem_stream_t stream_) {
  EEM_CHECK_ARG(img && coords && grad_out, "eem_bilinear_sample_backward: NULL pointer");
  EEM_CHECK_ARG(grad_img || grad_coords, "eem_bilinear_sample_backward: at least one gradient output is required");
  EEM_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "eem_bilinear_sample_backward: sizes must be > 0");
  cudaStream_t stream = as_stream(stream_);
  if (grad_img) EEM_CHECK_CUDA(cudaMemsetAsync(grad_img, 0, (size_t)N * C * H * W * sizeof(float), stream));
  const int64_t total = (int64_t)N * Ho * Wo;
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (cap > 0 && blocks > cap) blocks = cap;
  bilinear_sample_backward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(img, coords, grad_out, N, C, H, W, Ho, Wo, grad_img, grad_coords);
  EEM_CHECK_LAUNCH("bilinear_sample_backward_kernel");
  return EEM_OK;
}

int eem_bilinear_resize_backward(const float* grad_out, int B, int C, int h, int w, int H, int W, int align_corners,
                                 float scale0, float scale1, float scale_rest, float* grad_in, eem_stream_t stream_) {
  EEM_CHECK_ARG(grad_out && grad_in, "eem_bilinear_resize_backward: NULL pointer");
  EEM_CHECK_ARG(B > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "eem_bilinear_resize_backward: sizes must be > 0");
  const int64_t cells = (int64_t)B * C * h * w;
  int64_t blocks = ceil_div(cells, 8);
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  bilinear_resize_backward_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream_)>>>(grad_out, B, C, h, w, H, W, align_corners ? 1 : 0,
                                                                                    scale0, scale1, scale_rest, grad_in);
  EEM_CHECK_LAUNCH("bilinear_resize_backward_kernel");
  return EEM_OK;
}

int eem_batched_gemm_f32(const float* A, const float* B, float* C, int batch, int M, int N, int K, int64_t lda, int64_t ldb,
                         int64_t ldc, int64_t strideA, int64_t strideB, int64_t strideC, int b_transposed, float alpha,
                         int accumulate, eem_stream_t stream_) {
  EEM_CHECK_ARG(A && B && C, "eem_batched_gemm_f32: NULL pointer");
  EEM_CHECK_ARG(batch > 0 && M > 0 && N > 0 && K > 0, "eem_batched_gemm_f32: sizes must be > 0");
  EEM_CHECK_ARG(batch <= 65535 && (M + GT - 1) / GT <= 65535, "eem_batched_gemm_f32: batch or M too large for one launch");
  dim3 grid((unsigned)((N + GT - 1) / GT), (unsigned)((M + GT - 1) / GT), (unsigned)batch);
  if (b_transposed)
    batched_gemm_f32_kernel<true><<<grid, 256, 0, as_stream(stream_)>>>(A, B, C, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, alpha, accumulate);
  else
    batched_gemm_f32_kernel<false><<<grid, 256, 0, as_stream(stream_)>>>(A, B, C, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, alpha, accumulate);
  EEM_CHECK_LAUNCH("batched_gemm_f32_kernel");
  return EEM_OK;
}

}  // extern "C"
