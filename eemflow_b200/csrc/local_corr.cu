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
};

__global__ void __launch_bounds__(kThreads)
local_corr_kernel(const __grid_constant__ LocalCorrParams p) {
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
    for (int t = threadIdx.x; t < CC * TH * TW; t += kThreads) {
      const int c = t / (TH * TW), r = (t / TW) % TH, x = t % TW;
      const int gy = y0 + r, gx = x0 + x;
      float v = 0.f;
      if (c < nc && gy < p.H && gx < p.W) v = __ldg(f1 + (int64_t)(c0 + c) * plane + (int64_t)gy * p.W + gx);
      s1[c][r][x] = v;
    }
    for (int t = threadIdx.x; t < CC * F2H * F2W; t += kThreads) {
      const int c = t / (F2H * F2W), r = (t / F2W) % F2H, x = t % F2W;
      const int gy = y0 + r - MD, gx = x0 + x - MD;
      float v = 0.f;
      if (c < nc && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
        v = __ldg(f2 + (int64_t)(c0 + c) * plane + (int64_t)gy * p.W + gx);
      s2[c][r][x] = v;
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
  dim3 grid((unsigned)ceil_div(W, TW), (unsigned)ceil_div(H, TH), (unsigned)B);
  local_corr_kernel<<<grid, kThreads, 0, as_stream(stream_)>>>(p);
  EEM_CHECK_LAUNCH("local_corr_kernel");
  return EEM_OK;
}
