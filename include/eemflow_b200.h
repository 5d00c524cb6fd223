/*
 * eemflow_b200.h -- C ABI of libeemflow_b200.so
 *
 * B200 (sm_100a) kernels for the data-parallel hot path of boomluo02/EEMFlow:
 * event voxelization -> correlation (all-pairs pyramid + local 9x9) -> window
 * lookup -> backward warp / flow upsampling.  Every entry point takes raw DEVICE
 * pointers, plain sizes and a CUDA stream; the caller owns all memory (inputs,
 * outputs and workspaces).  Nothing here allocates, frees or retains memory,
 * and nothing synchronises the stream.  All functions return 0 on success and
 * a negative eem_status otherwise; eem_last_error_string() describes the last
 * failure on the calling thread.
 *
 * The "replaces" notes cite the reference (paths relative to the EEMFlow tree)
 * so a maintainer can see which Python/ATen code each call stands in for.
 * INTEGRATION.md shows the ctypes binding used by the Python drop-in classes.
 */
#ifndef EEMFLOW_B200_H_
#define EEMFLOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EEM_API __attribute__((visibility("default")))
#else
#define EEM_API
#endif

/* cudaStream_t / CUstream, passed as an opaque pointer (NULL = legacy default stream). */
typedef void* eem_stream_t;

enum eem_status {
  EEM_OK = 0,
  EEM_ERR_BAD_ARG = -1,       /* non-positive size, NULL pointer, unsupported flag       */
  EEM_ERR_MISALIGNED = -2,    /* pointer not aligned as documented                       */
  EEM_ERR_WORKSPACE = -3,     /* workspace NULL or smaller than *_workspace_bytes()      */
  EEM_ERR_CUDA = -4,          /* CUDA runtime/driver error (message holds the CUDA text) */
  EEM_ERR_UNSUPPORTED = -5    /* shape outside what the kernel family implements         */
};

EEM_API int eem_version(void);
EEM_API const char* eem_last_error_string(void);
/* Number of SMs of the current device (used by callers to size benches); <0 on error. */
EEM_API int eem_sm_count(void);
/* Number of CUDA kernels this library has launched in this process so far (bench.py reports the
 * difference over its timed region as "gpu_launches"). */
EEM_API long long eem_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * K1  event voxelization (time-bilinear polarity voting)
 * replaces: utils/transformers.py:56-112 (EventSequenceToVoxelGrid_Pytorch.__call__, voting
 *           part; identical copies utils_luo/event_utils.py:185-241, loader/loader_utils.py:469-525)
 *
 * events : [n_total, 4] float64 rows (ts, x, y, p) exactly as EventSequence.features holds them,
 *          32-byte aligned.  Windows are concatenated; window w owns rows
 *          [offsets[w], offsets[w+1]).  offsets is a DEVICE array of n_windows+1 int64.
 * grid   : [n_windows, num_bins, height, width] float32.  The call zero-fills it first.
 * mode   : EEM_VOXEL_ATOMIC        fp32 red.global.add, order of adds not defined
 *          EEM_VOXEL_DETERMINISTIC stable sort by voxel, sequential fp32 adds in the reference's
 *                                  order (all "left" votes in event order, then all "right"
 *                                  votes): bit-exact against the CPU reference.
 * max_events_per_window : host-side upper bound of offsets[w+1]-offsets[w] (sizes the launch).
 * normalize : non-zero fuses K2 (mean / unbiased std over the non-zero voxels, utils/transformers.py:114-122)
 *           into the call; stats_out (optional DEVICE [n_windows,3] float64: count, mean, std) as for K2.
 * Atomic mode picks between two layouts by itself: direct RED.ADD.F32 into the grid, or -- when a call
 * has at least two events per voxel -- a scratch "pair" layout that needs ONE 8-byte vector RED
 * per event followed by a streaming combine pass (see csrc/voxelize.cu); both need the workspace size
 * returned by eem_voxelize_workspace_bytes for the same arguments.
 * dropped : optional DEVICE int64 counter (may be NULL); incremented once per vote whose flat
 *           index x + y*W + bin*W*H falls outside the grid (the reference raises IndexError
 *           there).  In-range flat indices are voted exactly like the reference, including its
 *           row wrap-around for x >= width.
 * ------------------------------------------------------------------------------------------ */
enum { EEM_VOXEL_ATOMIC = 0, EEM_VOXEL_DETERMINISTIC = 1 };

EEM_API size_t eem_voxelize_workspace_bytes(int64_t n_total, int n_windows, int num_bins,
                                            int height, int width, int mode, int normalize);
EEM_API int eem_voxelize(const double* events, const int64_t* offsets, int n_windows,
                         int64_t n_total, int64_t max_events_per_window, int num_bins,
                         int height, int width, int mode, int normalize, float* grid,
                         int64_t* dropped, double* stats_out, void* workspace,
                         size_t workspace_bytes, eem_stream_t stream);

/* K1 with EventSequence's `timestamp_multiplier` (loader/loader_utils.py:367-368) applied on the fly: stamps are
 * multiplied by `timestamp_multiplier` in float64 as they are loaded (the relative conversion of :393-397 is implied:
 * the voting subtracts the window's first stamp anyway, with the same roundings).  Together with windows that are merely
 * CONSECUTIVE row ranges this voxelizes the dt4 concatenation of loader/MVSEC.py:245-262 (four frames' events joined and
 * scaled by 1e6) without a host-side concatenate / scale pass.  timestamp_multiplier == 1 is eem_voxelize. */
EEM_API int eem_voxelize_scaled(const double* events, double timestamp_multiplier, const int64_t* offsets,
                                int n_windows, int64_t n_total, int64_t max_events_per_window, int num_bins,
                                int height, int width, int mode, int normalize, float* grid,
                                int64_t* dropped, double* stats_out, void* workspace,
                                size_t workspace_bytes, eem_stream_t stream);

/* K1 (packed columns)  Same voting, events given as separate columns -- 13 B/event instead of 32:
 *   t : float64 [N] with the values of EventSequence.features[:,0] (t_is_ns = 0), or the raw int64
 *       nanosecond stamps of an HREM .npz (t_is_ns = 1); the kernel then applies the reference's
 *       float64 chain t*1e-9 (loader/loader_utils.py:34), *1e6 (EventSequence timestamp_multiplier,
 *       :367-368) and "- first stamp" (:393-397) itself, bit for bit.
 *   x, y : int16 [N];  p : int8 [N] (0/1 or -1/+1; 0 votes as -1, utils/transformers.py:82).
 * Everything else as eem_voxelize (same workspace size). */
EEM_API int eem_voxelize_soa(const void* t, int t_is_ns, const int16_t* x, const int16_t* y,
                             const int8_t* p, const int64_t* offsets, int n_windows,
                             int64_t n_total, int64_t max_events_per_window, int num_bins,
                             int height, int width, int mode, int normalize, float* grid,
                             int64_t* dropped, double* stats_out, void* workspace,
                             size_t workspace_bytes, eem_stream_t stream);

/* K2  voxel-grid normalisation: per window, mean / unbiased std over the NON-ZERO voxels, then
 *     v = (v - mean) / std on the non-zero voxels (v - mean when !(std > 0), which includes the
 *     NaN std of a single non-zero voxel).  In place.
 * replaces: utils/transformers.py:114-122
 * stats_out: optional DEVICE [n_windows, 3] float64 (count, mean, std) for inspection; may be NULL. */
EEM_API size_t eem_voxel_normalize_workspace_bytes(int n_windows, int64_t voxels_per_window);
EEM_API int eem_voxel_normalize(float* grid, int n_windows, int64_t voxels_per_window,
                                double* stats_out, void* workspace, size_t workspace_bytes,
                                eem_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3  all-pairs correlation pyramid
 * replaces: model/corr.py:13-27 (CorrBlock.__init__) and :52-60 (CorrBlock.corr)
 *
 * fmap1, fmap2 : [B, D, H, W] float32 contiguous (NCHW), 16-byte aligned.
 * levels[l]    : DEVICE pointer to level l, [B*H*W, H_l*W_l] float32 with H_0=H, W_0=W and
 *                H_{l+1} = H_l/2, W_{l+1} = W_l/2 (floor), l < num_levels (<= 8).  `levels` itself
 *                is a HOST array of num_levels device pointers.  A level with H_l*W_l == 0 is skipped.
 *     level_l[b*P + i, j] = (1/sqrt(D)) * sum_d fmap1[b,d,i] * pool^l(fmap2)[b,d,j]
 *   which equals avg_pool2d(2,2) applied l times to level 0 (pooling is linear and acts on j only).
 * precision    : EEM_CORR_FP32  CUDA-core fp32 FMA (<= 1e-5 abs against the reference)
 *                EEM_CORR_TF32  tcgen05.mma kind::tf32, TMA-fed, TMEM accumulators
 * ------------------------------------------------------------------------------------------ */
enum { EEM_CORR_FP32 = 0, EEM_CORR_TF32 = 1 };

EEM_API size_t eem_corr_pyramid_workspace_bytes(int B, int D, int H, int W, int num_levels);
EEM_API int eem_corr_pyramid(const float* fmap1, const float* fmap2, int B, int D, int H, int W,
                             int num_levels, float* const* levels, int precision,
                             void* workspace, size_t workspace_bytes, eem_stream_t stream);

/* K4  avg_pool2d(kernel 2, stride 2, floor) over planes: in [n_planes, h, w] -> out [n_planes, h/2, w/2].
 * replaces: model/corr.py:25-27 when a caller pools an existing volume itself. */
EEM_API int eem_avg_pool2x2(const float* in, int64_t n_planes, int h, int w, float* out,
                            eem_stream_t stream);

/* K5  multi-level (2r+1)^2 bilinear window lookup
 * replaces: model/corr.py:29-50 (CorrBlock.__call__) + model/model_utils.py:7-15 (bilinear_sampler)
 *
 * levels : HOST array of num_levels DEVICE pointers laid out as K3 writes them.
 * coords : [B, 2, H, W] float32 (channel 0 = x, channel 1 = y), pixel units of level 0.
 * out    : [B, num_levels*(2r+1)^2, H, W] float32 contiguous;
 *          channel l*(2r+1)^2 + a*(2r+1) + b samples level l at (x/2^l + a - r, y/2^l + b - r)
 *          (the reference's transposed window: a moves x, b moves y), zeros outside the map. */
EEM_API int eem_corr_lookup(const float* const* levels, int B, int H, int W, int num_levels,
                            int radius, const float* coords, float* out, eem_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3p / K5p  fp16 WORKING pyramid: the same pyramid as K3 (TF32 contraction, fp32 accumulation), rounded once to
 * fp16 and laid out for the window lookup instead of for inspection -- half the bytes written by the GEMM and read
 * by each of the ~12 lookups per pair that follow (model/eraft.py:140-142).
 * replaces: model/corr.py:13-27 + :52-60 (build) and :29-50 (lookup) when the caller only consumes the lookups.
 *
 * packed : [B*H*W, row_elems] fp16 (uint16 storage), 32-byte aligned.  Row (b*P + i) holds ALL levels of source
 *          position i back to back; level l starts at element level_offset[l] and stores its H_l x W_l map as
 *          4x4-pixel tiles (tile-row-major; row-major inside a tile; 16 fp16 = one 32-byte sector per tile):
 *            element(y, x) = level_offset[l] + (((y/4) * ceil(W_l/4) + x/4) * 16 + (y%4)*4 + x%4)
 *          cells of the last tile row/column beyond the map are zeros; each level is padded to a multiple of
 *          32 elements (level_elems[l]).  eem_corr_packed_layout returns these numbers (HOST arrays of num_levels).
 * eem_corr_pyramid_packed needs H*W % 4 == 0, D % 32 == 0, D <= 256 (EEM_ERR_UNSUPPORTED otherwise).
 * eem_corr_lookup_packed has the output contract of eem_corr_lookup (K5).
 * eem_corr_pyramid_unpack expands `packed` into the f32 level tensors of K3 (levels: HOST array of DEVICE pointers).
 * ------------------------------------------------------------------------------------------ */
EEM_API int eem_corr_packed_layout(int H, int W, int num_levels, int64_t* level_offset,
                                   int64_t* level_elems, int64_t* row_elems);
EEM_API size_t eem_corr_pyramid_packed_workspace_bytes(int B, int D, int H, int W, int num_levels);
EEM_API int eem_corr_pyramid_packed(const float* fmap1, const float* fmap2, int B, int D, int H, int W,
                                    int num_levels, void* packed, void* workspace,
                                    size_t workspace_bytes, eem_stream_t stream);
EEM_API int eem_corr_lookup_packed(const void* packed, int B, int H, int W, int num_levels, int radius,
                                   const float* coords, float* out, eem_stream_t stream);
EEM_API int eem_corr_pyramid_unpack(const void* packed, int B, int H, int W, int num_levels,
                                    float* const* levels, eem_stream_t stream);

/* K5b generic pixel-coordinate bilinear sampler
 * replaces: bilinear_sampler (model/model_utils.py:7-21) = normalise with (S-1) + F.grid_sample(align_corners=True)
 * img [N,C,H,W], coords [N,Ho,Wo,2] (x, y) in pixels -> out [N,C,Ho,Wo]; mask_out (optional, [N,Ho,Wo,1])
 * is the reference's strict-interior mask (model_utils.py:17-19). */
EEM_API int eem_bilinear_sample(const float* img, const float* coords, int N, int C, int H, int W,
                                int Ho, int Wo, float* out, float* mask_out, eem_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K6  local (2*max_disp+1)^2 correlation with fused 1/C and fused channel selection
 * replaces: spatial_correlation_sampler.SpatialCorrelationSampler(1, 9, 1, 0, 1) as called from
 *           model/EEMFlow/EEMFlow.py:14-23 / EEMFlow+.py:16-25, the "/ c" there, and the
 *           torch.index_select that follows (EEMFlow.py:160, EEMFlow+.py:178,190,202,214,226)
 *
 * f1, f2 : [B, C, H, W] float32.
 * out    : [B, n_out, H, W];  out[b,k,y,x] = scale * sum_c f1[b,c,y,x] * f2[b,c,y+dy,x+dx]
 *          with channel id ch = (dy+md)*(2md+1) + (dx+md), zero outside f2;
 *          index == NULL: n_out must be (2md+1)^2 and k = ch; else ch = index[k] (HOST array).
 * scale  : 1.0f for the raw sampler output, 1.0f/C for Correlation.forward. */
EEM_API int eem_local_corr(const float* f1, const float* f2, int B, int C, int H, int W,
                           int max_disp, const int* index, int n_out, float scale, float* out,
                           eem_stream_t stream);

/* K6t  the same operator with the channel contraction on the tensor cores (tcgen05 kind::tf32, fp32 accumulate):
 * a banded GEMM per 4-row x 32-pixel block of f1 against the 12 x 40 window of f2 around it, operands read by TMA
 * straight from the NCHW maps (csrc/local_corr_tc.cu).  Opt-in: products are rounded to TF32 (the sampler the
 * reference calls computes them in fp32), results agree with eem_local_corr to ~1e-3 relative.
 * Same arguments and errors as eem_local_corr; additionally EEM_ERR_UNSUPPORTED unless W % 4 == 0 and f1 / f2 are
 * 16-byte aligned (eem_local_corr_tf32_supported() answers that without a device). */
EEM_API int eem_local_corr_tf32(const float* f1, const float* f2, int B, int C, int H, int W,
                                int max_disp, const int* index, int n_out, float scale, float* out,
                                eem_stream_t stream);
EEM_API int eem_local_corr_tf32_supported(int B, int C, int H, int W, int max_disp);

/* ------------------------------------------------------------------------------------------
 * K7  backward warp (bilinear, zero padding)
 * replaces: EEMFlow_cdc.warp (model/EEMFlow/EEMFlow+.py:137-149)          -> EEM_WARP_EXACT
 *           tensor_tools.torch_warp (utils_luo/tools.py:2262-2306)         -> EEM_WARP_HALFPIX
 *           WarpingLayer_no_div.forward (model/EEMFlow/cdc_utils.py:50-78) -> EEM_WARP_HALFPIX + EEM_MASK_GE1
 *           tensor_tools.torch_warp_mask (utils_luo/tools.py:2217-2259)    -> EEM_WARP_HALFPIX + EEM_MASK_9999
 *
 * Both conventions normalise with (W-1): g = 2*(px+u)/max(W-1,1) - 1.  EXACT un-normalises with
 * align_corners=True ((g+1)/2*(W-1)); HALFPIX with align_corners=False (((g+1)*W-1)/2), i.e.
 * the reference's effective sample position px'*W/(W-1) - 0.5.
 * x    : [B, C, H, W], flow : [B, 2, H, W] (u, v), out : [B, C, H, W].
 * mask_out : optional [B, 1, H, W] receiving the 0/1 validity mask (may be NULL). */
enum { EEM_WARP_EXACT = 0, EEM_WARP_HALFPIX = 1 };
enum { EEM_MASK_NONE = 0, EEM_MASK_GE1 = 1, EEM_MASK_9999 = 2 };

EEM_API int eem_backwarp(const float* x, const float* flow, int B, int C, int H, int W,
                         int convention, int mask_mode, float* out, float* mask_out,
                         eem_stream_t stream);

/* K7b fused CDC blend: out = warp_HALFPIX(flow_init, inter_flow) * (1 - m) + flow_init * m
 * replaces: model/EEMFlow/cdc_utils.py:173.  flow_init, inter_flow, out: [B,2,H,W]; m: [B,1,H,W]. */
EEM_API int eem_warp_blend(const float* flow_init, const float* inter_flow, const float* m,
                           int B, int H, int W, float* out, eem_stream_t stream);

/* K7c fused per-level chains of EEMFlow_cdc (same arithmetic as the separate calls, one launch and one flow round trip less):
 *   eem_upsample_flow_warp : flow_out = upsample2d_flow_as(coarse_flow, [H,W], if_rate) WITHOUT the in-place side effect
 *                            (scale0 = W/w, scale1 = H/h), out = WarpingLayer_no_div(x, flow_out)
 *                            replaces: model/EEMFlow/cdc_utils.py:156-160 (K8 followed by K7 HALFPIX + MASK_GE1)
 *   eem_blend_flow_warp    : flow_out = K7b blend(flow_init, inter_flow, m), out = EEMFlow_cdc.warp(x, flow_out)
 *                            replaces: cdc_utils.py:173 followed by EEMFlow+.py:137-149 (K7b followed by K7 EXACT)
 * coarse_flow [B,2,h,w]; flow_init, inter_flow, flow_out [B,2,H,W]; m [B,1,H,W]; x, out [B,C,H,W]. */
EEM_API int eem_upsample_flow_warp(const float* coarse_flow, int h, int w, float scale0, float scale1,
                                   const float* x, int B, int C, int H, int W, float* flow_out, float* out,
                                   eem_stream_t stream);
EEM_API int eem_blend_flow_warp(const float* flow_init, const float* inter_flow, const float* m,
                                const float* x, int B, int C, int H, int W, float* flow_out, float* out,
                                eem_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K8  bilinear resize of flow / meshflow maps
 * replaces: upsample2d_flow_as (model/EEMFlow/cdc_utils.py:80-103; utils_luo/tools.py:3215-3229)
 *             -> align_corners = 1, scale0 = W/w, scale1 = H/h when if_rate else 1
 *           EEMFlow.upsample_flow (model/EEMFlow/EEMFlow.py:118-120), HREM ground-truth upsample
 *           (loader/HREM.py:264-268), upflow8 (model/model_utils.py:30-32)
 *             -> align_corners = 0 (resp. 1 with scale 8)
 * in : [B, C, h, w] -> out : [B, C, H, W]; channel 0 is multiplied by scale0, channel 1 by
 * scale1, further channels by scale_rest. */
EEM_API int eem_bilinear_resize(const float* in, int B, int C, int h, int w, float* out, int H,
                                int W, int align_corners, float scale0, float scale1,
                                float scale_rest, eem_stream_t stream);

/* In-place per-channel scale of channels 0 and 1 of [B, C, h, w]; reproduces the side effect of
 * upsample2d_flow_as(if_rate=True) on its input (model/EEMFlow/cdc_utils.py:85-86). */
EEM_API int eem_scale_uv_inplace(float* flow, int B, int C, int h, int w, float scale0,
                                 float scale1, eem_stream_t stream);

/* K8m  the same two operations for up to 8 maps of one batch in ONE launch each: EEMFlow_cdc resizes its five flow
 * predictions to the input size with five upsample2d_flow_as calls (model/EEMFlow/EEMFlow+.py:231-232).
 * ins/outs/flows, hs, ws, scale0, scale1: HOST arrays of n_maps entries (device pointers / sizes / per-map scales). */
EEM_API int eem_bilinear_resize_multi(const float* const* ins, const int* hs, const int* ws, int n_maps, int B,
                                      int C, float* const* outs, int H, int W, int align_corners,
                                      const float* scale0, const float* scale1, float scale_rest,
                                      eem_stream_t stream);
EEM_API int eem_scale_uv_inplace_multi(float* const* flows, const int* hs, const int* ws, int n_maps, int B, int C,
                                       const float* scale0, const float* scale1, eem_stream_t stream);

/* K9  replicate padding.  replaces: InputPadder.pad (utils/image_utils.py:126-139, F.pad 'replicate')
 * in [B,C,H,W] -> out [B,C,H+top+bottom,W+left+right]. */
EEM_API int eem_replicate_pad(const float* in, int B, int C, int H, int W, int left, int right,
                              int top, int bottom, float* out, eem_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Backward (adjoint) kernels -- training drop-in.  In the reference these
 * gradients come from autograd through spatial_correlation_sampler's backward (what the dead
 * extension model/IRRPWC/correlation_package/correlation_cuda_kernel.cu:117-298 implemented),
 * F.grid_sample and F.interpolate.  Either gradient output pointer may be NULL to skip it.
 * ------------------------------------------------------------------------------------------ */
/* Backward of K5 (training path of CorrBlock.__call__, model/corr.py:29-50 under autograd): the gradient of
 * every pyramid level given the gradient of the lookup output.  grad_levels[l] ([B*H*W, H_l*W_l], HOST array
 * of DEVICE pointers, same layout as K3's levels) is written IN FULL -- zeros outside each position's
 * (2r+2)^2 footprint -- so no memset is needed.  Deterministic (gather, no atomics).  coords get no
 * gradient: the callers detach them (model/eraft.py:141). */
EEM_API int eem_corr_lookup_backward(const float* grad_out, const float* coords, int B, int H, int W,
                                     int num_levels, int radius, float* const* grad_levels,
                                     eem_stream_t stream);

/* Backward of K4: grad_in [n_planes, h, w] (+)= 0.25 * grad_out [n_planes, h/2, w/2] under each coarse cell,
 * 0 in the odd last row/column the floor crops.  accumulate != 0 adds into grad_in. */
EEM_API int eem_avg_pool2x2_backward(const float* grad_out, int64_t n_planes, int h, int w, float* grad_in,
                                     int accumulate, eem_stream_t stream);

/* Batched exact-fp32 GEMM used by the backward of K3 (the gradient of torch.matmul in model/corr.py:58 that the
 * reference obtains from autograd during training, train_mvsec.py:251-258):
 *   C[b] (M x N, row pitch ldc) = alpha * A[b] (M x K, row pitch lda) * op(B[b])  (+ C[b] when accumulate != 0)
 *   op(B) = B given K x N row-major (b_transposed == 0) or B^T with B given N x K row-major (b_transposed != 0);
 *   batch strides in elements.  d fmap1 = sum_l pool^l(fmap2) . dV_l^T and d pool^l(fmap2) = fmap1 . dV_l. */
EEM_API int eem_batched_gemm_f32(const float* A, const float* B, float* C, int batch, int M, int N, int K,
                                 int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
                                 int64_t strideC, int b_transposed, float alpha, int accumulate,
                                 eem_stream_t stream);

/* The same product on the tensor cores (tcgen05 kind::tf32: 10-bit-mantissa operands, fp32 accumulate; TMA-staged
 * operands, TMEM accumulators) -- the precision torch gives the backward of torch.matmul (model/corr.py:58) when TF32
 * is allowed, and what the TF32 forward (EEM_CORR_TF32) uses.  Same arguments as eem_batched_gemm_f32.  Requirements
 * (EEM_ERR_UNSUPPORTED otherwise; eem_batched_gemm_tf32_supported answers without touching the GPU): M % 32 == 0,
 * M <= 256, lda / ldb / strideA / strideB multiples of 4 elements, A and B 16-byte aligned. */
EEM_API int eem_batched_gemm_tf32(const float* A, const float* B, float* C, int batch, int M, int N, int K,
                                  int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
                                  int64_t strideC, int b_transposed, float alpha, int accumulate,
                                  eem_stream_t stream);
/* Several products summed into ONE result in one launch (one K loop over the segments, no read-modify-write of C):
 *   C[b] = alpha * sum_s A_s[b] (M x K_s) * op(B_s[b])   (+ C[b]);   n_seg <= 6; every segment obeys the rules above.
 * This is d fmap1 = sum_l pool^l(fmap2) . dV_l^T over the pyramid levels (model/corr.py:13-27 under autograd). */
EEM_API int eem_batched_gemm_tf32_multi(const float* const* A, const float* const* B, float* C, int n_seg, int batch,
                                        int M, int N, const int* K, const int64_t* lda, const int64_t* ldb,
                                        int64_t ldc, const int64_t* strideA, const int64_t* strideB, int64_t strideC,
                                        int b_transposed, float alpha, int accumulate, eem_stream_t stream);
EEM_API int eem_batched_gemm_tf32_supported(int batch, int M, int N, int K, int64_t lda, int64_t ldb,
                                            int64_t strideA, int64_t strideB, int b_transposed);

/* grad_out [B,n_out,H,W] -> grad_f1, grad_f2 [B,C,H,W]; same index/scale meaning as eem_local_corr. */
EEM_API int eem_local_corr_backward(const float* f1, const float* f2, const float* grad_out, int B,
                                    int C, int H, int W, int max_disp, const int* index, int n_out,
                                    float scale, float* grad_f1, float* grad_f2, eem_stream_t stream);
/* grad_out [B,C,H,W] -> grad_x [B,C,H,W] (zero-filled by the call, accumulated with RED.ADD) and
 * grad_flow [B,2,H,W]; the 0/1 validity mask of mask_mode is treated as a constant. */
EEM_API int eem_backwarp_backward(const float* x, const float* flow, const float* grad_out, int B,
                                  int C, int H, int W, int convention, int mask_mode, float* grad_x,
                                  float* grad_flow, eem_stream_t stream);
/* Backward of eem_bilinear_sample (model/model_utils.py:7-21 under autograd; the reference differentiates
 * F.grid_sample): grad_out [N,C,Ho,Wo] -> grad_img [N,C,H,W] (zero-filled here, 4-tap scatter) and / or
 * grad_coords [N,Ho,Wo,2] (pixel units).  Either output may be NULL. */
EEM_API int eem_bilinear_sample_backward(const float* img, const float* coords, const float* grad_out, int N, int C,
                                         int H, int W, int Ho, int Wo, float* grad_img, float* grad_coords,
                                         eem_stream_t stream);

/* grad_out [B,C,H,W] -> grad_in [B,C,h,w]; adjoint of eem_bilinear_resize with the same scales. */
EEM_API int eem_bilinear_resize_backward(const float* grad_out, int B, int C, int h, int w, int H,
                                         int W, int align_corners, float scale0, float scale1,
                                         float scale_rest, float* grad_in, eem_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Loader / evaluation helpers next to the hot path
 * ------------------------------------------------------------------------------------------------ */

/* Per-pixel event mask of each window: mask[win][y][x] = 1 iff an event falls into pixel bin (x, y).
 * replaces: loader/MVSEC.py:133-142 (np.histogram2d(x, y, bins=(W,H), range=[[0,W],[0,H]]).T > 0):
 * bin = floor(v) for 0 <= v < S, v == S joins the last bin, anything else is ignored.
 * events/offsets as for eem_voxelize (16-byte aligned rows); mask: [n_windows, height, width] uint8. */
EEM_API int eem_event_mask(const double* events, const int64_t* offsets, int n_windows,
                           int64_t max_events_per_window, int height, int width, uint8_t* mask,
                           eem_stream_t stream);

/* Sum of a voxel grid over its bins, accumulated bin by bin in fp32 (bit-exact with numpy):
 * replaces: loader/HREM.py:238-239 (np.sum(event_volume_old, axis=0)).
 * grid [n_windows, num_bins, height, width] -> out [n_windows, height, width]. */
EEM_API int eem_voxel_bin_sum(const float* grid, int64_t n_windows, int num_bins, int height, int width,
                              float* out, eem_stream_t stream);

/* Masked end-point-error statistics of a batch, replaces: test_mvsec.py:291-346 (Test.flow_error).
 * flow_gt, flow_pred: [B, 2, H, W]; event_img: [B, H, W] or NULL ("dense" evaluation); only rows
 * < max_row count (190 for the MVSEC car sequences, H otherwise).  A pixel counts when gt is finite,
 * |gt| > 0 and (event_img > 0 when given).  stats[b] = { n_points, #(EE < 1), #(EE < 3 or EE < 0.1*|gt|),
 * sum EE, sum |gt| } as doubles; EE and |gt| are formed in fp32 exactly as numpy does. */
EEM_API int eem_flow_error(const float* flow_gt, const float* flow_pred, const float* event_img, int B,
                           int height, int width, int max_row, double* stats, eem_stream_t stream);

/* Dense flow -> mesh flow, replaces: loader/HREM.py:41-101 (motion_propagate): per vertex the median
 * (element n/2) of 4*radius clamped samples, then a 5x5 median over the mesh with replicated borders.
 * fflow: [B, H, W, 2] float32 (u, v interleaved, 8-byte aligned); mesh: [B, 2, mesh_size, mesh_size]. */
EEM_API int eem_motion_propagate(const float* fflow, int B, int height, int width, int mesh_size, int radius,
                                 float* mesh, eem_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EEMFLOW_B200_H_ */
