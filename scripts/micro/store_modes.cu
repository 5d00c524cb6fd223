// Micro-experiment: DRAM read traffic and duration of a write-only streaming kernel under different store
// flavours / vector widths / row alignments (B200).  Build: nvcc -arch=sm_100a -O3 -o store_modes store_modes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template <int MODE, int V>
__global__ void __launch_bounds__(256) fill(float* out, int H, int W, int planes) {
  const int X = (blockIdx.x * 32 + (threadIdx.x & 31)) * V;
  const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (X >= W || Y >= H) return;
  for (int p = blockIdx.z; p < planes; p += gridDim.z) {
    float* o = out + (size_t)p * H * W + (size_t)Y * W + X;
    float v = (float)(p + X);
    if (V == 1) {
      if (MODE == 0) *o = v;
      else if (MODE == 1) asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(o), "f"(v) : "memory");
      else if (MODE == 2) asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(o), "f"(v) : "memory");
      else asm volatile("st.global.wt.f32 [%0], %1;" ::"l"(o), "f"(v) : "memory");
    } else if (V == 2) {
      if (MODE == 0) *reinterpret_cast<float2*>(o) = make_float2(v, v);
      else if (MODE == 1) asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(o), "f"(v), "f"(v) : "memory");
      else if (MODE == 2) asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(o), "f"(v), "f"(v) : "memory");
      else asm volatile("st.global.wt.v2.f32 [%0], {%1,%2};" ::"l"(o), "f"(v), "f"(v) : "memory");
    } else {
      if (MODE == 0) *reinterpret_cast<float4*>(o) = make_float4(v, v, v, v);
      else if (MODE == 1) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
      else if (MODE == 2) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
      else asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    }
  }
}
template <int MODE, int V>
void run(const char* name, float* out, float* flush, size_t flush_n, int H, int W, int planes, int gz) {
  dim3 grid((W / V + 31) / 32, (H + 7) / 8, gz);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int it = 0; it < 6; ++it) {
    cudaMemsetAsync(flush, it, flush_n);
    cudaEventRecord(a);
    fill<MODE, V><<<grid, 256>>>(out, H, W, planes);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  double mb = (double)planes * H * W * 4 / 1e6;
  printf("%-34s W=%d V=%d gz=%d: %7.1f us  %7.1f GB/s (%.1f MB)\n", name, W, V, gz, best * 1e3, mb / best, mb);
}
int main() {
  const int H = 260, planes = 64;
  float *out, *flush; size_t flush_n = 512u << 20;
  cudaMalloc(&out, (size_t)planes * H * 352 * 4 + 256); cudaMalloc(&flush, flush_n);
  for (int W : {346, 352}) {
    for (int gz : {6, 16, 64}) {
      run<0, 1>("st.global", out, flush, flush_n, H, W, planes, gz);
      run<1, 1>("st.global.L1::no_allocate", out, flush, flush_n, H, W, planes, gz);
      run<0, 2>("st.global", out, flush, flush_n, H, W, planes, gz);
      run<1, 2>("st.global.L1::no_allocate", out, flush, flush_n, H, W, planes, gz);
      run<2, 2>("st.global.cs", out, flush, flush_n, H, W, planes, gz);
      run<3, 2>("st.global.wt", out, flush, flush_n, H, W, planes, gz);
      if (W % 4 == 0) { run<0, 4>("st.global", out, flush, flush_n, H, W, planes, gz); run<1, 4>("st.global.L1::no_allocate", out, flush, flush_n, H, W, planes, gz); }
    }
  }
  // a plain memset of the same size for scale
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaMemsetAsync(flush, 1, flush_n); cudaEventRecord(a); cudaMemsetAsync(out, 0, (size_t)planes * H * 346 * 4); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); printf("cudaMemset same size: %.1f us\n", ms * 1e3);
  return 0;
}
