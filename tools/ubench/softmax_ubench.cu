// Micro-benchmark behind the attention-kernel design (profiles/README.md): cycles per 128-column softmax pass of one
// warp (one thread per row, S in TMEM) with individual pieces removed, at 1 and 2 resident CTAs per SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../mix-of-show_b200/csrc softmax_ubench.cu -o softmax_ubench
#include <cstdio>
#include <cuda_runtime.h>
#include "tc.cuh"
using namespace mos;

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MODE bits: 1 = TMEM loads, 2 = MUFU exp (else FMUL), 4 = max tracking, 8 = pack + st.shared, 16 = 32-wide loads
template <int MODE>
__global__ void __launch_bounds__(128, 2) pass_kernel(float* out, long long* cyc, int iters, float c, float nm) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&holder, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
  const int r = threadIdx.x;
  const uint32_t sP = smem_u32(smem_raw);
  float s4[4] = {0, 0, 0, 0}, m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
  uint32_t v[2][16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { v[0][i] = __float_as_uint(0.01f * (i + lane)); v[1][i] = __float_as_uint(-0.02f * (i + warp)); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE & 1) tmem_ld16(trow, v[0]);
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      if (MODE & 1) {
        tmem_ld_wait();
        if (ch + 1 < 8) tmem_ld16(trow + (ch + 1) * 16, v[(ch + 1) & 1]);
      }
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float x0 = __uint_as_float(v[ch & 1][i]), x1 = __uint_as_float(v[ch & 1][i + 1]);
        if (MODE & 4) { m4[i & 3] = fmaxf(m4[i & 3], x0); m4[(i + 1) & 3] = fmaxf(m4[(i + 1) & 3], x1); }
        float e0 = fmaf(x0, c, nm), e1 = fmaf(x1, c, nm);
        if (MODE & 2) { e0 = ex2a(e0); e1 = ex2a(e1); } else { e0 *= 1.0001f; e1 *= 0.9999f; }
        s4[i & 3] += e0; s4[(i + 1) & 3] += e1;
        pk[i >> 1] = pack_bf16x2(e0, e1);
      }
      if (MODE & 8) {
        const uint32_t rowp = sP + (ch >> 2) * 16384 + r * 128;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c16 = (ch & 3) * 2 + g;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + ((c16 ^ (r & 7)) << 4)), "r"(pk[g * 4]),
                       "r"(pk[g * 4 + 1]), "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3]) : "memory");
        }
      } else {
#pragma unroll
        for (int g = 0; g < 8; ++g) s4[g & 3] += __uint_as_float(pk[g]);
      }
      if (!(MODE & 1)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[(ch + 1) & 1][i] ^= (uint32_t)(ch + it) << 3;   // keep the values moving
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * 128 + threadIdx.x] = s4[0] + s4[1] + s4[2] + s4[3] + m4[0] + m4[1] + m4[2] + m4[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// pure MUFU throughput: warps x 32 lanes, 8 independent chains
__global__ void mufu_kernel(float* out, long long* cyc, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ex2a(a[i]) - 1.0f;
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int ctas_per_sm) {
  float* out; long long* cyc;
  const int grid = 148 * ctas_per_sm, iters = 200;
  cudaMalloc(&out, grid * 128 * 4); cudaMalloc(&cyc, grid * 8);
  cudaFuncSetAttribute(pass_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int k = 0; k < 2; ++k) pass_kernel<MODE><<<grid, 128, 100 * 1024>>>(out, cyc, iters, 0.3f, -0.1f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < grid; ++i) s += (double)h[i];
  printf("%-44s CTAs/SM %d: %8.0f cycles per 128-column pass per warp  (%s)\n", name, ctas_per_sm, s / grid / iters,
         cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int c = 1; c <= 2; ++c) {
    run<1 | 2 | 4 | 8>("full pass (ld, max, exp, sum, pack, sts)", c);
    run<1 | 4 | 8>("no MUFU", c);
    run<2 | 4 | 8>("no TMEM loads", c);
    run<1 | 2 | 4>("no st.shared", c);
    run<2>("exp + sum + pack only", c);
    run<1>("TMEM loads + fma/sum/pack only", c);
  }
  for (int warps = 4; warps <= 32; warps *= 2) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * warps * 32 * 4); cudaMalloc(&cyc, 148 * 8);
    mufu_kernel<<<148, warps * 32>>>(out, cyc, 2000);
    mufu_kernel<<<148, warps * 32>>>(out, cyc, 2000);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; ++i) s += (double)h[i];
    printf("MUFU.EX2 only, %2d warps/SM: %.2f ex2 per clock per SM\n", warps, 2000.0 * 8 * warps * 32 / (s / 148));
    cudaFree(out); cudaFree(cyc);
  }
  return 0;
}
