// norm.cu — K5: GroupNorm(32) statistics / apply(+SiLU) over NHWC bf16 and LayerNorm over token rows.
// HBM-bound glue: 128-bit loads, fp32 statistics, warp-shuffle / shared-memory reductions, deterministic
// (no floating-point atomics to global memory: per-chunk partials are summed in a fixed order).
// Replaces diffusers GroupNorm / LayerNorm call sites inside ResnetBlock2D, Transformer2DModel and
// BasicTransformerBlock (reached from mixofshow/pipelines/pipeline_edlora.py:277).
#include <stdlib.h>

#include "common.h"
#include "tc.cuh"

namespace mos {

constexpr int GN_GROUPS = 32;

// partial[b][chunk][g][2] = (sum, sumsq) over rows [chunk*rows_per_chunk, ...) of batch b
template <bool F16>
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int HW, int C,
                                int rows_per_chunk, float* __restrict__ partial) {
  extern __shared__ float red[];  // [blockDim][16]: per-thread channel sums / sums of squares
  pdl_wait();
  pdl_launch_dependents();
  const int oct = C / 8;
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int lanes = blockDim.x / oct;  // row lanes; blockDim is a multiple of oct
  const int o = threadIdx.x % oct, rl = threadIdx.x / oct;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  const __nv_bfloat16* base = x + ((long long)b * HW) * ldx + o * 8;
  for (int rb = r0 + rl; rb < r1; rb += 4 * lanes) {
    uint4 u4[4];   // four independent 16-byte loads in flight per thread
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = rb + k * lanes;
      u4[k] = r < r1 ? __ldg(reinterpret_cast<const uint4*>(base + (long long)r * ldx)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t w[4] = {u4[k].x, u4[k].y, u4[k].z, u4[k].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = unpack16x2<F16>(w[i]);
        s[2 * i] += f.x;
        q[2 * i] += f.x * f.x;
        s[2 * i + 1] += f.y;
        q[2 * i + 1] += f.y * f.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x * 16 + i] = s[i];
    red[threadIdx.x * 16 + 8 + i] = q[i];
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {  // fixed summation order -> bitwise reproducible statistics
    const int g = threadIdx.x, cpg = C / GN_GROUPS;
    float gs = 0.f, gq = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      for (int l = 0; l < lanes; ++l) {
        const float* t = red + (l * oct + (c >> 3)) * 16 + (c & 7);
        gs += t[0];
        gq += t[8];
      }
    }
    float* dst = partial + (((long long)b * nchunks + chunk) * GN_GROUPS + g) * 2;
    dst[0] = gs;
    dst[1] = gq;
  }
}

template <bool F16>
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int HW, int C,
                                const float* __restrict__ partial, int nchunks, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int silu_act, int rows_per_block,
                                __nv_bfloat16* __restrict__ y, long long ldy) {
  __shared__ float mean[GN_GROUPS], rstd[GN_GROUPS];
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const int cpg = C / GN_GROUPS;
  if (threadIdx.x < GN_GROUPS * 4) {   // blockDim >= 160 always
    // 4 threads per group sum the per-chunk partials (fixed order -> reproducible), then a shuffle tree
    const int g = threadIdx.x >> 2, sub = threadIdx.x & 3;
    float s = 0.f, q = 0.f;
    for (int c = sub; c < nchunks; c += 4) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partial + (((long long)b * nchunks + c) * GN_GROUPS + g) * 2));
      s += v.x;
      q += v.y;
    }
#pragma unroll
    for (int d = 2; d > 0; d >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, d);
      q += __shfl_xor_sync(0xffffffffu, q, d);
    }
    if (sub == 0) {
      const float n = (float)HW * (float)cpg;
      const float m = s / n;
      const float var = fmaxf(q / n - m * m, 0.f);
      mean[g] = m;
      rstd[g] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  const int oct = C / 8;
  const int lanes = blockDim.x / oct;
  const int o = threadIdx.x % oct, rl = threadIdx.x / oct;
  if (rl >= lanes) return;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c = o * 8 + i, g = c / cpg;
    float ga = __ldg(gamma + c) * rstd[g];
    sc[i] = ga;
    sh[i] = __ldg(beta + c) - mean[g] * ga;
  }
  const int r0 = blockIdx.x * rows_per_block, r1 = min(HW, r0 + rows_per_block);
  const __nv_bfloat16* xb = x + ((long long)b * HW) * ldx + o * 8;
  __nv_bfloat16* yb = y + ((long long)b * HW) * ldy + o * 8;
  for (int rb = r0 + rl; rb < r1; rb += 4 * lanes) {
    uint4 u4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = rb + k * lanes;
      if (r < r1) u4[k] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)r * ldx));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = rb + k * lanes;
      if (r < r1) {
        const uint32_t w[4] = {u4[k].x, u4[k].y, u4[k].z, u4[k].w};
        float v[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 f = unpack16x2<F16>(w[i]);
          v[2 * i] = f.x * sc[2 * i] + sh[2 * i];
          v[2 * i + 1] = f.y * sc[2 * i + 1] + sh[2 * i + 1];
        }
        if (silu_act) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = silu(v[i]);
        }
        uint4 out;
        out.x = pack16x2<F16>(v[0], v[1]);
        out.y = pack16x2<F16>(v[2], v[3]);
        out.z = pack16x2<F16>(v[4], v[5]);
        out.w = pack16x2<F16>(v[6], v[7]);
        *reinterpret_cast<uint4*>(yb + (long long)r * ldy) = out;
      }
    }
  }
}

// One warp per row; C <= 1280 (5 octets per lane), two-pass statistics in registers.
template <bool F16>
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long M, int C,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 __nv_bfloat16* __restrict__ y, long long ldy) {
  pdl_wait();
  pdl_launch_dependents();
  const long long row = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int oct = C / 8;
  float v[5][8];
  float s = 0.f;
  const __nv_bfloat16* xr = x + row * ldx;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    int o = lane + k * 32;
    if (o < oct) {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + o * 8));
      uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = unpack16x2<F16>(w[i]);
        v[k][2 * i] = f.x;
        v[k][2 * i + 1] = f.y;
        s += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    int o = lane + k * 32;
    if (o < oct) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float d = v[k][i] - mean;
        q += d * d;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
  const float rstd = rsqrtf(q / (float)C + eps);
  __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    int o = lane + k * 32;
    if (o < oct) {
      float r[8];
      float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8));
      float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8 + 4));
      float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + o * 8));
      float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + o * 8 + 4));
      float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = (v[k][i] - mean) * rstd * gg[i] + bb[i];
      uint4 out;
      out.x = pack16x2<F16>(r[0], r[1]);
      out.y = pack16x2<F16>(r[2], r[3]);
      out.z = pack16x2<F16>(r[4], r[5]);
      out.w = pack16x2<F16>(r[6], r[7]);
      *reinterpret_cast<uint4*>(yr + o * 8) = out;
    }
  }
}

// ---------------------------------------------------------------------------------------------- fused GroupNorm
// Single-kernel GroupNorm(+SiLU): every block keeps its rows in registers, publishes per-chunk partial statistics,
// passes a per-sample grid barrier (monotonic ticket counter, all blocks co-resident by construction: the host caps
// the grid at the occupancy limit), reduces the partials in a fixed order and normalises from registers.  One read
// of x, one write of y, one launch.
constexpr int GN_MAXR = 8;   // rows per thread kept in registers

template <bool F16>
__global__ void __launch_bounds__(320, 2)
gn_fused_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int HW, int C, int rows_per_chunk,
                float* __restrict__ partial, unsigned int* __restrict__ counters, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int silu_act, __nv_bfloat16* __restrict__ y, long long ldy) {
  extern __shared__ float red[];  // [blockDim][16]
  __shared__ float mean[GN_GROUPS], rstd[GN_GROUPS];
  pdl_wait();
  pdl_launch_dependents();
  const int oct = C / 8;
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int lanes = blockDim.x / oct;
  const int o = threadIdx.x % oct, rl = threadIdx.x / oct;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const int cpg = C / GN_GROUPS;
  uint4 rows[GN_MAXR];
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  const __nv_bfloat16* base = x + ((long long)b * HW) * ldx + o * 8;
#pragma unroll
  for (int k = 0; k < GN_MAXR; ++k) {
    const int r = r0 + rl + k * lanes;
    if (r < r1) rows[k] = __ldg(reinterpret_cast<const uint4*>(base + (long long)r * ldx));
  }
#pragma unroll
  for (int k = 0; k < GN_MAXR; ++k) {
    const int r = r0 + rl + k * lanes;
    if (r < r1) {
      const uint32_t w[4] = {rows[k].x, rows[k].y, rows[k].z, rows[k].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack16x2<F16>(w[i]);
        s[2 * i] += f.x;
        q[2 * i] += f.x * f.x;
        s[2 * i + 1] += f.y;
        q[2 * i + 1] += f.y * f.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x * 16 + i] = s[i];
    red[threadIdx.x * 16 + 8 + i] = q[i];
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {
    const int g = threadIdx.x;
    float gs = 0.f, gq = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      for (int l = 0; l < lanes; ++l) {
        const float* t = red + (l * oct + (c >> 3)) * 16 + (c & 7);
        gs += t[0];
        gq += t[8];
      }
    }
    float* dst = partial + (((long long)b * nchunks + chunk) * GN_GROUPS + g) * 2;
    dst[0] = gs;
    dst[1] = gq;
    __threadfence();
  }
  __syncthreads();
  // ---- per-sample grid barrier (sense reversal: count returns to 0, generation increments; zero-initialised once)
  if (threadIdx.x == 0) {
    unsigned int* count = counters + b;
    unsigned int* gen = counters + 32 + b;
    unsigned int my_gen, cur;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(my_gen) : "l"(gen) : "memory");
    const unsigned int old = atomicAdd(count, 1u);
    if (old == (unsigned)nchunks - 1u) {
      *count = 0u;
      __threadfence();
      atomicAdd(gen, 1u);
    } else {
      long long t0 = clock64();
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(gen) : "memory");
        if (clock64() - t0 > 4000000000LL) {
          printf("mos: groupnorm grid-barrier timeout\n");
          __trap();
        }
      } while (cur == my_gen);
    }
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS * 4) {
    const int g = threadIdx.x >> 2, sub = threadIdx.x & 3;
    float ss = 0.f, qq = 0.f;
    for (int c = sub; c < nchunks; c += 4) {
      const float* src = partial + (((long long)b * nchunks + c) * GN_GROUPS + g) * 2;
      float a0, a1;
      asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(a0) : "l"(src) : "memory");   // bypass stale L1
      asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(a1) : "l"(src + 1) : "memory");
      ss += a0;
      qq += a1;
    }
#pragma unroll
    for (int d = 2; d > 0; d >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, d);
      qq += __shfl_xor_sync(0xffffffffu, qq, d);
    }
    if (sub == 0) {
      const float n = (float)HW * (float)cpg;
      const float m = ss / n;
      const float var = fmaxf(qq / n - m * m, 0.f);
      mean[g] = m;
      rstd[g] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = o * 8 + i, g = c / cpg;
    const float ga = __ldg(gamma + c) * rstd[g];
    sc[i] = ga;
    sh[i] = __ldg(beta + c) - mean[g] * ga;
  }
  __nv_bfloat16* yb = y + ((long long)b * HW) * ldy + o * 8;
#pragma unroll
  for (int k = 0; k < GN_MAXR; ++k) {
    const int r = r0 + rl + k * lanes;
    if (r < r1) {
      const uint32_t w[4] = {rows[k].x, rows[k].y, rows[k].z, rows[k].w};
      float v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack16x2<F16>(w[i]);
        v[2 * i] = f.x * sc[2 * i] + sh[2 * i];
        v[2 * i + 1] = f.y * sc[2 * i + 1] + sh[2 * i + 1];
      }
      if (silu_act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = silu(v[i]);
      }
      uint4 out;
      out.x = pack16x2<F16>(v[0], v[1]);
      out.y = pack16x2<F16>(v[2], v[3]);
      out.z = pack16x2<F16>(v[4], v[5]);
      out.w = pack16x2<F16>(v[6], v[7]);
      *reinterpret_cast<uint4*>(yb + (long long)r * ldy) = out;
    }
  }
}


// ---------------------------------------------------------------------------------------------- backward (training)
// GroupNorm(+SiLU) backward for the ED-LoRA training step (trainer_edlora.py:237 reached through loss.backward()):
// gamma / beta are frozen, so only dx is produced.   z = xhat*gamma + beta, y = act(z), dz = dy * act'(z),
//   dx = rstd * (dz*gamma - mean_g(dz*gamma) - xhat * mean_g(dz*gamma*xhat))        (means over the group's HW*cpg)
__device__ __forceinline__ float silu_grad(float z) {
  const float sg = 1.0f / (1.0f + __expf(-z));
  return sg * (1.0f + z * (1.0f - sg));
}

__device__ __forceinline__ void gn_load_stats(const float* __restrict__ partial, int nchunks, int b, int HW, int cpg,
                                              float eps, float* mean, float* rstd) {
  if (threadIdx.x < GN_GROUPS * 4) {
    const int g = threadIdx.x >> 2, sub = threadIdx.x & 3;
    float s = 0.f, q = 0.f;
    for (int c = sub; c < nchunks; c += 4) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partial + (((long long)b * nchunks + c) * GN_GROUPS + g) * 2));
      s += v.x;
      q += v.y;
    }
#pragma unroll
    for (int d = 2; d > 0; d >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, d);
      q += __shfl_xor_sync(0xffffffffu, q, d);
    }
    if (sub == 0) {
      const float n = (float)HW * (float)cpg;
      const float m = s / n;
      mean[g] = m;
      rstd[g] = rsqrtf(fmaxf(q / n - m * m, 0.f) + eps);
    }
  }
}

// partial2[b][chunk][g][2] = (sum dz*gamma, sum dz*gamma*xhat)
__global__ void gn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                     const __nv_bfloat16* __restrict__ dy, long long lddy, int HW, int C,
                                     const float* __restrict__ partial, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, int silu_act, int rows_per_chunk,
                                     float* __restrict__ partial2) {
  extern __shared__ float red[];  // [blockDim][16]
  __shared__ float mean[GN_GROUPS], rstd[GN_GROUPS];
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int cpg = C / GN_GROUPS;
  gn_load_stats(partial, nchunks, b, HW, cpg, eps, mean, rstd);
  __syncthreads();
  const int oct = C / 8;
  const int lanes = blockDim.x / oct;
  const int o = threadIdx.x % oct, rl = threadIdx.x / oct;
  float ga[8], be[8], mu[8], rs[8], a[8], bb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = o * 8 + i, g = c / cpg;
    ga[i] = __ldg(gamma + c);
    be[i] = __ldg(beta + c);
    mu[i] = mean[g];
    rs[i] = rstd[g];
    a[i] = bb[i] = 0.f;
  }
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const __nv_bfloat16* xb = x + ((long long)b * HW) * ldx + o * 8;
  const __nv_bfloat16* db = dy + ((long long)b * HW) * lddy + o * 8;
  for (int r = r0 + rl; r < r1; r += lanes) {
    const uint4 ux = __ldg(reinterpret_cast<const uint4*>(xb + (long long)r * ldx));
    const uint4 ud = __ldg(reinterpret_cast<const uint4*>(db + (long long)r * lddy));
    const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wd[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fx = unpack_bf16x2(wx[i]), fd = unpack_bf16x2(wd[i]);
      const float xv[2] = {fx.x, fx.y}, dv[2] = {fd.x, fd.y};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = 2 * i + h;
        const float xh = (xv[h] - mu[k]) * rs[k];
        float dz = dv[h];
        if (silu_act) dz *= silu_grad(xh * ga[k] + be[k]);
        const float t = dz * ga[k];
        a[k] += t;
        bb[k] += t * xh;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x * 16 + i] = a[i];
    red[threadIdx.x * 16 + 8 + i] = bb[i];
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {  // fixed summation order
    const int g = threadIdx.x;
    float gs = 0.f, gq = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      for (int l = 0; l < lanes; ++l) {
        const float* t = red + (l * oct + (c >> 3)) * 16 + (c & 7);
        gs += t[0];
        gq += t[8];
      }
    }
    float* dst = partial2 + (((long long)b * nchunks + chunk) * GN_GROUPS + g) * 2;
    dst[0] = gs;
    dst[1] = gq;
  }
}

__global__ void gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                    const __nv_bfloat16* __restrict__ dy, long long lddy, int HW, int C,
                                    const float* __restrict__ partial, const float* __restrict__ partial2,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                    int silu_act, int rows_per_chunk, const __nv_bfloat16* __restrict__ add,
                                    long long ldadd, __nv_bfloat16* __restrict__ dx, long long lddx) {
  __shared__ float mean[GN_GROUPS], rstd[GN_GROUPS], ma[GN_GROUPS], mb[GN_GROUPS];
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y, nchunks = gridDim.x;
  const int cpg = C / GN_GROUPS;
  gn_load_stats(partial, nchunks, b, HW, cpg, eps, mean, rstd);
  if (threadIdx.x < GN_GROUPS * 4) {
    const int g = threadIdx.x >> 2, sub = threadIdx.x & 3;
    float s = 0.f, q = 0.f;
    for (int c = sub; c < nchunks; c += 4) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partial2 + (((long long)b * nchunks + c) * GN_GROUPS + g) * 2));
      s += v.x;
      q += v.y;
    }
#pragma unroll
    for (int d = 2; d > 0; d >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, d);
      q += __shfl_xor_sync(0xffffffffu, q, d);
    }
    if (sub == 0) {
      const float n = (float)HW * (float)cpg;
      ma[g] = s / n;
      mb[g] = q / n;
    }
  }
  __syncthreads();
  const int oct = C / 8;
  const int lanes = blockDim.x / oct;
  const int o = threadIdx.x % oct, rl = threadIdx.x / oct;
  float ga[8], be[8], mu[8], rs[8], A[8], Bm[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = o * 8 + i, g = c / cpg;
    ga[i] = __ldg(gamma + c);
    be[i] = __ldg(beta + c);
    mu[i] = mean[g];
    rs[i] = rstd[g];
    A[i] = ma[g];
    Bm[i] = mb[g];
  }
  const int r0 = blockIdx.x * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const __nv_bfloat16* xb = x + ((long long)b * HW) * ldx + o * 8;
  const __nv_bfloat16* db = dy + ((long long)b * HW) * lddy + o * 8;
  const __nv_bfloat16* ab = add ? add + ((long long)b * HW) * ldadd + o * 8 : nullptr;
  __nv_bfloat16* ob = dx + ((long long)b * HW) * lddx + o * 8;
  for (int r = r0 + rl; r < r1; r += lanes) {
    const uint4 ux = __ldg(reinterpret_cast<const uint4*>(xb + (long long)r * ldx));
    const uint4 ud = __ldg(reinterpret_cast<const uint4*>(db + (long long)r * lddy));
    uint4 ua = make_uint4(0, 0, 0, 0);
    if (ab) ua = __ldg(reinterpret_cast<const uint4*>(ab + (long long)r * ldadd));
    const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wd[4] = {ud.x, ud.y, ud.z, ud.w}, wa[4] = {ua.x, ua.y, ua.z, ua.w};
    float v[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fx = unpack_bf16x2(wx[i]), fd = unpack_bf16x2(wd[i]), fa = unpack_bf16x2(wa[i]);
      const float xv[2] = {fx.x, fx.y}, dv[2] = {fd.x, fd.y}, av[2] = {fa.x, fa.y};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = 2 * i + h;
        const float xh = (xv[h] - mu[k]) * rs[k];
        float dz = dv[h];
        if (silu_act) dz *= silu_grad(xh * ga[k] + be[k]);
        v[k] = rs[k] * (dz * ga[k] - A[k] - xh * Bm[k]) + av[h];
      }
    }
    uint4 out;
    out.x = pack_bf16x2(v[0], v[1]);
    out.y = pack_bf16x2(v[2], v[3]);
    out.z = pack_bf16x2(v[4], v[5]);
    out.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(ob + (long long)r * lddx) = out;
  }
}

// LayerNorm backward, one warp per row (C <= 1280); statistics recomputed from x in registers.
__global__ void layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                     const __nv_bfloat16* __restrict__ dy, long long lddy, long long M, int C,
                                     const float* __restrict__ gamma, float eps, const __nv_bfloat16* __restrict__ add,
                                     long long ldadd, __nv_bfloat16* __restrict__ dx, long long lddx) {
  pdl_wait();
  pdl_launch_dependents();
  const long long row = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int oct = C / 8;
  float v[5][8], g[5][8];
  float s = 0.f;
  const __nv_bfloat16* xr = x + row * ldx;
  const __nv_bfloat16* dr = dy + row * lddy;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int o = lane + k * 32;
    if (o < oct) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + o * 8));
      const uint4 d = __ldg(reinterpret_cast<const uint4*>(dr + o * 8));
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + o * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const uint32_t w[4] = {u.x, u.y, u.z, u.w}, wd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16x2(w[i]), fd = unpack_bf16x2(wd[i]);
        v[k][2 * i] = f.x;
        v[k][2 * i + 1] = f.y;
        g[k][2 * i] = fd.x * gg[2 * i];
        g[k][2 * i + 1] = fd.y * gg[2 * i + 1];
        s += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (lane + k * 32 < oct) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[k][i] - mean;
        q += d * d;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
  const float rstd = rsqrtf(q / (float)C + eps);
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (lane + k * 32 < oct) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[k][i] = (v[k][i] - mean) * rstd;   // xhat
        sa += g[k][i];
        sb += g[k][i] * v[k][i];
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, d);
    sb += __shfl_xor_sync(0xffffffffu, sb, d);
  }
  sa /= (float)C;
  sb /= (float)C;
  __nv_bfloat16* outr = dx + row * lddx;
  const __nv_bfloat16* ar = add ? add + row * ldadd : nullptr;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int o = lane + k * 32;
    if (o < oct) {
      float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (ar) {
        const uint4 ua = __ldg(reinterpret_cast<const uint4*>(ar + o * 8));
        const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack_bf16x2(wa[i]);
          a8[2 * i] = f.x;
          a8[2 * i + 1] = f.y;
        }
      }
      float r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = rstd * (g[k][i] - sa - v[k][i] * sb) + a8[i];
      uint4 out;
      out.x = pack_bf16x2(r[0], r[1]);
      out.y = pack_bf16x2(r[2], r[3]);
      out.z = pack_bf16x2(r[4], r[5]);
      out.w = pack_bf16x2(r[6], r[7]);
      *reinterpret_cast<uint4*>(outr + o * 8) = out;
    }
  }
}

static int gn_block_threads(int C) {
  int oct = C / 8;
  int k = 320 / oct;
  if (k < 1) k = 1;
  return oct * k;
}

}  // namespace mos

using namespace mos;

// GroupNorm(32) + optional SiLU:  y[b, r, c] = act((x - mean_bg) * rstd_bg * gamma_c + beta_c)
// x: bf16 [B, HW, ldx] (first C channels), y: bf16 [B, HW, ldy]; partial: fp32 workspace [B, nchunks, 32, 2],
// nchunks = *nchunks_io (0 = choose; the chosen value is returned through the pointer).
extern "C" int mos_groupnorm_fwd(const void* x, int64_t ldx, int32_t B, int32_t HW, int32_t C, const float* gamma,
                                 const float* beta, float eps, int32_t silu_act, float* partial,
                                 int32_t partial_capacity_floats, void* y, int64_t ldy, int32_t act_dtype,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(x && y && gamma && beta && partial, "mos_groupnorm_fwd: NULL pointer");
  MOS_CHECK_DTYPE(act_dtype, "mos_groupnorm_fwd");
  const bool f16 = act_dtype == MOS_DT_F16;
  MOS_CHECK_ARG(C % 32 == 0 && C % 8 == 0 && C <= 2560 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C,
                "mos_groupnorm_fwd: bad C=%d ldx=%lld ldy=%lld", C, (long long)ldx, (long long)ldy);
  const int threads = gn_block_threads(C);
  const int lanes = threads / (C / 8);
  // ---- fused single-launch path: grid capped at the co-resident capacity (the kernel contains a grid barrier)
  static int capacity = 0, use_fused = -1;
  if (use_fused < 0) {
    // measured slower than the two-launch path on B200 (grid barrier + 2 blocks/SM): opt-in only
    const char* e = getenv("MOS_GN_FUSED");
    use_fused = (e && e[0] == '1') ? 1 : 0;
  }
  if (use_fused && capacity == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    MOS_CHECK_CUDA(cudaGetDevice(&dev));
    MOS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MOS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel<false>, 320, 320 * 16 * sizeof(float)));
    capacity = sms * (per_sm > 0 ? per_sm : 1);
  }
  if (use_fused) {
    int nchunks = capacity / B;
    if (nchunks > (int)ceil_div(HW, lanes)) nchunks = (int)ceil_div(HW, lanes);
    if (nchunks < 1) nchunks = 1;
    int rows_per_chunk = (int)ceil_div(HW, nchunks);
    nchunks = (int)ceil_div(HW, rows_per_chunk);
    const long long need = (long long)B * nchunks * GN_GROUPS * 2 + 64;   // + per-sample ticket counters
    if (rows_per_chunk <= GN_MAXR * lanes && B <= 32 && need <= partial_capacity_floats && B * nchunks <= capacity) {
      unsigned int* counters = reinterpret_cast<unsigned int*>(partial + partial_capacity_floats - 64);
      MOS_CHECK_CUDA(launch_pdl(f16 ? gn_fused_kernel<true> : gn_fused_kernel<false>, dim3(nchunks, B), dim3(threads), threads * 16 * sizeof(float), stream,
                                reinterpret_cast<const __nv_bfloat16*>(x), (long long)ldx, (int)HW, (int)C,
                                rows_per_chunk, partial, counters, gamma, beta, eps, (int)silu_act,
                                reinterpret_cast<__nv_bfloat16*>(y), (long long)ldy));
      return MOS_OK;
    }
  }
  // ---- two-launch fallback (very large maps)
  // aim for ~4 blocks per SM overall
  int nchunks = (int)ceil_div(2368, B);     // up to ~16 blocks per SM in flight: latency-bound kernels want parallelism
  int min_rows = 4 * (threads / (C / 8));
  if (nchunks > (int)ceil_div(HW, min_rows)) nchunks = (int)ceil_div(HW, min_rows);
  {
    const long long cap = ((long long)partial_capacity_floats - 64) / ((long long)B * GN_GROUPS * 2);
    if (nchunks > cap) nchunks = (int)cap;
  }
  if (nchunks < 1) nchunks = 1;
  int rows_per_chunk = (int)ceil_div(HW, nchunks);
  nchunks = (int)ceil_div(HW, rows_per_chunk);
  MOS_CHECK_ARG((long long)B * nchunks * GN_GROUPS * 2 <= partial_capacity_floats,
                "mos_groupnorm_fwd: partial workspace too small (need %lld floats)",
                (long long)B * nchunks * GN_GROUPS * 2);
  MOS_CHECK_CUDA(launch_pdl(f16 ? gn_stats_kernel<true> : gn_stats_kernel<false>, dim3(nchunks, B), dim3(threads),
                            threads * 16 * sizeof(float), stream, reinterpret_cast<const __nv_bfloat16*>(x), (long long)ldx,
                            (int)HW, (int)C, rows_per_chunk, partial));
  MOS_CHECK_CUDA(launch_pdl(f16 ? gn_apply_kernel<true> : gn_apply_kernel<false>, dim3(nchunks, B), dim3(threads), 0, stream,
                            reinterpret_cast<const __nv_bfloat16*>(x), (long long)ldx, (int)HW, (int)C,
                            (const float*)partial, nchunks, gamma, beta, eps, (int)silu_act, rows_per_chunk,
                            reinterpret_cast<__nv_bfloat16*>(y), (long long)ldy));
  return MOS_OK;
}

extern "C" int mos_layernorm_fwd(const void* x, int64_t ldx, int64_t M, int32_t C, const float* gamma,
                                 const float* beta, float eps, void* y, int64_t ldy, int32_t act_dtype, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(x && y && gamma && beta, "mos_layernorm_fwd: NULL pointer");
  MOS_CHECK_DTYPE(act_dtype, "mos_layernorm_fwd");
  MOS_CHECK_ARG(C % 8 == 0 && C <= 1280 && ldx % 8 == 0 && ldy % 8 == 0, "mos_layernorm_fwd: bad C=%d", C);
  const int warps = 8;
  MOS_CHECK_CUDA(launch_pdl(act_dtype == MOS_DT_F16 ? layernorm_kernel<true> : layernorm_kernel<false>,
                            dim3((unsigned)ceil_div(M, warps)), dim3(warps * 32), 0, stream,
                            reinterpret_cast<const __nv_bfloat16*>(x), (long long)ldx, (long long)M, (int)C, gamma, beta,
                            eps, reinterpret_cast<__nv_bfloat16*>(y), (long long)ldy));
  return MOS_OK;
}


// GroupNorm(+SiLU) backward (frozen affine):  dx = d/dx [ act(GN(x)) ] . dy  (+ add).  Statistics are recomputed from x.
// workspace: fp32, >= 2 * B * nchunks * 64 floats (the entry point picks nchunks to fit `workspace_floats`).
extern "C" int mos_groupnorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, int32_t B, int32_t HW,
                                 int32_t C, const float* gamma, const float* beta, float eps, int32_t silu_act,
                                 float* workspace, int32_t workspace_floats, const void* add, int64_t ldadd, void* dx,
                                 int64_t lddx, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(x && dy && dx && gamma && beta && workspace, "mos_groupnorm_bwd: NULL pointer");
  MOS_CHECK_ARG(C % 32 == 0 && C <= 2560 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && ldx >= C && lddy >= C &&
                    lddx >= C && (!add || (ldadd % 8 == 0 && ldadd >= C)),
                "mos_groupnorm_bwd: bad C=%d / pitches", C);
  const int threads = gn_block_threads(C);
  int nchunks = (int)ceil_div(1184, B);
  const int min_rows = 4 * (threads / (C / 8));
  if (nchunks > (int)ceil_div(HW, min_rows)) nchunks = (int)ceil_div(HW, min_rows);
  {
    const long long cap = (long long)workspace_floats / ((long long)B * GN_GROUPS * 4);
    if (nchunks > cap) nchunks = (int)cap;
  }
  MOS_CHECK_ARG(nchunks >= 1, "mos_groupnorm_bwd: workspace too small (need >= %d floats)", B * GN_GROUPS * 4);
  const int rows_per_chunk = (int)ceil_div(HW, nchunks);
  nchunks = (int)ceil_div(HW, rows_per_chunk);
  float* p1 = workspace;
  float* p2 = workspace + (long long)B * nchunks * GN_GROUPS * 2;
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* db = reinterpret_cast<const __nv_bfloat16*>(dy);
  MOS_CHECK_CUDA(launch_pdl(gn_stats_kernel<false>, dim3(nchunks, B), dim3(threads), threads * 16 * sizeof(float), stream, xb,
                            (long long)ldx, (int)HW, (int)C, rows_per_chunk, p1));
  MOS_CHECK_CUDA(launch_pdl(gn_bwd_reduce_kernel, dim3(nchunks, B), dim3(threads), threads * 16 * sizeof(float), stream,
                            xb, (long long)ldx, db, (long long)lddy, (int)HW, (int)C, (const float*)p1, gamma, beta, eps,
                            (int)silu_act, rows_per_chunk, p2));
  MOS_CHECK_CUDA(launch_pdl(gn_bwd_apply_kernel, dim3(nchunks, B), dim3(threads), 0, stream, xb, (long long)ldx, db,
                            (long long)lddy, (int)HW, (int)C, (const float*)p1, (const float*)p2, gamma, beta, eps,
                            (int)silu_act, rows_per_chunk, reinterpret_cast<const __nv_bfloat16*>(add), (long long)ldadd,
                            reinterpret_cast<__nv_bfloat16*>(dx), (long long)lddx));
  return MOS_OK;
}

// LayerNorm backward (frozen affine): dx = d/dx LN(x) . dy (+ add); one warp per row.
extern "C" int mos_layernorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, int64_t M, int32_t C,
                                 const float* gamma, float eps, const void* add, int64_t ldadd, void* dx, int64_t lddx,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(x && dy && dx && gamma, "mos_layernorm_bwd: NULL pointer");
  MOS_CHECK_ARG(C % 8 == 0 && C <= 1280 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && (!add || ldadd % 8 == 0),
                "mos_layernorm_bwd: bad C=%d / pitches", C);
  const int warps = 8;
  MOS_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel, dim3((unsigned)ceil_div(M, warps)), dim3(warps * 32), 0, stream,
                            reinterpret_cast<const __nv_bfloat16*>(x), (long long)ldx,
                            reinterpret_cast<const __nv_bfloat16*>(dy), (long long)lddy, (long long)M, (int)C, gamma, eps,
                            reinterpret_cast<const __nv_bfloat16*>(add), (long long)ldadd,
                            reinterpret_cast<__nv_bfloat16*>(dx), (long long)lddx));
  return MOS_OK;
}
