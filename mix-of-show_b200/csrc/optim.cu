// optim.cu — fused flat AdamW step for the ED-LoRA training state (train_edlora.py:57,129-143).
// The trainable state is ONE flat fp32 buffer [concept embedding rows | text-encoder LoRA | UNet LoRA] (+2 logged
// scalars on the wire); after the single NCCL all-reduce of the flat gradient this kernel applies
// the 1/world mean, AdamW with the three learning-rate groups (weight decay 0.01 on all groups, as the reference's
// param groups) and produces Norm_mean of the concept rows (train_edlora.py:138-140) — one launch per optimiser step.
#include "common.h"
#include "tc.cuh"

namespace mos {

struct AdamGroups {
  long long end[3];   // exclusive end offset of group 0 (embedding rows), 1 (text-encoder LoRA), 2 (UNet LoRA)
  float lr[3];
};

__global__ void flat_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, long long n, AdamGroups grp, float beta1, float beta2,
                                  float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
  pdl_wait();
  pdl_launch_dependents();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float lr = i < grp.end[0] ? grp.lr[0] : (i < grp.end[1] ? grp.lr[1] : grp.lr[2]);
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);                     // decoupled weight decay
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

// norm_out[0] = mean_r ||p[r, :]||_2 over the first `rows` rows of width `dim` (one block)
__global__ void row_norm_mean_kernel(const float* __restrict__ p, int rows, int dim, float* __restrict__ norm_out) {
  __shared__ float sh[32];
  __shared__ float total;
  if (threadIdx.x == 0) total = 0.f;
  __syncthreads();
  for (int r = 0; r < rows; ++r) {
    float acc = 0.f;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
      const float x = p[(long long)r * dim + c];
      acc += x * x;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
      total += sqrtf(s);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) norm_out[0] = total / (float)rows;
}


// Re-pack every LoRA pair of the flat training state into the operand layouts of the forward and backward GEMMs
// (one launch per optimiser step).  table [n_mod][8] = {D fp32 [4,K], U fp32 [N,4], K, N,
//   fwd down rows (bf16, 4 rows of pitch K inside the module group's [16,K] block),
//   fwd up rows (fp32 [N,4], pre-scaled by alpha),
//   bwd "down" (bf16 [16,N]: rows 0..3 = U^T), bwd "up" (fp32 [K,4] = alpha D^T)}.
__global__ void lora_pack_kernel(const long long* __restrict__ table, float alpha) {
  pdl_wait();
  pdl_launch_dependents();
  const long long* e = table + (long long)blockIdx.y * 8;
  const float* D = reinterpret_cast<const float*>(e[0]);
  const float* U = reinterpret_cast<const float*>(e[1]);
  const int K = (int)e[2], N = (int)e[3];
  __nv_bfloat16* fdown = reinterpret_cast<__nv_bfloat16*>(e[4]);
  float* fup = reinterpret_cast<float*>(e[5]);
  __nv_bfloat16* bdown = reinterpret_cast<__nv_bfloat16*>(e[6]);
  float* bup = reinterpret_cast<float*>(e[7]);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * K; i += gridDim.x * blockDim.x) {
    const int r = i / K, k = i - r * K;
    const float v = D[i];
    fdown[i] = __float2bfloat16(v);
    if (bup) bup[k * 4 + r] = alpha * v;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * N; i += gridDim.x * blockDim.x) {
    const int n = i >> 2, r = i & 3;
    const float v = U[i];
    fup[i] = alpha * v;
    if (bdown) bdown[(long long)r * N + n] = __float2bfloat16(v);
  }
}

}  // namespace mos

using namespace mos;

extern "C" int mos_flat_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                   const int64_t* group_end, const float* group_lr, float beta1, float beta2,
                                   float eps, float weight_decay, int64_t step, float grad_scale, int32_t emb_rows,
                                   int32_t emb_dim, float* norm_mean_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && group_end && group_lr && n > 0 && step >= 1,
                "mos_flat_adamw_step: bad arguments");
  MOS_CHECK_ARG(group_end[0] <= group_end[1] && group_end[1] <= group_end[2] && group_end[2] == n,
                "mos_flat_adamw_step: group offsets must be increasing and end at n");
  AdamGroups grp;
  for (int i = 0; i < 3; ++i) {
    grp.end[i] = group_end[i];
    grp.lr[i] = group_lr[i];
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  long long blocks = ceil_div(n, 256 * 4);
  if (blocks > 1184) blocks = 1184;
  MOS_CHECK_CUDA(launch_pdl(flat_adamw_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, params, grads, exp_avg,
                            exp_avg_sq, (long long)n, grp, beta1, beta2, eps, weight_decay, (float)bc1,
                            (float)sqrt(bc2), grad_scale));
  if (norm_mean_out && emb_rows > 0) {
    MOS_CHECK_ARG((int64_t)emb_rows * emb_dim <= group_end[0], "mos_flat_adamw_step: embedding rows exceed group 0");
    row_norm_mean_kernel<<<1, 256, 0, stream>>>(params, emb_rows, emb_dim, norm_mean_out);
    MOS_CHECK_LAUNCH();
  }
  return MOS_OK;
}


extern "C" int mos_lora_pack(const int64_t* table_dev, int32_t n_modules, float alpha, void* stream) {
  MOS_CHECK_ARG(table_dev && n_modules > 0, "mos_lora_pack: bad arguments");
  MOS_CHECK_CUDA(launch_pdl(lora_pack_kernel, dim3(8, (unsigned)n_modules), dim3(256), 0,
                            reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const long long*>(table_dev), alpha));
  return MOS_OK;
}
