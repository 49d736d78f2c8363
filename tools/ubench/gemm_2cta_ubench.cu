// ROUND-2 CANDIDATE, NOT YET RUN ON A GPU (the round-1 GPU budget ended first).  Stand-alone experiment, not linked into
// libmos_sm100.so and not part of the test suite: run it under `timeout` on a B200 before adopting anything from it.
//
//   C[M, N] (bf16) = A[M, K] W[N, K]^T       with tcgen05.mma.cta_group::2 (UMMA 256 x 160 x 16)
//
// Why: profiles/README.md — the long-K GEMMs / 3x3 convolutions of the denoise step run at 0.36 us per 64-deep k-block
// with the 1-CTA kernel, which is the L2->SM fabric limit (128 CTAs x 36 KB per k-block).  A CTA pair shares the W tile:
// each CTA fetches its own 128 rows of A (16 KB) but only HALF of the W tile (80 of 160 rows, 10 KB) per k-block, i.e.
// 26 KB instead of 36 KB for the same MMA work per SM.  PTX forms follow the CUTLASS headers in this image
// (cute/arch/copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_2D, cutlass/arch/barrier.h umma_arrive_multicast_2x1SM,
// cute/arch/tmem_allocator_sm100.hpp Allocator2Sm).
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I../../mix-of-show_b200/csrc \
//        gemm_2cta_ubench.cu ../../mix-of-show_b200/csrc/common.cu -o gemm_2cta_ubench
//   timeout 60 ./gemm_2cta_ubench            # prints max |error| vs a CUDA-core reference and TFLOP/s per shape
//
// Structure per CTA (192 threads): warp 0 = TMA producer (A rows of this CTA, W half of this CTA; transaction bytes of
// BOTH CTAs land on the LEADER's full barrier), warp 1 = TMEM allocator (both CTAs) + MMA issuer (leader only; commits
// multicast to both CTAs' empty / accumulator-ready barriers), warps 2..5 = epilogue (each CTA drains its own 128 rows).
// One 256 x 160 tile per CTA pair (non-persistent: this is a mainloop experiment).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.h"
#include "tc.cuh"

using namespace mos;

constexpr int BM = 128, BN = 160, BK = 64, STAGES = 7;
constexpr int A_BYTES = BM * BK * 2;            // 16384
constexpr int BH_BYTES = (BN / 2) * BK * 2;     // 10240: this CTA's half of the W tile
constexpr int STAGE_BYTES = A_BYTES + BH_BYTES; // 26624 (a multiple of 1024: every tile stays 1024-byte aligned)
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;     // cute::Sm100MmaPeerBitMask: address of the same object in CTA 0

__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  // executed by both CTAs; the mbarrier address has its peer bit cleared, so the bytes count on the leader's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {   // same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                 __nv_bfloat16* __restrict__ C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int n_tiles = N / BN;
  const int tn = pair % n_tiles, tm = pair / n_tiles;
  const int m0 = tm * 2 * BM + rank * BM;           // this CTA's 128 rows of the 256-row pair tile
  const int n0 = tn * BN;
  const int kb_total = K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);      // leader: one arrive.expect_tx covering the bytes of both CTAs (unused in CTA 1)
      mbar_init(&empty_bar[s], 1);     // one multicast tcgen05.commit from the leader's MMA thread
    }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(&tmem_holder, 256);
  tc_fence_before();
  cluster_sync_all();                  // peers see initialised barriers before any remote complete_tx / commit
  tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * STAGE_BYTES);
        tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BK, m0);
        tma_load_2d_2sm(sa + A_BYTES, &tmB, &full_bar[stage], kb * BK, n0 + rank * (BN / 2));
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(2 * BM, BN, 1);      // M = 256 over the CTA pair
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        uint8_t* sa = smem + stage * STAGE_BYTES;
        const uint64_t adesc = make_desc_sw128(smem_u32(sa));
        const uint64_t bdesc = make_desc_sw128(smem_u32(sa + A_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16_2cta(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit_2cta(&empty_bar[stage], 0b11);           // frees the slot in BOTH CTAs
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_2cta(&acc_bar, 0b11);                      // accumulators of both CTAs are complete
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    if (lane == 0) mbar_wait(&acc_bar, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t trow = tmem + (uint32_t(q * 32) << 16);
    const long long m = (long long)m0 + r;
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
      uint32_t v[16];
      tmem_ld16(trow + c * 16, v);
      tmem_ld_wait();
      if (m < M) {
        uint4 u0, u1;
        u0.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]));
        u0.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
        u0.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]));
        u0.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
        u1.x = pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9]));
        u1.y = pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11]));
        u1.z = pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13]));
        u1.w = pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15]));
        __nv_bfloat16* dst = C + m * N + n0 + c * 16;
        *reinterpret_cast<uint4*>(dst) = u0;
        *reinterpret_cast<uint4*>(dst + 8) = u1;
      }
    }
    tc_fence_before();
  }
  cluster_sync_all();                  // nobody exits (or frees TMEM) while the peer may still touch this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem, 256);
  }
}

__global__ void ref_gemm_kernel(const __nv_bfloat16* A, const __nv_bfloat16* W, int M, int N, int K, float* C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * N) return;
  const int m = (int)(idx / N), n = (int)(idx % N);
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(long long)m * K + k]) * __bfloat162float(W[(long long)n * K + k]);
  C[idx] = acc;
}
__global__ void fill_kernel(__nv_bfloat16* x, long long n, unsigned seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)i * 2654435761u + seed;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  x[i] = __float2bfloat16(((h & 0xFFFF) / 65536.0f - 0.5f));
}
__global__ void maxerr_kernel(const __nv_bfloat16* c, const float* ref, long long n, float* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = fabsf(__bfloat162float(c[i]) - ref[i]);
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(e));      // e >= 0: the int order is the float order
  atomicMax(reinterpret_cast<int*>(out + 1), __float_as_int(fabsf(ref[i])));
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static int run(int M, int N, int K) {
  if (M % (2 * BM) || N % BN || K % BK) { printf("skip %d %d %d\n", M, N, K); return 0; }
  __nv_bfloat16 *A, *W, *C; float *R, *err;
  CK(cudaMalloc(&A, (size_t)M * K * 2)); CK(cudaMalloc(&W, (size_t)N * K * 2)); CK(cudaMalloc(&C, (size_t)M * N * 2));
  CK(cudaMalloc(&R, (size_t)M * N * 4)); CK(cudaMalloc(&err, 8)); CK(cudaMemset(err, 0, 8)); CK(cudaMemset(C, 0xFF, (size_t)M * N * 2));
  fill_kernel<<<(unsigned)(((long long)M * K + 255) / 256), 256>>>(A, (long long)M * K, 1u);
  fill_kernel<<<(unsigned)(((long long)N * K + 255) / 256), 256>>>(W, (long long)N * K, 2u);
  CUtensorMap tmA, tmB;
  { uint64_t d[2] = {(uint64_t)K, (uint64_t)M}, s[1] = {(uint64_t)K * 2}; uint32_t b[2] = {BK, BM};
    if (encode_tmap(&tmA, A, 2, 2, d, s, b, 3)) { printf("tmap A: %s\n", mos_last_error()); return 1; } }
  { uint64_t d[2] = {(uint64_t)K, (uint64_t)N}, s[1] = {(uint64_t)K * 2}; uint32_t b[2] = {BK, BN / 2};
    if (encode_tmap(&tmB, W, 2, 2, d, s, b, 3)) { printf("tmap B: %s\n", mos_last_error()); return 1; } }
  const size_t smem = STAGES * STAGE_BYTES + 1024;
  CK(cudaFuncSetAttribute(gemm_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = 2u * (M / (2 * BM)) * (N / BN);
  gemm_2cta_kernel<<<grid, 192, smem>>>(tmA, tmB, M, N, K, C);
  CK(cudaDeviceSynchronize());
  ref_gemm_kernel<<<(unsigned)(((long long)M * N + 255) / 256), 256>>>(A, W, M, N, K, R);
  maxerr_kernel<<<(unsigned)(((long long)M * N + 255) / 256), 256>>>(C, R, (long long)M * N, err);
  float h[2]; CK(cudaMemcpy(h, err, 8, cudaMemcpyDeviceToHost));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 20; ++i) gemm_2cta_kernel<<<grid, 192, smem>>>(tmA, tmB, M, N, K, C);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms / 20 * 1e3, tf = 2.0 * M * N * K / (us * 1e-6) / 1e12;
  const int waves = (int)((grid / 2 + 73) / 74);
  printf("M=%5d N=%5d K=%5d: max|err| %.4f (max|ref| %.2f)  %8.1f us  %7.1f TFLOP/s  %.3f us per k-block and wave (%u CTA pairs)\n",
         M, N, K, h[0], h[1], us, tf, us / (K / BK) / waves, grid / 2);
  cudaFree(A); cudaFree(W); cudaFree(C); cudaFree(R); cudaFree(err);
  return 0;
}

int main() {
  int rc = 0;
  rc |= run(256, 160, 64);          // one pair, one k-block: the smallest possible correctness case
  rc |= run(512, 320, 640);
  rc |= run(8192, 320, 2880);       // FLOPs of the res-64 3x3 conv 320 -> 320 (1-CTA kernel: 29 us cold / 0.36 us per k-block)
  rc |= run(8192, 640, 5760);
  rc |= run(8192, 2560, 320);
  rc |= run(16384, 1280, 1280);
  return rc;
}
