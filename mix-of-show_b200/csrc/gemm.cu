// gemm.cu — K1/K4: warp-specialised tcgen05 GEMM and implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[m,n] = epilogue( sum_k A[m,k] W[n,k] )        A, W bf16 K-major; fp32 accumulation in TMEM
//
// One CTA computes one 128 x 160 output tile (160 divides every SD1.5 channel count: 320/640/1280/...),
// optionally one split of the K range.  Roles (192 threads):
//   warp 0      TMA producer   cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x (160|176) x 16)
//   warps 2..5  epilogue       tcgen05.ld (32x32b) -> registers -> fused epilogue -> global
// LoRA fusion (edlora.py:244-246): the rank-padded down matrix [16, K] rides along as 16 extra B rows, so the
// same MMA also produces t = x * down^T in TMEM columns 160..175; the epilogue adds t * (alpha*up)^T.
// Convolution: the A tile is a TW x TH x TB pixel patch of the NHWC activation fetched by a 4-D tensor map at
// the tap-shifted coordinate; TMA out-of-bounds zero fill implements the padding.
#include "common.h"
#include "tc.cuh"

namespace mos {

constexpr int BM = 128;
constexpr int BN = 160;
constexpr int BK = 64;
constexpr int LORA_N = 16;
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BM * BK * 2;               // 16384
constexpr int B_STAGE_BYTES = BN * BK * 2;               // 20480
constexpr int L_STAGE_BYTES = LORA_N * BK * 2;           // 2048
constexpr int TMEM_COLS = 256;
constexpr int MAX_DYN_SMEM = 227 * 1024 - 2048;  // leave room for the static barriers

struct GemmDev {
  int M, N;
  int kb_total;        // number of 64-wide k blocks over the whole reduction (conv: 9 * C/64)
  int kb_per_split;
  int stages;
  int conv, H, W, B, kc_per_tap, TW, TH, TB, tiles_w, tiles_h;
  int lora;
  int geglu;
  int out_mode;
  int splits;
  float* partial;
  const float* bias;
  const float* bias_batch;
  long long rows_per_batch;
  long long bias_batch_ld;
  const __nv_bfloat16* residual;
  long long ldr;
  const float* lora_up;
  long long lora_seg;
  void* out;
  long long ldc;
  void* seg_ptr[3];
  int seg_kind[3];
  long long seg_rows_pad[3];
  int heads, head_dim, dpad, dv_pad;
  long long tokens_per_batch;
};

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}

__global__ void __launch_bounds__(192, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmL, const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B; dynamic smem base is only guaranteed 16 B aligned.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = B_STAGE_BYTES + (p.lora ? L_STAGE_BYTES : 0);
  const int stage_bytes = A_STAGE_BYTES + b_bytes;

  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m_tile = blockIdx.y;
  const int split = blockIdx.z;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.lora) tma_prefetch_desc(&tmL);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_holder, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  // tile origin
  int m0 = m_tile * BM;
  int cb0 = 0, ch0 = 0, cw0 = 0;
  if (p.conv) {
    int tw_i = m_tile % p.tiles_w;
    int th_i = (m_tile / p.tiles_w) % p.tiles_h;
    int tb_i = m_tile / (p.tiles_w * p.tiles_h);
    cw0 = tw_i * p.TW;
    ch0 = th_i * p.TH;
    cb0 = tb_i * p.TB;
  }

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        uint8_t* sb = sa + A_STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
        if (p.conv) {
          int tap = kb / p.kc_per_tap;
          int kc = kb - tap * p.kc_per_tap;
          int kh = tap / 3, kw = tap - kh * 3;
          tma_load_4d(sa, &tmA, &full_bar[stage], kc * BK, cw0 + kw - 1, ch0 + kh - 1, cb0);
        } else {
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
        }
        tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
        if (p.lora) tma_load_2d(sb + B_STAGE_BYTES, &tmL, &full_bar[stage], kb * BK, 0);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, p.lora ? BN + LORA_N : BN, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        uint8_t* sa = smem + stage * stage_bytes;
        uint64_t adesc = make_desc_sw128(smem_u32(sa));
        uint64_t bdesc = make_desc_sw128(smem_u32(sa + A_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);

    long long m;
    bool valid;
    if (p.conv) {
      int tw = r % p.TW;
      int th = (r / p.TW) % p.TH;
      int tb = r / (p.TW * p.TH);
      int b = cb0 + tb, h = ch0 + th, w = cw0 + tw;
      valid = (b < p.B) && (h < p.H);
      m = ((long long)b * p.H + h) * p.W + w;
    } else {
      m = m0 + r;
      valid = m < p.M;
    }

    if (p.splits > 1) {
      float* dst = p.partial + ((long long)split * p.M + m) * p.N + n0;
#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(trow + c * 16, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<uint4*>(dst + c * 16 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    } else {
      float t[16];
      if (p.lora) {
        uint32_t tv[16];
        tmem_ld16(trow + BN, tv);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = __uint_as_float(tv[j]);
      }
      const float* bb = nullptr;
      if (p.bias_batch) bb = p.bias_batch + (valid ? (m / p.rows_per_batch) : 0) * p.bias_batch_ld;

      if (p.geglu) {
        // tile columns [0,80) = a, [80,160) = gate for the same 80 outputs
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldc + blockIdx.x * (BN / 2);
#pragma unroll 1
        for (int c = 0; c < (BN / 2) / 16; ++c) {
          uint32_t va[16], vg[16];
          tmem_ld16(trow + c * 16, va);
          tmem_ld16(trow + BN / 2 + c * 16, vg);
          tmem_ld_wait();
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            int na = n0 + c * 16 + j, ng = na + BN / 2;
            float a = __uint_as_float(va[j]), g = __uint_as_float(vg[j]);
            if (p.bias) {
              a += __ldg(p.bias + na);
              g += __ldg(p.bias + ng);
            }
            if (p.lora) {
              float4 ua = __ldg(reinterpret_cast<const float4*>(p.lora_up) + na);
              float4 ug = __ldg(reinterpret_cast<const float4*>(p.lora_up) + ng);
              a += t[0] * ua.x + t[1] * ua.y + t[2] * ua.z + t[3] * ua.w;
              g += t[0] * ug.x + t[1] * ug.y + t[2] * ug.z + t[3] * ug.w;
            }
            o[j] = a * gelu_erf(g);
          }
          if (valid) {
            store_bf16x8(orow + c * 16, o);
            store_bf16x8(orow + c * 16 + 8, o + 8);
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
          uint32_t v[16];
          tmem_ld16(trow + c * 16, v);
          tmem_ld_wait();
          const int nc = n0 + c * 16;
          float o[16];
          float tt[4] = {0.f, 0.f, 0.f, 0.f};
          if (p.lora) {
            const int sidx = (int)(nc / p.lora_seg);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              tt[i] = sidx == 0 ? t[i] : sidx == 1 ? t[4 + i] : sidx == 2 ? t[8 + i] : t[12 + i];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(v[j]);
            if (p.bias) a += __ldg(p.bias + nc + j);
            if (bb) a += __ldg(bb + nc + j);
            if (p.lora) {
              float4 u = __ldg(reinterpret_cast<const float4*>(p.lora_up) + nc + j);
              a += tt[0] * u.x + tt[1] * u.y + tt[2] * u.z + tt[3] * u.w;
            }
            o[j] = a;
          }
          if (valid) {
          if (p.residual) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m * p.ldr + nc);
            uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float2 f = unpack_bf16x2(rr[j]);
              o[2 * j] += f.x;
              o[2 * j + 1] += f.y;
            }
          }
          if (p.out_mode == MOS_OUT_BF16) {
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldc + nc;
            store_bf16x8(orow, o);
            store_bf16x8(orow + 8, o + 8);
          } else if (p.out_mode == MOS_OUT_F32) {
            float* orow = reinterpret_cast<float*>(p.out) + m * p.ldc + nc;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(orow + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          } else {  // MOS_OUT_HEADS
            const int seg_len = p.heads * p.head_dim;
            const long long b = m / p.tokens_per_batch;
            const long long tok = m - b * p.tokens_per_batch;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              int n = nc + half * 8;
              int seg = n / seg_len;
              int cc = n - seg * seg_len;
              int head = cc / p.head_dim;
              int j0 = cc - head * p.head_dim;
              __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p.seg_ptr[seg]);
              long long bh = b * p.heads + head;
              if (p.seg_kind[seg] == MOS_SEG_ROWS) {
                store_bf16x8(base + (bh * p.seg_rows_pad[seg] + tok) * p.dpad + j0, o + half * 8);
              } else {
                __nv_bfloat16* d = base + (bh * p.dv_pad + j0) * p.seg_rows_pad[seg] + tok;
#pragma unroll
                for (int e = 0; e < 8; ++e) d[(long long)e * p.seg_rows_pad[seg]] = __float2bfloat16(o[half * 8 + e]);
              }
            }
          }
          }  // valid
          __syncwarp();
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- split-K finalize
__global__ void splitk_finalize_kernel(const float* __restrict__ partial, int splits, long long M, long long N,
                                       const float* __restrict__ bias, const float* __restrict__ bias_batch,
                                       long long rows_per_batch, long long bias_batch_ld,
                                       const __nv_bfloat16* __restrict__ residual,
                                       long long ldr, __nv_bfloat16* __restrict__ out, long long ldc) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 columns
  long long n4 = N / 4;
  if (idx >= M * n4) return;
  long long m = idx / n4;
  long long n = (idx - m * n4) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    float4 v = __ldg(reinterpret_cast<const float4*>(partial + ((long long)s * M + m) * N + n));
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  if (bias) {
    float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  if (bias_batch) {
    float4 b = __ldg(reinterpret_cast<const float4*>(bias_batch + (m / rows_per_batch) * bias_batch_ld + n));
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  if (residual) {
    uint2 r = __ldg(reinterpret_cast<const uint2*>(residual + m * ldr + n));
    float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
    acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
  }
  uint2 o;
  o.x = pack_bf16x2(acc.x, acc.y);
  o.y = pack_bf16x2(acc.z, acc.w);
  *reinterpret_cast<uint2*>(out + m * ldc + n) = o;
}

static bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace mos

using namespace mos;

extern "C" int mos_gemm_bf16(const mos_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(a != nullptr, "mos_gemm_bf16: args is NULL");
  MOS_CHECK_ARG(a->A && a->W, "mos_gemm_bf16: A/W is NULL");
  MOS_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "mos_gemm_bf16: bad shape M=%lld N=%lld K=%lld",
                (long long)a->M, (long long)a->N, (long long)a->K);
  MOS_CHECK_ARG(a->N % BN == 0, "mos_gemm_bf16: N=%lld must be a multiple of %d", (long long)a->N, BN);
  MOS_CHECK_ARG(a->K % BK == 0, "mos_gemm_bf16: K=%lld must be a multiple of %d", (long long)a->K, BK);
  MOS_CHECK_ARG(is_aligned(a->A, 16) && is_aligned(a->W, 16), "mos_gemm_bf16: A/W must be 16-byte aligned");
  const int splits = a->splits > 0 ? a->splits : 1;
  const bool lora = a->lora_down != nullptr;
  if (splits > 1) {
    MOS_CHECK_ARG(a->partial != nullptr, "mos_gemm_bf16: split-K needs a partial workspace");
    MOS_CHECK_ARG(!lora && !a->geglu && a->out_mode == MOS_OUT_BF16,
                  "mos_gemm_bf16: split-K cannot be combined with lora / geglu / head-split output");
  } else {
    MOS_CHECK_ARG(a->out_mode == MOS_OUT_HEADS || a->out != nullptr, "mos_gemm_bf16: out is NULL");
  }
  if (lora) {
    MOS_CHECK_ARG(a->lora_up != nullptr && a->lora_seg > 0 && a->lora_seg % 16 == 0 && !a->conv,
                  "mos_gemm_bf16: bad LoRA arguments");
    MOS_CHECK_ARG(a->N / a->lora_seg <= 4, "mos_gemm_bf16: at most 4 LoRA segments");
  }
  if (a->out_mode == MOS_OUT_HEADS) {
    MOS_CHECK_ARG(a->heads > 0 && a->head_dim % 8 == 0 && a->N % (a->heads * a->head_dim) == 0 &&
                      a->N / (a->heads * a->head_dim) <= 3 && a->tokens_per_batch > 0,
                  "mos_gemm_bf16: bad head-split arguments");
  }
  if (a->geglu) MOS_CHECK_ARG(a->out_mode == MOS_OUT_BF16, "mos_gemm_bf16: geglu needs bf16 row-major output");

  GemmDev p;
  memset(&p, 0, sizeof(p));
  CUtensorMap tmA, tmB, tmL;
  memset(&tmL, 0, sizeof(tmL));
  p.M = (int)a->M;
  p.N = (int)a->N;
  p.conv = a->conv;
  int m_tiles;
  if (a->conv) {
    MOS_CHECK_ARG(a->B > 0 && a->H > 0 && a->Wd > 0 && a->C == a->K, "mos_gemm_bf16: bad conv geometry");
    MOS_CHECK_ARG((int64_t)a->B * a->H * a->Wd == a->M, "mos_gemm_bf16: conv M != B*H*W");
    int TW = 1;
    while (TW * 2 <= 128 && a->Wd % (TW * 2) == 0) TW *= 2;
    // choose TH (power of two) maximising the fraction of valid rows in the 128-row tile
    int best_th = 1;
    double best_eff = -1;
    for (int TH = 1; TH * TW <= 128; TH *= 2) {
      int TB = 128 / (TW * TH);
      double eff = ((double)a->H / (ceil_div(a->H, TH) * TH)) * ((double)a->B / (ceil_div(a->B, TB) * TB));
      if (eff > best_eff + 1e-9) {
        best_eff = eff;
        best_th = TH;
      }
    }
    p.TW = TW;
    p.TH = best_th;
    p.TB = 128 / (TW * best_th);
    p.H = a->H;
    p.W = a->Wd;
    p.B = a->B;
    p.tiles_w = a->Wd / TW;
    p.tiles_h = (int)ceil_div(a->H, p.TH);
    int tiles_b = (int)ceil_div(a->B, p.TB);
    m_tiles = p.tiles_w * p.tiles_h * tiles_b;
    p.kc_per_tap = (int)(a->K / BK);
    p.kb_total = 9 * p.kc_per_tap;
    uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->Wd, (uint64_t)a->H, (uint64_t)a->B};
    const uint64_t pitch = (uint64_t)(a->lda > 0 ? a->lda : a->C);
    MOS_CHECK_ARG(pitch >= (uint64_t)a->C && pitch % 8 == 0, "mos_gemm_bf16: conv pixel pitch %llu invalid",
                  (unsigned long long)pitch);
    uint64_t str[3] = {pitch * 2, (uint64_t)a->Wd * pitch * 2, (uint64_t)a->H * a->Wd * pitch * 2};
    uint32_t box[4] = {BK, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TB};
    int rc = encode_tmap(&tmA, a->A, 2, 4, dims, str, box, 3);
    if (rc) return rc;
    uint64_t wd[2] = {(uint64_t)a->K * 9, (uint64_t)a->N};
    uint64_t ws[1] = {(uint64_t)a->K * 9 * 2};
    uint32_t wb[2] = {BK, BN};
    rc = encode_tmap(&tmB, a->W, 2, 2, wd, ws, wb, 3);
    if (rc) return rc;
  } else {
    MOS_CHECK_ARG(a->lda >= a->K && a->lda % 8 == 0, "mos_gemm_bf16: lda=%lld invalid", (long long)a->lda);
    m_tiles = (int)ceil_div(a->M, BM);
    p.kb_total = (int)(a->K / BK);
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->lda * 2};
    uint32_t box[2] = {BK, BM};
    int rc = encode_tmap(&tmA, a->A, 2, 2, dims, str, box, 3);
    if (rc) return rc;
    uint64_t wd[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t ws[1] = {(uint64_t)a->K * 2};
    uint32_t wb[2] = {BK, BN};
    rc = encode_tmap(&tmB, a->W, 2, 2, wd, ws, wb, 3);
    if (rc) return rc;
    if (lora) {
      uint64_t ld[2] = {(uint64_t)a->K, LORA_N};
      uint64_t ls[1] = {(uint64_t)a->K * 2};
      uint32_t lb[2] = {BK, LORA_N};
      rc = encode_tmap(&tmL, a->lora_down, 2, 2, ld, ls, lb, 3);
      if (rc) return rc;
    }
  }
  MOS_CHECK_ARG(splits <= p.kb_total, "mos_gemm_bf16: splits=%d > k blocks=%d", splits, p.kb_total);
  p.splits = splits;
  p.kb_per_split = (int)ceil_div(p.kb_total, splits);
  MOS_CHECK_ARG((long long)p.kb_per_split * (splits - 1) < p.kb_total, "mos_gemm_bf16: empty split");
  p.lora = lora ? 1 : 0;
  p.geglu = a->geglu;
  p.out_mode = a->out_mode;
  p.partial = a->partial;
  p.bias = a->bias;
  p.bias_batch = a->bias_batch;
  p.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : 1;
  p.bias_batch_ld = a->bias_batch_ld > 0 ? a->bias_batch_ld : a->N;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.ldr = a->ldr;
  p.lora_up = a->lora_up;
  p.lora_seg = a->lora_seg > 0 ? a->lora_seg : a->N;
  p.out = a->out;
  p.ldc = a->ldc;
  for (int i = 0; i < 3; ++i) {
    p.seg_ptr[i] = a->seg_ptr[i];
    p.seg_kind[i] = a->seg_kind[i];
    p.seg_rows_pad[i] = a->seg_rows_pad[i];
  }
  p.heads = a->heads;
  p.head_dim = a->head_dim;
  p.dpad = a->dpad;
  p.dv_pad = a->dv_pad;
  p.tokens_per_batch = a->tokens_per_batch > 0 ? a->tokens_per_batch : 1;

  const int stage_bytes = A_STAGE_BYTES + B_STAGE_BYTES + (lora ? L_STAGE_BYTES : 0);
  int stages = a->stages > 0 ? a->stages : 5;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  while (stages * stage_bytes + 1024 > MAX_DYN_SMEM) --stages;
  if (stages > p.kb_per_split) stages = p.kb_per_split < 2 ? 2 : p.kb_per_split;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + 1024;

  static bool configured = false;
  if (!configured) {
    MOS_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM));
    configured = true;
  }
  dim3 grid((unsigned)(a->N / BN), (unsigned)m_tiles, (unsigned)splits);
  gemm_kernel<<<grid, 192, smem_bytes, stream>>>(tmA, tmB, tmL, p);
  MOS_CHECK_LAUNCH();
  return MOS_OK;
}

extern "C" int mos_splitk_finalize(const float* partial, int32_t splits, int64_t M, int64_t N, const float* bias,
                                   const float* bias_batch, int64_t rows_per_batch, int64_t bias_batch_ld,
                                   const void* residual, int64_t ldr, void* out, int64_t ldc, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(partial && out && splits >= 1 && M > 0 && N > 0 && N % 4 == 0, "mos_splitk_finalize: bad arguments");
  long long total = M * (N / 4);
  int threads = 256;
  long long blocks = ceil_div(total, threads);
  splitk_finalize_kernel<<<(unsigned)blocks, threads, 0, stream>>>(
      partial, splits, M, N, bias, bias_batch, rows_per_batch > 0 ? rows_per_batch : 1,
      bias_batch_ld > 0 ? bias_batch_ld : N, reinterpret_cast<const __nv_bfloat16*>(residual), ldr, reinterpret_cast<__nv_bfloat16*>(out), ldc);
  MOS_CHECK_LAUNCH();
  return MOS_OK;
}
