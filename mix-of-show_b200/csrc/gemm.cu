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
constexpr int TMEM_COLS = 512;            // two accumulator stages of 256 columns
constexpr int ACC_STRIDE = 256;
constexpr int STG_PITCH = BN * 2 + 16;    // padded row pitch of the epilogue staging tile (bank-conflict free)
constexpr int STG_BYTES = BM * STG_PITCH;  // 43008
constexpr int EPI_SMEM_BYTES = STG_BYTES + 4 * BN * 4 + BN * 16;
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
  int n_tiles, total_items, nbatch;
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

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // epilogue warps only
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

struct TileCoord {
  int n0, m0, cb0, ch0, cw0, split;
};
__device__ __forceinline__ TileCoord item_coord(const GemmDev& p, int w) {
  TileCoord t;
  t.split = w % p.splits;
  const int tt = w / p.splits;
  const int tn = tt % p.n_tiles;
  const int tm = tt / p.n_tiles;
  t.n0 = tn * BN;
  t.m0 = tm * BM;
  t.cb0 = t.ch0 = t.cw0 = 0;
  if (p.conv) {
    t.cw0 = (tm % p.tiles_w) * p.TW;
    t.ch0 = ((tm / p.tiles_w) % p.tiles_h) * p.TH;
    t.cb0 = (tm / (p.tiles_w * p.tiles_h)) * p.TB;
  }
  return t;
}
// row r of a tile -> global output row m (and validity)
__device__ __forceinline__ bool row_coord(const GemmDev& p, const TileCoord& t, int r, long long& m, int& b) {
  if (p.conv) {
    const int tw = r % p.TW, th = (r / p.TW) % p.TH, tb = r / (p.TW * p.TH);
    b = t.cb0 + tb;
    const int h = t.ch0 + th, w = t.cw0 + tw;
    m = ((long long)b * p.H + h) * p.W + w;
    return (b < p.B) && (h < p.H);
  }
  m = (long long)t.m0 + r;
  b = (int)(m / p.rows_per_batch);
  return m < p.M;
}

__global__ void __launch_bounds__(192, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmL, const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B; dynamic smem base is only guaranteed 16 B aligned.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = B_STAGE_BYTES + (p.lora ? L_STAGE_BYTES : 0);
  const int stage_bytes = A_STAGE_BYTES + b_bytes;
  uint8_t* stg = smem + p.stages * stage_bytes;                    // epilogue staging tile [128][STG_PITCH]
  float* cb_s = reinterpret_cast<float*>(stg + STG_BYTES);          // [4][BN] bias (+ per-batch bias)
  float4* up_s = reinterpret_cast<float4*>(cb_s + 4 * BN);          // [BN] LoRA up rows (pre-scaled by alpha)

  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.lora) tma_prefetch_desc(&tmL);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_holder, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  // Everything above touched no global memory written by the previous kernel in the stream.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x) {
        const TileCoord t = item_coord(p, w);
        const int kb_begin = t.split * p.kb_per_split;
        const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          if (p.conv) {
            const int tap = kb / p.kc_per_tap;
            const int kc = kb - tap * p.kc_per_tap;
            const int kh = tap / 3, kw = tap - kh * 3;
            tma_load_4d(sa, &tmA, &full_bar[stage], kc * BK, t.cw0 + kw - 1, t.ch0 + kh - 1, t.cb0);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, t.m0);
          }
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, t.n0);
          if (p.lora) tma_load_2d(sb + B_STAGE_BYTES, &tmL, &full_bar[stage], kb * BK, 0);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, p.lora ? BN + LORA_N : BN, 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++it) {
        const TileCoord t = item_coord(p, w);
        const int kb_begin = t.split * p.kb_per_split;
        const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          uint8_t* sa = smem + stage * stage_bytes;
          const uint64_t adesc = make_desc_sw128(smem_u32(sa));
          const uint64_t bdesc = make_desc_sw128(smem_u32(sa + A_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;        // tile row owned by this thread
    const int et = threadIdx.x - 64;    // 0..127
    const bool staged = (p.splits == 1) && (p.out_mode == MOS_OUT_BF16);
    const int out_cols = p.geglu ? BN / 2 : BN;
    const int chunks_per_row = out_cols / 8;   // 16-byte chunks
    int it = 0;
    for (int w = blockIdx.x; w < p.total_items; w += gridDim.x, ++it) {
      const TileCoord t = item_coord(p, w);
      const int acc = it & 1;
      long long m;
      int b;
      const bool valid = row_coord(p, t, r, m, b);
      int b_lo;
      {
        long long m_first;
        row_coord(p, t, 0, m_first, b_lo);
      }
      // ---- 1. stage bias (+ per-batch bias) and LoRA-up rows of this tile's columns
      if (p.splits == 1) {
        for (int i = et; i < 4 * BN; i += 128) {
          const int j = i / BN, n = i - j * BN;
          float v = p.bias ? __ldg(p.bias + t.n0 + n) : 0.f;
          if (p.bias_batch && (b_lo + j) < p.nbatch) v += __ldg(p.bias_batch + (long long)(b_lo + j) * p.bias_batch_ld + t.n0 + n);
          cb_s[i] = v;
        }
        if (p.lora)
          for (int i = et; i < BN; i += 128) up_s[i] = __ldg(reinterpret_cast<const float4*>(p.lora_up) + t.n0 + i);
      }
      // ---- 2. residual tile -> staging (coalesced 16-byte cp.async), hidden behind the mainloop
      if (staged && p.residual) {
        for (int i = et; i < BM * chunks_per_row; i += 128) {
          const int rr = i / chunks_per_row, ch = i - rr * chunks_per_row;
          long long mm;
          int bb_;
          if (row_coord(p, t, rr, mm, bb_))
            cp_async16(stg + rr * STG_PITCH + ch * 16, p.residual + mm * p.ldr + (p.geglu ? t.n0 / 2 : t.n0) + ch * 8);
        }
        cp_async_wait_all();
      }
      epi_bar();
      // ---- 3. accumulators ready?
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + acc * ACC_STRIDE + (uint32_t(q * 32) << 16);
      const int bsel = min(max(b - b_lo, 0), 3);
      const float* cb = cb_s + bsel * BN;
      uint8_t* srow = stg + r * STG_PITCH;

      if (p.splits > 1) {
        float* dst = p.partial + ((long long)t.split * p.M + m) * p.N + t.n0;
#pragma unroll 1
        for (int c = 0; c < BN / 16; c += 2) {
          uint32_t v0[16], v1[16];
          tmem_ld16(trow + c * 16, v0);
          tmem_ld16(trow + c * 16 + 16, v1);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              *reinterpret_cast<uint4*>(dst + c * 16 + j) = make_uint4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
              *reinterpret_cast<uint4*>(dst + c * 16 + 16 + j) = make_uint4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
            }
          }
        }
      } else {
        float t4[16];
        if (p.lora) {
          uint32_t tv[16];
          tmem_ld16(trow + BN, tv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) t4[j] = __uint_as_float(tv[j]);
        }
        if (p.geglu) {
          // tile columns [0,80) = a, [80,160) = gate for the same 80 outputs
#pragma unroll 1
          for (int c = 0; c < (BN / 2) / 16; ++c) {
            uint32_t va[16], vg[16];
            tmem_ld16(trow + c * 16, va);
            tmem_ld16(trow + BN / 2 + c * 16, vg);
            tmem_ld_wait();
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int na = c * 16 + j, ng = na + BN / 2;
              float a = __uint_as_float(va[j]) + cb[na];
              float g = __uint_as_float(vg[j]) + cb[ng];
              if (p.lora) {
                const float4 ua = up_s[na], ug = up_s[ng];
                a += t4[0] * ua.x + t4[1] * ua.y + t4[2] * ua.z + t4[3] * ua.w;
                g += t4[0] * ug.x + t4[1] * ug.y + t4[2] * ug.z + t4[3] * ug.w;
              }
              o[j] = a * gelu_erf(g);
            }
            store_bf16x8(reinterpret_cast<__nv_bfloat16*>(srow + c * 32), o);
            store_bf16x8(reinterpret_cast<__nv_bfloat16*>(srow + c * 32 + 16), o + 8);
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < BN / 16; c += 2) {
            uint32_t vv[2][16];
            tmem_ld16(trow + c * 16, vv[0]);
            tmem_ld16(trow + c * 16 + 16, vv[1]);
            tmem_ld_wait();
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              const int nl = (c + hlf) * 16;       // column inside the tile
              const int nc = t.n0 + nl;            // global column
              float o[16];
              float tt[4] = {0.f, 0.f, 0.f, 0.f};
              if (p.lora) {
                const int sidx = (int)(nc / p.lora_seg);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  tt[i] = sidx == 0 ? t4[i] : sidx == 1 ? t4[4 + i] : sidx == 2 ? t4[8 + i] : t4[12 + i];
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float a = __uint_as_float(vv[hlf][j]) + cb[nl + j];
                if (p.lora) {
                  const float4 u = up_s[nl + j];
                  a += tt[0] * u.x + tt[1] * u.y + tt[2] * u.z + tt[3] * u.w;
                }
                o[j] = a;
              }
              if (p.out_mode == MOS_OUT_BF16) {
                if (p.residual) {
                  const uint4 r0 = *reinterpret_cast<const uint4*>(srow + nl * 2);
                  const uint4 r1 = *reinterpret_cast<const uint4*>(srow + nl * 2 + 16);
                  const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float2 f = unpack_bf16x2(rr[j]);
                    o[2 * j] += f.x;
                    o[2 * j + 1] += f.y;
                  }
                }
                store_bf16x8(reinterpret_cast<__nv_bfloat16*>(srow + nl * 2), o);
                store_bf16x8(reinterpret_cast<__nv_bfloat16*>(srow + nl * 2 + 16), o + 8);
              } else if (p.out_mode == MOS_OUT_F32) {
                if (valid) {
                  float* orow = reinterpret_cast<float*>(p.out) + m * p.ldc + nc;
#pragma unroll
                  for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(orow + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                }
              } else if (valid) {  // MOS_OUT_HEADS
                const int seg_len = p.heads * p.head_dim;
                const long long bb = m / p.tokens_per_batch;
                const long long tok = m - bb * p.tokens_per_batch;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  const int n = nc + half * 8;
                  const int seg = n / seg_len;
                  const int cc = n - seg * seg_len;
                  const int head = cc / p.head_dim;
                  const int j0 = cc - head * p.head_dim;
                  __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p.seg_ptr[seg]);
                  const long long bh = bb * p.heads + head;
                  if (p.seg_kind[seg] == MOS_SEG_ROWS) {
                    store_bf16x8(base + (bh * p.seg_rows_pad[seg] + tok) * p.dpad + j0, o + half * 8);
                  } else {
                    __nv_bfloat16* d = base + (bh * p.dv_pad + j0) * p.seg_rows_pad[seg] + tok;
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                      d[(long long)e * p.seg_rows_pad[seg]] = __float2bfloat16(o[half * 8 + e]);
                  }
                }
              }
            }
            __syncwarp();
          }
        }
      }
      // ---- 4. accumulator drained: hand it back to the MMA warp (next-but-one tile)
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      // ---- 5. coalesced write-out of the staged bf16 tile
      if (staged) {
        epi_bar();
        __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out);
        const int ocol0 = p.geglu ? (t.n0 / 2) : t.n0;
        for (int i = et; i < BM * chunks_per_row; i += 128) {
          const int rr = i / chunks_per_row, ch = i - rr * chunks_per_row;
          long long mm;
          int bb_;
          if (row_coord(p, t, rr, mm, bb_)) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * STG_PITCH + ch * 16);
            *reinterpret_cast<uint4*>(obase + mm * p.ldc + ocol0 + ch * 8) = v;
          }
        }
      }
      epi_bar();   // staging / bias tables are reused by the next item
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
  pdl_wait();
  pdl_launch_dependents();
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
      if (TB > 4 && TH * 2 * TW <= 128) continue;   // the epilogue stages per-batch bias for <= 4 batches per tile
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

  if (a->bias_batch && !a->conv)
    MOS_CHECK_ARG(p.rows_per_batch >= 32, "mos_gemm_bf16: bias_batch needs rows_per_batch >= 32 in plain mode");
  p.n_tiles = (int)(a->N / BN);
  p.total_items = p.n_tiles * m_tiles * splits;
  p.nbatch = a->conv ? a->B : (int)ceil_div(a->M, p.rows_per_batch);

  const int stage_bytes = A_STAGE_BYTES + B_STAGE_BYTES + (lora ? L_STAGE_BYTES : 0);
  int stages = a->stages > 0 ? a->stages : MAX_STAGES;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  while (stages * stage_bytes + EPI_SMEM_BYTES + 1024 > MAX_DYN_SMEM) --stages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + EPI_SMEM_BYTES + 1024;

  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    MOS_CHECK_CUDA(cudaGetDevice(&dev));
    MOS_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    MOS_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM));
  }
  // persistent: one CTA per SM, each loops over its share of (tile, split) work items
  const int grid = p.total_items < num_sms ? p.total_items : num_sms;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MOS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel, tmA, tmB, tmL, p));
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
  MOS_CHECK_CUDA(launch_pdl(splitk_finalize_kernel, dim3((unsigned)blocks), dim3(threads), 0, stream, partial,
                            (int)splits, (long long)M, (long long)N, bias, bias_batch,
                            (long long)(rows_per_batch > 0 ? rows_per_batch : 1),
                            (long long)(bias_batch_ld > 0 ? bias_batch_ld : N),
                            reinterpret_cast<const __nv_bfloat16*>(residual), (long long)ldr,
                            reinterpret_cast<__nv_bfloat16*>(out), (long long)ldc));
  return MOS_OK;
}
