// gemm.cu — K1/K4: warp-specialised tcgen05 GEMM and implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[m,n] = epilogue( sum_k A[m,k] W[n,k] )        A, W bf16 K-major; fp32 accumulation in TMEM
//
// Persistent kernel: one CTA per SM loops over 128 x 160 output tiles (160 divides every SD1.5 channel count).
// CTAs are grouped in thread-block clusters of CX x CM (CX = 2 along N, CM = 1/2/4 along M): the A tile is shared by
// the CX column neighbours and the W tile by the CM row neighbours, so every CTA fetches only 1/CX of A and 1/CM of W
// from L2 and TMA-multicasts it to its peers (the mainloop is L2-bandwidth bound without this: 73 FLOP/B per tile).
// Roles (192 threads):
//   warp 0      TMA producer   cp.async.bulk.tensor(.multicast) -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x (160|176) x 16), 2 accumulator stages
//   warps 2..5  epilogue       tcgen05.ld (32x32b) -> registers -> fused epilogue -> smem staging -> coalesced stores;
//                              runs concurrently with the next tile's mainloop
// LoRA fusion (edlora.py:244-246): the rank-padded down matrix [16, K] rides along as 16 extra B rows, so the
// same MMA also produces t = x * down^T in TMEM columns 160..175; the epilogue adds t * (alpha*up)^T.
// Convolution: the A tile is a TW x TH x TB pixel patch of the NHWC activation fetched by a 4-D tensor map at
// the tap-shifted coordinate; TMA out-of-bounds zero fill implements the padding.
#include <stdlib.h>

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
constexpr int EPI_THREADS = 256;           // 8 epilogue warps: 2 per TMEM lane quadrant, each half of the tile columns
constexpr int NUM_THREADS = 64 + EPI_THREADS;
constexpr int MAX_DYN_SMEM = 227 * 1024 - 2048;  // leave room for the static barriers

struct GemmDev {
  int M, N;
  int kb_total;        // number of 64-wide k blocks over the whole reduction (conv: 9 * C/64)
  int kb_per_split;
  int stages;
  int conv, H, W, B, kc_per_tap, TW, TH, TB, lgTW, lgTH, tiles_w, tiles_h;
  int half_dim;        // conv, CX == 2: which box dimension (1 = W, 2 = H, 3 = B) is split between the A halves
  int half_off;        // coordinate offset of the second half along that dimension
  int lora;
  int geglu;
  int out_mode;
  int splits;
  int n_tiles, m_tiles, total_super, nbatch;
  int cx, cm;          // cluster shape (N x M)
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
  int accum;           // MOS_OUT_F32: out += result (Gram accumulation)
  unsigned long long* tl;   // optional timeline buffer (mos_debug_set_timeline)
  int w_static;        // W tiles may be requested before griddepcontrol.wait
  int w_f16;           // W (and the LoRA rows) are fp16 instead of bf16 (Gram products of fp16 activations)
};

template <bool F16>
__device__ __forceinline__ void store16x8(__nv_bfloat16* dst, const float* v) {
  uint4 u;
  u.x = pack16x2<F16>(v[0], v[1]);
  u.y = pack16x2<F16>(v[2], v[3]);
  u.z = pack16x2<F16>(v[4], v[5]);
  u.w = pack16x2<F16>(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}

__device__ __forceinline__ void epi_bar() {  // epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x - 64) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---- optional in-kernel timeline (profiling aid): when a buffer is registered through mos_debug_set_timeline, the
// first 8 CTAs of every gemm launch record %globaltimer stamps (ns) at their phase boundaries.  The pointer travels in
// the kernel parameters (constant bank): a __device__ global would cost an L2 round trip at every stamp site, on the
// critical path of the TMA / MMA threads.
#define stamp(slot)                                                       \
  do {                                                                    \
    if (p.tl != nullptr && blockIdx.x < 8) {                              \
      unsigned long long t_;                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));              \
      p.tl[blockIdx.x * 8 + (slot)] = t_;                                 \
    }                                                                     \
  } while (0)

struct TileCoord {
  int n0, m0, cb0, ch0, cw0, split;
};
// super-item ws (shared by the whole cluster) + cluster rank -> this CTA's tile
__device__ __forceinline__ TileCoord item_coord(const GemmDev& p, int ws, int nx, int my) {
  TileCoord t;
  t.split = ws % p.splits;
  const int tt = ws / p.splits;
  const int sn = p.n_tiles / p.cx;
  const int tn = (tt % sn) * p.cx + nx;
  const int tm = (tt / sn) * p.cm + my;
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
    const int tw = r & (p.TW - 1), th = (r >> p.lgTW) & (p.TH - 1), tb = r >> (p.lgTW + p.lgTH);
    b = t.cb0 + tb;
    const int h = t.ch0 + th, w = t.cw0 + tw;
    m = ((long long)b * p.H + h) * p.W + w;
    return (b < p.B) && (h < p.H);
  }
  m = (long long)t.m0 + r;
  b = (int)(m / p.rows_per_batch);
  return m < p.M;
}

// same without the batch index (no 64-bit division on the copy loops)
__device__ __forceinline__ bool row_m(const GemmDev& p, const TileCoord& t, int r, long long& m) {
  if (p.conv) {
    const int tw = r & (p.TW - 1), th = (r >> p.lgTW) & (p.TH - 1), tb = r >> (p.lgTW + p.lgTH);
    const int b = t.cb0 + tb, h = t.ch0 + th, w = t.cw0 + tw;
    m = ((long long)b * p.H + h) * p.W + w;
    return (b < p.B) && (h < p.H);
  }
  m = (long long)t.m0 + r;
  return m < p.M;
}

// coalesced copy between the padded staging tile and global rows; CPR = 16-byte chunks per row (20 or 10)
template <int CPR, bool TO_GLOBAL>
__device__ __forceinline__ void stage_copy(const GemmDev& p, const TileCoord& t, uint8_t* stg, __nv_bfloat16* gbase,
                                           long long ld, int col0, int et) {
#pragma unroll 4
  const int nthr = (int)blockDim.x - 64;
  for (int i = et; i < BM * CPR; i += nthr) {
    const int rr = i / CPR, ch = i - rr * CPR;
    long long mm;
    if (row_m(p, t, rr, mm)) {
      uint8_t* s = stg + rr * STG_PITCH + ch * 16;
      __nv_bfloat16* g = gbase + mm * ld + col0 + ch * 8;
      if (TO_GLOBAL)
        *reinterpret_cast<uint4*>(g) = *reinterpret_cast<const uint4*>(s);
      else
        cp_async16(s, g);
    }
  }
}

template <bool F16>   // 16-bit type of A, of the row / head-split outputs and of the residual: fp16 or bf16
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmL, const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B; the dynamic smem base offset is identical in every CTA of the
  // cluster (same kernel, same static smem), which the multicast addressing relies on.
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
  const int csize = p.cx * p.cm;
  const int crank = csize > 1 ? (int)cluster_ctarank() : 0;
  const int nx = crank % p.cx, my = crank / p.cx;
  const int cluster_id = blockIdx.x / csize;
  const int num_clusters = gridDim.x / csize;
  if (threadIdx.x == 0) stamp(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.lora) tma_prefetch_desc(&tmL);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], csize);   // one tcgen05.commit arrival from every CTA of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], blockDim.x - 64);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_holder, TMEM_COLS);
  tc_fence_before();
  if (csize > 1)
    cluster_sync_all();   // peers must see initialised barriers before any remote arrive / multicast lands
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  // Everything above touched no global memory written by the previous kernel in the stream.
  if (threadIdx.x == 0) stamp(1);
  const bool is_producer = warp == 0 && lane == 0;
  if (!is_producer) pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      // multicast masks: A goes to the CX column neighbours (same my), W to the CM row neighbours (same nx)
      uint16_t mask_a = 0, mask_b = 0;
      for (int j = 0; j < p.cx; ++j) mask_a |= (uint16_t)(1u << (my * p.cx + j));
      for (int j = 0; j < p.cm; ++j) mask_b |= (uint16_t)(1u << (j * p.cx + nx));
      const int a_rows = BM / p.cx;           // rows of A this CTA fetches
      const int b_rows = BN / p.cm;           // rows of W this CTA fetches
      // The weights are static: request the W tiles of the first ring pass BEFORE waiting for the previous kernel, so
      // that their HBM round trip overlaps the predecessor's tail (the 1.72 GB of weights stream from HBM every step;
      // the activations behind the dependency come from L2).  A (and the LoRA rows, which the training step rewrites)
      // follow after griddepcontrol.wait and complete the same mbarrier transaction.
      int pre = 0;
      if (p.w_static && csize == 1 && cluster_id < p.total_super) {
        const TileCoord t0 = item_coord(p, cluster_id, nx, my);
        const int kb0 = t0.split * p.kb_per_split;
        pre = min(p.stages, min(p.kb_total, kb0 + p.kb_per_split) - kb0);
        for (int i = 0; i < pre; ++i) {
          mbar_expect_tx(&full_bar[i], (uint32_t)stage_bytes);
          tma_load_2d(smem + i * stage_bytes + A_STAGE_BYTES, &tmB, &full_bar[i], (kb0 + i) * BK, t0.n0);
        }
      }
      pdl_wait();
      stamp(2);
      int stage = 0;
      uint32_t phase = 0;
      int issued = 0;
      for (int ws = cluster_id; ws < p.total_super; ws += num_clusters) {
        const TileCoord t = item_coord(p, ws, nx, my);
        const int kb_begin = t.split * p.kb_per_split;
        const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++issued) {
          const bool w_requested = issued < pre;     // W tile already in flight, transaction bytes already expected
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (!w_requested) {
            mbar_wait(&empty_bar[stage], phase ^ 1);   // every CTA of the cluster has released this slot
            mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          }
          uint8_t* sa_dst = sa + nx * a_rows * 128;
          if (p.conv) {
            const int tap = kb / p.kc_per_tap;
            const int kc = kb - tap * p.kc_per_tap;
            const int kh = tap / 3, kw = tap - kh * 3;
            int cw = t.cw0 + kw - 1, chh = t.ch0 + kh - 1, cb = t.cb0;
            if (nx == 1) {
              if (p.half_dim == 1) cw += p.half_off;
              else if (p.half_dim == 2) chh += p.half_off;
              else cb += p.half_off;
            }
            if (p.cx > 1) tma_load_4d_mc(sa_dst, &tmA, &full_bar[stage], kc * BK, cw, chh, cb, mask_a);
            else tma_load_4d(sa_dst, &tmA, &full_bar[stage], kc * BK, cw, chh, cb);
          } else {
            if (p.cx > 1) tma_load_2d_mc(sa_dst, &tmA, &full_bar[stage], kb * BK, t.m0 + nx * a_rows, mask_a);
            else tma_load_2d(sa_dst, &tmA, &full_bar[stage], kb * BK, t.m0);
          }
          if (p.cm > 1)
            tma_load_2d_mc(sb + my * b_rows * 128, &tmB, &full_bar[stage], kb * BK, t.n0 + my * b_rows, mask_b);
          else if (!w_requested)
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
      const uint32_t idesc = make_idesc_ab(BM, p.lora ? BN + LORA_N : BN, F16 ? 0 : 1, p.w_f16 ? 0 : 1);
      const uint16_t mask_all = (uint16_t)((1u << csize) - 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int ws = cluster_id; ws < p.total_super; ws += num_clusters, ++it) {
        const TileCoord t = item_coord(p, ws, nx, my);
        const int kb_begin = t.split * p.kb_per_split;
        const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (it == 0 && kb == kb_begin) stamp(3);
          uint8_t* sa = smem + stage * stage_bytes;
          const uint64_t adesc = make_desc_sw128(smem_u32(sa));
          const uint64_t bdesc = make_desc_sw128(smem_u32(sa + A_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in every CTA of the cluster) once these MMAs retire
          if (csize > 1) umma_commit_mc(&empty_bar[stage], mask_all);
          else umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int chalf0 = (warp - 2) >> 2;      // first column half this warp handles
    const int chalf_step = ((int)blockDim.x - 64) >> 7;   // 1 (4 epilogue warps: both halves) or 2 (8 warps)
    const int r = q * 32 + lane;             // tile row owned by this thread
    const int et = threadIdx.x - 64;         // 0..255
    const bool staged = (p.splits == 1) && (p.out_mode == MOS_OUT_BF16);
    int it = 0;
    for (int ws = cluster_id; ws < p.total_super; ws += num_clusters, ++it) {
      const TileCoord t = item_coord(p, ws, nx, my);
      const int acc = it & 1;
      long long m;
      int b;
      const bool valid = row_coord(p, t, r, m, b);
      int b_lo;
      {
        long long m_first;
        row_coord(p, t, 0, m_first, b_lo);
      }
      // ---- 1. residual tile -> staging (coalesced 16-byte cp.async, in flight while the bias tables are staged)
      if (staged && p.residual) {
        __nv_bfloat16* rbase = const_cast<__nv_bfloat16*>(p.residual);
        if (p.geglu) stage_copy<BN / 16, false>(p, t, stg, rbase, p.ldr, t.n0 / 2, et);
        else stage_copy<BN / 8, false>(p, t, stg, rbase, p.ldr, t.n0, et);
      }
      // ---- 2. bias (+ per-batch bias) and LoRA-up rows of this tile's columns
      if (p.splits == 1) {
        for (int i = et; i < 4 * BN; i += (int)blockDim.x - 64) {
          const int j = i / BN, n = i - j * BN;
          float v = p.bias ? __ldg(p.bias + t.n0 + n) : 0.f;
          if (p.bias_batch && (b_lo + j) < p.nbatch)
            v += __ldg(p.bias_batch + (long long)(b_lo + j) * p.bias_batch_ld + t.n0 + n);
          cb_s[i] = v;
        }
        if (p.lora)
          for (int i = et; i < BN; i += (int)blockDim.x - 64)
            up_s[i] = __ldg(reinterpret_cast<const float4*>(p.lora_up) + t.n0 + i);
      }
      cp_async_wait_all();
      epi_bar();
      if (et == 0 && it == 0) stamp(4);
      // ---- 3. accumulators ready?  (one lane per warp polls: 256 spinning threads would steal issue slots from the
      //         single-thread TMA / MMA roles)
      if (lane == 0) mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      if (et == 0 && it == 0) stamp(5);
      const uint32_t trow = tmem_base + acc * ACC_STRIDE + (uint32_t(q * 32) << 16);
      const int bsel = min(max(b - b_lo, 0), 3);
      const float* cb = cb_s + bsel * BN;
      uint8_t* srow = stg + r * STG_PITCH;

      for (int chalf = chalf0; chalf < 2; chalf += chalf_step) {
      if (p.splits > 1) {
        float* dst = p.partial + ((long long)t.split * p.M + m) * p.N + t.n0;
#pragma unroll 1
        for (int c = chalf * 5; c < chalf * 5 + 5; ++c) {
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
        float t4[16];
        if (p.lora) {
          uint32_t tv[16];
          tmem_ld16(trow + BN, tv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) t4[j] = __uint_as_float(tv[j]);
        }
        if (p.geglu) {
          // tile columns [0,80) = a, [80,160) = gate for the same 80 outputs; this warp: outputs [40*chalf, +40)
#pragma unroll 1
          for (int c = chalf * 5; c < chalf * 5 + 5; ++c) {
            uint32_t va[8], vg[8];
            tmem_ld8(trow + c * 8, va);
            tmem_ld8(trow + BN / 2 + c * 8, vg);
            tmem_ld_wait();
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int na = c * 8 + j, ng = na + BN / 2;
              float a = __uint_as_float(va[j]) + cb[na];
              float g = __uint_as_float(vg[j]) + cb[ng];
              if (p.lora) {
                const float4 ua = up_s[na], ug = up_s[ng];
                a += t4[0] * ua.x + t4[1] * ua.y + t4[2] * ua.z + t4[3] * ua.w;
                g += t4[0] * ug.x + t4[1] * ug.y + t4[2] * ug.z + t4[3] * ug.w;
              }
              o[j] = a * gelu_erf(g);
            }
            store16x8<F16>(reinterpret_cast<__nv_bfloat16*>(srow + c * 16), o);
          }
        } else {
#pragma unroll 1
          for (int c = chalf * 5; c < chalf * 5 + 5; ++c) {
            uint32_t vv[16];
            tmem_ld16(trow + c * 16, vv);
            tmem_ld_wait();
            const int nl = c * 16;               // column inside the tile
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
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const float4 cbv = *reinterpret_cast<const float4*>(cb + nl + j4);
              o[j4 + 0] = __uint_as_float(vv[j4 + 0]) + cbv.x;
              o[j4 + 1] = __uint_as_float(vv[j4 + 1]) + cbv.y;
              o[j4 + 2] = __uint_as_float(vv[j4 + 2]) + cbv.z;
              o[j4 + 3] = __uint_as_float(vv[j4 + 3]) + cbv.w;
            }
            if (p.lora) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 u = up_s[nl + j];
                o[j] += tt[0] * u.x + tt[1] * u.y + tt[2] * u.z + tt[3] * u.w;
              }
            }
            if (p.out_mode == MOS_OUT_BF16) {
              if (p.residual) {
                const uint4 r0 = *reinterpret_cast<const uint4*>(srow + nl * 2);
                const uint4 r1 = *reinterpret_cast<const uint4*>(srow + nl * 2 + 16);
                const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 f = unpack16x2<F16>(rr[j]);
                  o[2 * j] += f.x;
                  o[2 * j + 1] += f.y;
                }
              }
              store16x8<F16>(reinterpret_cast<__nv_bfloat16*>(srow + nl * 2), o);
              store16x8<F16>(reinterpret_cast<__nv_bfloat16*>(srow + nl * 2 + 16), o + 8);
            } else if (p.out_mode == MOS_OUT_F32) {
              if (valid) {
                float* orow = reinterpret_cast<float*>(p.out) + m * p.ldc + nc;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  float4 r4 = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                  if (p.accum) {
                    const float4 old = *reinterpret_cast<const float4*>(orow + j);
                    r4.x += old.x; r4.y += old.y; r4.z += old.z; r4.w += old.w;
                  }
                  *reinterpret_cast<float4*>(orow + j) = r4;
                }
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
                  store16x8<F16>(base + (bh * p.seg_rows_pad[seg] + tok) * p.dpad + j0, o + half * 8);
                } else {
                  __nv_bfloat16* d = base + (bh * p.dv_pad + j0) * p.seg_rows_pad[seg] + tok;
#pragma unroll
                  for (int e = 0; e < 8; ++e)
                    reinterpret_cast<uint16_t*>(d)[(long long)e * p.seg_rows_pad[seg]] = cvt16<F16>(o[half * 8 + e]);
                }
              }
            }
            __syncwarp();
          }
        }
      }
      }  // column halves
      // ---- 4. accumulator drained: hand it back to the MMA warp (next-but-one tile)
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      if (et == 0 && it == 0) stamp(6);
      // ---- 5. coalesced write-out of the staged bf16 tile
      if (staged) {
        epi_bar();
        __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out);
        if (p.geglu) stage_copy<BN / 16, true>(p, t, stg, obase, p.ldc, t.n0 / 2, et);
        else stage_copy<BN / 8, true>(p, t, stg, obase, p.ldc, t.n0, et);
      }
      epi_bar();   // staging / bias tables are reused by the next item
      if (et == 0 && it == 0) stamp(7);
    }
    tc_fence_before();
  }

  // no CTA may exit while a peer can still multicast into its smem or arrive on its barriers
  if (csize > 1)
    cluster_sync_all();
  else
    __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- split-K finalize
template <bool F16>
__global__ void splitk_finalize_kernel(const float* __restrict__ partial, int splits, long long M, long long N,
                                       const float* __restrict__ bias, const float* __restrict__ bias_batch,
                                       long long rows_per_batch, long long bias_batch_ld,
                                       const __nv_bfloat16* __restrict__ residual, long long ldr,
                                       __nv_bfloat16* __restrict__ out, long long ldc) {
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
    float2 a = unpack16x2<F16>(r.x), b = unpack16x2<F16>(r.y);
    acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
  }
  uint2 o;
  o.x = pack16x2<F16>(acc.x, acc.y);
  o.y = pack16x2<F16>(acc.z, acc.w);
  *reinterpret_cast<uint2*>(out + m * ldc + n) = o;
}

static unsigned long long* g_timeline_host = nullptr;
static bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

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
  MOS_CHECK_DTYPE(a->a_dtype, "mos_gemm_bf16 (a_dtype)");
  MOS_CHECK_DTYPE(a->w_dtype, "mos_gemm_bf16 (w_dtype)");
  const bool f16 = a->a_dtype == MOS_DT_F16;
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
  p.n_tiles = (int)(a->N / BN);
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
    p.lgTW = ilog2(p.TW);
    p.lgTH = ilog2(p.TH);
    p.H = a->H;
    p.W = a->Wd;
    p.B = a->B;
    p.tiles_w = a->Wd / TW;
    p.tiles_h = (int)ceil_div(a->H, p.TH);
    int tiles_b = (int)ceil_div(a->B, p.TB);
    m_tiles = p.tiles_w * p.tiles_h * tiles_b;
    p.kc_per_tap = (int)(a->K / BK);
    p.kb_total = 9 * p.kc_per_tap;
  } else {
    MOS_CHECK_ARG(a->lda >= a->K && a->lda % 8 == 0, "mos_gemm_bf16: lda=%lld invalid", (long long)a->lda);
    m_tiles = (int)ceil_div(a->M, BM);
    p.kb_total = (int)(a->K / BK);
  }
  p.m_tiles = m_tiles;
  // ---- cluster shape: CX column neighbours share A, CM row neighbours share W
  // Measured on B200 (profiles/README.md): with 4 smem stages the multicast hand-shake (remote slot release + multicast
  // latency) costs more than the L2 traffic it saves (0.44 vs 0.36 us per k-block on the 320->320 conv), so clusters
  // are opt-in (MOS_GEMM_CLUSTER=1) until the stage budget grows (2-CTA UMMA, next round).
  static int use_cluster = -1;
  if (use_cluster < 0) {
    const char* e = getenv("MOS_GEMM_CLUSTER");
    use_cluster = (e && e[0] == '1') ? 1 : 0;
  }
  p.cx = p.cm = 1;
  if (use_cluster) {
    p.cx = (p.n_tiles % 2 == 0) ? 2 : 1;
    p.cm = (m_tiles % 4 == 0) ? 4 : (m_tiles % 2 == 0) ? 2 : 1;
  }
  const int csize = p.cx * p.cm;

  // ---- tensor maps (the A box is 1/CX of the tile rows, the W box 1/CM of the tile columns)
  if (a->conv) {
    uint32_t bw = (uint32_t)p.TW, bh = (uint32_t)p.TH, bb = (uint32_t)p.TB;
    if (p.cx == 2) {
      if (bb >= 2) { bb /= 2; p.half_dim = 3; p.half_off = (int)bb; }
      else if (bh >= 2) { bh /= 2; p.half_dim = 2; p.half_off = (int)bh; }
      else { bw /= 2; p.half_dim = 1; p.half_off = (int)bw; }
    }
    uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->Wd, (uint64_t)a->H, (uint64_t)a->B};
    const uint64_t pitch = (uint64_t)(a->lda > 0 ? a->lda : a->C);
    MOS_CHECK_ARG(pitch >= (uint64_t)a->C && pitch % 8 == 0, "mos_gemm_bf16: conv pixel pitch %llu invalid",
                  (unsigned long long)pitch);
    uint64_t str[3] = {pitch * 2, (uint64_t)a->Wd * pitch * 2, (uint64_t)a->H * a->Wd * pitch * 2};
    uint32_t box[4] = {BK, bw, bh, bb};
    int rc = encode_tmap(&tmA, a->A, 2, 4, dims, str, box, 3);
    if (rc) return rc;
    uint64_t wd[2] = {(uint64_t)a->K * 9, (uint64_t)a->N};
    uint64_t ws[1] = {(uint64_t)a->K * 9 * 2};
    uint32_t wb[2] = {BK, (uint32_t)(BN / p.cm)};
    rc = encode_tmap(&tmB, a->W, 2, 2, wd, ws, wb, 3);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->lda * 2};
    uint32_t box[2] = {BK, (uint32_t)(BM / p.cx)};
    int rc = encode_tmap(&tmA, a->A, 2, 2, dims, str, box, 3);
    if (rc) return rc;
    uint64_t wd[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t ws[1] = {(uint64_t)a->K * 2};
    uint32_t wb[2] = {BK, (uint32_t)(BN / p.cm)};
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
  p.accum = a->accumulate;
  p.tl = g_timeline_host;
  p.w_static = a->w_static;
  p.w_f16 = a->w_dtype == MOS_DT_F16;
  if (a->bias_batch && !a->conv)
    MOS_CHECK_ARG(p.rows_per_batch >= 32, "mos_gemm_bf16: bias_batch needs rows_per_batch >= 32 in plain mode");
  p.total_super = (p.n_tiles / p.cx) * (m_tiles / p.cm) * splits;
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
    MOS_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM));
    MOS_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM));
  }
  // persistent: one CTA per SM; each cluster loops over its share of (super-tile, split) work items
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  static int epi_warps = 0;
  if (epi_warps == 0) {
    const char* e = getenv("MOS_GEMM_EPI_WARPS");
    epi_warps = (e && e[0] == '4') ? 4 : 8;
  }
  cfg.blockDim = dim3(64 + 32 * epi_warps);
  cfg.dynamicSmemBytes = (size_t)MAX_DYN_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = (unsigned)csize;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // co-resident clusters per cluster size (GPC packing)
  if (max_clusters[csize] == 0) {
    int n = 0;
    cfg.gridDim = dim3((unsigned)(csize * (num_sms / csize)));
    cfg.numAttrs = 2;
    if (csize > 1) {
      cudaLaunchAttribute only_cluster[1] = {attr[1]};
      cfg.attrs = only_cluster;
      cfg.numAttrs = 1;
      MOS_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_kernel<false>, &cfg));
      cfg.attrs = attr;
      MOS_CHECK_ARG(n > 0, "mos_gemm_bf16: cluster size %d cannot be scheduled", csize);
    } else {
      n = num_sms;
    }
    max_clusters[csize] = n;
  }
  int clusters = max_clusters[csize];
  if (clusters > p.total_super) clusters = p.total_super;
  cfg.gridDim = dim3((unsigned)(clusters * csize));
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.numAttrs = csize > 1 ? 2 : 1;   // no cluster attribute at all for unclustered launches
  if (f16) MOS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<true>, tmA, tmB, tmL, p));
  else MOS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<false>, tmA, tmB, tmL, p));
  return MOS_OK;
}

extern "C" int mos_debug_set_timeline(void* buf) {
  mos::g_timeline_host = reinterpret_cast<unsigned long long*>(buf);
  return MOS_OK;
}

extern "C" int mos_splitk_finalize(const float* partial, int32_t splits, int64_t M, int64_t N, const float* bias,
                                   const float* bias_batch, int64_t rows_per_batch, int64_t bias_batch_ld,
                                   const void* residual, int64_t ldr, void* out, int64_t ldc, int32_t act_dtype,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(partial && out && splits >= 1 && M > 0 && N > 0 && N % 4 == 0, "mos_splitk_finalize: bad arguments");
  MOS_CHECK_DTYPE(act_dtype, "mos_splitk_finalize");
  long long total = M * (N / 4);
  int threads = 256;
  long long blocks = ceil_div(total, threads);
  MOS_CHECK_CUDA(launch_pdl(act_dtype == MOS_DT_F16 ? splitk_finalize_kernel<true> : splitk_finalize_kernel<false>,
                            dim3((unsigned)blocks), dim3(threads), 0, stream, partial,
                            (int)splits, (long long)M, (long long)N, bias, bias_batch,
                            (long long)(rows_per_batch > 0 ? rows_per_batch : 1),
                            (long long)(bias_batch_ld > 0 ? bias_batch_ld : N),
                            reinterpret_cast<const __nv_bfloat16*>(residual), (long long)ldr,
                            reinterpret_cast<__nv_bfloat16*>(out), (long long)ldc));
  return MOS_OK;
}
