// attention.cu — K2/K3: flash attention on tcgen05 for the SD1.5 head sizes d = 40 / 80 / 160 (8 heads).
//
//   O = softmax(Q K^T * scale) V     per (batch, head);  self-attention (N x N) and cross-attention (N x 77)
//
// Replaces xformers.ops.memory_efficient_attention / attn.get_attention_scores + bmm at
// mixofshow/models/edlora.py:77-83,151-156 and pipeline_regionally_t2iadapter.py:111-116.
//
// One CTA = one 128-query tile of one (batch, head).  192 threads:
//   warps 0..3  softmax: tcgen05.ld S from TMEM, online softmax (exp2, warp-free: one thread owns one row),
//               P -> bf16 -> 128B-swizzled smem; O accumulated in registers from the per-tile P.V product
//   warp 4      TMA producer (Q once, K / V^T ring)
//   warp 5      TMEM allocator + tcgen05.mma issuer: S_j = Q K_j^T (TMEM, double buffered), PV_j = P_j V_j
// S_{j+1} is issued before PV_j so the tensor pipe overlaps the softmax of the next tile.
// Layouts (written by the QKV GEMM epilogue): Q,K [B*H, rows, DP] (DP = d padded to 64, pad = 0),
// V^T [B*H, DV, nk8] (keys contiguous), so every MMA operand is K-major SWIZZLE_128B.
#include "common.h"
#include "tc.cuh"

namespace mos {

// WIDE: d = 160 cross-attention variant (nk <= 128): one 128-key tile, single-buffered, so that the probability maps of
// the controller path come from a single kv tile at every head size.
template <int D, bool WIDE = false>
struct AttnCfg {
  static constexpr int KSTEPS = (D + 15) / 16;
  static constexpr int DP = ((D + 63) / 64) * 64;
  static constexpr int QCH = DP / 64;
  static constexpr int DV = ((D + 15) / 16) * 16;
  static constexpr int BKV = (D <= 80 || WIDE) ? 128 : 64;
  static constexpr int KVCH = BKV / 64;
  static constexpr int STAGES = WIDE ? 1 : 2;
  // d = 40: single S / P buffers and 256 TMEM columns so that TWO CTAs share an SM (the softmax of one hides the
  // MMA / TMEM latency of the other); larger head sizes keep double buffering and one CTA per SM.
  static constexpr int SB = (D <= 40 || WIDE) ? 1 : 2;
  static constexpr int PB = (D <= 40 || WIDE) ? 1 : 2;
  static constexpr int MINB = D <= 40 ? 2 : 1;
  static constexpr int Q_BYTES = QCH * 128 * 128;
  static constexpr int K_BYTES = QCH * BKV * 128;
  static constexpr int V_BYTES = KVCH * DV * 128;
  static constexpr int P_BYTES = KVCH * 128 * 128;
  static constexpr int O_STRIDE = ((DV + 63) / 64) * 64;
  static constexpr int S_COL0 = 0;
  static constexpr int O_COL0 = SB * BKV;
  static constexpr int TMEM_COLS = (O_COL0 + 2 * O_STRIDE <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * (K_BYTES + V_BYTES) + PB * P_BYTES + 1024;
  static_assert(O_COL0 + 2 * O_STRIDE <= TMEM_COLS, "TMEM budget");
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnDev {
  int nq, nk, heads;
  float scale_log2;  // scale * log2(e)
  __nv_bfloat16* out;
  long long ldo;
  float* probs;  // optional [B*H, nq, nk] fp32 (single kv tile only)
  float* lse2;   // optional [B*H, nq]: log2-domain log-sum-exp of scale*S (saved for the backward kernels)
  float* pcols;  // optional [B*H, nq, 2]: probabilities at key columns pos[b][0..1] (single kv tile only)
  const int* pos;
};

template <int D, bool WIDE>
__global__ void __launch_bounds__(192, AttnCfg<D, WIDE>::MINB)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const AttnDev p) {
  using C = AttnCfg<D, WIDE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + C::Q_BYTES;
  uint8_t* sP = sKV + C::STAGES * (C::K_BYTES + C::V_BYTES);

  __shared__ uint64_t q_full, kv_full[C::STAGES], kv_empty[C::STAGES];
  __shared__ uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2], o_full[2], o_empty[2];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int bh = blockIdx.y;
  const int T = (p.nk + C::BKV - 1) / C::BKV;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&q_full, 1);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
      mbar_init(&p_full[s], 128);
      mbar_init(&p_empty[s], 1);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_empty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(&tmem_holder, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  pdl_wait();               // Q / K / V^T come from the previous kernel in the stream
  pdl_launch_dependents();

  if (warp == 4) {
    // ================================================================= TMA producer
    if (lane == 0) {
      mbar_expect_tx(&q_full, C::Q_BYTES);
#pragma unroll
      for (int c = 0; c < C::QCH; ++c) tma_load_3d(sQ + c * 16384, &tmQ, &q_full, c * 64, q0, bh);
      for (int j = 0; j < T; ++j) {
        const int st = j % C::STAGES;
        mbar_wait(&kv_empty[st], ((j / C::STAGES) & 1) ^ 1);
        uint8_t* sK = sKV + st * (C::K_BYTES + C::V_BYTES);
        uint8_t* sV = sK + C::K_BYTES;
        mbar_expect_tx(&kv_full[st], C::K_BYTES + C::V_BYTES);
#pragma unroll
        for (int c = 0; c < C::QCH; ++c)
          tma_load_3d(sK + c * (C::BKV * 128), &tmK, &kv_full[st], c * 64, j * C::BKV, bh);
#pragma unroll
        for (int c = 0; c < C::KVCH; ++c)
          tma_load_3d(sV + c * (C::DV * 128), &tmV, &kv_full[st], j * C::BKV + c * 64, 0, bh);
      }
    }
  } else if (warp == 5) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(128, C::BKV, 1);
      const uint32_t idesc_o = make_idesc(128, C::DV, 1);
      mbar_wait(&q_full, 0);
      for (int j = 0; j <= T; ++j) {
        if (j < T) {
          const int st = j % C::STAGES, sb = j % C::SB;
          mbar_wait(&kv_full[st], (j / C::STAGES) & 1);
          mbar_wait(&s_empty[sb], ((j / C::SB) & 1) ^ 1);
          tc_fence_after();
          uint8_t* sK = sKV + st * (C::K_BYTES + C::V_BYTES);
#pragma unroll
          for (int kk = 0; kk < C::KSTEPS; ++kk) {
            uint64_t ad = make_desc_sw128(smem_u32(sQ + (kk >> 2) * 16384)) + 2 * (kk & 3);
            uint64_t bd = make_desc_sw128(smem_u32(sK + (kk >> 2) * (C::BKV * 128))) + 2 * (kk & 3);
            umma_bf16(tmem + C::S_COL0 + sb * C::BKV, ad, bd, idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[sb]);
        }
        if (j >= 1) {
          const int jj = j - 1, pb = jj % C::PB, ob = jj & 1, st = jj % C::STAGES;
          mbar_wait(&p_full[pb], (jj / C::PB) & 1);
          mbar_wait(&o_empty[ob], ((jj >> 1) & 1) ^ 1);
          tc_fence_after();
          uint8_t* sV = sKV + st * (C::K_BYTES + C::V_BYTES) + C::K_BYTES;
          uint8_t* sPb = sP + pb * C::P_BYTES;
          const int kv_valid = min(C::BKV, p.nk - jj * C::BKV);
          const int ksteps = (kv_valid + 15) >> 4;
          for (int kk = 0; kk < ksteps; ++kk) {
            uint64_t ad = make_desc_sw128(smem_u32(sPb + (kk >> 2) * 16384)) + 2 * (kk & 3);
            uint64_t bd = make_desc_sw128(smem_u32(sV + (kk >> 2) * (C::DV * 128))) + 2 * (kk & 3);
            umma_bf16(tmem + C::O_COL0 + ob * C::O_STRIDE, ad, bd, idesc_o, kk > 0 ? 1u : 0u);
          }
          umma_commit(&o_full[ob]);
          umma_commit(&p_empty[pb]);
          umma_commit(&kv_empty[st]);
        }
      }
    }
  } else {
    // ================================================================= softmax / accumulate (warps 0..3)
    const int r = warp * 32 + lane;  // query row in tile == TMEM lane
    const uint32_t trow = tmem + (uint32_t(warp * 32) << 16);
    const int q_idx = q0 + r;
    float acc[C::DV];
#pragma unroll
    for (int i = 0; i < C::DV; ++i) acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;

    auto accumulate = [&](int jj, float a) {
      const int ob = jj & 1;
      if (lane == 0) mbar_wait(&o_full[ob], (jj >> 1) & 1);   // one polling lane per warp
      __syncwarp();
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < C::DV / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(trow + C::O_COL0 + ob * C::O_STRIDE + c * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[c * 16 + i] = acc[c * 16 + i] * a + __uint_as_float(v[i]);
      }
      tc_fence_before();
      mbar_arrive(&o_empty[ob]);
    };

    for (int j = 0; j < T; ++j) {
      const int sb = j % C::SB, pbuf = j % C::PB;
      const int kv_valid = min(C::BKV, p.nk - j * C::BKV);
      if (lane == 0) mbar_wait(&s_full[sb], (j / C::SB) & 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t ts = trow + C::S_COL0 + sb * C::BKV;
      // pass 1: row max   (full tiles take the mask-free path: this kernel is issue-bound, ncu: 69 % issue active)
      const bool full = kv_valid == C::BKV;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < C::BKV / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(ts + c * 32, v);
        tmem_ld_wait();
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = ex2_approx(m_run - m_new);
      // pass 2: probabilities -> bf16 -> swizzled smem
      if (lane == 0) mbar_wait(&p_empty[pbuf], ((j / C::PB) & 1) ^ 1);
      __syncwarp();
      uint8_t* sPb = sP + pbuf * C::P_BYTES;
      float rs = 0.f;
      float pc0 = 0.f, pc1 = 0.f;
      int pos0 = -1, pos1 = -1;
      if (p.pcols != nullptr) {
        const int bb = bh / p.heads;
        pos0 = __ldg(p.pos + bb * 2) - j * C::BKV;
        pos1 = __ldg(p.pos + bb * 2 + 1) - j * C::BKV;
      }
#pragma unroll 1
      for (int c = 0; c < C::BKV / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(ts + c * 32, v);
        tmem_ld_wait();
        float pv[32];
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            pv[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
            rs += pv[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
            pv[i] = (c * 32 + i < kv_valid) ? e : 0.f;
            rs += pv[i];
          }
        }
        if (p.pcols != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            pc0 = (c * 32 + i == pos0) ? pv[i] : pc0;
            pc1 = (c * 32 + i == pos1) ? pv[i] : pc1;
          }
        }
        uint8_t* rowp = sPb + (c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(pv[g * 8 + 0], pv[g * 8 + 1]);
          u.y = pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]);
          u.z = pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]);
          u.w = pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]);
          const int c16 = (c & 1) * 4 + g;  // 16-byte chunk inside the 128-byte row
          *reinterpret_cast<uint4*>(rowp + ((c16 ^ (r & 7)) << 4)) = u;
        }
      }
      if (p.pcols != nullptr && T == 1 && q_idx < p.nq) {
        const float inv = 1.0f / rs;
        *reinterpret_cast<float2*>(p.pcols + ((long long)bh * p.nq + q_idx) * 2) = make_float2(pc0 * inv, pc1 * inv);
      }
      if (p.probs != nullptr && T == 1) {
        // normalised probabilities for the attention controller (edlora.py:81-82): single kv tile, so l = rs
        const float inv = 1.0f / rs;
        float* prow = p.probs + ((long long)bh * p.nq + q_idx) * p.nk;
#pragma unroll 1
        for (int c = 0; c < C::BKV / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(ts + c * 32, v);
          tmem_ld_wait();
          if (q_idx < p.nq) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < kv_valid) prow[c * 32 + i] = exp2f(__uint_as_float(v[i]) * p.scale_log2 - m_new) * inv;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(&s_empty[sb]);
      fence_proxy_async_smem();
      mbar_arrive(&p_full[pbuf]);
      l_run = l_run * alpha + rs;
      m_run = m_new;
      if (j >= 1) accumulate(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    accumulate(T - 1, alpha_prev);

    if (q_idx < p.nq) {
      const float inv = 1.0f / l_run;
      const int b = bh / p.heads, h = bh - b * p.heads;
      if (p.lse2 != nullptr) p.lse2[(long long)bh * p.nq + q_idx] = m_run + log2f(l_run);
      __nv_bfloat16* orow = p.out + ((long long)b * p.nq + q_idx) * p.ldo + h * D;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        uint4 u;
        u.x = pack_bf16x2(acc[c * 8 + 0] * inv, acc[c * 8 + 1] * inv);
        u.y = pack_bf16x2(acc[c * 8 + 2] * inv, acc[c * 8 + 3] * inv);
        u.z = pack_bf16x2(acc[c * 8 + 4] * inv, acc[c * 8 + 5] * inv);
        u.w = pack_bf16x2(acc[c * 8 + 6] * inv, acc[c * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + c * 8) = u;
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, C::TMEM_COLS);
  }
}

template <int D, bool WIDE>
static int launch_attn(const void* Q, const void* K, const void* Vt, void* out, int64_t ldo, float* probs, int BH,
                       int heads, int nq, int nk, int nk8, float scale, cudaStream_t stream, float* lse2 = nullptr,
                       float* pcols = nullptr, const int* pos = nullptr) {
  using C = AttnCfg<D, WIDE>;
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)C::DP, (uint64_t)nq, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)C::DP * 2, (uint64_t)nq * C::DP * 2};
    uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap(&tmQ, Q, 2, 3, dims, str, box, 3);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)C::DP, (uint64_t)nk, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)C::DP * 2, (uint64_t)nk * C::DP * 2};
    uint32_t box[3] = {64, (uint32_t)C::BKV, 1};
    int rc = encode_tmap(&tmK, K, 2, 3, dims, str, box, 3);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)nk8, (uint64_t)C::DV, (uint64_t)BH};
    uint64_t str[2] = {(uint64_t)nk8 * 2, (uint64_t)C::DV * nk8 * 2};
    uint32_t box[3] = {64, (uint32_t)C::DV, 1};
    int rc = encode_tmap(&tmV, Vt, 2, 3, dims, str, box, 3);
    if (rc) return rc;
  }
  AttnDev p;
  p.nq = nq;
  p.nk = nk;
  p.heads = heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.probs = probs;
  p.lse2 = lse2;
  p.pcols = pcols;
  p.pos = pos;
  static bool configured = false;
  if (!configured) {
    MOS_CHECK_CUDA(cudaFuncSetAttribute(attn_kernel<D, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(nq, 128), (unsigned)BH);
  MOS_CHECK_CUDA(launch_pdl(attn_kernel<D, WIDE>, grid, dim3(192), (size_t)C::SMEM_BYTES, stream, tmQ, tmK, tmV, p));
  return MOS_OK;
}

}  // namespace mos

using namespace mos;

extern "C" int mos_attention_fwd(const void* Q, const void* K, const void* Vt, void* out, int64_t ldo, float* probs,
                                 int32_t batch, int32_t heads, int32_t head_dim, int32_t nq, int32_t nk,
                                 int32_t nk8, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(Q && K && Vt && out, "mos_attention_fwd: NULL pointer");
  MOS_CHECK_ARG(batch > 0 && heads > 0 && nq > 0 && nk > 0, "mos_attention_fwd: bad shape");
  MOS_CHECK_ARG(nk8 >= nk && nk8 % 8 == 0, "mos_attention_fwd: nk8=%d must be >= nk=%d and a multiple of 8", nk8, nk);
  MOS_CHECK_ARG(ldo >= (int64_t)heads * head_dim && ldo % 8 == 0, "mos_attention_fwd: bad ldo");
  const int BH = batch * heads;
  if (probs) MOS_CHECK_ARG(nk <= 128, "mos_attention_fwd: probs output needs a single kv tile (nk <= 128)");
  switch (head_dim) {
    case 40: return launch_attn<40, false>(Q, K, Vt, out, ldo, probs, BH, heads, nq, nk, nk8, scale, stream);
    case 80: return launch_attn<80, false>(Q, K, Vt, out, ldo, probs, BH, heads, nq, nk, nk8, scale, stream);
    case 160:
      if (nk <= 128) return launch_attn<160, true>(Q, K, Vt, out, ldo, probs, BH, heads, nq, nk, nk8, scale, stream);
      return launch_attn<160, false>(Q, K, Vt, out, ldo, probs, BH, heads, nq, nk, nk8, scale, stream);
    default: return set_err(MOS_EUNSUPPORTED, "mos_attention_fwd: head_dim %d not in {40, 80, 160}", head_dim);
  }
}


// Training forward: same kernel, additionally saves the log2-domain log-sum-exp [B*H, nq] for mos_attention_bwd and
// (cross-attention, nk <= 128) the per-head probabilities at the two concept-token columns pos[b][0..1] -> pcols
// [B*H, nq, 2] for the attention regulariser (trainer_edlora.py:263-313).
extern "C" int mos_attention_fwd_train(const void* Q, const void* K, const void* Vt, void* out, int64_t ldo, float* lse2,
                                       float* pcols, const int32_t* pos, int32_t batch, int32_t heads,
                                       int32_t head_dim, int32_t nq, int32_t nk, int32_t nk8, float scale,
                                       void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MOS_CHECK_ARG(Q && K && Vt && out && lse2, "mos_attention_fwd_train: NULL pointer");
  MOS_CHECK_ARG(batch > 0 && heads > 0 && nq > 0 && nk > 0, "mos_attention_fwd_train: bad shape");
  MOS_CHECK_ARG(nk8 >= nk && nk8 % 8 == 0, "mos_attention_fwd_train: nk8 must be >= nk and a multiple of 8");
  MOS_CHECK_ARG(ldo >= (int64_t)heads * head_dim && ldo % 8 == 0, "mos_attention_fwd_train: bad ldo");
  MOS_CHECK_ARG(!pcols == !pos, "mos_attention_fwd_train: pcols and pos go together");
  if (pcols) MOS_CHECK_ARG(nk <= 128, "mos_attention_fwd_train: pcols needs a single kv tile (nk <= 128)");
  const int BH = batch * heads;
  const int* ip = reinterpret_cast<const int*>(pos);
  switch (head_dim) {
    case 40: return launch_attn<40, false>(Q, K, Vt, out, ldo, nullptr, BH, heads, nq, nk, nk8, scale, stream, lse2, pcols, ip);
    case 80: return launch_attn<80, false>(Q, K, Vt, out, ldo, nullptr, BH, heads, nq, nk, nk8, scale, stream, lse2, pcols, ip);
    case 160:
      if (nk <= 128)
        return launch_attn<160, true>(Q, K, Vt, out, ldo, nullptr, BH, heads, nq, nk, nk8, scale, stream, lse2, pcols, ip);
      return launch_attn<160, false>(Q, K, Vt, out, ldo, nullptr, BH, heads, nq, nk, nk8, scale, stream, lse2, pcols, ip);
    default: return set_err(MOS_EUNSUPPORTED, "mos_attention_fwd_train: head_dim %d not in {40, 80, 160}", head_dim);
  }
}
