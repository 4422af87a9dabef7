// b200mm — multi-head self-attention forward / backward for the short sequences of the ViT+BERT path
// (L = 197/257/577 vision tokens, 77..86 text / cross-modal tokens; head_dim 32/64/80).
//
// Reference arithmetic:
//   ViT : nn.MultiheadAttention inside ResidualAttentionBlock.attention, antmmf/modules/vision/backbone/clip/model.py:245-251
//         softmax(q k^T / sqrt(hd)) v, no mask, no dropout
//   BERT: BertSelfAttention.forward, antmmf/modules/vision/backbone/clip/modeling_bert.py:134-172
//         softmax(q k^T / sqrt(hd) + (1-mask)*-10000) v; the additive key bias is passed as `key_bias` [B, L] (fp32)
//
// Design: because a whole head's K and V (L x hd bf16 each, <= 92 KB) fit in shared memory, one CTA owns one
// (batch, head): K/V are staged once, each warp then streams 16-row query tiles against them with an online softmax
// (FlashAttention-2 register layout, mma.sync.m16n8k16 bf16 -> fp32). S and P never touch HBM; only O and the
// log-sum-exp per row are written. Backward is two kernels with the same shape and no atomics:
//   dq : per 16-query tile, recompute P, dP = dO V^T, dS = P*(dP - D), dQ = dS K        (also writes D = rowsum(dO*O))
//   dkv: per 16-key tile, recompute P^T against all queries, dV = P^T dO, dK = dS^T Q
// Round-1 note: these run on the legacy tensor path (HMMA); the tcgen05/TMEM version is planned.
//
// Layout: q/k/v live in one fused activation buffer [B, L, ld] (ld = 3*W for ViT's packed in_proj, same for the fused
// BERT q/k/v GEMM): head h of q at column q_off + h*hd, k at k_off + h*hd, v at v_off + h*hd. O is [B, L, ldo].
#include "common.cuh"

#include <stdlib.h>

#include <algorithm>

namespace b200mm {

struct AttnParams {
  const __nv_bfloat16* qkv;
  int64_t ld;
  int32_t q_off, k_off, v_off;
  __nv_bfloat16* o;  // fwd: output; bwd: forward output (read)
  int64_t ldo;
  float* lse;            // [B, H, L] natural-log LSE of the scaled+biased scores
  const float* key_bias; // [B, L] additive, or null
  int32_t B, H, L;
  float scale;
  // backward only
  const __nv_bfloat16* d_o;  // [B, L, ldo]
  __nv_bfloat16* dqkv;       // [B, L, ld] same column offsets as qkv
  float* dsum;               // [B, H, L]  D = rowsum(dO * O)
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int HD>
struct Cfg {
  static constexpr int PITCH = HD + 8;   // bf16 elements; +16 B keeps ldmatrix rows on distinct banks
  static constexpr int KSTEPS = HD / 16; // k16 steps over the head dimension
  static constexpr int NT_O = HD / 8;    // n8 tiles over the head dimension
  static constexpr int VEC_PER_ROW = HD / 8;
};

// cooperative copy of `rows` rows (global row stride ld_g) of one head into smem [rows_pad][PITCH]; rows >= L are zero
template <int HD>
__device__ __forceinline__ void stage_rows(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld_g, int L, int rows_pad,
                                           int tid, int nthreads) {
  constexpr int V = Cfg<HD>::VEC_PER_ROW;
  for (int i = tid; i < rows_pad * V; i += nthreads) {
    const int r = i / V, c = (i - r * V) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < L) v = *reinterpret_cast<const uint4*>(src + r * ld_g + c);
    *reinterpret_cast<uint4*>(dst + r * Cfg<HD>::PITCH + c) = v;
  }
}

// A-operand fragments (16 rows x HD) from a [16][PITCH] smem tile
template <int HD>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[Cfg<HD>::KSTEPS][4], const __nv_bfloat16* tile, int lane) {
  const int r = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int cofs = (lane >> 4) * 8;
#pragma unroll
  for (int ks = 0; ks < Cfg<HD>::KSTEPS; ++ks) ldsm_x4(a[ks], smem_u32(tile + r * Cfg<HD>::PITCH + ks * 16 + cofs));
}

// acc[8][4] (16 x 64) = A(16 x HD) * Bsm[n0 .. n0+64)[HD]^T ; Bsm rows are the n index, HD contiguous (non-transposed ldmatrix)
template <int HD>
__device__ __forceinline__ void mma_a_bT(float (&acc)[8][4], const uint32_t (&a)[Cfg<HD>::KSTEPS][4], const __nv_bfloat16* bsm,
                                         int n0, int n_valid16, int lane) {
  const int r = (lane & 7) + (lane >> 4) * 8;
  const int cofs = ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int np = 0; np < 4; ++np) {  // pairs of n8 tiles = 16 rows of bsm
    if (np < n_valid16) {
#pragma unroll
      for (int ks = 0; ks < Cfg<HD>::KSTEPS; ++ks) {
        uint32_t b[4];
        ldsm_x4(b, smem_u32(bsm + (n0 + np * 16 + r) * Cfg<HD>::PITCH + ks * 16 + cofs));
        mma_bf16_16816(acc[2 * np], a[ks], b[0], b[1]);
        mma_bf16_16816(acc[2 * np + 1], a[ks], b[2], b[3]);
      }
    }
  }
}

// out[NT_O][4] (16 x HD) += P(16 x 64, as A fragments) * Bsm[k0 .. k0+64)[HD] ; Bsm rows are the k index (transposed ldmatrix)
template <int HD>
__device__ __forceinline__ void mma_p_b(float (&out)[Cfg<HD>::NT_O][4], const uint32_t (&pa)[4][4], const __nv_bfloat16* bsm, int k0,
                                        int k_valid16, int lane) {
  const int r = (lane & 7) + ((lane >> 3) & 1) * 8;
  const int cofs = (lane >> 4) * 8;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (s < k_valid16) {
#pragma unroll
      for (int nt = 0; nt < Cfg<HD>::NT_O; nt += 2) {
        uint32_t b[4];
        ldsm_x4_t(b, smem_u32(bsm + (k0 + s * 16 + r) * Cfg<HD>::PITCH + nt * 8 + cofs));
        mma_bf16_16816(out[nt], pa[s], b[0], b[1]);
        if (nt + 1 < Cfg<HD>::NT_O) mma_bf16_16816(out[nt + 1], pa[s], b[2], b[3]);
      }
    }
  }
}

// C fragments of a 16 x 64 tile -> bf16 A fragments for the next mma (FlashAttention-2 register reuse)
__device__ __forceinline__ void c_to_a(uint32_t (&pa)[4][4], const float (&c)[8][4]) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    pa[s][0] = pack_bf16x2(c[2 * s][0], c[2 * s][1]);
    pa[s][1] = pack_bf16x2(c[2 * s][2], c[2 * s][3]);
    pa[s][2] = pack_bf16x2(c[2 * s + 1][0], c[2 * s + 1][1]);
    pa[s][3] = pack_bf16x2(c[2 * s + 1][2], c[2 * s + 1][3]);
  }
}

// write a 16 x HD fp32 C-fragment tile as bf16 to global through the warp's smem tile (16-byte coalesced stores)
template <int HD>
__device__ __forceinline__ void store_tile(const float (&c)[Cfg<HD>::NT_O][4], __nv_bfloat16* tile, __nv_bfloat16* gdst, int64_t ld_g,
                                           int row0, int L, int lane) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < Cfg<HD>::NT_O; ++nt) {
    *reinterpret_cast<uint32_t*>(tile + g * Cfg<HD>::PITCH + nt * 8 + 2 * t) = pack_bf16x2(c[nt][0], c[nt][1]);
    *reinterpret_cast<uint32_t*>(tile + (g + 8) * Cfg<HD>::PITCH + nt * 8 + 2 * t) = pack_bf16x2(c[nt][2], c[nt][3]);
  }
  __syncwarp();
  constexpr int V = Cfg<HD>::VEC_PER_ROW;
  for (int i = lane; i < 16 * V; i += 32) {
    const int r = i / V, cc = (i - r * V) * 8;
    if (row0 + r < L) *reinterpret_cast<uint4*>(gdst + static_cast<int64_t>(row0 + r) * ld_g + cc) = *reinterpret_cast<const uint4*>(tile + r * Cfg<HD>::PITCH + cc);
  }
  __syncwarp();
}

constexpr float LOG2E = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256) attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int Lp = (p.L + 15) & ~15;
  __nv_bfloat16* ksm = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* vsm = ksm + Lp * Cfg<HD>::PITCH;
  float* bias_sm = reinterpret_cast<float*>(vsm + Lp * Cfg<HD>::PITCH);  // [Lp], pre-multiplied by log2e; -inf past L
  __nv_bfloat16* wtiles = reinterpret_cast<__nv_bfloat16*>(bias_sm + Lp);

  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const __nv_bfloat16* base = p.qkv + static_cast<int64_t>(b) * p.L * p.ld + h * HD;
  stage_rows<HD>(ksm, base + p.k_off, p.ld, p.L, Lp, threadIdx.x, blockDim.x);
  stage_rows<HD>(vsm, base + p.v_off, p.ld, p.L, Lp, threadIdx.x, blockDim.x);
  for (int i = threadIdx.x; i < Lp; i += blockDim.x)
    bias_sm[i] = i < p.L ? (p.key_bias ? p.key_bias[static_cast<int64_t>(b) * p.L + i] * LOG2E : 0.f) : -INFINITY;
  __syncthreads();

  __nv_bfloat16* tile = wtiles + warp * 16 * Cfg<HD>::PITCH;
  const float sl2 = p.scale * LOG2E;
  const int g = lane >> 2, t = lane & 3;
  const int n_qtiles = Lp / 16;
  for (int qt = warp; qt < n_qtiles; qt += nwarps) {
    const int row0 = qt * 16;
    stage_rows<HD>(tile, base + p.q_off + static_cast<int64_t>(row0) * p.ld, p.ld, p.L - row0, 16, lane, 32);
    __syncwarp();
    uint32_t qa[Cfg<HD>::KSTEPS][4];
    load_a_frags<HD>(qa, tile, lane);
    float o[Cfg<HD>::NT_O][4];
#pragma unroll
    for (int i = 0; i < Cfg<HD>::NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    for (int kc = 0; kc < Lp; kc += 64) {
      const int nv16 = min(4, (Lp - kc) / 16);
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      mma_a_bT<HD>(s, qa, ksm, kc, nv16, lane);
      float cmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < 2 * nv16) {
          const float b0 = bias_sm[kc + nt * 8 + 2 * t], b1 = bias_sm[kc + nt * 8 + 2 * t + 1];
          s[nt][0] = fmaf(s[nt][0], sl2, b0); s[nt][1] = fmaf(s[nt][1], sl2, b1);
          s[nt][2] = fmaf(s[nt][2], sl2, b0); s[nt][3] = fmaf(s[nt][3], sl2, b1);
          cmax[0] = fmaxf(cmax[0], fmaxf(s[nt][0], s[nt][1]));
          cmax[1] = fmaxf(cmax[1], fmaxf(s[nt][2], s[nt][3]));
        } else {
          s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = -INFINITY;
        }
      }
      float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        cmax[r] = fmaxf(cmax[r], __shfl_xor_sync(0xffffffffu, cmax[r], 1));
        cmax[r] = fmaxf(cmax[r], __shfl_xor_sync(0xffffffffu, cmax[r], 2));
        const float m_new = fmaxf(m_run[r], cmax[r]);
        corr[r] = exp2f(m_run[r] - m_new);  // m_run = -inf on the first chunk -> 0
        m_run[r] = m_new;
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - m_run[0]); s[nt][1] = exp2f(s[nt][1] - m_run[0]);
        s[nt][2] = exp2f(s[nt][2] - m_run[1]); s[nt][3] = exp2f(s[nt][3] - m_run[1]);
        rs[0] += s[nt][0] + s[nt][1];
        rs[1] += s[nt][2] + s[nt][3];
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
        rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
        l_run[r] = l_run[r] * corr[r] + rs[r];
      }
#pragma unroll
      for (int i = 0; i < Cfg<HD>::NT_O; ++i) {
        o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
      }
      uint32_t pa[4][4];
      c_to_a(pa, s);
      mma_p_b<HD>(o, pa, vsm, kc, nv16, lane);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
    for (int i = 0; i < Cfg<HD>::NT_O; ++i) {
      o[i][0] *= inv0; o[i][1] *= inv0; o[i][2] *= inv1; o[i][3] *= inv1;
    }
    if (t == 0) {
      float* lse = p.lse + (static_cast<int64_t>(b) * p.H + h) * p.L;
      if (row0 + g < p.L) lse[row0 + g] = (m_run[0] + log2f(l_run[0])) / LOG2E;
      if (row0 + g + 8 < p.L) lse[row0 + g + 8] = (m_run[1] + log2f(l_run[1])) / LOG2E;
    }
    store_tile<HD>(o, tile, p.o + static_cast<int64_t>(b) * p.L * p.ldo + h * HD, p.ldo, row0, p.L, lane);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward, dQ (and D = rowsum(dO*O))
// ------------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int Lp = (p.L + 15) & ~15;
  __nv_bfloat16* ksm = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* vsm = ksm + Lp * Cfg<HD>::PITCH;
  float* bias_sm = reinterpret_cast<float*>(vsm + Lp * Cfg<HD>::PITCH);
  __nv_bfloat16* wtiles = reinterpret_cast<__nv_bfloat16*>(bias_sm + Lp);

  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const __nv_bfloat16* base = p.qkv + static_cast<int64_t>(b) * p.L * p.ld + h * HD;
  stage_rows<HD>(ksm, base + p.k_off, p.ld, p.L, Lp, threadIdx.x, blockDim.x);
  stage_rows<HD>(vsm, base + p.v_off, p.ld, p.L, Lp, threadIdx.x, blockDim.x);
  for (int i = threadIdx.x; i < Lp; i += blockDim.x)
    bias_sm[i] = i < p.L ? (p.key_bias ? p.key_bias[static_cast<int64_t>(b) * p.L + i] * LOG2E : 0.f) : -INFINITY;
  __syncthreads();

  __nv_bfloat16* tile = wtiles + warp * 2 * 16 * Cfg<HD>::PITCH;  // two tiles per warp: Q / dO (then O for D)
  __nv_bfloat16* tile2 = tile + 16 * Cfg<HD>::PITCH;
  const float sl2 = p.scale * LOG2E;
  const int g = lane >> 2, t = lane & 3;
  const int64_t bh = static_cast<int64_t>(b) * p.H + h;
  const __nv_bfloat16* obase = p.o + static_cast<int64_t>(b) * p.L * p.ldo + h * HD;
  const __nv_bfloat16* dobase = p.d_o + static_cast<int64_t>(b) * p.L * p.ldo + h * HD;
  for (int qt = warp; qt < Lp / 16; qt += nwarps) {
    const int row0 = qt * 16;
    // D = rowsum(dO * O): stage O and dO, two lanes per row
    stage_rows<HD>(tile, obase + static_cast<int64_t>(row0) * p.ldo, p.ldo, p.L - row0, 16, lane, 32);
    stage_rows<HD>(tile2, dobase + static_cast<int64_t>(row0) * p.ldo, p.ldo, p.L - row0, 16, lane, 32);
    __syncwarp();
    {
      const int r = lane >> 1, half = lane & 1;
      float acc = 0.f;
      for (int c = half * (HD / 2); c < (half + 1) * (HD / 2); ++c)
        acc += __bfloat162float(tile[r * Cfg<HD>::PITCH + c]) * __bfloat162float(tile2[r * Cfg<HD>::PITCH + c]);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (half == 0 && row0 + r < p.L) p.dsum[bh * p.L + row0 + r] = acc;
      // broadcast the two rows this thread needs (g, g+8)
      const float d0 = __shfl_sync(0xffffffffu, acc, 2 * g), d1 = __shfl_sync(0xffffffffu, acc, 2 * (g + 8));
      uint32_t da[Cfg<HD>::KSTEPS][4];
      load_a_frags<HD>(da, tile2, lane);
      __syncwarp();
      stage_rows<HD>(tile, base + p.q_off + static_cast<int64_t>(row0) * p.ld, p.ld, p.L - row0, 16, lane, 32);
      __syncwarp();
      uint32_t qa[Cfg<HD>::KSTEPS][4];
      load_a_frags<HD>(qa, tile, lane);
      const float* lse = p.lse + bh * p.L;
      const float l0 = row0 + g < p.L ? lse[row0 + g] * LOG2E : 0.f;
      const float l1 = row0 + g + 8 < p.L ? lse[row0 + g + 8] * LOG2E : 0.f;
      float dq[Cfg<HD>::NT_O][4];
#pragma unroll
      for (int i = 0; i < Cfg<HD>::NT_O; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
      for (int kc = 0; kc < Lp; kc += 64) {
        const int nv16 = min(4, (Lp - kc) / 16);
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
          dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
        }
        mma_a_bT<HD>(s, qa, ksm, kc, nv16, lane);
        mma_a_bT<HD>(dp, da, vsm, kc, nv16, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          if (nt < 2 * nv16) {
            const float b0 = bias_sm[kc + nt * 8 + 2 * t], b1 = bias_sm[kc + nt * 8 + 2 * t + 1];
            const float p0 = exp2f(fmaf(s[nt][0], sl2, b0) - l0), p1 = exp2f(fmaf(s[nt][1], sl2, b1) - l0);
            const float p2 = exp2f(fmaf(s[nt][2], sl2, b0) - l1), p3 = exp2f(fmaf(s[nt][3], sl2, b1) - l1);
            s[nt][0] = p0 * (dp[nt][0] - d0) * p.scale; s[nt][1] = p1 * (dp[nt][1] - d0) * p.scale;
            s[nt][2] = p2 * (dp[nt][2] - d1) * p.scale; s[nt][3] = p3 * (dp[nt][3] - d1) * p.scale;
          } else {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
          }
        }
        uint32_t dsa[4][4];
        c_to_a(dsa, s);
        mma_p_b<HD>(dq, dsa, ksm, kc, nv16, lane);
      }
      store_tile<HD>(dq, tile, p.dqkv + static_cast<int64_t>(b) * p.L * p.ld + p.q_off + h * HD, p.ld, row0, p.L, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward, dK and dV (one warp per 16-key tile, streaming over all queries of the head)
// ------------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256) attn_bwd_dkv_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int Lp = (p.L + 15) & ~15;
  __nv_bfloat16* qsm = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* dosm = qsm + Lp * Cfg<HD>::PITCH;
  float* lse_sm = reinterpret_cast<float*>(dosm + Lp * Cfg<HD>::PITCH);  // [Lp] * log2e; +inf past L (-> p = 0)
  float* d_sm = lse_sm + Lp;
  __nv_bfloat16* wtiles = reinterpret_cast<__nv_bfloat16*>(d_sm + Lp);

  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t bh = static_cast<int64_t>(b) * p.H + h;
  const __nv_bfloat16* base = p.qkv + static_cast<int64_t>(b) * p.L * p.ld + h * HD;
  stage_rows<HD>(qsm, base + p.q_off, p.ld, p.L, Lp, threadIdx.x, blockDim.x);
  stage_rows<HD>(dosm, p.d_o + static_cast<int64_t>(b) * p.L * p.ldo + h * HD, p.ldo, p.L, Lp, threadIdx.x, blockDim.x);
  for (int i = threadIdx.x; i < Lp; i += blockDim.x) {
    lse_sm[i] = i < p.L ? p.lse[bh * p.L + i] * LOG2E : INFINITY;
    d_sm[i] = i < p.L ? p.dsum[bh * p.L + i] : 0.f;
  }
  __syncthreads();

  __nv_bfloat16* tile = wtiles + warp * 16 * Cfg<HD>::PITCH;
  const float sl2 = p.scale * LOG2E;
  const int g = lane >> 2, t = lane & 3;
  for (int kt = warp; kt < Lp / 16; kt += nwarps) {
    const int row0 = kt * 16;
    stage_rows<HD>(tile, base + p.k_off + static_cast<int64_t>(row0) * p.ld, p.ld, p.L - row0, 16, lane, 32);
    __syncwarp();
    uint32_t ka[Cfg<HD>::KSTEPS][4], va[Cfg<HD>::KSTEPS][4];
    load_a_frags<HD>(ka, tile, lane);
    __syncwarp();
    stage_rows<HD>(tile, base + p.v_off + static_cast<int64_t>(row0) * p.ld, p.ld, p.L - row0, 16, lane, 32);
    __syncwarp();
    load_a_frags<HD>(va, tile, lane);
    // additive bias of this thread's two key rows (rows past L contribute nothing)
    float kb[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = row0 + g + 8 * r;
      kb[r] = key < p.L ? (p.key_bias ? p.key_bias[static_cast<int64_t>(b) * p.L + key] * LOG2E : 0.f) : -INFINITY;
    }
    float dk[Cfg<HD>::NT_O][4], dv[Cfg<HD>::NT_O][4];
#pragma unroll
    for (int i = 0; i < Cfg<HD>::NT_O; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int qc = 0; qc < Lp; qc += 64) {
      const int nv16 = min(4, (Lp - qc) / 16);
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      }
      mma_a_bT<HD>(st, ka, qsm, qc, nv16, lane);     // S^T tile: keys x queries
      mma_a_bT<HD>(dpt, va, dosm, qc, nv16, lane);   // dP^T tile
      uint32_t pa[4][4], dsa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float pv[4], dsv[4];
        if (nt < 2 * nv16) {
          const int q0 = qc + nt * 8 + 2 * t;
          const float lq0 = lse_sm[q0], lq1 = lse_sm[q0 + 1];
          const float dq0 = d_sm[q0], dq1 = d_sm[q0 + 1];
          pv[0] = exp2f(fmaf(st[nt][0], sl2, kb[0]) - lq0); pv[1] = exp2f(fmaf(st[nt][1], sl2, kb[0]) - lq1);
          pv[2] = exp2f(fmaf(st[nt][2], sl2, kb[1]) - lq0); pv[3] = exp2f(fmaf(st[nt][3], sl2, kb[1]) - lq1);
          dsv[0] = pv[0] * (dpt[nt][0] - dq0) * p.scale; dsv[1] = pv[1] * (dpt[nt][1] - dq1) * p.scale;
          dsv[2] = pv[2] * (dpt[nt][2] - dq0) * p.scale; dsv[3] = pv[3] * (dpt[nt][3] - dq1) * p.scale;
        } else {
          pv[0] = pv[1] = pv[2] = pv[3] = 0.f;
          dsv[0] = dsv[1] = dsv[2] = dsv[3] = 0.f;
        }
        const int s = nt >> 1, hi = (nt & 1) * 2;
        pa[s][hi] = pack_bf16x2(pv[0], pv[1]);   pa[s][hi + 1] = pack_bf16x2(pv[2], pv[3]);
        dsa[s][hi] = pack_bf16x2(dsv[0], dsv[1]); dsa[s][hi + 1] = pack_bf16x2(dsv[2], dsv[3]);
      }
      mma_p_b<HD>(dv, pa, dosm, qc, nv16, lane);
      mma_p_b<HD>(dk, dsa, qsm, qc, nv16, lane);
    }
    __nv_bfloat16* dbase = p.dqkv + static_cast<int64_t>(b) * p.L * p.ld + h * HD;
    store_tile<HD>(dk, tile, dbase + p.k_off, p.ld, row0, p.L, lane);
    store_tile<HD>(dv, tile, dbase + p.v_off, p.ld, row0, p.L, lane);
  }
}

// tcgen05 path (attention_tc.cu): returns 1 if it handled the call, 0 if the shape needs the legacy kernels below
int attention_fwd_tc(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, void* o, int64_t ldo, float* lse,
                     const float* key_bias, int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale, cudaStream_t stream);

int attention_bwd_tc(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* o, const void* d_o, int64_t ldo,
                     const float* lse, const float* key_bias, void* dqkv, float* dsum, int32_t B, int32_t H, int32_t L, int32_t head_dim,
                     float scale, cudaStream_t stream);

static int pick_warps(int L, int max_warps) {
  const int tiles = (L + 15) / 16;
  const int rounds = (tiles + max_warps - 1) / max_warps;
  return (tiles + rounds - 1) / rounds;
}

template <int HD>
static int launch_attn(int which, const AttnParams& p, cudaStream_t stream) {
  const int Lp = (p.L + 15) & ~15;
  const int tiles_per_warp = which == 1 ? 2 : 1;
  const size_t fixed = static_cast<size_t>(2) * Lp * Cfg<HD>::PITCH * 2 + (which == 2 ? 2 : 1) * Lp * 4;
  const size_t per_warp = static_cast<size_t>(tiles_per_warp) * 16 * Cfg<HD>::PITCH * 2;
  const size_t limit = 227 * 1024;
  if (fixed + per_warp > limit) {
    set_last_error("attention: L=%d head_dim=%d needs %zu B of shared memory (> 227 KB)", p.L, HD, fixed + per_warp);
    return B200MM_ERR_SHAPE;
  }
  const int max_warps = static_cast<int>(std::min<size_t>(8, (limit - fixed) / per_warp));
  const int nwarps = pick_warps(p.L, max_warps);
  const size_t smem = fixed + nwarps * per_warp;
  void (*kern)(const AttnParams) = which == 0 ? attn_fwd_kernel<HD> : (which == 1 ? attn_bwd_dq_kernel<HD> : attn_bwd_dkv_kernel<HD>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) {
    set_last_error("attention: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return B200MM_ERR_LAUNCH;
  }
  kern<<<p.B * p.H, nwarps * 32, smem, stream>>>(p);
  return check_launch("attention kernel");
}

static int dispatch_attn(int which, int hd, const AttnParams& p, cudaStream_t stream) {
  switch (hd) {
    case 16: return launch_attn<16>(which, p, stream);
    case 32: return launch_attn<32>(which, p, stream);
    case 64: return launch_attn<64>(which, p, stream);
    case 80: return launch_attn<80>(which, p, stream);
    default:
      set_last_error("attention: head_dim %d not supported (16, 32, 64, 80)", hd);
      return B200MM_ERR_SHAPE;
  }
}

int attention_fwd_v3(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, void* o, int64_t ldo, float* lse,
                     const float* key_bias, int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale, cudaStream_t stream);
int attention_bwd_v3(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* o, const void* d_o, int64_t ldo,
                     const float* lse, const float* key_bias, void* dqkv, float* dsum, int32_t B, int32_t H, int32_t L, int32_t head_dim,
                     float scale, cudaStream_t stream);

static int check_common(const char* who, const void* qkv, int64_t ld, const void* o, int64_t ldo, int32_t B, int32_t H, int32_t L,
                        int32_t hd, int32_t q_off, int32_t k_off, int32_t v_off) {
  B200MM_REQUIRE(B > 0 && H > 0 && L > 0, B200MM_ERR_SHAPE, "%s: B=%d H=%d L=%d", who, B, H, L);
  B200MM_REQUIRE(static_cast<int64_t>(B) * H < (1ll << 31), B200MM_ERR_SHAPE, "%s: B*H too large", who);
  B200MM_REQUIRE(qkv && o, B200MM_ERR_SHAPE, "%s: null pointer", who);
  B200MM_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0 && hd % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0,
                 B200MM_ERR_ALIGN, "%s: pitches/offsets must be multiples of 8 elements and bases 16B aligned", who);
  return B200MM_OK;
}

}  // namespace b200mm

using namespace b200mm;

extern "C" int b200mm_attention_fwd(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, void* o, int64_t ldo,
                                    float* lse, const float* key_bias, int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale,
                                    void* stream) {
  int rc = check_common("attention_fwd", qkv, ld, o, ldo, B, H, L, head_dim, q_off, k_off, v_off);
  if (rc) return rc;
  B200MM_REQUIRE(lse != nullptr, B200MM_ERR_SHAPE, "attention_fwd: lse is required");
  if (!getenv("B200MM_ATTN_OLD"))
    return attention_fwd_v3(qkv, ld, q_off, k_off, v_off, o, ldo, lse, key_bias, B, H, L, head_dim, scale, reinterpret_cast<cudaStream_t>(stream));
  rc = attention_fwd_tc(qkv, ld, q_off, k_off, v_off, o, ldo, lse, key_bias, B, H, L, head_dim, scale, reinterpret_cast<cudaStream_t>(stream));
  if (rc != 0) return rc < 0 ? rc : B200MM_OK;
  AttnParams p{};
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv); p.ld = ld; p.q_off = q_off; p.k_off = k_off; p.v_off = v_off;
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.ldo = ldo; p.lse = lse; p.key_bias = key_bias;
  p.B = B; p.H = H; p.L = L; p.scale = scale;
  return dispatch_attn(0, head_dim, p, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b200mm_attention_bwd(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* o,
                                    const void* d_o, int64_t ldo, const float* lse, const float* key_bias, void* dqkv, float* dsum,
                                    int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale, void* stream) {
  int rc = check_common("attention_bwd", qkv, ld, o, ldo, B, H, L, head_dim, q_off, k_off, v_off);
  if (rc) return rc;
  B200MM_REQUIRE(lse && d_o && dqkv && dsum, B200MM_ERR_SHAPE, "attention_bwd: null pointer");
  B200MM_REQUIRE((reinterpret_cast<uintptr_t>(d_o) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0, B200MM_ERR_ALIGN,
                 "attention_bwd: d_o/dqkv must be 16B aligned");
  if (!getenv("B200MM_ATTN_OLD"))
    return attention_bwd_v3(qkv, ld, q_off, k_off, v_off, o, d_o, ldo, lse, key_bias, dqkv, dsum, B, H, L, head_dim, scale,
                            reinterpret_cast<cudaStream_t>(stream));
  rc = attention_bwd_tc(qkv, ld, q_off, k_off, v_off, o, d_o, ldo, lse, key_bias, dqkv, dsum, B, H, L, head_dim, scale,
                        reinterpret_cast<cudaStream_t>(stream));
  if (rc != 0) return rc < 0 ? rc : B200MM_OK;
  AttnParams p{};
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv); p.ld = ld; p.q_off = q_off; p.k_off = k_off; p.v_off = v_off;
  p.o = const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(o)); p.ldo = ldo;
  p.lse = const_cast<float*>(lse); p.key_bias = key_bias;
  p.B = B; p.H = H; p.L = L; p.scale = scale;
  p.d_o = reinterpret_cast<const __nv_bfloat16*>(d_o);
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.dsum = dsum;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  rc = dispatch_attn(1, head_dim, p, s);
  if (rc) return rc;
  return dispatch_attn(2, head_dim, p, s);
}
