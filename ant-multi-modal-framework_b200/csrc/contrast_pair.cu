// b200mm — grouped contrastive-loss kernels for the opt-in "two_sided" backend (b200mm_contrast_lse_partials_pair /
// b200mm_contrast_softgrad_pair, include/b200mm.h): both directions of a symmetric loss per launch, two-sided softmax-gradient tiles,
// device-resident scalars.
//
// The kernel is the warp-specialised single-CTA tcgen05 GEMM of gemm_tcgen05.cu (TMA producer warp, single-lane MMA issuer, TMEM
// double-buffered accumulators, epilogue warps) instantiated for the two contrastive epilogues, plus: a second problem (own tensor maps and
// ContrastParams) whose tiles follow the first one's in the persistent schedule, the column-LSE term of the two-sided gradient, alpha / coef
// read from device memory. It lives in its OWN translation unit on purpose: it was written after round 2's GPU budget was spent and has not
// run on hardware yet, and gemm_tcgen05.cu — every verified GEMM of the path — stays byte-identical to the build the GPU tests passed on.
#include "common.cuh"

#include <stdlib.h>

#include <type_traits>

namespace b200mm {
namespace pair {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = one 128B swizzle row
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int SLAB_BYTES = 64 * BK * 2;  // MN-major operands arrive as 64(mn) x 64(k) slabs of 8 KB
constexpr int EPI_WARPS = 12;                 // warp e: TMEM lane quarter e % 4; 32-column chunks c with c % 3 == e / 4
constexpr int EPI_PARTS = EPI_WARPS / 4;      // (3 warps per scheduler: the epilogue is latency-bound, not issue-bound)
constexpr int EPI_CHUNKS = BN / 32;
constexpr int GEMM_THREADS = 128 + EPI_WARPS * 32;
constexpr int EPI_STAGE_PITCH = 33;           // fp32 words per staged row (32 columns + 1: conflict-free both ways)
constexpr int EPI_STAGE_BYTES = 32 * EPI_STAGE_PITCH * 4;  // per epilogue warp
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * EPI_STAGE_BYTES + 1024;  // +1024: manual 1 KB alignment

// CTA-pair variant (cta_group::2): one 256 x 256 macro-tile per pair. Each CTA holds its 128 rows of A and HALF of the B
// tile (128 of the 256 N rows); the leader's tcgen05.mma reads both halves, so per-CTA operand traffic out of shared memory
// drops from 12 KB to 8 KB per k-step and the smaller stages allow a 6-deep ring.
template <int CG>
struct PairCfg {
  static constexpr int BN_CTA = BN / CG;
  static constexpr int B_BYTES = BN_CTA * BK * 2;
  static constexpr int STAGE = A_STAGE_BYTES + B_BYTES;
  static constexpr int NSTAGES = CG == 2 ? 5 : 3;  // what fits next to the 12 epilogue staging tiles (CG == 1 only serves M <= 128)
  static constexpr int SMEM = NSTAGES * STAGE + EPI_WARPS * EPI_STAGE_BYTES + 1024;
};
constexpr uint32_t TMEM_COLS = 512;

struct EpiParams {
  void* D;
  int64_t ldd;
  int32_t d_f32;
  float alpha;
  const __nv_bfloat16* bias;
  int32_t act;
  __nv_bfloat16* aux_out;
  const __nv_bfloat16* dact_in;
  int64_t ld_dact;
  const __nv_bfloat16* residual;
  int64_t ldr;
  uint32_t drop_thr;     // fused dropout (before the residual add): drop iff hash < drop_thr; 0 = off
  float drop_inv_keep;   // 1 / (1 - p)
  uint64_t drop_seed;
};

// epilogues of the contrastive-loss GEMMs (EPI_LSE / EPI_SOFTGRAD); z = alpha * <a_m, b_n> is the logit
struct ContrastParams {
  float* part_max;        // EPI_LSE: [M, n_tiles] running max of z over the tile's columns
  float* part_sum;        // EPI_LSE: [M, n_tiles] sum exp(z - part_max)
  float* diag;            // EPI_LSE: [M] z at column n == m + diag_off (the positive pair)
  int64_t diag_off;
  const float* row_lse;   // EPI_SOFTGRAD: [M] log-sum-exp each row is normalised with
  float coef;             // EPI_SOFTGRAD: dL/dz = coef * (exp(z - row_lse[m]) - diag_sub * [n == m + diag_off])
  float diag_sub;
  int32_t diag_zero;      // EPI_SOFTGRAD: force dL/dz = 0 on the diagonal (MIL-NCE text->video block)
  float* dscale;          // EPI_SOFTGRAD: += sum dL/dz * z  (gradient of the log-temperature), or null
  int64_t n_valid;        // EPI_SOFTGRAD: columns >= n_valid are padding (G = 0 there)
  int32_t* rank_out;      // EPI_RANK: [M] += #{n : z[m,n] > row_lse[m], n != positive column}  (row_lse doubles as the reference logit)
  const int32_t* gt_col;  // EPI_RANK: [M] positive column of each row, or null -> m + diag_off
  // EPI_SOFTGRAD, two-sided form: with col_lse the same logit tile also carries the gradient of the TRANSPOSED block (the other
  // direction of a symmetric loss, whose rows are this block's columns):
  //   dL/dz = coef * ( wr * exp(z - row_lse[m]) + wc * exp(z - col_lse[n]) - diag_sub * [diag] ),  wr / wc = 0 on the diagonal if *_diag_zero
  const float* col_lse;   // [N] or null (one-sided form above)
  int32_t row_diag_zero, col_diag_zero;
  int32_t diag_exact;     // 1: the stored G leaves the -diag_sub term out (dscale still counts it): the caller adds -diag_sub*coef*alpha*b_{m+off}
                          // to the row gradient in fp32, so the one large entry of a row is not rounded to bf16
  // device-resident scalars (no host sync to read a parameter or the upstream gradient): alpha = *alpha_dev, coef *= *coef_dev
  const float* alpha_dev;
  const float* coef_dev;
  void* D;                // EPI_SOFTGRAD output of THIS problem (grouped launches carry two problems, see GemmParams::n_prob)
};

enum { EPI_STD = 0, EPI_LSE = 1, EPI_SOFTGRAD = 2, EPI_RANK = 3 };

// epilogue wait policy: B200MM_GEMM_WAIT_NS=0 spins, otherwise the suspend hint in ns (default below)
static uint32_t gemm_wait_ns() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200MM_GEMM_WAIT_NS");
    v = e ? atoi(e) : 2000;
  }
  return static_cast<uint32_t>(v);
}

struct GemmParams {
  int64_t M, N, K;
  int32_t m_tiles, n_tiles, splits, kb_total, kb_per_split;
  float* partial;  // split-K: f32 [splits][M][N]; nullptr when splits == 1
  uint32_t mn_lbo, mn_sbo, k_lbo, k_sbo;  // descriptor byte offsets (defaults below; env-overridable for bring-up)
  uint32_t wait_ns;  // > 0: epilogue warps wait for the accumulator with a suspending try_wait (hint in ns) instead of spinning
  EpiParams epi;
  ContrastParams con;
  // grouped launch of the contrastive epilogues: problem 1 (same M, N, K, pitches of D) has its own operands (tmA2 / tmB2) and ContrastParams;
  // its tiles follow problem 0's in the persistent schedule, so one launch fills the SMs where two half-empty waves ran before
  int32_t n_prob;
  ContrastParams con2;
};

// Full epilogue on 8 consecutive columns of one row (n % 8 == 0). Shared by the GEMM epilogue warps and the split-K
// reduction kernel.
__device__ __forceinline__ void epi_apply8(float (&v)[8], int64_t m, int64_t n, const EpiParams& e) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= e.alpha;
  if (e.bias != nullptr) {
    uint4 b = *reinterpret_cast<const uint4*>(e.bias + n);
    float2 f0 = unpack_bf16x2(b.x), f1 = unpack_bf16x2(b.y), f2 = unpack_bf16x2(b.z), f3 = unpack_bf16x2(b.w);
    v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y;
    v[4] += f2.x; v[5] += f2.y; v[6] += f3.x; v[7] += f3.y;
  }
  if (e.aux_out != nullptr && e.dact_in == nullptr) {
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(e.aux_out + m * e.ldd + n) = o;
  }
  if (e.dact_in != nullptr) {
    uint4 u = *reinterpret_cast<const uint4*>(e.dact_in + m * e.ld_dact + n);
    float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
    float x[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= apply_dact(e.act, x[j]);
    if (e.aux_out != nullptr) {  // with dact_in the auxiliary output is the recomputed activation act(u)
      uint4 o;
      o.x = pack_bf16x2(apply_act(e.act, x[0]), apply_act(e.act, x[1])); o.y = pack_bf16x2(apply_act(e.act, x[2]), apply_act(e.act, x[3]));
      o.z = pack_bf16x2(apply_act(e.act, x[4]), apply_act(e.act, x[5])); o.w = pack_bf16x2(apply_act(e.act, x[6]), apply_act(e.act, x[7]));
      *reinterpret_cast<uint4*>(e.aux_out + m * e.ldd + n) = o;
    }
  } else if (e.act != B200MM_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(e.act, v[j]);
  }
  if (e.drop_thr != 0u) {
    const uint32_t key = drop_stream_key(e.drop_seed, static_cast<uint64_t>(m));
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = drop_keep(key, static_cast<uint32_t>(n) + j, e.drop_thr) ? v[j] * e.drop_inv_keep : 0.f;
  }
  if (e.residual != nullptr) {
    uint4 r = *reinterpret_cast<const uint4*>(e.residual + m * e.ldr + n);
    float2 f0 = unpack_bf16x2(r.x), f1 = unpack_bf16x2(r.y), f2 = unpack_bf16x2(r.z), f3 = unpack_bf16x2(r.w);
    v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y;
    v[4] += f2.x; v[5] += f2.y; v[6] += f3.x; v[7] += f3.y;
  }
  if (e.d_f32) {
    float* d = reinterpret_cast<float*>(e.D) + m * e.ldd + n;
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.D) + m * e.ldd + n) = o;
  }
}

// compile-time flavour encoding: bit0 bias, bit1 aux_out, bit2 residual, bit3 dact, bit4 f32 out, bit5 alpha != 1, bits 6-7 act, bit8 dropout
__host__ __device__ constexpr int flavor_bits(bool bias, bool aux, bool res, bool dact, bool f32, bool scale, int act, bool drop = false) {
  return (bias ? 1 : 0) | (aux ? 2 : 0) | (res ? 4 : 0) | (dact ? 8 : 0) | (f32 ? 16 : 0) | (scale ? 32 : 0) | (act << 6) | (drop ? 256 : 0);
}

// Lean per-8-column epilogue for the GEMM warps: every address and flag is resolved by the caller once per tile / chunk;
// here only arithmetic, one optional aux store and the output store remain.
struct EpiFlags {
  bool has_bias, has_aux, has_res, has_dact, f32, scale, drop;
  int act;
};
struct EpiDrop {  // dropout of this thread's 8 columns: the row's stream key, first column, threshold, 1 / (1 - p)
  uint32_t key, col, thr;
  float inv_keep;
};
__device__ __forceinline__ void epi_lean8(float (&v)[8], const EpiFlags& f, float alpha, const uint4& bias8, const uint4& ext8, void* dptr,
                                          __nv_bfloat16* auxptr, const EpiDrop& dr) {
  if (f.scale) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= alpha;
  }
  if (f.has_bias) {
    const float2 f0 = unpack_bf16x2(bias8.x), f1 = unpack_bf16x2(bias8.y), f2 = unpack_bf16x2(bias8.z), f3 = unpack_bf16x2(bias8.w);
    v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y;
    v[4] += f2.x; v[5] += f2.y; v[6] += f3.x; v[7] += f3.y;
  }
  if (f.has_aux && !f.has_dact) {
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(auxptr) = o;
  }
  if (f.has_dact || f.has_res) {
    const float2 x0 = unpack_bf16x2(ext8.x), x1 = unpack_bf16x2(ext8.y), x2 = unpack_bf16x2(ext8.z), x3 = unpack_bf16x2(ext8.w);
    const float x[8] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y, x3.x, x3.y};
    if (f.has_dact) {
      // with aux_out the activation act(u) itself is emitted next to D = acc * act'(u): both share the sigmoid / Gaussian terms, and
      // the backward pass gets the activated hidden (operand of the following weight gradient) without a separate recompute kernel
      float g[8];
      if (f.act == B200MM_ACT_QUICKGELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sg = sigmoid_1702(x[j]);
          g[j] = x[j] * sg;
          v[j] *= sg * fmaf(1.702f * x[j], 1.f - sg, 1.f);
        }
      } else if (f.act == B200MM_ACT_GELU_ERF) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float cdf, expo;
          gauss_cdf_pdf(x[j], cdf, expo);
          g[j] = x[j] * cdf;
          v[j] *= fmaf(x[j] * 0.3989422804014327f, expo, cdf);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = x[j];
      }
      if (f.has_aux) {
        uint4 o;
        o.x = pack_bf16x2(g[0], g[1]); o.y = pack_bf16x2(g[2], g[3]);
        o.z = pack_bf16x2(g[4], g[5]); o.w = pack_bf16x2(g[6], g[7]);
        *reinterpret_cast<uint4*>(auxptr) = o;
      }
    } else {
      if (f.act == B200MM_ACT_QUICKGELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_quickgelu(v[j]);
      } else if (f.act == B200MM_ACT_GELU_ERF) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_gelu_erf(v[j]);
      }
      if (f.drop) {  // dropout(dense(x)) BEFORE the residual joins (BertSelfOutput / BertOutput)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = drop_keep(dr.key, dr.col + j, dr.thr) ? v[j] * dr.inv_keep : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += x[j];
    }
  } else {
    if (f.act == B200MM_ACT_QUICKGELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = act_quickgelu(v[j]);
    } else if (f.act == B200MM_ACT_GELU_ERF) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = act_gelu_erf(v[j]);
    }
    if (f.drop) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = drop_keep(dr.key, dr.col + j, dr.thr) ? v[j] * dr.inv_keep : 0.f;
    }
  }
  if (f.f32) {
    float* d = reinterpret_cast<float*>(dptr);
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dptr) = o;
  }
}

struct TileCoord {
  int32_t m_blk, n_blk, split, kb0, kb1;
};
__device__ __forceinline__ TileCoord decode_tile(int64_t t, const GemmParams& p) {
  TileCoord c;
  int64_t per_split = static_cast<int64_t>(p.m_tiles) * p.n_tiles;
  if (t >= per_split * p.splits) t -= per_split * p.splits;  // second problem of a grouped launch: same tile grid
  c.split = static_cast<int32_t>(t / per_split);
  int64_t rem = t - c.split * per_split;
  c.m_blk = static_cast<int32_t>(rem / p.n_tiles);
  c.n_blk = static_cast<int32_t>(rem - static_cast<int64_t>(c.m_blk) * p.n_tiles);
  c.kb0 = c.split * p.kb_per_split;
  c.kb1 = min(p.kb_total, c.kb0 + p.kb_per_split);
  return c;
}

// FL >= 0 bakes the epilogue flags (see flavor_bits) into the kernel so that the per-element code has no flag branches;
// FL == -1 keeps them as runtime values (any combination, edge flavours).
template <bool A_MN, bool B_MN, int EPI, int CG, int FL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
contrast_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmB2, const GemmParams p) {
  using PC = PairCfg<CG>;
  constexpr int STAGES = PC::NSTAGES;           // shadows the 1-CTA constants inside this kernel
  constexpr int STAGE_BYTES = PC::STAGE;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  // 1 KB alignment for the 128B swizzle; offset arithmetic on the __shared__ array keeps the shared address space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], EPI_WARPS * 32 * CG);  // pair: both CTAs' epilogue threads report to the leader
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) tmem_alloc_cg2(&tmem_base_smem, TMEM_COLS);
    else tmem_alloc(&tmem_base_smem, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const int64_t tile0 = blockIdx.x / CG, tile_stride = gridDim.x / CG;  // a pair walks the macro-tile list together

  const int64_t tiles_per_prob = static_cast<int64_t>(p.m_tiles) * p.n_tiles * p.splits;
  const int64_t total_tiles = tiles_per_prob * (EPI == EPI_STD ? 1 : p.n_prob);

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t t = tile0; t < total_tiles; t += tile_stride) {
      const TileCoord tc = decode_tile(t, p);
      const bool second = EPI != EPI_STD && t >= tiles_per_prob;
      const CUtensorMap* const pA = second ? &tmA2 : &tmA;
      const CUtensorMap* const pB = second ? &tmB2 : &tmB;
      const int32_t m0 = tc.m_blk * (BM * CG) + static_cast<int32_t>(cta_rank) * BM;
      const int32_t n0 = tc.n_blk * BN + static_cast<int32_t>(cta_rank) * PC::BN_CTA;  // this CTA's share of the B tile
      for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
        mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          const int32_t k0 = kb * BK;
          if constexpr (CG == 1) {
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            if constexpr (!A_MN) {
              tma_load_2d(pA, &full_bar[stage], sa, k0, m0);
            } else {
#pragma unroll
              for (int s = 0; s < BM / 64; ++s) tma_load_2d(pA, &full_bar[stage], sa + s * SLAB_BYTES, m0 + 64 * s, k0);
            }
            if constexpr (!B_MN) {
              tma_load_2d(pB, &full_bar[stage], sb, k0, n0);
            } else {
#pragma unroll
              for (int s = 0; s < BN / 64; ++s) tma_load_2d(pB, &full_bar[stage], sb + s * SLAB_BYTES, n0 + 64 * s, k0);
            }
          } else {
            // both CTAs land their bytes on the LEADER's full barrier; only the leader arms it (for the pair's total)
            const uint32_t lead_bar = map_to_cta(&full_bar[stage], 0);
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], STAGE_BYTES * CG);
            if constexpr (!A_MN) {
              tma_load_2d_cg2(&tmA, lead_bar, sa, k0, m0);
            } else {
#pragma unroll
              for (int s = 0; s < BM / 64; ++s) tma_load_2d_cg2(&tmA, lead_bar, sa + s * SLAB_BYTES, m0 + 64 * s, k0);
            }
            if constexpr (!B_MN) {
              tma_load_2d_cg2(&tmB, lead_bar, sb, k0, n0);
            } else {
#pragma unroll
              for (int s = 0; s < PC::BN_CTA / 64; ++s) tma_load_2d_cg2(&tmB, lead_bar, sb + s * SLAB_BYTES, n0 + 64 * s, k0);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ===================== MMA issuer (leader CTA of a pair issues for both) =====================
    constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t t = tile0; t < total_tiles; t += tile_stride) {
      const TileCoord tc = decode_tile(t, p);
      mbar_wait_relaxed(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
        mbar_wait_relaxed(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_base = a_base + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 k-elements = 32 B inside the 128B swizzle row; 8-row groups 1024 B apart (SBO).
            // MN-major: 16 k-rows = 2048 B; 8-k-row groups 1024 B apart (SBO); 64-wide mn slabs 8 KB apart (LBO).
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_base + k * 2048, p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc_sw128(a_base + k * 32, p.k_lbo, p.k_sbo);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_base + k * 2048, p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc_sw128(b_base + k * 32, p.k_lbo, p.k_sbo);
            if constexpr (CG == 2) umma_bf16_cg2(d_tmem, adesc, bdesc, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > tc.kb0 || k > 0) ? 1u : 0u);
          }
          // smem stage reusable (in both CTAs of a pair) once these MMAs have read it
          if constexpr (CG == 2) umma_commit_cg2(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) {  // accumulator complete
        if constexpr (CG == 2) umma_commit_cg2(&tmem_full_bar[acc]);
        else umma_commit(&tmem_full_bar[acc]);
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;
    const int quarter = ew & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
    const int cpart = ew >> 2;   // this warp's 32-column chunks: c = cpart, cpart + 3, (cpart + 6)
    float* stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES) + ew * (32 * EPI_STAGE_PITCH);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t t = tile0; t < total_tiles; t += tile_stride) {
      const TileCoord tc = decode_tile(t, p);
      const ContrastParams& con = (EPI != EPI_STD && t >= tiles_per_prob) ? p.con2 : p.con;
      const float alpha = (EPI != EPI_STD && con.alpha_dev != nullptr) ? *con.alpha_dev : p.epi.alpha;
      const int64_t m_cta = static_cast<int64_t>(tc.m_blk) * (BM * CG) + cta_rank * BM;  // first row of this CTA's 128
      const int64_t m = m_cta + quarter * 32 + lane;
      const int64_t n0 = static_cast<int64_t>(tc.n_blk) * BN;
      // Everything that does not depend on the chunk is resolved once per tile: flags, the 4 row pointers of this lane
      // (row = it*8 + lane/4, 8 columns starting at (lane%4)*8) for D, aux_out and the side input (dact_in if set, else
      // residual), and row validity. The side input is fetched one chunk ahead (the first chunk before the accumulator is
      // even ready), so its HBM latency overlaps the MMA / the previous chunk's math.
      const int rr = lane >> 2, cg = (lane & 3) * 8;
      const int64_t m_base = m_cta + quarter * 32;
      EpiFlags fl{};
      // row pointers of this lane for it = 0 and the byte / element step to the next `it` (8 rows further down)
      const __nv_bfloat16* ext0 = nullptr;
      uint8_t* d0 = nullptr;
      __nv_bfloat16* aux0 = nullptr;  // aux_out shares D's pitch but is always bf16
      int64_t d_step = 0, ext_step = 0, aux_step = 0;
      uint32_t row_ok = 0;
      uint32_t drop_key[4] = {0u, 0u, 0u, 0u};  // dropout stream keys of this lane's four rows
      uint4 nxt[4];
      const bool is_partial = (EPI == EPI_STD) && p.partial != nullptr;
      if constexpr (EPI == EPI_STD) {
        const int esz = is_partial || p.epi.d_f32 ? 4 : 2;
        if (!is_partial) {
          fl.has_bias = p.epi.bias != nullptr; fl.has_aux = p.epi.aux_out != nullptr; fl.has_res = p.epi.residual != nullptr;
          fl.has_dact = p.epi.dact_in != nullptr; fl.f32 = p.epi.d_f32 != 0; fl.scale = p.epi.alpha != 1.f; fl.act = p.epi.act;
          fl.drop = p.epi.drop_thr != 0u;
        } else {
          fl.f32 = true;
        }
        if constexpr (FL >= 0) {  // the host guarantees that the runtime arguments match the baked-in flavour
          fl.has_bias = (FL & 1) != 0; fl.has_aux = (FL & 2) != 0; fl.has_res = (FL & 4) != 0; fl.has_dact = (FL & 8) != 0;
          fl.f32 = (FL & 16) != 0; fl.scale = (FL & 32) != 0; fl.act = (FL >> 6) & 3; fl.drop = (FL & 256) != 0;
        }
        const __nv_bfloat16* ext = is_partial ? nullptr : (fl.has_dact ? p.epi.dact_in : p.epi.residual);
        const int64_t ld_ext = fl.has_dact ? p.epi.ld_dact : p.epi.ldr;
        uint8_t* dbase = is_partial ? reinterpret_cast<uint8_t*>(p.partial + static_cast<int64_t>(tc.split) * p.M * p.N)
                                    : reinterpret_cast<uint8_t*>(p.epi.D);
        const int64_t ldd = is_partial ? p.N : p.epi.ldd;
        const int64_t mm0 = m_base + rr;
        d0 = dbase + (mm0 * ldd + n0 + cg) * esz;
        d_step = 8 * ldd * esz;
        if (fl.has_aux) { aux0 = p.epi.aux_out + mm0 * ldd + n0 + cg; aux_step = 8 * ldd; }
        if (ext != nullptr) { ext0 = ext + mm0 * ld_ext + n0 + cg; ext_step = 8 * ld_ext; }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t mm = mm0 + it * 8;
          if (mm < p.M) row_ok |= 1u << it;
          if (fl.drop) drop_key[it] = drop_stream_key(p.epi.drop_seed, static_cast<uint64_t>(mm));
          nxt[it] = make_uint4(0u, 0u, 0u, 0u);
          if (ext != nullptr && mm < p.M && n0 + cpart * 32 + cg < p.N) nxt[it] = *reinterpret_cast<const uint4*>(ext0 + it * ext_step + cpart * 32);
        }
      }
      // bias of this lane's 8 columns: like the side input it is fetched one chunk ahead (first chunk: before the accumulator is ready)
      uint4 nxt_bias = make_uint4(0u, 0u, 0u, 0u);
      if constexpr (EPI == EPI_STD) {
        if (fl.has_bias && n0 + cpart * 32 + cg < p.N) nxt_bias = *reinterpret_cast<const uint4*>(p.epi.bias + n0 + cpart * 32 + cg);
      }
      if (p.wait_ns) mbar_wait_suspend(&tmem_full_bar[acc], acc_phase, p.wait_ns);
      else mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      if constexpr (EPI == EPI_STD) {
        // TMEM gives each thread one row (32 consecutive columns). Going to HBM like that would touch 32 different lines per
        // warp instruction, so the 32x32 chunk is transposed through a warp-private smem tile: afterwards 4 lanes cover
        // 64 contiguous bytes of one row and every load/store of the fused epilogue is sector-exact.
        const int esz = fl.f32 ? 4 : 2;
        const bool has_ext = ext0 != nullptr;
        // interior tiles (all 128 rows and 256 columns valid) take the predicate-free instantiation of the chunk loop
        auto chunk_loop = [&](auto full_tag) {
          constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll 1
          for (int c = cpart; c < EPI_CHUNKS; c += EPI_PARTS) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait32(r);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) stage[lane * EPI_STAGE_PITCH + j] = __uint_as_float(r[j]);
            __syncwarp();
            const bool col_ok = FULL || n0 + c * 32 + cg < p.N;
            uint4 cur[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) cur[it] = nxt[it];
            const bool next_ok = c + EPI_PARTS < EPI_CHUNKS && (FULL || n0 + (c + EPI_PARTS) * 32 + cg < p.N);
            if (has_ext && next_ok) {
#pragma unroll
              for (int it = 0; it < 4; ++it)
                if (FULL || (row_ok & (1u << it))) nxt[it] = *reinterpret_cast<const uint4*>(ext0 + it * ext_step + (c + EPI_PARTS) * 32);
            }
            const uint4 bias8 = nxt_bias;
            if (fl.has_bias && next_ok) nxt_bias = *reinterpret_cast<const uint4*>(p.epi.bias + n0 + (c + EPI_PARTS) * 32 + cg);
            if (col_ok) {
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                if (FULL || (row_ok & (1u << it))) {
                  float v[8];
                  const float* srow = stage + (it * 8 + rr) * EPI_STAGE_PITCH + cg;
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = srow[j];
                  const EpiDrop dr{drop_key[it], static_cast<uint32_t>(n0) + c * 32 + cg, p.epi.drop_thr, p.epi.drop_inv_keep};
                  epi_lean8(v, fl, p.epi.alpha, bias8, cur[it], d0 + it * d_step + c * 32 * esz, aux0 + it * aux_step + c * 32, dr);
                }
              }
            }
          }
        };
        if (m_cta + BM <= p.M && n0 + BN <= p.N) chunk_loop(std::true_type{});
        else chunk_loop(std::false_type{});
      } else if constexpr (EPI == EPI_LSE) {
        // online (max, sum-exp) over this tile's columns of row m; the diagonal logit is captured on the way
        float mx = -INFINITY, sm = 0.f;
        const int64_t dcol = m + con.diag_off;
#pragma unroll 1
        for (int c = cpart; c < EPI_CHUNKS; c += EPI_PARTS) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait32(r);
          const int64_t nb = n0 + c * 32;
          if (nb < p.N) {
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float z = nb + j < p.N ? __uint_as_float(r[j]) * alpha : -INFINITY;
              r[j] = __float_as_uint(z);
              cm = fmaxf(cm, z);
            }
            const float nm = fmaxf(mx, cm);
            float cs = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) cs += __expf(__uint_as_float(r[j]) - nm);
            sm = sm * __expf(mx - nm) + cs;
            mx = nm;
            if (m < p.M && dcol >= nb && dcol < nb + 32 && dcol < p.N) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j == dcol) con.diag[m] = __uint_as_float(r[j]);
            }
          }
        }
        if (m < p.M) {
          con.part_max[m * (EPI_PARTS * p.n_tiles) + EPI_PARTS * tc.n_blk + cpart] = mx;
          con.part_sum[m * (EPI_PARTS * p.n_tiles) + EPI_PARTS * tc.n_blk + cpart] = sm;
        }
      } else if constexpr (EPI == EPI_RANK) {
        // retrieval rank of the positive: how many logits of row m beat the reference logit (strictly), the positive itself excluded
        const float ref = m < p.M ? con.row_lse[m] : INFINITY;
        const int64_t dcol = m < p.M ? (con.gt_col != nullptr ? static_cast<int64_t>(con.gt_col[m]) : m + con.diag_off) : -1;
        int cnt = 0;
#pragma unroll 1
        for (int c = cpart; c < EPI_CHUNKS; c += EPI_PARTS) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait32(r);
          const int64_t nb = n0 + c * 32;
          if (nb < p.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              cnt += (nb + j < p.N && nb + j != dcol && __uint_as_float(r[j]) * alpha > ref) ? 1 : 0;
          }
        }
        if (m < p.M && cnt) atomicAdd(con.rank_out + m, cnt);
      } else {
        const float lse = m < p.M ? con.row_lse[m] : 0.f;
        const int64_t dcol = m + con.diag_off;
        const float coef = con.coef_dev != nullptr ? con.coef * *con.coef_dev : con.coef;
        const bool two_sided = con.col_lse != nullptr;
        float ds_acc = 0.f;
        __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(con.D) + m * p.epi.ldd;
#pragma unroll 1
        for (int c = cpart; c < EPI_CHUNKS; c += EPI_PARTS) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait32(r);
          if (m < p.M) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int64_t n = n0 + c * 32 + g * 8;
              if (n < p.N) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float z = __uint_as_float(r[g * 8 + j]) * alpha;
                  const bool on_diag = (n + j == dcol);
                  float gz;
                  if (two_sided) {
                    // this logit is also entry (n, m) of the transposed block, normalised there by col_lse[n]: both softmax terms at once
                    const bool valid = n + j < con.n_valid;
                    const float er = (on_diag && con.row_diag_zero) ? 0.f : __expf(z - lse);
                    const float ec = (!valid || (on_diag && con.col_diag_zero)) ? 0.f : __expf(z - __ldg(con.col_lse + (valid ? n + j : 0)));
                    gz = coef * (er + ec - (on_diag ? con.diag_sub : 0.f));
                  } else {
                    gz = coef * (__expf(z - lse) - (on_diag ? con.diag_sub : 0.f));
                  }
                  if ((on_diag && con.diag_zero) || n + j >= con.n_valid) gz = 0.f;
                  ds_acc += gz * z;
                  if (two_sided && con.diag_exact && on_diag) gz += coef * con.diag_sub;
                  v[j] = gz * alpha;  // dL/d<a_m, b_n>
                }
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(drow + n) = o;
              }
            }
          }
        }
        if (con.dscale != nullptr) {
          ds_acc = warp_sum(ds_acc);
          if (lane == 0) atomicAdd(con.dscale, ds_acc);
        }
      }
      tc_fence_before();
      if (CG == 1 || cta_rank == 0) mbar_arrive(&tmem_empty_bar[acc]);
      else mbar_arrive_cluster(map_to_cta(&tmem_empty_bar[acc], 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // the peer may still be signalling our barriers / reading our smem half
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool A_MN, bool B_MN, int EPI, int CG, int FL = -1>
static int launch_gemm_cg(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream, const CUtensorMap* tmA2 = nullptr,
                          const CUtensorMap* tmB2 = nullptr) {
  auto kern = contrast_pair_kernel<A_MN, B_MN, EPI, CG, FL>;
  constexpr int smem = PairCfg<CG>::SMEM;
  static bool attr_set = false;  // benign race: idempotent attribute
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(gemm smem=%d): %s", smem, cudaGetErrorString(e));
      return B200MM_ERR_LAUNCH;
    }
    attr_set = true;
  }
  const int64_t total_tiles = static_cast<int64_t>(p.m_tiles) * p.n_tiles * p.splits * (EPI == EPI_STD || p.n_prob < 1 ? 1 : p.n_prob);
  const int units = sm_count() / CG;  // CTAs (CG == 1) or CTA pairs
  const int grid = static_cast<int>(total_tiles < units ? total_tiles : units) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmA2 ? *tmA2 : tmA, tmB2 ? *tmB2 : tmB, p);
  if (e != cudaSuccess) {
    set_last_error("contrast_pair_kernel<CG=%d> launch: %s", CG, cudaGetErrorString(e));
    return B200MM_ERR_LAUNCH;
  }
  return check_launch("contrast_pair_kernel");
}

template <bool A_MN, bool B_MN, int EPI = EPI_STD>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream, const CUtensorMap* tmA2 = nullptr,
                       const CUtensorMap* tmB2 = nullptr) {
  return launch_gemm_cg<A_MN, B_MN, EPI, 1>(tmA, tmB, p, stream, tmA2, tmB2);
}

}  // namespace pair
}  // namespace b200mm

using namespace b200mm;
using namespace b200mm::pair;

static int setup_plain(GemmParams& p, CUtensorMap& tmA, CUtensorMap& tmB, const void* a, int64_t lda, const void* b, int64_t ldb,
                       int32_t b_mn, int64_t M, int64_t N, int64_t K, float alpha) {
  B200MM_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), B200MM_ERR_SHAPE,
                 "contrast: M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  B200MM_REQUIRE(a && b, B200MM_ERR_SHAPE, "contrast: null operand");
  p = GemmParams{};
  p.n_prob = 1;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = static_cast<int32_t>(ceil_div(M, BM));
  p.n_tiles = static_cast<int32_t>(ceil_div(N, BN));
  p.kb_total = static_cast<int32_t>(ceil_div(K, BK));
  p.kb_per_split = p.kb_total;
  p.splits = 1;
  p.partial = nullptr;
  p.mn_lbo = SLAB_BYTES; p.mn_sbo = 1024; p.k_lbo = 16; p.k_sbo = 1024;
  p.wait_ns = gemm_wait_ns();
  p.epi.alpha = alpha;
  int rc = make_tmap_2d_bf16(&tmA, a, K, M, lda, BK, BM);
  if (rc) return rc;
  if (!b_mn) return make_tmap_2d_bf16(&tmB, b, K, N, ldb, BK, BN);
  return make_tmap_2d_bf16(&tmB, b, N, K, ldb, 64, BK);  // b stored [K, N] row-major (e.g. the MoCo queue [dim, K_queue])
}

// ---- grouped (two problems per launch) forms for symmetric losses: problem 0 = rows a0 x columns b0, problem 1 = rows a1 x columns b1
static int setup_pair(GemmParams& p, CUtensorMap (&tm)[4], const void* a0, int64_t lda0, const void* b0, int64_t ldb0, const void* a1, int64_t lda1,
                      const void* b1, int64_t ldb1, int64_t M, int64_t N, int64_t K, float alpha) {
  int rc = setup_plain(p, tm[0], tm[1], a0, lda0, b0, ldb0, 0, M, N, K, alpha);
  if (rc) return rc;
  GemmParams q;
  rc = setup_plain(q, tm[2], tm[3], a1, lda1, b1, ldb1, 0, M, N, K, alpha);
  if (rc) return rc;
  p.n_prob = 2;
  return B200MM_OK;
}

extern "C" int b200mm_contrast_lse_partials_pair(const void* a0, int64_t lda0, const void* b0, int64_t ldb0, const void* a1, int64_t lda1,
                                                 const void* b1, int64_t ldb1, int64_t M, int64_t N, int64_t K, float alpha,
                                                 const float* alpha_dev, int64_t diag_off, float* part_max0, float* part_sum0, float* diag0,
                                                 float* part_max1, float* part_sum1, float* diag1, void* stream) {
  GemmParams p;
  CUtensorMap tm[4];
  int rc = setup_pair(p, tm, a0, lda0, b0, ldb0, a1, lda1, b1, ldb1, M, N, K, alpha);
  if (rc) return rc;
  B200MM_REQUIRE(part_max0 && part_sum0 && diag0 && part_max1 && part_sum1 && diag1, B200MM_ERR_SHAPE, "contrast_lse_partials_pair: null output");
  p.con.part_max = part_max0; p.con.part_sum = part_sum0; p.con.diag = diag0; p.con.diag_off = diag_off; p.con.alpha_dev = alpha_dev;
  p.con2 = p.con;
  p.con2.part_max = part_max1; p.con2.part_sum = part_sum1; p.con2.diag = diag1;
  return launch_gemm<false, false, EPI_LSE>(tm[0], tm[1], p, reinterpret_cast<cudaStream_t>(stream), &tm[2], &tm[3]);
}

extern "C" int b200mm_contrast_softgrad_pair(const void* a0, int64_t lda0, const void* b0, int64_t ldb0, const void* a1, int64_t lda1, const void* b1,
                                             int64_t ldb1, int64_t M, int64_t N, int64_t K, int64_t n_valid, float alpha, const float* alpha_dev,
                                             int64_t diag_off, const float* row_lse0, const float* col_lse0, const float* row_lse1,
                                             const float* col_lse1, float coef, const float* coef_dev, float diag_sub, int32_t row_diag_zero0,
                                             int32_t col_diag_zero0, int32_t row_diag_zero1, int32_t col_diag_zero1, void* G0, void* G1, int64_t ldg,
                                             float* dscale, void* stream) {
  GemmParams p;
  CUtensorMap tm[4];
  int rc = setup_pair(p, tm, a0, lda0, b0, ldb0, a1, lda1, b1, ldb1, M, N, K, alpha);
  if (rc) return rc;
  B200MM_REQUIRE(row_lse0 && col_lse0 && row_lse1 && col_lse1 && G0 && G1 && N % 8 == 0 && ldg % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(G0) & 15) == 0 && (reinterpret_cast<uintptr_t>(G1) & 15) == 0,
                 B200MM_ERR_ALIGN, "contrast_softgrad_pair: row / column LSEs required; G 16B aligned, N and ldg multiples of 8");
  p.epi.D = G0; p.epi.ldd = ldg;
  p.con.D = G0; p.con.row_lse = row_lse0; p.con.col_lse = col_lse0; p.con.coef = coef; p.con.coef_dev = coef_dev; p.con.diag_sub = diag_sub;
  p.con.row_diag_zero = row_diag_zero0; p.con.col_diag_zero = col_diag_zero0; p.con.diag_off = diag_off; p.con.dscale = dscale;
  p.con.n_valid = n_valid; p.con.alpha_dev = alpha_dev; p.con.diag_exact = 1;
  p.con2 = p.con;
  p.con2.D = G1; p.con2.row_lse = row_lse1; p.con2.col_lse = col_lse1; p.con2.row_diag_zero = row_diag_zero1; p.con2.col_diag_zero = col_diag_zero1;
  p.con2.dscale = nullptr;  // every logit pair is visited by both problems: the log-temperature gradient is taken from problem 0 only
  return launch_gemm<false, false, EPI_SOFTGRAD>(tm[0], tm[1], p, reinterpret_cast<cudaStream_t>(stream), &tm[2], &tm[3]);
}

