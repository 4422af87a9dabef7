// b200mm — tcgen05 / TMEM attention for head_dim 64 (the ViT-L/14, ViT-B/16 and BERT shapes).
//
// Same arithmetic contract as attention.cu (reference: nn.MultiheadAttention in clip/model.py:245-251 and
// BertSelfAttention.forward, clip/modeling_bert.py:134-172). The legacy mma.sync path tops out near 120 TFLOP/s on
// B200 (measured), so the contractions move to the 5th-gen tensor cores:
//
//   persistent CTA (one per SM) loops over (batch, head) items; per item K and V ([L,64] each) are TMA-loaded once into
//   128B-swizzled smem and the queries are processed in 128-row tiles:
//     S = Q_tile K^T     tcgen05.mma 128 x Lk x 64 (SS), fp32 S in TMEM (Lk <= 320 columns)
//     softmax            16 warps (4 per scheduler to hide TMEM/MUFU latency): warp = (TMEM lane quarter, column partition);
//                        two passes over S with tcgen05.ld, row max/sum combined through smem, P written as bf16 into
//                        K-major swizzled smem tiles
//     O = P V            tcgen05.mma 128 x 64 x Lk, V read MN-major from the same [key][64] smem image
//     epilogue           O * 1/rowsum -> bf16 -> smem transpose -> coalesced stores; LSE per row
//   warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..19 = softmax/epilogue.
//   K/V/Q loads of the next tile/item run ahead through mbarrier rings; QK^T of tile t+1 overlaps the epilogue of t.
#include "common.cuh"

#include <stdlib.h>

#include <algorithm>

namespace b200mm {

constexpr int AT_HD = 64;
constexpr int AT_SM_WARPS = 16;               // softmax warps: TMEM lane quarter = w % 4, column partition = w / 4
constexpr int AT_SM_THREADS = AT_SM_WARPS * 32;
constexpr int AT_THREADS = 128 + AT_SM_THREADS;
constexpr int AT_MAX_LK = 320;
constexpr float AT_LOG2E = 1.4426950408889634f;

struct AttnTcParams {
  __nv_bfloat16* o;
  int64_t ldo;
  float* lse;
  const float* key_bias;
  int32_t B, H, L, Lk;  // Lk = L rounded up to 16
  int32_t q_off, k_off, v_off;
  float scale;
  int32_t short_max;  // tiles with <= short_max valid rows take the replicated-rows path (32, or 0 = off)
};

enum { BAR_K_FULL = 0, BAR_K_EMPTY, BAR_V_FULL, BAR_V_EMPTY, BAR_Q_FULL0, BAR_Q_FULL1, BAR_Q_EMPTY0, BAR_Q_EMPTY1, BAR_S_FULL, BAR_P_FULL,
       BAR_O_FULL, BAR_O_EMPTY, BAR_COUNT };

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Order in which an item's query tiles are processed: the last tile (the short remainder, e.g. 1 row of 257) goes second, so
// that the item ENDS with a full tile whose softmax hides the TMA latency of the next item's K / Q loads.
__device__ __forceinline__ int tile_order(int k, int n) { return n < 3 ? k : (k == 0 ? 0 : (k == 1 ? n - 1 : k - 1)); }

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a [rows][128 B] tile with the 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[BAR_COUNT];
  __shared__ uint32_t tmem_base_smem;

  // 1 KB alignment for the 128B swizzle; offset arithmetic on the __shared__ array keeps the shared address space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int Lk = p.Lk;
  const int kv_bytes = Lk * 128;                     // one [Lk][64] bf16 image
  const int kv_pad = (kv_bytes + 1023) & ~1023;
  const int n_ptiles = (Lk + 63) / 64;
  uint8_t* k_sm = smem;
  uint8_t* v_sm = k_sm + kv_pad;
  uint8_t* q_sm = v_sm + kv_pad;                     // 2 x 16 KB
  uint8_t* p_sm = q_sm + 2 * 16384;                  // n_ptiles x 16 KB
  float* bias_sm = reinterpret_cast<float*>(p_sm + n_ptiles * 16384);  // [Lk] key bias * log2e, -inf for key >= L
  float* red_max = bias_sm + Lk;       // [4][128] per-partition row maxima
  float* red_sum = red_max + 4 * 128;  // [4][128] per-partition row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < BAR_COUNT; ++i) mbar_init(&bars[i], (i == BAR_P_FULL || i == BAR_O_EMPTY) ? AT_SM_THREADS : 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_s = tmem_base;          // S: columns [0, Lk)
  const uint32_t tmem_o = tmem_base + 384;    // O: columns [384, 448)

  const int n_items = p.B * p.H;
  const int n_qt = (p.L + 127) / 128;
  const int n1 = Lk < 256 ? Lk : 256, n2 = Lk - n1;

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t item_cnt = 0, tile_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_cnt) {
      const int b = item / p.H, h = item - b * p.H;
      const int32_t row0 = b * p.L;
      for (int tk = 0; tk < n_qt; ++tk, ++tile_cnt) {
        const int t = tile_order(tk, n_qt);
        if (tk == 0) {
          mbar_wait(&bars[BAR_K_EMPTY], (item_cnt & 1) ^ 1);
          if (lane == 0) {
            mbar_expect_tx(&bars[BAR_K_FULL], kv_bytes);
            tma_load_2d(&tmKV, &bars[BAR_K_FULL], k_sm, p.k_off + h * AT_HD, row0);
            tma_load_2d(&tmKV, &bars[BAR_K_FULL], k_sm + kv_bytes / 2, p.k_off + h * AT_HD, row0 + Lk / 2);
          }
        }
        const int qb = tile_cnt & 1;
        mbar_wait(&bars[BAR_Q_EMPTY0 + qb], ((tile_cnt >> 1) & 1) ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&bars[BAR_Q_FULL0 + qb], 16384);
          tma_load_2d(&tmQ, &bars[BAR_Q_FULL0 + qb], q_sm + qb * 16384, p.q_off + h * AT_HD, row0 + t * 128);
        }
        if (tk == 0) {
          mbar_wait(&bars[BAR_V_EMPTY], (item_cnt & 1) ^ 1);
          if (lane == 0) {
            mbar_expect_tx(&bars[BAR_V_FULL], kv_bytes);
            tma_load_2d(&tmKV, &bars[BAR_V_FULL], v_sm, p.v_off + h * AT_HD, row0);
            tma_load_2d(&tmKV, &bars[BAR_V_FULL], v_sm + kv_bytes / 2, p.v_off + h * AT_HD, row0 + Lk / 2);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s1 = make_idesc_bf16(128, n1, 0, 0);
    const uint32_t idesc_s2 = make_idesc_bf16(128, n2 > 0 ? n2 : 16, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, AT_HD, 0, 1);  // A = P K-major, B = V MN-major
    uint32_t item_cnt = 0, tile_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_cnt) {
      for (int tk = 0; tk < n_qt; ++tk, ++tile_cnt) {
        const int qb = tile_cnt & 1;
        mbar_wait(&bars[BAR_Q_FULL0 + qb], (tile_cnt >> 1) & 1);
        if (tk == 0) mbar_wait(&bars[BAR_K_FULL], item_cnt & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t qa = smem_u32(q_sm + qb * 16384), kb = smem_u32(k_sm);
#pragma unroll
          for (int k = 0; k < AT_HD / 16; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(qa + k * 32, 16, 1024);
            umma_bf16(tmem_s, adesc, make_smem_desc_sw128(kb + k * 32, 16, 1024), idesc_s1, k > 0);
            if (n2 > 0) umma_bf16(tmem_s + 256, adesc, make_smem_desc_sw128(kb + 256 * 128 + k * 32, 16, 1024), idesc_s2, k > 0);
          }
          umma_commit(&bars[BAR_S_FULL]);
          umma_commit(&bars[BAR_Q_EMPTY0 + qb]);
          if (tk == n_qt - 1) umma_commit(&bars[BAR_K_EMPTY]);
        }
        __syncwarp();
        mbar_wait(&bars[BAR_P_FULL], tile_cnt & 1);
        if (tk == 0) mbar_wait(&bars[BAR_V_FULL], item_cnt & 1);
        mbar_wait(&bars[BAR_O_EMPTY], (tile_cnt & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t pa = smem_u32(p_sm), vb = smem_u32(v_sm);
          const int ksteps = Lk / 16;
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t adesc = make_smem_desc_sw128(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(vb + kk * 2048, 8192, 1024);
            umma_bf16(tmem_o, adesc, bdesc, idesc_pv, kk > 0);
          }
          umma_commit(&bars[BAR_O_FULL]);
          if (tk == n_qt - 1) umma_commit(&bars[BAR_V_EMPTY]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue =====================
    // 16 warps: warp sw owns TMEM lanes / query rows [32*(sw%4), +32) and the 16-column chunks cc with cc % 4 == sw / 4.
    // (One warp per scheduler cannot hide the TMEM / MUFU / ALU latencies: 4 per scheduler can.)
    const int sw = warp - 4;
    const int quarter = sw & 3, part = sw >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.scale * AT_LOG2E;
    const int n_chunks = Lk / 16;
    const bool has_bias = p.key_bias != nullptr;
    uint32_t tile_cnt = 0;
    auto sm_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(AT_SM_THREADS) : "memory"); };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / p.H, h = item - b * p.H;
      sm_sync();  // previous item's readers are done with bias_sm
      for (int i = threadIdx.x - 128; i < Lk; i += AT_SM_THREADS)
        bias_sm[i] = i < p.L ? (has_bias ? p.key_bias[static_cast<int64_t>(b) * p.L + i] * AT_LOG2E : 0.f) : -INFINITY;
      sm_sync();
      for (int tk = 0; tk < n_qt; ++tk, ++tile_cnt) {
        const int t = tile_order(tk, n_qt);
        const int q_row = t * 128 + r;
        const bool warp_active = t * 128 + quarter * 32 < p.L;  // warp-uniform: any valid row in this warp
        mbar_wait(&bars[BAR_S_FULL], tile_cnt & 1);
        tc_fence_after();
        // ---- pass 1: partial row maximum of s*c + bias over this warp's chunks
        float mx = -INFINITY;
        if (warp_active) {
          float m0 = -INFINITY, m1 = -INFINITY;
          auto pass1 = [&](const uint32_t (&v)[16], int cc) {
            if (has_bias || cc * 16 + 16 > p.L) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                m0 = fmaxf(m0, fmaf(__uint_as_float(v[j]), c, bias_sm[cc * 16 + j]));
                m1 = fmaxf(m1, fmaf(__uint_as_float(v[j + 1]), c, bias_sm[cc * 16 + j + 1]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                m0 = fmaxf(m0, __uint_as_float(v[j]) * c);
                m1 = fmaxf(m1, __uint_as_float(v[j + 1]) * c);
              }
            }
          };
          for (int cc = part; cc < n_chunks; cc += 4) {
            uint32_t v[16];
            tmem_ld_32x16(tmem_s + lane_addr + cc * 16, v);
            tmem_ld_wait16(v);
            pass1(v, cc);
          }
          red_max[part * 128 + r] = fmaxf(m0, m1);
        }
        sm_sync();
        float sum = 0.f;
        if (warp_active) {
          mx = fmaxf(fmaxf(red_max[r], red_max[128 + r]), fmaxf(red_max[256 + r], red_max[384 + r]));
          // ---- pass 2: p = exp2(s*c + bias - mx) -> bf16 -> K-major swizzled P tiles (this warp's chunks)
          float s0 = 0.f, s1 = 0.f;
          auto pass2 = [&](const uint32_t (&v)[16], int cc) {
            float e[16];
            if (has_bias || cc * 16 + 16 > p.L) {
#pragma unroll
              for (int j = 0; j < 16; ++j) e[j] = fast_exp2(fmaf(__uint_as_float(v[j]), c, bias_sm[cc * 16 + j]) - mx);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) e[j] = fast_exp2(fmaf(__uint_as_float(v[j]), c, -mx));
            }
#pragma unroll
            for (int j = 0; j < 16; j += 2) { s0 += e[j]; s1 += e[j + 1]; }
            uint8_t* ptile = p_sm + (cc >> 2) * 16384;
            const int chunk0 = (cc & 3) * 2;
            uint4 o;
            o.x = pack_bf16x2(e[0], e[1]); o.y = pack_bf16x2(e[2], e[3]);
            o.z = pack_bf16x2(e[4], e[5]); o.w = pack_bf16x2(e[6], e[7]);
            *reinterpret_cast<uint4*>(ptile + sw128_off(r, chunk0)) = o;
            o.x = pack_bf16x2(e[8], e[9]); o.y = pack_bf16x2(e[10], e[11]);
            o.z = pack_bf16x2(e[12], e[13]); o.w = pack_bf16x2(e[14], e[15]);
            *reinterpret_cast<uint4*>(ptile + sw128_off(r, chunk0 + 1)) = o;
          };
          for (int cc = part; cc < n_chunks; cc += 4) {
            uint32_t v[16];
            tmem_ld_32x16(tmem_s + lane_addr + cc * 16, v);
            tmem_ld_wait16(v);
            pass2(v, cc);
          }
          red_sum[part * 128 + r] = s0 + s1;
        }
        // make the generic-proxy smem writes visible to the tensor core (async proxy), release S, publish P
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[BAR_P_FULL]);
        sm_sync();  // red_sum complete
        if (warp_active) sum = (red_sum[r] + red_sum[128 + r]) + (red_sum[256 + r] + red_sum[384 + r]);
        // ---- epilogue: O / sum -> bf16; partition `part` converts O columns [16*part, +16) of its rows, staged in P tile 0
        //      (free once PV has completed), then all 16 warps write the tile out with coalesced 128-byte rows
        mbar_wait(&bars[BAR_O_FULL], tile_cnt & 1);
        tc_fence_after();
        if (warp_active) {
          const float inv = 1.f / sum;
          uint32_t v[16];
          tmem_ld_32x16(tmem_o + lane_addr + part * 16, v);
          tmem_ld_wait16(v);
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
          *reinterpret_cast<uint4*>(p_sm + sw128_off(r, part * 2)) = o;
          o.x = pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
          o.y = pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
          o.z = pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
          o.w = pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
          *reinterpret_cast<uint4*>(p_sm + sw128_off(r, part * 2 + 1)) = o;
          if (part == 0 && q_row < p.L) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.L + q_row] = (mx + log2f(sum)) / AT_LOG2E;
        }
        tc_fence_before();
        mbar_arrive(&bars[BAR_O_EMPTY]);
        sm_sync();  // staged O tile complete
        {
          // 8 lanes x 16 B = one 128-byte output row; warp sw writes rows [8*sw, 8*sw+8)
          __nv_bfloat16* obase = p.o + (static_cast<int64_t>(b) * p.L + t * 128) * p.ldo + h * AT_HD;
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int row = sw * 8 + it * 4 + (lane >> 3);
            const int ch = lane & 7;
            if (t * 128 + row < p.L)
              *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(row) * p.ldo + ch * 8) = *reinterpret_cast<const uint4*>(p_sm + sw128_off(row, ch));
          }
        }
        sm_sync();  // staged tile consumed before the next tile's P stores overwrite it
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================================
// forward, second generation: key-blocked online softmax so that the softmax warps never wait for the tensor core.
//
//   keys are split into block a (first <= 128) and block b (the rest, <= 192); per 128-query tile
//     S_a = Q K_a^T, S_b = Q K_b^T          separate TMEM buffers (columns [0,128) and [128,320))
//     softmax_a: m = rowmax(S_a), P_a = exp2(S_a c - m)                         -> O  = P_a V_a
//     softmax_b: m' = max(m, rowmax(S_b)), O *= exp2(m - m') (in TMEM, skipped when no row of the warp moved), P_b   -> O += P_b V_b
//   S values are loaded from TMEM ONCE and stay in registers between the max and the exp; the four warps that share a
//   row quarter meet on a 128-thread named barrier (they sit on the same scheduler anyway).
//   Schedule of the 16 softmax warps:  a(0) b(0) | a(1) epi(0) b(1) | a(2) epi(1) b(2) ...   while the MMA thread runs
//   PV_a(g) S_a(g+1) PV_b(g) S_b(g+1) behind them; O is double-buffered in TMEM (columns [320,448)). K/V halves are released
//   separately (K_a after the item's last S_a ...), so the next item's K/V arrive one tile period ahead without extra smem.
// =====================================================================================================================
enum { F2_KA_FULL = 0, F2_KA_EMPTY, F2_KB_FULL, F2_KB_EMPTY, F2_VA_FULL, F2_VA_EMPTY, F2_VB_FULL, F2_VB_EMPTY, F2_Q_FULL0, F2_Q_FULL1,
       F2_Q_EMPTY0, F2_Q_EMPTY1, F2_SA_FULL, F2_SB_FULL, F2_PA_FULL, F2_PB_FULL, F2_PA_EMPTY, F2_PB_EMPTY, F2_O_FULL0, F2_O_FULL1,
       F2_O_EMPTY0, F2_O_EMPTY1, F2_COUNT };

constexpr int A2_THREADS = 96 + AT_SM_THREADS;  // warps 0..2 = TMA / MMA / TMEM allocator, warps 3..18 = softmax (19 warps: 104 registers each)

__global__ void __launch_bounds__(A2_THREADS, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ32,
                    const __grid_constant__ CUtensorMap tmKa, const __grid_constant__ CUtensorMap tmKb, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[F2_COUNT];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int Lk = p.Lk;
  const int n_chunks = Lk / 16;
  const int na = n_chunks < 8 ? n_chunks : 8, nb = n_chunks - na;  // 16-key chunks of block a / b
  const int ka_bytes = na * 16 * 128, kb_bytes = nb * 16 * 128;
  const int kv_pad = (Lk * 128 + 1023) & ~1023;
  const int n_ptiles = (Lk + 63) / 64;
  uint8_t* k_sm = smem;
  uint8_t* v_sm = k_sm + kv_pad;
  uint8_t* q_sm = v_sm + kv_pad;                     // 2 x 16 KB
  uint8_t* p_sm = q_sm + 2 * 16384;                  // n_ptiles x 16 KB
  uint8_t* o_sm = p_sm + n_ptiles * 16384;           // 16 KB staging tile of the output epilogue
  float* bias_sm = reinterpret_cast<float*>(o_sm + 16384);  // [2][Lk] key bias * log2e, -inf for key >= L
  float* red_max = bias_sm + 2 * Lk;    // [2][4][128]
  float* red_sum = red_max + 2 * 512;   // [2][4][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKa);
    tma_prefetch_desc(&tmKb);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < F2_COUNT; ++i)
      mbar_init(&bars[i], (i == F2_PA_FULL || i == F2_PB_FULL || i == F2_O_EMPTY0 || i == F2_O_EMPTY1) ? AT_SM_WARPS : 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_sa = tmem_base, tmem_sb = tmem_base + 128, tmem_o0 = tmem_base + 320;  // O buffers at 320 and 384

  const int n_items = p.B * p.H;
  const int n_qt = (p.L + 127) / 128;
  const int my_items = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int G = my_items * n_qt;  // tiles this CTA processes, in order: g -> (item g / n_qt, query tile g % n_qt)

  if (warp == 0) {
    // ===================== TMA producer =====================
    for (int g = 0; g < G; ++g) {
      const int it = g / n_qt, tk = g - it * n_qt;
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / p.H, h = item - b * p.H;
      const int32_t row0 = b * p.L;
      const uint32_t ipar = (it & 1) ^ 1;
      if (tk == 0) {
        mbar_wait_relaxed(&bars[F2_KA_EMPTY], ipar);
        if (lane == 0) {
          mbar_expect_tx(&bars[F2_KA_FULL], ka_bytes);
          tma_load_2d(&tmKa, &bars[F2_KA_FULL], k_sm, p.k_off + h * AT_HD, row0);
        }
        if (nb > 0) {
          mbar_wait_relaxed(&bars[F2_KB_EMPTY], ipar);
          if (lane == 0) {
            mbar_expect_tx(&bars[F2_KB_FULL], kb_bytes);
            tma_load_2d(&tmKb, &bars[F2_KB_FULL], k_sm + ka_bytes, p.k_off + h * AT_HD, row0 + na * 16);
          }
        }
      }
      const int qb = g & 1;
      mbar_wait_relaxed(&bars[F2_Q_EMPTY0 + qb], ((g >> 1) & 1) ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&bars[F2_Q_FULL0 + qb], 16384);
        if (p.L - tk * 128 > p.short_max) {
          tma_load_2d(&tmQ, &bars[F2_Q_FULL0 + qb], q_sm + qb * 16384, p.q_off + h * AT_HD, row0 + tk * 128);
        } else {
          // short tile (<= 32 valid query rows, e.g. the 257th token): the same 32 rows go to all four lane quarters, so that the
          // softmax warps of every scheduler can share the columns of these few rows instead of one quarter doing all the work
#pragma unroll
          for (int rep = 0; rep < 4; ++rep)
            tma_load_2d(&tmQ32, &bars[F2_Q_FULL0 + qb], q_sm + qb * 16384 + rep * 4096, p.q_off + h * AT_HD, row0 + tk * 128);
        }
      }
      if (tk == 0) {
        mbar_wait_relaxed(&bars[F2_VA_EMPTY], ipar);
        if (lane == 0) {
          mbar_expect_tx(&bars[F2_VA_FULL], ka_bytes);
          tma_load_2d(&tmKa, &bars[F2_VA_FULL], v_sm, p.v_off + h * AT_HD, row0);
        }
        if (nb > 0) {
          mbar_wait_relaxed(&bars[F2_VB_EMPTY], ipar);
          if (lane == 0) {
            mbar_expect_tx(&bars[F2_VB_FULL], kb_bytes);
            tma_load_2d(&tmKb, &bars[F2_VB_FULL], v_sm + ka_bytes, p.v_off + h * AT_HD, row0 + na * 16);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_sa = make_idesc_bf16(128, na * 16, 0, 0);
    const uint32_t idesc_sb = make_idesc_bf16(128, nb > 0 ? nb * 16 : 16, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, AT_HD, 0, 1);  // A = P K-major, B = V MN-major
    const uint32_t pa = smem_u32(p_sm), vb = smem_u32(v_sm), kb = smem_u32(k_sm);
    auto issue_s = [&](int g, bool blk_b) {
      const int it = g / n_qt, tk = g - it * n_qt;
      const int qb = g & 1;
      if (!blk_b) mbar_wait_relaxed(&bars[F2_Q_FULL0 + qb], (g >> 1) & 1);
      if (tk == 0) mbar_wait_relaxed(&bars[blk_b ? F2_KB_FULL : F2_KA_FULL], it & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t qa = smem_u32(q_sm + qb * 16384);
        const uint32_t kk0 = blk_b ? kb + ka_bytes : kb;
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          umma_bf16(blk_b ? tmem_sb : tmem_sa, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(kk0 + k * 32, 16, 1024),
                    blk_b ? idesc_sb : idesc_sa, k > 0);
        umma_commit(&bars[blk_b ? F2_SB_FULL : F2_SA_FULL]);
        if (blk_b || nb == 0) umma_commit(&bars[F2_Q_EMPTY0 + qb]);
        if (tk == n_qt - 1) umma_commit(&bars[blk_b ? F2_KB_EMPTY : F2_KA_EMPTY]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int g, bool blk_b) {
      const int it = g / n_qt, tk = g - it * n_qt;
      const int ob = g & 1;
      mbar_wait_relaxed(&bars[blk_b ? F2_PB_FULL : F2_PA_FULL], g & 1);
      if (tk == 0) mbar_wait_relaxed(&bars[blk_b ? F2_VB_FULL : F2_VA_FULL], it & 1);
      if (!blk_b) mbar_wait_relaxed(&bars[F2_O_EMPTY0 + ob], ((g >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const int k0 = blk_b ? na : 0, k1 = blk_b ? n_chunks : na;
        for (int kk = k0; kk < k1; ++kk)
          umma_bf16(tmem_o0 + ob * 64, make_smem_desc_sw128(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(vb + kk * 2048, 8192, 1024), idesc_pv, kk > 0);
        umma_commit(&bars[blk_b ? F2_PB_EMPTY : F2_PA_EMPTY]);
        if (blk_b || nb == 0) umma_commit(&bars[F2_O_FULL0 + ob]);
        if (tk == n_qt - 1) umma_commit(&bars[blk_b ? F2_VB_EMPTY : F2_VA_EMPTY]);
      }
      __syncwarp();
    };
    if (G > 0) {
      issue_s(0, false);
      if (nb > 0) issue_s(0, true);
    }
    for (int g = 0; g < G; ++g) {
      issue_pv(g, false);
      if (g + 1 < G) issue_s(g + 1, false);
      if (nb > 0) {
        issue_pv(g, true);
        if (g + 1 < G) issue_s(g + 1, true);
      }
    }
  } else if (warp >= 3) {
    // ===================== softmax + epilogue =====================
    // TMEM lane quarter = warp % 4 (hardware rule); the four warps of a quarter take the column partitions 0..3
    const int quarter = warp & 3, part = (warp - 3) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.scale * AT_LOG2E;
    const bool has_bias = p.key_bias != nullptr;
    auto full_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(AT_SM_THREADS) : "memory"); };
    auto quarter_sync = [&] { asm volatile("bar.sync %0, 128;" ::"r"(2 + quarter) : "memory"); };
    auto load_bias = [&](int buf, int b) {
      for (int i = threadIdx.x - 96; i < Lk; i += AT_SM_THREADS)
        bias_sm[buf * Lk + i] = i < p.L ? (has_bias ? p.key_bias[static_cast<int64_t>(b) * p.L + i] * AT_LOG2E : 0.f) : -INFINITY;
    };
    if (!has_bias) {
      load_bias(0, 0);
      full_sync();
    }
    uint32_t xcnt = 0;             // max-exchange counter (double-buffered red_max)
    float m_run = 0.f, psum = 0.f;  // running max (scaled, log2 domain) and this partition's share of the row sum
    float m_fin = 0.f, ps_fin = 0.f;

    // one key block of one tile: S chunks of this partition -> registers -> max -> (rescale O) -> P
    auto softmax_block = [&](int g, bool blk_b, bool active, const float* bias) {
      mbar_wait(&bars[blk_b ? F2_SB_FULL : F2_SA_FULL], g & 1);
      tc_fence_after();
      const int nblk = blk_b ? nb : na;
      const int kc0 = blk_b ? na : 0;
      const uint32_t sbase = (blk_b ? tmem_sb : tmem_sa) + lane_addr;
      if (active) {
        uint32_t v[3][16];
        // (a partition with fewer chunks re-reads its last one: unconditional loads keep v[][] in registers)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int ci = part + 4 * i < nblk ? part + 4 * i : part;
          tmem_ld_32x16(sbase + ci * 16, v[i]);
        }
        tmem_ld_wait16(v[0]);
        tmem_ld_wait16(v[1]);
        tmem_ld_wait16(v[2]);
        // ---- partial row maximum (scaled log2 domain)
        float mloc = -INFINITY;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (part + 4 * i < nblk) {
            const int kc = kc0 + part + 4 * i;
            if (has_bias || kc * 16 + 16 > p.L) {
              float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float z0 = fmaf(__uint_as_float(v[i][j]), c, bias[kc * 16 + j]);
                const float z1 = fmaf(__uint_as_float(v[i][j + 1]), c, bias[kc * 16 + j + 1]);
                v[i][j] = __float_as_uint(z0);
                v[i][j + 1] = __float_as_uint(z1);
                m0 = fmaxf(m0, z0);
                m1 = fmaxf(m1, z1);
              }
              mloc = fmaxf(mloc, fmaxf(m0, m1));
            } else {
              float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i][j]), __uint_as_float(v[i][j + 1])));
                m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i][j + 2]), __uint_as_float(v[i][j + 3])));
              }
              mloc = fmaxf(mloc, fmaxf(m0, m1) * c);
            }
          }
        }
        float* rm = red_max + (xcnt & 1) * 512;
        ++xcnt;
        rm[part * 128 + r] = mloc;
        quarter_sync();
        const float bm = fmaxf(fmaxf(rm[r], rm[128 + r]), fmaxf(rm[256 + r], rm[384 + r]));
        float m_new = bm, alpha = 1.f;
        if (blk_b) {
          m_new = fmaxf(m_run, bm);
          alpha = fast_exp2(m_run - m_new);
          psum *= alpha;
        } else {
          psum = 0.f;
        }
        m_run = m_new;
        // ---- P = exp2(z - m) as bf16 into the K-major swizzled tiles; PV of the previous tile must be done with them
        mbar_wait(&bars[blk_b ? F2_PB_EMPTY : F2_PA_EMPTY], (g & 1) ^ 1);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (part + 4 * i < nblk) {
            const int kc = kc0 + part + 4 * i;
            float e[16];
            if (has_bias || kc * 16 + 16 > p.L) {
#pragma unroll
              for (int j = 0; j < 16; ++j) e[j] = fast_exp2(__uint_as_float(v[i][j]) - m_new);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) e[j] = fast_exp2(fmaf(__uint_as_float(v[i][j]), c, -m_new));
            }
#pragma unroll
            for (int j = 0; j < 16; j += 2) { s0 += e[j]; s1 += e[j + 1]; }
            uint8_t* ptile = p_sm + (kc >> 2) * 16384;
            const int chunk0 = (kc & 3) * 2;
            uint4 o;
            o.x = pack_bf16x2(e[0], e[1]); o.y = pack_bf16x2(e[2], e[3]);
            o.z = pack_bf16x2(e[4], e[5]); o.w = pack_bf16x2(e[6], e[7]);
            *reinterpret_cast<uint4*>(ptile + sw128_off(r, chunk0)) = o;
            o.x = pack_bf16x2(e[8], e[9]); o.y = pack_bf16x2(e[10], e[11]);
            o.z = pack_bf16x2(e[12], e[13]); o.w = pack_bf16x2(e[14], e[15]);
            *reinterpret_cast<uint4*>(ptile + sw128_off(r, chunk0 + 1)) = o;
          }
        }
        psum += s0 + s1;
        if (blk_b && !__all_sync(0xffffffffu, alpha == 1.f)) {
          // O holds P_a V_a normalised with the old maximum: bring this partition's 16 columns to the new one
          mbar_wait(&bars[F2_PA_EMPTY], g & 1);  // PV_a(g) has completed
          tc_fence_after();
          uint32_t o[16];
          const uint32_t oaddr = tmem_o0 + (g & 1) * 64 + lane_addr + part * 16;
          tmem_ld_32x16(oaddr, o);
          tmem_ld_wait16(o);
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
          tmem_st_32x16(oaddr, o);
          tmem_st_wait();
        }
      } else {
        mbar_wait(&bars[blk_b ? F2_PB_EMPTY : F2_PA_EMPTY], (g & 1) ^ 1);
      }
      // generic-proxy smem writes -> async proxy; TMEM accesses ordered before the hand-over; one arrival per warp
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[blk_b ? F2_PB_FULL : F2_PA_FULL]);
    };

    // O / rowsum -> bf16 -> global (each thread: its row, this partition's 16 columns = one 32-byte sector), LSE
    auto epilogue = [&](int g, bool active, float m_row, float ps_row) {
      const int it = g / n_qt, tk = g - it * n_qt;
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / p.H, h = item - b * p.H;
      const int ob = g & 1;
      mbar_wait(&bars[F2_O_FULL0 + ob], (g >> 1) & 1);
      tc_fence_after();
      if (active) {
        float* rs = red_sum + ob * 512;
        rs[part * 128 + r] = ps_row;
        quarter_sync();
        const float sum = (rs[r] + rs[128 + r]) + (rs[256 + r] + rs[384 + r]);
        const float inv = 1.f / sum;
        uint32_t v[16];
        tmem_ld_32x16(tmem_o0 + ob * 64 + lane_addr + part * 16, v);
        tmem_ld_wait16(v);
        const int q_row = tk * 128 + r;
        // a thread's 16 columns are only 32 B of its row: the quarter's 32 rows are staged as bf16 and written back by the same
        // four warps as full 128-byte rows (direct 32-byte stores cost 32 lines per warp instruction)
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
        *reinterpret_cast<uint4*>(o_sm + sw128_off(r, part * 2)) = o;
        o.x = pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
        *reinterpret_cast<uint4*>(o_sm + sw128_off(r, part * 2 + 1)) = o;
        if (part == 0 && q_row < p.L) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.L + q_row] = (m_row + log2f(sum)) / AT_LOG2E;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[F2_O_EMPTY0 + ob]);
        quarter_sync();
        __nv_bfloat16* obase = p.o + (static_cast<int64_t>(b) * p.L + tk * 128) * p.ldo + h * AT_HD;
#pragma unroll
        for (int it2 = 0; it2 < 2; ++it2) {
          const int row = quarter * 32 + part * 8 + it2 * 4 + (lane >> 3);
          const int ch = lane & 7;
          if (tk * 128 + row < p.L)
            *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(row) * p.ldo + ch * 8) = *reinterpret_cast<const uint4*>(o_sm + sw128_off(row, ch));
        }
        // (the next use of these staging rows is a full tile later: the quarter meets at two max exchanges in between)
        return;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[F2_O_EMPTY0 + ob]);
    };

    // ---- short tile (<= 32 valid rows, replicated in all four lane quarters): warp wid owns 16-key chunk wid of the block, every
    //      lane = one of the 32 rows; row maxima / sums meet across all 16 warps; P goes to rows 0..31 of the tiles
    const int wid = part * 4 + quarter;
    auto softmax_block_short = [&](int g, bool blk_b, const float* bias) {
      mbar_wait(&bars[blk_b ? F2_SB_FULL : F2_SA_FULL], g & 1);
      tc_fence_after();
      const int nblk = blk_b ? nb : na;
      const int kc = (blk_b ? na : 0) + wid;
      const bool has = wid < nblk;
      uint32_t v[16];
      tmem_ld_32x16((blk_b ? tmem_sb : tmem_sa) + lane_addr + (has ? wid : 0) * 16, v);
      tmem_ld_wait16(v);
      float mloc = -INFINITY;
      if (has) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float z = fmaf(__uint_as_float(v[j]), c, bias[kc * 16 + j]);
          v[j] = __float_as_uint(z);
          mloc = fmaxf(mloc, z);
        }
      }
      float* rm = red_max + (xcnt & 1) * 512;
      ++xcnt;
      rm[wid * 32 + lane] = mloc;
      full_sync();
      float bm = rm[lane];
#pragma unroll
      for (int w = 1; w < 16; ++w) bm = fmaxf(bm, rm[w * 32 + lane]);
      float m_new = bm, alpha = 1.f;
      if (blk_b) {
        m_new = fmaxf(m_run, bm);
        alpha = fast_exp2(m_run - m_new);
        psum *= alpha;
      } else {
        psum = 0.f;
      }
      m_run = m_new;
      mbar_wait(&bars[blk_b ? F2_PB_EMPTY : F2_PA_EMPTY], (g & 1) ^ 1);
      if (has) {
        float e[16], s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) e[j] = fast_exp2(__uint_as_float(v[j]) - m_new);
#pragma unroll
        for (int j = 0; j < 16; j += 2) { s0 += e[j]; s1 += e[j + 1]; }
        psum += s0 + s1;
        uint8_t* ptile = p_sm + (kc >> 2) * 16384;
        const int chunk0 = (kc & 3) * 2;
        uint4 o;
        o.x = pack_bf16x2(e[0], e[1]); o.y = pack_bf16x2(e[2], e[3]);
        o.z = pack_bf16x2(e[4], e[5]); o.w = pack_bf16x2(e[6], e[7]);
        *reinterpret_cast<uint4*>(ptile + sw128_off(lane, chunk0)) = o;
        o.x = pack_bf16x2(e[8], e[9]); o.y = pack_bf16x2(e[10], e[11]);
        o.z = pack_bf16x2(e[12], e[13]); o.w = pack_bf16x2(e[14], e[15]);
        *reinterpret_cast<uint4*>(ptile + sw128_off(lane, chunk0 + 1)) = o;
      }
      if (blk_b && quarter == 0 && !__all_sync(0xffffffffu, alpha == 1.f)) {
        mbar_wait(&bars[F2_PA_EMPTY], g & 1);  // PV_a(g) has completed
        tc_fence_after();
        uint32_t o[16];
        const uint32_t oaddr = tmem_o0 + (g & 1) * 64 + part * 16;  // lanes 0..31
        tmem_ld_32x16(oaddr, o);
        tmem_ld_wait16(o);
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
        tmem_st_32x16(oaddr, o);
        tmem_st_wait();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[blk_b ? F2_PB_FULL : F2_PA_FULL]);
    };
    auto epilogue_short = [&](int g, float m_row, float ps_row) {
      const int it = g / n_qt, tk = g - it * n_qt;
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / p.H, h = item - b * p.H;
      const int ob = g & 1;
      mbar_wait(&bars[F2_O_FULL0 + ob], (g >> 1) & 1);
      tc_fence_after();
      float* rs = red_sum + ob * 512;
      rs[wid * 32 + lane] = ps_row;
      full_sync();
      if (quarter == 0) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) sum += rs[w * 32 + lane];
        const float inv = 1.f / sum;
        uint32_t v[16];
        tmem_ld_32x16(tmem_o0 + ob * 64 + part * 16, v);
        tmem_ld_wait16(v);
        const int q_row = tk * 128 + lane;
        if (q_row < p.L) {
          __nv_bfloat16* orow = p.o + (static_cast<int64_t>(b) * p.L + q_row) * p.ldo + h * AT_HD + part * 16;
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
          *reinterpret_cast<uint4*>(orow) = o;
          o.x = pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
          o.y = pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
          o.z = pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
          o.w = pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
          *reinterpret_cast<uint4*>(orow + 8) = o;
          if (part == 0) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.L + q_row] = (m_row + log2f(sum)) / AT_LOG2E;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[F2_O_EMPTY0 + ob]);
    };

    bool prev_active = false, prev_short = false;
    for (int g = 0; g < G; ++g) {
      const int it = g / n_qt, tk = g - it * n_qt;
      const bool short_tile = p.L - tk * 128 <= p.short_max;
      const bool active = tk * 128 + quarter * 32 < p.L;  // warp-uniform: any valid query row in this warp
      const float* bias = bias_sm;
      if (has_bias) {
        if (tk == 0) {
          const int item = blockIdx.x + it * gridDim.x;
          load_bias(it & 1, item / p.H);
          full_sync();
        }
        bias = bias_sm + (it & 1) * Lk;
      }
      if (short_tile) softmax_block_short(g, false, bias);
      else softmax_block(g, false, active, bias);
      if (g > 0) {
        if (prev_short) epilogue_short(g - 1, m_fin, ps_fin);
        else epilogue(g - 1, prev_active, m_fin, ps_fin);
      }
      if (nb > 0) {
        if (short_tile) softmax_block_short(g, true, bias);
        else softmax_block(g, true, active, bias);
      }
      m_fin = m_run;
      ps_fin = psum;
      prev_active = active;
      prev_short = short_tile;
    }
    if (G > 0) {
      if (prev_short) epilogue_short(G - 1, m_fin, ps_fin);
      else epilogue(G - 1, prev_active, m_fin, ps_fin);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static size_t attn_tc2_smem_bytes(int Lk) {
  const int kv_pad = (Lk * 128 + 1023) & ~1023;
  const int n_ptiles = (Lk + 63) / 64;
  return static_cast<size_t>(2) * kv_pad + 2 * 16384 + static_cast<size_t>(n_ptiles + 1) * 16384 + 2 * Lk * 4 + 4 * 512 * 4 + 1024;
}

static size_t attn_tc_smem_bytes(int Lk) {
  const int kv_pad = (Lk * 128 + 1023) & ~1023;
  const int n_ptiles = (Lk + 63) / 64;
  return static_cast<size_t>(2) * kv_pad + 2 * 16384 + static_cast<size_t>(n_ptiles) * 16384 + Lk * 4 + 2 * 4 * 128 * 4 + 1024;
}

// returns 1 if handled, 0 if the shape is not supported by the tcgen05 path (caller falls back), <0 on error
int attention_fwd_tc(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, void* o, int64_t ldo, float* lse,
                     const float* key_bias, int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale, cudaStream_t stream) {
  const int Lk = (L + 15) & ~15;
  if (head_dim != AT_HD || Lk > AT_MAX_LK || Lk < 16) return 0;
  if (getenv("B200MM_ATTN_LEGACY")) return 0;
  const int64_t T = static_cast<int64_t>(B) * L;
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_2d_bf16(&tmQ, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmKV, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, Lk / 2);
  if (rc) return rc;
  AttnTcParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.ldo = ldo; p.lse = lse; p.key_bias = key_bias;
  p.B = B; p.H = H; p.L = L; p.Lk = Lk; p.q_off = q_off; p.k_off = k_off; p.v_off = v_off; p.scale = scale;
  p.short_max = getenv("B200MM_ATTN_NOSHORT") ? 0 : 32;
  const int grid = std::min(B * H, sm_count());
  // sequences of <= 128 keys are a single key block: nothing to pipeline, the first-generation kernel (one S buffer, two passes) is
  // measured faster there (BERT, L = 77: 0.065 vs 0.074 ms at 256 x 12 heads); longer ones take the key-blocked kernel
  if (!getenv("B200MM_ATTN_FWD_V1") && Lk > 128) {
    const int n_chunks = Lk / 16, na = n_chunks < 8 ? n_chunks : 8, nb = n_chunks - na;
    CUtensorMap tmKa, tmKb;
    rc = make_tmap_2d_bf16(&tmKa, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, na * 16);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tmKb, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, nb > 0 ? nb * 16 : 16);
    if (rc) return rc;
    const size_t smem2 = attn_tc2_smem_bytes(Lk);
    cudaError_t e2 = cudaFuncSetAttribute(attn_fwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2));
    if (e2 != cudaSuccess) {
      set_last_error("attention_fwd_tc2: cudaFuncSetAttribute(%zu): %s", smem2, cudaGetErrorString(e2));
      return B200MM_ERR_LAUNCH;
    }
    CUtensorMap tmQ32;
    rc = make_tmap_2d_bf16(&tmQ32, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, 32);
    if (rc) return rc;
    attn_fwd_tc2_kernel<<<grid, A2_THREADS, smem2, stream>>>(tmQ, tmQ32, tmKa, tmKb, p);
    rc = check_launch("attn_fwd_tc2_kernel");
    return rc ? rc : 1;
  }
  const size_t smem = attn_tc_smem_bytes(Lk);
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) {
    set_last_error("attention_fwd_tc: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return B200MM_ERR_LAUNCH;
  }
  attn_fwd_tc_kernel<<<grid, AT_THREADS, smem, stream>>>(tmQ, tmKV, p);
  rc = check_launch("attn_fwd_tc_kernel");
  return rc ? rc : 1;
}


// =====================================================================================================================
// backward: one kernel per (batch, head) item producing dQ, dK, dV (no recomputation pass, no atomics)
//
//   key tiles j (128 keys = MMA M / TMEM lanes)  x  query chunks i (<= 96 queries = TMEM columns):
//     S^T_ji = K_j Q_i^T,  dP^T_ji = V_j dO_i^T                      tcgen05, fp32 in TMEM
//     P^T = exp2(S^T c + bias_key - lse_q),  dS^T = P^T (dP^T - D_q) scale     16 elementwise warps, bf16 tiles in smem
//     dV_j += P^T dO_i,  dK_j += dS^T Q_i                               A = the K-major smem tiles
//     dQ_i += dS_ji K_j                                                 A = the SAME dS^T tile read MN-major (transposed view)
//   TMEM (512 columns, exactly full): S^T 96 | dP^T 96 | dV 64 | dK 64 | dQ_0..2 3x64.
//   Q and dO of the item stay resident in smem ([Lq][64] images, used K-major for S^T/dP^T and MN-major for dK/dV);
//   K_j / V_j tiles are double-buffered.
// =====================================================================================================================
constexpr int AB_CW = 96;  // query-chunk width (TMEM columns of S^T / dP^T)

struct AttnBwdTcParams {
  const float* lse;
  const float* dsum;
  const float* key_bias;
  __nv_bfloat16* dqkv;
  int64_t ld;
  int32_t B, H, L, Lq;
  int32_t q_off, k_off, v_off;
  float scale;
};

enum { BB_QD_FULL = 0, BB_QD_EMPTY, BB_KV_FULL0, BB_KV_FULL1, BB_KV_EMPTY0, BB_KV_EMPTY1, BB_SD_FULL, BB_SD_EMPTY, BB_PDS_FULL, BB_PDS_EMPTY,
       BB_DKV_FULL, BB_DKV_EMPTY, BB_DQ_FULL, BB_DQ_EMPTY, BB_COUNT };

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQrows, const __grid_constant__ CUtensorMap tmQtile,
                   const __grid_constant__ CUtensorMap tmDOrows, const AttnBwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[BB_COUNT];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int Lq = p.Lq;
  const int img_bytes = Lq * 128;
  const int img_pad = (img_bytes + 1023) & ~1023;
  uint8_t* q_sm = smem;                    // [Lq][64] Q image
  uint8_t* do_sm = q_sm + img_pad;         // [Lq][64] dO image
  uint8_t* k_sm = do_sm + img_pad;         // 2 x [128][64]
  uint8_t* v_sm = k_sm + 2 * 16384;        // 2 x [128][64]
  uint8_t* pt_sm = v_sm + 2 * 16384;       // P^T  : 2 sub-tiles [128 keys][64 queries]
  uint8_t* ds_sm = pt_sm + 2 * 16384;      // dS^T : 2 sub-tiles
  float* lse_sm = reinterpret_cast<float*>(ds_sm + 2 * 16384);  // [Lq] lse * log2e (+inf for q >= L)
  float* d_sm = lse_sm + Lq;                                      // [Lq] D (0 for q >= L)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQrows);
    tma_prefetch_desc(&tmQtile);
    tma_prefetch_desc(&tmDOrows);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < BB_COUNT; ++i)
      mbar_init(&bars[i], (i == BB_SD_EMPTY || i == BB_PDS_FULL || i == BB_DKV_EMPTY || i == BB_DQ_EMPTY) ? AT_SM_WARPS : 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 96, tm_dv = tmem_base + 192, tm_dk = tmem_base + 256, tm_dq = tmem_base + 320;

  const int n_items = p.B * p.H;
  const int n_kt = (p.L + 127) / 128;
  const int n_qc = (Lq + AB_CW - 1) / AB_CW;  // <= 3

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t item_cnt = 0, tile_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_cnt) {
      const int b = item / p.H, h = item - b * p.H;
      const int32_t row0 = b * p.L;
      mbar_wait(&bars[BB_QD_EMPTY], (item_cnt & 1) ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&bars[BB_QD_FULL], 2 * img_bytes);
        tma_load_2d(&tmQrows, &bars[BB_QD_FULL], q_sm, p.q_off + h * AT_HD, row0);
        tma_load_2d(&tmQrows, &bars[BB_QD_FULL], q_sm + img_bytes / 2, p.q_off + h * AT_HD, row0 + Lq / 2);
        tma_load_2d(&tmDOrows, &bars[BB_QD_FULL], do_sm, h * AT_HD, row0);
        tma_load_2d(&tmDOrows, &bars[BB_QD_FULL], do_sm + img_bytes / 2, h * AT_HD, row0 + Lq / 2);
      }
      __syncwarp();
      for (int j = 0; j < n_kt; ++j, ++tile_cnt) {
        const int jb = tile_cnt & 1;
        mbar_wait(&bars[BB_KV_EMPTY0 + jb], ((tile_cnt >> 1) & 1) ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&bars[BB_KV_FULL0 + jb], 2 * 16384);
          tma_load_2d(&tmQtile, &bars[BB_KV_FULL0 + jb], k_sm + jb * 16384, p.k_off + h * AT_HD, row0 + j * 128);
          tma_load_2d(&tmQtile, &bars[BB_KV_FULL0 + jb], v_sm + jb * 16384, p.v_off + h * AT_HD, row0 + j * 128);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Per item the steps s = (key tile j, query chunk i) run as a software pipeline: S^T/dP^T of step s+1 are issued as soon as the
    // elementwise warps have pulled step s out of TMEM (SD_EMPTY), i.e. BEFORE the accumulating MMAs of step s, so the tensor pipe
    // works on dV/dK/dQ(s) and S/dP(s+1) while the elementwise warps are busy with the exponentials of step s / s+1.
    const uint32_t idesc_acc = make_idesc_bf16(128, AT_HD, 0, 1);  // dV / dK : A K-major (P^T, dS^T), B MN-major (dO, Q)
    const uint32_t idesc_dq = make_idesc_bf16(128, AT_HD, 1, 1);   // dQ      : A = dS^T read MN-major, B = K_j MN-major
    const int n_steps = n_kt * n_qc;
    uint32_t item_cnt = 0, tile_cnt0 = 0, step_cnt0 = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_cnt, tile_cnt0 += n_kt, step_cnt0 += n_steps) {
      mbar_wait_relaxed(&bars[BB_QD_FULL], item_cnt & 1);
      mbar_wait_relaxed(&bars[BB_DQ_EMPTY], (item_cnt & 1) ^ 1);  // previous item's dQ accumulators have been read out
      auto issue_sd = [&](int s) {
        const int j = s / n_qc, i = s - j * n_qc;
        const uint32_t tile_cnt = tile_cnt0 + j;
        const int jb = tile_cnt & 1;
        if (i == 0) mbar_wait_relaxed(&bars[BB_KV_FULL0 + jb], (tile_cnt >> 1) & 1);
        const int q0 = i * AB_CW;
        const int w = min(AB_CW, Lq - q0);
        const uint32_t idesc_sd = make_idesc_bf16(128, w, 0, 0);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t ka = smem_u32(k_sm + jb * 16384), va = smem_u32(v_sm + jb * 16384);
          const uint32_t qb = smem_u32(q_sm) + q0 * 128, dob = smem_u32(do_sm) + q0 * 128;
#pragma unroll
          for (int k = 0; k < AT_HD / 16; ++k) {
            umma_bf16(tm_s, make_smem_desc_sw128(ka + k * 32, 16, 1024), make_smem_desc_sw128(qb + k * 32, 16, 1024), idesc_sd, k > 0);
            umma_bf16(tm_dp, make_smem_desc_sw128(va + k * 32, 16, 1024), make_smem_desc_sw128(dob + k * 32, 16, 1024), idesc_sd, k > 0);
          }
          umma_commit(&bars[BB_SD_FULL]);
        }
        __syncwarp();
      };
      issue_sd(0);
      for (int s = 0; s < n_steps; ++s) {
        const int j = s / n_qc, i = s - j * n_qc;
        const uint32_t tile_cnt = tile_cnt0 + j, step_cnt = step_cnt0 + s;
        const int jb = tile_cnt & 1;
        if (s + 1 < n_steps) {
          mbar_wait_relaxed(&bars[BB_SD_EMPTY], step_cnt & 1);  // S^T/dP^T of step s are in registers: the TMEM buffers are free
          issue_sd(s + 1);
        }
        const int q0 = i * AB_CW;
        const int w = min(AB_CW, Lq - q0);
        mbar_wait_relaxed(&bars[BB_PDS_FULL], step_cnt & 1);
        if (i == 0) mbar_wait_relaxed(&bars[BB_DKV_EMPTY], (tile_cnt & 1) ^ 1);  // previous tile's dV/dK have been read out
        tc_fence_after();
        if (lane == 0) {
          const uint32_t ka = smem_u32(k_sm + jb * 16384);
          const uint32_t pa = smem_u32(pt_sm), dsa = smem_u32(ds_sm);
          const uint32_t qb = smem_u32(q_sm) + q0 * 128, dob = smem_u32(do_sm) + q0 * 128;
          for (int kk = 0; kk < w / 16; ++kk) {
            const uint32_t aoff = (kk >> 2) * 16384 + (kk & 3) * 32;
            umma_bf16(tm_dv, make_smem_desc_sw128(pa + aoff, 16, 1024), make_smem_desc_sw128(dob + kk * 2048, 8192, 1024), idesc_acc,
                      (i > 0 || kk > 0));
            umma_bf16(tm_dk, make_smem_desc_sw128(dsa + aoff, 16, 1024), make_smem_desc_sw128(qb + kk * 2048, 8192, 1024), idesc_acc,
                      (i > 0 || kk > 0));
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // reduction over the 128 keys of tile j
            umma_bf16(tm_dq + i * 64, make_smem_desc_sw128(dsa + kk * 2048, 16384, 1024), make_smem_desc_sw128(ka + kk * 2048, 8192, 1024),
                      idesc_dq, (j > 0 || kk > 0));
          umma_commit(&bars[BB_PDS_EMPTY]);
          if (i == n_qc - 1) {
            umma_commit(&bars[BB_DKV_FULL]);
            umma_commit(&bars[BB_KV_EMPTY0 + jb]);
            if (j == n_kt - 1) {
              umma_commit(&bars[BB_DQ_FULL]);
              umma_commit(&bars[BB_QD_EMPTY]);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== elementwise + epilogues (16 warps) =====================
    const int sw = warp - 4;
    const int quarter = sw & 3, part = sw >> 2;
    const int r = quarter * 32 + lane;  // key row inside the tile / TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float c = p.scale * AT_LOG2E;
    uint32_t item_cnt = 0, tile_cnt = 0, step_cnt = 0;
    auto sm_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(AT_SM_THREADS) : "memory"); };
    // staged-tile writer: 16 fp32 accumulator columns [16*part, +16) of row r -> bf16 into a swizzled [128][128 B] tile
    auto stage16 = [&](uint8_t* tile, const uint32_t (&v)[16]) {
      uint4 o;
      o.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1])); o.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
      o.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5])); o.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
      *reinterpret_cast<uint4*>(tile + sw128_off(r, part * 2)) = o;
      o.x = pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9])); o.y = pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11]));
      o.z = pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13])); o.w = pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15]));
      *reinterpret_cast<uint4*>(tile + sw128_off(r, part * 2 + 1)) = o;
    };
    // coalesced write-out of a staged tile: warp sw writes rows [8*sw, +8), 8 lanes x 16 B per 128-byte row
    auto write_tile = [&](const uint8_t* tile, __nv_bfloat16* gbase, int rows_valid) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int row = sw * 8 + it * 4 + (lane >> 3);
        const int ch = lane & 7;
        if (row < rows_valid)
          *reinterpret_cast<uint4*>(gbase + static_cast<int64_t>(row) * p.ld + ch * 8) = *reinterpret_cast<const uint4*>(tile + sw128_off(row, ch));
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_cnt) {
      const int b = item / p.H, h = item - b * p.H;
      const int64_t bh = static_cast<int64_t>(b) * p.H + h;
      sm_sync();  // previous item's readers are done with lse_sm / d_sm
      for (int i = threadIdx.x - 128; i < Lq; i += AT_SM_THREADS) {
        lse_sm[i] = i < p.L ? p.lse[bh * p.L + i] * AT_LOG2E : INFINITY;
        d_sm[i] = i < p.L ? p.dsum[bh * p.L + i] : 0.f;
      }
      sm_sync();
      for (int j = 0; j < n_kt; ++j, ++tile_cnt) {
        const int key = j * 128 + r;
        const bool warp_active = j * 128 + quarter * 32 < p.L;
        const float kb = key < p.L ? (p.key_bias ? p.key_bias[static_cast<int64_t>(b) * p.L + key] * AT_LOG2E : 0.f) : -INFINITY;
        for (int i = 0; i < n_qc; ++i, ++step_cnt) {
          const int q0 = i * AB_CW;
          const int w = min(AB_CW, Lq - q0);
          mbar_wait(&bars[BB_SD_FULL], step_cnt & 1);
          tc_fence_after();
          // this partition's columns of the chunk, in units of 8 queries (balanced over the 4 partitions: <= 3 units each)
          const int units = w / 8, ubase = units >> 2, urem = units & 3;
          const int my_units = ubase + (part < urem ? 1 : 0);
          const int u0 = part * ubase + (part < urem ? part : urem);
          uint4 pk[3], dk[3];
          if (warp_active) {
            uint32_t sv[3][8], dv[3][8];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const int u = k < my_units ? u0 + k : u0;  // (unconditional loads keep the arrays in registers)
              tmem_ld_32x8(tm_s + lane_addr + u * 8, sv[k]);
              tmem_ld_32x8(tm_dp + lane_addr + u * 8, dv[k]);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              tmem_ld_wait8(sv[k]);
              tmem_ld_wait8(dv[k]);
            }
            // S^T / dP^T of this step now live in registers: hand the TMEM buffers back to the tensor core
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BB_SD_EMPTY]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (k < my_units) {
                const int q = q0 + (u0 + k) * 8;
                const float4 l0 = *reinterpret_cast<const float4*>(lse_sm + q), l1 = *reinterpret_cast<const float4*>(lse_sm + q + 4);
                const float4 d0 = *reinterpret_cast<const float4*>(d_sm + q), d1 = *reinterpret_cast<const float4*>(d_sm + q + 4);
                const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                float pe[8], de[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                  pe[x] = fast_exp2(fmaf(__uint_as_float(sv[k][x]), c, kb) - ls[x]);
                  de[x] = pe[x] * (__uint_as_float(dv[k][x]) - dd[x]) * p.scale;
                }
                pk[k] = make_uint4(pack_bf16x2(pe[0], pe[1]), pack_bf16x2(pe[2], pe[3]), pack_bf16x2(pe[4], pe[5]), pack_bf16x2(pe[6], pe[7]));
                dk[k] = make_uint4(pack_bf16x2(de[0], de[1]), pack_bf16x2(de[2], de[3]), pack_bf16x2(de[4], de[5]), pack_bf16x2(de[6], de[7]));
              }
            }
          } else {
            // keys >= L: rows of dS^T must be exactly zero (they are reduced over in dQ); P^T rows only feed unused dV rows
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BB_SD_EMPTY]);
#pragma unroll
            for (int k = 0; k < 3; ++k) pk[k] = dk[k] = make_uint4(0u, 0u, 0u, 0u);
          }
          // the accumulating MMAs of the previous step must be done with the P^T / dS^T tiles before they are overwritten
          mbar_wait(&bars[BB_PDS_EMPTY], (step_cnt & 1) ^ 1);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            if (k < my_units) {
              const int col = (u0 + k) * 8;
              const uint32_t off = (col >> 6) * 16384 + sw128_off(r, (col & 63) >> 3);
              *reinterpret_cast<uint4*>(pt_sm + off) = pk[k];
              *reinterpret_cast<uint4*>(ds_sm + off) = dk[k];
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BB_PDS_FULL]);
        }
        // ---- tile epilogue: dV_j, dK_j -> bf16 rows of dqkv (staged through the now idle P^T / dS^T sub-tile 0)
        mbar_wait(&bars[BB_DKV_FULL], tile_cnt & 1);
        tc_fence_after();
        if (warp_active) {
          uint32_t a[16], bq[16];
          tmem_ld_32x16(tm_dv + lane_addr + part * 16, a);
          tmem_ld_32x16(tm_dk + lane_addr + part * 16, bq);
          tmem_ld_wait16(a);
          tmem_ld_wait16(bq);
          stage16(pt_sm, a);
          stage16(ds_sm, bq);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BB_DKV_EMPTY]);
        sm_sync();
        {
          __nv_bfloat16* gk = p.dqkv + (static_cast<int64_t>(b) * p.L + j * 128) * p.ld + h * AT_HD;
          const int rows_valid = min(128, p.L - j * 128);
          write_tile(pt_sm, gk + p.v_off, rows_valid);
          write_tile(ds_sm, gk + p.k_off, rows_valid);
        }
        sm_sync();
      }
      // ---- item epilogue: dQ chunks (TMEM lane = query index inside the chunk)
      mbar_wait(&bars[BB_DQ_FULL], item_cnt & 1);
      tc_fence_after();
      for (int i = 0; i < n_qc; ++i) {
        const int q0 = i * AB_CW;
        const int w = min(AB_CW, Lq - q0);
        if (quarter * 32 < w) {
          uint32_t a[16];
          tmem_ld_32x16(tm_dq + i * 64 + lane_addr + part * 16, a);
          tmem_ld_wait16(a);
          stage16(pt_sm, a);
        }
        if (i == n_qc - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BB_DQ_EMPTY]);
        }
        sm_sync();
        write_tile(pt_sm, p.dqkv + (static_cast<int64_t>(b) * p.L + q0) * p.ld + p.q_off + h * AT_HD, min(w, p.L - q0));
        sm_sync();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// D[b,h,l] = sum_c dO[b,l,h*64+c] * O[b,l,h*64+c]; one warp per token row, 8 lanes per head
__global__ void __launch_bounds__(256) attn_dsum_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int64_t ldo,
                                                        float* __restrict__ dsum, int32_t B, int32_t H, int32_t L) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= static_cast<int64_t>(B) * L) return;
  const int64_t b = row / L;
  const int l = static_cast<int>(row - b * L);
  for (int h0 = 0; h0 < H; h0 += 4) {
    const int h = h0 + (lane >> 3);
    float acc = 0.f;
    if (h < H) {
      const int64_t off = row * ldo + h * AT_HD + (lane & 7) * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(o + off);
      const uint4 g = *reinterpret_cast<const uint4*>(d_o + off);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x = unpack_bf16x2(aw[k]), y = unpack_bf16x2(gw[k]);
        acc += x.x * y.x + x.y * y.y;
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (h < H && (lane & 7) == 0) dsum[(b * H + h) * L + l] = acc;
  }
}

int attention_bwd_tc(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* o, const void* d_o, int64_t ldo,
                     const float* lse, const float* key_bias, void* dqkv, float* dsum, int32_t B, int32_t H, int32_t L, int32_t head_dim,
                     float scale, cudaStream_t stream) {
  const int Lq = (L + 15) & ~15;
  if (head_dim != AT_HD || Lq > 3 * AB_CW || Lq < 16) return 0;
  if (getenv("B200MM_ATTN_LEGACY") || getenv("B200MM_ATTN_BWD_LEGACY")) return 0;
  const int64_t T = static_cast<int64_t>(B) * L;
  attn_dsum_kernel<<<static_cast<int>(ceil_div(T, 8)), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(o),
                                                                        reinterpret_cast<const __nv_bfloat16*>(d_o), ldo, dsum, B, H, L);
  int rc = check_launch("attn_dsum_kernel");
  if (rc) return rc;
  CUtensorMap tmQrows, tmQtile, tmDOrows;
  rc = make_tmap_2d_bf16(&tmQrows, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, Lq / 2);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmQtile, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmDOrows, d_o, static_cast<uint64_t>(ldo), static_cast<uint64_t>(T), static_cast<uint64_t>(ldo), 64, Lq / 2);
  if (rc) return rc;
  AttnBwdTcParams p;
  p.lse = lse; p.dsum = dsum; p.key_bias = key_bias; p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv); p.ld = ld;
  p.B = B; p.H = H; p.L = L; p.Lq = Lq; p.q_off = q_off; p.k_off = k_off; p.v_off = v_off; p.scale = scale;
  const int img_pad = (Lq * 128 + 1023) & ~1023;
  const size_t smem = static_cast<size_t>(2) * img_pad + 8 * 16384 + static_cast<size_t>(2) * Lq * 4 + 1024;
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) {
    set_last_error("attention_bwd_tc: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return B200MM_ERR_LAUNCH;
  }
  const int grid = std::min(B * H, sm_count());
  attn_bwd_tc_kernel<<<grid, AT_THREADS, smem, stream>>>(tmQrows, tmQtile, tmDOrows, p);
  rc = check_launch("attn_bwd_tc_kernel");
  return rc ? rc : 1;
}

}  // namespace b200mm
