// b200mm — merged tcgen05 / TMEM attention backward for head_dim 64 and sequences of up to 288 tokens (ViT-B/16, ViT-L/14, BERT, the
// cross-modal stage): dQ, dK and dV of a (batch, head) item in ONE kernel, no recomputation pass and no atomics, because all three
// accumulators of the item fit the 512 TMEM columns. Every other shape takes the recompute kernels of attention_v3.cu.
//
// Reference arithmetic: autograd of nn.MultiheadAttention (clip/model.py:245-251) / BertSelfAttention.forward (clip/modeling_bert.py:134-172).
#include "common.cuh"

#include <stdlib.h>

#include <algorithm>

namespace b200mm {

constexpr int AT_HD = 64;
constexpr int AT_SM_WARPS = 16;               // elementwise warps: TMEM lane quarter = w % 4, column partition = w / 4
constexpr int AT_SM_THREADS = AT_SM_WARPS * 32;
constexpr int AT_THREADS = 128 + AT_SM_THREADS;
constexpr float AT_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ bool at_elect_one() {
  uint32_t q;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(q));
  return q != 0;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a [rows][128 B] tile with the 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }


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
    // Issue is done by ONE ELECTED lane (elect.sync) over warp-uniform operands, so every tcgen05.mma / commit compiles to a single
    // uniform-datapath instruction; under a plain `lane == 0` predicate ptxas wraps each of the 28 MMAs of a step in a vote / broadcast
    // loop (~25 instructions), which made this warp - not the tensor pipe - the bottleneck of the kernel. Descriptors are base + offset.
    const uint32_t idesc_acc = make_idesc_bf16(128, AT_HD, 0, 1);  // dV / dK : A K-major (P^T, dS^T), B MN-major (dO, Q)
    const uint32_t idesc_dq = make_idesc_bf16(128, AT_HD, 1, 1);   // dQ      : A = dS^T read MN-major, B = K_j MN-major
    const uint64_t kd0 = make_smem_desc_sw128(smem_u32(k_sm), 16, 1024), vd0 = make_smem_desc_sw128(smem_u32(v_sm), 16, 1024);
    const uint64_t qd0 = make_smem_desc_sw128(smem_u32(q_sm), 16, 1024), dod0 = make_smem_desc_sw128(smem_u32(do_sm), 16, 1024);
    const uint64_t ptd0 = make_smem_desc_sw128(smem_u32(pt_sm), 16, 1024), dsd0 = make_smem_desc_sw128(smem_u32(ds_sm), 16, 1024);
    const uint64_t qmn0 = make_smem_desc_sw128(smem_u32(q_sm), 8192, 1024), domn0 = make_smem_desc_sw128(smem_u32(do_sm), 8192, 1024);
    const uint64_t dsmn0 = make_smem_desc_sw128(smem_u32(ds_sm), 16384, 1024), kmn0 = make_smem_desc_sw128(smem_u32(k_sm), 8192, 1024);
    const int n_steps = n_kt * n_qc;
    const int w_last = Lq - (n_qc - 1) * AB_CW;
    const uint32_t idesc_sd_full = make_idesc_bf16(128, AB_CW, 0, 0), idesc_sd_last = make_idesc_bf16(128, w_last, 0, 0);
    uint32_t item_cnt = 0, tile_cnt0 = 0, step_cnt0 = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_cnt, tile_cnt0 += n_kt, step_cnt0 += n_steps) {
      mbar_wait(&bars[BB_QD_FULL], item_cnt & 1);
      mbar_wait(&bars[BB_DQ_EMPTY], (item_cnt & 1) ^ 1);  // previous item's dQ accumulators have been read out
      // S^T / dP^T of step (j, i)
      auto issue_sd = [&](int j, int i) {
        const uint32_t tile_cnt = tile_cnt0 + j;
        const int jb = tile_cnt & 1;
        if (i == 0) mbar_wait(&bars[BB_KV_FULL0 + jb], (tile_cnt >> 1) & 1);
        const uint32_t idesc_sd = i == n_qc - 1 ? idesc_sd_last : idesc_sd_full;
        const uint64_t ka = kd0 + jb * 1024, va = vd0 + jb * 1024;             // 16 KB tiles, descriptor units of 16 B
        const uint64_t qb = qd0 + i * (AB_CW * 8), dob = dod0 + i * (AB_CW * 8);  // AB_CW rows of 128 B
        tc_fence_after();
        if (at_elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_HD / 16; ++k) {
            umma_bf16(tm_s, ka + 2 * k, qb + 2 * k, idesc_sd, k > 0);
            umma_bf16(tm_dp, va + 2 * k, dob + 2 * k, idesc_sd, k > 0);
          }
          umma_commit(&bars[BB_SD_FULL]);
        }
        __syncwarp();
      };
      issue_sd(0, 0);
      int j = 0, i = 0;
      for (int s = 0; s < n_steps; ++s) {
        const uint32_t tile_cnt = tile_cnt0 + j, step_cnt = step_cnt0 + s;
        const int jb = tile_cnt & 1;
        if (s + 1 < n_steps) {
          mbar_wait(&bars[BB_SD_EMPTY], step_cnt & 1);  // S^T/dP^T of step s are in registers: the TMEM buffers are free
          if (i + 1 == n_qc) issue_sd(j + 1, 0);
          else issue_sd(j, i + 1);
        }
        const int nk = (i == n_qc - 1 ? w_last : AB_CW) >> 4;
        const uint64_t dob = domn0 + i * (AB_CW * 8), qb = qmn0 + i * (AB_CW * 8);
        const uint64_t ka = kmn0 + jb * 1024;
        if (i == 0) mbar_wait(&bars[BB_DKV_EMPTY], (tile_cnt & 1) ^ 1);  // previous tile's dV/dK have been read out
        mbar_wait(&bars[BB_PDS_FULL], step_cnt & 1);
        tc_fence_after();
        if (at_elect_one()) {
          for (int kk = 0; kk < nk; ++kk) {
            const uint32_t aoff = (kk >> 2) * 1024 + (kk & 3) * 2;
            umma_bf16(tm_dv, ptd0 + aoff, dob + kk * 128, idesc_acc, (i > 0 || kk > 0));
            umma_bf16(tm_dk, dsd0 + aoff, qb + kk * 128, idesc_acc, (i > 0 || kk > 0));
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // reduction over the 128 keys of tile j
            umma_bf16(tm_dq + i * 64, dsmn0 + kk * 128, ka + kk * 128, idesc_dq, (j > 0 || kk > 0));
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
        if (++i == n_qc) { i = 0; ++j; }
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
        lse_sm[i] = i < p.L ? -p.lse[bh * p.L + i] * AT_LOG2E : -INFINITY;  // negated: exponent = s c + (key bias + lse_sm)
        d_sm[i] = i < p.L ? -p.dsum[bh * p.L + i] * p.scale : 0.f;          // negated and scaled: dS = P (dP scale + d_sm)
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
            const float2 c2 = make_float2(c, c), sc2 = make_float2(p.scale, p.scale), kb2 = make_float2(kb, kb);
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
                // packed fp32x2 arithmetic: per pair one FADD2 (key bias + -lse), one FFMA2 (exponent), two ex2, one FFMA2 (dP scale - D scale),
                // one FMUL2, two packs
                const float2 ls[4] = {make_float2(l0.x, l0.y), make_float2(l0.z, l0.w), make_float2(l1.x, l1.y), make_float2(l1.z, l1.w)};
                const float2 dd[4] = {make_float2(d0.x, d0.y), make_float2(d0.z, d0.w), make_float2(d1.x, d1.y), make_float2(d1.z, d1.w)};
                uint32_t pw[4], dw[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const float2 z = __ffma2_rn(make_float2(__uint_as_float(sv[k][2 * x]), __uint_as_float(sv[k][2 * x + 1])), c2, __fadd2_rn(ls[x], kb2));
                  const float2 pe = make_float2(fast_exp2(z.x), fast_exp2(z.y));
                  const float2 u = __ffma2_rn(make_float2(__uint_as_float(dv[k][2 * x]), __uint_as_float(dv[k][2 * x + 1])), sc2, dd[x]);
                  const float2 de = __fmul2_rn(pe, u);
                  pw[x] = pack_bf16x2(pe.x, pe.y);
                  dw[x] = pack_bf16x2(de.x, de.y);
                }
                pk[k] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                dk[k] = make_uint4(dw[0], dw[1], dw[2], dw[3]);
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

// returns 1 if handled, 0 if the shape is outside this kernel (the caller takes the recompute kernels), < 0 on error; dsum = rowsum(dO * O)
// must already be in place (attn3_dsum_kernel)
int attention_bwd_merged(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* d_o, int64_t ldo,
                     const float* lse, const float* key_bias, void* dqkv, const float* dsum, int32_t B, int32_t H, int32_t L, int32_t head_dim,
                     float scale, cudaStream_t stream) {
  const int Lq = (L + 15) & ~15;
  if (head_dim != AT_HD || Lq > 3 * AB_CW || Lq < 16) return 0;
  const int64_t T = static_cast<int64_t>(B) * L;
  int rc;
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
