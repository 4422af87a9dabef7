// b200mm — tcgen05 / TMEM attention for head_dim 64 (the ViT-L/14, ViT-B/16 and BERT shapes).
//
// Same arithmetic contract as attention.cu (reference: nn.MultiheadAttention in clip/model.py:245-251 and
// BertSelfAttention.forward, clip/modeling_bert.py:134-172). The legacy mma.sync path tops out near 120 TFLOP/s on
// B200 (measured), so the contractions move to the 5th-gen tensor cores:
//
//   persistent CTA (one per SM) loops over (batch, head) items; per item K and V ([L,64] each) are TMA-loaded once into
//   128B-swizzled smem and the queries are processed in 128-row tiles:
//     S = Q_tile K^T     tcgen05.mma 128 x Lk x 64 (SS), fp32 S in TMEM (Lk <= 320 columns)
//     softmax            4 warps, thread == row == TMEM lane: two passes over S with tcgen05.ld (no cross-thread
//                        reductions), P written as bf16 into K-major swizzled smem tiles
//     O = P V            tcgen05.mma 128 x 64 x Lk, V read MN-major from the same [key][64] smem image
//     epilogue           O * 1/rowsum -> bf16 -> smem transpose -> coalesced stores; LSE per row
//   warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..7 = softmax/epilogue.
//   K/V/Q loads of the next tile/item run ahead through mbarrier rings; QK^T of tile t+1 overlaps the epilogue of t.
#include "common.cuh"

#include <stdlib.h>

#include <algorithm>

namespace b200mm {

constexpr int AT_HD = 64;
constexpr int AT_THREADS = 256;
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
};

enum { BAR_K_FULL = 0, BAR_K_EMPTY, BAR_V_FULL, BAR_V_EMPTY, BAR_Q_FULL0, BAR_Q_FULL1, BAR_Q_EMPTY0, BAR_Q_EMPTY1, BAR_S_FULL, BAR_P_FULL,
       BAR_O_FULL, BAR_O_EMPTY, BAR_COUNT };

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a [rows][128 B] tile with the 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[BAR_COUNT];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int Lk = p.Lk;
  const int kv_bytes = Lk * 128;                     // one [Lk][64] bf16 image
  const int kv_pad = (kv_bytes + 1023) & ~1023;
  const int n_ptiles = (Lk + 63) / 64;
  uint8_t* k_sm = smem;
  uint8_t* v_sm = k_sm + kv_pad;
  uint8_t* q_sm = v_sm + kv_pad;                     // 2 x 16 KB
  uint8_t* p_sm = q_sm + 2 * 16384;                  // n_ptiles x 16 KB
  float* bias_sm = reinterpret_cast<float*>(p_sm + n_ptiles * 16384);  // [Lk] key bias * log2e, -inf for key >= L

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < BAR_COUNT; ++i) mbar_init(&bars[i], (i == BAR_P_FULL || i == BAR_O_EMPTY) ? 128 : 1);
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
      for (int t = 0; t < n_qt; ++t, ++tile_cnt) {
        if (t == 0) {
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
        if (t == 0) {
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
      for (int t = 0; t < n_qt; ++t, ++tile_cnt) {
        const int qb = tile_cnt & 1;
        mbar_wait(&bars[BAR_Q_FULL0 + qb], (tile_cnt >> 1) & 1);
        if (t == 0) mbar_wait(&bars[BAR_K_FULL], item_cnt & 1);
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
          if (t == n_qt - 1) umma_commit(&bars[BAR_K_EMPTY]);
        }
        __syncwarp();
        mbar_wait(&bars[BAR_P_FULL], tile_cnt & 1);
        if (t == 0) mbar_wait(&bars[BAR_V_FULL], item_cnt & 1);
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
          if (t == n_qt - 1) umma_commit(&bars[BAR_V_EMPTY]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue (thread == query row == TMEM lane) =====================
    const int sw = warp - 4;
    const int r = sw * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(sw * 32) << 16;
    const float c = p.scale * AT_LOG2E;
    uint32_t tile_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / p.H, h = item - b * p.H;
      // key bias of this batch row (shared by its H heads, but items of one CTA are strided, so reload per item)
      asm volatile("bar.sync 1, 128;" ::: "memory");  // previous item's readers are done with bias_sm
      for (int i = threadIdx.x - 128; i < Lk; i += 128)
        bias_sm[i] = i < p.L ? (p.key_bias ? p.key_bias[static_cast<int64_t>(b) * p.L + i] * AT_LOG2E : 0.f) : -INFINITY;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int t = 0; t < n_qt; ++t, ++tile_cnt) {
        const int q_row = t * 128 + r;
        const bool warp_active = t * 128 + sw * 32 < p.L;  // warp-uniform: any valid row in this warp
        mbar_wait(&bars[BAR_S_FULL], tile_cnt & 1);
        tc_fence_after();
        float mx = -INFINITY, sum = 0.f;
        if (warp_active) {
          // pass 1: row maximum of s*c + bias
          for (int j0 = 0; j0 < Lk; j0 += 32) {
            if (j0 + 32 <= Lk) {
              uint32_t v[32];
              tmem_ld_32x32(tmem_s + lane_addr + j0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaf(__uint_as_float(v[j]), c, bias_sm[j0 + j]));
            } else {
              uint32_t v[16];
              tmem_ld_32x16(tmem_s + lane_addr + j0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fmaf(__uint_as_float(v[j]), c, bias_sm[j0 + j]));
            }
          }
          // pass 2: p = exp2(s*c + bias - mx) -> bf16 -> K-major swizzled P tiles
          for (int j0 = 0; j0 < Lk; j0 += 32) {
            uint8_t* ptile = p_sm + (j0 >> 6) * 16384;
            const int chunk0 = (j0 & 63) >> 3;
            if (j0 + 32 <= Lk) {
              uint32_t v[32];
              tmem_ld_32x32(tmem_s + lane_addr + j0, v);
              tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  e[j] = fast_exp2(fmaf(__uint_as_float(v[g * 8 + j]), c, bias_sm[j0 + g * 8 + j]) - mx);
                  sum += e[j];
                }
                uint4 o;
                o.x = pack_bf16x2(e[0], e[1]); o.y = pack_bf16x2(e[2], e[3]);
                o.z = pack_bf16x2(e[4], e[5]); o.w = pack_bf16x2(e[6], e[7]);
                *reinterpret_cast<uint4*>(ptile + sw128_off(r, chunk0 + g)) = o;
              }
            } else {
              uint32_t v[16];
              tmem_ld_32x16(tmem_s + lane_addr + j0, v);
              tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  e[j] = fast_exp2(fmaf(__uint_as_float(v[g * 8 + j]), c, bias_sm[j0 + g * 8 + j]) - mx);
                  sum += e[j];
                }
                uint4 o;
                o.x = pack_bf16x2(e[0], e[1]); o.y = pack_bf16x2(e[2], e[3]);
                o.z = pack_bf16x2(e[4], e[5]); o.w = pack_bf16x2(e[6], e[7]);
                *reinterpret_cast<uint4*>(ptile + sw128_off(r, chunk0 + g)) = o;
              }
            }
          }
        }
        // make the generic-proxy smem writes visible to the tensor core (async proxy), release S, publish P
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bars[BAR_P_FULL]);
        // ---- epilogue: O / sum -> bf16, transposed through this warp's 32 rows of P tile 0 (free once PV has completed)
        mbar_wait(&bars[BAR_O_FULL], tile_cnt & 1);
        tc_fence_after();
        if (warp_active) {
          const float inv = 1.f / sum;
          uint8_t* otile = p_sm;  // [128][128 B] swizzled, rows sw*32 .. +31 belong to this warp
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_o + lane_addr + half * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv);
              o.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv);
              o.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv);
              o.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(otile + sw128_off(r, half * 4 + g)) = o;
            }
          }
          if (q_row < p.L) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.L + q_row] = (mx + log2f(sum)) / AT_LOG2E;
        }
        tc_fence_before();
        mbar_arrive(&bars[BAR_O_EMPTY]);
        if (warp_active) {
          __syncwarp();
          // 8 lanes x 16 B = one 128-byte output row; 4 rows per warp instruction
          __nv_bfloat16* obase = p.o + (static_cast<int64_t>(b) * p.L + t * 128) * p.ldo + h * AT_HD;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = sw * 32 + it * 4 + (lane >> 3);
            const int ch = lane & 7;
            if (t * 128 + row < p.L)
              *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(row) * p.ldo + ch * 8) = *reinterpret_cast<const uint4*>(p_sm + sw128_off(row, ch));
          }
          __syncwarp();
        }
        // the next tile's P stores reuse p_sm: all 4 warps must have finished reading their staged O rows. P tile 0 rows are
        // warp-private in both uses (rows sw*32..+31), so a warp-level sync above is sufficient.
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

static size_t attn_tc_smem_bytes(int Lk) {
  const int kv_pad = (Lk * 128 + 1023) & ~1023;
  const int n_ptiles = (Lk + 63) / 64;
  return static_cast<size_t>(2) * kv_pad + 2 * 16384 + static_cast<size_t>(n_ptiles) * 16384 + Lk * 4 + 1024;
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
  const size_t smem = attn_tc_smem_bytes(Lk);
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) {
    set_last_error("attention_fwd_tc: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return B200MM_ERR_LAUNCH;
  }
  const int grid = std::min(B * H, sm_count());
  attn_fwd_tc_kernel<<<grid, AT_THREADS, smem, stream>>>(tmQ, tmKV, p);
  rc = check_launch("attn_fwd_tc_kernel");
  return rc ? rc : 1;
}

}  // namespace b200mm
