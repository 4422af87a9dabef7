// b200mm — multi-head self-attention on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), every shape of the path:
// head_dim any multiple of 16 up to 128 (ViT-B/L 64, ViT-H 80, M2-Encoder-10B 128), any sequence length.
//
// Reference arithmetic (unchanged contract of b200mm_attention_fwd / _bwd, include/b200mm.h):
//   ViT : nn.MultiheadAttention inside ResidualAttentionBlock.attention, antmmf/modules/vision/backbone/clip/model.py:245-251
//   BERT: BertSelfAttention.forward, antmmf/modules/vision/backbone/clip/modeling_bert.py:134-172 (additive key bias)
//   M2  : MultiheadAttention.forward, prj/M2_Encoder/vlmo/torchscale/component/multihead_attention.py:85-150
//
// Forward (attn3_fwd_kernel): persistent CTA per SM, 16 warps in four warpgroups with re-balanced register budgets (setmaxnreg):
//   WG0  warp 0 TMA producer, warp 1 tcgen05.mma issuer, warp 2 TMEM allocator
//   WG1/WG2  two softmax groups, ONE THREAD PER QUERY ROW (TMEM lane = row): no shuffles, no shared-memory exchange, no named
//            barriers anywhere in the softmax. The two groups work on two different 128-query tiles ("slots") in flight, so the exp
//            work of one tile hides the tensor-core round trip of the other.
//   WG3  epilogue: O / rowsum -> bf16 -> coalesced stores, log-sum-exp
//   per slot and key block (<= 128 keys):   S = Q K_j^T (SS-MMA, fp32 in TMEM)  ->  row max, P = exp2(S c - m) as bf16 written back
//   into the SAME TMEM columns  ->  O += P V_j (TS-MMA: A operand read from TMEM, V read MN-major from the [key][hd] smem image).
//   The running maximum is lazy: O is only rescaled (in TMEM, by the softmax thread itself) when a block raises a row maximum by more
//   than 2^8, so that in steady state nobody touches O between the first PV and the epilogue; exactness is unaffected (the final
//   O / l and LSE are those of the reference maximum).
//   K/V of a (batch, head) stay resident in shared memory across its query tiles when two items fit (ViT-L/14: 2 x 68 KB), otherwise
//   key blocks stream through a ring. head_dim > 64 is held as two 64-column (128-byte, 128B-swizzled) slabs per row.
//   A last query tile of <= 32 rows (the 257th token of ViT-L/14) is placed in a different lane quarter for successive items, which
//   balances its exp work over the four SM sub-partitions.
//
// Backward = two kernels of the same skeleton (recompute instead of atomics: bit-reproducible, no fp32 workspace):
//   attn3_bwd_dkv_kernel  unit = 128-key tile (TMEM lane = key), query blocks stream:  S^T = K Q_i^T, dP^T = V dO_i^T  ->
//                         P^T = exp2(S^T c + bias_k - lse_q), dS^T = P^T (dP^T - D_q) scale  (bf16, written back over S^T / dP^T)
//                         ->  dV += P^T dO_i, dK += dS^T Q_i (TS-MMA)
//   attn3_bwd_dq_kernel   unit = 128-query tile (TMEM lane = query, lse / D are per-thread scalars), key blocks stream:
//                         S = Q K_j^T, dP = dO V_j^T -> dS -> dQ += dS K_j
//   S/dP are double-buffered in TMEM, 12 elementwise warps split the COLUMNS of a block (the backward needs no row reduction), each
//   warp keeps the bf16 results inside its own column range so no warp ever overwrites scores another warp has not read yet.
#include "common.cuh"

#include <stdlib.h>

#include <algorithm>

namespace b200mm {

constexpr float A3_LOG2E = 1.4426950408889634f;
constexpr float A3_LN2 = 0.6931471805599453f;
constexpr int A3_MAX_STAGES = 8;
constexpr int A3_THREADS = 512;
constexpr int A3_FWD_NCH = 6;              // 16-column chunks of a forward key block held in registers per softmax thread (KB <= 96)
constexpr float A3_RESCALE_THRESHOLD = 8.f;  // log2 units: P stays <= 2^8, exact in bf16 range and fp32 sums

// ---------------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float a3_fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// Every mbarrier wait of these kernels goes through here. A wait that outlives any legitimate latency (seconds) traps, so that a protocol
// bug surfaces as a CUDA error instead of a hung GPU; build with -DA3_WATCHDOG=2 to also print which barrier / role / parity it was
// (the printf costs a stack frame and ~25 % of the kernel time, bring-up only), -DA3_WATCHDOG=0 removes the counter.
#ifndef A3_WATCHDOG
#define A3_WATCHDOG 1
#endif
__device__ __forceinline__ void a3_wait(uint64_t* bar, uint32_t parity, int tag) {
#if A3_WATCHDOG
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins == (1u << 24)) {
#if A3_WATCHDOG > 1
      printf("b200mm attention: mbarrier wait timed out (tag %d, parity %u, block %d, thread %d)\n", tag, parity, blockIdx.x, threadIdx.x);
#endif
      __trap();
    }
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
  (void)tag;
}
// ex2 whose position in the instruction stream is pinned (volatile): the softmax loops issue a whole 16-column chunk of MUFU ops back to
// back and only then consume the PREVIOUS chunk, so the ~40-cycle MUFU latency (measured, tools/ubench) is never exposed. Left to itself
// ptxas sinks each ex2 next to its consumer to save registers, and the single warp per scheduler then stalls on every pair.
__device__ __forceinline__ float a3_ex2v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void a3_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a [rows][128 B] tile with the 128B swizzle
__device__ __forceinline__ uint32_t a3_sw128(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (128 rows x 16 bf16 per instruction = 8 TMEM columns) is read from tensor memory
__device__ __forceinline__ void a3_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One elected lane of a converged warp (elect.sync). Code under `if (a3_elect_one())` that only touches warp-uniform values is compiled for
// the uniform datapath: tcgen05.mma / tcgen05.commit / TMA issue become single UTCHMMA / UTCBAR / UTMALDG instructions, instead of the
// vote + broadcast loop ptxas emits around them under an ordinary `lane == 0` predicate (15-25 instructions per MMA, measured: the
// issuing warp, not the tensor pipe, was the bottleneck of 48-cycle MMAs).
__device__ __forceinline__ bool a3_elect_one() {
  uint32_t q;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(q));
  return q != 0;
}
__device__ __forceinline__ void a3_tma_load(const CUtensorMap* m, uint64_t* bar, uint32_t smem_dst, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :
               : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void a3_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ float2 a3_u2f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// Where block j of unit g lives in the operand ring (K/V blocks in the forward and the dQ kernel, Q/dO blocks in the dK/dV kernel).
//   resident: the inner operand of an item is loaded once and shared by the item's units; two item slots of n_in stages each
//   streamed: every (unit, block) is a fill; with `pairs` the two units in flight interleave (g0,0) (g1,0) (g0,1) (g1,1) ...
struct A3Ring {
  int32_t n_in, n_outer, NS, resident, pairs;
};
struct A3Pos {
  int stage;
  uint32_t fill;
  bool first, last;
};
__device__ __forceinline__ A3Pos a3_ring_pos(const A3Ring& rg, int g, int j, int G) {
  A3Pos r;
  if (rg.resident) {
    const int it = g / rg.n_outer, tk = g - it * rg.n_outer;
    r.stage = (it & 1) * rg.n_in + j;
    r.fill = static_cast<uint32_t>(it >> 1);
    r.first = tk == 0;
    r.last = tk == rg.n_outer - 1;
  } else {
    int seq;
    if (rg.pairs) {
      const int base = (g >> 1) * 2 * rg.n_in;
      seq = ((g | 1) < G) ? base + 2 * j + (g & 1) : base + j;
    } else {
      seq = g * rg.n_in + j;
    }
    r.stage = seq % rg.NS;
    r.fill = static_cast<uint32_t>(seq / rg.NS);
    r.first = r.last = true;
  }
  return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
struct A3FwdParams {
  __nv_bfloat16* o;
  int64_t ldo;
  float* lse;
  const float* key_bias;
  int32_t B, H, L, Lk;       // Lk = L rounded up to 16
  int32_t hd, n_slabs;       // head dim, 64-column slabs per row
  int32_t n_qt, r_last;      // query tiles per item, rows of the last one
  int32_t KB, n_blk, kb_last;  // key block, blocks per item, keys (multiple of 16) of the last block
  int32_t NS, resident;
  int32_t q_off, k_off, v_off;
  float scale_log2;  // softmax scale * log2(e)
  uint32_t drop_thr;     // DROP kernels: probability (q, k) of item (b, h) is zeroed iff hash < drop_thr (common.cuh)
  float drop_inv_keep;   // 1 / (1 - p), applied to O in the epilogue
  uint64_t drop_seed;
};

template <bool HAS_BIAS, bool DROP>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn3_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQt, const __grid_constant__ CUtensorMap tmKV,
                 const A3FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t k_full[A3_MAX_STAGES], k_empty[A3_MAX_STAGES], v_full[A3_MAX_STAGES], v_empty[A3_MAX_STAGES];
  __shared__ __align__(8) uint64_t q_full[2], q_empty[2], s_full[2], p_full[2], pv_done[2], o_full[2], o_empty[2];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = p.n_slabs * p.KB * 128;
  const int slab_stride = p.KB * 128;  // between the 64-column slabs of a K/V stage
  const int qslot_bytes = p.n_slabs * 16384;
  uint8_t* k_sm = smem;
  uint8_t* v_sm = k_sm + p.NS * stage_bytes;
  uint8_t* q_sm = v_sm + p.NS * stage_bytes;
  uint8_t* e_sm = q_sm + 2 * qslot_bytes;                                     // epilogue staging: 4 warps x [n_slabs][32 rows][128 B]
  float2* stats = reinterpret_cast<float2*>(e_sm + p.n_slabs * 16384);        // [slot][parity][128] (row max (log2 domain), row sum)
  float* bias_sm = reinterpret_cast<float*>(stats + 512);                     // [slot][Lk] key bias * log2e, -inf for key >= L

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmQt);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < A3_MAX_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  // one arrival per softmax warp
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // TMEM columns: S/P of slot w at 128 w, O of slot w at 256 + 128 w
  const int n_items = p.B * p.H;
  const int my_items = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int G = my_items * p.n_qt;  // units of this CTA in order: g -> (item g / n_qt, query tile g % n_qt)
  const A3Ring ring{p.n_blk, p.n_qt, p.NS, p.resident, 1};
  const bool rot_tail = p.n_qt > 1 && p.r_last <= 32;  // short last tile: lane quarter rotates with the item

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int stage16 = stage_bytes >> 4, slab16 = slab_stride >> 4;
    if (warp == 0) {
      // ===================== TMA producer =====================
      // units are visited in pairs (the two slots in flight); ring cursors advance incrementally (no divisions in the loop)
      int it = 0, tk = 0;                          // (item, tile) of the pair's first unit
      int ks = 0, vs = 0;                          // streamed rings: next stage
      uint32_t kpar = 1, vpar = 1;                 // ... and the parity to wait for on its empty barrier
      const uint32_t q_s = smem_u32(q_sm), k_s = smem_u32(k_sm), v_s = smem_u32(v_sm);
      for (int g0 = 0; g0 < G; g0 += 2) {
        const int n_u = G - g0 < 2 ? G - g0 : 2;
        const int tk1 = tk + 1 == p.n_qt ? 0 : tk + 1, it1 = tk + 1 == p.n_qt ? it + 1 : it;
        for (int j = 0; j < p.n_blk; ++j) {
          for (int pass = 0; pass < 2; ++pass) {  // pass 0: (Q and) K of both units, pass 1: V of both units
            for (int x = 0; x < n_u; ++x) {
              const int itu = x ? it1 : it, tku = x ? tk1 : tk;
              const int item = blockIdx.x + itu * gridDim.x;
              const int b = item / p.H, h = item - b * p.H;
              const int32_t row0 = b * p.L;
              if (pass == 0 && j == 0) {
                a3_wait(&q_empty[x], ((g0 >> 1) & 1) ^ 1, 17 /*q_empty*/);
                const bool tail = tku == p.n_qt - 1;
                const int rows = tail ? p.r_last : 128;
                const int rot = (tail && rot_tail) ? (itu & 3) : 0;
                if (a3_elect_one()) {
                  mbar_expect_tx(&q_full[x], p.n_slabs * rows * 128);
                  for (int sl = 0; sl < p.n_slabs; ++sl)
                    a3_tma_load(tail ? &tmQt : &tmQ, &q_full[x], q_s + x * qslot_bytes + sl * 16384 + rot * 4096, p.q_off + h * p.hd + 64 * sl,
                                row0 + tku * 128);
                }
                __syncwarp();
              }
              int st;
              uint32_t par;
              bool load = true;
              if (p.resident) {
                st = (itu & 1) * p.n_blk + j;
                par = ((itu >> 1) & 1) ^ 1;
                load = tku == 0;
              } else if (pass == 0) {
                st = ks; par = kpar;
                if (++ks == p.NS) { ks = 0; kpar ^= 1; }
              } else {
                st = vs; par = vpar;
                if (++vs == p.NS) { vs = 0; vpar ^= 1; }
              }
              if (load) {
                uint64_t* full = pass == 0 ? &k_full[st] : &v_full[st];
                a3_wait(pass == 0 ? &k_empty[st] : &v_empty[st], par, 1 /*kv_empty*/);
                const uint32_t dst = (pass == 0 ? k_s : v_s) + st * stage_bytes;
                const int col = (pass == 0 ? p.k_off : p.v_off) + h * p.hd;
                if (a3_elect_one()) {
                  mbar_expect_tx(full, stage_bytes);
                  for (int sl = 0; sl < p.n_slabs; ++sl) a3_tma_load(&tmKV, full, dst + sl * slab_stride, col + 64 * sl, row0 + j * p.KB);
                }
                __syncwarp();
              }
            }
          }
        }
        // advance the pair cursor by two units
        for (int x = 0; x < 2; ++x)
          if (++tk == p.n_qt) { tk = 0; ++it; }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      const uint32_t idesc_pv = make_idesc_bf16(128, p.hd, 0, 1);  // A = P (TMEM, K-major), B = V MN-major
      const uint32_t idesc_s_full = make_idesc_bf16(128, p.KB, 0, 0), idesc_s_last = make_idesc_bf16(128, p.kb_last, 0, 0);
      const int ksteps = p.hd / 16;
      const uint64_t qd0 = make_smem_desc_sw128(smem_u32(q_sm), 16, 1024), kd0 = make_smem_desc_sw128(smem_u32(k_sm), 16, 1024);
      const uint64_t vd0 = make_smem_desc_sw128(smem_u32(v_sm), slab_stride, 1024);
      const int qslot16 = qslot_bytes >> 4;
      int it = 0, tk = 0;
      int ks = 0, vs = 0;
      uint32_t kpar = 0, vpar = 0;
      int it1 = 0, tk1 = 0;
      uint32_t upar = 0;   // (g >> 1) & 1 of the current pair
      uint32_t cpar0 = 0;  // parity of (u * n_blk) for the current pair
      // ring cursor of the next K / V block of slot x (resident: derived from the unit; streamed: the running ring position)
      auto k_pos = [&](int x, int j, int& st, uint32_t& par, bool& last) {
        if (p.resident) {
          st = ((x ? it1 : it) & 1) * p.n_blk + j;
          par = ((x ? it1 : it) >> 1) & 1;
          last = (x ? tk1 : tk) == p.n_qt - 1;
        } else {
          st = ks; par = kpar; last = true;
          if (++ks == p.NS) { ks = 0; kpar ^= 1; }
        }
      };
      auto v_pos = [&](int x, int j, int& st, uint32_t& par, bool& last) {
        if (p.resident) {
          st = ((x ? it1 : it) & 1) * p.n_blk + j;
          par = ((x ? it1 : it) >> 1) & 1;
          last = (x ? tk1 : tk) == p.n_qt - 1;
        } else {
          st = vs; par = vpar; last = true;
          if (++vs == p.NS) { vs = 0; vpar ^= 1; }
        }
      };
      // S(x, 0): first block of a unit
      auto issue_s0 = [&](int x) {
        int st;
        uint32_t par;
        bool last;
        k_pos(x, 0, st, par, last);
        const uint64_t qd = qd0 + x * qslot16, kd = kd0 + st * stage16;
        const uint32_t idesc = p.n_blk == 1 ? idesc_s_last : idesc_s_full;
        const uint32_t td = tmem_base + x * 128;
        a3_wait(&k_full[st], par, 3 /*k_full*/);
        a3_wait(&q_full[x], upar, 2 /*q_full*/);
        tc_fence_after();
        if (a3_elect_one()) {
          for (int k4 = 0; k4 < ksteps; ++k4)
            umma_bf16(td, qd + (k4 >> 2) * 1024 + (k4 & 3) * 2, kd + (k4 >> 2) * slab16 + (k4 & 3) * 2, idesc, k4 > 0);
          umma_commit(&s_full[x]);
          if (last) umma_commit(&k_empty[st]);
          if (p.n_blk == 1) umma_commit(&q_empty[x]);
        }
        __syncwarp();
      };
      // PV(x, j) followed by S(x, j+1). Everything that does not depend on the softmax (operand barriers, descriptors) is settled BEFORE
      // the wait on P: each mbarrier test costs ~90 cycles even when it succeeds, and the softmax group of this slot sits idle from its
      // arrival on p_full until S(j+1) lands.
      auto issue_pv_s = [&](int x, int j) {
        int vst, kst = 0;
        uint32_t vpr, kpr = 0;
        bool vlast, klast = false;
        v_pos(x, j, vst, vpr, vlast);
        const bool more = j + 1 < p.n_blk;
        if (more) k_pos(x, j + 1, kst, kpr, klast);
        const int nk = (j == p.n_blk - 1 ? p.kb_last : p.KB) >> 4;
        const uint64_t vd = vd0 + vst * stage16;
        const uint32_t to = tmem_base + 256 + x * 128, ts = tmem_base + x * 128;
        const uint64_t qd = qd0 + x * qslot16, kd = kd0 + kst * stage16;
        const uint32_t idesc = j + 1 == p.n_blk - 1 ? idesc_s_last : idesc_s_full;
        a3_wait(&v_full[vst], vpr, 5 /*v_full*/);
        if (j == 0) a3_wait(&o_empty[x], upar ^ 1, 6 /*o_empty*/);
        if (more) a3_wait(&k_full[kst], kpr, 3 /*k_full*/);
        a3_wait(&p_full[x], (cpar0 + j) & 1, 4 /*p_full*/);
        tc_fence_after();
        if (a3_elect_one()) {
          for (int kk = 0; kk < nk; ++kk) a3_umma_ts(to, ts + kk * 8, vd + kk * 128, idesc_pv, (j > 0 || kk > 0));
          umma_commit(&pv_done[x]);
          if (vlast) umma_commit(&v_empty[vst]);
          if (!more) umma_commit(&o_full[x]);
          if (more) {
            for (int k4 = 0; k4 < ksteps; ++k4)
              umma_bf16(ts, qd + (k4 >> 2) * 1024 + (k4 & 3) * 2, kd + (k4 >> 2) * slab16 + (k4 & 3) * 2, idesc, k4 > 0);
            umma_commit(&s_full[x]);
            if (klast) umma_commit(&k_empty[kst]);
            if (j + 1 == p.n_blk - 1) umma_commit(&q_empty[x]);
          }
        }
        __syncwarp();
      };
      for (int g0 = 0; g0 < G; g0 += 2) {
        const int n_u = G - g0 < 2 ? G - g0 : 2;
        tk1 = tk + 1 == p.n_qt ? 0 : tk + 1;
        it1 = tk + 1 == p.n_qt ? it + 1 : it;
        for (int x = 0; x < n_u; ++x) issue_s0(x);
        for (int j = 0; j < p.n_blk; ++j)
          for (int x = 0; x < n_u; ++x) issue_pv_s(x, j);
        for (int x = 0; x < 2; ++x)
          if (++tk == p.n_qt) { tk = 0; ++it; }
        upar ^= 1;
        cpar0 = (cpar0 + p.n_blk) & 1;
      }
    }
  } else if (warp < 12) {
    // ===================== softmax: one thread per query row, slot w =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
    const int w = (warp >> 2) - 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + w * 128 + lane_addr, t_o = tmem_base + 256 + w * 128 + lane_addr;
    const float cs_ = p.scale_log2;
    float* bias = bias_sm + w * p.Lk;
    int bias_item = -1;
    for (int g = w; g < G; g += 2) {
      const int it = g / p.n_qt, tk = g - it * p.n_qt;
      const int u = g >> 1;
      const bool tail = tk == p.n_qt - 1;
      const int rot = (tail && rot_tail) ? (it & 3) : 0;
      const int rows = tail ? p.r_last : 128;
      const bool active = rot_tail && tail ? quarter == rot : quarter * 32 < rows;  // warp-uniform: any valid row in this warp
      // dropout on the probabilities: stream = (batch, head) item, index = query * L + key. The row sum l stays that of the un-dropped
      // softmax (dropout follows the normalisation, modeling_bert.py:155-158); the 1 / (1 - p) factor is applied to O in the epilogue.
      uint32_t drop_key = 0u, drop_row = 0u;
      if (DROP) {
        drop_key = drop_stream_key(p.drop_seed, static_cast<uint64_t>(blockIdx.x + it * gridDim.x));
        drop_row = static_cast<uint32_t>(tk * 128 + r - 32 * rot) * static_cast<uint32_t>(p.L);
      }
      if (HAS_BIAS && bias_item != it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int b = item / p.H;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");  // the group is done with the previous item's bias
        for (int i = threadIdx.x & 127; i < p.Lk; i += 128) bias[i] = i < p.L ? p.key_bias[static_cast<int64_t>(b) * p.L + i] * A3_LOG2E : -INFINITY;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
        bias_item = it;
      }
      float m_ref = 0.f, l = 0.f;
      for (int j = 0; j < p.n_blk; ++j) {
        const uint32_t cs = static_cast<uint32_t>(u) * p.n_blk + j;
        const int kbj = j == p.n_blk - 1 ? p.kb_last : p.KB;
        a3_wait(&s_full[w], cs & 1, 7 /*s_full*/);
        tc_fence_after();
        if (active) {
          uint32_t s[A3_FWD_NCH][16];
#pragma unroll
          for (int c = 0; c < A3_FWD_NCH; ++c)
            if (c * 16 < kbj) tmem_ld_32x16(t_s + c * 16, s[c]);
#pragma unroll
          for (int c = 0; c < A3_FWD_NCH; ++c)
            if (c * 16 < kbj) tmem_ld_wait16(s[c]);
          // ---- block maximum (log2 domain)
          float mb;
          if (HAS_BIAS) {
            const float* bj = bias + j * p.KB;
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int c = 0; c < A3_FWD_NCH; ++c)
              if (c * 16 < kbj) {
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                  const float4 bv = *reinterpret_cast<const float4*>(bj + c * 16 + q);
                  const float z0 = fmaf(__uint_as_float(s[c][q]), cs_, bv.x), z1 = fmaf(__uint_as_float(s[c][q + 1]), cs_, bv.y);
                  const float z2 = fmaf(__uint_as_float(s[c][q + 2]), cs_, bv.z), z3 = fmaf(__uint_as_float(s[c][q + 3]), cs_, bv.w);
                  s[c][q] = __float_as_uint(z0);
                  s[c][q + 1] = __float_as_uint(z1);
                  s[c][q + 2] = __float_as_uint(z2);
                  s[c][q + 3] = __float_as_uint(z3);
                  m0 = a3_fmax3(m0, z0, z1);
                  m1 = a3_fmax3(m1, z2, z3);
                }
              }
            mb = fmaxf(m0, m1);
          } else {
            if (j == p.n_blk - 1 && p.Lk != p.L) {
              // keys >= L of the last 16-key chunk are padding: -inf before the maximum, exp2 gives exact zeros
              const int c_last = (kbj >> 4) - 1, k0 = j * p.KB + c_last * 16;
#pragma unroll
              for (int c = 0; c < A3_FWD_NCH; ++c)
                if (c == c_last) {
#pragma unroll
                  for (int q = 0; q < 16; ++q)
                    if (k0 + q >= p.L) s[c][q] = 0xff800000u;
                }
            }
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int c = 0; c < A3_FWD_NCH; ++c)
              if (c * 16 < kbj) {
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                  m0 = a3_fmax3(m0, __uint_as_float(s[c][q]), __uint_as_float(s[c][q + 1]));
                  m1 = a3_fmax3(m1, __uint_as_float(s[c][q + 2]), __uint_as_float(s[c][q + 3]));
                }
              }
            mb = fmaxf(m0, m1) * cs_;
          }
          if (j == 0) {
            m_ref = mb;
          } else {
            const bool need = mb > m_ref + A3_RESCALE_THRESHOLD;
            if (__any_sync(0xffffffffu, need)) {
              // rare: bring O (and the running sum) of this warp's rows to the new reference maximum; PV(j-1) must have completed
              a3_wait(&pv_done[w], (cs - 1) & 1, 8 /*pv_done*/);
              tc_fence_after();
              const float m_new = need ? mb : m_ref;
              const float alpha = fast_ex2(m_ref - m_new);
              l *= alpha;
              m_ref = m_new;
              for (int c = 0; c < p.hd / 16; ++c) {
                uint32_t ov[16];
                tmem_ld_32x16(t_o + c * 16, ov);
                tmem_ld_wait16(ov);
#pragma unroll
                for (int q = 0; q < 16; ++q) ov[q] = __float_as_uint(__uint_as_float(ov[q]) * alpha);
                tmem_st_32x16(t_o + c * 16, ov);
              }
              tmem_st_wait();
            }
          }
          // ---- P = exp2(z - m_ref) as bf16, written back over the first kbj/2 columns of S. Software-pipelined by one chunk: the 16 ex2
          //      of chunk c are issued before chunk c-1 is summed, packed and stored.
          float2 acc = make_float2(0.f, 0.f);
          const float2 c2 = make_float2(HAS_BIAS ? 1.f : cs_, HAS_BIAS ? 1.f : cs_), nm2 = make_float2(-m_ref, -m_ref);
          float e[2][16];
#pragma unroll
          for (int c = 0; c <= A3_FWD_NCH; ++c) {
            if (c < A3_FWD_NCH && c * 16 < kbj) {
#pragma unroll
              for (int q = 0; q < 16; q += 2) {
                const float2 z = __ffma2_rn(a3_u2f2(s[c < A3_FWD_NCH ? c : 0][q], s[c < A3_FWD_NCH ? c : 0][q + 1]), c2, nm2);
                e[c & 1][q] = a3_ex2v(z.x);
                e[c & 1][q + 1] = a3_ex2v(z.y);
              }
            }
            if (c > 0 && (c - 1) * 16 < kbj) {
              uint32_t pk[8];
#pragma unroll
              for (int q = 0; q < 16; q += 2) {
                acc = __fadd2_rn(acc, make_float2(e[(c - 1) & 1][q], e[(c - 1) & 1][q + 1]));
                if (DROP) {
                  const uint32_t idx = drop_row + static_cast<uint32_t>(j * p.KB + (c - 1) * 16 + q);
                  pk[q >> 1] = pack_bf16x2(drop_keep(drop_key, idx, p.drop_thr) ? e[(c - 1) & 1][q] : 0.f,
                                           drop_keep(drop_key, idx + 1, p.drop_thr) ? e[(c - 1) & 1][q + 1] : 0.f);
                } else {
                  pk[q >> 1] = pack_bf16x2(e[(c - 1) & 1][q], e[(c - 1) & 1][q + 1]);
                }
              }
              a3_tmem_st8(t_s + (c - 1) * 8, pk);
            }
          }
          l += acc.x + acc.y;
          if (j == p.n_blk - 1) stats[(w * 2 + (u & 1)) * 128 + r] = make_float2(m_ref, l);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[w]);
      }
    }
  } else {
    // ===================== epilogue: O / rowsum -> bf16 -> global, LSE =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint8_t* stg = e_sm + quarter * (p.n_slabs * 4096);  // this warp's [n_slabs][32][128 B]
    const int cpr = p.hd / 8;                            // 16-byte chunks per output row
    for (int g = 0; g < G; ++g) {
      const int w = g & 1, u = g >> 1;
      const int it = g / p.n_qt, tk = g - it * p.n_qt;
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / p.H, h = item - b * p.H;
      const bool tail = tk == p.n_qt - 1;
      const int rot = (tail && rot_tail) ? (it & 3) : 0;
      const int rows = tail ? p.r_last : 128;
      const bool active = rot_tail && tail ? quarter == rot : quarter * 32 < rows;
      a3_wait(&o_full[w], u & 1, 9 /*o_full*/);
      tc_fence_after();
      if (active) {
        const float2 st = stats[(w * 2 + (u & 1)) * 128 + r];
        const float inv = DROP ? p.drop_inv_keep / st.y : 1.f / st.y;
        const uint32_t t_o = tmem_base + 256 + w * 128 + lane_addr;
        for (int c = 0; c < p.hd / 16; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(t_o + c * 16, v);
          tmem_ld_wait16(v);
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
          uint8_t* slab = stg + (c >> 2) * 4096;
          *reinterpret_cast<uint4*>(slab + a3_sw128(lane, (c & 3) * 2)) = o;
          o.x = pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
          o.y = pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
          o.z = pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
          o.w = pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
          *reinterpret_cast<uint4*>(slab + a3_sw128(lane, (c & 3) * 2 + 1)) = o;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[w]);
        // rows of this warp: tile row = r - 32 rot
        const int tr0 = (quarter - rot) * 32;
        const int q_row = tk * 128 + tr0 + lane;
        if (tr0 + lane < rows) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.L + q_row] = (st.x + log2f(st.y)) * A3_LN2;
        __nv_bfloat16* obase = p.o + (static_cast<int64_t>(b) * p.L + tk * 128 + tr0) * p.ldo + h * p.hd;
        const int n_rows = rows - tr0 < 32 ? rows - tr0 : 32;
        for (int idx = lane; idx < n_rows * cpr; idx += 32) {
          const int row = idx / cpr, ch = idx - row * cpr;
          *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(row) * p.ldo + ch * 8) =
              *reinterpret_cast<const uint4*>(stg + (ch >> 3) * 4096 + a3_sw128(row, ch & 7));
        }
        __syncwarp();
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[w]);
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


// ---------------------------------------------------------------------------------------------------------------------
// backward: one skeleton, two modes
//   outer tile (TMEM lanes) = two A operands X0, X1 [128 rows][hd]; inner blocks (TMEM columns) = two B operands Y0, Y1 [NB rows][hd]
//     T0 = X0 Y0^T, T1 = X1 Y1^T      (SS-MMA, double-buffered stages in TMEM)
//     elementwise: F0 (bf16, over T0), F1 (bf16, over T1)
//     acc0 += F0 Y1 (dK/dV mode only), acc1 += F1 Y0      (TS-MMA, Y read MN-major)
//   MODE_DKV: X = (K_t, V_t), Y = (Q_i, dO_i): T0 = S^T, T1 = dP^T, F0 = P^T, F1 = dS^T, acc0 = dV, acc1 = dK
//   MODE_DQ : X = (Q_t, dO_t), Y = (K_j, V_j): T0 = S,   T1 = dP,   F1 = dS,             acc1 = dQ
// ---------------------------------------------------------------------------------------------------------------------
enum { A3_MODE_DKV = 0, A3_MODE_DQ = 1 };
constexpr int A3_EW_WARPS = 12;

struct A3BwdParams {
  const float* lse;
  const float* dsum;
  const float* key_bias;
  __nv_bfloat16* dqkv;
  int64_t ld;
  int32_t B, H, L, Lk;  // Lk = L rounded up to 16 (inner extent)
  int32_t hd, n_slabs;
  int32_t n_outer, r_last;     // outer tiles per item, rows of the last one
  int32_t NB, n_in, nb_last;   // inner block, blocks per item, columns (multiple of 16) of the last block
  int32_t NS, resident;
  int32_t x0_off, x1_off, y0_off, y1_off;  // column offsets of the four operands inside their matrices
  int32_t q_off, k_off, v_off;
  float scale, scale_log2;
  uint32_t drop_thr;     // DROP kernels: see A3FwdParams
  float drop_inv_keep;
  uint64_t drop_seed;
};

template <int MODE, bool HAS_BIAS, bool DROP>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn3_bwd_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX0t, const __grid_constant__ CUtensorMap tmX1,
                 const __grid_constant__ CUtensorMap tmX1t, const __grid_constant__ CUtensorMap tmY0, const __grid_constant__ CUtensorMap tmY1,
                 const A3BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t y_full[A3_MAX_STAGES], y_empty[A3_MAX_STAGES];
  __shared__ __align__(8) uint64_t x_full[2], x_empty[2], sd_full[2], pds_full[2], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int ytile_bytes = p.n_slabs * p.NB * 128;  // one of Y0 / Y1
  const int stage_bytes = 2 * ytile_bytes;
  const int yslab_stride = p.NB * 128;
  const int xtile_bytes = p.n_slabs * 16384;
  uint8_t* y_sm = smem;
  uint8_t* x_sm = y_sm + p.NS * stage_bytes;                                // [2 slots][X0, X1]
  float* colc0 = reinterpret_cast<float*>(x_sm + 4 * xtile_bytes);           // DKV: -lse*log2e per query (-inf padding); DQ: key bias*log2e
  float* colc1 = colc0 + p.Lk;                                               // DKV: -D*scale per query

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmY0);
    tma_prefetch_desc(&tmY1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < A3_MAX_STAGES; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
      mbar_init(&sd_full[i], 1);
      mbar_init(&pds_full[i], A3_EW_WARPS);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, A3_EW_WARPS);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // TMEM columns: stage t: T0 at 2 t NB, T1 at (2 t + 1) NB; accumulators from 4 NB (acc0 then acc1 in dK/dV mode)
  const uint32_t t_acc = tmem_base + 4 * p.NB;
  const uint32_t t_acc1 = MODE == A3_MODE_DKV ? t_acc + p.hd : t_acc;

  const int n_items = p.B * p.H;
  const int my_items = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int G = my_items * p.n_outer;
  const int F = G * p.n_in;  // flat (unit, block) steps of this CTA
  const A3Ring ring{p.n_in, p.n_outer, p.NS, p.resident, 0};
  const bool rot_tail = p.n_outer > 1 && p.r_last <= 32;

  const int stage16 = stage_bytes >> 4, yslab16 = yslab_stride >> 4, ytile16 = ytile_bytes >> 4, xtile16 = xtile_bytes >> 4;
  if (warp == 0) {
    // ===================== TMA producer =====================
    int it = 0, tk = 0;
    int ys = 0;
    uint32_t ypar = 1;
    const uint32_t x_s = smem_u32(x_sm), y_s = smem_u32(y_sm);
    for (int g = 0; g < G; ++g) {
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / p.H, h = item - b * p.H;
      const int32_t row0 = b * p.L;
      const int xs = g & 1;
      a3_wait(&x_empty[xs], ((g >> 1) & 1) ^ 1, 18 /*x_empty*/);
      {
        const bool tail = tk == p.n_outer - 1;
        const int rows = tail ? p.r_last : 128;
        const int rot = (tail && rot_tail) ? (it & 3) : 0;
        const uint32_t dst = x_s + xs * 2 * xtile_bytes + rot * 4096;
        if (a3_elect_one()) {
          mbar_expect_tx(&x_full[xs], 2 * p.n_slabs * rows * 128);
          for (int sl = 0; sl < p.n_slabs; ++sl) {
            a3_tma_load(tail ? &tmX0t : &tmX0, &x_full[xs], dst + sl * 16384, p.x0_off + h * p.hd + 64 * sl, row0 + tk * 128);
            a3_tma_load(tail ? &tmX1t : &tmX1, &x_full[xs], dst + xtile_bytes + sl * 16384, p.x1_off + h * p.hd + 64 * sl, row0 + tk * 128);
          }
        }
        __syncwarp();
      }
      if (!p.resident || tk == 0) {
        for (int i = 0; i < p.n_in; ++i) {
          int st;
          uint32_t par;
          if (p.resident) {
            st = (it & 1) * p.n_in + i;
            par = ((it >> 1) & 1) ^ 1;
          } else {
            st = ys; par = ypar;
            if (++ys == p.NS) { ys = 0; ypar ^= 1; }
          }
          a3_wait(&y_empty[st], par, 10 /*y_empty*/);
          const uint32_t dst = y_s + st * stage_bytes;
          if (a3_elect_one()) {
            mbar_expect_tx(&y_full[st], stage_bytes);
            for (int sl = 0; sl < p.n_slabs; ++sl) {
              a3_tma_load(&tmY0, &y_full[st], dst + sl * yslab_stride, p.y0_off + h * p.hd + 64 * sl, row0 + i * p.NB);
              a3_tma_load(&tmY1, &y_full[st], dst + ytile_bytes + sl * yslab_stride, p.y1_off + h * p.hd + 64 * sl, row0 + i * p.NB);
            }
          }
          __syncwarp();
        }
      }
      if (++tk == p.n_outer) { tk = 0; ++it; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_acc = make_idesc_bf16(128, p.hd, 0, 1);  // A = F (TMEM, K-major), B = Y MN-major
    const uint32_t idesc_t_full = make_idesc_bf16(128, p.NB, 0, 0), idesc_t_last = make_idesc_bf16(128, p.nb_last, 0, 0);
    const int ksteps = p.hd / 16;
    const uint64_t xd0 = make_smem_desc_sw128(smem_u32(x_sm), 16, 1024), yd0 = make_smem_desc_sw128(smem_u32(y_sm), 16, 1024);
    const uint64_t ymn0 = make_smem_desc_sw128(smem_u32(y_sm), yslab_stride, 1024);
    // two cursors walk the flat (unit, block) sequence: T (two steps ahead) and acc
    struct Cur {
      int g, it, tk, i, ys;
      uint32_t ypar;
    };
    Cur ct{0, 0, 0, 0, 0, 0}, ca{0, 0, 0, 0, 0, 0};
    auto advance = [&](Cur& c) {
      if (!p.resident && ++c.ys == p.NS) { c.ys = 0; c.ypar ^= 1; }
      if (++c.i == p.n_in) {
        c.i = 0;
        ++c.g;
        if (++c.tk == p.n_outer) { c.tk = 0; ++c.it; }
      }
    };
    auto stage_of = [&](const Cur& c, int& st, uint32_t& par) {
      if (p.resident) { st = (c.it & 1) * p.n_in + c.i; par = (c.it >> 1) & 1; }
      else { st = c.ys; par = c.ypar; }
    };
    // T(f): S / dP of flat step f (cursor ct)
    auto t_waits = [&](int& st) {
      uint32_t par;
      stage_of(ct, st, par);
      if (ct.i == 0) a3_wait(&x_full[ct.g & 1], (ct.g >> 1) & 1, 11 /*x_full*/);
      a3_wait(&y_full[st], par, 12 /*y_full*/);
    };
    auto t_issue = [&](int f, int st) {  // inside an elected region
      const int t = f & 1, xs = ct.g & 1;
      const uint32_t idesc = ct.i == p.n_in - 1 ? idesc_t_last : idesc_t_full;
      const uint64_t xd = xd0 + xs * 2 * xtile16, yd = yd0 + st * stage16;
      const uint32_t t0 = tmem_base + 2 * t * p.NB, t1 = t0 + p.NB;
      for (int k4 = 0; k4 < ksteps; ++k4) {
        const int ao = (k4 >> 2) * 1024 + (k4 & 3) * 2, bo = (k4 >> 2) * yslab16 + (k4 & 3) * 2;
        umma_bf16(t0, xd + ao, yd + bo, idesc, k4 > 0);
        umma_bf16(t1, xd + xtile16 + ao, yd + ytile16 + bo, idesc, k4 > 0);
      }
      umma_commit(&sd_full[t]);
    };
    auto issue_t = [&](int f) {
      int st;
      t_waits(st);
      tc_fence_after();
      if (a3_elect_one()) t_issue(f, st);
      __syncwarp();
      advance(ct);
    };
    // acc(f) followed by T(f+2). The operand barriers of T(f+2) are settled before the wait on the elementwise warps when that cannot
    // deadlock (its ring stage is not the one acc(f) is about to release): every mbarrier test costs ~90 cycles, and the elementwise warps
    // of this stage idle until T(f+2) lands.
    // (with one block per unit T(f+2) belongs to the unit after next, whose operand slot is only released by acc(f) itself)
    const bool prewait = p.n_in >= 2 && (p.resident || p.NS >= 3);
    auto issue_acc_t = [&](int f) {
      const int t = f & 1, xs = ca.g & 1;
      const bool has_t = f + 2 < F;
      int st, tst = 0;
      uint32_t par;
      stage_of(ca, st, par);
      const int n_ch = (ca.i == p.n_in - 1 ? p.nb_last : p.NB) >> 4;
      const uint64_t yd = ymn0 + st * stage16;
      const uint32_t t0 = tmem_base + 2 * t * p.NB, t1 = t0 + p.NB;
      if (has_t && prewait) t_waits(tst);
      if (ca.i == 0) a3_wait(&acc_empty, (ca.g & 1) ^ 1, 14 /*acc_empty*/);  // the previous unit's accumulators have been read out
      a3_wait(&pds_full[t], (f >> 1) & 1, 13 /*pds_full*/);
      tc_fence_after();
      if (a3_elect_one()) {
        for (int wi = 0; wi < 3; ++wi) {
          const int c0 = (wi * n_ch) / 3, c1 = ((wi + 1) * n_ch) / 3;
          for (int c = c0; c < c1; ++c) {
            const uint32_t acol = 16 * c0 + 8 * (c - c0);  // where the owner warp stored bf16 chunk c
            const uint32_t accum = ca.i > 0 || c > 0;
            if (MODE == A3_MODE_DKV) a3_umma_ts(t_acc, t0 + acol, yd + ytile16 + c * 128, idesc_acc, accum);
            a3_umma_ts(t_acc1, t1 + acol, yd + c * 128, idesc_acc, accum);
          }
        }
        if (!p.resident || ca.tk == p.n_outer - 1) umma_commit(&y_empty[st]);
        if (ca.i == p.n_in - 1) {
          umma_commit(&acc_full);
          umma_commit(&x_empty[xs]);
        }
        if (has_t && prewait) t_issue(f + 2, tst);
      }
      __syncwarp();
      advance(ca);
      if (has_t && prewait) advance(ct);
      if (has_t && !prewait) issue_t(f + 2);
    };
    if (F > 0) issue_t(0);
    if (F > 1) issue_t(1);
    for (int f = 0; f < F; ++f) issue_acc_t(f);
  } else if (warp >= 4) {
    // ===================== elementwise + epilogue: 3 warps per lane quarter, each owns a column range of the block =====================
    const int quarter = warp & 3, wi = (warp >> 2) - 1;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2), sc2 = make_float2(p.scale, p.scale);
    auto ew_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(A3_EW_WARPS * 32) : "memory"); };
    int const_item = -1;
    int f = 0;
    for (int g = 0; g < G; ++g) {
      const int it = g / p.n_outer, tk = g - it * p.n_outer;
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / p.H, h = item - b * p.H;
      const int64_t bh = static_cast<int64_t>(b) * p.H + h;
      const bool tail = tk == p.n_outer - 1;
      const int rot = (tail && rot_tail) ? (it & 3) : 0;
      const int rows = tail ? p.r_last : 128;
      const bool active = rot_tail && tail ? quarter == rot : quarter * 32 < rows;
      const int trow = r - 32 * rot;          // row inside the tile
      const int orow = tk * 128 + trow;       // row inside the item (a key in dK/dV mode, a query in dQ mode)
      const bool row_ok = trow >= 0 && trow < rows;
      // ---- per-item column constants
      if ((MODE == A3_MODE_DKV || HAS_BIAS) && const_item != it) {
        ew_sync();
        for (int i = threadIdx.x - 128; i < p.Lk; i += A3_EW_WARPS * 32) {
          if (MODE == A3_MODE_DKV) {
            colc0[i] = i < p.L ? -p.lse[bh * p.L + i] * A3_LOG2E : -INFINITY;
            colc1[i] = i < p.L ? -p.dsum[bh * p.L + i] * p.scale : 0.f;
          } else {
            colc0[i] = i < p.L ? p.key_bias[static_cast<int64_t>(b) * p.L + i] * A3_LOG2E : -INFINITY;
          }
        }
        ew_sync();
        const_item = it;
      }
      // ---- per-row constants
      float rc0 = 0.f, rc1 = 0.f;
      if (MODE == A3_MODE_DKV) {
        if (HAS_BIAS) rc0 = row_ok ? p.key_bias[static_cast<int64_t>(b) * p.L + orow] * A3_LOG2E : 0.f;
      } else {
        rc0 = row_ok ? -p.lse[bh * p.L + orow] * A3_LOG2E : 0.f;
        rc1 = row_ok ? -p.dsum[bh * p.L + orow] * p.scale : 0.f;
      }
      const float2 r0 = make_float2(rc0, rc0), r1 = make_float2(rc1, rc1);
      // dropout on the probabilities, regenerated from the seed: index = query * L + key (row = key in dK/dV mode, query in dQ mode)
      uint32_t drop_key = 0u;
      if (DROP) drop_key = drop_stream_key(p.drop_seed, static_cast<uint64_t>(item));
      for (int i = 0; i < p.n_in; ++i, ++f) {
        const int t = f & 1;
        const int nbj = i == p.n_in - 1 ? p.nb_last : p.NB;
        const int n_ch = nbj >> 4;
        const int c0 = (wi * n_ch) / 3, c1 = ((wi + 1) * n_ch) / 3;
        const uint32_t t0 = tmem_base + 2 * t * p.NB + lane_addr, t1 = t0 + p.NB;
        a3_wait(&sd_full[t], (f >> 1) & 1, 15 /*sd_full*/);
        tc_fence_after();
        if (active) {
          uint32_t a[2][16], d[2][16];
#pragma unroll
          for (int k = 0; k < 2; ++k)
            if (c0 + k < c1) {
              tmem_ld_32x16(t0 + (c0 + k) * 16, a[k]);
              tmem_ld_32x16(t1 + (c0 + k) * 16, d[k]);
            }
#pragma unroll
          for (int k = 0; k < 2; ++k)
            if (c0 + k < c1) {
              tmem_ld_wait16(a[k]);
              tmem_ld_wait16(d[k]);
            }
          // phase 1: exponents of both chunks, then all ex2 back to back (the MUFU latency hides behind its own throughput)
          float pe[2][16];
#pragma unroll
          for (int k = 0; k < 2; ++k)
            if (c0 + k < c1) {
              const int col0 = i * p.NB + (c0 + k) * 16;  // inner index of the chunk's first column
              if (MODE == A3_MODE_DQ && !HAS_BIAS && i == p.n_in - 1 && col0 + 16 > p.L) {
#pragma unroll
                for (int x = 0; x < 16; ++x)
                  if (col0 + x >= p.L) a[k][x] = 0xff800000u;  // padded keys: exp2(-inf) = 0
              }
#pragma unroll
              for (int x = 0; x < 16; x += 4) {
                float2 za, zb;
                if (MODE == A3_MODE_DKV) {
                  const float4 l4 = *reinterpret_cast<const float4*>(colc0 + col0 + x);
                  float2 la = make_float2(l4.x, l4.y), lb = make_float2(l4.z, l4.w);
                  if (HAS_BIAS) {
                    la = __fadd2_rn(la, r0);
                    lb = __fadd2_rn(lb, r0);
                  }
                  za = __ffma2_rn(a3_u2f2(a[k][x], a[k][x + 1]), c2, la);
                  zb = __ffma2_rn(a3_u2f2(a[k][x + 2], a[k][x + 3]), c2, lb);
                } else {
                  za = __ffma2_rn(a3_u2f2(a[k][x], a[k][x + 1]), c2, r0);
                  zb = __ffma2_rn(a3_u2f2(a[k][x + 2], a[k][x + 3]), c2, r0);
                  if (HAS_BIAS) {
                    const float4 b4 = *reinterpret_cast<const float4*>(colc0 + col0 + x);
                    za = __fadd2_rn(za, make_float2(b4.x, b4.y));
                    zb = __fadd2_rn(zb, make_float2(b4.z, b4.w));
                  }
                }
                pe[k][x] = za.x;
                pe[k][x + 1] = za.y;
                pe[k][x + 2] = zb.x;
                pe[k][x + 3] = zb.y;
              }
            }
#pragma unroll
          for (int k = 0; k < 2; ++k)
            if (c0 + k < c1) {
#pragma unroll
              for (int x = 0; x < 16; ++x) pe[k][x] = a3_ex2v(pe[k][x]);
            }
          // phase 2: dS = P (dP scale - D scale), pack, write back into this warp's own columns
#pragma unroll
          for (int k = 0; k < 2; ++k)
            if (c0 + k < c1) {
              const int col0 = i * p.NB + (c0 + k) * 16;
              uint32_t pk[8], dk[8];
#pragma unroll
              for (int x = 0; x < 16; x += 4) {
                float2 ua, ub;
                if (MODE == A3_MODE_DKV) {
                  const float4 d4 = *reinterpret_cast<const float4*>(colc1 + col0 + x);
                  ua = __ffma2_rn(a3_u2f2(d[k][x], d[k][x + 1]), sc2, make_float2(d4.x, d4.y));
                  ub = __ffma2_rn(a3_u2f2(d[k][x + 2], d[k][x + 3]), sc2, make_float2(d4.z, d4.w));
                } else {
                  ua = __ffma2_rn(a3_u2f2(d[k][x], d[k][x + 1]), sc2, r1);
                  ub = __ffma2_rn(a3_u2f2(d[k][x + 2], d[k][x + 3]), sc2, r1);
                }
                float2 pa = make_float2(pe[k][x], pe[k][x + 1]), pb = make_float2(pe[k][x + 2], pe[k][x + 3]);
                if (DROP) {
                  // dS = P o (m dP / (1-p) - D) scale,  dV += (P o m / (1-p))^T dO  with m the keep mask: recompute ua / ub with the masked dP
                  const uint32_t L_ = static_cast<uint32_t>(p.L), o_ = static_cast<uint32_t>(orow), cc = static_cast<uint32_t>(col0 + x);
                  bool kp[4];
#pragma unroll
                  for (int y = 0; y < 4; ++y)
                    kp[y] = drop_keep(drop_key, MODE == A3_MODE_DKV ? (cc + y) * L_ + o_ : o_ * L_ + cc + y, p.drop_thr);
                  const float ik = p.drop_inv_keep;
                  float2 dterm_a, dterm_b;
                  if (MODE == A3_MODE_DKV) {
                    const float4 d4 = *reinterpret_cast<const float4*>(colc1 + col0 + x);
                    dterm_a = make_float2(d4.x, d4.y);
                    dterm_b = make_float2(d4.z, d4.w);
                  } else {
                    dterm_a = r1;
                    dterm_b = r1;
                  }
                  ua.x = fmaf(kp[0] ? __uint_as_float(d[k][x]) * ik : 0.f, p.scale, dterm_a.x);
                  ua.y = fmaf(kp[1] ? __uint_as_float(d[k][x + 1]) * ik : 0.f, p.scale, dterm_a.y);
                  ub.x = fmaf(kp[2] ? __uint_as_float(d[k][x + 2]) * ik : 0.f, p.scale, dterm_b.x);
                  ub.y = fmaf(kp[3] ? __uint_as_float(d[k][x + 3]) * ik : 0.f, p.scale, dterm_b.y);
                  const float2 da_ = __fmul2_rn(pa, ua), db_ = __fmul2_rn(pb, ub);
                  if (MODE == A3_MODE_DKV) {
                    pk[x >> 1] = pack_bf16x2(kp[0] ? pa.x * ik : 0.f, kp[1] ? pa.y * ik : 0.f);
                    pk[(x >> 1) + 1] = pack_bf16x2(kp[2] ? pb.x * ik : 0.f, kp[3] ? pb.y * ik : 0.f);
                  }
                  dk[x >> 1] = pack_bf16x2(da_.x, da_.y);
                  dk[(x >> 1) + 1] = pack_bf16x2(db_.x, db_.y);
                  continue;
                }
                const float2 da = __fmul2_rn(pa, ua), db = __fmul2_rn(pb, ub);
                if (MODE == A3_MODE_DKV) {
                  pk[x >> 1] = pack_bf16x2(pa.x, pa.y);
                  pk[(x >> 1) + 1] = pack_bf16x2(pb.x, pb.y);
                }
                dk[x >> 1] = pack_bf16x2(da.x, da.y);
                dk[(x >> 1) + 1] = pack_bf16x2(db.x, db.y);
              }
              const uint32_t scol = 16 * c0 + 8 * k;  // bf16 results stay inside this warp's own column range
              if (MODE == A3_MODE_DKV) a3_tmem_st8(t0 + scol, pk);
              a3_tmem_st8(t1 + scol, dk);
            }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_full[t]);
      }
      // ---- unit epilogue: accumulators -> bf16 rows of dqkv; warp wi takes the 16-column chunks e = wi, wi + 3, ...
      a3_wait(&acc_full, g & 1, 16 /*acc_full*/);
      tc_fence_after();
      if (active) {
        const int n_e = (MODE == A3_MODE_DKV ? 2 : 1) * (p.hd >> 4);
        __nv_bfloat16* grow = p.dqkv + (static_cast<int64_t>(b) * p.L + orow) * p.ld + h * p.hd;
        for (int e = wi; e < n_e; e += 3) {
          uint32_t v[16];
          tmem_ld_32x16(t_acc + lane_addr + e * 16, v);
          tmem_ld_wait16(v);
          int col;
          if (MODE == A3_MODE_DKV) col = e < (p.hd >> 4) ? p.v_off + e * 16 : p.k_off + e * 16 - p.hd;
          else col = p.q_off + e * 16;
          if (row_ok) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]));
            o.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
            o.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]));
            o.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
            *reinterpret_cast<uint4*>(grow + col) = o;
            o.x = pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9]));
            o.y = pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11]));
            o.z = pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13]));
            o.w = pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15]));
            *reinterpret_cast<uint4*>(grow + col + 8) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// D[b,h,l] = sum_c dO[b,l,h*hd+c] * O[b,l,h*hd+c] (HBM-bound: 2 T W bf16 read once). One warp per token row; P = hd/8 lanes of 16 bytes
// cover a head, 32/P whole heads per warp step (coalesced), segmented shuffle reduction inside each head.
__global__ void __launch_bounds__(256) attn3_dsum_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int64_t ldo,
                                                         float* __restrict__ dsum, int64_t T, int32_t H, int32_t L, int32_t hd) {
  const int lane = threadIdx.x & 31;
  const int P = hd >> 3, hpi = 32 / P;
  const int hl = lane / P, cl = lane - hl * P;
  for (int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5); row < T; row += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t b = row / L;
    const int l = static_cast<int>(row - b * L);
    for (int h0 = 0; h0 < H; h0 += hpi) {
      const int h = h0 + hl;
      const bool on = hl < hpi && h < H;
      float acc = 0.f;
      if (on) {
        const int64_t off = row * ldo + static_cast<int64_t>(h) * hd + cl * 8;
        const uint4 a = *reinterpret_cast<const uint4*>(o + off);
        const uint4 g = *reinterpret_cast<const uint4*>(d_o + off);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 x = unpack_bf16x2(aw[k]), y = unpack_bf16x2(gw[k]);
          acc = fmaf(x.x, y.x, acc);
          acc = fmaf(x.y, y.y, acc);
        }
      }
#pragma unroll
      for (int st = 1; st < 16; st <<= 1) {
        const float v = __shfl_down_sync(0xffffffffu, acc, st);
        if (cl + st < P) acc += v;
      }
      if (on && cl == 0) dsum[(b * H + h) * L + l] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host: tiling plan + launch
// ---------------------------------------------------------------------------------------------------------------------
constexpr int A3_SMEM_MAX = 232448 - 1024;  // 227 KB opt-in shared memory per CTA minus the static part (barriers)

struct A3Plan {
  int Lk, n_slabs, n_outer, r_last, KB, n_blk, kb_last, NS, resident;
  size_t smem;
};

// inner-block size: blocks of <= cap columns, balanced, multiples of 16
static void a3_split_blocks(int Lk, int cap, int& KB, int& n_blk, int& kb_last) {
  n_blk = (Lk + cap - 1) / cap;
  KB = (((Lk + n_blk - 1) / n_blk) + 15) & ~15;
  if (KB > cap) KB = cap;
  n_blk = (Lk + KB - 1) / KB;
  kb_last = Lk - (n_blk - 1) * KB;
}

static bool a3_plan_fwd(int L, int hd, bool has_bias, A3Plan& pl) {
  pl.Lk = (L + 15) & ~15;
  pl.n_slabs = (hd + 63) / 64;
  pl.n_outer = (L + 127) / 128;
  pl.r_last = L - (pl.n_outer - 1) * 128;
  const size_t fixed = static_cast<size_t>(3) * pl.n_slabs * 16384 + 4096 + (has_bias ? 2 * pl.Lk * 4 : 0) + 1024;
  static const int caps[] = {16 * A3_FWD_NCH, 80, 64, 48, 32, 16};
  for (int cap : caps) {
    a3_split_blocks(pl.Lk, cap, pl.KB, pl.n_blk, pl.kb_last);
    const size_t stage = static_cast<size_t>(pl.n_slabs) * pl.KB * 128;
    if (pl.n_outer > 1 && 2 * pl.n_blk <= A3_MAX_STAGES && fixed + 4 * pl.n_blk * stage <= A3_SMEM_MAX) {
      pl.resident = 1;
      pl.NS = 2 * pl.n_blk;
      pl.smem = fixed + 4 * pl.n_blk * stage;
      return true;
    }
    const int ns = static_cast<int>(std::min<size_t>(A3_MAX_STAGES, (A3_SMEM_MAX - fixed) / (2 * stage)));
    if (ns >= 2 && (ns >= 3 || cap <= 80)) {
      pl.resident = 0;
      pl.NS = ns;
      pl.smem = fixed + 2 * ns * stage;
      return true;
    }
  }
  return false;
}

int attention_fwd_v3(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, void* o, int64_t ldo, float* lse,
                     const float* key_bias, int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale, float drop_p, uint64_t drop_seed,
                     cudaStream_t stream) {
  B200MM_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || L <= 65535), B200MM_ERR_SHAPE, "attention_fwd: drop_p=%f L=%d", drop_p, L);
  B200MM_REQUIRE(head_dim % 16 == 0 && head_dim >= 16 && head_dim <= 128, B200MM_ERR_SHAPE,
                 "attention_fwd: head_dim %d not supported (multiples of 16 up to 128)", head_dim);
  A3Plan pl;
  B200MM_REQUIRE(a3_plan_fwd(L, head_dim, key_bias != nullptr, pl), B200MM_ERR_SHAPE, "attention_fwd: no tiling for L=%d head_dim=%d", L,
                 head_dim);
  const int64_t T = static_cast<int64_t>(B) * L;
  CUtensorMap tmQ, tmQt, tmKV;
  int rc = make_tmap_2d_bf16(&tmQ, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmQt, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, pl.r_last);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmKV, qkv, static_cast<uint64_t>(ld), static_cast<uint64_t>(T), static_cast<uint64_t>(ld), 64, pl.KB);
  if (rc) return rc;
  A3FwdParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.ldo = ldo; p.lse = lse; p.key_bias = key_bias;
  p.B = B; p.H = H; p.L = L; p.Lk = pl.Lk; p.hd = head_dim; p.n_slabs = pl.n_slabs;
  p.n_qt = pl.n_outer; p.r_last = pl.r_last; p.KB = pl.KB; p.n_blk = pl.n_blk; p.kb_last = pl.kb_last;
  p.NS = pl.NS; p.resident = pl.resident; p.q_off = q_off; p.k_off = k_off; p.v_off = v_off;
  p.scale_log2 = scale * A3_LOG2E;
  const bool drop = drop_p > 0.f;
  p.drop_thr = drop ? drop_threshold(drop_p) : 0u;
  p.drop_inv_keep = drop ? 1.f / (1.f - drop_p) : 1.f;
  p.drop_seed = drop_seed;
  const int grid = std::min(B * H, sm_count());
  auto kern = drop ? (key_bias ? attn3_fwd_kernel<true, true> : attn3_fwd_kernel<false, true>)
                   : (key_bias ? attn3_fwd_kernel<true, false> : attn3_fwd_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem));
  if (e != cudaSuccess) {
    set_last_error("attention_fwd: cudaFuncSetAttribute(%zu): %s", pl.smem, cudaGetErrorString(e));
    return B200MM_ERR_LAUNCH;
  }
  kern<<<grid, A3_THREADS, pl.smem, stream>>>(tmQ, tmQt, tmKV, p);
  return check_launch("attn3_fwd_kernel");
}


static bool a3_plan_bwd(int mode, int L, int hd, A3Plan& pl) {
  pl.Lk = (L + 15) & ~15;
  pl.n_slabs = (hd + 63) / 64;
  pl.n_outer = (L + 127) / 128;
  pl.r_last = L - (pl.n_outer - 1) * 128;
  const size_t fixed = static_cast<size_t>(4) * pl.n_slabs * 16384 + 2 * pl.Lk * 4 + 1024;
  const int tmem_cap = ((512 - (mode == A3_MODE_DKV ? 2 : 1) * hd) / 4) & ~31;  // 4 NB + accumulators <= 512 columns, 32-column aligned
  static const int caps[] = {96, 64, 32};
  for (int cap : caps) {
    if (cap > tmem_cap) continue;
    a3_split_blocks(pl.Lk, cap, pl.KB, pl.n_blk, pl.kb_last);
    const size_t stage = static_cast<size_t>(2) * pl.n_slabs * pl.KB * 128;
    if (pl.n_outer > 1 && 2 * pl.n_blk <= A3_MAX_STAGES && fixed + 2 * pl.n_blk * stage <= A3_SMEM_MAX) {
      pl.resident = 1;
      pl.NS = 2 * pl.n_blk;
      pl.smem = fixed + 2 * pl.n_blk * stage;
      return true;
    }
    const int ns = static_cast<int>(std::min<size_t>(A3_MAX_STAGES, (A3_SMEM_MAX - fixed) / stage));
    if (ns >= 2) {
      pl.resident = 0;
      pl.NS = ns;
      pl.smem = fixed + ns * stage;
      return true;
    }
  }
  return false;
}

template <int MODE>
static int a3_launch_bwd(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* d_o, int64_t ldo,
                         const float* lse, const float* dsum, const float* key_bias, void* dqkv, int32_t B, int32_t H, int32_t L, int32_t hd,
                         float scale, float drop_p, uint64_t drop_seed, cudaStream_t stream) {
  A3Plan pl;
  B200MM_REQUIRE(a3_plan_bwd(MODE, L, hd, pl), B200MM_ERR_SHAPE, "attention_bwd: no tiling for L=%d head_dim=%d", L, hd);
  const uint64_t T = static_cast<uint64_t>(B) * L;
  // operand -> (matrix, pitch, column offset): dK/dV mode X = (K, V), Y = (Q, dO); dQ mode X = (Q, dO), Y = (K, V)
  const void* x0m = qkv; uint64_t x0ld = ld; int x0off = MODE == A3_MODE_DKV ? k_off : q_off;
  const void* x1m = MODE == A3_MODE_DKV ? qkv : d_o; uint64_t x1ld = MODE == A3_MODE_DKV ? ld : ldo; int x1off = MODE == A3_MODE_DKV ? v_off : 0;
  const void* y0m = qkv; uint64_t y0ld = ld; int y0off = MODE == A3_MODE_DKV ? q_off : k_off;
  const void* y1m = MODE == A3_MODE_DKV ? d_o : qkv; uint64_t y1ld = MODE == A3_MODE_DKV ? ldo : ld; int y1off = MODE == A3_MODE_DKV ? 0 : v_off;
  CUtensorMap tmX0, tmX0t, tmX1, tmX1t, tmY0, tmY1;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmX0, x0m, x0ld, T, x0ld, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmX0t, x0m, x0ld, T, x0ld, 64, pl.r_last))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmX1, x1m, x1ld, T, x1ld, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmX1t, x1m, x1ld, T, x1ld, 64, pl.r_last))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmY0, y0m, y0ld, T, y0ld, 64, pl.KB))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmY1, y1m, y1ld, T, y1ld, 64, pl.KB))) return rc;
  A3BwdParams p;
  p.lse = lse; p.dsum = dsum; p.key_bias = key_bias; p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv); p.ld = ld;
  p.B = B; p.H = H; p.L = L; p.Lk = pl.Lk; p.hd = hd; p.n_slabs = pl.n_slabs; p.n_outer = pl.n_outer; p.r_last = pl.r_last;
  p.NB = pl.KB; p.n_in = pl.n_blk; p.nb_last = pl.kb_last; p.NS = pl.NS; p.resident = pl.resident;
  p.x0_off = x0off; p.x1_off = x1off; p.y0_off = y0off; p.y1_off = y1off;
  p.q_off = q_off; p.k_off = k_off; p.v_off = v_off; p.scale = scale; p.scale_log2 = scale * A3_LOG2E;
  const bool drop = drop_p > 0.f;
  p.drop_thr = drop ? drop_threshold(drop_p) : 0u;
  p.drop_inv_keep = drop ? 1.f / (1.f - drop_p) : 1.f;
  p.drop_seed = drop_seed;
  const int grid = std::min(B * H, sm_count());
  auto kern = drop ? (key_bias ? attn3_bwd_kernel<MODE, true, true> : attn3_bwd_kernel<MODE, false, true>)
                   : (key_bias ? attn3_bwd_kernel<MODE, true, false> : attn3_bwd_kernel<MODE, false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem));
  if (e != cudaSuccess) {
    set_last_error("attention_bwd: cudaFuncSetAttribute(%zu): %s", pl.smem, cudaGetErrorString(e));
    return B200MM_ERR_LAUNCH;
  }
  kern<<<grid, A3_THREADS, pl.smem, stream>>>(tmX0, tmX0t, tmX1, tmX1t, tmY0, tmY1, p);
  return check_launch(MODE == A3_MODE_DKV ? "attn3_bwd_kernel<dkv>" : "attn3_bwd_kernel<dq>");
}

int attention_bwd_merged(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* d_o, int64_t ldo,
                         const float* lse, const float* key_bias, void* dqkv, const float* dsum, int32_t B, int32_t H, int32_t L, int32_t head_dim,
                         float scale, cudaStream_t stream);  // attention_tc.cu

int attention_bwd_v3(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* o, const void* d_o, int64_t ldo,
                     const float* lse, const float* key_bias, void* dqkv, float* dsum, int32_t B, int32_t H, int32_t L, int32_t head_dim,
                     float scale, float drop_p, uint64_t drop_seed, cudaStream_t stream) {
  B200MM_REQUIRE(head_dim % 16 == 0 && head_dim >= 16 && head_dim <= 128, B200MM_ERR_SHAPE,
                 "attention_bwd: head_dim %d not supported (multiples of 16 up to 128)", head_dim);
  B200MM_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || L <= 65535), B200MM_ERR_SHAPE, "attention_bwd: drop_p=%f L=%d", drop_p, L);
  const int64_t T = static_cast<int64_t>(B) * L;
  attn3_dsum_kernel<<<static_cast<unsigned>(std::min<int64_t>(ceil_div(T, 8), 148 * 64)), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(o),
                                                                                    reinterpret_cast<const __nv_bfloat16*>(d_o), ldo, dsum, T, H, L,
                                                                                    head_dim);
  int rc = check_launch("attn3_dsum_kernel");
  if (rc) return rc;
  // head_dim 64 and <= 288 tokens: dQ, dK and dV of an item fit the 512 TMEM columns together -> one merged kernel, 5 GEMMs and one exp
  // per score instead of the 7 + 2 of the recompute pair (attention_tc.cu)
  // (with dropout the mask is regenerated by the recompute pair; the merged kernel keeps the p = 0 arithmetic)
  if (drop_p == 0.f) {
    rc = attention_bwd_merged(qkv, ld, q_off, k_off, v_off, d_o, ldo, lse, key_bias, dqkv, dsum, B, H, L, head_dim, scale, stream);
    if (rc != 0) return rc < 0 ? rc : B200MM_OK;
  }
  rc = a3_launch_bwd<A3_MODE_DKV>(qkv, ld, q_off, k_off, v_off, d_o, ldo, lse, dsum, key_bias, dqkv, B, H, L, head_dim, scale, drop_p, drop_seed,
                                  stream);
  if (rc) return rc;
  return a3_launch_bwd<A3_MODE_DQ>(qkv, ld, q_off, k_off, v_off, d_o, ldo, lse, dsum, key_bias, dqkv, B, H, L, head_dim, scale, drop_p, drop_seed,
                                   stream);
}

}  // namespace b200mm

// ---------------------------------------------------------------------------------------------------------------------
// C-ABI (include/b200mm.h)
// ---------------------------------------------------------------------------------------------------------------------
using namespace b200mm;

static int attn_check_common(const char* who, const void* qkv, int64_t ld, const void* o, int64_t ldo, int32_t B, int32_t H, int32_t L, int32_t hd,
                             int32_t q_off, int32_t k_off, int32_t v_off) {
  B200MM_REQUIRE(B > 0 && H > 0 && L > 0, B200MM_ERR_SHAPE, "%s: B=%d H=%d L=%d", who, B, H, L);
  B200MM_REQUIRE(static_cast<int64_t>(B) * H < (1ll << 31) && static_cast<int64_t>(B) * L < (1ll << 31), B200MM_ERR_SHAPE,
                 "%s: B*H or B*L too large", who);
  B200MM_REQUIRE(qkv && o, B200MM_ERR_SHAPE, "%s: null pointer", who);
  B200MM_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0 && hd % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0,
                 B200MM_ERR_ALIGN, "%s: pitches/offsets must be multiples of 8 elements and bases 16B aligned", who);
  return B200MM_OK;
}

extern "C" int b200mm_attention_fwd_dropout(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, void* o, int64_t ldo,
                                            float* lse, const float* key_bias, int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale,
                                            float drop_p, uint64_t drop_seed, void* stream) {
  int rc = attn_check_common("attention_fwd", qkv, ld, o, ldo, B, H, L, head_dim, q_off, k_off, v_off);
  if (rc) return rc;
  B200MM_REQUIRE(lse != nullptr, B200MM_ERR_SHAPE, "attention_fwd: lse is required");
  return attention_fwd_v3(qkv, ld, q_off, k_off, v_off, o, ldo, lse, key_bias, B, H, L, head_dim, scale, drop_p, drop_seed,
                          reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b200mm_attention_fwd(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, void* o, int64_t ldo, float* lse,
                                    const float* key_bias, int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale, void* stream) {
  return b200mm_attention_fwd_dropout(qkv, ld, q_off, k_off, v_off, o, ldo, lse, key_bias, B, H, L, head_dim, scale, 0.f, 0, stream);
}

extern "C" int b200mm_attention_bwd_dropout(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* o,
                                            const void* d_o, int64_t ldo, const float* lse, const float* key_bias, void* dqkv, float* dsum,
                                            int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale, float drop_p, uint64_t drop_seed,
                                            void* stream) {
  int rc = attn_check_common("attention_bwd", qkv, ld, o, ldo, B, H, L, head_dim, q_off, k_off, v_off);
  if (rc) return rc;
  B200MM_REQUIRE(lse && d_o && dqkv && dsum, B200MM_ERR_SHAPE, "attention_bwd: null pointer");
  B200MM_REQUIRE((reinterpret_cast<uintptr_t>(d_o) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0, B200MM_ERR_ALIGN,
                 "attention_bwd: d_o/dqkv must be 16B aligned");
  return attention_bwd_v3(qkv, ld, q_off, k_off, v_off, o, d_o, ldo, lse, key_bias, dqkv, dsum, B, H, L, head_dim, scale, drop_p, drop_seed,
                          reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b200mm_attention_bwd(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* o, const void* d_o,
                                    int64_t ldo, const float* lse, const float* key_bias, void* dqkv, float* dsum, int32_t B, int32_t H,
                                    int32_t L, int32_t head_dim, float scale, void* stream) {
  return b200mm_attention_bwd_dropout(qkv, ld, q_off, k_off, v_off, o, d_o, ldo, lse, key_bias, dqkv, dsum, B, H, L, head_dim, scale, 0.f, 0,
                                      stream);
}
