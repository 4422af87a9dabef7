// b200mm — small HBM-bound helpers around the GEMMs: activation recompute, bias / positional-embedding gradient
// reductions, embedding scatter-add, L2 row normalisation, dtype casts, patch extraction (im2row).
#include "common.cuh"

#include <algorithm>

namespace b200mm {

__device__ __forceinline__ void unpack8e(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8e(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// y = act(x), n8 = number of 8-element vectors
__global__ void __launch_bounds__(256) act_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t n8, int act) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n8; i += gridDim.x * 256ll) {
    float v[8];
    unpack8e(x[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(act, v[j]);
    y[i] = pack8e(v);
  }
}

// y[r, c] = keep(seed, r, c) ? x[r, c] * inv_keep : 0 — inverted dropout with the counter-based mask of common.cuh; one thread per
// 8 consecutive columns of a row (one 16-byte access each way), the row key is hashed once per thread
__global__ void __launch_bounds__(256) dropout_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y, int64_t ldy,
                                                      int64_t rows, int32_t c8, uint32_t thr, float inv_keep, uint64_t seed) {
  const int64_t total = rows * c8;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int64_t r = i / c8;
    const uint32_t c = static_cast<uint32_t>(i - r * c8) * 8u;
    const uint32_t key = drop_stream_key(seed, static_cast<uint64_t>(r));
    float v[8];
    unpack8e(*reinterpret_cast<const uint4*>(x + r * ldx + c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = drop_keep(key, c + j, thr) ? v[j] * inv_keep : 0.f;
    *reinterpret_cast<uint4*>(y + r * ldy + c) = pack8e(v);
  }
}

// keep[b, h, q, k] = 1 iff the attention kernels keep probability (q, k) of item (b, h) under this seed (test / debugging aid)
__global__ void __launch_bounds__(256) attention_dropout_mask_kernel(uint8_t* __restrict__ keep, int64_t items, int32_t L, uint32_t thr, uint64_t seed) {
  const int64_t per = static_cast<int64_t>(L) * L, total = items * per;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int64_t item = i / per;
    keep[i] = drop_keep(drop_stream_key(seed, static_cast<uint64_t>(item)), static_cast<uint32_t>(i - item * per), thr) ? 1 : 0;
  }
}

// out[(row % period), col] += in[row, col]   (fp32 atomics; caller zero-fills out)
//   period == 1 : bias gradient   db[n] = sum_m dY[m, n]
//   period == L : positional-embedding gradient  dpos[l, :] = sum_b ds[b*L + l, :]
// block = 32 column groups (8 cols each) x 8 row lanes; grid = (col blocks, period, repeat chunks)
__global__ void __launch_bounds__(256) rowsum_periodic_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                                              int64_t rows, int32_t W, int64_t period, int64_t reps_per_chunk) {
  __shared__ float red[8][32 * 8 + 1];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cg) * 8;
  const int64_t pidx = blockIdx.y;
  const int64_t reps = rows / period;
  const int64_t r0 = blockIdx.z * reps_per_chunk, r1 = min(reps, r0 + reps_per_chunk);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < W) {
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      float v[8];
      unpack8e(*reinterpret_cast<const uint4*>(in + (r * period + pidx) * W + col), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  const int ocol = blockIdx.x * 256 + c;
  if (ocol < W) atomicAdd(out + pidx * W + ocol, s);
}

// out[ids[row], :] += in[row, :] (fp32 atomics), rows with ids == skip_id are dropped (nn.Embedding padding_idx)
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const __nv_bfloat16* __restrict__ in, const int64_t* __restrict__ ids,
                                                               float* __restrict__ out, int64_t rows, int32_t W, int64_t skip_id,
                                                               int64_t n_out_rows) {
  const int64_t w8 = W / 8;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < rows * w8; i += gridDim.x * 256ll) {
    const int64_t row = i / w8;
    const int col = static_cast<int>(i - row * w8) * 8;
    const int64_t id = ids[row];
    if (id == skip_id || id < 0 || id >= n_out_rows) continue;
    float v[8];
    unpack8e(*reinterpret_cast<const uint4*>(in + row * W + col), v);
    float* o = out + id * W + col;
    // two 16-byte vector reductions instead of eight scalar atomics
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
  }
}

// same operation for a handful of output rows (token_type_embeddings: 2): every input row would hit the same few addresses, so each
// thread keeps one accumulator set per output row over its slice of the input and issues NOUT x 2 vector reductions at the end.
// block = 32 column groups (8 cols) x 8 row lanes; grid = (column blocks, row chunks)
template <int NOUT>
__global__ void __launch_bounds__(256) scatter_add_few_kernel(const __nv_bfloat16* __restrict__ in, const int64_t* __restrict__ ids,
                                                              float* __restrict__ out, int64_t rows, int32_t W, int64_t skip_id,
                                                              int64_t rows_per_chunk, int32_t n_out) {
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cg) * 8;
  if (col >= W) return;
  const int64_t r0 = blockIdx.y * rows_per_chunk;
  const int64_t r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  float acc[NOUT][8];
#pragma unroll
  for (int o = 0; o < NOUT; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
  for (int64_t row = r0 + rl; row < r1; row += 8) {
    const int64_t id = ids[row];
    if (id == skip_id || id < 0 || id >= n_out) continue;
    float v[8];
    unpack8e(*reinterpret_cast<const uint4*>(in + row * W + col), v);
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (id == o) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o][j] += v[j];
      }
  }
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    if (o >= n_out) break;
    float* dst = out + static_cast<int64_t>(o) * W + col;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc[o][0]), "f"(acc[o][1]), "f"(acc[o][2]), "f"(acc[o][3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(acc[o][4]), "f"(acc[o][5]), "f"(acc[o][6]), "f"(acc[o][7]) : "memory");
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n,
                                                            float scale) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) y[i] = __float2bfloat16(x[i] * scale);
}

// L2 row normalisation  y = x / ||x||   (cn_model.py:217-218, F.normalize in univl_video_base.py:114,158)
// one warp per row; inv_norm saved for backward
__global__ void __launch_bounds__(256) rownorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          float* __restrict__ inv_norm, int64_t rows, int32_t W, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  float sq = 0.f;
  for (int c = lane * 8; c < W; c += 256) {
    float v[8];
    unpack8e(*reinterpret_cast<const uint4*>(x + row * W + c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) sq += v[j] * v[j];
  }
  sq = warp_sum(sq);
  const float inv = 1.f / fmaxf(sqrtf(sq), eps);
  if (lane == 0) inv_norm[row] = inv;
  for (int c = lane * 8; c < W; c += 256) {
    float v[8];
    unpack8e(*reinterpret_cast<const uint4*>(x + row * W + c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= inv;
    *reinterpret_cast<uint4*>(y + row * W + c) = pack8e(v);
  }
}

// dx = (dy - yhat * <yhat, dy>) * inv_norm, yhat = x * inv_norm recomputed in fp32 from x
__global__ void __launch_bounds__(256) rownorm_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                          const float* __restrict__ inv_norm, __nv_bfloat16* __restrict__ dx,
                                                          int64_t rows, int32_t W) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float inv = inv_norm[row];
  float dot = 0.f;
  for (int c = lane * 8; c < W; c += 256) {
    float v[8];
    unpack8e(*reinterpret_cast<const uint4*>(x + row * W + c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) dot += v[j] * inv * dy[row * W + c + j];
  }
  dot = warp_sum(dot);
  for (int c = lane * 8; c < W; c += 256) {
    float v[8], o[8];
    unpack8e(*reinterpret_cast<const uint4*>(x + row * W + c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (dy[row * W + c + j] - v[j] * inv * dot) * inv;
    *reinterpret_cast<uint4*>(dx + row * W + c) = pack8e(o);
  }
}

// Patch extraction for the ViT stem (conv with stride == kernel == p, clip/model.py:289-295,310-312):
//   out[b*(Np+1) + 1 + (gy*g + gx), (c*p + dy)*p + dx] = image[b, c, gy*p + dy, gx*p + dx]
//   row b*(Np+1) + 0 (the class-token slot) and the K padding columns [3*p*p, Kp) are zero, so that the patch GEMM
//   writes straight into the [B, Np+1, width] token buffer.
__global__ void __launch_bounds__(256) im2row_kernel(const __nv_bfloat16* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                     int64_t B, int32_t C, int32_t H, int32_t Wd, int32_t p, int32_t Kp) {
  const int g_y = H / p, g_x = Wd / p;
  const int64_t L = static_cast<int64_t>(g_y) * g_x + 1;
  const int64_t total = B * L * Kp;
  const int K = C * p * p;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int k = static_cast<int>(i % Kp);
    const int64_t row = i / Kp;
    const int64_t b = row / L;
    const int l = static_cast<int>(row - b * L);
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (l > 0 && k < K) {
      const int pi = l - 1;
      const int gy = pi / g_x, gx = pi - gy * g_x;
      const int c = k / (p * p);
      const int rem = k - c * p * p;
      const int dy = rem / p, dx = rem - dy * p;
      v = img[((b * C + c) * H + gy * p + dy) * Wd + gx * p + dx];
    }
    out[i] = v;
  }
}

// out[r] = scale * <a[r,:], b[r,:]>  (MoCo positive logits: einsum("bh,bh->b"), univl_video_ret.py:292-296); one warp per row
__global__ void __launch_bounds__(256) rowdot_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                     float* __restrict__ out, int64_t rows, int32_t W, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  float acc = 0.f;
  for (int c = lane * 8; c < W; c += 256) {
    float x[8], y[8];
    unpack8e(*reinterpret_cast<const uint4*>(a + row * W + c), x);
    unpack8e(*reinterpret_cast<const uint4*>(b + row * W + c), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(x[j], y[j], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc * scale;
}

// momentum (EMA) update of a key-encoder parameter kept in fp32: pk = m*pk + (1-m)*pq  (moco_utils.py:55-69).
// fp32 master copy on purpose: with m = 0.9999 the increment is below bf16 resolution.
template <typename TQ>
__global__ void __launch_bounds__(256) ema_update_kernel(float* __restrict__ pk, const TQ* __restrict__ pq, int64_t n, float m) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    float q;
    if constexpr (sizeof(TQ) == 2) q = __bfloat162float(pq[i]);
    else q = pq[i];
    pk[i] = fmaf(pk[i], m, q * (1.f - m));
  }
}

static inline int grid_for(int64_t work_items, int threads = 256, int waves = 8) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(work_items, threads), static_cast<int64_t>(sm_count()) * waves)));
}

// masked mean over the P positions (frames x grid cells) of each clip: y[r, :] = sum_p valid[r,p] * x[r,p,:] / count[r]
// (frame pooling of UnivlVideoBase.forward_img_encoder, prj/base_vtp/roi_univl/univl/model/univl_video_base.py:91-95);
// pad[r,p] != 0 marks a padded position; inv_count[r] = 1 / #valid is kept for backward. One thread per 8 columns.
__global__ void __launch_bounds__(256) masked_mean_fwd_kernel(const uint4* __restrict__ x, const uint8_t* __restrict__ pad, uint4* __restrict__ y,
                                                              float* __restrict__ inv_count, int64_t R, int32_t P, int32_t W8) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < R * W8; i += gridDim.x * 256ll) {
    const int64_t r = i / W8;
    const int c = static_cast<int>(i - r * W8);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    for (int p = 0; p < P; ++p) {
      if (pad != nullptr && pad[r * P + p]) continue;
      ++cnt;
      float v[8];
      unpack8e(x[(r * P + p) * W8 + c], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    const float inv = 1.f / static_cast<float>(cnt);  // an all-padded clip gives inf/nan like the reference's 0/0
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    y[i] = pack8e(acc);
    if (c == 0) inv_count[r] = inv;
  }
}

// dx[r,p,:] = valid[r,p] * inv_count[r] * dy[r,:]
__global__ void __launch_bounds__(256) masked_mean_bwd_kernel(const uint4* __restrict__ dy, const uint8_t* __restrict__ pad,
                                                              const float* __restrict__ inv_count, uint4* __restrict__ dx, int64_t R, int32_t P,
                                                              int32_t W8) {
  const int64_t total = R * P * W8;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int64_t rp = i / W8;
    const int c = static_cast<int>(i - rp * W8);
    const int64_t r = rp / P;
    float v[8];
    if (pad != nullptr && pad[rp]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    } else {
      unpack8e(dy[r * W8 + c], v);
      const float inv = inv_count[r];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= inv;
    }
    dx[i] = pack8e(v);
  }
}

}  // namespace b200mm

using namespace b200mm;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int b200mm_act_fwd(const void* x, void* y, int64_t n, int32_t act, void* stream) {
  B200MM_REQUIRE(n >= 0 && n % 8 == 0, B200MM_ERR_SHAPE, "act_fwd: n=%lld must be a multiple of 8", (long long)n);
  if (n == 0) return B200MM_OK;
  B200MM_REQUIRE(ALIGNED16(x) && ALIGNED16(y), B200MM_ERR_ALIGN, "act_fwd: pointers must be 16B aligned");
  act_fwd_kernel<<<grid_for(n / 8), 256, 0, STREAM(stream)>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), n / 8, act);
  return check_launch("act_fwd_kernel");
}

extern "C" int b200mm_dropout(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int32_t cols, float p, uint64_t seed, void* stream) {
  B200MM_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= cols && ldy >= cols, B200MM_ERR_SHAPE,
                 "dropout: rows=%lld cols=%d ldx=%lld ldy=%lld (cols and pitches must be multiples of 8)", (long long)rows, cols, (long long)ldx,
                 (long long)ldy);
  B200MM_REQUIRE(p >= 0.f && p < 1.f, B200MM_ERR_SHAPE, "dropout: p=%f must be in [0, 1)", p);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(x && y && ALIGNED16(x) && ALIGNED16(y), B200MM_ERR_ALIGN, "dropout: pointers must be 16B aligned");
  dropout_kernel<<<grid_for(rows * (cols / 8)), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__nv_bfloat16*>(y),
                                                                         ldy, rows, cols / 8, drop_threshold(p), 1.f / (1.f - p), seed);
  return check_launch("dropout_kernel");
}

extern "C" int b200mm_attention_dropout_mask(uint8_t* keep, int32_t B, int32_t H, int32_t L, float drop_p, uint64_t drop_seed, void* stream) {
  B200MM_REQUIRE(keep && B > 0 && H > 0 && L > 0 && L <= 65535 && drop_p >= 0.f && drop_p < 1.f, B200MM_ERR_SHAPE,
                 "attention_dropout_mask: B=%d H=%d L=%d p=%f", B, H, L, drop_p);
  const int64_t items = static_cast<int64_t>(B) * H;
  attention_dropout_mask_kernel<<<grid_for(items * L * L), 256, 0, STREAM(stream)>>>(keep, items, L, drop_threshold(drop_p), drop_seed);
  return check_launch("attention_dropout_mask_kernel");
}

extern "C" int b200mm_rowsum_periodic(const void* in, float* out, int64_t rows, int32_t W, int64_t period, void* stream) {
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0 && period > 0 && rows % period == 0 && period <= 65535, B200MM_ERR_SHAPE,
                 "rowsum_periodic: rows=%lld W=%d period=%lld", (long long)rows, W, (long long)period);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(ALIGNED16(in), B200MM_ERR_ALIGN, "rowsum_periodic: input must be 16B aligned");
  const int64_t reps = rows / period;
  const int col_blocks = static_cast<int>(ceil_div(W, 256));
  int64_t want_chunks = std::max<int64_t>(1, static_cast<int64_t>(sm_count()) * 4 / (col_blocks * period));
  want_chunks = std::min<int64_t>(want_chunks, std::max<int64_t>(1, reps / 32));
  want_chunks = std::min<int64_t>(want_chunks, 65535);
  const int64_t per = ceil_div(reps, want_chunks);
  dim3 grid(col_blocks, static_cast<unsigned>(period), static_cast<unsigned>(ceil_div(reps, per)));
  rowsum_periodic_kernel<<<grid, 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, rows, W, period, per);
  return check_launch("rowsum_periodic_kernel");
}

extern "C" int b200mm_scatter_add_rows(const void* in, const int64_t* ids, float* out, int64_t rows, int32_t W, int64_t skip_id,
                                       int64_t n_out_rows, void* stream) {
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "scatter_add_rows: rows=%lld W=%d", (long long)rows, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(ALIGNED16(in) && ALIGNED16(out), B200MM_ERR_ALIGN, "scatter_add_rows: input and output must be 16B aligned");
  if (n_out_rows <= 4) {
    const int64_t rows_per_chunk = 512;
    dim3 grid(static_cast<unsigned>((W / 8 + 31) / 32), static_cast<unsigned>((rows + rows_per_chunk - 1) / rows_per_chunk));
    if (n_out_rows <= 2)
      scatter_add_few_kernel<2><<<grid, 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(in), ids, out, rows, W, skip_id, rows_per_chunk,
                                                                  static_cast<int32_t>(n_out_rows));
    else
      scatter_add_few_kernel<4><<<grid, 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(in), ids, out, rows, W, skip_id, rows_per_chunk,
                                                                  static_cast<int32_t>(n_out_rows));
    return check_launch("scatter_add_few_kernel");
  }
  scatter_add_rows_kernel<<<grid_for(rows * (W / 8)), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(in), ids, out, rows,
                                                                                  W, skip_id, n_out_rows);
  return check_launch("scatter_add_rows_kernel");
}

extern "C" int b200mm_cast_f32_bf16(const float* x, void* y, int64_t n, float scale, void* stream) {
  if (n <= 0) return B200MM_OK;
  cast_f32_bf16_kernel<<<grid_for(n), 256, 0, STREAM(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n, scale);
  return check_launch("cast_f32_bf16_kernel");
}

extern "C" int b200mm_rownorm_fwd(const void* x, void* y, float* inv_norm, int64_t rows, int32_t W, float eps, void* stream) {
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "rownorm_fwd: rows=%lld W=%d", (long long)rows, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(ALIGNED16(x) && ALIGNED16(y), B200MM_ERR_ALIGN, "rownorm_fwd: pointers must be 16B aligned");
  rownorm_fwd_kernel<<<static_cast<int>(ceil_div(rows, 8)), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), inv_norm, rows, W, eps);
  return check_launch("rownorm_fwd_kernel");
}

extern "C" int b200mm_rownorm_bwd(const float* dy, const void* x, const float* inv_norm, void* dx, int64_t rows, int32_t W, void* stream) {
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "rownorm_bwd: rows=%lld W=%d", (long long)rows, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(ALIGNED16(x) && ALIGNED16(dx), B200MM_ERR_ALIGN, "rownorm_bwd: pointers must be 16B aligned");
  rownorm_bwd_kernel<<<static_cast<int>(ceil_div(rows, 8)), 256, 0, STREAM(stream)>>>(
      dy, reinterpret_cast<const __nv_bfloat16*>(x), inv_norm, reinterpret_cast<__nv_bfloat16*>(dx), rows, W);
  return check_launch("rownorm_bwd_kernel");
}

extern "C" int b200mm_im2row(const void* img, void* out, int64_t B, int32_t C, int32_t H, int32_t Wd, int32_t p, int32_t Kp, void* stream) {
  B200MM_REQUIRE(B >= 0 && C > 0 && p > 0 && H % p == 0 && Wd % p == 0 && Kp >= C * p * p && Kp % 8 == 0, B200MM_ERR_SHAPE,
                 "im2row: B=%lld C=%d H=%d W=%d p=%d Kp=%d", (long long)B, C, H, Wd, p, Kp);
  if (B == 0) return B200MM_OK;
  const int64_t L = static_cast<int64_t>(H / p) * (Wd / p) + 1;
  im2row_kernel<<<grid_for(B * L * Kp), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(img),
                                                                 reinterpret_cast<__nv_bfloat16*>(out), B, C, H, Wd, p, Kp);
  return check_launch("im2row_kernel");
}

extern "C" int b200mm_rowdot(const void* a, const void* b, float* out, int64_t rows, int32_t W, float scale, void* stream) {
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "rowdot: rows=%lld W=%d", (long long)rows, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(ALIGNED16(a) && ALIGNED16(b), B200MM_ERR_ALIGN, "rowdot: pointers must be 16B aligned");
  rowdot_kernel<<<static_cast<int>(ceil_div(rows, 8)), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(a),
                                                                                 reinterpret_cast<const __nv_bfloat16*>(b), out, rows, W, scale);
  return check_launch("rowdot_kernel");
}

extern "C" int b200mm_ema_update(float* pk, const void* pq, int32_t pq_is_bf16, int64_t n, float m, void* stream) {
  if (n <= 0) return B200MM_OK;
  B200MM_REQUIRE(pk && pq, B200MM_ERR_SHAPE, "ema_update: null pointer");
  if (pq_is_bf16)
    ema_update_kernel<<<grid_for(n), 256, 0, STREAM(stream)>>>(pk, reinterpret_cast<const __nv_bfloat16*>(pq), n, m);
  else
    ema_update_kernel<<<grid_for(n), 256, 0, STREAM(stream)>>>(pk, reinterpret_cast<const float*>(pq), n, m);
  return check_launch("ema_update_kernel");
}

extern "C" int b200mm_masked_mean_fwd(const void* x, const uint8_t* pad, void* y, float* inv_count, int64_t R, int32_t P, int32_t W, void* stream) {
  B200MM_REQUIRE(x && y && inv_count && R > 0 && P > 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "masked_mean_fwd: R=%lld P=%d W=%d (W %% 8 == 0)",
                 (long long)R, P, W);
  const int64_t n = R * (W / 8);
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  masked_mean_fwd_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4*>(x), pad, reinterpret_cast<uint4*>(y),
                                                                                     inv_count, R, P, W / 8);
  return check_launch("masked_mean_fwd_kernel");
}

extern "C" int b200mm_masked_mean_bwd(const void* dy, const uint8_t* pad, const float* inv_count, void* dx, int64_t R, int32_t P, int32_t W,
                                      void* stream) {
  B200MM_REQUIRE(dy && dx && inv_count && R > 0 && P > 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "masked_mean_bwd: R=%lld P=%d W=%d (W %% 8 == 0)",
                 (long long)R, P, W);
  const int64_t n = R * P * (W / 8);
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  masked_mean_bwd_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4*>(dy), pad, inv_count,
                                                                                     reinterpret_cast<uint4*>(dx), R, P, W / 8);
  return check_launch("masked_mean_bwd_kernel");
}
