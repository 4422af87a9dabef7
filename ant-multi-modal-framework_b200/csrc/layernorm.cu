// b200mm — fused LayerNorm forward / backward (HBM-bound; one warp per row, 16-byte accesses, fp32 statistics).
//
// Reference arithmetic: antmmf/modules/vision/backbone/clip/model.py:213-219 (ViT LayerNorm computes in fp32, eps 1e-5)
// and torch.nn.LayerNorm as used by modeling_bert.py:63,83,179,231 (eps 1e-12).
//
// forward : y = (s - mean) * rstd * w + b,  s = x (+ add0[row % add0_period] + add1[row is first of its period])
//           `add0`/`add1` implement the ViT stem  x + positional_embedding (+ class_embedding on token 0)
//           (clip/model.py:313-324) and the BERT embedding sum when rows are pre-gathered; s is optionally written out.
// backward: dx = rstd * (dy*w - mean(dy*w) - xhat * mean(dy*w*xhat)) (+ dadd),  dw += dy*xhat, db += dy (fp32 atomics)
//
// Algorithmic bytes per row of width W (bf16): fwd 2W read + 2W write (+2W if s is written); bwd 4W read + 2W write.
#include "common.cuh"

#include <algorithm>
#include <type_traits>

namespace b200mm {

constexpr int LN_WARPS = 8;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

struct LnFwdParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* add0;  // [add_period, W] or null
  const __nv_bfloat16* add1;  // [W] added to rows with row % add_period == 0, or null
  int64_t add_period;
  const int64_t* gather_ids;   // if set, row r of x is x[gather_ids[r]] (embedding lookup, bit-exact int64 indexing)
  const __nv_bfloat16* add2;   // [n, W] table indexed by add2_ids[row] (token-type embedding), or null
  const int64_t* add2_ids;
  const __nv_bfloat16* w;
  const __nv_bfloat16* b;
  __nv_bfloat16* y;
  __nv_bfloat16* s_out;  // pre-LN sum, or null
  float* mean;
  float* rstd;
  int64_t rows;
  int32_t W;
  float eps;
};

// VPL = 16-byte vectors per lane; lane l, slot i covers columns (i*32 + l)*8 .. +7
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32, VPL <= 4 ? 2 : 1) ln_fwd_kernel(const LnFwdParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  uint4 wq[VPL], bq[VPL];  // affine parameters stay packed (bf16) in registers: occupancy matters more than the unpack cost
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 8;
    wq[i] = bq[i] = make_uint4(0u, 0u, 0u, 0u);
    if (col < p.W) {
      wq[i] = *reinterpret_cast<const uint4*>(p.w + col);
      bq[i] = *reinterpret_cast<const uint4*>(p.b + col);
    }
  }
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp; row < p.rows; row += warps_total) {
    float v[VPL][8];
    float sum = 0.f;
    const int64_t prow = p.add_period > 0 ? row % p.add_period : 0;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 8;
      if (col < p.W) {
        const int64_t xrow = p.gather_ids != nullptr ? p.gather_ids[row] : row;
        unpack8(*reinterpret_cast<const uint4*>(p.x + xrow * p.W + col), v[i]);
        if (p.add2 != nullptr) {
          float a[8];
          unpack8(*reinterpret_cast<const uint4*>(p.add2 + p.add2_ids[row] * p.W + col), a);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += a[j];
        }
        if (p.add0 != nullptr) {
          float a[8];
          unpack8(*reinterpret_cast<const uint4*>(p.add0 + prow * p.W + col), a);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += a[j];
        }
        if (p.add1 != nullptr && prow == 0) {
          float a[8];
          unpack8(*reinterpret_cast<const uint4*>(p.add1 + col), a);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += a[j];
        }
        if (p.s_out != nullptr) {
          // the stored sum is what backward re-reads, so normalise exactly the bf16-rounded values
          uint4 o = pack8(v[i]);
          *reinterpret_cast<uint4*>(p.s_out + row * p.W + col) = o;
          unpack8(o, v[i]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i][j];
      }
    }
    const float mean = warp_sum(sum) / p.W;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 8;
      if (col < p.W) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / p.W + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[row] = mean;
      if (p.rstd) p.rstd[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 8;
      if (col < p.W) {
        float o[8], wv[8], bv[8];
        unpack8(wq[i], wv);
        unpack8(bq[i], bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * wv[j] + bv[j];
        *reinterpret_cast<uint4*>(p.y + row * p.W + col) = pack8(o);
      }
    }
  }
}

// Lean forward for the common case (no fused adds, no gather, row width an exact multiple of 256): ~8 instructions per
// element instead of ~23 in the general kernel (ncu: the general kernel is issue-bound, not HBM-bound).
__device__ __forceinline__ void unpack8_fast(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_plain_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                                     const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows,
                                                                     float eps) {
  constexpr int W = VPL * 256;
  constexpr float inv_w = 1.f / W;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float wv[VPL][8], bv[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    unpack8_fast(*reinterpret_cast<const uint4*>(w + i * 256 + lane * 8), wv[i]);
    unpack8_fast(*reinterpret_cast<const uint4*>(b + i * 256 + lane * 8), bv[i]);
  }
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp;
  const int64_t row_stride = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  const __nv_bfloat16* xr = x + row0 * W + lane * 8;
  __nv_bfloat16* yr = y + row0 * W + lane * 8;
  for (int64_t row = row0; row < rows; row += row_stride, xr += row_stride * W, yr += row_stride * W) {
    uint4 q[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) q[i] = *reinterpret_cast<const uint4*>(xr + i * 256);
    float v[VPL][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      unpack8_fast(q[i], v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
    const float mean = warp_sum(sum) * inv_w;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] -= mean;
        sq = fmaf(v[i][j], v[i][j], sq);
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_w + eps);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(v[i][j] * rstd, wv[i][j], bv[i][j]);
      *reinterpret_cast<uint4*>(yr + i * 256) = pack8(o);
    }
  }
}

struct LnBwdParams {
  const __nv_bfloat16* dy;
  const __nv_bfloat16* x;  // the LN input (the pre-LN sum)
  const float* mean;
  const float* rstd;
  const __nv_bfloat16* w;
  const __nv_bfloat16* dadd;  // optional gradient added to dx (residual branch), [rows, W]
  __nv_bfloat16* dx;
  float* dw;  // [W] fp32, accumulated with atomics (caller zero-fills)
  float* db;
  int64_t rows;
  int32_t W;
};

template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const LnBwdParams p) {
  __shared__ float red[LN_WARPS][32 * 8 + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  uint4 wq[VPL];
  float dwv[VPL][8], dbv[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 8;
    wq[i] = make_uint4(0u, 0u, 0u, 0u);
    if (col < p.W) wq[i] = *reinterpret_cast<const uint4*>(p.w + col);
#pragma unroll
    for (int j = 0; j < 8; ++j) { dwv[i][j] = 0.f; dbv[i][j] = 0.f; }
  }
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp; row < p.rows; row += warps_total) {
    const float mean = p.mean[row], rstd = p.rstd[row];
    uint4 xq[VPL], dyq[VPL];  // kept packed; xhat and dy*w are recomputed in the second pass (register diet -> 2 CTAs/SM)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 8;
      xq[i] = dyq[i] = make_uint4(0u, 0u, 0u, 0u);
      if (col < p.W) {
        xq[i] = *reinterpret_cast<const uint4*>(p.x + row * p.W + col);
        dyq[i] = *reinterpret_cast<const uint4*>(p.dy + row * p.W + col);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 8;
      if (col < p.W) {
        float xv[8], dyv[8], wv[8];
        unpack8(xq[i], xv);
        unpack8(dyq[i], dyv);
        unpack8(wq[i], wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float g = dyv[j] * wv[j];
          s1 += g;
          s2 += g * xh;
          dwv[i][j] += dyv[j] * xh;
          dbv[i][j] += dyv[j];
        }
      }
    }
    s1 = warp_sum(s1) / p.W;
    s2 = warp_sum(s2) / p.W;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 8;
      if (col < p.W) {
        float xv[8], dyv[8], wv[8], o[8];
        unpack8(xq[i], xv);
        unpack8(dyq[i], dyv);
        unpack8(wq[i], wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dyv[j] * wv[j] - s1 - (xv[j] - mean) * rstd * s2);
        if (p.dadd != nullptr) {
          float a[8];
          unpack8(*reinterpret_cast<const uint4*>(p.dadd + row * p.W + col), a);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += a[j];
        }
        *reinterpret_cast<uint4*>(p.dx + row * p.W + col) = pack8(o);
      }
    }
  }
  // cross-warp reduction of the per-lane column partials, then one atomic per column per CTA
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = pass == 0 ? dwv[i][j] : dbv[i][j];
      __syncthreads();
      const int c = threadIdx.x;  // 256 threads <-> 256 columns of this slot
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) acc += red[w][c];
      const int col = i * 256 + c;
      if (col < p.W) atomicAdd((pass == 0 ? p.dw : p.db) + col, acc);
    }
  }
}

// Lean backward for rows of exactly VPL*256 columns: no predicates, hoisted pointers, shift-based bf16 unpack.
template <int VPL, bool HAS_DADD>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_plain_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                                     const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ dadd,
                                                                     __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                                                                     int64_t rows) {
  __shared__ float red[LN_WARPS][32 * 8 + 1];
  constexpr int W = VPL * 256;
  constexpr float inv_w = 1.f / W;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float wv[VPL][8], dwv[VPL][8], dbv[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    unpack8_fast(*reinterpret_cast<const uint4*>(w + i * 256 + lane * 8), wv[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { dwv[i][j] = 0.f; dbv[i][j] = 0.f; }
  }
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp;
  const int64_t row_stride = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  int64_t off = row0 * W + lane * 8;
  for (int64_t row = row0; row < rows; row += row_stride, off += row_stride * W) {
    uint4 xq[VPL], dyq[VPL], daq[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xq[i] = *reinterpret_cast<const uint4*>(x + off + i * 256);
      dyq[i] = *reinterpret_cast<const uint4*>(dy + off + i * 256);
      if (HAS_DADD) daq[i] = *reinterpret_cast<const uint4*>(dadd + off + i * 256);
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[VPL][8], g[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float dyv[8];
      unpack8_fast(xq[i], xh[i]);
      unpack8_fast(dyq[i], dyv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xh[i][j] - mean) * rstd;
        g[i][j] = dyv[j] * wv[i][j];
        s1 += g[i][j];
        s2 = fmaf(g[i][j], xh[i][j], s2);
        dwv[i][j] = fmaf(dyv[j], xh[i][j], dwv[i][j]);
        dbv[i][j] += dyv[j];
      }
    }
    s1 = warp_sum(s1) * inv_w;
    s2 = warp_sum(s2) * inv_w;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * fmaf(-xh[i][j], s2, g[i][j] - s1);
      if (HAS_DADD) {
        float a[8];
        unpack8_fast(daq[i], a);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += a[j];
      }
      *reinterpret_cast<uint4*>(dx + off + i * 256) = pack8(o);
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = pass == 0 ? dwv[i][j] : dbv[i][j];
      __syncthreads();
      const int c = threadIdx.x;
      float acc = 0.f;
#pragma unroll
      for (int ww = 0; ww < LN_WARPS; ++ww) acc += red[ww][c];
      atomicAdd((pass == 0 ? dw : db) + i * 256 + c, acc);
    }
  }
}

template <typename P, typename F>
static int ln_dispatch(int W, F&& launch) {
  const int vpl = static_cast<int>(ceil_div(W, 256));
  switch (vpl) {
    case 1: launch(std::integral_constant<int, 1>{}); break;
    case 2: launch(std::integral_constant<int, 2>{}); break;
    case 3: launch(std::integral_constant<int, 3>{}); break;
    case 4: launch(std::integral_constant<int, 4>{}); break;
    case 5: launch(std::integral_constant<int, 5>{}); break;
    case 6: launch(std::integral_constant<int, 6>{}); break;
    case 8: case 7: launch(std::integral_constant<int, 8>{}); break;
    default:
      set_last_error("layernorm: width %d not supported (max 2048)", W);
      return B200MM_ERR_SHAPE;
  }
  return B200MM_OK;
}

}  // namespace b200mm

using namespace b200mm;

extern "C" int b200mm_layernorm_fwd(const void* x, const void* add0, const void* add1, int64_t add_period, const void* w,
                                    const void* b, void* y, void* s_out, float* mean, float* rstd, int64_t rows, int32_t W,
                                    float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "layernorm_fwd: rows=%lld W=%d (W %% 8 != 0)", (long long)rows, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(x && w && b && y, B200MM_ERR_SHAPE, "layernorm_fwd: null pointer");
  B200MM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w) |
                   reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(add0) | reinterpret_cast<uintptr_t>(add1) |
                   reinterpret_cast<uintptr_t>(s_out)) & 15) == 0,
                 B200MM_ERR_ALIGN, "layernorm_fwd: pointers must be 16B aligned");
  B200MM_REQUIRE((add0 == nullptr && add1 == nullptr) || add_period > 0, B200MM_ERR_SHAPE, "layernorm_fwd: add_period must be > 0");
  if (!add0 && !add1 && !s_out && mean && rstd && W % 256 == 0 && W / 256 >= 1 && W / 256 <= 5) {
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div(rows, LN_WARPS), static_cast<int64_t>(sm_count()) * 4));
    auto xx = reinterpret_cast<const __nv_bfloat16*>(x);
    auto ww = reinterpret_cast<const __nv_bfloat16*>(w);
    auto bb = reinterpret_cast<const __nv_bfloat16*>(b);
    auto yy = reinterpret_cast<__nv_bfloat16*>(y);
    switch (W / 256) {
      case 1: ln_fwd_plain_kernel<1><<<grid, LN_WARPS * 32, 0, stream>>>(xx, ww, bb, yy, mean, rstd, rows, eps); break;
      case 2: ln_fwd_plain_kernel<2><<<grid, LN_WARPS * 32, 0, stream>>>(xx, ww, bb, yy, mean, rstd, rows, eps); break;
      case 3: ln_fwd_plain_kernel<3><<<grid, LN_WARPS * 32, 0, stream>>>(xx, ww, bb, yy, mean, rstd, rows, eps); break;
      case 4: ln_fwd_plain_kernel<4><<<grid, LN_WARPS * 32, 0, stream>>>(xx, ww, bb, yy, mean, rstd, rows, eps); break;
      default: ln_fwd_plain_kernel<5><<<grid, LN_WARPS * 32, 0, stream>>>(xx, ww, bb, yy, mean, rstd, rows, eps); break;
    }
    return check_launch("ln_fwd_plain_kernel");
  }
  LnFwdParams p{reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(add0),
                reinterpret_cast<const __nv_bfloat16*>(add1), add_period, nullptr, nullptr, nullptr,
                reinterpret_cast<const __nv_bfloat16*>(w),
                reinterpret_cast<const __nv_bfloat16*>(b), reinterpret_cast<__nv_bfloat16*>(y),
                reinterpret_cast<__nv_bfloat16*>(s_out), mean, rstd, rows, W, eps};
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div(rows, LN_WARPS), static_cast<int64_t>(sm_count()) * 8));
  int rc = ln_dispatch<LnFwdParams>(W, [&](auto vpl) { ln_fwd_kernel<decltype(vpl)::value><<<grid, LN_WARPS * 32, 0, stream>>>(p); });
  if (rc) return rc;
  return check_launch("ln_fwd_kernel");
}

extern "C" int b200mm_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* w,
                                    const void* dadd, void* dx, float* dw, float* db, int64_t rows, int32_t W, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "layernorm_bwd: rows=%lld W=%d", (long long)rows, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(dy && x && mean && rstd && w && dx && dw && db, B200MM_ERR_SHAPE, "layernorm_bwd: null pointer");
  B200MM_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) |
                   reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dadd)) & 15) == 0,
                 B200MM_ERR_ALIGN, "layernorm_bwd: pointers must be 16B aligned");
  if (W % 256 == 0 && W / 256 <= 5) {
    const int grid = static_cast<int>(std::min<int64_t>(ceil_div(rows, LN_WARPS), static_cast<int64_t>(sm_count()) * 2));
    auto a0 = reinterpret_cast<const __nv_bfloat16*>(dy);
    auto a1 = reinterpret_cast<const __nv_bfloat16*>(x);
    auto a2 = reinterpret_cast<const __nv_bfloat16*>(w);
    auto a3 = reinterpret_cast<const __nv_bfloat16*>(dadd);
    auto a4 = reinterpret_cast<__nv_bfloat16*>(dx);
#define B200MM_LNB(V)                                                                                                          \
  if (dadd) ln_bwd_plain_kernel<V, true><<<grid, LN_WARPS * 32, 0, stream>>>(a0, a1, mean, rstd, a2, a3, a4, dw, db, rows);   \
  else ln_bwd_plain_kernel<V, false><<<grid, LN_WARPS * 32, 0, stream>>>(a0, a1, mean, rstd, a2, a3, a4, dw, db, rows);
    switch (W / 256) {
      case 1: B200MM_LNB(1) break;
      case 2: B200MM_LNB(2) break;
      case 3: B200MM_LNB(3) break;
      case 4: B200MM_LNB(4) break;
      default: B200MM_LNB(5) break;
    }
#undef B200MM_LNB
    return check_launch("ln_bwd_plain_kernel");
  }
  LnBwdParams p{reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), mean, rstd,
                reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<const __nv_bfloat16*>(dadd),
                reinterpret_cast<__nv_bfloat16*>(dx), dw, db, rows, W};
  // few, fat CTAs: each ends with W atomics per reduction, so keep the CTA count near 2 per SM
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div(rows, LN_WARPS), static_cast<int64_t>(sm_count()) * 2));
  int rc = ln_dispatch<LnBwdParams>(W, [&](auto vpl) { ln_bwd_kernel<decltype(vpl)::value><<<grid, LN_WARPS * 32, 0, stream>>>(p); });
  if (rc) return rc;
  return check_launch("ln_bwd_kernel");
}

// BertEmbeddings.forward (antmmf/modules/vision/backbone/clip/modeling_bert.py:86-103):
//   y = LN(word[ids] + pos[row % L] + type[type_ids]); the bf16 sum is written to s_out for backward.
extern "C" int b200mm_embed_layernorm_fwd(const void* word, const int64_t* ids, const void* pos, int64_t L, const void* type,
                                          const int64_t* type_ids, const void* w, const void* b, void* y, void* s_out, float* mean,
                                          float* rstd, int64_t rows, int32_t W, float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0 && L > 0, B200MM_ERR_SHAPE, "embed_layernorm_fwd: rows=%lld W=%d L=%lld", (long long)rows, W, (long long)L);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(word && ids && pos && type && type_ids && w && b && y, B200MM_ERR_SHAPE, "embed_layernorm_fwd: null pointer");
  B200MM_REQUIRE(((reinterpret_cast<uintptr_t>(word) | reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(type) |
                   reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y) |
                   reinterpret_cast<uintptr_t>(s_out)) & 15) == 0,
                 B200MM_ERR_ALIGN, "embed_layernorm_fwd: pointers must be 16B aligned");
  LnFwdParams p{reinterpret_cast<const __nv_bfloat16*>(word), reinterpret_cast<const __nv_bfloat16*>(pos), nullptr, L, ids,
                reinterpret_cast<const __nv_bfloat16*>(type), type_ids, reinterpret_cast<const __nv_bfloat16*>(w),
                reinterpret_cast<const __nv_bfloat16*>(b), reinterpret_cast<__nv_bfloat16*>(y),
                reinterpret_cast<__nv_bfloat16*>(s_out), mean, rstd, rows, W, eps};
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div(rows, LN_WARPS), static_cast<int64_t>(sm_count()) * 8));
  int rc = ln_dispatch<LnFwdParams>(W, [&](auto vpl) { ln_fwd_kernel<decltype(vpl)::value><<<grid, LN_WARPS * 32, 0, stream>>>(p); });
  if (rc) return rc;
  return check_launch("ln_fwd_kernel(embed)");
}
