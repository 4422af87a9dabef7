// b200mm — sub-LayerNorm kernels of the M²-Encoder (BEiT-3 multiway) blocks: LayerNorm fused with the activation that
// feeds it, for rows of ANY width that is a multiple of 8 (the FFN sub-LN normalises the 4·W-wide hidden: 3072 / 4096).
//
// Reference arithmetic (prj/M2_Encoder):
//   FeedForwardNetwork.forward  vlmo/torchscale/component/feedforward_network.py:117-128   fc2(ffn_layernorm(gelu(fc1 x)))
//   MultiheadAttention.forward  vlmo/torchscale/component/multihead_attention.py:148-149   inner_attn_ln on merged heads (act = none)
//
// forward : y = LN(act(u)) * w + b;  mean / rstd of act(u) are kept (fp32) for the backward
// backward: g = act(u) recomputed from u (never stored), xhat = (g - mean) * rstd,
//           dg = rstd * (dy*w - mean(dy*w) - xhat * mean(dy*w*xhat)),  du = dg * act'(u),  dw += dy*xhat, db += dy
//
// HBM-bound. One CTA per row (grid-stride over rows), 16-byte accesses, the row lives in registers between the passes.
// Algorithmic bytes per row (bf16): fwd 2W read + 2W write; bwd 4W read + 2W write. The unfused reference chain
// (gelu, LN, and in backward LN', gelu') moves 3x that.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace b200mm {

__device__ __forceinline__ void sl_unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 sl_pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

template <int ACT>
__device__ __forceinline__ float sl_act(float x) {
  return ACT == B200MM_ACT_QUICKGELU ? act_quickgelu(x) : (ACT == B200MM_ACT_GELU_ERF ? act_gelu_erf(x) : x);
}
template <int ACT>
__device__ __forceinline__ float sl_dact(float x) {
  return ACT == B200MM_ACT_QUICKGELU ? dact_quickgelu(x) : (ACT == B200MM_ACT_GELU_ERF ? dact_gelu_erf(x) : 1.f);
}

// act(x) and act'(x) together: QuickGELU shares the sigmoid, erf-GELU the Gaussian cdf / pdf pair
template <int ACT>
__device__ __forceinline__ void sl_act_dact(float x, float& a, float& d) {
  if (ACT == B200MM_ACT_QUICKGELU) {
    const float s = sigmoid_1702(x);
    a = x * s;
    d = s * fmaf(1.702f * x, 1.f - s, 1.f);
  } else if (ACT == B200MM_ACT_GELU_ERF) {
    float cdf, expo;
    gauss_cdf_pdf(x, cdf, expo);
    a = x * cdf;
    d = fmaf(x * 0.3989422804014327f, expo, cdf);
  } else {
    a = x;
    d = 1.f;
  }
}

// CTA-wide sum through one smem slot per warp; `red` must not be reused before the NEXT __syncthreads of the caller
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) t += red[i];
  return t;
}

// thread t, slot i covers columns (i*THREADS + t)*8 .. +7
template <int VPT, int THREADS, int ACT, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB) act_ln_fwd_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ w,
                                                             const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y,
                                                             float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int32_t W,
                                                             float eps) {
  __shared__ float red_a[THREADS / 32], red_b[THREADS / 32];
  // affine parameters as fp32 in shared memory (2 x 4 B x row width <= 32 KB): the output pass reads them with LDS.128 instead of
  // unpacking bf16 registers for every row (2 of ~33 issue slots per element in the first version, cuobjdump -sass)
  // Only with an activation: there the kernel is issue-bound and the saved slots pay (0.59 -> 0.42 ms at 100 864 x 4096); without one it
  // is HBM-bound and the 32 KB of smem would cost occupancy (0.32 -> 0.35 ms), so the bf16 registers stay (profiles/r01g_subln_sweep.log).
  constexpr bool kSmemWB = ACT != B200MM_ACT_NONE && VPT * THREADS * 8 * 8 <= 32768;
  __shared__ __align__(16) float w_s[kSmemWB ? VPT * THREADS * 8 : 4], b_s[kSmemWB ? VPT * THREADS * 8 : 4];
  const float inv_w = 1.f / static_cast<float>(W);
  uint4 wq[kSmemWB ? 1 : VPT], bq[kSmemWB ? 1 : VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int col = (i * THREADS + threadIdx.x) * 8;
    uint4 wv4 = make_uint4(0, 0, 0, 0), bv4 = make_uint4(0, 0, 0, 0);
    if (col < W) {
      wv4 = *reinterpret_cast<const uint4*>(w + col);
      bv4 = *reinterpret_cast<const uint4*>(b + col);
    }
    if (kSmemWB) {  // each thread stores (and later reads) only its own columns: no barrier needed
      float wf[8], bf[8];
      sl_unpack8(wv4, wf);
      sl_unpack8(bv4, bf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        w_s[col + j] = wf[j];
        b_s[col + j] = bf[j];
      }
    } else {
      wq[i] = wv4;
      bq[i] = bv4;
    }
  }
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const __nv_bfloat16* ur = u + row * W;
    uint4 q[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int col = (i * THREADS + threadIdx.x) * 8;
      q[i] = make_uint4(0, 0, 0, 0);
      if (col < W) q[i] = *reinterpret_cast<const uint4*>(ur + col);
    }
    float v[VPT][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      sl_unpack8(q[i], v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = sl_act<ACT>(v[i][j]);  // act(0) = 0 for every supported activation: padding slots add nothing
        sum += v[i][j];
      }
    }
    const float mean = block_sum<THREADS>(sum, red_a) * inv_w;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      if ((i * THREADS + threadIdx.x) * 8 < W) {  // padding slots hold act(0) = 0 and must not add mean^2
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[i][j] -= mean;
          sq = fmaf(v[i][j], v[i][j], sq);
        }
      }
    }
    const float rstd = rsqrtf(block_sum<THREADS>(sq, red_b) * inv_w + eps);
    if (threadIdx.x == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    __nv_bfloat16* yr = y + row * W;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int col = (i * THREADS + threadIdx.x) * 8;
      if (col < W) {
        float wv[8], bv[8], o[8];
        if (kSmemWB) {
          *reinterpret_cast<float4*>(&wv[0]) = *reinterpret_cast<const float4*>(&w_s[col]);
          *reinterpret_cast<float4*>(&wv[4]) = *reinterpret_cast<const float4*>(&w_s[col + 4]);
          *reinterpret_cast<float4*>(&bv[0]) = *reinterpret_cast<const float4*>(&b_s[col]);
          *reinterpret_cast<float4*>(&bv[4]) = *reinterpret_cast<const float4*>(&b_s[col + 4]);
        } else {
          sl_unpack8(wq[i], wv);
          sl_unpack8(bq[i], bv);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(v[i][j] * rstd, wv[j], bv[j]);
        *reinterpret_cast<uint4*>(yr + col) = sl_pack8(o);
      }
    }
  }
}

template <int VPT, int THREADS, int ACT, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB) act_ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ u,
                                                             const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                             const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ du,
                                                             float* __restrict__ dw, float* __restrict__ db, int64_t rows, int32_t W) {
  __shared__ float red1[2][THREADS / 32], red2[2][THREADS / 32];
  __shared__ __align__(16) float w_s[VPT * THREADS * 8];  // fp32 LN weights (<= 32 KB), thread-private columns: no barrier
  const float inv_w = 1.f / static_cast<float>(W);
  float dwv[VPT][8], dbv[VPT][8];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int col = (i * THREADS + threadIdx.x) * 8;
    uint4 wv4 = make_uint4(0, 0, 0, 0);
    if (col < W) wv4 = *reinterpret_cast<const uint4*>(w + col);
    float wf[8];
    sl_unpack8(wv4, wf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      w_s[col + j] = wf[j];
      dwv[i][j] = 0.f;
      dbv[i][j] = 0.f;
    }
  }
  int par = 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x, par ^= 1) {
    const int64_t off = row * W;
    uint4 uq[VPT], dq[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int col = (i * THREADS + threadIdx.x) * 8;
      uq[i] = dq[i] = make_uint4(0, 0, 0, 0);
      if (col < W) {
        uq[i] = *reinterpret_cast<const uint4*>(u + off + col);
        dq[i] = *reinterpret_cast<const uint4*>(dy + off + col);
      }
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[VPT][8], gy[VPT][8], da[VPT][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float uv[8], dyv[8], wv[8];
      sl_unpack8(uq[i], uv);
      sl_unpack8(dq[i], dyv);
      const int col = (i * THREADS + threadIdx.x) * 8;
      *reinterpret_cast<float4*>(&wv[0]) = *reinterpret_cast<const float4*>(&w_s[col]);
      *reinterpret_cast<float4*>(&wv[4]) = *reinterpret_cast<const float4*>(&w_s[col + 4]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float g;
        sl_act_dact<ACT>(uv[j], g, da[i][j]);  // act(u) and act'(u) from ONE evaluation of the shared sub-expressions
        xh[i][j] = (g - mean) * rstd;
        gy[i][j] = dyv[j] * wv[j];  // dy and w are zero in padding slots, so those add nothing below
        s1 += gy[i][j];
        s2 = fmaf(gy[i][j], xh[i][j], s2);
        dwv[i][j] = fmaf(dyv[j], xh[i][j], dwv[i][j]);
        dbv[i][j] += dyv[j];
      }
    }
    // two independent reductions behind ONE barrier: each warp posts both partials, then everyone sums them
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) {
      red1[par][threadIdx.x >> 5] = s1;
      red2[par][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    s1 = 0.f;
    s2 = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) {
      s1 += red1[par][i];
      s2 += red2[par][i];
    }
    s1 *= inv_w;
    s2 *= inv_w;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int col = (i * THREADS + threadIdx.x) * 8;
      if (col < W) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * fmaf(-xh[i][j], s2, gy[i][j] - s1) * da[i][j];
        *reinterpret_cast<uint4*>(du + off + col) = sl_pack8(o);
      }
    }
  }
  // every thread owns its columns: no cross-thread reduction inside the CTA, one atomic per column and CTA
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int col = (i * THREADS + threadIdx.x) * 8;
    if (col < W) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(dw + col + j, dwv[i][j]);
        atomicAdd(db + col + j, dbv[i][j]);
      }
    }
  }
}

// y[r, :] = keep[r] ? x[r, :] : 0   (Encoder.forward zeroes padded token rows: architecture/encoder.py:440)
__global__ void __launch_bounds__(256) mask_rows_kernel(const uint4* __restrict__ x, const uint8_t* __restrict__ drop, uint4* __restrict__ y,
                                                        int64_t rows, int32_t w8) {
  const int64_t n = rows * w8;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    const int64_t r = i / w8;
    y[i] = drop[r] ? make_uint4(0, 0, 0, 0) : x[i];
  }
}

}  // namespace b200mm

using namespace b200mm;

#define SL_STREAM(s) reinterpret_cast<cudaStream_t>(s)

namespace {

// launch shape: THREADS x VPT vectors cover the row. Defaults measured on B200 (tools/subln_bench.py); B200MM_SUBLN_FWD / _BWD =
// "<threads>,<ctas_per_sm>" override them for bring-up sweeps.
struct SubLnCfg { int threads, per_sm; };
static SubLnCfg subln_cfg(const char* env, int def_threads, int def_per_sm) {
  SubLnCfg c{def_threads, def_per_sm};
  if (const char* e = getenv(env)) {
    int t = 0, p = 0;
    if (sscanf(e, "%d,%d", &t, &p) == 2 && (t == 128 || t == 256 || t == 512) && p > 0 && p <= 32) c = SubLnCfg{t, p};
  }
  return c;
}

template <int ACT>
int sub_ln_fwd(const __nv_bfloat16* u, const __nv_bfloat16* w, const __nv_bfloat16* b, __nv_bfloat16* y, float* mean, float* rstd, int64_t rows,
               int32_t W, float eps, cudaStream_t st) {
  // CTAs per SM: 6 x 32.8 KB of smem is what fits with the fp32 affine copy (8 would run as two uneven waves: 0.50 vs 0.42 ms)
  SubLnCfg c = subln_cfg("B200MM_SUBLN_FWD", W <= 4096 ? 128 : 256, ACT != B200MM_ACT_NONE && W > 2048 ? 6 : 8);
  if (ceil_div(W, c.threads * 8) > (c.threads == 512 ? 2 : 4)) c.threads = W <= 4096 ? 128 : 256;  // override does not fit this width
  B200MM_REQUIRE(W <= 8192, B200MM_ERR_SHAPE, "act_layernorm_fwd: width %d not supported (max 8192)", W);
  const int grid = static_cast<int>(std::min<int64_t>(rows, static_cast<int64_t>(sm_count()) * c.per_sm));
  const int v = static_cast<int>(ceil_div(W, c.threads * 8));
#define SL_F(V, T, MB) act_ln_fwd_kernel<V, T, ACT, MB><<<grid, T, 0, st>>>(u, w, b, y, mean, rstd, rows, W, eps)
  if (c.threads == 128) { if (v == 1) SL_F(1, 128, 1); else if (v == 2) SL_F(2, 128, 1); else if (v == 3) SL_F(3, 128, 1); else SL_F(4, 128, 1); }
  else if (c.threads == 256) { if (v == 1) SL_F(1, 256, 1); else if (v == 2) SL_F(2, 256, 1); else if (v == 3) SL_F(3, 256, 1); else SL_F(4, 256, 1); }
  else { if (v == 1) SL_F(1, 512, 2); else SL_F(2, 512, 1); }
#undef SL_F
  return check_launch("act_ln_fwd_kernel");
}

template <int ACT>
int sub_ln_bwd(const __nv_bfloat16* dy, const __nv_bfloat16* u, const float* mean, const float* rstd, const __nv_bfloat16* w, __nv_bfloat16* du,
               float* dw, float* db, int64_t rows, int32_t W, cudaStream_t st) {
  // register budget: 2 accumulators + xhat + dy*w + act' per column held by the thread -> at most 16 columns per thread up to 4096
  // (32 up to 8192); each CTA ends with 2*W atomics, so the CTA count stays at a few per SM
  // measured at 100 864 x 4096 (profiles/r01e_subln_sweep.log): with an activation the kernel is MUFU/issue-bound and wants the most
  // warps (512 threads, 2 CTAs/SM: 0.77 ms vs 1.30 ms at 256 threads); without one 256 threads x 2 CTAs/SM is best (68 % of HBM peak)
  const int def_threads = W <= 2048 ? 128 : (ACT != B200MM_ACT_NONE ? 512 : 256);
  SubLnCfg c = subln_cfg("B200MM_SUBLN_BWD", def_threads, W <= 2048 ? 4 : 2);
  if (ceil_div(W, c.threads * 8) > (c.threads == 512 ? 2 : (c.threads == 256 ? 4 : 2))) c.threads = W <= 2048 ? 128 : 256;
  B200MM_REQUIRE(W <= 8192, B200MM_ERR_SHAPE, "act_layernorm_bwd: width %d not supported (max 8192)", W);
  const int grid = static_cast<int>(std::min<int64_t>(rows, static_cast<int64_t>(sm_count()) * c.per_sm));
  const int v = static_cast<int>(ceil_div(W, c.threads * 8));
#define SL_B(V, T, MB) act_ln_bwd_kernel<V, T, ACT, MB><<<grid, T, 0, st>>>(dy, u, mean, rstd, w, du, dw, db, rows, W)
  if (c.threads == 128) { if (v == 1) SL_B(1, 128, 1); else SL_B(2, 128, 1); }
  else if (c.threads == 256) { if (v == 1) SL_B(1, 256, 1); else if (v == 2) SL_B(2, 256, 1); else SL_B(4, 256, 1); }
  else { if (v == 1) SL_B(1, 512, 2); else SL_B(2, 512, 1); }
#undef SL_B
  return check_launch("act_ln_bwd_kernel");
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int b200mm_act_layernorm_fwd(const void* u, int32_t act, const void* w, const void* b, void* y, float* mean, float* rstd,
                                        int64_t rows, int32_t W, float eps, void* stream) {
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "act_layernorm_fwd: rows=%lld W=%d (W %% 8 != 0)", (long long)rows, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(u && w && b && y && mean && rstd, B200MM_ERR_SHAPE, "act_layernorm_fwd: null pointer");
  B200MM_REQUIRE(al16(u) && al16(w) && al16(b) && al16(y), B200MM_ERR_ALIGN, "act_layernorm_fwd: pointers must be 16B aligned");
  auto uu = reinterpret_cast<const __nv_bfloat16*>(u);
  auto ww = reinterpret_cast<const __nv_bfloat16*>(w);
  auto bb = reinterpret_cast<const __nv_bfloat16*>(b);
  auto yy = reinterpret_cast<__nv_bfloat16*>(y);
  switch (act) {
    case B200MM_ACT_NONE: return sub_ln_fwd<B200MM_ACT_NONE>(uu, ww, bb, yy, mean, rstd, rows, W, eps, SL_STREAM(stream));
    case B200MM_ACT_QUICKGELU: return sub_ln_fwd<B200MM_ACT_QUICKGELU>(uu, ww, bb, yy, mean, rstd, rows, W, eps, SL_STREAM(stream));
    case B200MM_ACT_GELU_ERF: return sub_ln_fwd<B200MM_ACT_GELU_ERF>(uu, ww, bb, yy, mean, rstd, rows, W, eps, SL_STREAM(stream));
    default: B200MM_REQUIRE(false, B200MM_ERR_SHAPE, "act_layernorm_fwd: unknown activation %d", act);
  }
}

extern "C" int b200mm_act_layernorm_bwd(const void* dy, const void* u, int32_t act, const float* mean, const float* rstd, const void* w, void* du,
                                        float* dw, float* db, int64_t rows, int32_t W, void* stream) {
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "act_layernorm_bwd: rows=%lld W=%d (W %% 8 != 0)", (long long)rows, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(dy && u && mean && rstd && w && du && dw && db, B200MM_ERR_SHAPE, "act_layernorm_bwd: null pointer");
  B200MM_REQUIRE(al16(dy) && al16(u) && al16(w) && al16(du), B200MM_ERR_ALIGN, "act_layernorm_bwd: pointers must be 16B aligned");
  auto dd = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto uu = reinterpret_cast<const __nv_bfloat16*>(u);
  auto ww = reinterpret_cast<const __nv_bfloat16*>(w);
  auto oo = reinterpret_cast<__nv_bfloat16*>(du);
  switch (act) {
    case B200MM_ACT_NONE: return sub_ln_bwd<B200MM_ACT_NONE>(dd, uu, mean, rstd, ww, oo, dw, db, rows, W, SL_STREAM(stream));
    case B200MM_ACT_QUICKGELU: return sub_ln_bwd<B200MM_ACT_QUICKGELU>(dd, uu, mean, rstd, ww, oo, dw, db, rows, W, SL_STREAM(stream));
    case B200MM_ACT_GELU_ERF: return sub_ln_bwd<B200MM_ACT_GELU_ERF>(dd, uu, mean, rstd, ww, oo, dw, db, rows, W, SL_STREAM(stream));
    default: B200MM_REQUIRE(false, B200MM_ERR_SHAPE, "act_layernorm_bwd: unknown activation %d", act);
  }
}

extern "C" int b200mm_mask_rows(const void* x, const uint8_t* drop, void* y, int64_t rows, int32_t W, void* stream) {
  B200MM_REQUIRE(rows >= 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "mask_rows: rows=%lld W=%d (W %% 8 != 0)", (long long)rows, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(x && drop && y, B200MM_ERR_SHAPE, "mask_rows: null pointer");
  B200MM_REQUIRE(al16(x) && al16(y), B200MM_ERR_ALIGN, "mask_rows: pointers must be 16B aligned");
  const int64_t n = rows * (W / 8);
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div(n, 256), static_cast<int64_t>(sm_count()) * 16));
  mask_rows_kernel<<<grid, 256, 0, SL_STREAM(stream)>>>(reinterpret_cast<const uint4*>(x), drop, reinterpret_cast<uint4*>(y), rows, W / 8);
  return check_launch("mask_rows_kernel");
}
