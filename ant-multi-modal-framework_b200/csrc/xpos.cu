// b200mm — XPOS rotary position embedding applied in place to the q and k sections of the fused QKV projection output
// (the optional "RoPE" of the hot path: prj/M2_Encoder/vlmo/torchscale/component/xpos_relative_position.py:15-62, used by
// MultiheadAttention.forward, component/multihead_attention.py:112-118; off in every shipped config, args.xpos_rel_pos = False).
//
//   out[2i]   = cs[l,i] * x[2i]   - sn[l,i] * x[2i+1]          cs = cos(l * inv_freq_i) * scale[l,i]   (q: scale, k: 1/scale)
//   out[2i+1] = cs[l,i] * x[2i+1] + sn[l,i] * x[2i]            sn = sin(l * inv_freq_i) * scale[l,i]
// per head, "rotate every two" pairing (:17-21). The tables [L, hd/2] (fp32) are built on the host with the reference's own torch
// expressions, so their values are bit-identical; the backward pass is the same map with sn negated (transpose of the 2x2 blocks).
// HBM-bound: reads and writes 2/3 of the [T, 3W] buffer once; the tables stay in L2.
#include "common.cuh"

#include <algorithm>

namespace b200mm {

__global__ void __launch_bounds__(256) xpos_apply_kernel(__nv_bfloat16* __restrict__ qkv, int64_t ld, int32_t q_off, int32_t k_off,
                                                         const float* __restrict__ q_cs, const float* __restrict__ q_sn,
                                                         const float* __restrict__ k_cs, const float* __restrict__ k_sn, int64_t T, int32_t L,
                                                         int32_t W, int32_t hd, float sn_sign) {
  const int w8 = W / 8;                    // 16-byte vectors per section and row
  const int64_t n = T * 2 * w8;            // q and k sections
  const int half = hd / 2;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    const int64_t row = i / (2 * w8);
    const int rem = static_cast<int>(i - row * 2 * w8);
    const int sec = rem / w8, v = rem - sec * w8;
    const int col = v * 8;                 // column inside the section; a vector never straddles heads (hd % 8 == 0)
    const int pair0 = (col % hd) / 2;      // first of the 4 (even, odd) pairs of this vector
    const int l = static_cast<int>(row % L);
    const float* cs = (sec == 0 ? q_cs : k_cs) + static_cast<int64_t>(l) * half + pair0;
    const float* sn = (sec == 0 ? q_sn : k_sn) + static_cast<int64_t>(l) * half + pair0;
    __nv_bfloat16* p = qkv + row * ld + (sec == 0 ? q_off : k_off) + col;
    uint4 u = *reinterpret_cast<uint4*>(p);
    uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16x2(w[j]);
      const float c = cs[j], s = sn[j] * sn_sign;
      w[j] = pack_bf16x2(fmaf(c, x.x, -s * x.y), fmaf(c, x.y, s * x.x));
    }
    *reinterpret_cast<uint4*>(p) = u;
  }
}

}  // namespace b200mm

using namespace b200mm;

extern "C" int b200mm_xpos_apply(void* qkv, int64_t ld, int32_t q_off, int32_t k_off, const float* q_cos, const float* q_sin,
                                 const float* k_cos, const float* k_sin, int64_t T, int32_t L, int32_t H, int32_t hd, int32_t backward,
                                 void* stream) {
  B200MM_REQUIRE(T >= 0 && L > 0 && H > 0 && hd > 0 && hd % 8 == 0 && T % L == 0, B200MM_ERR_SHAPE,
                 "xpos_apply: T=%lld L=%d H=%d hd=%d (hd %% 8 != 0 or T %% L != 0)", (long long)T, L, H, hd);
  if (T == 0) return B200MM_OK;
  B200MM_REQUIRE(qkv && q_cos && q_sin && k_cos && k_sin, B200MM_ERR_SHAPE, "xpos_apply: null pointer");
  B200MM_REQUIRE(ld % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0, B200MM_ERR_ALIGN,
                 "xpos_apply: qkv must be 16B aligned with ld, q_off, k_off multiples of 8");
  const int W = H * hd;
  const int64_t n = T * 2 * (W / 8);
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), static_cast<int64_t>(sm_count()) * 16)));
  xpos_apply_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<__nv_bfloat16*>(qkv), ld, q_off, k_off, q_cos,
                                                                            q_sin, k_cos, k_sin, T, L, W, hd, backward ? -1.f : 1.f);
  return check_launch("xpos_apply_kernel");
}
