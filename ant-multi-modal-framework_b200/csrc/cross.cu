// b200mm — helpers of the stage-2 (cross-modal) retrieval path of prj/base_vtp:
//   * pair gather: builds the [text ; visual] token rows of MANY (text, video) pairs in one pass, replacing the reference's
//     unsqueeze/repeat/view/cat copies per block of 5 texts (univl_video_ret.py:52-78) and per mined row (:101-131);
//   * ReLU forward / backward of the similarity head (univl_video_ret.py:24-28);
//   * MIL-NCE on an explicit square score matrix with optional row weights, forward + gradient
//     (get_mil_nce_loss, univl_video_ret.py:146-197 as called from forward_stage2 :403-433).
// All HBM/L2-bound byte and small-matrix work; the encoder layers in between run on the tcgen05 GEMM / attention kernels.
#include "common.cuh"

#include <algorithm>

namespace b200mm {

// out[r, :] = src[ids[r], :]  (bit-exact row copy; ids outside [0, n_src) give a zero row). W % 8 == 0.
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ src, const int64_t* __restrict__ ids, uint4* __restrict__ out,
                                                          int64_t rows, int64_t n_src, int32_t w8) {
  // one warp per output row, lanes stride over the 16-byte vectors of the row
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * 256) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const int64_t id = ids[r];
    const bool ok = id >= 0 && id < n_src;
    const uint4* s = src + (ok ? id : 0) * w8;
    uint4* o = out + r * w8;
    for (int c = lane; c < w8; c += 32) o[c] = ok ? s[c] : make_uint4(0, 0, 0, 0);
  }
}

__device__ __forceinline__ uint32_t relu2(uint32_t v) {
  // two bf16 lanes: clear a lane when its sign bit is set (−0 → +0 as well, like max(x, 0))
  const uint32_t lo = (v & 0x00008000u) ? 0u : (v & 0x0000ffffu);
  const uint32_t hi = (v & 0x80000000u) ? 0u : (v & 0xffff0000u);
  return lo | hi;
}
__device__ __forceinline__ uint32_t relu_mask2(uint32_t x, uint32_t dy) {
  // pass dy where x > 0 (strictly: torch's threshold_backward), per bf16 lane
  const uint32_t xl = x & 0xffffu, xh = x >> 16;
  const bool pl = !(xl & 0x8000u) && (xl & 0x7fffu) != 0, ph = !(xh & 0x8000u) && (xh & 0x7fffu) != 0;
  return (pl ? (dy & 0x0000ffffu) : 0u) | (ph ? (dy & 0xffff0000u) : 0u);
}

__global__ void __launch_bounds__(256) relu_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t n8) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n8; i += gridDim.x * 256ll) {
    uint4 v = x[i];
    v.x = relu2(v.x); v.y = relu2(v.y); v.z = relu2(v.z); v.w = relu2(v.w);
    y[i] = v;
  }
}
__global__ void __launch_bounds__(256) relu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, uint4* __restrict__ dx, int64_t n8) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n8; i += gridDim.x * 256ll) {
    const uint4 a = x[i], g = dy[i];
    uint4 o;
    o.x = relu_mask2(a.x, g.x); o.y = relu_mask2(a.y, g.y); o.z = relu_mask2(a.z, g.z); o.w = relu_mask2(a.w, g.w);
    dx[i] = o;
  }
}

// One warp per column/row index j:  Z_j = sum_i e^{S[i,j]} + sum_{k != j} e^{S[j,k]},  lse_j = log Z_j,
// loss_sum += w_j * (lse_j - S[j,j]).  S is fp32 [B, ld]; B is small (stage-2 batch), the matrix stays in L2.
__global__ void __launch_bounds__(256) mil_nce_matrix_fwd_kernel(const float* __restrict__ S, int64_t ld, const float* __restrict__ w,
                                                                 float* __restrict__ lse, float* __restrict__ loss_sum, int32_t B) {
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * 256 + threadIdx.x) >> 5;
  if (j >= B) return;
  float m = -INFINITY;
  for (int i = lane; i < B; i += 32) {
    m = fmaxf(m, S[static_cast<int64_t>(i) * ld + j]);
    if (i != j) m = fmaxf(m, S[static_cast<int64_t>(j) * ld + i]);
  }
  m = warp_max(m);
  float z = 0.f;
  for (int i = lane; i < B; i += 32) {
    z += expf(S[static_cast<int64_t>(i) * ld + j] - m);
    if (i != j) z += expf(S[static_cast<int64_t>(j) * ld + i] - m);
  }
  z = warp_sum(z);
  if (lane == 0) {
    const float l = m + logf(z);
    lse[j] = l;
    atomicAdd(loss_sum, (w ? w[j] : 1.f) * (l - S[static_cast<int64_t>(j) * ld + j]));
  }
}

// dS[a,b] = c * ( w_b e^{S_ab - lse_b} + [a != b] w_a e^{S_ab - lse_a} - [a == b] w_a ),  c = *gout / B
__global__ void __launch_bounds__(256) mil_nce_matrix_bwd_kernel(const float* __restrict__ S, int64_t ld, const float* __restrict__ w,
                                                                 const float* __restrict__ lse, const float* __restrict__ gout,
                                                                 float* __restrict__ dS, int32_t B) {
  const int64_t n = static_cast<int64_t>(B) * B;
  const float c = gout[0] / static_cast<float>(B);
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    const int a = static_cast<int>(i / B), b = static_cast<int>(i % B);
    const float s = S[static_cast<int64_t>(a) * ld + b];
    const float wa = w ? w[a] : 1.f, wb = w ? w[b] : 1.f;
    float g = wb * expf(s - lse[b]);
    g += a != b ? wa * expf(s - lse[a]) : -wa;
    dS[i] = c * g;
  }
}

}  // namespace b200mm

using namespace b200mm;

namespace {
inline bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int grid_1d(int64_t n_threads, int per_sm = 16) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_threads, 256), static_cast<int64_t>(sm_count()) * per_sm)));
}
}  // namespace

extern "C" int b200mm_gather_rows(const void* src, const int64_t* ids, void* out, int64_t rows, int64_t n_src, int32_t W, void* stream) {
  B200MM_REQUIRE(rows >= 0 && n_src > 0 && W > 0 && W % 8 == 0, B200MM_ERR_SHAPE, "gather_rows: rows=%lld n_src=%lld W=%d (W %% 8 != 0)",
                 (long long)rows, (long long)n_src, W);
  if (rows == 0) return B200MM_OK;
  B200MM_REQUIRE(src && ids && out, B200MM_ERR_SHAPE, "gather_rows: null pointer");
  B200MM_REQUIRE(a16(src) && a16(out), B200MM_ERR_ALIGN, "gather_rows: pointers must be 16B aligned");
  gather_rows_kernel<<<grid_1d(rows * 32), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(src), ids, reinterpret_cast<uint4*>(out), rows, n_src, W / 8);
  return check_launch("gather_rows_kernel");
}

extern "C" int b200mm_relu_fwd(const void* x, void* y, int64_t n, void* stream) {
  B200MM_REQUIRE(n >= 0 && n % 8 == 0, B200MM_ERR_SHAPE, "relu_fwd: n=%lld must be a multiple of 8", (long long)n);
  if (n == 0) return B200MM_OK;
  B200MM_REQUIRE(a16(x) && a16(y), B200MM_ERR_ALIGN, "relu_fwd: pointers must be 16B aligned");
  relu_fwd_kernel<<<grid_1d(n / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), n / 8);
  return check_launch("relu_fwd_kernel");
}

extern "C" int b200mm_relu_bwd(const void* dy, const void* x, void* dx, int64_t n, void* stream) {
  B200MM_REQUIRE(n >= 0 && n % 8 == 0, B200MM_ERR_SHAPE, "relu_bwd: n=%lld must be a multiple of 8", (long long)n);
  if (n == 0) return B200MM_OK;
  B200MM_REQUIRE(a16(dy) && a16(x) && a16(dx), B200MM_ERR_ALIGN, "relu_bwd: pointers must be 16B aligned");
  relu_bwd_kernel<<<grid_1d(n / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(dx), n / 8);
  return check_launch("relu_bwd_kernel");
}

extern "C" int b200mm_mil_nce_matrix_fwd(const float* S, int64_t ld, const float* w, float* lse, float* loss_sum, int32_t B, void* stream) {
  B200MM_REQUIRE(B > 0 && ld >= B, B200MM_ERR_SHAPE, "mil_nce_matrix_fwd: B=%d ld=%lld", B, (long long)ld);
  B200MM_REQUIRE(S && lse && loss_sum, B200MM_ERR_SHAPE, "mil_nce_matrix_fwd: null pointer");
  mil_nce_matrix_fwd_kernel<<<static_cast<int>(ceil_div(static_cast<int64_t>(B) * 32, 256)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      S, ld, w, lse, loss_sum, B);
  return check_launch("mil_nce_matrix_fwd_kernel");
}

extern "C" int b200mm_mil_nce_matrix_bwd(const float* S, int64_t ld, const float* w, const float* lse, const float* gout, float* dS, int32_t B,
                                         void* stream) {
  B200MM_REQUIRE(B > 0 && ld >= B, B200MM_ERR_SHAPE, "mil_nce_matrix_bwd: B=%d ld=%lld", B, (long long)ld);
  B200MM_REQUIRE(S && lse && gout && dS, B200MM_ERR_SHAPE, "mil_nce_matrix_bwd: null pointer");
  mil_nce_matrix_bwd_kernel<<<grid_1d(static_cast<int64_t>(B) * B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(S, ld, w, lse, gout, dS, B);
  return check_launch("mil_nce_matrix_bwd_kernel");
}
