"""Tensor-level wrappers of the C-ABI (one Python function per kernel entry point; no arithmetic happens here).

torch is used for allocation, stream handles and dtype/shape checks only. Every function requires CUDA tensors and
raises otherwise: the product path has no CPU fallback.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_QUICKGELU  # noqa: F401

BF16 = torch.bfloat16

# bookkeeping for bench.py: number of b200mm kernels launched, and (when enabled) CUDA-event timing of every GEMM launch
LAUNCHES = 0
GEMM_PROFILE = None  # set to a list to collect (flops, start_event, stop_event) per tcgen05 GEMM launch


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _req(t, name, dtype=BF16, dim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.B200mmError(f"b200mm: `{name}` must be a CUDA tensor (no CPU fallback exists)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.B200mmError(f"b200mm: `{name}` must be {dtype}, got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise _lib.B200mmError(f"b200mm: `{name}` must be {dim}-D, got shape {tuple(t.shape)}")
    return t


def _row_major_2d(t, name):
    _req(t, name, BF16, 2)
    if t.stride(1) != 1:
        raise _lib.B200mmError(f"b200mm: `{name}` must have unit inner stride, got strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


# ---------------------------------------------------------------------------------------------------------------------
# dropout seeds: every dropout site of a step draws one 64-bit seed = hash(base seed, running counter) on the HOST (no device RNG state,
# no sync); the site keeps it for backward / recompute. `manual_seed` re-bases the sequence (default: torch.initial_seed() at first use).
# ---------------------------------------------------------------------------------------------------------------------
_DROP_BASE = None
_DROP_COUNTER = 0


def manual_seed(seed):
    global _DROP_BASE, _DROP_COUNTER
    _DROP_BASE, _DROP_COUNTER = int(seed) & 0xFFFFFFFFFFFFFFFF, 0


def next_dropout_seed():
    """A fresh 64-bit seed (splitmix64 of base + counter). Ranks of a data-parallel job decorrelate through their base seed."""
    global _DROP_BASE, _DROP_COUNTER
    if _DROP_BASE is None:
        manual_seed(torch.initial_seed())
    _DROP_COUNTER += 1
    z = (_DROP_BASE + 0x9E3779B97F4A7C15 * _DROP_COUNTER) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def get_dropout_state():
    """(base seed, counter) — with set_dropout_state the way to replay a forward pass with the same masks (GradCache pass 2)."""
    if _DROP_BASE is None:
        manual_seed(torch.initial_seed())
    return _DROP_BASE, _DROP_COUNTER


def set_dropout_state(state):
    global _DROP_BASE, _DROP_COUNTER
    _DROP_BASE, _DROP_COUNTER = state


_SM = {}


def sm_count():
    d = torch.cuda.current_device()
    if d not in _SM:
        _SM[d] = torch.cuda.get_device_properties(d).multi_processor_count
    return _SM[d]


def pick_splits(M, N, K, max_splits=32):
    """Split-K factor for GEMMs with few output tiles (weight gradients): maximise wave efficiency over the SMs."""
    tiles = math.ceil(M / 128) * math.ceil(N / 256)
    kb = math.ceil(K / 64)
    sms = sm_count()
    if tiles >= sms or kb < 16:
        return 1
    best, best_eff = 1, 0.0
    for s in range(1, max_splits + 1):
        if kb // s < 8:
            break
        ctas = tiles * s
        eff = ctas / (math.ceil(ctas / sms) * sms)
        if eff > best_eff + 0.03:
            best, best_eff = s, eff
    return best


def gemm(a, b, *, a_mn=False, b_mn=False, bias=None, act=ACT_NONE, aux_out=False, dact_in=None, residual=None, alpha=1.0,
         out=None, out_f32=False, splits=None, drop=None):
    """D = epilogue(alpha * A·B^T) on the tcgen05 GEMM (b200mm_gemm_bf16).

    a: [M, K] (a_mn=False) or [K, M] (a_mn=True, reduction index slow);  b: [N, K] (b_mn=False) or [K, N] (b_mn=True).
    Returns D [M, N] (bf16, or f32 if out_f32) — and the pre-activation copy if aux_out.
    drop = (p, seed): inverted dropout on act(alpha*acc + bias) before the residual add, mask of `dropout(x, p, seed)`.
    """
    lib = _lib.load()
    lda = _row_major_2d(a, "a")
    ldb = _row_major_2d(b, "b")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise _lib.B200mmError(f"b200mm.gemm: reduction dims differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if out_f32 else BF16)
    else:
        _req(out, "out", torch.float32 if out_f32 else BF16, 2)
    aux = torch.empty((M, N), device=a.device, dtype=BF16) if aux_out else None
    if splits is None:
        splits = pick_splits(M, N, K)
    ws = None
    ws_bytes = 0
    if splits > 1:
        ws_bytes = lib.b200mm_gemm_workspace_bytes(M, N, splits)
        ws = torch.empty(ws_bytes // 4, device=a.device, dtype=torch.float32)
    g = _lib.GemmArgs()
    g.A, g.lda, g.a_mn = a.data_ptr(), lda, int(a_mn)
    g.B, g.ldb, g.b_mn = b.data_ptr(), ldb, int(b_mn)
    g.D, g.ldd, g.d_f32 = out.data_ptr(), out.stride(0), int(out_f32)
    g.M, g.N, g.K = M, N, K
    g.alpha = alpha
    g.bias = _req(bias, "bias", BF16, 1).data_ptr() if bias is not None else None
    g.act = act
    g.aux_out = aux.data_ptr() if aux is not None else None
    if dact_in is not None:
        g.dact_in, g.ld_dact = dact_in.data_ptr(), _row_major_2d(dact_in, "dact_in")
    if residual is not None:
        g.residual, g.ldr = residual.data_ptr(), _row_major_2d(residual, "residual")
    g.splits = splits
    if drop is not None and drop[0] > 0.0:
        g.drop_p, g.drop_seed = float(drop[0]), int(drop[1]) & 0xFFFFFFFFFFFFFFFF
    g.workspace = ws.data_ptr() if ws is not None else None
    g.workspace_bytes = ws_bytes
    prof = GEMM_PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(lib.b200mm_gemm_bf16(ctypes.byref(g), _stream()), "b200mm_gemm_bf16")
    if prof is not None:
        e1.record()
        prof.append((2.0 * M * N * K, e0, e1, splits, (M, N, K, int(a_mn), int(b_mn), int(bias is not None), act, int(aux_out),
                                                       int(dact_in is not None), int(residual is not None), int(out_f32))))
    _count(2 if splits > 1 else 1)
    return (out, aux) if aux_out else out


def layernorm_fwd(x, w, b, eps, add0=None, add1=None, add_period=0, want_sum=False):
    """y = LN(x + add0[row % period] + add1·[row % period == 0]); returns (y, s_or_None, mean, rstd)."""
    lib = _lib.load()
    _req(x, "x", BF16, 2)
    rows, W = x.shape
    y = torch.empty_like(x)
    s = torch.empty_like(x) if want_sum else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _lib.check(
        lib.b200mm_layernorm_fwd(_ptr(x), _ptr(add0), _ptr(add1), add_period, _ptr(_req(w, "w", BF16, 1)), _ptr(_req(b, "b", BF16, 1)),
                                 _ptr(y), _ptr(s), _ptr(mean), _ptr(rstd), rows, W, eps, _stream()),
        "b200mm_layernorm_fwd")
    _count(1)
    return y, s, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, w, dw32, db32, dadd=None):
    """dx = LN'(dy) (+ dadd); dw32/db32 (f32 [W]) are accumulated into."""
    lib = _lib.load()
    _req(dy, "dy", BF16, 2)
    rows, W = dy.shape
    dx = torch.empty_like(dy)
    _lib.check(
        lib.b200mm_layernorm_bwd(_ptr(dy), _ptr(_req(x, "x", BF16, 2)), _ptr(mean), _ptr(rstd), _ptr(w), _ptr(dadd), _ptr(dx),
                                 _ptr(_req(dw32, "dw", torch.float32)), _ptr(_req(db32, "db", torch.float32)), rows, W, _stream()),
        "b200mm_layernorm_bwd")
    _count(1)
    return dx


def act_layernorm_fwd(u, act, w, b, eps):
    """y = LN(act(u)) * w + b for rows of any width (sub-LN of the M2-Encoder blocks); returns (y, mean, rstd)."""
    lib = _lib.load()
    _req(u, "u", BF16, 2)
    if not u.is_contiguous():
        raise _lib.B200mmError("b200mm.act_layernorm_fwd: u must be contiguous")
    rows, W = u.shape
    y = torch.empty_like(u)
    mean = torch.empty(rows, device=u.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=u.device, dtype=torch.float32)
    _lib.check(lib.b200mm_act_layernorm_fwd(_ptr(u), act, _ptr(_req(w, "w", BF16, 1)), _ptr(_req(b, "b", BF16, 1)), _ptr(y), _ptr(mean),
                                            _ptr(rstd), rows, W, eps, _stream()), "b200mm_act_layernorm_fwd")
    _count(1)
    return y, mean, rstd


def act_layernorm_bwd(dy, u, act, mean, rstd, w, dw32, db32):
    """du = LN'(dy) * act'(u) (act(u) recomputed inside); dw32/db32 (f32 [W]) are accumulated into."""
    lib = _lib.load()
    _req(dy, "dy", BF16, 2)
    _req(u, "u", BF16, 2)
    if not (dy.is_contiguous() and u.is_contiguous()):
        raise _lib.B200mmError("b200mm.act_layernorm_bwd: dy and u must be contiguous")
    rows, W = dy.shape
    du = torch.empty_like(dy)
    _lib.check(lib.b200mm_act_layernorm_bwd(_ptr(dy), _ptr(u), act, _ptr(mean), _ptr(rstd), _ptr(w), _ptr(du),
                                            _ptr(_req(dw32, "dw", torch.float32)), _ptr(_req(db32, "db", torch.float32)), rows, W, _stream()),
               "b200mm_act_layernorm_bwd")
    _count(1)
    return du


def mask_rows(x, drop, inplace=False):
    """y[r, :] = 0 where drop[r] (uint8 [rows]) else x[r, :]."""
    lib = _lib.load()
    _req(x, "x", BF16, 2)
    _req(drop, "drop", torch.uint8, 1)
    if not x.is_contiguous():
        raise _lib.B200mmError("b200mm.mask_rows: x must be contiguous")
    y = x if inplace else torch.empty_like(x)
    _lib.check(lib.b200mm_mask_rows(_ptr(x), _ptr(drop), _ptr(y), x.shape[0], x.shape[1], _stream()), "b200mm_mask_rows")
    _count(1)
    return y


def embed_layernorm_fwd(word, ids, pos, L, type_table, type_ids, w, b, eps):
    lib = _lib.load()
    _req(word, "word", BF16, 2)
    _req(ids, "ids", torch.int64)
    _req(type_ids, "type_ids", torch.int64)
    rows, W = ids.numel(), word.shape[1]
    y = torch.empty((rows, W), device=word.device, dtype=BF16)
    s = torch.empty_like(y)
    mean = torch.empty(rows, device=word.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    _lib.check(
        lib.b200mm_embed_layernorm_fwd(_ptr(word), _ptr(ids), _ptr(_req(pos, "pos", BF16, 2)), L, _ptr(_req(type_table, "type", BF16, 2)),
                                       _ptr(type_ids), _ptr(w), _ptr(b), _ptr(y), _ptr(s), _ptr(mean), _ptr(rstd), rows, W, eps, _stream()),
        "b200mm_embed_layernorm_fwd")
    _count(1)
    return y, s, mean, rstd


def dropout(x, p, seed, out=None):
    """y = x * keep(seed, row, col) / (1 - p) on a bf16 [rows, cols] matrix (counter-based mask: the same call on the gradient is the
    backward pass; the GEMM epilogue with drop=(p, seed) yields the same mask). out may be x (in place)."""
    lib = _lib.load()
    ldx = _row_major_2d(x, "x")
    rows, cols = x.shape
    y = torch.empty((rows, cols), device=x.device, dtype=BF16) if out is None else out
    ldy = _row_major_2d(y, "out")
    _lib.check(lib.b200mm_dropout(_ptr(x), ldx, _ptr(y), ldy, rows, cols, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()), "b200mm_dropout")
    _count(1)
    return y


def attention_dropout_mask(B, H, L, p, seed, device="cuda"):
    """bool [B, H, L, L]: which attention probabilities attention_fwd / attention_bwd keep under drop=(p, seed) (test / debugging aid)."""
    lib = _lib.load()
    keep = torch.empty((B, H, L, L), device=device, dtype=torch.uint8)
    _lib.check(lib.b200mm_attention_dropout_mask(_ptr(keep), B, H, L, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()),
               "b200mm_attention_dropout_mask")
    _count(1)
    return keep.bool()


def _drop_args(drop):
    if drop is None or drop[0] <= 0.0:
        return 0.0, 0
    return float(drop[0]), int(drop[1]) & 0xFFFFFFFFFFFFFFFF


def attention_fwd(qkv, B, L, H, hd, key_bias=None, q_off=0, k_off=None, v_off=None, drop=None):
    """qkv: [B*L, ld] fused projection; returns (o [B*L, H*hd], lse [B, H, L]). drop = (p, seed): dropout on the probabilities."""
    lib = _lib.load()
    ld = _row_major_2d(qkv, "qkv")
    W = H * hd
    k_off = W if k_off is None else k_off
    v_off = 2 * W if v_off is None else v_off
    o = torch.empty((B * L, W), device=qkv.device, dtype=BF16)
    lse = torch.empty((B, H, L), device=qkv.device, dtype=torch.float32)
    if key_bias is not None:
        _req(key_bias, "key_bias", torch.float32, 2)
    dp, ds = _drop_args(drop)
    _lib.check(
        lib.b200mm_attention_fwd_dropout(_ptr(qkv), ld, q_off, k_off, v_off, _ptr(o), W, _ptr(lse), _ptr(key_bias), B, H, L, hd,
                                         1.0 / math.sqrt(hd), dp, ds, _stream()),
        "b200mm_attention_fwd")
    _count(1)
    return o, lse


def attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=None, q_off=0, k_off=None, v_off=None, drop=None):
    lib = _lib.load()
    ld = _row_major_2d(qkv, "qkv")
    W = H * hd
    k_off = W if k_off is None else k_off
    v_off = 2 * W if v_off is None else v_off
    _req(d_o, "d_o", BF16, 2)
    if not d_o.is_contiguous():
        raise _lib.B200mmError("b200mm.attention_bwd: d_o must be contiguous")
    dqkv = torch.empty_like(qkv)
    dsum = torch.empty_like(lse)
    dp, ds = _drop_args(drop)
    _lib.check(
        lib.b200mm_attention_bwd_dropout(_ptr(qkv), ld, q_off, k_off, v_off, _ptr(o), _ptr(d_o), W, _ptr(lse), _ptr(key_bias), _ptr(dqkv),
                                         _ptr(dsum), B, H, L, hd, 1.0 / math.sqrt(hd), dp, ds, _stream()),
        "b200mm_attention_bwd")
    _count(2)
    return dqkv


def act_fwd(x, act):
    lib = _lib.load()
    _req(x, "x", BF16)
    y = torch.empty_like(x)
    _lib.check(lib.b200mm_act_fwd(_ptr(x), _ptr(y), x.numel(), act, _stream()), "b200mm_act_fwd")
    _count(1)
    return y


def rowsum_periodic(x, out32, period=1):
    """out32[(row % period), :] += x[row, :]"""
    lib = _lib.load()
    _req(x, "x", BF16, 2)
    _req(out32, "out", torch.float32)
    _lib.check(lib.b200mm_rowsum_periodic(_ptr(x), _ptr(out32), x.shape[0], x.shape[1], period, _stream()), "b200mm_rowsum_periodic")
    _count(1)


def scatter_add_rows(x, ids, out32, skip_id=-1):
    lib = _lib.load()
    _req(x, "x", BF16, 2)
    _req(ids, "ids", torch.int64)
    _req(out32, "out", torch.float32, 2)
    _lib.check(lib.b200mm_scatter_add_rows(_ptr(x), _ptr(ids), _ptr(out32), x.shape[0], x.shape[1], skip_id, out32.shape[0], _stream()),
               "b200mm_scatter_add_rows")
    _count(1)


def cast_f32_bf16(x32, scale=1.0):
    lib = _lib.load()
    _req(x32, "x", torch.float32)
    y = torch.empty(x32.shape, device=x32.device, dtype=BF16)
    _lib.check(lib.b200mm_cast_f32_bf16(_ptr(x32), _ptr(y), x32.numel(), scale, _stream()), "b200mm_cast_f32_bf16")
    _count(1)
    return y


def rownorm_fwd(x, eps=1e-12):
    lib = _lib.load()
    _req(x, "x", BF16, 2)
    y = torch.empty_like(x)
    inv = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
    _lib.check(lib.b200mm_rownorm_fwd(_ptr(x), _ptr(y), _ptr(inv), x.shape[0], x.shape[1], eps, _stream()), "b200mm_rownorm_fwd")
    _count(1)
    return y, inv


def rownorm_bwd(dy32, x, inv):
    lib = _lib.load()
    _req(dy32, "dy", torch.float32, 2)
    dx = torch.empty_like(x)
    _lib.check(lib.b200mm_rownorm_bwd(_ptr(dy32), _ptr(x), _ptr(inv), _ptr(dx), x.shape[0], x.shape[1], _stream()), "b200mm_rownorm_bwd")
    _count(1)
    return dx


def im2row(img, p, Kp):
    """[B, C, H, W] bf16 -> [B*(Np+1), Kp] patch matrix with a zero row per image for the class token."""
    lib = _lib.load()
    _req(img, "img", BF16, 4)
    if not img.is_contiguous():
        raise _lib.B200mmError("b200mm.im2row: image must be contiguous NCHW")
    B, C, H, W = img.shape
    L = (H // p) * (W // p) + 1
    out = torch.empty((B * L, Kp), device=img.device, dtype=BF16)
    _lib.check(lib.b200mm_im2row(_ptr(img), _ptr(out), B, C, H, W, p, Kp, _stream()), "b200mm_im2row")
    _count(1)
    return out


NO_DIAG = -(1 << 40)  # diag_off that never matches a column: "this block has no positive column"


def contrast_lse_partials(a, b, alpha, diag_off, b_mn=False):
    """Per-row (max, sum-exp) partials of z = alpha * a·b^T over 256-column tiles + the diagonal logit.
    b: [N, K] rows, or with b_mn=True stored [K, N] (the MoCo queue layout [dim, K_queue])."""
    lib = _lib.load()
    lda = _row_major_2d(a, "a")
    ldb = _row_major_2d(b, "b")
    M, K = a.shape
    N = b.shape[1] if b_mn else b.shape[0]
    nt = lib.b200mm_contrast_num_tiles(N)
    pmax = torch.empty((M, nt), device=a.device, dtype=torch.float32)
    psum = torch.empty_like(pmax)
    diag = torch.zeros(M, device=a.device, dtype=torch.float32)
    _lib.check(lib.b200mm_contrast_lse_partials(_ptr(a), lda, _ptr(b), ldb, int(b_mn), M, N, K, alpha, diag_off, _ptr(pmax), _ptr(psum), _ptr(diag),
                                                _stream()), "b200mm_contrast_lse_partials")
    _count(1)
    return pmax, psum, diag


def contrast_lse_merge(partsA, partsB, diag, sub_diag, loss_sum):
    lib = _lib.load()
    maxA, sumA = partsA
    M = maxA.shape[0]
    lse = torch.empty(M, device=maxA.device, dtype=torch.float32)
    maxB, sumB = partsB if partsB is not None else (None, None)
    _lib.check(lib.b200mm_contrast_lse_merge(_ptr(maxA), _ptr(sumA), maxA.shape[1], _ptr(maxB), _ptr(sumB),
                                             maxB.shape[1] if maxB is not None else 0, _ptr(diag), int(sub_diag), _ptr(lse),
                                             _ptr(loss_sum), M, _stream()), "b200mm_contrast_lse_merge")
    _count(1)
    return lse


def contrast_softgrad(a, b, n_valid, alpha, diag_off, row_lse, coef, diag_sub, diag_zero, dscale, b_mn=False):
    """G = alpha*coef*(exp(z - row_lse) - diag_sub*[diag]) as bf16 [M, N]; dscale (f32 scalar tensor or None) += sum dL/dz * z.
    b may carry zero padding rows up to a multiple of 8; columns >= n_valid are zeroed."""
    lib = _lib.load()
    lda = _row_major_2d(a, "a")
    ldb = _row_major_2d(b, "b")
    M, K = a.shape
    N = b.shape[1] if b_mn else b.shape[0]
    G = torch.empty((M, N), device=a.device, dtype=BF16)
    _lib.check(lib.b200mm_contrast_softgrad(_ptr(a), lda, _ptr(b), ldb, int(b_mn), M, N, K, n_valid, alpha, diag_off, _ptr(row_lse), coef, diag_sub,
                                            int(diag_zero), _ptr(G), N, _ptr(dscale), _stream()), "b200mm_contrast_softgrad")
    _count(1)
    return G


def contrast_lse_partials_pair(a0, b0, a1, b1, alpha, diag_off, alpha_dev=None):
    """Both directions of a symmetric loss in one grouped launch: ((pmax0, psum0, diag0), (pmax1, psum1, diag1)) for z0 = alpha a0·b0^T and
    z1 = alpha a1·b1^T (same shapes). alpha_dev: device f32 scalar that replaces `alpha` (no host read of the temperature)."""
    lib = _lib.load()
    M, K = a0.shape
    N = b0.shape[0]
    if tuple(a1.shape) != (M, K) or tuple(b1.shape) != (N, K):
        raise _lib.B200mmError("b200mm.contrast_lse_partials_pair: the two problems must have the same shapes")
    nt = lib.b200mm_contrast_num_tiles(N)
    outs = []
    for _ in range(2):
        pmax = torch.empty((M, nt), device=a0.device, dtype=torch.float32)
        outs.append((pmax, torch.empty_like(pmax), torch.zeros(M, device=a0.device, dtype=torch.float32)))
    if alpha_dev is not None:
        _req(alpha_dev, "alpha_dev", torch.float32)
    _lib.check(lib.b200mm_contrast_lse_partials_pair(_ptr(a0), _row_major_2d(a0, "a0"), _ptr(b0), _row_major_2d(b0, "b0"), _ptr(a1),
                                                     _row_major_2d(a1, "a1"), _ptr(b1), _row_major_2d(b1, "b1"), M, N, K, float(alpha),
                                                     _ptr(alpha_dev), diag_off, _ptr(outs[0][0]), _ptr(outs[0][1]), _ptr(outs[0][2]),
                                                     _ptr(outs[1][0]), _ptr(outs[1][1]), _ptr(outs[1][2]), _stream()),
               "b200mm_contrast_lse_partials_pair")
    _count(1)
    return outs[0], outs[1]


def contrast_softgrad_pair(a0, b0, a1, b1, n_valid, alpha, diag_off, row_lse0, col_lse0, row_lse1, col_lse1, coef, diag_sub, zero_flags, dscale,
                           alpha_dev=None, coef_dev=None):
    """Two-sided softmax-gradient tiles of both directions in one grouped launch (include/b200mm.h: b200mm_contrast_softgrad_pair):
    G_k[m, n] = alpha*coef*(wr exp(z_k - row_lse_k[m]) + wc exp(z_k - col_lse_k[n]) - diag_sub [diag]) as bf16 [M, N];
    zero_flags = (row_diag_zero0, col_diag_zero0, row_diag_zero1, col_diag_zero1); dscale (f32 [1] or None) += sum dL/dz z over problem 0."""
    lib = _lib.load()
    M, K = a0.shape
    N = b0.shape[0]
    if tuple(a1.shape) != (M, K) or tuple(b1.shape) != (N, K):
        raise _lib.B200mmError("b200mm.contrast_softgrad_pair: the two problems must have the same shapes")
    for t, nm, n_ in ((row_lse0, "row_lse0", M), (row_lse1, "row_lse1", M), (col_lse0, "col_lse0", n_valid), (col_lse1, "col_lse1", n_valid)):
        _req(t, nm, torch.float32, 1)
        if t.numel() < n_ or not t.is_contiguous():
            raise _lib.B200mmError(f"b200mm.contrast_softgrad_pair: {nm} must be contiguous with >= {n_} entries")
    G0 = torch.empty((M, N), device=a0.device, dtype=BF16)
    G1 = torch.empty((M, N), device=a0.device, dtype=BF16)
    _lib.check(lib.b200mm_contrast_softgrad_pair(_ptr(a0), _row_major_2d(a0, "a0"), _ptr(b0), _row_major_2d(b0, "b0"), _ptr(a1),
                                                 _row_major_2d(a1, "a1"), _ptr(b1), _row_major_2d(b1, "b1"), M, N, K, n_valid, float(alpha),
                                                 _ptr(alpha_dev), diag_off, _ptr(row_lse0), _ptr(col_lse0), _ptr(row_lse1), _ptr(col_lse1),
                                                 float(coef), _ptr(coef_dev), float(diag_sub), *[int(z) for z in zero_flags], _ptr(G0), _ptr(G1), N,
                                                 _ptr(dscale), _stream()), "b200mm_contrast_softgrad_pair")
    _count(1)
    return G0, G1


def contrast_rank(a, b, alpha, ref, diag_off=0, gt_col=None, b_mn=False):
    """rank[m] = #{n != pos(m): alpha*<a_m, b_n> > ref[m]} (int32 [M]); pos(m) = gt_col[m] (int32) or m + diag_off.
    The [M, N] similarity only exists tile by tile in TMEM."""
    lib = _lib.load()
    lda = _row_major_2d(a, "a")
    ldb = _row_major_2d(b, "b")
    M, K = a.shape
    N = b.shape[1] if b_mn else b.shape[0]
    _req(ref, "ref", torch.float32, 1)
    if gt_col is not None:
        _req(gt_col, "gt_col", torch.int32, 1)
    rank = torch.zeros(M, device=a.device, dtype=torch.int32)
    _lib.check(lib.b200mm_contrast_rank(_ptr(a), lda, _ptr(b), ldb, int(b_mn), M, N, K, alpha, diag_off, _ptr(gt_col), _ptr(ref), _ptr(rank),
                                        _stream()), "b200mm_contrast_rank")
    _count(1)
    return rank


def masked_mean_fwd(x, pad):
    """x [R, P, W] bf16, pad [R, P] bool/uint8 (True = padded) or None -> (y [R, W] bf16, inv_count [R] f32)."""
    lib = _lib.load()
    _req(x, "x", BF16, 3)
    x = x.contiguous()
    R, P, W = x.shape
    if pad is not None:
        pad = _req(pad, "pad", None, 2).to(torch.uint8).contiguous()
    y = torch.empty((R, W), device=x.device, dtype=BF16)
    inv = torch.empty(R, device=x.device, dtype=torch.float32)
    _lib.check(lib.b200mm_masked_mean_fwd(_ptr(x), _ptr(pad), _ptr(y), _ptr(inv), R, P, W, _stream()), "b200mm_masked_mean_fwd")
    _count(1)
    return y, inv, pad


def masked_mean_bwd(dy, pad, inv, P):
    lib = _lib.load()
    _req(dy, "dy", BF16, 2)
    dy = dy.contiguous()
    R, W = dy.shape
    dx = torch.empty((R, P, W), device=dy.device, dtype=BF16)
    _lib.check(lib.b200mm_masked_mean_bwd(_ptr(dy), _ptr(pad), _ptr(inv), _ptr(dx), R, P, W, _stream()), "b200mm_masked_mean_bwd")
    _count(1)
    return dx


def rowdot(a, b, scale=1.0):
    """out[r] = scale * <a[r], b[r]> (f32)."""
    lib = _lib.load()
    _req(a, "a", BF16, 2)
    _req(b, "b", BF16, 2)
    out = torch.empty(a.shape[0], device=a.device, dtype=torch.float32)
    _lib.check(lib.b200mm_rowdot(_ptr(a.contiguous()), _ptr(b.contiguous()), _ptr(out), a.shape[0], a.shape[1], scale, _stream()), "b200mm_rowdot")
    _count(1)
    return out


def ema_update(pk32, pq, m):
    """pk32 (f32, in place) = m * pk32 + (1 - m) * pq (bf16 or f32)."""
    lib = _lib.load()
    _req(pk32, "pk", torch.float32)
    if not (pq.is_cuda and pq.dtype in (BF16, torch.float32)) or pq.numel() != pk32.numel():
        raise _lib.B200mmError("b200mm.ema_update: pq must be a CUDA bf16/f32 tensor of the same size")
    if not (pk32.is_contiguous() and pq.is_contiguous()):
        raise _lib.B200mmError("b200mm.ema_update: tensors must be contiguous")
    _lib.check(lib.b200mm_ema_update(_ptr(pk32), _ptr(pq), int(pq.dtype == BF16), pk32.numel(), m, _stream()), "b200mm_ema_update")
    _count(1)


def gather_rows(src, ids):
    """out[r, :] = src[ids[r], :] (bf16 [n, W], int64 [rows]) — bit-exact row gather."""
    lib = _lib.load()
    _req(src, "src", BF16, 2)
    _req(ids, "ids", torch.int64, 1)
    if not (src.is_contiguous() and ids.is_contiguous()):
        raise _lib.B200mmError("b200mm.gather_rows: src and ids must be contiguous")
    out = torch.empty((ids.numel(), src.shape[1]), device=src.device, dtype=BF16)
    _lib.check(lib.b200mm_gather_rows(_ptr(src), _ptr(ids), _ptr(out), ids.numel(), src.shape[0], src.shape[1], _stream()), "b200mm_gather_rows")
    _count(1)
    return out


def relu_fwd(x):
    lib = _lib.load()
    _req(x, "x", BF16)
    if not x.is_contiguous():
        raise _lib.B200mmError("b200mm.relu_fwd: x must be contiguous")
    y = torch.empty_like(x)
    _lib.check(lib.b200mm_relu_fwd(_ptr(x), _ptr(y), x.numel(), _stream()), "b200mm_relu_fwd")
    _count(1)
    return y


def relu_bwd(dy, x):
    lib = _lib.load()
    _req(dy, "dy", BF16)
    _req(x, "x", BF16)
    if not (x.is_contiguous() and dy.is_contiguous()):
        raise _lib.B200mmError("b200mm.relu_bwd: tensors must be contiguous")
    dx = torch.empty_like(x)
    _lib.check(lib.b200mm_relu_bwd(_ptr(dy), _ptr(x), _ptr(dx), x.numel(), _stream()), "b200mm_relu_bwd")
    _count(1)
    return dx


def mil_nce_matrix_fwd(S, w=None):
    """S f32 [B, B] (row stride free), w f32 [B] or None -> (lse [B], loss_sum scalar tensor = sum_j w_j (lse_j - S_jj))."""
    lib = _lib.load()
    _req(S, "S", torch.float32, 2)
    if S.shape[0] != S.shape[1] or S.stride(1) != 1:
        raise _lib.B200mmError(f"b200mm.mil_nce_matrix_fwd: S must be square with unit inner stride, got {tuple(S.shape)} / {S.stride()}")
    if w is not None:
        _req(w, "w", torch.float32, 1)
    B = S.shape[0]
    lse = torch.empty(B, device=S.device, dtype=torch.float32)
    loss_sum = torch.zeros((), device=S.device, dtype=torch.float32)
    _lib.check(lib.b200mm_mil_nce_matrix_fwd(_ptr(S), S.stride(0), _ptr(w), _ptr(lse), _ptr(loss_sum), B, _stream()), "b200mm_mil_nce_matrix_fwd")
    _count(1)
    return lse, loss_sum


def mil_nce_matrix_bwd(S, w, lse, gout):
    lib = _lib.load()
    B = S.shape[0]
    dS = torch.empty((B, B), device=S.device, dtype=torch.float32)
    _lib.check(lib.b200mm_mil_nce_matrix_bwd(_ptr(S), S.stride(0), _ptr(w), _ptr(lse), _ptr(_req(gout, "gout", torch.float32)), _ptr(dS), B, _stream()),
               "b200mm_mil_nce_matrix_bwd")
    _count(1)
    return dS


def xpos_apply(qkv, tables, B, L, H, hd, backward=False, q_off=0, k_off=None):
    """In place: XPOS rotation + scale of the q and k sections of qkv [B*L, ld]; tables = (q_cos, q_sin, k_cos, k_sin), f32 [L, hd/2]."""
    lib = _lib.load()
    ld = _row_major_2d(qkv, "qkv")
    k_off = H * hd if k_off is None else k_off
    for t in tables:
        _req(t, "xpos table", torch.float32, 2)
        if tuple(t.shape) != (L, hd // 2) or not t.is_contiguous():
            raise _lib.B200mmError(f"b200mm.xpos_apply: tables must be contiguous [L, hd/2] = [{L}, {hd // 2}], got {tuple(t.shape)}")
    _lib.check(lib.b200mm_xpos_apply(_ptr(qkv), ld, q_off, k_off, _ptr(tables[0]), _ptr(tables[1]), _ptr(tables[2]), _ptr(tables[3]), B * L, L, H, hd,
                                     int(backward), _stream()), "b200mm_xpos_apply")
    _count(1)
    return qkv
