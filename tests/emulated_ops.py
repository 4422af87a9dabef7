"""TEST INFRASTRUCTURE — torch (CPU) stand-ins for the b200mm.ops kernel wrappers, used ONLY by the `not gpu` host-logic
tests to run the autograd glue of b200mm.functional / b200mm.modules (saved-tensor order, gradient routing, parameter
naming, masking) without a GPU. They follow the documented contract of each C-ABI entry point (include/b200mm.h) in fp32
and round to bf16 where the kernels do. Never imported by the package: the product path has no CPU route.
"""
import contextlib
import math

import torch
import torch.nn.functional as F

BF = torch.bfloat16
ACT_NONE, ACT_QUICKGELU, ACT_GELU_ERF = 0, 1, 2


def _act(act, x):
    return x * torch.sigmoid(1.702 * x) if act == ACT_QUICKGELU else (F.gelu(x) if act == ACT_GELU_ERF else x)


def _dact(act, x):
    with torch.enable_grad():  # called from inside autograd backward, where grad mode is off
        x = x.detach().clone().requires_grad_()
        (g,) = torch.autograd.grad(_act(act, x).sum(), x)
    return g


def _keep(rows, cols, p, seed):
    from oracle import restated

    return restated.dropout_keep(seed, rows, cols, p)


def dropout(x, p, seed, out=None):
    y = (x.float() * _keep(x.shape[0], x.shape[1], p, seed) / (1.0 - p)).to(BF)
    if out is not None:
        out.copy_(y)
        return out
    return y


def gemm(a, b, *, a_mn=False, b_mn=False, bias=None, act=ACT_NONE, aux_out=False, dact_in=None, residual=None, alpha=1.0, out=None,
         out_f32=False, splits=None, drop=None):
    A = a.float().t() if a_mn else a.float()
    Bm = b.float() if b_mn else b.float().t()
    d = alpha * (A @ Bm)
    if bias is not None:
        d = d + bias.float()
    aux = None
    if dact_in is not None:
        aux = _act(act, dact_in.float()).to(BF)
        d = d * _dact(act, dact_in.float())
    else:
        aux = d.to(BF)
        d = _act(act, d)
    if drop is not None and drop[0] > 0:
        d = d * _keep(d.shape[0], d.shape[1], drop[0], drop[1]) / (1.0 - drop[0])
    if residual is not None:
        d = d + residual.float()
    d = d if out_f32 else d.to(BF)
    if out is not None:  # the kernel writes into a caller-provided (possibly strided) buffer
        out.copy_(d)
        d = out
    return (d, aux) if aux_out else d


def layernorm_fwd(x, w, b, eps, add0=None, add1=None, add_period=0, want_sum=False):
    s = x.float()
    if add0 is not None or add1 is not None:
        rows = s.shape[0]
        r = torch.arange(rows) % add_period
        if add0 is not None:
            s = s + add0.float()[r]
        if add1 is not None:
            s = s + (r == 0).float()[:, None] * add1.float()[None, :]
        s = s.to(BF).float()
    mean = s.mean(-1)
    rstd = torch.rsqrt(s.var(-1, unbiased=False) + eps)
    y = ((s - mean[:, None]) * rstd[:, None] * w.float() + b.float()).to(BF)
    return y, (s.to(BF) if want_sum else None), mean, rstd


def layernorm_bwd(dy, x, mean, rstd, w, dw32, db32, dadd=None):
    xh = (x.float() - mean[:, None]) * rstd[:, None]
    gy = dy.float() * w.float()
    dx = rstd[:, None] * (gy - gy.mean(-1, keepdim=True) - xh * (gy * xh).mean(-1, keepdim=True))
    if dadd is not None:
        dx = dx + dadd.float()
    dw32 += (dy.float() * xh).sum(0)
    db32 += dy.float().sum(0)
    return dx.to(BF)


def act_layernorm_fwd(u, act, w, b, eps):
    g = _act(act, u.float())
    mean = g.mean(-1)
    rstd = torch.rsqrt(g.var(-1, unbiased=False) + eps)
    return ((g - mean[:, None]) * rstd[:, None] * w.float() + b.float()).to(BF), mean, rstd


def act_layernorm_bwd(dy, u, act, mean, rstd, w, dw32, db32):
    g = _act(act, u.float())
    xh = (g - mean[:, None]) * rstd[:, None]
    gy = dy.float() * w.float()
    dg = rstd[:, None] * (gy - gy.mean(-1, keepdim=True) - xh * (gy * xh).mean(-1, keepdim=True))
    dw32 += (dy.float() * xh).sum(0)
    db32 += dy.float().sum(0)
    return (dg * _dact(act, u.float())).to(BF)


def mask_rows(x, drop, inplace=False):
    y = x * (1 - drop.to(x.dtype))[:, None]
    if inplace:
        x.copy_(y)
        return x
    return y


def embed_layernorm_fwd(word, ids, pos, L, type_table, type_ids, w, b, eps):
    rows = ids.numel()
    s = (word.float()[ids] + pos.float()[torch.arange(rows) % L] + type_table.float()[type_ids]).to(BF)
    y, _, mean, rstd = layernorm_fwd(s, w, b, eps)
    return y, s, mean, rstd


def _attn(qkv, B, L, H, hd, key_bias, drop=None):
    W = H * hd
    q, k, v = (qkv[:, i * W:(i + 1) * W].reshape(B, L, H, hd).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) / math.sqrt(hd)
    if key_bias is not None:
        s = s + key_bias[:, None, None, :]
    pr = torch.softmax(s, -1)
    if drop is not None and drop[0] > 0:
        from oracle import restated

        pr = pr * restated.attention_dropout_keep(drop[1], B, H, L, drop[0]) / (1.0 - drop[0])
    return (pr @ v).transpose(1, 2).reshape(B * L, W), torch.logsumexp(s, -1)


def attention_fwd(qkv, B, L, H, hd, key_bias=None, drop=None, **kw):
    o, lse = _attn(qkv.float(), B, L, H, hd, key_bias, drop)
    return o.to(BF), lse


def attention_bwd(qkv, o, d_o, lse, B, L, H, hd, key_bias=None, drop=None, **kw):
    with torch.enable_grad():
        x = qkv.float().detach().requires_grad_()
        out, _ = _attn(x, B, L, H, hd, key_bias, drop)
        (g,) = torch.autograd.grad(out, x, d_o.float())
    return g.to(BF)


def act_fwd(x, act):
    return _act(act, x.float()).to(BF)


def rowsum_periodic(x, out32, period=1):
    rows, W = x.shape
    out32.view(period, W).add_(x.float().view(rows // period, period, W).sum(0))


def scatter_add_rows(x, ids, out32, skip_id=-1):
    keep = ids != skip_id
    out32.index_add_(0, ids[keep], x.float()[keep])


def cast_f32_bf16(x32, scale=1.0):
    return (x32 * scale).to(BF)


def rownorm_fwd(x, eps=1e-12):
    inv = 1.0 / x.float().norm(dim=-1).clamp_min(eps)
    return (x.float() * inv[:, None]).to(BF), inv


def rownorm_bwd(dy32, x, inv):
    y = x.float() * inv[:, None]
    return ((dy32 - y * (dy32 * y).sum(-1, keepdim=True)) * inv[:, None]).to(BF)


def im2row(img, p, Kp):
    B, C, H, W = img.shape
    g = H // p
    x = img.reshape(B, C, g, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * (W // p), C * p * p)
    out = torch.zeros(B, g * (W // p) + 1, Kp, dtype=BF)
    out[:, 1:, : C * p * p] = x
    return out.reshape(-1, Kp)


def masked_mean_fwd(x, pad):
    R, P, W = x.shape
    if pad is None:
        inv = torch.full((R,), 1.0 / P)
        return (x.float().mean(1)).to(BF), inv, None
    pad8 = pad.to(torch.uint8)
    keep = (pad8 == 0).float()
    inv = 1.0 / keep.sum(1)
    return ((x.float() * keep[:, :, None]).sum(1) * inv[:, None]).to(BF), inv, pad8


def masked_mean_bwd(dy, pad, inv, P):
    R, W = dy.shape
    keep = torch.ones(R, P) if pad is None else (pad == 0).float()
    return (dy.float()[:, None, :] * (keep * inv[:, None])[:, :, None]).to(BF)


def xpos_apply(qkv, tables, B, L, H, hd, backward=False, q_off=0, k_off=None):
    W = H * hd
    k_off = W if k_off is None else k_off
    q_cos, q_sin, k_cos, k_sin = tables
    sign = -1.0 if backward else 1.0
    for off, cs, sn in ((q_off, q_cos, q_sin), (k_off, k_cos, k_sin)):
        x = qkv[:, off:off + W].float().reshape(B, L, H, hd // 2, 2)
        c, s = cs[None, :, None, :], sn[None, :, None, :] * sign
        out = torch.stack((c * x[..., 0] - s * x[..., 1], c * x[..., 1] + s * x[..., 0]), dim=-1)
        qkv[:, off:off + W] = out.reshape(B * L, W).to(qkv.dtype)
    return qkv


def gather_rows(src, ids):
    ok = (ids >= 0) & (ids < src.shape[0])
    out = src[ids.clamp(0, src.shape[0] - 1)].clone()
    out[~ok] = 0
    return out


def relu_fwd(x):
    return torch.relu(x)


def relu_bwd(dy, x):
    return dy * (x > 0).to(dy.dtype)


def _mil_parts(S, w):
    B = S.shape[0]
    eye = torch.eye(B, dtype=torch.bool)
    both = torch.cat([S.t(), S.masked_fill(eye, float("-inf"))], dim=1)
    lse = torch.logsumexp(both, dim=1)
    ww = torch.ones(B) if w is None else w
    return lse, (ww * (lse - torch.diagonal(S))).sum()


def mil_nce_matrix_fwd(S, w=None):
    return _mil_parts(S.float(), w)


def mil_nce_matrix_bwd(S, w, lse, gout):
    with torch.enable_grad():
        x = S.float().detach().requires_grad_()
        _, total = _mil_parts(x, w)
        (g,) = torch.autograd.grad(total / S.shape[0], x)
    return g * gout


NO_DIAG = -(1 << 40)
_TILE = 256  # column tile of the contrastive epilogues (b200mm_contrast_num_tiles)


def _logits(a, b, alpha, b_mn):
    return alpha * (a.float() @ (b.float() if b_mn else b.float().t()))


def contrast_lse_partials(a, b, alpha, diag_off, b_mn=False):
    z = _logits(a, b, alpha, b_mn)
    M, N = z.shape
    nt = (N + _TILE - 1) // _TILE
    pmax = torch.full((M, nt), float("-inf"))
    psum = torch.zeros(M, nt)
    for t in range(nt):
        blk = z[:, t * _TILE:(t + 1) * _TILE]
        pmax[:, t] = blk.max(dim=1).values
        psum[:, t] = torch.exp(blk - pmax[:, t:t + 1]).sum(dim=1)
    diag = torch.zeros(M)
    cols = torch.arange(M) + diag_off
    ok = (cols >= 0) & (cols < N)
    diag[ok] = z[torch.arange(M)[ok], cols[ok]]
    return pmax, psum, diag


def contrast_lse_merge(partsA, partsB, diag, sub_diag, loss_sum):
    pm = partsA[0] if partsB is None else torch.cat([partsA[0], partsB[0]], dim=1)
    ps = partsA[1] if partsB is None else torch.cat([partsA[1], partsB[1]], dim=1)
    m = pm.max(dim=1).values
    sm = (ps * torch.exp(pm - m[:, None])).sum(dim=1)
    sub = int(sub_diag)
    if sub > 0:
        sm = sm - torch.exp(diag - m)
    elif sub < 0:
        nm = torch.maximum(m, diag)
        sm = sm * torch.exp(m - nm) + torch.exp(diag - nm)
        m = nm
    lse = m + torch.log(sm)
    loss_sum += (lse - diag).sum()
    return lse


def contrast_softgrad(a, b, n_valid, alpha, diag_off, row_lse, coef, diag_sub, diag_zero, dscale, b_mn=False):
    z = _logits(a, b, alpha, b_mn)
    M, N = z.shape
    dz = coef * torch.exp(z - row_lse[:, None])
    cols = torch.arange(M) + diag_off
    ok = (cols >= 0) & (cols < N)
    r = torch.arange(M)[ok]
    dz[r, cols[ok]] -= coef * diag_sub
    if diag_zero:
        dz[r, cols[ok]] = 0.0
    dz[:, n_valid:] = 0.0
    if dscale is not None:
        dscale += (dz * z).sum()
    return (alpha * dz).to(BF)


def contrast_lse_partials_pair(a0, b0, a1, b1, alpha, diag_off, alpha_dev=None):
    al = float(alpha_dev) if alpha_dev is not None else alpha
    return contrast_lse_partials(a0, b0, al, diag_off), contrast_lse_partials(a1, b1, al, diag_off)


def contrast_softgrad_pair(a0, b0, a1, b1, n_valid, alpha, diag_off, row_lse0, col_lse0, row_lse1, col_lse1, coef, diag_sub, zero_flags, dscale,
                           alpha_dev=None, coef_dev=None):
    al = float(alpha_dev) if alpha_dev is not None else alpha
    cf = coef * (float(coef_dev) if coef_dev is not None else 1.0)
    outs = []
    for k, (a, b, rl, cl) in enumerate(((a0, b0, row_lse0, col_lse0), (a1, b1, row_lse1, col_lse1))):
        z = _logits(a, b, al, False)
        M, N = z.shape
        er = torch.exp(z - rl[:, None])
        clp = torch.zeros(N)
        clp[:n_valid] = cl[:n_valid]
        ec = torch.exp(z - clp[None, :])
        cols = torch.arange(M) + diag_off
        ok = (cols >= 0) & (cols < N)
        r = torch.arange(M)[ok]
        if zero_flags[2 * k]:
            er[r, cols[ok]] = 0.0
        if zero_flags[2 * k + 1]:
            ec[r, cols[ok]] = 0.0
        dz = cf * (er + ec)
        dz[:, n_valid:] = 0.0
        full = dz.clone()
        full[r, cols[ok]] -= cf * diag_sub
        if k == 0 and dscale is not None:
            dscale += (full * z).sum()
        outs.append((al * dz).to(BF))  # the stored tiles leave the -diag_sub term to the caller (exact, fp32)
    return outs[0], outs[1]


def contrast_rank(a, b, alpha, ref, diag_off=0, gt_col=None, b_mn=False):
    z = _logits(a, b, alpha, b_mn)
    M, N = z.shape
    pos = gt_col.long() if gt_col is not None else torch.arange(M) + diag_off
    beats = z > ref[:, None]
    ok = (pos >= 0) & (pos < N)
    beats[torch.arange(M)[ok], pos[ok]] = False
    return beats.sum(dim=1).to(torch.int32)


def rowdot(a, b, scale=1.0):
    return scale * (a.float() * b.float()).sum(-1)


def ema_update(pk32, pq, m):
    pk32.mul_(m).add_(pq.float(), alpha=1.0 - m)


@contextlib.contextmanager
def patched():
    """Swaps the kernel wrappers of b200mm.ops for the stand-ins above (restored on exit)."""
    from b200mm import ops

    names = [n for n, f in globals().items() if (callable(f) or n == "NO_DIAG") and not n.startswith("_") and hasattr(ops, n) and n != "patched"]
    saved = {n: getattr(ops, n) for n in names}
    try:
        for n in names:
            setattr(ops, n, globals()[n])
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
