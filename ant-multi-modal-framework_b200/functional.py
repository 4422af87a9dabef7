"""Block-level autograd glue: each Function runs a fixed sequence of b200mm kernels forward and backward.

Block granularity (instead of per-op Functions) gives explicit control over what is kept for backward. Per ViT block
the saved set is {x, qkv, o, lse, x_mid, u, LN statistics} = 10 T·W bf16 elements; LayerNorm outputs and the activated
MLP hidden are recomputed in backward (HBM-bound, ~5 % of the block's GEMM time). With ``checkpoint=True`` only x is
kept and the whole block forward is re-run in backward.

Nothing here computes: all arithmetic is in the CUDA kernels behind b200mm.ops. torch supplies allocation, views,
autograd bookkeeping and the occasional tiny copy (torch.cat of q/k/v weights, class-token row gather).
"""
import torch
from torch.autograd import Function

from . import ops
from .ops import ACT_GELU_ERF, ACT_NONE, ACT_QUICKGELU  # noqa: F401

BF16 = torch.bfloat16


class _VecGrads:
    """fp32 accumulators for the small per-block vector gradients (biases, LN affine) in ONE zero-filled buffer,
    converted to bf16 views with a single cast kernel."""

    def __init__(self, device, sizes):
        self.sizes = sizes
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        self.buf = torch.zeros(self.offsets[-1], device=device, dtype=torch.float32)

    def __getitem__(self, i):
        return self.buf[self.offsets[i] : self.offsets[i + 1]]

    def finish(self):
        out = ops.cast_f32_bf16(self.buf)
        return [out[self.offsets[i] : self.offsets[i + 1]] for i in range(len(self.sizes))]


def _wgrad(dy, x):
    """dW[out, in] = dY^T X with both operands read MN-major (no transposes through HBM)."""
    return ops.gemm(dy, x, a_mn=True, b_mn=True)


# ======================================================================================================================
# ViT residual attention block (pre-LN) — antmmf/modules/vision/backbone/clip/model.py:227-256
# ======================================================================================================================
def _vit_block_forward(x, p, B, L, H, eps, keep_ln=False):
    (ln1_w, ln1_b, in_w, in_b, out_w, out_b, ln2_w, ln2_b, fc_w, fc_b, proj_w, proj_b) = p
    W = x.shape[1]
    h1, _, mean1, rstd1 = ops.layernorm_fwd(x, ln1_w, ln1_b, eps)
    qkv = ops.gemm(h1, in_w, bias=in_b)
    if not keep_ln:
        del h1
        h1 = None
    o, lse = ops.attention_fwd(qkv, B, L, H, W // H)
    x_mid = ops.gemm(o, out_w, bias=out_b, residual=x)
    h2, _, mean2, rstd2 = ops.layernorm_fwd(x_mid, ln2_w, ln2_b, eps)
    g, u = ops.gemm(h2, fc_w, bias=fc_b, act=ACT_QUICKGELU, aux_out=True)
    if not keep_ln:
        del h2
        h2 = None
    y = ops.gemm(g, proj_w, bias=proj_b, residual=x_mid)
    return y, (mean1, rstd1, qkv, o, lse, x_mid, mean2, rstd2, u), g, (h1, h2)


class VitBlockFn(Function):
    """Memory-for-time knobs (both default off): `keep_act` keeps the activated MLP hidden g (8·T·W bytes per block; since the dgrad GEMM
    emits g next to du it only saves that epilogue's extra store, ≈ 0.08 ms per block at T = 263 168), `keep_ln` keeps both LayerNorm
    outputs (4·T·W bytes per block) and saves the two LN recomputes of the backward (≈ 0.37 ms per block): twice the time per byte."""

    @staticmethod
    def forward(ctx, x, ln1_w, ln1_b, in_w, in_b, out_w, out_b, ln2_w, ln2_b, fc_w, fc_b, proj_w, proj_b, B, L, H, eps, checkpoint,
                keep_act, keep_ln=False):
        p = (ln1_w, ln1_b, in_w, in_b, out_w, out_b, ln2_w, ln2_b, fc_w, fc_b, proj_w, proj_b)
        keep_act = bool(keep_act) and not checkpoint
        keep_ln = bool(keep_ln) and not checkpoint
        y, saved, g, hs = _vit_block_forward(x, p, B, L, H, eps, keep_ln)
        ctx.meta = (B, L, H, eps, checkpoint, keep_act, keep_ln)
        extra = ((g,) if keep_act else ()) + (hs if keep_ln else ())
        if checkpoint:
            ctx.save_for_backward(x, *p)
        else:
            ctx.save_for_backward(x, *p, *saved, *extra)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, L, H, eps, checkpoint, keep_act, keep_ln = ctx.meta
        t = ctx.saved_tensors
        x, p = t[0], t[1:13]
        (ln1_w, ln1_b, in_w, in_b, out_w, out_b, ln2_w, ln2_b, fc_w, fc_b, proj_w, proj_b) = p
        g = h1 = h2 = None
        if checkpoint:
            _, saved, g, _ = _vit_block_forward(x, p, B, L, H, eps)
        else:
            saved = t[13:22]
            k = 22
            if keep_act:
                g = t[k]
                k += 1
            if keep_ln:
                h1, h2 = t[k], t[k + 1]
        mean1, rstd1, qkv, o, lse, x_mid, mean2, rstd2, u = saved
        W = x.shape[1]
        dy = dy.contiguous()
        vg = _VecGrads(x.device, [W, W, 3 * W, W, W, W, 4 * W, W])  # ln1 w,b | in_b | out_b | ln2 w,b | fc_b | proj_b
        # ---- MLP branch
        if g is None:
            # the dgrad GEMM emits the activated hidden act(u) next to du (same sigmoid): no separate recompute pass over [T, 4W]
            du, g = ops.gemm(dy, proj_w, b_mn=True, act=ACT_QUICKGELU, dact_in=u, aux_out=True)
        else:
            du = ops.gemm(dy, proj_w, b_mn=True, act=ACT_QUICKGELU, dact_in=u)
        d_proj_w = _wgrad(dy, g)
        del g
        ops.rowsum_periodic(dy, vg[7])
        if h2 is None:
            h2, _, _, _ = ops.layernorm_fwd(x_mid, ln2_w, ln2_b, eps)
        d_fc_w = _wgrad(du, h2)
        del h2
        ops.rowsum_periodic(du, vg[6])
        dh2 = ops.gemm(du, fc_w, b_mn=True)
        del du
        dx_mid = ops.layernorm_bwd(dh2, x_mid, mean2, rstd2, ln2_w, vg[4], vg[5], dadd=dy)
        del dh2
        # ---- attention branch
        d_out_w = _wgrad(dx_mid, o)
        ops.rowsum_periodic(dx_mid, vg[3])
        d_o = ops.gemm(dx_mid, out_w, b_mn=True)
        dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, L, H, W // H)
        del d_o
        if h1 is None:
            h1, _, _, _ = ops.layernorm_fwd(x, ln1_w, ln1_b, eps)
        d_in_w = _wgrad(dqkv, h1)
        del h1
        ops.rowsum_periodic(dqkv, vg[2])
        dh1 = ops.gemm(dqkv, in_w, b_mn=True)
        del dqkv
        dx = ops.layernorm_bwd(dh1, x, mean1, rstd1, ln1_w, vg[0], vg[1], dadd=dx_mid)
        d_ln1_w, d_ln1_b, d_in_b, d_out_b, d_ln2_w, d_ln2_b, d_fc_b, d_proj_b = vg.finish()
        return (dx, d_ln1_w, d_ln1_b, d_in_w, d_in_b, d_out_w, d_out_b, d_ln2_w, d_ln2_b, d_fc_w, d_fc_b, d_proj_w, d_proj_b,
                None, None, None, None, None, None, None)


# ======================================================================================================================
# ViT stem: conv1 patch embedding + class token + positional embedding + ln_pre — clip/model.py:309-324
# ======================================================================================================================
class VitStemFn(Function):
    @staticmethod
    def forward(ctx, image, conv_w, cls, pos, ln_w, ln_b, eps):
        Bn, C, Hh, Ww = image.shape
        width, _, p, _ = conv_w.shape
        K = C * p * p
        Kp = (K + 63) // 64 * 64
        L = (Hh // p) * (Ww // p) + 1
        if pos.shape[0] != L:
            raise ops._lib.B200mmError(f"positional_embedding has {pos.shape[0]} rows, image gives {L} tokens")
        patches = ops.im2row(image, p, Kp)
        w2d = torch.zeros((width, Kp), device=image.device, dtype=BF16)
        w2d[:, :K] = conv_w.reshape(width, K)
        raw = ops.gemm(patches, w2d)
        del patches
        x0, s, mean, rstd = ops.layernorm_fwd(raw, ln_w, ln_b, eps, add0=pos, add1=cls, add_period=L, want_sum=True)
        ctx.save_for_backward(image, s, mean, rstd, ln_w)
        ctx.meta = (p, K, Kp, L, tuple(conv_w.shape))
        return x0

    @staticmethod
    def backward(ctx, dx0):
        image, s, mean, rstd, ln_w = ctx.saved_tensors
        p, K, Kp, L, wshape = ctx.meta
        W = s.shape[1]
        vg = _VecGrads(s.device, [W, W, L * W])
        ds = ops.layernorm_bwd(dx0.contiguous(), s, mean, rstd, ln_w, vg[0], vg[1])
        ops.rowsum_periodic(ds, vg[2].view(L, W), period=L)
        patches = ops.im2row(image, p, Kp)
        d_w2d = _wgrad(ds, patches)  # class-token rows of `patches` are zero, so they drop out
        d_conv = d_w2d[:, :K].reshape(wshape)
        d_ln_w, d_ln_b, d_pos = vg.finish()
        d_pos = d_pos.view(L, W)
        return None, d_conv, d_pos[0].clone(), d_pos, d_ln_w, d_ln_b, None


# ======================================================================================================================
# Class-token head: ln_post(x[:, 0]) @ proj — clip/model.py:330-333;  text: seq[:, 0] @ text_projection — cn_model.py:210
# ======================================================================================================================
class ClsHeadFn(Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, proj, B, L, eps):
        W = x.shape[1]
        xc = x.view(B, L, W)[:, 0, :].contiguous()
        if ln_w is not None:
            h, _, mean, rstd = ops.layernorm_fwd(xc, ln_w, ln_b, eps)
        else:
            h, mean, rstd = xc, None, None
        out = ops.gemm(h, proj, b_mn=True)  # proj is [W, E] = row-major [K, N]
        ctx.meta = (B, L, ln_w is not None)
        ctx.save_for_backward(xc, h, mean, rstd, ln_w, proj)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, L, has_ln = ctx.meta
        xc, h, mean, rstd, ln_w, proj = ctx.saved_tensors
        W = xc.shape[1]
        dout = dout.contiguous()
        dh = ops.gemm(dout, proj)  # [B, W] = dout [B, E] · proj[W, E]^T
        d_proj = ops.gemm(h, dout, a_mn=True, b_mn=True)  # [W, E] = h^T dout
        d_ln_w = d_ln_b = None
        if has_ln:
            vg = _VecGrads(xc.device, [W, W])
            dxc = ops.layernorm_bwd(dh, xc, mean, rstd, ln_w, vg[0], vg[1])
            d_ln_w, d_ln_b = vg.finish()
        else:
            dxc = dh
        dx = torch.zeros((B * L, W), device=xc.device, dtype=BF16)
        dx.view(B, L, W)[:, 0, :] = dxc
        return dx, d_ln_w, d_ln_b, d_proj, None, None, None


# ======================================================================================================================
# BERT layer (post-LN) — antmmf/modules/vision/backbone/clip/modeling_bert.py:253-270
# ======================================================================================================================
def _drop_sites(drop):
    """drop = None or (p_hidden, p_attn, seed_attn, seed_self_out, seed_out) -> the (p, seed) pairs of the three dropout sites of a BERT
    layer: attention probabilities (modeling_bert.py:158), BertSelfOutput (:180), BertOutput (:232); None where p = 0."""
    if drop is None:
        return None, None, None
    p_h, p_a, s_a, s_so, s_o = drop
    return ((p_a, s_a) if p_a > 0 else None), ((p_h, s_so) if p_h > 0 else None), ((p_h, s_o) if p_h > 0 else None)


def _bert_layer_forward(x, p, key_bias, B, L, H, eps, drop=None):
    (qkv_w, qkv_b, o_w, o_b, ln1_w, ln1_b, i_w, i_b, d_w, d_b, ln2_w, ln2_b) = p
    Hd = x.shape[1]
    d_attn, d_so, d_out = _drop_sites(drop)
    qkv = ops.gemm(x, qkv_w, bias=qkv_b)
    c, lse = ops.attention_fwd(qkv, B, L, H, Hd // H, key_bias=key_bias, drop=d_attn)
    s1 = ops.gemm(c, o_w, bias=o_b, residual=x, drop=d_so)  # dropout(dense(c)) + x in the epilogue
    x1, _, mean1, rstd1 = ops.layernorm_fwd(s1, ln1_w, ln1_b, eps)
    g, u = ops.gemm(x1, i_w, bias=i_b, act=ACT_GELU_ERF, aux_out=True)
    s2 = ops.gemm(g, d_w, bias=d_b, residual=x1, drop=d_out)
    del g, x1
    y, _, mean2, rstd2 = ops.layernorm_fwd(s2, ln2_w, ln2_b, eps)
    return y, (qkv, c, lse, s1, mean1, rstd1, u, s2, mean2, rstd2)


class BertLayerFn(Function):
    @staticmethod
    def forward(ctx, x, q_w, q_b, k_w, k_b, v_w, v_b, o_w, o_b, ln1_w, ln1_b, i_w, i_b, d_w, d_b, ln2_w, ln2_b, key_bias, B, L, H,
                eps, checkpoint, drop=None):
        qkv_w = torch.cat([q_w, k_w, v_w], dim=0)
        qkv_b = torch.cat([q_b, k_b, v_b], dim=0)
        p = (qkv_w, qkv_b, o_w, o_b, ln1_w, ln1_b, i_w, i_b, d_w, d_b, ln2_w, ln2_b)
        y, saved = _bert_layer_forward(x, p, key_bias, B, L, H, eps, drop)
        ctx.meta = (B, L, H, eps, checkpoint, key_bias is not None)
        ctx.drop = drop  # probabilities + host-side seeds: the masks are regenerated, never stored
        kb = (key_bias,) if key_bias is not None else ()
        if checkpoint:
            ctx.save_for_backward(x, *p, *kb)
        else:
            ctx.save_for_backward(x, *p, *kb, *saved)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, L, H, eps, checkpoint, has_kb = ctx.meta
        t = ctx.saved_tensors
        x, p = t[0], t[1:13]
        key_bias = t[13] if has_kb else None
        rest = t[13 + int(has_kb) :]
        (qkv_w, qkv_b, o_w, o_b, ln1_w, ln1_b, i_w, i_b, d_w, d_b, ln2_w, ln2_b) = p
        d_attn, d_so, d_out = _drop_sites(ctx.drop)
        if checkpoint:
            _, rest = _bert_layer_forward(x, p, key_bias, B, L, H, eps, ctx.drop)
        qkv, c, lse, s1, mean1, rstd1, u, s2, mean2, rstd2 = rest
        Hd = x.shape[1]
        I = i_w.shape[0]
        vg = _VecGrads(x.device, [3 * Hd, Hd, Hd, Hd, I, Hd, Hd, Hd])  # qkv_b | o_b | ln1 w,b | i_b | d_b | ln2 w,b
        ds2 = ops.layernorm_bwd(dy.contiguous(), s2, mean2, rstd2, ln2_w, vg[6], vg[7])
        # gradient w.r.t. the dense output = the sum's gradient through the SAME mask; the residual branch keeps the unmasked ds2
        dm2 = ops.dropout(ds2, *d_out) if d_out is not None else ds2
        du, g = ops.gemm(dm2, d_w, b_mn=True, act=ACT_GELU_ERF, dact_in=u, aux_out=True)  # g = gelu(u) recomputed in the epilogue
        d_d_w = _wgrad(dm2, g)
        del g
        ops.rowsum_periodic(dm2, vg[5])
        del dm2
        x1, _, _, _ = ops.layernorm_fwd(s1, ln1_w, ln1_b, eps)
        d_i_w = _wgrad(du, x1)
        del x1
        ops.rowsum_periodic(du, vg[4])
        dx1 = ops.gemm(du, i_w, b_mn=True, residual=ds2)  # + ds2: x1 also feeds the output residual
        del du, ds2
        ds1 = ops.layernorm_bwd(dx1, s1, mean1, rstd1, ln1_w, vg[2], vg[3])
        del dx1
        dm1 = ops.dropout(ds1, *d_so) if d_so is not None else ds1
        d_o_w = _wgrad(dm1, c)
        ops.rowsum_periodic(dm1, vg[1])
        dc = ops.gemm(dm1, o_w, b_mn=True)
        del dm1
        dqkv = ops.attention_bwd(qkv, c, dc, lse, B, L, H, Hd // H, key_bias=key_bias, drop=d_attn)
        del dc
        d_qkv_w = _wgrad(dqkv, x)
        ops.rowsum_periodic(dqkv, vg[0])
        dx = ops.gemm(dqkv, qkv_w, b_mn=True, residual=ds1)  # + ds1: x also feeds the attention residual
        d_qkv_b, d_o_b, d_ln1_w, d_ln1_b, d_i_b, d_d_b, d_ln2_w, d_ln2_b = vg.finish()
        dq_w, dk_w, dv_w = d_qkv_w[:Hd], d_qkv_w[Hd : 2 * Hd], d_qkv_w[2 * Hd :]
        dq_b, dk_b, dv_b = d_qkv_b[:Hd], d_qkv_b[Hd : 2 * Hd], d_qkv_b[2 * Hd :]
        return (dx, dq_w, dq_b, dk_w, dk_b, dv_w, dv_b, d_o_w, d_o_b, d_ln1_w, d_ln1_b, d_i_w, d_i_b, d_d_w, d_d_b, d_ln2_w, d_ln2_b,
                None, None, None, None, None, None, None)


# ======================================================================================================================
# BERT embeddings — modeling_bert.py:66-103 (inputs_embeds variant: prj/base_vtp/.../clip_text_encoder.py:36-60)
# ======================================================================================================================
class BertEmbeddingsFn(Function):
    @staticmethod
    def forward(ctx, word_or_embeds, ids, pos_table, type_table, type_ids, ln_w, ln_b, L, eps, padding_idx, from_embeds, drop=None):
        # from_embeds: word_or_embeds is [rows, H] (already embedded tokens), ids = arange(rows)
        # drop = (p, seed): BertEmbeddings.dropout on the LayerNorm output (modeling_bert.py:101), in place
        y, s, mean, rstd = ops.embed_layernorm_fwd(word_or_embeds, ids, pos_table, L, type_table, type_ids, ln_w, ln_b, eps)
        if drop is not None and drop[0] > 0:
            ops.dropout(y, drop[0], drop[1], out=y)
        else:
            drop = None
        ctx.drop = drop
        ctx.save_for_backward(s, mean, rstd, ln_w, ids, type_ids)
        ctx.meta = (L, padding_idx, from_embeds, word_or_embeds.shape[0], pos_table.shape[0], type_table.shape[0])
        return y

    @staticmethod
    def backward(ctx, dy):
        s, mean, rstd, ln_w, ids, type_ids = ctx.saved_tensors
        L, padding_idx, from_embeds, n_word, n_pos, n_type = ctx.meta
        Hd = s.shape[1]
        vg = _VecGrads(s.device, [Hd, Hd, n_pos * Hd, n_type * Hd])
        dy = dy.contiguous()
        if ctx.drop is not None:
            dy = ops.dropout(dy, ctx.drop[0], ctx.drop[1])
        ds = ops.layernorm_bwd(dy, s, mean, rstd, ln_w, vg[0], vg[1])
        ops.rowsum_periodic(ds, vg[2].view(n_pos, Hd)[:L], period=L)
        ops.scatter_add_rows(ds, type_ids, vg[3].view(n_type, Hd))
        if from_embeds:
            d_word = ds
        else:
            acc = torch.zeros((n_word, Hd), device=s.device, dtype=torch.float32)
            ops.scatter_add_rows(ds, ids, acc, skip_id=padding_idx if padding_idx is not None else -1)
            d_word = ops.cast_f32_bf16(acc)
        d_ln_w, d_ln_b, d_pos, d_type = vg.finish()
        return d_word, None, d_pos.view(n_pos, Hd), d_type.view(n_type, Hd), None, d_ln_w, d_ln_b, None, None, None, None, None


# ======================================================================================================================
# L2 row normalisation — cn_model.py:217-218
# ======================================================================================================================
class RowNormFn(Function):
    @staticmethod
    def forward(ctx, x):
        y, inv = ops.rownorm_fwd(x)
        ctx.save_for_backward(x, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, inv = ctx.saved_tensors
        return ops.rownorm_bwd(dy.float().contiguous(), x, inv)


# ======================================================================================================================
# M²-Encoder (BEiT-3 multiway) encoder layer: pre-LN + sub-LN — prj/M2_Encoder/vlmo/torchscale/architecture/encoder.py:113-168,
# attention component/multihead_attention.py:66-154, FFN component/feedforward_network.py:117-128
# ======================================================================================================================
def _m2_layer_forward(x, p, key_bias, B, L, H, eps, xpos=None):
    (ln1_w, ln1_b, qkv_w, qkv_b, iln_w, iln_b, o_w, o_b, ln2_w, ln2_b, fc1_w, fc1_b, fln_w, fln_b, fc2_w, fc2_b) = p
    W = x.shape[1]
    h1, _, mean1, rstd1 = ops.layernorm_fwd(x, ln1_w, ln1_b, eps)
    qkv = ops.gemm(h1, qkv_w, bias=qkv_b)
    del h1
    if xpos is not None:  # optional rotary embedding (XPOS, multihead_attention.py:112-118), in place on the q / k sections
        ops.xpos_apply(qkv, xpos, B, L, H, W // H)
    # the reference scales q by hd^-0.5 before q·k^T (multihead_attention.py:95); the kernel applies the same factor to the scores
    a, lse = ops.attention_fwd(qkv, B, L, H, W // H, key_bias=key_bias)
    a_n, _, mean_i, rstd_i = ops.layernorm_fwd(a, iln_w, iln_b, eps)  # inner_attn_ln on the merged heads (:148-149)
    x_mid = ops.gemm(a_n, o_w, bias=o_b, residual=x)
    del a_n
    h2, _, mean2, rstd2 = ops.layernorm_fwd(x_mid, ln2_w, ln2_b, eps)
    u = ops.gemm(h2, fc1_w, bias=fc1_b)
    del h2
    g_n, mean_f, rstd_f = ops.act_layernorm_fwd(u, ACT_GELU_ERF, fln_w, fln_b, eps)  # ffn_layernorm(gelu(u)) in one pass
    y = ops.gemm(g_n, fc2_w, bias=fc2_b, residual=x_mid)
    return y, (mean1, rstd1, qkv, a, lse, mean_i, rstd_i, x_mid, mean2, rstd2, u, mean_f, rstd_f), g_n


class M2EncoderLayerFn(Function):
    """One multiway expert (A or B) of an EncoderLayer applied to every token (split_position −1 / 0)."""

    @staticmethod
    def forward(ctx, x, ln1_w, ln1_b, q_w, q_b, k_w, k_b, v_w, v_b, iln_w, iln_b, o_w, o_b, ln2_w, ln2_b, fc1_w, fc1_b, fln_w, fln_b,
                fc2_w, fc2_b, key_bias, B, L, H, eps, checkpoint, keep_act, xpos=None):
        qkv_w = torch.cat([q_w, k_w, v_w], dim=0)
        qkv_b = torch.cat([q_b, k_b, v_b], dim=0)
        p = (ln1_w, ln1_b, qkv_w, qkv_b, iln_w, iln_b, o_w, o_b, ln2_w, ln2_b, fc1_w, fc1_b, fln_w, fln_b, fc2_w, fc2_b)
        y, saved, g_n = _m2_layer_forward(x, p, key_bias, B, L, H, eps, xpos)
        keep_act = bool(keep_act) and not checkpoint
        ctx.meta = (B, L, H, eps, checkpoint, key_bias is not None, keep_act)
        ctx.xpos = xpos  # constant f32 tables (no gradient), kept by reference
        kb = (key_bias,) if key_bias is not None else ()
        if checkpoint:
            ctx.save_for_backward(x, *p, *kb)
        elif keep_act:  # memory for time: the normalised FFN hidden (4W per token) is kept instead of re-running gelu + sub-LN
            ctx.save_for_backward(x, *p, *kb, *saved, g_n)
        else:
            ctx.save_for_backward(x, *p, *kb, *saved)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, L, H, eps, checkpoint, has_kb, keep_act = ctx.meta
        t = ctx.saved_tensors
        x, p = t[0], t[1:17]
        key_bias = t[17] if has_kb else None
        rest = t[17 + int(has_kb):]
        (ln1_w, ln1_b, qkv_w, qkv_b, iln_w, iln_b, o_w, o_b, ln2_w, ln2_b, fc1_w, fc1_b, fln_w, fln_b, fc2_w, fc2_b) = p
        g_n = None
        if checkpoint:
            _, rest, g_n = _m2_layer_forward(x, p, key_bias, B, L, H, eps, ctx.xpos)
        elif keep_act:
            rest, g_n = rest[:-1], rest[-1]
        mean1, rstd1, qkv, a, lse, mean_i, rstd_i, x_mid, mean2, rstd2, u, mean_f, rstd_f = rest
        W = x.shape[1]
        F_ = fc1_w.shape[0]
        dy = dy.contiguous()
        # ln1 w,b | qkv_b | iln w,b | o_b | ln2 w,b | fc1_b | fln w,b | fc2_b
        vg = _VecGrads(x.device, [W, W, 3 * W, W, W, W, W, W, F_, F_, F_, W])
        # ---- FFN branch: fc2(ffn_ln(gelu(fc1 h2)))
        if g_n is None:
            g_n, _, _ = ops.act_layernorm_fwd(u, ACT_GELU_ERF, fln_w, fln_b, eps)  # recomputed unless the keep-activation policy kept it
        d_fc2_w = _wgrad(dy, g_n)
        del g_n
        ops.rowsum_periodic(dy, vg[11])
        dg_n = ops.gemm(dy, fc2_w, b_mn=True)
        du = ops.act_layernorm_bwd(dg_n, u, ACT_GELU_ERF, mean_f, rstd_f, fln_w, vg[9], vg[10])  # through sub-LN and gelu' at once
        del dg_n
        h2, _, _, _ = ops.layernorm_fwd(x_mid, ln2_w, ln2_b, eps)
        d_fc1_w = _wgrad(du, h2)
        del h2
        ops.rowsum_periodic(du, vg[8])
        dh2 = ops.gemm(du, fc1_w, b_mn=True)
        del du
        dx_mid = ops.layernorm_bwd(dh2, x_mid, mean2, rstd2, ln2_w, vg[6], vg[7], dadd=dy)
        del dh2
        # ---- attention branch: out_proj(inner_ln(attn(qkv(ln1 x))))
        a_n, _, _, _ = ops.layernorm_fwd(a, iln_w, iln_b, eps)
        d_o_w = _wgrad(dx_mid, a_n)
        del a_n
        ops.rowsum_periodic(dx_mid, vg[5])
        da_n = ops.gemm(dx_mid, o_w, b_mn=True)
        da = ops.layernorm_bwd(da_n, a, mean_i, rstd_i, iln_w, vg[3], vg[4])
        del da_n
        dqkv = ops.attention_bwd(qkv, a, da, lse, B, L, H, W // H, key_bias=key_bias)
        del da
        if ctx.xpos is not None:  # gradient w.r.t. the rotated q / k -> w.r.t. the projection output (transposed rotation)
            ops.xpos_apply(dqkv, ctx.xpos, B, L, H, W // H, backward=True)
        h1, _, _, _ = ops.layernorm_fwd(x, ln1_w, ln1_b, eps)
        d_qkv_w = _wgrad(dqkv, h1)
        del h1
        ops.rowsum_periodic(dqkv, vg[2])
        dh1 = ops.gemm(dqkv, qkv_w, b_mn=True)
        del dqkv
        dx = ops.layernorm_bwd(dh1, x, mean1, rstd1, ln1_w, vg[0], vg[1], dadd=dx_mid)
        d_ln1_w, d_ln1_b, d_qkv_b, d_iln_w, d_iln_b, d_o_b, d_ln2_w, d_ln2_b, d_fc1_b, d_fln_w, d_fln_b, d_fc2_b = vg.finish()
        dq_w, dk_w, dv_w = d_qkv_w[:W], d_qkv_w[W: 2 * W], d_qkv_w[2 * W:]
        dq_b, dk_b, dv_b = d_qkv_b[:W], d_qkv_b[W: 2 * W], d_qkv_b[2 * W:]
        return (dx, d_ln1_w, d_ln1_b, dq_w, dq_b, dk_w, dk_b, dv_w, dv_b, d_iln_w, d_iln_b, d_o_w, d_o_b, d_ln2_w, d_ln2_b, d_fc1_w, d_fc1_b,
                d_fln_w, d_fln_b, d_fc2_w, d_fc2_b, None, None, None, None, None, None, None, None)


class LayerNormFn(Function):
    """Plain LayerNorm over all rows (Encoder.layer_norm, architecture/encoder.py:469-470)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, _, mean, rstd = ops.layernorm_fwd(x, w, b, eps)
        ctx.save_for_backward(x, mean, rstd, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, w = ctx.saved_tensors
        W = x.shape[1]
        vg = _VecGrads(x.device, [W, W])
        dx = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, w, vg[0], vg[1])
        dw, db = vg.finish()
        return dx, dw, db, None


class MaskRowsFn(Function):
    """x * (1 − padding) on token rows (Encoder.forward, architecture/encoder.py:440); the gradient is masked the same way."""

    @staticmethod
    def forward(ctx, x, drop):
        ctx.save_for_backward(drop)
        return ops.mask_rows(x, drop)

    @staticmethod
    def backward(ctx, dy):
        (drop,) = ctx.saved_tensors
        return ops.mask_rows(dy.contiguous(), drop), None


class ClsLinearFn(Function):
    """ITCHead on the CLS rows: x[:, 0] @ weight^T, weight [E, W] bias-free (vlmo/modules/heads.py:17-24, vlmo_module.py:346,389)."""

    @staticmethod
    def forward(ctx, x, weight, B, L):
        W = x.shape[1]
        xc = x.view(B, L, W)[:, 0, :].contiguous()
        ctx.save_for_backward(xc, weight)
        ctx.meta = (B, L)
        return ops.gemm(xc, weight)

    @staticmethod
    def backward(ctx, dout):
        xc, weight = ctx.saved_tensors
        B, L = ctx.meta
        dout = dout.contiguous()
        dxc = ops.gemm(dout, weight, b_mn=True)  # [B, W] = dout [B, E] · weight [E, W]
        d_w = ops.gemm(dout, xc, a_mn=True, b_mn=True)  # [E, W] = dout^T xc
        dx = torch.zeros((B * L, xc.shape[1]), device=xc.device, dtype=BF16)
        dx.view(B, L, -1)[:, 0, :] = dxc
        return dx, d_w, None, None


class M2VisionEmbedFn(Function):
    """VisionEmbedding (conv patch-embed with bias, CLS prepended; component/embedding.py:67-83) + positional expert A from
    position 2 (embedding.py:93-110). `pos` is the [L, W] slice embed_positions.A.weight[2:L+2]."""

    @staticmethod
    def forward(ctx, image, conv_w, conv_b, cls, pos):
        Bn, C, Hh, Ww = image.shape
        width, _, p, _ = conv_w.shape
        K = C * p * p
        Kp = (K + 63) // 64 * 64
        L = (Hh // p) * (Ww // p) + 1
        if pos.shape[0] != L:
            raise ops._lib.B200mmError(f"embed_positions.A gives {pos.shape[0]} rows, image gives {L} tokens")
        patches = ops.im2row(image, p, Kp)  # one zero row per image for the CLS slot
        w2d = torch.zeros((width, Kp), device=image.device, dtype=BF16)
        w2d[:, :K] = conv_w.reshape(width, K)
        raw = ops.gemm(patches, w2d)
        del patches
        add0 = pos.clone()
        add0[1:] += conv_b  # the conv bias belongs to the patch rows only ([L, W] parameter-sized op)
        one = torch.ones(width, device=image.device, dtype=BF16)
        zero = torch.zeros(width, device=image.device, dtype=BF16)
        _, s, _, _ = ops.layernorm_fwd(raw, one, zero, 1e-5, add0=add0, add1=cls.reshape(-1), add_period=L, want_sum=True)
        ctx.save_for_backward(image)
        ctx.meta = (p, K, Kp, L, tuple(conv_w.shape), tuple(cls.shape))
        return s

    @staticmethod
    def backward(ctx, ds):
        (image,) = ctx.saved_tensors
        p, K, Kp, L, wshape, cshape = ctx.meta
        ds = ds.contiguous()
        W = ds.shape[1]
        vg = _VecGrads(ds.device, [L * W])
        ops.rowsum_periodic(ds, vg[0].view(L, W), period=L)
        patches = ops.im2row(image, p, Kp)
        d_w2d = _wgrad(ds, patches)
        d_conv = d_w2d[:, :K].reshape(wshape)
        dadd32 = vg[0].view(L, W)
        d_bias = dadd32[1:].sum(0).to(BF16)  # [L, W] parameter-sized reduction
        (d_add,) = vg.finish()
        d_add = d_add.view(L, W)
        return None, d_conv, d_bias, d_add[0].reshape(cshape).clone(), d_add


class M2TextEmbedFn(Function):
    """TextEmbedding lookup + positional expert B from position 2, padded rows zeroed (BEiT3.forward model/BEiT3.py:64-67,
    Encoder.forward_embedding architecture/encoder.py:353-363 and :440). `pos` = embed_positions.B.weight[2:L+2]."""

    @staticmethod
    def forward(ctx, word, ids, pos, drop):
        Bn, L = ids.shape
        W = word.shape[1]
        flat = ids.reshape(-1).contiguous()
        ztype = torch.zeros((1, W), device=word.device, dtype=BF16)
        zids = torch.zeros_like(flat)
        one = torch.ones(W, device=word.device, dtype=BF16)
        zero = torch.zeros(W, device=word.device, dtype=BF16)
        _, s, _, _ = ops.embed_layernorm_fwd(word, flat, pos.contiguous(), L, ztype, zids, one, zero, 1e-5)
        if drop is not None:
            s = ops.mask_rows(s, drop, inplace=True)
        ctx.save_for_backward(flat, drop)
        ctx.meta = (L, word.shape[0])
        return s

    @staticmethod
    def backward(ctx, ds):
        flat, drop = ctx.saved_tensors
        L, n_word = ctx.meta
        ds = ds.contiguous()
        if drop is not None:
            ds = ops.mask_rows(ds, drop)
        W = ds.shape[1]
        vg = _VecGrads(ds.device, [L * W])
        ops.rowsum_periodic(ds, vg[0].view(L, W), period=L)
        acc = torch.zeros((n_word, W), device=ds.device, dtype=torch.float32)
        ops.scatter_add_rows(ds, flat, acc)
        d_word = ops.cast_f32_bf16(acc)
        (d_pos,) = vg.finish()
        return d_word, None, d_pos.view(L, W), None


# ======================================================================================================================
# M²-Encoder layer in three pieces, for a multiway split INSIDE a sequence (fused vision + language input, split_position > 0:
# multiway_network.py:38-45): the per-token parts run per expert on that expert's rows, attention runs on the joint sequence.
# ======================================================================================================================
class M2PreAttnFn(Function):
    """x -> qkv = Linear_qkv(LN(x))  (self_attn_layer_norm + q/k/v projections of ONE expert, encoder.py:138-139, multihead_attention.py:91-93)."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, q_w, q_b, k_w, k_b, v_w, v_b, eps):
        qkv_w = torch.cat([q_w, k_w, v_w], dim=0)
        qkv_b = torch.cat([q_b, k_b, v_b], dim=0)
        h, _, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, eps)
        qkv = ops.gemm(h, qkv_w, bias=qkv_b)
        ctx.save_for_backward(x, mean, rstd, ln_w, ln_b, qkv_w)
        ctx.eps = eps
        return qkv

    @staticmethod
    def backward(ctx, dqkv):
        x, mean, rstd, ln_w, ln_b, qkv_w = ctx.saved_tensors
        W = x.shape[1]
        dqkv = dqkv.contiguous()
        vg = _VecGrads(x.device, [W, W, 3 * W])
        h, _, _, _ = ops.layernorm_fwd(x, ln_w, ln_b, ctx.eps)
        d_w = _wgrad(dqkv, h)
        del h
        ops.rowsum_periodic(dqkv, vg[2])
        dh = ops.gemm(dqkv, qkv_w, b_mn=True)
        dx = ops.layernorm_bwd(dh, x, mean, rstd, ln_w, vg[0], vg[1])
        d_ln_w, d_ln_b, d_b = vg.finish()
        return (dx, d_ln_w, d_ln_b, d_w[:W], d_b[:W], d_w[W: 2 * W], d_b[W: 2 * W], d_w[2 * W:], d_b[2 * W:], None)


class AttentionFn(Function):
    """softmax(q k^T / sqrt(hd) + key_bias) v on a fused projection buffer qkv [B*L, 3W] -> [B*L, W]."""

    @staticmethod
    def forward(ctx, qkv, key_bias, B, L, H):
        W = qkv.shape[1] // 3
        o, lse = ops.attention_fwd(qkv, B, L, H, W // H, key_bias=key_bias)
        ctx.save_for_backward(qkv, o, lse, key_bias)
        ctx.meta = (B, L, H, W)
        return o

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, lse, key_bias = ctx.saved_tensors
        B, L, H, W = ctx.meta
        return ops.attention_bwd(qkv, o, d_o.contiguous(), lse, B, L, H, W // H, key_bias=key_bias), None, None, None, None


class XposFn(Function):
    """Out-of-place XPOS on the q / k sections of qkv (see M2EncoderLayerFn for the in-place use inside the fused layer)."""

    @staticmethod
    def forward(ctx, qkv, tables, B, L, H):
        ctx.tables, ctx.meta = tables, (B, L, H, qkv.shape[1] // 3 // H)
        return ops.xpos_apply(qkv.clone(), tables, B, L, H, ctx.meta[3])

    @staticmethod
    def backward(ctx, g):
        B, L, H, hd = ctx.meta
        return ops.xpos_apply(g.clone().contiguous(), ctx.tables, B, L, H, hd, backward=True), None, None, None, None


class M2PostAttnFn(Function):
    """(a, x) -> y: inner_attn_ln, out_proj + residual, final_layer_norm, fc1, gelu + ffn_layernorm, fc2 + residual of ONE expert
    (multihead_attention.py:148-151, encoder.py:149-167, feedforward_network.py:117-128)."""

    @staticmethod
    def forward(ctx, a, x, iln_w, iln_b, o_w, o_b, ln2_w, ln2_b, fc1_w, fc1_b, fln_w, fln_b, fc2_w, fc2_b, eps):
        a_n, _, mean_i, rstd_i = ops.layernorm_fwd(a, iln_w, iln_b, eps)
        x_mid = ops.gemm(a_n, o_w, bias=o_b, residual=x)
        del a_n
        h2, _, mean2, rstd2 = ops.layernorm_fwd(x_mid, ln2_w, ln2_b, eps)
        u = ops.gemm(h2, fc1_w, bias=fc1_b)
        del h2
        g_n, mean_f, rstd_f = ops.act_layernorm_fwd(u, ACT_GELU_ERF, fln_w, fln_b, eps)
        y = ops.gemm(g_n, fc2_w, bias=fc2_b, residual=x_mid)
        ctx.save_for_backward(a, iln_w, iln_b, o_w, ln2_w, ln2_b, fc1_w, fln_w, fln_b, fc2_w, mean_i, rstd_i, x_mid, mean2, rstd2, u, mean_f, rstd_f)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        (a, iln_w, iln_b, o_w, ln2_w, ln2_b, fc1_w, fln_w, fln_b, fc2_w, mean_i, rstd_i, x_mid, mean2, rstd2, u, mean_f, rstd_f) = ctx.saved_tensors
        eps = ctx.eps
        W, F_ = a.shape[1], fc1_w.shape[0]
        dy = dy.contiguous()
        vg = _VecGrads(a.device, [W, W, W, W, W, F_, F_, F_, W])  # iln w,b | o_b | ln2 w,b | fc1_b | fln w,b | fc2_b
        g_n, _, _ = ops.act_layernorm_fwd(u, ACT_GELU_ERF, fln_w, fln_b, eps)
        d_fc2_w = _wgrad(dy, g_n)
        del g_n
        ops.rowsum_periodic(dy, vg[8])
        dg_n = ops.gemm(dy, fc2_w, b_mn=True)
        du = ops.act_layernorm_bwd(dg_n, u, ACT_GELU_ERF, mean_f, rstd_f, fln_w, vg[6], vg[7])
        del dg_n
        h2, _, _, _ = ops.layernorm_fwd(x_mid, ln2_w, ln2_b, eps)
        d_fc1_w = _wgrad(du, h2)
        del h2
        ops.rowsum_periodic(du, vg[5])
        dh2 = ops.gemm(du, fc1_w, b_mn=True)
        del du
        dx_mid = ops.layernorm_bwd(dh2, x_mid, mean2, rstd2, ln2_w, vg[3], vg[4], dadd=dy)
        del dh2
        a_n, _, _, _ = ops.layernorm_fwd(a, iln_w, iln_b, eps)
        d_o_w = _wgrad(dx_mid, a_n)
        del a_n
        ops.rowsum_periodic(dx_mid, vg[2])
        da_n = ops.gemm(dx_mid, o_w, b_mn=True)
        da = ops.layernorm_bwd(da_n, a, mean_i, rstd_i, iln_w, vg[0], vg[1])
        d_iln_w, d_iln_b, d_o_b, d_ln2_w, d_ln2_b, d_fc1_b, d_fln_w, d_fln_b, d_fc2_b = vg.finish()
        return (da, dx_mid, d_iln_w, d_iln_b, d_o_w, d_o_b, d_ln2_w, d_ln2_b, d_fc1_w, d_fc1_b, d_fln_w, d_fln_b, d_fc2_w, d_fc2_b, None)


class GatherRowsFn(Function):
    """out[r] = table[ids[r]]; backward sums the row gradients back per table row (interleaving / splitting token rows of two experts)."""

    @staticmethod
    def forward(ctx, table, ids):
        ctx.save_for_backward(ids)
        ctx.n = table.shape[0]
        return ops.gather_rows(table.contiguous(), ids)

    @staticmethod
    def backward(ctx, dy):
        (ids,) = ctx.saved_tensors
        acc = torch.zeros((ctx.n, dy.shape[1]), device=dy.device, dtype=torch.float32)
        ops.scatter_add_rows(dy.contiguous(), ids, acc)
        return ops.cast_f32_bf16(acc), None
