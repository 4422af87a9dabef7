"""Per-kernel parity (GPU): every C-ABI entry point against the CPU oracle's arithmetic (oracle/restated.py) or a plain
fp32 torch restatement of the same op, on the SAME bf16-rounded inputs.

Tolerances: kernels accumulate in fp32, so with f32 outputs the bar is 2e-5 relative to the tensor scale; bf16 outputs
add one rounding (<= 2^-8 relative per element).
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import restated

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    import b200mm

    assert b200mm._lib.load().b200mm_check_device() == 0, b200mm._lib.load().b200mm_last_error()
    return b200mm.ops


def rnd(*shape, scale=1.0, seed=None):
    g = torch.Generator().manual_seed(seed if seed is not None else sum(shape) + len(shape))
    return (torch.randn(*shape, generator=g) * scale).to(BF)


def close(got, ref, tol, what=""):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = ref.abs().max().clamp_min(1e-6)
    err = (got - ref).abs().max() / scale
    assert torch.isfinite(got).all(), what
    assert float(err) <= tol, f"{what}: max err / scale = {float(err):.3e} > {tol}"


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (200, 264, 136), (77, 768, 768), (1000, 1032, 520)])
def test_gemm_layouts(ops, a_mn, b_mn, M, N, K):
    # MN-major operands keep the mn index contiguous: the row pitch must be a multiple of 8 elements (16 B), so an
    # odd M/N lives in a padded buffer
    M8, N8 = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    a_full = rnd(*((K, M8) if a_mn else (M, K))).cuda()
    b_full = rnd(*((K, N8) if b_mn else (N, K))).cuda()
    a = a_full[:, :M] if a_mn else a_full
    b = b_full[:, :N] if b_mn else b_full
    A = a.float().cpu().t() if a_mn else a.float().cpu()
    Bm = b.float().cpu() if b_mn else b.float().cpu().t()
    ref = A @ Bm
    if N % 8:
        with pytest.raises(Exception):
            ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out_f32=True)
        return
    got = ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out_f32=True)
    close(got, ref, 2e-5, "gemm f32")
    got16 = ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn))
    close(got16, ref, 5e-3, "gemm bf16")


@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_epilogue(ops, act):
    M, N, K = 300, 520, 200
    a, w = rnd(M, K), rnd(N, K, scale=0.1)
    bias, res = rnd(N), rnd(M, N)
    pre = 0.5 * (a.float() @ w.float().t()) + bias.float()
    actf = {0: lambda x: x, 1: restated.quick_gelu, 2: restated.gelu_erf}[act]
    ref = actf(pre) + res.float()
    got, aux = ops.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), act=act, aux_out=True, residual=res.cuda(), alpha=0.5, out_f32=True)
    close(got, ref, 1e-4, "epilogue out")  # erff / __expf differ from the CPU libm by ~1e-6 abs
    close(aux, pre, 5e-3, "aux_out")
    # derivative epilogue: D = (A·W^T) * act'(u)
    u = rnd(M, N)
    uf = u.float().requires_grad_()
    actf(uf).sum().backward()
    ref_d = (a.float() @ w.float().t()) * uf.grad
    got_d = ops.gemm(a.cuda(), w.cuda(), act=act, dact_in=u.cuda(), out_f32=True)
    close(got_d, ref_d, 3e-5, "dact")
    # ... and with aux_out the same launch also returns the recomputed activation act(u) (bf16)
    got_d2, g = ops.gemm(a.cuda(), w.cuda(), act=act, dact_in=u.cuda(), aux_out=True, out_f32=True)
    assert torch.equal(got_d2, got_d)
    close(g, actf(u.float()).detach(), 5e-3, "act(u) from the dact epilogue")
    # the CTA-pair flavoured kernels (M > 128, B operand MN-major as in the dgrad of the MLP) and the bf16 output
    wt = w.t().contiguous()
    got_d3, g3 = ops.gemm(a.cuda(), wt.cuda(), b_mn=True, act=act, dact_in=u.cuda(), aux_out=True)
    close(got_d3, ref_d, 8e-3, "dact bf16, MN-major B")
    close(g3, actf(u.float()).detach(), 5e-3, "act(u), MN-major B")


@pytest.mark.parametrize("splits", [2, 5])
def test_gemm_splitk(ops, splits):
    T, O, I = 4096 + 40, 200, 136  # wgrad shape: dW[O, I] = dY^T X
    dy, x, bias = rnd(T, O), rnd(T, I), rnd(I)
    ref = dy.float().t() @ x.float() + bias.float()
    got = ops.gemm(dy.cuda(), x.cuda(), a_mn=True, b_mn=True, bias=bias.cuda(), splits=splits, out_f32=True)
    close(got, ref, 2e-5, "split-K")


def test_gemm_strided_rows(ops):
    # A taken with a row pitch (class-token rows of a [B, L, W] buffer)
    B, L, W, E = 9, 5, 64, 48
    x = rnd(B * L, W).cuda()
    w = rnd(E, W).cuda()
    a = x.view(B, L * W)[:, :W]
    got = ops.gemm(a, w, out_f32=True)
    close(got, x.view(B, L, W)[:, 0].float().cpu() @ w.float().cpu().t(), 2e-5)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("W", [64, 768, 1024, 1280])
def test_layernorm_fwd_bwd(ops, W):
    rows, eps = 37 * 5, 1e-5
    x, w, b, dy, dadd = rnd(rows, W), rnd(W).add(1), rnd(W), rnd(rows, W), rnd(rows, W)
    xf, wf, bf_ = x.float().requires_grad_(), w.float().requires_grad_(), b.float().requires_grad_()
    ref = restated.layer_norm(xf, wf, bf_, eps)
    ref.backward(dy.float())
    y, s, mean, rstd = ops.layernorm_fwd(x.cuda(), w.cuda(), b.cuda(), eps)
    close(y, ref, 8e-3, "ln fwd")
    dw = torch.zeros(W, device="cuda")
    db = torch.zeros(W, device="cuda")
    dx = ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, w.cuda(), dw, db, dadd=dadd.cuda())
    close(dx, xf.grad + dadd.float(), 8e-3, "ln dx")
    close(dw, wf.grad, 1e-4, "ln dw")
    close(db, bf_.grad, 1e-4, "ln db")


def test_layernorm_stem_adds(ops):
    B, L, W = 6, 5, 64
    x, pos, cls, w, b = rnd(B * L, W), rnd(L, W), rnd(W), rnd(W).add(1), rnd(W)
    s_ref = x.float().view(B, L, W) + pos.float()
    s_ref[:, 0] += cls.float()
    y, s, _, _ = ops.layernorm_fwd(x.cuda(), w.cuda(), b.cuda(), 1e-5, add0=pos.cuda(), add1=cls.cuda(), add_period=L, want_sum=True)
    close(s, s_ref.view(B * L, W), 8e-3, "stem sum")
    close(y, restated.layer_norm(s.float().cpu(), w.float(), b.float(), 1e-5), 8e-3, "stem ln")


def test_embed_layernorm_bit_exact_indexing(ops):
    V, Lmax, H, B, L = 300, 32, 64, 7, 11
    word, pos, typ, w, b = rnd(V, H), rnd(Lmax, H), rnd(2, H), rnd(H).add(1), rnd(H)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, V, (B, L), generator=g)
    tt = torch.randint(0, 2, (B, L), generator=g)
    y, s, mean, rstd = ops.embed_layernorm_fwd(word.cuda(), ids.cuda(), pos.cuda(), L, typ.cuda(), tt.cuda(), w.cuda(), b.cuda(), 1e-12)
    s_ref = (word.float()[ids] + pos.float()[:L][None] + typ.float()[tt]).view(B * L, H)
    # the gathered sum is exact in fp32 and rounded once to bf16: indexing must be bit-exact
    assert torch.equal(s.cpu(), s_ref.to(BF))
    close(y, restated.layer_norm(s_ref.to(BF).float(), w.float(), b.float(), 1e-12), 8e-3, "embed ln")


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,L,H,hd,masked", [(3, 5, 2, 32, False), (2, 77, 3, 64, True), (2, 257, 2, 64, False), (2, 50, 2, 80, True),
                                             (1, 577, 1, 80, False), (2, 86, 12, 64, True), (2, 12, 2, 16, True),
                                             (3, 197, 2, 64, False), (2, 288, 2, 64, True), (2, 300, 1, 64, False), (1, 128, 1, 64, False),
                                             (150, 257, 2, 64, False), (2, 577, 2, 64, False), (2, 200, 2, 128, False),
                                             (3, 160, 4, 96, True), (1, 1000, 2, 64, True), (40, 86, 12, 64, True)])
def test_attention_fwd_bwd(ops, B, L, H, hd, masked):
    """Every shape runs on the tcgen05 kernels (attention_v3.cu; backward of head_dim 64 with <= 288 tokens on the merged kernel of
    attention_tc.cu, everything else on the recompute pair): head_dim 16..128, one to eight key blocks, resident and streamed K/V, short
    last tiles. (150, 257, 2, 64) and (40, 86, 12, 64) have more (batch, head) items than SMs, so the persistent kernels loop and recycle
    their mbarrier phases and operand rings."""
    W = H * hd
    qkv = rnd(B * L, 3 * W, scale=1.0, seed=B * L + hd)
    d_o = rnd(B * L, W, seed=7)
    key_bias = None
    if masked:
        mask = torch.ones(B, L)
        mask[0, L // 2 :] = 0
        mask[-1, L - 3 :] = 0
        key_bias = (1.0 - mask) * -10000.0
    qf = qkv.float().view(B, L, 3 * W).requires_grad_()
    ref = restated.mha(qf[..., :W], qf[..., W : 2 * W], qf[..., 2 * W :], H, key_bias)
    ref.backward(d_o.float().view(B, L, W))
    kb = key_bias.cuda() if key_bias is not None else None
    o, lse = ops.attention_fwd(qkv.cuda(), B, L, H, hd, key_bias=kb)
    close(o, ref.reshape(B * L, W), 1e-2, "attn fwd")
    # LSE against the definition
    q = qkv.float().view(B, L, 3, H, hd)
    sc = torch.einsum("blhd,bmhd->bhlm", q[:, :, 0], q[:, :, 1]) / math.sqrt(hd)
    if key_bias is not None:
        sc = sc + key_bias[:, None, None, :]
    close(lse, torch.logsumexp(sc, -1), 2e-3, "lse")
    dqkv = ops.attention_bwd(qkv.cuda(), o, d_o.cuda(), lse, B, L, H, hd, key_bias=kb)
    close(dqkv, qf.grad.reshape(B * L, 3 * W), 2e-2, "attn bwd")


# ------------------------------------------------------------------------------------------------------------------
def test_small_helpers(ops):
    x = rnd(40, 72)
    for act, f in [(1, restated.quick_gelu), (2, restated.gelu_erf)]:
        close(ops.act_fwd(x.cuda(), act), f(x.float()), 8e-3, "act")
    rows, W, L = 35 * 6, 136, 35
    x = rnd(rows, W)
    out = torch.zeros(1, W, device="cuda")
    ops.rowsum_periodic(x.cuda(), out, 1)
    close(out[0], x.float().sum(0), 1e-5, "colsum")
    out = torch.zeros(L, W, device="cuda")
    ops.rowsum_periodic(x.cuda(), out, L)
    close(out, x.float().view(6, L, W).sum(0), 1e-5, "periodic sum")
    ids = torch.randint(0, 9, (rows,), generator=torch.Generator().manual_seed(1))
    out = torch.zeros(9, W, device="cuda")
    ops.scatter_add_rows(x.cuda(), ids.cuda(), out, skip_id=0)
    ref = torch.zeros(9, W).index_add_(0, ids, x.float())
    ref[0] = 0
    close(out, ref, 1e-5, "scatter")
    # few output rows (token_type_embeddings): register-accumulating kernel; ragged row count, skip id, 2 and 3 output rows
    for n_out, skip in [(2, -1), (3, 1)]:
        big = rnd(1500 + n_out, W)
        ids2 = torch.randint(0, n_out, (big.shape[0],), generator=torch.Generator().manual_seed(n_out))
        out2 = torch.zeros(n_out, W, device="cuda")
        ops.scatter_add_rows(big.cuda(), ids2.cuda(), out2, skip_id=skip)
        ref2 = torch.zeros(n_out, W).index_add_(0, ids2, big.float())
        if skip >= 0:
            ref2[skip] = 0
        close(out2, ref2, 1e-5, f"scatter few rows ({n_out})")
    y, inv = ops.rownorm_fwd(x.cuda())
    xf = x.float().requires_grad_()
    refn = xf / xf.norm(dim=-1, keepdim=True)
    close(y, refn, 8e-3, "rownorm")
    dy = torch.randn(rows, W, generator=torch.Generator().manual_seed(2))
    refn.backward(dy)
    close(ops.rownorm_bwd(dy.cuda(), x.cuda(), inv), xf.grad, 8e-3, "rownorm bwd")
    f = torch.randn(1000)
    assert torch.equal(ops.cast_f32_bf16(f.cuda(), 0.5).cpu(), (f * 0.5).to(BF))


def test_im2row_bit_exact(ops):
    B, C, H, W, p, Kp = 3, 3, 32, 48, 8, 256
    img = rnd(B, C, H, W)
    got = ops.im2row(img.cuda(), p, Kp).cpu()
    L = (H // p) * (W // p) + 1
    ref = torch.zeros(B, L, Kp, dtype=BF)
    ref[:, 1:, : C * p * p] = restated.patchify(img.float(), p).to(BF)
    assert torch.equal(got, ref.view(B * L, Kp))


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,E,off", [(6, 6, 32, 0), (200, 1000, 64, 300), (128, 520, 768, 128)])
def test_contrast_lse_and_softgrad(ops, M, N, E, off):
    alpha = 14.3
    a = F.normalize(torch.randn(M, E, generator=torch.Generator().manual_seed(M)), dim=-1).to(BF)
    Npad = (N + 7) // 8 * 8
    b = torch.zeros(Npad, E, dtype=BF)
    b[:N] = F.normalize(torch.randn(N, E, generator=torch.Generator().manual_seed(N + 1)), dim=-1).to(BF)
    z = alpha * a.float() @ b[:N].float().t()
    lse_ref = torch.logsumexp(z, 1)
    diag_ref = z[torch.arange(M), torch.arange(M) + off]
    pmax, psum, diag = ops.contrast_lse_partials(a.cuda(), b.cuda()[:N], alpha, off)
    loss_sum = torch.zeros(1, device="cuda")
    lse = ops.contrast_lse_merge((pmax, psum), None, diag, False, loss_sum)
    close(lse, lse_ref, 1e-5, "lse")
    close(diag, diag_ref, 1e-5, "diag")
    close(loss_sum, (lse_ref - diag_ref).sum().view(1), 1e-4, "loss sum")
    coef = 0.37
    dscale = torch.zeros(1, device="cuda")
    G = ops.contrast_softgrad(a.cuda(), b.cuda(), N, alpha, off, lse, coef, 1.0, False, dscale)
    gz = coef * torch.softmax(z, 1)
    gz[torch.arange(M), torch.arange(M) + off] -= coef
    Gref = torch.zeros(M, Npad)
    Gref[:, :N] = gz * alpha
    close(G, Gref, 8e-3, "softgrad G")
    close(dscale, (gz * z).sum().view(1), 2e-3, "dscale")
    # MIL-NCE style merge of two blocks minus the duplicated positive
    lse2 = ops.contrast_lse_merge((pmax, psum), (pmax, psum), diag, True, None)
    close(lse2, torch.log(2 * torch.exp(lse_ref) - torch.exp(diag_ref)), 1e-5, "mil merge")
