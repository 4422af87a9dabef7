"""Parity at BASELINE.json's full sizes (configs[1]: ViT-L/14 + BERT-base, 1024 pairs per GPU; configs[2]: 8192 x 8192
contrastive) through size-independent properties — the oracle cannot run these shapes in seconds, the properties can be
checked exactly or against closed forms:

  GEMM        row-permutation equivariance (bit-exact: the K reduction order of a row does not depend on its position)
              and agreement with the oracle arithmetic on a random sample of output entries
  LayerNorm   unit-affine output has per-row mean 0 / variance 1; LN is idempotent up to bf16 rounding
  attention   with V = 1 every output is 1 (rows of P sum to 1); permuting keys and values together leaves O unchanged
  embedding   bit-exact gather at the full 21128-word vocabulary
  contrastive symmetric: loss(a, b) == loss(b, a); for orthonormal-ish one-hot features the loss has a closed form
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
T_VIT, W_VIT, T_BERT, H_BERT = 1024 * 257, 1024, 1024 * 77, 768


@pytest.fixture(scope="module")
def ops():
    import b200mm

    return b200mm.ops


def test_gemm_full_size_permutation_and_sample(ops):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(T_VIT, W_VIT, device="cuda", generator=g).to(BF)
    w = (torch.randn(4096, W_VIT, device="cuda", generator=g) * 0.03).to(BF)
    bias = torch.randn(4096, device="cuda", generator=g).to(BF)
    y = ops.gemm(a, w, bias=bias, act=ops.ACT_QUICKGELU)
    perm = torch.randperm(T_VIT, device="cuda", generator=g)
    y_p = ops.gemm(a[perm].contiguous(), w, bias=bias, act=ops.ACT_QUICKGELU)
    assert torch.equal(y_p, y[perm]), "row permutation changed results (tile position leaks into the arithmetic)"
    rows = torch.randint(0, T_VIT, (512,), device="cuda", generator=g)
    ref = a[rows].float() @ w.float().t() + bias.float()
    ref = ref * torch.sigmoid(1.702 * ref)
    err = (y[rows].float() - ref).abs().max() / ref.abs().max()
    assert float(err) < 6e-3, float(err)
    # weight gradient shape (reduction over all 263168 tokens, split-K): sampled rows of dW = dY^T X against fp32
    dy = torch.randn(T_VIT, 1024, device="cuda", generator=g).to(BF)
    dw = ops.gemm(dy, a, a_mn=True, b_mn=True, out_f32=True)
    ref_dw = dy[:, :8].float().t() @ a.float()
    err = (dw[:8] - ref_dw).abs().max() / ref_dw.abs().max()
    assert float(err) < 1e-3, float(err)


def test_layernorm_full_size_properties(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn(T_VIT, W_VIT, device="cuda", generator=g) * 3 + 1).to(BF)
    one, zero = torch.ones(W_VIT, device="cuda", dtype=BF), torch.zeros(W_VIT, device="cuda", dtype=BF)
    y, _, mean, rstd = ops.layernorm_fwd(x, one, zero, 1e-5)
    yf = y.float()
    assert float(yf.mean(1).abs().max()) < 2e-2
    assert float((yf.var(1, unbiased=False) - 1).abs().max()) < 3e-2
    torch.testing.assert_close(mean, x.float().mean(1), rtol=1e-4, atol=1e-4)
    y2, _, _, _ = ops.layernorm_fwd(y, one, zero, 1e-5)
    assert float((y2.float() - yf).abs().max()) < 4e-2  # idempotent up to bf16 rounding of y


def test_attention_full_size_properties(ops):
    B, L, H, hd = 1024, 257, 16, 64
    W = H * hd
    g = torch.Generator(device="cuda").manual_seed(2)
    qkv = torch.randn(B * L, 3 * W, device="cuda", generator=g).to(BF)
    qkv[:, 2 * W :] = 1.0  # V = 1  ->  O = sum_j P_ij = 1
    o, lse = ops.attention_fwd(qkv, B, L, H, hd)
    assert float((o.float() - 1).abs().max()) < 1.5e-2
    assert torch.isfinite(lse).all()
    # permute keys and values of every sequence together: O is unchanged (softmax is permutation invariant over keys)
    qkv = torch.randn(64 * L, 3 * W, device="cuda", generator=g).to(BF)
    o1, lse1 = ops.attention_fwd(qkv, 64, L, H, hd)
    perm = torch.randperm(L, device="cuda", generator=g)
    q3 = qkv.view(64, L, 3 * W).clone()
    q3[:, :, W:] = q3[:, perm, W:]
    o2, lse2 = ops.attention_fwd(q3.view(64 * L, 3 * W), 64, L, H, hd)
    assert float((o1.float() - o2.float()).abs().max()) < 2e-2
    torch.testing.assert_close(lse1, lse2, rtol=1e-4, atol=1e-4)


def test_embedding_full_vocab_bit_exact(ops):
    V, Hd, B, L = 21128, 768, 1024, 77
    g = torch.Generator(device="cuda").manual_seed(3)
    word = torch.randn(V, Hd, device="cuda", generator=g).to(BF)
    pos = torch.zeros(512, Hd, device="cuda", dtype=BF)
    typ = torch.zeros(2, Hd, device="cuda", dtype=BF)
    ids = torch.randint(0, V, (B * L,), device="cuda", generator=g)
    ids[-1] = V - 1
    tt = torch.zeros(B * L, dtype=torch.long, device="cuda")
    one, zero = torch.ones(Hd, device="cuda", dtype=BF), torch.zeros(Hd, device="cuda", dtype=BF)
    _, s, _, _ = ops.embed_layernorm_fwd(word, ids, pos, L, typ, tt, one, zero, 1e-12)
    assert torch.equal(s, word[ids])


def test_contrastive_global_batch_8192(ops):
    import b200mm
    from b200mm.contrastive import clip_contrastive_loss

    Bg, E = 8192, 768
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.nn.functional.normalize(torch.randn(Bg, E, device="cuda", generator=g), dim=-1).to(BF)
    b = torch.nn.functional.normalize(torch.randn(Bg, E, device="cuda", generator=g), dim=-1).to(BF)
    ls = torch.tensor(math.log(1 / 0.07), device="cuda")
    l_ab = clip_contrastive_loss(a.clone().requires_grad_(), b.clone().requires_grad_(), ls)
    l_ba = clip_contrastive_loss(b, a, ls)
    assert abs(float(l_ab) - float(l_ba)) < 1e-5 * float(l_ab)
    # oracle on the same rounded features, fp32 on the GPU (2 x 268 MB of logits: fine here, impossible for the CPU oracle in seconds)
    z = float(ls.exp()) * a.float() @ b.float().t()
    d = z.diagonal()
    ref = 0.5 * ((torch.logsumexp(z, 1) - d).mean() + (torch.logsumexp(z, 0) - d).mean())
    assert abs(float(l_ab) - float(ref)) < 2e-5 * float(ref), (float(l_ab), float(ref))
    # identical, exactly orthogonal one-hot features: logits = s*I  ->  loss = log(1 + (B-1) e^{-s}) exactly
    eye = torch.zeros(1024, E, device="cuda", dtype=BF)
    eye[torch.arange(768), torch.arange(768)] = 1.0  # 768 orthonormal rows, the remaining 256 rows are zero vectors
    l_eye = clip_contrastive_loss(eye[:768].contiguous(), eye[:768].contiguous(), ls)
    s = math.exp(float(ls))
    assert abs(float(l_eye) - math.log(1 + 767 * math.exp(-s))) < 1e-6 + 1e-4 * math.log(1 + 767 * math.exp(-s))
