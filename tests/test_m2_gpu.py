"""GPU parity of the M²-Encoder (BEiT-3 multiway) path — SURVEY.md §8 rows M1-M3 — through the C-ABI:
the sub-LayerNorm kernels against fp32 torch arithmetic, and the whole ITC path (infer_image / infer_text / loss /
gradients) against the golden vectors of the unmodified reference (tests/golden/m2_tiny.pt) and the CPU oracle."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import restated

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-12))


@pytest.mark.parametrize("W", [64, 256, 1000, 3072, 4096, 6144])
@pytest.mark.parametrize("act", ["none", "gelu", "quickgelu"])
def test_act_layernorm_fwd_bwd(W, act):
    from b200mm import ops

    torch.manual_seed(W)
    rows = 301
    code = {"none": ops.ACT_NONE, "gelu": ops.ACT_GELU_ERF, "quickgelu": ops.ACT_QUICKGELU}[act]
    fn = {"none": lambda t: t, "gelu": F.gelu, "quickgelu": restated.quick_gelu}[act]
    u = (1.5 * torch.randn(rows, W, device="cuda")).to(BF)
    w = (1 + 0.2 * torch.randn(W, device="cuda")).to(BF)
    b = (0.2 * torch.randn(W, device="cuda")).to(BF)
    dy = torch.randn(rows, W, device="cuda").to(BF)
    y, mean, rstd = ops.act_layernorm_fwd(u, code, w, b, 1e-5)
    uf = u.float().requires_grad_()
    wf, bf = w.float().requires_grad_(), b.float().requires_grad_()
    g = fn(uf)
    ref = F.layer_norm(g, (W,), wf, bf, 1e-5)
    # fp32 statistics; the bf16 output carries one rounding: 2^-8 of the output scale
    assert float((y.float() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())
    torch.testing.assert_close(mean, g.mean(-1).detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rstd, torch.rsqrt(g.var(-1, unbiased=False) + 1e-5).detach(), rtol=1e-4, atol=1e-5)
    ref.backward(dy.float())
    dw = torch.zeros(W, device="cuda")
    db = torch.zeros(W, device="cuda")
    du = ops.act_layernorm_bwd(dy, u, code, mean, rstd, w, dw, db)
    assert float((du.float() - uf.grad).abs().max()) <= 2 ** -7 * float(uf.grad.abs().max())
    assert rel_l2(du, uf.grad) < 4e-3
    torch.testing.assert_close(dw, wf.grad, rtol=1e-3, atol=1e-3 * float(wf.grad.abs().max()))
    torch.testing.assert_close(db, bf.grad, rtol=1e-3, atol=1e-3 * float(bf.grad.abs().max()))


def test_act_layernorm_rejects_bad_shapes():
    import b200mm
    from b200mm import ops

    u = torch.zeros(4, 12, device="cuda", dtype=BF)
    with pytest.raises(b200mm.B200mmError, match="W % 8"):
        ops.act_layernorm_fwd(u, ops.ACT_NONE, torch.ones(12, device="cuda", dtype=BF), torch.zeros(12, device="cuda", dtype=BF), 1e-5)
    u = torch.zeros(2, 16384, device="cuda", dtype=BF)
    with pytest.raises(b200mm.B200mmError, match="not supported"):
        ops.act_layernorm_fwd(u, ops.ACT_NONE, torch.ones(16384, device="cuda", dtype=BF), torch.zeros(16384, device="cuda", dtype=BF), 1e-5)


def test_mask_rows_is_exact():
    from b200mm import ops

    x = torch.randn(77, 136, device="cuda").to(BF)
    drop = (torch.rand(77, device="cuda") < 0.4).to(torch.uint8)
    y = ops.mask_rows(x, drop)
    assert torch.equal(y, x * (1 - drop.to(BF))[:, None])
    assert torch.equal(ops.mask_rows(x.clone(), drop, inplace=True), y)


def _build(fx, checkpoint=False, keep=0):
    from b200mm.modules import M2Encoder

    c = fx["config"]
    m = M2Encoder(image_size=c["img"], patch_size=c["patch"], vocab_size=c["vocab"], encoder_embed_dim=c["W"], encoder_attention_heads=c["heads"],
                  encoder_layers=c["layers"], beit3_vl_layers=c["vl_layers"], out_embed_dim=c["out_dim"], max_text_len=c["L"],
                  max_source_positions=c.get("max_source_positions", 1024), xpos_rel_pos=c.get("xpos", False))
    missing, unexpected = m.load_state_dict(fx["state_dict"], strict=False)
    assert not unexpected and all(k.startswith(("norm.", "pooler.")) for k in missing)
    m = m.cuda().to(BF).train()
    m.set_grad_checkpointing(checkpoint)
    m.set_keep_activation(keep)
    return m


def _oracle(sd_src, image, ids, masks, heads, device="cpu", dtype=torch.float32, xpos=None):
    sd = {k: (v.detach().to(BF).to(dtype).to(device).requires_grad_(True) if torch.is_floating_point(v) else v.to(device)) for k, v in sd_src.items()}
    h_i, f_i, fv_i = restated.m2_infer_image(sd, image.to(BF).to(dtype).to(device), heads, xpos)
    h_t, f_t, fv_t = restated.m2_infer_text(sd, ids.to(device), masks.to(device), heads, xpos)
    sdf = {k: v.float() for k, v in sd.items() if k.startswith("logit")}
    loss = restated.m2_itc_loss(sdf, f_i.float(), f_t.float(), fv_i.float(), fv_t.float())
    loss.backward()
    return sd, (h_i, h_t, f_i, f_t, fv_i, fv_t), loss


def _check_against(m, fx_sd, image, ids, masks, heads, golden=None, xpos=None, min_grads=60):
    sd16, (o_hi, o_ht, o_fi, o_ft, o_fvi, o_fvt), o_loss = _oracle(fx_sd, image, ids, masks, heads, xpos=xpos)
    # calibrator: the same oracle arithmetic in bf16 torch eager on the GPU (the reference modules after .cuda().bfloat16())
    sdb, _, _ = _oracle(fx_sd, image, ids, masks, heads, device="cuda", dtype=BF, xpos=xpos)
    # the module applies the reference's inception normalisation (x - 0.5) / 0.5 itself (vlmo_module.py:385); the golden tensors are the
    # already normalised backbone inputs, so the loader-side image is its inverse image
    raw = (image * 0.5 + 0.5)
    img = m.infer_image({"image": [raw.cuda()]})
    txt = m.infer_text({"text_ids": ids.cuda(), "text_masks": masks.cuda()})
    assert img["cls_feats"].dtype == BF and img["cls_feats"].is_cuda
    # hidden states after the final layer_norm; padded text rows are garbage-in-garbage-out in the reference too (they only
    # see valid keys) and are compared as well — they are deterministic functions of the same inputs
    assert rel_l2(img["image_feats"], o_hi) < 1.5e-2, rel_l2(img["image_feats"], o_hi)
    assert rel_l2(txt["text_hidden"], o_ht) < 1.5e-2, rel_l2(txt["text_hidden"], o_ht)
    for got, ref in [(img["cls_feats"], o_fi), (txt["cls_feats"], o_ft), (img["cls_vlffn_feats"], o_fvi), (txt["cls_vlffn_feats"], o_fvt)]:
        assert rel_l2(got, ref) < 1.5e-2, rel_l2(got, ref)
    if golden is not None:  # + weight rounding
        assert rel_l2(img["cls_feats"], golden["img_f"]) < 2e-2 and rel_l2(txt["cls_feats"], golden["txt_f"]) < 2e-2
        assert rel_l2(img["cls_vlffn_feats"], golden["img_fv"]) < 2e-2 and rel_l2(txt["cls_vlffn_feats"], golden["txt_fv"]) < 2e-2
    loss = m.itc_loss(raw.cuda(), ids.cuda(), masks.cuda())
    assert abs(float(loss) - float(o_loss)) < 2e-2 * max(1.0, abs(float(o_loss))), (float(loss), float(o_loss))
    if golden is not None:
        assert abs(float(loss) - float(golden["loss"])) < 3e-2 * max(1.0, abs(float(golden["loss"])))
    loss.backward()
    worst, eager = {}, {}
    with_grad = 0
    for n, p in m.named_parameters():
        ref = sd16[n].grad if n in sd16 else None
        if ref is None or float(ref.abs().max()) == 0.0:
            # experts / members the path does not touch (e.g. text expert B of backbone_vl, norm, pooler, mask_token): no gradient
            assert p.grad is None or float(p.grad.float().abs().max()) == 0.0, n
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        with_grad += 1
        scale = float(ref.abs().max())
        if scale < 1e-5:  # analytically-zero gradients (k_proj biases)
            assert float(p.grad.float().abs().max()) < 1e-3, n
            continue
        if n.startswith("logit"):
            assert abs(float(p.grad) - float(ref)) < 3e-2 * max(1.0, abs(float(ref))), (n, float(p.grad), float(ref))
            continue
        worst[n] = rel_l2(p.grad, ref)
        eager[n] = rel_l2(sdb[n].grad, ref)
        assert worst[n] < max(3e-2, 2.0 * eager[n]), (n, worst[n], eager[n])
    assert with_grad > min_grads
    med = sorted(worst.values())[len(worst) // 2]
    med_eager = sorted(eager.values())[len(eager) // 2]
    assert med < max(2.5e-2, 2.0 * med_eager), (med, med_eager)


@pytest.mark.parametrize("checkpoint,keep", [(False, 0), (True, 0), (False, 3)])
def test_m2_encoder_matches_reference_golden(golden_dir, checkpoint, keep):
    """hd 32 (mma.sync attention path), L 17 / 12, width 64: the golden vectors of the unmodified reference classes."""
    fx = torch.load(os.path.join(golden_dir, "m2_tiny.pt"), weights_only=False)
    m = _build(fx, checkpoint, keep)
    _check_against(m, fx["state_dict"], fx["image"], fx["ids"], fx["masks"], fx["config"]["heads"], golden=fx)


@pytest.mark.parametrize("checkpoint", [False, True])
def test_m2_encoder_xpos_matches_reference_golden(golden_dir, checkpoint):
    """args.xpos_rel_pos = True: XPOS rotary embedding applied in place to q / k of the fused projection (the optional RoPE of the path),
    forward and transposed in backward — vs the golden vectors of the unmodified reference run with XPOS (odd lengths 17 / 11)."""
    fx = torch.load(os.path.join(golden_dir, "m2_tiny_xpos.pt"), weights_only=False)
    m = _build(fx, checkpoint)
    _check_against(m, fx["state_dict"], fx["image"], fx["ids"], fx["masks"], fx["config"]["heads"], golden=fx, xpos=512, min_grads=40)


def test_xpos_apply_kernel_matches_reference_expression():
    """b200mm_xpos_apply vs the reference's apply_rotary_pos_emb expression (oracle restated.xpos), forward and the transposed map."""
    from b200mm import ops
    from b200mm.modules.beit3 import xpos_tables

    torch.manual_seed(0)
    B, L, H, hd = 3, 197, 4, 64
    W = H * hd
    qkv = torch.randn(B * L, 3 * W, device="cuda").to(BF)
    tabs = xpos_tables(L, hd, 512, "cuda")
    ref = qkv.float().cpu().clone()
    for sec, down in ((0, False), (1, True)):
        x = ref[:, sec * W:(sec + 1) * W].reshape(B, L, H, hd).transpose(1, 2).reshape(B * H, L, hd)
        y = restated.xpos(x, 512, downscale=down).view(B, H, L, hd).transpose(1, 2).reshape(B * L, W)
        ref[:, sec * W:(sec + 1) * W] = y
    out = ops.xpos_apply(qkv.clone(), tabs, B, L, H, hd)
    assert torch.equal(out[:, 2 * W:], qkv[:, 2 * W:])  # v untouched
    assert float((out.float().cpu() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())
    # transposed map: <R x, g> == <x, R^T g>
    g = torch.randn_like(qkv)
    rt_g = ops.xpos_apply(g.clone(), tabs, B, L, H, hd, backward=True)
    lhs = float((out.float()[:, :2 * W] * g.float()[:, :2 * W]).sum())
    rhs = float((qkv.float()[:, :2 * W] * rt_g.float()[:, :2 * W]).sum())
    assert abs(lhs - rhs) < 2e-2 * max(abs(lhs), 1.0), (lhs, rhs)


def test_m2_encoder_hd64_matches_oracle():
    """head_dim 64 (tcgen05 attention with the −inf-equivalent key bias), width 256, ffn sub-LN width 1024, 197 / 52 tokens."""
    from b200mm.modules import M2Encoder

    torch.manual_seed(3)
    m = M2Encoder(image_size=224, patch_size=16, vocab_size=320, encoder_embed_dim=256, encoder_attention_heads=4, encoder_layers=2,
                  beit3_vl_layers=1, out_embed_dim=128, max_text_len=52, max_source_positions=64)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias") or "layer_norm" in n or "layernorm" in n or "_ln" in n or "token" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().to(BF).train()
    B, L = 9, 52
    image = torch.randn(B, 3, 224, 224, generator=g)
    ids = torch.randint(1, 320, (B, L), generator=g)
    lens = torch.randint(5, L + 1, (B,), generator=g)
    masks = (torch.arange(L)[None, :] < lens[:, None]).long()
    ids = ids * masks
    _check_against(m, sd, image, ids, masks, 4)


def test_m2_reference_interface():
    """Encoder.forward(token_embeddings=…, encoder_padding_mask=…, multiway_split_position=…) and BEiT3.forward keep the reference's
    keyword interface and dict keys (architecture/encoder.py:388-482, model/BEiT3.py:48-96). (The fused vision + language call is checked
    in tests/test_zz_m2_fused_gpu.py.)"""
    from b200mm.modules import M2Encoder

    m = M2Encoder(image_size=32, patch_size=8, vocab_size=64, encoder_embed_dim=64, encoder_attention_heads=2, encoder_layers=1,
                  beit3_vl_layers=1, out_embed_dim=32, max_text_len=8, max_source_positions=16).cuda().to(BF)
    ids = torch.randint(1, 64, (3, 8), device="cuda")
    pad = torch.zeros(3, 8, dtype=torch.long, device="cuda")
    pad[:, 6:] = 1
    out = m.backbone(textual_tokens=ids, text_padding_position=pad)
    assert out["encoder_out"].shape == (3, 8, 64) and out["multiway_split_position"] == 0
    out2 = m.backbone_vl(src_tokens=None, token_embeddings=out["encoder_out"], encoder_padding_mask=pad, multiway_split_position=-1)
    assert set(out2) >= {"encoder_out", "encoder_embedding", "encoder_padding_mask", "encoder_states", "l_aux", "multiway_split_position"}
    vis = m.backbone(visual_tokens=torch.randn(3, 3, 32, 32, device="cuda"))
    assert vis["encoder_out"].shape == (3, 17, 64)
