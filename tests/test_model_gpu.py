"""End-to-end parity (GPU): the b200mm modules against the committed golden vectors of the real reference
(tests/golden/, produced by oracle/make_golden.py) and against the CPU oracle on identical (bf16-rounded) weights.

Tolerance (north star: "within 1e-3 rel fp16/bf16"): the product path stores activations in bf16 (2^-9 relative rounding
per tensor), so the end-to-end comparison against the fp32 oracle uses rel-L2 error bounds that are a small multiple of
bf16 epsilon times sqrt(depth); the numbers are written next to each assert.
"""
import os

import pytest
import torch

from oracle import restated

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-12))


def _build(fx, checkpoint=False):
    from b200mm.modules import CNCLIP

    cfg = dict(fx["config"])
    m = CNCLIP(**cfg)
    m.load_state_dict(fx["state_dict"])
    m = m.cuda().to(BF)
    m.set_grad_checkpointing(checkpoint)
    m.train()
    return m, cfg


@pytest.mark.parametrize("name", ["cnclip_tiny.pt", "cnclip_tiny_h80.pt"])
@pytest.mark.parametrize("checkpoint", [False, True])
def test_cnclip_matches_reference_golden(golden_dir, name, checkpoint):
    fx = torch.load(os.path.join(golden_dir, name), weights_only=False)
    m, cfg = _build(fx, checkpoint)
    image, text = fx["image"].cuda(), fx["text"].cuda()
    # oracle on the same bf16-rounded weights/inputs: isolates kernel arithmetic from weight rounding
    sd16 = {k: (v.to(BF).float() if torch.is_floating_point(v) else v) for k, v in fx["state_dict"].items()}
    sd16 = {k: v.clone().requires_grad_(torch.is_floating_point(v)) for k, v in sd16.items()}
    vh = cfg["vision_width"] // cfg["vision_head_width"]
    o_img, o_txt, o_logits, _ = restated.cnclip_forward(sd16, fx["image"].to(BF).float(), fx["text"], vh, cfg["text_num_attention_heads"])
    o_loss = restated.symmetric_info_nce(o_logits)
    o_loss.backward()

    # calibrator: the same oracle arithmetic executed in bf16 by torch eager on the GPU (what the reference's modules do
    # after .cuda().bfloat16()); the product path must not be further from the fp32 oracle than ~2x this
    sdb = {k: (v.detach().to(BF).cuda().requires_grad_(True) if torch.is_floating_point(v) else v.cuda()) for k, v in fx["state_dict"].items()}
    _, _, e_logits, _ = restated.cnclip_forward(sdb, image.to(BF), text, vh, cfg["text_num_attention_heads"])
    restated.symmetric_info_nce(e_logits.float()).backward()

    img, txt = m.encode_normalized(image, text)
    assert img.dtype == BF and img.is_cuda
    # forward: bf16 pipeline vs fp32 oracle. 2 layers deep: rel-L2 <= 1.5e-2 (~4 bf16 eps * sqrt(depth) with LN gain)
    assert rel_l2(img, o_img) < 1.5e-2, rel_l2(img, o_img)
    assert rel_l2(txt, o_txt) < 1.5e-2, rel_l2(txt, o_txt)
    # and against the golden output of the real reference (fp32 weights): adds weight rounding
    assert rel_l2(img, fx["image_features"]) < 2e-2
    assert rel_l2(txt, fx["text_features"]) < 2e-2

    loss = m.contrastive_loss(image, text)
    assert abs(float(loss) - float(o_loss)) < 2e-2 * max(1.0, abs(float(o_loss))), (float(loss), float(o_loss))
    assert abs(float(loss) - float(fx["loss"])) < 3e-2 * max(1.0, abs(float(fx["loss"])))
    loss.backward()
    worst, eager_err = {}, {}
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        assert torch.isfinite(p.grad).all(), n
        ref = sd16[n].grad
        scale = float(ref.abs().max())
        if scale < 1e-6:  # analytically-zero gradients (key biases): only require them to stay negligible
            assert float(p.grad.float().abs().max()) < 1e-3, n
            continue
        if n == "logit_scale":
            # scalar = sum_ij dL/dz_ij * z_ij with heavy cancellation: bound the error by 1e-2 of the un-cancelled magnitude
            zz = o_logits.detach()
            dz = (torch.softmax(zz, 1) + torch.softmax(zz, 0) - 2 * torch.eye(zz.shape[0])) / (2 * zz.shape[0])
            assert abs(float(p.grad) - float(ref)) < 1e-2 * float((dz * zz).abs().sum()), (float(p.grad), float(ref))
            continue
        worst[n] = rel_l2(p.grad, ref)
        eager_err[n] = rel_l2(sdb[n].grad, ref)
        assert worst[n] < max(3e-2, 2.0 * eager_err[n]), (n, worst[n], eager_err[n])
    # the bulk of the parameters: median error no worse than 2x the eager-bf16 median
    med = sorted(worst.values())[len(worst) // 2]
    med_eager = sorted(eager_err.values())[len(eager_err) // 2]
    assert med < max(2.5e-2, 2.0 * med_eager), (med, med_eager)
    # padding_idx row of the word embeddings receives no gradient (nn.Embedding(padding_idx=0), modeling_bert.py:71-73)
    assert float(m.bert.embeddings.word_embeddings.weight.grad[0].abs().max()) == 0.0


def test_forward_returns_reference_tuple(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "cnclip_tiny.pt"), weights_only=False)
    m, _ = _build(fx)
    img, txt, lpi, lpt = m(fx["image"].cuda(), fx["text"].cuda())
    assert lpi.shape == (6, 6) and torch.equal(lpt, lpi.t())
    # logits = exp(logit_scale) * cosine: compare the cosines with an absolute bound (features carry ~1e-2 relative bf16 error)
    alpha = float(fx["state_dict"]["logit_scale"].exp())
    assert float((lpi.float().cpu() - fx["logits_per_image"]).abs().max()) / alpha < 1.5e-2
    lpi.diagonal().sum().backward()
    assert m.logit_scale.grad is not None and m.visual.proj.grad is not None


def test_mil_nce_matches_reference_known_answers(golden_dir):
    from b200mm.contrastive import mil_nce_loss

    fx = torch.load(os.path.join(golden_dir, "losses.pt"), weights_only=False)
    for key in ["mil_b4", "mil_b37"]:
        c = fx[key]
        t16, v16 = c["t"].to(BF), c["v"].to(BF)
        # oracle on the rounded inputs (the kernel's contract) ...
        tf, vf = t16.float().requires_grad_(), v16.float().requires_grad_()
        ref = restated.mil_nce_n1(restated.l1_simi_matrix(tf, vf, 1).view(tf.shape[0], vf.shape[0]))
        ref.backward()
        t = t16.cuda().requires_grad_()
        v = v16.cuda().requires_grad_()
        loss = mil_nce_loss(v, t)
        assert abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref))), (float(loss), float(ref))
        # ... and the reference's own known answer on unrounded inputs (bf16 input rounding only)
        assert abs(float(loss) - float(c["loss"])) < 2e-2 * max(1.0, abs(float(c["loss"])))
        loss.backward()
        assert rel_l2(t.grad, tf.grad) < 1e-2 and rel_l2(v.grad, vf.grad) < 1e-2


@pytest.mark.parametrize("key", ["mil_b3_n2", "mil_b6_n3", "mil_b9_n4"])
def test_mil_nce_n_clips_matches_reference(golden_dir, key):
    """n_clips > 1 (forward_stage1 + get_mil_nce_loss of the unmodified reference, golden) and a larger seeded case vs the oracle"""
    from b200mm.contrastive import mil_nce_loss

    c = torch.load(os.path.join(golden_dir, "losses.pt"), weights_only=False)[key]
    n = c["n"]
    cases = [(c["t"], c["v"], c["loss"])]
    g = torch.Generator().manual_seed(n)
    cases.append((torch.nn.functional.normalize(torch.randn(300, 64, generator=g), dim=-1),
                  torch.nn.functional.normalize(torch.randn(300 * n, 64, generator=g), dim=-1), None))
    for t32, v32, golden in cases:
        t16, v16 = t32.to(BF), v32.to(BF)
        tf, vf = t16.float().requires_grad_(), v16.float().requires_grad_()
        ref = restated.mil_nce_clips(restated.l1_simi_matrix(tf, vf, n))
        ref.backward()
        t, v = t16.cuda().requires_grad_(), v16.cuda().requires_grad_()
        loss = mil_nce_loss(v, t, n_clips=n)
        assert abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref))), (float(loss), float(ref))
        if golden is not None:
            assert abs(float(loss) - float(golden)) < 2e-2 * max(1.0, abs(float(golden)))
        loss.backward()
        assert rel_l2(t.grad, tf.grad) < 1e-2 and rel_l2(v.grad, vf.grad) < 1e-2, (rel_l2(t.grad, tf.grad), rel_l2(v.grad, vf.grad))


def test_no_cpu_fallback():
    import b200mm

    with pytest.raises(b200mm.B200mmError):
        b200mm.ops.gemm(torch.zeros(8, 8, dtype=BF), torch.zeros(8, 8, dtype=BF))


def test_cross_modal_stage2_path_matches_oracle():
    """UnivlVideoBase.prepare_cross_text / prepare_cross_visual / get_cross_output
    (prj/base_vtp/roi_univl/univl/model/univl_video_base.py:168-271) driven through the B200 text encoder exactly as the
    reference drives RobertBertEncoder: shared embeddings (input_ids path and inputs_embeds path with token_type 1 + [SEP]),
    concatenated [text ; clips ; SEP] sequence, additive (1-mask)*-10000 key mask, the same encoder layers, CLS @ text_projection."""
    from b200mm.encoders import B200RobertBertEncoder

    torch.manual_seed(0)
    Hd, heads, layers, E, B, Lt, nc = 128, 2, 2, 64, 5, 11, 4
    enc = B200RobertBertEncoder(pretrained=False, hidden_size=Hd, intermediate_size=256, num_hidden_layers=layers, num_attention_heads=heads,
                                vocab_size=300, max_position_embeddings=64, out_dim=E, hidden_dropout_prob=0.0,
                                attention_probs_dropout_prob=0.0).cuda().to(BF).train()
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if n.endswith("bias") or "LayerNorm.weight" in n:
                p.add_(0.1 * torch.randn_like(p))
    ids = torch.randint(1, 300, (B, Lt)).cuda()
    ids[:, 0] = 101
    cap_mask = torch.ones(B, Lt, dtype=torch.long).cuda()
    cap_mask[1, 7:] = 0
    clips = (0.5 * torch.randn(B, nc, Hd)).cuda().to(BF).requires_grad_()
    vis_pad = torch.zeros(B, nc, dtype=torch.bool).cuda()
    vis_pad[2, 3:] = True

    # ---- the reference's call sequence on the B200 modules
    cap_embed = enc.embeddings(input_ids=ids, token_type_ids=torch.zeros_like(ids))
    sep = enc.embeddings.word_embeddings(torch.full((B,), 102, dtype=torch.long, device="cuda")).unsqueeze(1)
    vis_in = torch.cat([clips, sep], 1)
    vis_embed = enc.embeddings(inputs_embeds=vis_in, token_type_ids=torch.ones(B, nc + 1, dtype=torch.long, device="cuda"))
    vis_mask = torch.cat([vis_pad.logical_not().long(), torch.ones(B, 1, dtype=torch.long, device="cuda")], 1)
    embed = torch.cat([cap_embed, vis_embed], 1)
    mask = torch.cat([cap_mask, vis_mask], 1)
    ext = (1.0 - mask.unsqueeze(1).unsqueeze(2).float()) * -10000.0
    seq = enc.encoder(embed, attention_mask=ext, head_mask=[None] * len(enc.encoder.layer))[0]
    pooled = seq[:, 0, :].float() @ enc.text_projection.float()
    pooled.square().sum().backward()

    # ---- oracle (fp32, same bf16-rounded parameters)
    sd = {k: v.detach().float().cpu().requires_grad_(torch.is_floating_point(v)) for k, v in enc.module.state_dict().items()}
    clips_o = clips.detach().float().cpu().requires_grad_()
    o_cap = restated.bert_embeddings(sd, "embeddings.", ids.cpu(), torch.zeros(B, Lt, dtype=torch.long))
    o_sep = sd["embeddings.word_embeddings.weight"][torch.full((B,), 102)].unsqueeze(1)
    o_vis = restated.bert_embeddings(sd, "embeddings.", None, torch.ones(B, nc + 1, dtype=torch.long), inputs_embeds=torch.cat([clips_o, o_sep], 1))
    o_embed = torch.cat([o_cap, o_vis], 1)
    o_seq = restated.bert_encoder(sd, "encoder.", o_embed, heads, mask.cpu().float())
    o_pooled = o_seq[:, 0, :] @ enc.text_projection.detach().float().cpu()
    o_pooled.square().sum().backward()
    valid = mask.bool().cpu()
    assert rel_l2(seq.float().cpu()[valid], o_seq[valid]) < 1.5e-2
    assert rel_l2(pooled, o_pooled) < 1.5e-2
    assert rel_l2(clips.grad, clips_o.grad) < 4e-2
    g_word = enc.embeddings.word_embeddings.weight.grad
    assert rel_l2(g_word, sd["embeddings.word_embeddings.weight"].grad) < 4e-2
    assert float(g_word[0].abs().max()) == 0.0  # padding_idx


def test_moco_queue_loss_ema_and_enqueue():
    """MocoUtils.moco_loss / momentum_update_key_encoder / dequeue_and_enqueue
    (prj/base_vtp/roi_univl/univl/model/moco_utils.py:55-108) on the fused kernels vs the oracle (restated.moco_nce, itself pinned to the
    reference's known answer in tests/golden/losses.pt)."""
    import torch.nn.functional as F

    from b200mm.moco import B200MocoUtils, moco_nce

    torch.manual_seed(0)
    N, E, K, T = 37, 64, 1024, 0.05
    q = F.normalize(torch.randn(N, E), dim=-1).to(BF)
    kp = F.normalize(q.float() + 0.5 * torch.randn(N, E), dim=-1).to(BF)
    queue = F.normalize(torch.randn(E, K), dim=0).to(BF)
    qf = q.float().requires_grad_()
    pos = (qf * kp.float()).sum(-1, keepdim=True)
    neg = qf @ queue.float()
    ref = restated.moco_nce(pos, neg, T)
    ref.backward()
    qc = q.cuda().requires_grad_()
    loss = moco_nce(qc, kp.cuda(), queue.cuda(), T)
    assert abs(float(loss) - float(ref)) < 2e-4 * max(1.0, abs(float(ref))), (float(loss), float(ref))
    loss.backward()
    assert rel_l2(qc.grad, qf.grad) < 1e-2, rel_l2(qc.grad, qf.grad)

    # module: same buffer names / shapes as the reference, EMA of the key encoder, ring-buffer enqueue
    enc = torch.nn.Linear(16, 16).cuda().to(BF)
    mu = B200MocoUtils(dict(hidden_size=E, K=K, M=0.99, T=T), txt_encoder=enc).cuda()
    assert set(dict(mu.named_buffers())) == {"txt_queue", "txt_queue_ptr"} and mu.txt_queue.shape == (E, K)
    w_k0 = mu.txt_encoder_k.weight.detach().clone()
    with torch.no_grad():
        enc.weight.add_(1.0)
    mu.momentum_update_key_encoder()
    expect = w_k0 * 0.99 + enc.weight.float() * 0.01
    torch.testing.assert_close(mu.txt_encoder_k.weight, expect, rtol=1e-6, atol=1e-6)
    keys = F.normalize(torch.randn(24, E), dim=-1).cuda().to(BF)
    l0 = mu.moco_loss_fused(qc.detach(), kp.cuda(), "txt")
    mu.dequeue_and_enqueue(None, keys)
    assert int(mu.txt_queue_ptr) == 24
    torch.testing.assert_close(mu.txt_queue[:, :24], keys.float().t())
    l1 = mu.moco_loss_fused(qc.detach(), kp.cuda(), "txt")
    negs = qc.detach().float() @ mu.txt_queue.to(BF).float()
    pos_c = (qc.detach().float() * kp.cuda().float()).sum(-1, keepdim=True)
    assert abs(float(l1) - float(restated.moco_nce(pos_c.cpu(), negs.cpu(), T))) < 2e-4 * max(1.0, float(l1))
    assert float(l0) != float(l1)


def test_video_heads_frame_pooling_matches_reference_arithmetic():
    """a11: forward_img_encoder / forward_text_encoder of univl_video_base.py:56-166 (frame mean-pool under the mask, img_proj, normalise)
    against the reference's own tensor expressions evaluated in fp32 on the same encoder outputs."""
    from b200mm import video

    torch.manual_seed(0)
    b, n_clips, n_frames, c, hidden = 3, 2, 4, 32, 48

    class FakeEncoder(torch.nn.Module):  # stands in for VitImageEncoder: returns the structures of clip_visual_encoder.py:73-94
        out_dim = c

        def __init__(self):
            super().__init__()
            self.feat = torch.nn.Parameter(torch.randn(b, n_clips * n_frames, c, 1, 1, device="cuda").to(BF))
            self.mask = torch.zeros(b, n_clips * n_frames, 1, 1, dtype=torch.bool, device="cuda")
            self.mask[0, 1] = self.mask[0, 2] = self.mask[2, 7] = True

        def forward(self, image, image_mask):
            return dict(grid_feature=self.feat, grid_mask=self.mask, grid_feature_with_pos=None)

    enc = FakeEncoder()
    img_proj = torch.nn.Parameter((torch.randn(c, hidden, device="cuda") * 0.2).to(BF))
    out = video.forward_img_encoder(enc, None, None, [n_clips] * b, [n_frames] * b, img_proj=img_proj)
    assert set(out) == {"visual_embed", "visual_mask", "visual_grid_shape", "clip_feature"}
    assert out["visual_embed"].shape == (b, n_clips, hidden) and out["clip_feature"].shape == (b * n_clips, hidden)
    assert out["visual_mask"].dtype == torch.bool and not out["visual_mask"].any() and tuple(out["visual_grid_shape"]) == (1, 1)
    # reference expressions (univl_video_base.py:70-73, :84-95, :114) in fp32
    feat32 = enc.feat.detach().float().requires_grad_()
    proj32 = img_proj.detach().float().requires_grad_()
    gf = torch.einsum("bnchw, cj -> bnjhw", feat32, proj32)
    gf = gf.contiguous().view(b * n_clips, n_frames, *gf.shape[2:])
    gm = enc.mask.contiguous().view(b * n_clips, n_frames, 1, 1)
    g = gf.transpose(1, 2).flatten(2)
    m = ~gm.flatten(1).unsqueeze(1).expand_as(g)
    clip = (g * m).sum(-1) / m.sum(-1)
    ref_tokens = clip.view(b, n_clips, hidden)
    ref_feat = torch.nn.functional.normalize(clip, p=2, dim=-1)
    assert rel_l2(out["visual_embed"], ref_tokens) < 1e-2
    assert rel_l2(out["clip_feature"], ref_feat) < 1e-2
    w = torch.randn_like(ref_feat)
    (out["clip_feature"].float() * w).sum().backward()
    (ref_feat * w).sum().backward()
    assert rel_l2(enc.feat.grad, feat32.grad) < 2e-2, rel_l2(enc.feat.grad, feat32.grad)
    assert rel_l2(img_proj.grad, proj32.grad) < 2e-2
    # masked frames receive exactly zero gradient
    assert float(enc.feat.grad[0, 1].abs().max()) == 0.0 and float(enc.feat.grad[2, 7].abs().max()) == 0.0

    class FakeText(torch.nn.Module):
        def forward(self, input_ids, attention_mask):
            torch.manual_seed(1)
            return torch.randn(4, 6, 16, device="cuda").to(BF), torch.randn(4, 16, device="cuda").to(BF)

    t = video.forward_text_encoder(FakeText(), torch.zeros(4, 6, dtype=torch.long, device="cuda"), torch.ones(4, 6, device="cuda"))
    assert set(t) == {"sequence_output", "pooled_output", "input_mask", "words_importance"} and t["words_importance"] is None
    assert float((t["pooled_output"].float().norm(dim=-1) - 1).abs().max()) < 1e-2


def test_memory_policies_give_the_same_step(golden_dir):
    """keep-activation / keep-LayerNorm / block checkpointing only change WHAT is kept for backward: loss identical, gradients equal up to
    the order of the fp32 atomics that accumulate the small vector gradients."""
    fx = torch.load(os.path.join(golden_dir, "cnclip_tiny.pt"), weights_only=False)
    image, text = fx["image"].cuda(), fx["text"].cuda()
    results = {}
    for policy in ["plain", "keep_ln", "keep_act", "keep_both", "checkpoint"]:
        m, _ = _build(fx, checkpoint=(policy == "checkpoint"))
        if policy in ("keep_ln", "keep_both"):
            m.visual.set_keep_layernorm(2)
        if policy in ("keep_act", "keep_both"):
            m.visual.set_keep_activation(1)
        loss = m.contrastive_loss(image, text)
        loss.backward()
        results[policy] = (float(loss), {n: p.grad.float().clone() for n, p in m.named_parameters()})
    base_loss, base = results["plain"]
    for policy, (l, grads) in results.items():
        assert abs(l - base_loss) <= 1e-5 * abs(base_loss), (policy, l, base_loss)  # the loss is summed with fp32 atomics (order varies)
        for n, g in grads.items():
            if float(base[n].abs().max()) < 1e-6:
                continue
            # keep_act reuses the forward's act(fp32 accumulator), the other policies recompute act(bf16-rounded pre-activation): one bf16
            # rounding apart in the c_proj weight gradient; everything else differs only by atomic order
            assert rel_l2(g, base[n]) < (5e-2 if "keep_act" in policy or policy == "keep_both" else 2e-2), (policy, n, rel_l2(g, base[n]))
