"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the UNMODIFIED reference (oracle/ref_loader.py).

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The fixtures are small (tiny model configs) and committed, because the reference tree does not travel to the GPU box.

Fixtures
  cnclip_tiny.pt   reference CNCLIP (ViT 2L/64w/patch 8 @32px + BERT 2L/64h, embed 32), fp32, dropout 0:
                   state_dict, inputs, forward outputs, symmetric-CE loss, gradients of every parameter
  cnclip_tiny_h80.pt  same with vision_head_width=80-style odd head dim (width 160, 2 heads of 80) — ViT-H head geometry
  m2_tiny.pt       reference BEiT3 (multiway, 2 layers) + backbone_vl Encoder (1 layer) + ITC heads composed exactly as
                   VLMo.infer_image / infer_text (vlmo_module.py:323-405): hiddens, both feature pairs, ITC loss, all gradients
  bert_dropout.pt  reference BertModel in training mode with dropout 0.1 / 0.2 and PRESET (counter-based) masks: output + all gradients
  losses.pt        get_mil_nce_loss / get_l1_simi_matrix / moco_loss known answers incl. SURVEY.md §8c (1)
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY = dict(
    embed_dim=32,
    image_resolution=32,
    vision_layers=2,
    vision_width=64,
    vision_patch_size=8,
    vocab_size=21128,  # FullTokenizer's [PAD]=0 lookup needs the bundled vocab; embeddings table sliced below
    text_attention_probs_dropout_prob=0.0,
    text_hidden_act="gelu",
    text_hidden_dropout_prob=0.0,
    text_hidden_size=64,
    text_initializer_range=0.02,
    text_intermediate_size=256,
    text_max_position_embeddings=64,
    text_num_attention_heads=2,
    text_num_hidden_layers=2,
    text_type_vocab_size=2,
    vision_head_width=32,
)


def synth_text(B, L, vocab, gen):
    """[CLS]=101 first, random ids, [SEP]=102 at the last valid position, [PAD]=0 after."""
    ids = torch.randint(1, vocab, (B, L), generator=gen)
    ids[:, 0] = 101
    lens = torch.randint(4, L + 1, (B,), generator=gen)
    for b in range(B):
        ids[b, lens[b] - 1] = 102
        ids[b, lens[b] :] = 0
    return ids


def make_cnclip(name, cfg, B, L, vocab_used):
    cfg = dict(cfg)
    cfg["vocab_size"] = vocab_used
    model = ref_loader.build_cnclip(cfg, seed=0, dropout=0.0)
    # give biases / LN params / logit-scale non-trivial values so that parity covers them
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias") or "LayerNorm.weight" in n or "ln_" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    model.train()
    gen = torch.Generator().manual_seed(1234)
    image = torch.randn(B, 3, cfg["image_resolution"], cfg["image_resolution"], generator=gen)
    text = synth_text(B, L, vocab_used, gen)
    img_f, txt_f, lpi, lpt = model(image, text)
    labels = torch.arange(B)
    loss = 0.5 * (F.cross_entropy(lpi, labels) + F.cross_entropy(lpt, labels))
    loss.backward()
    fx = {
        "config": cfg,
        "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "image": image,
        "text": text,
        "image_features": img_f.detach(),
        "text_features": txt_f.detach(),
        "logits_per_image": lpi.detach(),
        "loss": loss.detach(),
        "grads": {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None},
    }
    torch.save(fx, os.path.join(OUT, name))
    print(name, "loss", float(loss.detach()), "params", sum(p.numel() for p in model.parameters()))


def make_losses():
    fn = ref_loader.load_loss_functions()
    out = {}
    # SURVEY.md §8c known answer (1)
    torch.manual_seed(0)
    B, D = 4, 8
    t = F.normalize(torch.randn(B, D)).requires_grad_()
    v = F.normalize(torch.randn(B, D)).requires_grad_()
    sim = fn.get_l1_simi_matrix(None, t, v, 1, True).view(B, B)
    loss = fn.get_mil_nce_loss(None, sim, B, 1)
    loss.backward()
    out["mil_b4"] = dict(t=t.detach(), v=v.detach(), loss=loss.detach(), dt=t.grad.clone(), dv=v.grad.clone())
    # a larger, un-normalised case
    torch.manual_seed(1)
    B, D = 37, 24
    t = (0.7 * torch.randn(B, D)).requires_grad_()
    v = (0.7 * torch.randn(B, D)).requires_grad_()
    sim = fn.get_l1_simi_matrix(None, t, v, 1, True).view(B, B)
    loss = fn.get_mil_nce_loss(None, sim, B, 1)
    loss.backward()
    out["mil_b37"] = dict(t=t.detach(), v=v.detach(), loss=loss.detach(), dt=t.grad.clone(), dv=v.grad.clone())
    # n_clips > 1, exactly as forward_stage1 drives it (univl_video_ret.py:357-387): repeat the text rows, then get_mil_nce_loss
    for key, (B, n, D, seed) in {"mil_b3_n2": (3, 2, 8, 5), "mil_b6_n3": (6, 3, 16, 6), "mil_b9_n4": (9, 4, 24, 7)}.items():
        torch.manual_seed(seed)
        t = F.normalize(torch.randn(B, D)).requires_grad_()
        v = F.normalize(torch.randn(B * n, D)).requires_grad_()
        l1 = fn.get_l1_simi_matrix(None, t, v, n, True)                     # [B, B, n]
        mil = l1.unsqueeze(1).repeat([1, n, 1, 1]).view(B * n, B * n)
        loss = fn.get_mil_nce_loss(None, mil, B, n)
        loss.backward()
        out[key] = dict(t=t.detach(), v=v.detach(), n=n, loss=loss.detach(), dt=t.grad.clone(), dv=v.grad.clone())
    # moco
    import types

    torch.manual_seed(2)
    pos = torch.randn(9, 1)
    neg = torch.randn(9, 50)
    out["moco"] = dict(pos=pos, neg=neg, T=0.05, loss=fn.moco_loss(types.SimpleNamespace(T=0.05), pos, neg).detach())
    torch.save(out, os.path.join(OUT, "losses.pt"))
    print("losses.pt", {k: float(v["loss"]) for k, v in out.items()})


def make_retrieval():
    """Retrieval metrics of the unmodified reference on seeded embeddings (square 1:1 case and a 5-captions-per-image case)."""
    import numpy as np

    fn = ref_loader.load_retrieval_metric_functions()
    out = {}
    torch.manual_seed(3)
    t = F.normalize(torch.randn(48, 16) + 0.0, dim=-1)
    v = F.normalize(t + 0.9 * torch.randn(48, 16), dim=-1)
    sim = (t @ v.t()).numpy()
    out["square"] = dict(t=t, v=v, ranks=torch.from_numpy(np.asarray(fn._compute_retrieval_metrics(sim)).astype(np.int64)),
                         recall={k: float(x) for k, x in fn._cal_recall(sim).items()})
    # 12 images x 5 captions: texts 5i..5i+4 belong to image i
    torch.manual_seed(4)
    vis = F.normalize(torch.randn(12, 16), dim=-1)
    txt = F.normalize(vis.repeat_interleave(5, 0) + 0.8 * torch.randn(60, 16), dim=-1)
    t2v = [[i // 5] for i in range(60)]
    v2t = [list(range(5 * i, 5 * i + 5)) for i in range(12)]
    simm = (txt @ vis.t()).numpy()
    out["multi_gt"] = dict(t=txt, v=vis, t2v=t2v, v2t=v2t, metrics={k: float(x) for k, x in fn._cal_sym_recall(simm, t2v, v2t).items()})
    torch.save(out, os.path.join(OUT, "retrieval.pt"))
    print("retrieval.pt", out["square"]["recall"], out["multi_gt"]["metrics"])


def make_m2(name="m2_tiny.pt", W=64, heads=2, layers=2, vl_layers=1, img=32, patch=8, vocab=256, L=12, B=5, out_dim=32, xpos=False,
            max_source_positions=1024):
    """M²-Encoder path with the UNMODIFIED reference classes. VLMo itself is a LightningModule whose constructor loads
    tokenizers/timm (not importable here), so its infer_image / infer_text bodies (vlmo_module.py:323-405) are composed
    here line by line from the same sub-modules and parameter names (backbone, backbone_vl, itc_*_proj, logit_*scale)."""
    import copy

    import numpy as np

    m2 = ref_loader.load_m2()
    torch.manual_seed(0)
    args = m2.EncoderConfig(img_size=img, patch_size=patch, vocab_size=vocab, multiway=True, layernorm_embedding=False,
                            normalize_output=True, no_output_layer=True, drop_path_rate=0, encoder_embed_dim=W,
                            encoder_attention_heads=heads, encoder_layers=layers, encoder_ffn_embed_dim=4 * W,
                            checkpoint_activations=False, max_text_len=L, xpos_rel_pos=xpos, xpos_scale_base=512,
                            max_source_positions=max_source_positions)
    model = torch.nn.Module()
    model.backbone = m2.BEiT3(args)
    vl_args = copy.copy(args)
    vl_args.encoder_layers = vl_layers
    model.backbone_vl = m2.Encoder(vl_args)                       # vlmo_module.py:171-174
    model.itc_text_proj = m2.heads.ITCHead(W, out_dim)            # :186-194
    model.itc_image_proj = m2.heads.ITCHead(W, out_dim)
    model.itc_vl_text_proj = m2.heads.ITCHead(W, out_dim)
    model.itc_vl_image_proj = m2.heads.ITCHead(W, out_dim)
    model.logit_scale = torch.nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
    model.logit_vl_scale = torch.nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias") or "layer_norm" in n or "layernorm" in n or "_ln" in n or "token" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if "embed_positions" in n or "itc_" in n:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    model.train()
    gen = torch.Generator().manual_seed(1234)
    image = torch.randn(B, 3, img, img, generator=gen)
    ids = synth_text(B, L, vocab, gen)
    masks = (ids != 0).long()
    # ---- infer_image (:385-398)
    vffn = model.backbone(visual_tokens=image)["encoder_out"]
    vl_v = model.backbone_vl(src_tokens=None, token_embeddings=vffn, multiway_split_position=-1)["encoder_out"]
    img_f = model.itc_image_proj(vffn[:, 0])
    img_f = img_f / img_f.norm(dim=-1, keepdim=True)
    img_fv = model.itc_vl_image_proj(vl_v[:, 0])
    img_fv = img_fv / img_fv.norm(dim=-1, keepdim=True)
    # ---- infer_text (:333-355)
    pad = 1 - masks
    lffn = model.backbone(textual_tokens=ids, text_padding_position=pad)["encoder_out"]
    vl_t = model.backbone_vl(src_tokens=None, token_embeddings=lffn, encoder_padding_mask=pad, multiway_split_position=-1)["encoder_out"]
    txt_f = model.itc_text_proj(lffn[:, 0])
    txt_f = txt_f / txt_f.norm(dim=-1, keepdim=True)
    txt_fv = model.itc_vl_text_proj(vl_t[:, 0])
    txt_fv = txt_fv / txt_fv.norm(dim=-1, keepdim=True)
    # ---- similarity as m2_encoder.py:92-95, symmetric CE on both head pairs
    labels = torch.arange(B)
    lg = model.logit_scale.exp() * img_f @ txt_f.t()
    lgv = model.logit_vl_scale.exp() * img_fv @ txt_fv.t()
    loss = 0.5 * (F.cross_entropy(lg, labels) + F.cross_entropy(lg.t(), labels)) + 0.5 * (F.cross_entropy(lgv, labels) + F.cross_entropy(lgv.t(), labels))
    loss.backward()
    itc_grads = {n: p_.grad.detach().clone() for n, p_ in model.named_parameters() if p_.grad is not None}
    # ---- fused vision + language input through the SAME backbone (BEiT3.forward with both modalities: multiway split inside the sequence)
    for p_ in model.parameters():
        p_.grad = None
    fused = model.backbone(textual_tokens=ids, visual_tokens=image, text_padding_position=pad)
    assert fused["multiway_split_position"] == vffn.shape[1]
    fused_out = fused["encoder_out"]
    Rf = torch.randn(fused_out.shape, generator=torch.Generator().manual_seed(77))
    (fused_out * Rf * (1 - torch.cat([torch.zeros(B, vffn.shape[1], dtype=torch.long), pad], 1)).unsqueeze(-1)).sum().backward()
    fused_grads = {n: p_.grad.detach().clone() for n, p_ in model.named_parameters() if p_.grad is not None}
    fx = {
        "fused_hidden": fused_out.detach(), "fused_proj": Rf, "fused_grads": fused_grads,
        "config": dict(W=W, heads=heads, layers=layers, vl_layers=vl_layers, img=img, patch=patch, vocab=vocab, L=L, out_dim=out_dim,
                       xpos=xpos, max_source_positions=max_source_positions),
        "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "image": image, "ids": ids, "masks": masks,
        "image_hidden": vffn.detach(), "text_hidden": lffn.detach(),
        "img_f": img_f.detach(), "txt_f": txt_f.detach(), "img_fv": img_fv.detach(), "txt_fv": txt_fv.detach(),
        "logits": lg.detach(), "loss": loss.detach(),
        "grads": itc_grads,
    }
    torch.save(fx, os.path.join(OUT, name))
    print(name, "loss", float(loss.detach()), "params", sum(p.numel() for p in model.parameters()), "with grad", len(fx["grads"]))


def make_stage2():
    """Stage-2 cross-modal scoring with the UNMODIFIED reference functions (ref_loader.load_stage2 / build_stage2_self):
    `_cross_similarity` on a ragged (7 texts x 4 videos, block size 5 + 2) problem and `_cross_similarity_hard_mining` for both
    re-sampling methods, each followed by get_mil_nce_loss with the 'median' row weights of forward_stage2 (:414-433)."""
    out = {}
    me = ref_loader.build_stage2_self(hidden=64, heads=2, layers=2, inter=128, out_dim=32, seed=0)
    sd = {"cross_encoder." + k: v.detach().clone() for k, v in me.module.cross_encoder.state_dict().items()}
    sd["text_projection"] = me.module.text_encoder.text_projection.detach().clone()
    sd.update({"similarity_dense." + k: v.detach().clone() for k, v in me.similarity_dense.state_dict().items()})
    out["state_dict"] = sd
    out["config"] = dict(hidden=64, heads=2, layers=2, inter=128, out_dim=32)
    params = dict(me.named_parameters())

    def names(n):  # reference parameter name -> fixture state-dict name
        return n.replace("module.cross_encoder.", "cross_encoder.").replace("module.text_encoder.text_projection", "text_projection")

    torch.manual_seed(11)
    Bt, St, Bv, Sv, H = 7, 6, 4, 3, 64
    seq = torch.randn(Bt, St, H, requires_grad=True)
    vis = torch.randn(Bv, Sv, H, requires_grad=True)
    am = torch.ones(Bt, St, dtype=torch.long)
    am[2, 4:] = 0
    am[5, 3:] = 0
    vm = torch.ones(Bv, Sv, dtype=torch.long)
    vm[1, 2:] = 0
    logits = me._cross_similarity(seq, vis, am, vm, 1)
    logits.square().sum().backward()
    out["cross"] = dict(seq=seq.detach(), vis=vis.detach(), am=am, vm=vm, logits=logits.detach(), d_seq=seq.grad.clone(), d_vis=vis.grad.clone(),
                        grads={names(n): p.grad.clone() for n, p in params.items() if p.grad is not None})
    for method in ["top_k", "nearliest"]:
        for p in params.values():
            p.grad = None
        me.config.re_sample_method = method
        torch.manual_seed(12)
        B = 6
        seq = torch.randn(B, St, H, requires_grad=True)
        vis = torch.randn(B, Sv, H, requires_grad=True)
        am = torch.ones(B, St, dtype=torch.long)
        am[1, 4:] = 0
        vm = torch.ones(B, Sv, dtype=torch.long)
        vm[3, 1:] = 0
        l1 = torch.randn(B, B) + 2.0 * torch.eye(B)
        l1c = l1.clone()
        l2 = me._cross_similarity_hard_mining((vis, vm, None, 1, None), (seq, am, None, B, None), l1c)
        # forward_stage2 :414-433 verbatim (single process: beg_idx = 0)
        l1_diag = torch.diag(l1c[0:B, 0:B])
        l1_median, l1_minimum = torch.mean(l1_diag), torch.min(l1_diag)
        weight_vector = torch.ones(l1_diag.shape, dtype=l1_diag.dtype)
        for i in range(len(l1_diag)):
            if l1_diag[i] > l1_median:
                weight_vector[i] = max((l1_median - l1_minimum) / (l1_diag[i] - l1_minimum), 0.2)
        mil = l2.unsqueeze(1).repeat([1, 1, 1, 1]).view(B, B)
        loss = me.get_mil_nce_loss(mil, B, weight_vector=weight_vector)
        loss_unweighted = me.get_mil_nce_loss(mil, B).detach()
        loss.backward()
        out["hard_" + method] = dict(seq=seq.detach(), vis=vis.detach(), am=am, vm=vm, l1=l1, logits=l2.detach(), weights=weight_vector,
                                     loss=loss.detach(), loss_unweighted=loss_unweighted, d_seq=seq.grad.clone(), d_vis=vis.grad.clone(),
                                     grads={names(n): p.grad.clone() for n, p in params.items() if p.grad is not None})
    torch.save(out, os.path.join(OUT, "stage2.pt"))
    print("stage2.pt", tuple(out["cross"]["logits"].shape), {k: float(v["loss"]) for k, v in out.items() if k.startswith("hard")})


def make_bert_dropout(name="bert_dropout.pt", B=4, L=24, Hd=64, heads=2, layers=2, inter=128, vocab=300, p_hidden=0.1, p_attn=0.2, seed=77):
    """The UNMODIFIED reference BertModel in TRAINING mode with dropout p > 0. torch.nn.functional.dropout (what nn.Dropout.forward calls) is
    replaced for the duration of the forward by a function that applies y = x * keep / (1 - p) with PRESET masks, consumed in call order
    (embeddings; then per layer: attention probabilities, self-output, output — modeling_bert.py:101,158,180,232). The masks are the
    counter-based ones b200mm generates from its per-site seed sequence (b200mm.ops.manual_seed(seed) / next_dropout_seed), so the fixture pins
    (a) where the oracle places each dropout and (b) what the B200 modules must produce from that seed."""
    from b200mm import ops
    from oracle import restated

    ns = ref_loader.load()
    cfg = ns.configuration_bert.BertConfig(vocab_size_or_config_json_file=vocab, hidden_size=Hd, num_hidden_layers=layers, num_attention_heads=heads,
                                           intermediate_size=inter, hidden_dropout_prob=p_hidden, attention_probs_dropout_prob=p_attn,
                                           max_position_embeddings=32, layer_norm_eps=1e-12)
    torch.manual_seed(0)
    model = ns.modeling_bert.BertModel(cfg)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p_ in model.named_parameters():
            if n.endswith("bias") or "LayerNorm.weight" in n:
                p_.add_(0.1 * torch.randn(p_.shape, generator=g))
    model.train()
    gen = torch.Generator().manual_seed(1)
    ids = torch.randint(1, vocab, (B, L), generator=gen)
    mask = torch.ones(B, L, dtype=torch.long)
    mask[1, 17:] = 0
    ids[1, 17:] = 0
    ops.manual_seed(seed)
    seeds = [ops.next_dropout_seed() for _ in range(1 + 3 * layers)]
    calls = []

    def preset_dropout(x, p=0.5, training=True, inplace=False):
        i = len(calls)
        if x.dim() == 4:   # attention probabilities [B, heads, L, L]
            keep = restated.attention_dropout_keep(seeds[i], x.shape[0], x.shape[1], x.shape[2], p)
        else:              # hidden states [B, L, H]
            keep = restated.dropout_keep(seeds[i], x.shape[0] * x.shape[1], x.shape[2], p).view(x.shape)
        calls.append((tuple(x.shape), p))
        return x * keep.to(x.dtype) / (1.0 - p)

    real = F.dropout
    torch.nn.functional.dropout = preset_dropout
    try:
        out = model(ids, attention_mask=mask)[0]
    finally:
        torch.nn.functional.dropout = real
    assert len(calls) == len(seeds), calls
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(123))
    (out * probe).sum().backward()
    fx = {"config": dict(vocab=vocab, hidden=Hd, heads=heads, layers=layers, inter=inter, p_hidden=p_hidden, p_attn=p_attn, seed=seed),
          "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()}, "ids": ids, "mask": mask, "seeds": seeds, "calls": calls,
          "probe": probe, "out": out.detach(), "grads": {n: p_.grad.detach().clone() for n, p_ in model.named_parameters() if p_.grad is not None}}
    torch.save(fx, os.path.join(OUT, name))
    print(name, "dropout calls", calls[:4], "...", "out norm", float(out.norm()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--bert-dropout-only" in sys.argv:
        make_bert_dropout()
        sys.exit(0)
    if "--stage2-only" in sys.argv:
        make_stage2()
        sys.exit(0)
    if "--m2-only" in sys.argv:
        make_m2()
        make_m2("m2_tiny_xpos.pt", layers=1, vl_layers=1, L=11, B=4, xpos=True, max_source_positions=32)
        sys.exit(0)
    if "--m2-xpos-only" in sys.argv:
        make_m2("m2_tiny_xpos.pt", layers=1, vl_layers=1, L=11, B=4, xpos=True, max_source_positions=32)
        sys.exit(0)
    if "--retrieval-only" in sys.argv:
        make_retrieval()
        sys.exit(0)
    if "--losses-only" in sys.argv:
        make_losses()
        sys.exit(0)
    make_cnclip("cnclip_tiny.pt", TINY, B=6, L=16, vocab_used=512)
    h80 = dict(TINY, vision_width=160, vision_head_width=80, vision_layers=1, text_hidden_size=32,
               text_num_attention_heads=2, text_num_hidden_layers=1, text_intermediate_size=64, embed_dim=16,
               image_resolution=32, vision_patch_size=16)
    make_cnclip("cnclip_tiny_h80.pt", h80, B=5, L=12, vocab_used=128)
    make_losses()
    make_retrieval()
    make_m2()
    make_m2("m2_tiny_xpos.pt", layers=1, vl_layers=1, L=11, B=4, xpos=True, max_source_positions=32)
    make_stage2()
    make_bert_dropout()
