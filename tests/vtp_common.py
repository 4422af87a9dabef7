"""Shared by the CPU host-logic test and the GPU parity test of b200mm.vtp: a tiny base_vtp configuration, seeded inputs in the
reference's batch-dict layout, and the ORACLE composition of the same forward pass out of oracle/restated.py pieces (each pinned to
reference-generated golden vectors on its own: ViT / BERT a1-a8, frame pooling a11, cross input a12, MIL-NCE a15, stage 2 f3)."""
import torch
import torch.nn.functional as F

from oracle import restated

BF = torch.bfloat16
HID, HEADS, VIT_W, VIT_HEADS = 64, 2, 64, 2


def make_config(stage="stage1+stage2", hard=True, with_moco=False, re_sample="top_k", re_weight="median"):
    return dict(
        training_stage=stage, arch_type="clip", hidden_size=HID, with_moco=with_moco, hard_example_mining=hard, re_sample_method=re_sample,
        re_weight_method=re_weight, K=256, M=0.99, T=0.05,
        image_encoder=dict(type="B200VitImageEncoder", params=dict(model_name="-", input_resolution=32, patch_size=8, width=VIT_W, layers=1,
                                                                   out_dim=HID, head_width=VIT_W // VIT_HEADS, pretrained=False)),
        text_encoder=dict(type="B200RobertBertEncoder", params=dict(pretrained=False, hidden_size=HID, intermediate_size=128, num_hidden_layers=2,
                                                                    num_attention_heads=HEADS, vocab_size=200, max_position_embeddings=64,
                                                                    out_dim=HID, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)),
    )


def make_batch(B=6, n_clips=1, n_frames=2, L=10, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, n_clips * n_frames, 3, 32, 32, generator=g).to(BF)
    pad = torch.zeros(B, n_clips * n_frames, 32, 32, dtype=torch.bool)
    pad[1, -1] = True  # one fully padded frame
    ids = torch.randint(1, 200, (B, L), generator=g)
    ids[:, 0] = 101
    mask = torch.ones(B, L, dtype=torch.long)
    mask[2, 6:] = 0
    ids = ids * mask
    img_input = dict(image_data=image.to(device), image_pad_mask=pad.to(device), image_n_clips=[n_clips] * B, image_num_frames=[n_frames] * B)
    cap_input = dict(caption_raw_input_ids=ids.to(device), caption_input_ids=ids.to(device), caption_input_mask=mask.to(device))
    return img_input, cap_input


def randomize(model, seed=1):
    """Random-init transformers map every input to nearly the same embedding; widen the input-dependent parts so that scores differ
    across pairs (patch / word embeddings up, class / position embeddings down)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif "similarity_dense" in n:
                p.normal_(0.0, 0.1, generator=g)
            elif "conv1" in n:
                p.normal_(0.0, 0.3, generator=g)
            elif "word_embeddings" in n:
                p.normal_(0.0, 1.0, generator=g)
            elif "class_embedding" in n or "positional_embedding" in n or "position_embeddings" in n:
                p.mul_(0.1)


def oracle_forward(model, img_input, cap_input, config, chosen=None):
    """fp32 CPU oracle of B200VideoTextRetrieval.forward on the model's own (bf16-rounded) parameters.
    Returns (dict of losses / matrices, dict name -> leaf tensor for gradient comparison)."""
    leaves = {}

    def sd_of(mod, pfx):
        out = {}
        for k, v in mod.state_dict().items():
            t = v.detach().float().cpu()
            if torch.is_floating_point(t):
                t = t.requires_grad_()
                leaves[pfx + k] = t
            out[k] = t
        return out

    vit = sd_of(model.module.img_encoder.visual, "module.img_encoder.visual.")
    bert = sd_of(model.module.text_encoder.module, "module.text_encoder.module.")
    proj = model.module.text_encoder.text_projection.detach().float().cpu().requires_grad_()
    leaves["module.text_encoder.text_projection"] = proj
    image = img_input["image_data"].float().cpu()
    pad = img_input["image_pad_mask"].cpu()
    n_clips, n_frames = img_input["image_n_clips"][0], img_input["image_num_frames"][0]
    ids, mask = cap_input["caption_raw_input_ids"].cpu(), cap_input["caption_input_mask"].cpu()
    B = ids.shape[0]
    # --- a9/a11: frames through the ViT, masked mean per clip (univl_video_base.py:76-95), L2-normalise
    feat = restated.vit_forward(vit, image.reshape(-1, 3, 32, 32), VIT_HEADS).view(B * n_clips, n_frames, HID)
    keep = (~F.interpolate(pad.float(), size=(1, 1)).bool()).view(B * n_clips, n_frames, 1).float()
    clip = (feat * keep).sum(1) / keep.sum(1)
    clip_n = F.normalize(clip, dim=-1)
    # --- a10: text tower, CLS @ text_projection, normalise
    seq = restated.bert_forward(bert, ids, mask, HEADS)
    text_n = F.normalize(seq[:, 0] @ proj, dim=-1)
    out = {"text_n": text_n, "clip_n": clip_n}
    l1 = restated.l1_simi_matrix(text_n, clip_n, n_clips)                       # [B, B, n]
    out["l1_loss"] = restated.mil_nce_clips(l1)
    out["l1_simi"] = l1.logsumexp(-1).detach()
    if "stage2" in config["training_stage"]:
        sd2 = {"cross_encoder." + k[len("encoder."):]: v for k, v in bert.items() if k.startswith("encoder.")}
        sd2["text_projection"] = proj
        for k, v in model.similarity_dense.state_dict().items():
            t = v.detach().float().cpu().requires_grad_()
            leaves["similarity_dense." + k] = t
            sd2["similarity_dense." + k] = t
        # --- a12: cross inputs (univl_video_base.py:168-209)
        cap_embed = restated.bert_embeddings(bert, "embeddings.", ids, torch.zeros_like(ids))
        sep = bert["embeddings.word_embeddings.weight"][torch.full((B,), 102)].unsqueeze(1)
        vis_tokens = clip.view(B, n_clips, HID)
        vis_embed = restated.bert_embeddings(bert, "embeddings.", None, torch.ones(B, n_clips + 1, dtype=torch.long),
                                             inputs_embeds=torch.cat([vis_tokens, sep], 1))
        vis_mask = torch.ones(B, n_clips + 1, dtype=torch.long)
        if config["hard_example_mining"]:
            if chosen is None:
                chosen = restated.hard_mining_indices(out["l1_simi"], 0, B, config["re_sample_method"])
            l2 = restated.cross_similarity_hard_mining(sd2, cap_embed, mask, vis_embed, vis_mask, chosen, HEADS)
            w = restated.hard_mining_weights(torch.diagonal(out["l1_simi"]), config["re_sample_method"]) if config["re_weight_method"] == "median" else None
        else:
            l2 = restated.cross_similarity(sd2, cap_embed, mask, vis_embed, vis_mask, HEADS)
            w = None
        out["l2_simi"] = l2
        out["l2_loss"] = restated.mil_nce_matrix(l2, w)
        out["chosen"] = chosen
    return out, leaves
