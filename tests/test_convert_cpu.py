"""b200mm.convert(model): the injection entry of SURVEY.md §8b — swaps the reference's VisionTransformer / BertModel inside an already
built reference CNCLIP (loaded unmodified from /root/reference when present) and keeps its forward results. Runs over the torch
stand-ins of the kernels (tests/emulated_ops.py); skipped where the reference tree does not exist (the GPU box)."""
import pytest
import torch

from oracle import ref_loader
from tests import emulated_ops

BF = torch.bfloat16


def rel_l2(got, ref):
    got, ref = got.detach().float(), ref.detach().float()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-12))


@pytest.mark.skipif(not ref_loader.available(), reason="needs the reference tree")
def test_convert_swaps_reference_towers_and_keeps_results():
    import b200mm

    cfg = dict(embed_dim=32, image_resolution=32, vision_layers=2, vision_width=64, vision_patch_size=8, vocab_size=21128,
               text_attention_probs_dropout_prob=0.0, text_hidden_act="gelu", text_hidden_dropout_prob=0.0, text_hidden_size=64,
               text_initializer_range=0.02, text_intermediate_size=256, text_max_position_embeddings=64, text_num_attention_heads=2,
               text_num_hidden_layers=2, text_type_vocab_size=2, vision_head_width=32)
    ref = ref_loader.build_cnclip(cfg, seed=0).eval()
    g = torch.Generator().manual_seed(3)
    image = torch.randn(4, 3, 32, 32, generator=g)
    text = torch.randint(1, 21128, (4, 12), generator=g)
    text[:, 0] = 101
    text[1, 8:] = 0
    with torch.no_grad():
        ref_img, ref_txt, ref_logits, _ = ref(image, text)
    sd_before = {k: v.clone() for k, v in ref.state_dict().items()}
    out = b200mm.convert(ref)
    assert out is ref and sorted(ref._b200mm_converted) == ["bert", "visual"]
    assert type(ref.visual).__module__.startswith("b200mm") and type(ref.bert).__module__.startswith("b200mm")
    assert ref.visual.conv1.weight.dtype == BF and ref.logit_scale.dtype == torch.float32  # only the swapped modules are cast
    assert set(ref.state_dict()) == set(sd_before)
    for k, v in ref.state_dict().items():
        assert torch.equal(v.float(), sd_before[k].to(v.dtype).float()), k
    b200mm.convert(ref)  # idempotent: b200mm modules are not touched again
    assert ref._b200mm_converted == []
    ref = ref.to(BF)  # the rest of the model (text_projection, logit_scale) follows, as in a bf16 training run
    with emulated_ops.patched(), torch.no_grad():
        # the REFERENCE's own CNCLIP.forward now drives the B200 towers (cn_model.py:198-226)
        img, txt, logits, _ = ref(image.to(BF), text)
    assert rel_l2(img, ref_img) < 2e-2 and rel_l2(txt, ref_txt) < 2e-2
    assert float((logits.float() - ref_logits).abs().max()) / float(ref.logit_scale.exp()) < 2e-2


def test_convert_leaves_foreign_models_alone():
    import b200mm

    m = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.LayerNorm(4))
    assert b200mm.convert(m) is m and m._b200mm_converted == []


@pytest.mark.skipif(not ref_loader.available(), reason="needs the reference tree")
@pytest.mark.parametrize("xpos", [False, True])
def test_convert_swaps_reference_beit3_and_vl_encoder(xpos):
    """The M²-Encoder family: an unmodified reference BEiT3 + stand-alone torchscale Encoder (VLMo.backbone / backbone_vl) inside a
    holder module are swapped in place; the converted modules keep the reference's keyword interface, so the lines of
    VLMo.infer_image / infer_text (vlmo_module.py:333-343,385-391) run unchanged on them and reproduce the reference's hiddens."""
    import copy

    import b200mm

    m2 = ref_loader.load_m2()
    torch.manual_seed(0)
    args = m2.EncoderConfig(img_size=32, patch_size=8, vocab_size=128, multiway=True, layernorm_embedding=False, normalize_output=True,
                            no_output_layer=True, encoder_embed_dim=64, encoder_attention_heads=2, encoder_layers=2, encoder_ffn_embed_dim=256,
                            checkpoint_activations=False, max_text_len=10, max_source_positions=24, xpos_rel_pos=xpos)
    holder = torch.nn.Module()
    holder.backbone = m2.BEiT3(args)
    vl = copy.copy(args)
    vl.encoder_layers = 1
    holder.backbone_vl = m2.Encoder(vl)
    holder.eval()
    g = torch.Generator().manual_seed(5)
    image = torch.randn(3, 3, 32, 32, generator=g)
    ids = torch.randint(1, 128, (3, 10), generator=g)
    pad = torch.zeros(3, 10, dtype=torch.long)
    pad[1, 7:] = 1
    with torch.no_grad():
        r_img = holder.backbone(visual_tokens=image)["encoder_out"]
        r_txt = holder.backbone(textual_tokens=ids, text_padding_position=pad)["encoder_out"]
        r_vl = holder.backbone_vl(src_tokens=None, token_embeddings=r_txt, encoder_padding_mask=pad, multiway_split_position=-1)["encoder_out"]
    keys = set(holder.state_dict())
    b200mm.convert(holder)
    assert sorted(holder._b200mm_converted) == ["backbone", "backbone_vl"] and set(holder.state_dict()) == keys
    assert type(holder.backbone).__module__.startswith("b200mm") and type(holder.backbone_vl).__module__.startswith("b200mm")
    valid = (pad == 0).unsqueeze(-1)
    with emulated_ops.patched(), torch.no_grad():
        img = holder.backbone(visual_tokens=image.to(BF))["encoder_out"]
        txt = holder.backbone(textual_tokens=ids, text_padding_position=pad)["encoder_out"]
        out_vl = holder.backbone_vl(src_tokens=None, token_embeddings=txt, encoder_padding_mask=pad, multiway_split_position=-1)["encoder_out"]
    assert rel_l2(img, r_img) < 1.5e-2
    assert rel_l2(txt.float() * valid, r_txt * valid) < 1.5e-2
    assert rel_l2(out_vl.float() * valid, r_vl * valid) < 2e-2


@pytest.mark.skipif(not ref_loader.available(), reason="needs the reference tree")
def test_convert_rejects_torchscale_configurations_off_the_path():
    import b200mm

    m2 = ref_loader.load_m2()
    args = m2.EncoderConfig(img_size=32, patch_size=8, vocab_size=64, multiway=True, no_output_layer=True, encoder_embed_dim=64,
                            encoder_attention_heads=2, encoder_layers=1, encoder_ffn_embed_dim=128, deepnorm=True)
    with pytest.raises(NotImplementedError, match="sub-LN"):
        b200mm.convert(m2.BEiT3(args))
