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
