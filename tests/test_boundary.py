"""CPU checks of the drop-in boundary: the C-ABI library exports every symbol include/b200mm.h declares, the Python
host mirrors the reference's registry / module interfaces, and state-dict keys equal the reference's."""
import ctypes
import os
import re

import pytest
import torch

import b200mm
from b200mm import _lib
from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "b200mm.h")).read()
    declared = set(re.findall(r"\b(b200mm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200mm.h but not exported by libb200mm.so"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().b200mm_version() >= 100


def test_no_cpu_fallback_and_loud_errors():
    with pytest.raises(b200mm.B200mmError, match="CUDA tensor"):
        b200mm.ops.layernorm_fwd(torch.zeros(4, 8, dtype=torch.bfloat16), torch.ones(8, dtype=torch.bfloat16),
                                 torch.zeros(8, dtype=torch.bfloat16), 1e-5)
    from b200mm.modules import CNCLIP

    m = CNCLIP(embed_dim=16, image_resolution=32, vision_layers=1, vision_width=64, vision_patch_size=8, vocab_size=128,
               text_attention_probs_dropout_prob=0.0, text_hidden_act="gelu", text_hidden_dropout_prob=0.0, text_hidden_size=64,
               text_initializer_range=0.02, text_intermediate_size=128, text_max_position_embeddings=32, text_num_attention_heads=2,
               text_num_hidden_layers=1, text_type_vocab_size=2)
    with pytest.raises(b200mm.B200mmError):
        m.to(torch.bfloat16).encode_image(torch.zeros(2, 3, 32, 32))


def test_registry_construction_and_signatures():
    import inspect

    from b200mm.encoders import B200RobertBertEncoder, B200VitImageEncoder, install_as_reference_names
    from b200mm.registry import ModuleRegistry, TextEncoder, VisualEncoder, registry

    v = VisualEncoder(dict(type="B200VitImageEncoder", params=dict(model_name="-", input_resolution=32, patch_size=8, width=64, layers=1,
                                                                  out_dim=16, head_width=32, pretrained=False)))
    assert isinstance(v.module, B200VitImageEncoder) and v.module.out_dim == 16
    t = TextEncoder(dict(type="B200RobertBertEncoder", params=dict(pretrained=False, hidden_size=64, intermediate_size=128,
                                                                  num_hidden_layers=2, num_attention_heads=2, vocab_size=100, out_dim=16)))
    mod = t.module
    assert isinstance(mod, B200RobertBertEncoder) and len(mod.encoder.layer) == 2
    assert mod.embeddings.word_embeddings.weight.shape == (100, 64) and mod.text_projection.shape == (64, 16)
    # the reference constructor argument names (clip_visual_encoder.py:17-28, clip_text_encoder.py:133-159)
    assert list(inspect.signature(B200VitImageEncoder.__init__).parameters)[1:] == [
        "model_name", "input_resolution", "patch_size", "width", "layers", "out_dim", "head_width", "pretrained", "is_proj"]
    want = ["model_name", "pretrained", "num_segments", "model_type", "bert_model_name", "hidden_size", "intermediate_size",
            "num_hidden_layers", "start_hidden_layer", "num_attention_heads", "output_attentions", "output_hidden_states", "vocab_size",
            "gradient_checkpointing", "type_vocab_size", "max_position_embeddings", "hidden_act", "hidden_dropout_prob",
            "attention_probs_dropout_prob", "initializer_range", "layer_norm_eps", "is_proj", "out_dim"]
    assert list(inspect.signature(B200RobertBertEncoder.__init__).parameters)[1:] == want
    with pytest.raises(ValueError):
        ModuleRegistry.get("NoSuchEncoder")
    install_as_reference_names()
    assert ModuleRegistry.get("VitImageEncoder") is B200VitImageEncoder
    assert registry.get_loss_class("b200_clip_nce") is not None and registry.get_loss_class("b200_mil_nce") is not None


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["ViT-B-16", "ViT-L-14", "ViT-H-14"])
def test_state_dict_keys_equal_reference(name):
    from b200mm.modules import CNCLIP, CONFIGS

    ns = ref_loader.load()
    assert CONFIGS[name] == {k: v for k, v in ns.cn_model.CONFIGS[name].items()}
    if name != "ViT-B-16":  # building the big towers twice is slow; the size table equality above covers them
        return
    ours = CNCLIP(**CONFIGS[name]).state_dict()
    ref = ref_loader.build_cnclip(name).state_dict()
    assert set(ours) == set(ref)
    for k in ours:
        assert ours[k].shape == ref[k].shape, k


def test_golden_state_dict_loads_strict(golden_dir):
    from b200mm.modules import CNCLIP

    fx = torch.load(os.path.join(golden_dir, "cnclip_tiny.pt"), weights_only=False)
    m = CNCLIP(**fx["config"])
    missing, unexpected = m.load_state_dict(fx["state_dict"], strict=True)
    assert not missing and not unexpected


@pytest.mark.parametrize("name", ["m2_tiny.pt", "m2_tiny_xpos.pt"])
def test_m2_encoder_state_dict_keys_match_reference(golden_dir, name):
    """M2Encoder carries the reference's parameter / buffer names and shapes (prj/M2_Encoder: BEiT3 + backbone_vl Encoder + ITC heads;
    with XPOS also `self_attn.xpos.scale`), checked against the state dicts of the unmodified reference classes in tests/golden/."""
    from b200mm.modules import M2Encoder

    fx = torch.load(os.path.join(golden_dir, name), weights_only=False)
    c = fx["config"]
    m = M2Encoder(image_size=c["img"], patch_size=c["patch"], vocab_size=c["vocab"], encoder_embed_dim=c["W"], encoder_attention_heads=c["heads"],
                  encoder_layers=c["layers"], beit3_vl_layers=c["vl_layers"], out_embed_dim=c["out_dim"], max_text_len=c["L"],
                  max_source_positions=c.get("max_source_positions", 1024), xpos_rel_pos=c.get("xpos", False))
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in fx["state_dict"].items()}
    assert set(ref) <= set(ours), sorted(set(ref) - set(ours))[:5]
    extra = set(ours) - set(ref)
    assert all(k.startswith(("norm.", "pooler.")) for k in extra), sorted(extra)[:5]  # VLMo's own unused members (vlmo_module.py:176-179)
    for k in ref:
        assert ours[k] == ref[k], (k, ours[k], ref[k])
    m.load_state_dict(fx["state_dict"], strict=False)
    if ref_loader.available():  # and against the live reference classes
        m2 = ref_loader.load_m2()
        args = m2.EncoderConfig(img_size=c["img"], patch_size=c["patch"], vocab_size=c["vocab"], multiway=True, no_output_layer=True,
                                encoder_embed_dim=c["W"], encoder_attention_heads=c["heads"], encoder_layers=c["layers"],
                                encoder_ffn_embed_dim=4 * c["W"], max_text_len=c["L"], xpos_rel_pos=c.get("xpos", False),
                                max_source_positions=c.get("max_source_positions", 1024))
        live = {"backbone." + k: tuple(v.shape) for k, v in m2.BEiT3(args).state_dict().items()}
        assert all(ours[k] == s for k, s in live.items())


def test_cnclip_loader_api(tmp_path, golden_dir):
    """build_model / load / available_models and the single-tower wrappers of cn_model.py:229-420 (no download step)."""
    from b200mm.modules import CNCLIPImageEncoder, CNCLIPLanguageEncoder, available_models, build_model, load

    assert available_models() == ["ViT-B-16", "ViT-L-14", "ViT-L-14-336", "ViT-H-14"]
    fx = torch.load(os.path.join(golden_dir, "cnclip_tiny.pt"), weights_only=False)
    ckpt = tmp_path / "tiny.pt"
    torch.save({"state_dict": {"module." + k: v for k, v in fx["state_dict"].items()}}, ckpt)  # a DDP-style checkpoint
    m = load(str(ckpt), fx["config"], pretrained=True)
    assert not m.training
    for k, v in m.state_dict().items():
        assert torch.equal(v, fx["state_dict"][k]), k
    assert isinstance(CNCLIPImageEncoder(str(ckpt), fx["config"]).model, type(m))
    assert isinstance(CNCLIPLanguageEncoder(str(ckpt), fx["config"]).model, type(m))
    fresh = build_model(fx["config"])
    assert set(fresh.state_dict()) == set(fx["state_dict"])
    with pytest.raises(RuntimeError, match="local checkpoint path"):
        load("ViT-B-16", pretrained=True)
    with pytest.raises(RuntimeError, match="not found"):
        load("RN50", pretrained=False)
    bad = dict(fx["state_dict"])
    bad.pop("logit_scale")
    with pytest.raises(KeyError):
        build_model(fx["config"], bad)
