"""b200mm.convert(model) — swap the reference's ViT / BERT / CN-CLIP modules of an ALREADY BUILT model for their B200 equivalents
(SURVEY.md §8b "injection precedent": antmmf/utils/optim_utils.py:24-33,59-93 replaces nn.LayerNorm by FastLayerNorm with exactly this
recursive `named_children()` walk + `setattr(module, name, new_module)` + `load_state_dict(child.state_dict())`).

Matched by class name and structure (the reference classes are not importable from here):
  VisionTransformer  antmmf/modules/vision/backbone/clip/model.py:275-335    -> b200mm.modules.VisionTransformer
  BertModel          antmmf/modules/vision/backbone/clip/modeling_bert.py:421 -> b200mm.modules.BertModel
  BEiT3              prj/M2_Encoder/vlmo/torchscale/model/BEiT3.py:15         -> b200mm.modules.BEiT3   (multiway, sub-LN; args read from .args)
  Encoder            prj/M2_Encoder/vlmo/torchscale/architecture/encoder.py:171 (stand-alone, e.g. VLMo.backbone_vl) -> b200mm M2 Encoder
Parameters are copied by key (`load_state_dict`, strict), so the converted model produces the reference's results on the same inputs
within the bf16 bars of DESIGN.md §4; everything else in the model is left untouched.
"""
import torch
from torch import nn

from . import modules as M


def _vit_from(ref):
    width, patch = ref.conv1.weight.shape[0], ref.conv1.weight.shape[-1]
    layers = len(ref.transformer.resblocks)
    blk = ref.transformer.resblocks[0]
    heads = getattr(blk.attn, "num_heads", None) or getattr(blk, "n_head", None)
    out_dim = ref.proj.shape[1]
    new = M.VisionTransformer(input_resolution=ref.input_resolution, patch_size=patch, width=width, layers=layers, heads=heads, output_dim=out_dim)
    new.load_state_dict(ref.state_dict())
    return new


def _bert_from(ref):
    c = ref.config
    cfg = M.BertConfig(vocab_size_or_config_json_file=c.vocab_size, hidden_size=c.hidden_size, num_hidden_layers=c.num_hidden_layers,
                       num_attention_heads=c.num_attention_heads, intermediate_size=c.intermediate_size, hidden_act=c.hidden_act,
                       hidden_dropout_prob=c.hidden_dropout_prob, attention_probs_dropout_prob=c.attention_probs_dropout_prob,
                       max_position_embeddings=c.max_position_embeddings, type_vocab_size=c.type_vocab_size,
                       initializer_range=c.initializer_range, layer_norm_eps=c.layer_norm_eps)
    new = M.BertModel(cfg)
    new.load_state_dict(ref.state_dict())
    return new


def _m2_args_supported(a):
    bad = []
    if not getattr(a, "multiway", False):
        bad.append("multiway=False")
    if not getattr(a, "encoder_normalize_before", True) or not getattr(a, "subln", True) or getattr(a, "deepnorm", False):
        bad.append("not pre-LN + sub-LN")
    if getattr(a, "moe_freq", 0) or getattr(a, "rel_pos_buckets", 0) or getattr(a, "layernorm_embedding", False):
        bad.append("MoE / relative position bias / embedding LayerNorm")
    if not getattr(a, "no_output_layer", False) or getattr(a, "share_layer", False) or getattr(a, "share_attn", False):
        bad.append("output projection / shared layers")
    if bad:
        raise NotImplementedError("b200mm.convert: torchscale encoder configuration not on the hot path: " + ", ".join(bad))


def _beit3_from(ref):
    from .modules import beit3 as B3

    a = ref.args
    _m2_args_supported(a)
    new = B3.BEiT3(img_size=a.img_size, patch_size=a.patch_size, in_chans=a.in_chans, vocab_size=a.vocab_size,
                   encoder_embed_dim=a.encoder_embed_dim, encoder_attention_heads=a.encoder_attention_heads,
                   encoder_ffn_embed_dim=a.encoder_ffn_embed_dim, encoder_layers=len(ref.encoder.layers),
                   max_source_positions=a.max_source_positions, layernorm_eps=a.layernorm_eps, xpos_rel_pos=a.xpos_rel_pos,
                   xpos_scale_base=a.xpos_scale_base)
    new.load_state_dict(ref.state_dict())
    return new


def _m2_encoder_from(ref):
    from .modules import beit3 as B3

    a = ref.args
    _m2_args_supported(a)
    if ref.embed_positions is not None or ref.embed_tokens is not None:
        raise NotImplementedError("b200mm.convert: a torchscale Encoder with its own embeddings is converted through its BEiT3 owner")
    new = B3.Encoder(a.encoder_embed_dim, a.encoder_attention_heads, a.encoder_ffn_embed_dim, len(ref.layers), a.layernorm_eps,
                     xpos_rel_pos=a.xpos_rel_pos, xpos_scale_base=a.xpos_scale_base)
    new.load_state_dict(ref.state_dict())
    return new


def _is_ref(module, name):
    return type(module).__name__ == name and not type(module).__module__.startswith("b200mm")


def _convert_one(child):
    if _is_ref(child, "VisionTransformer") and hasattr(child, "conv1") and hasattr(child, "transformer") and hasattr(child, "class_embedding"):
        return _vit_from(child)
    if _is_ref(child, "BertModel") and hasattr(child, "embeddings") and hasattr(child, "encoder") and hasattr(child, "config"):
        return _bert_from(child)
    if _is_ref(child, "BEiT3") and hasattr(child, "vision_embed") and hasattr(child, "text_embed") and hasattr(child, "args"):
        return _beit3_from(child)
    if _is_ref(child, "Encoder") and hasattr(child, "layers") and hasattr(child, "args") and hasattr(child, "embed_positions"):
        return _m2_encoder_from(child)
    return None


def convert(model: nn.Module, dtype=torch.bfloat16, verbose=False):
    """In-place: returns `model` with every reference VisionTransformer / BertModel below it replaced (and `model` itself replaced if it
    is one of them — use the return value). New modules take the old module's device; floating parameters of the NEW modules are cast to
    `dtype` (None = keep fp32 master weights; the kernels cast per call)."""
    top = _convert_one(model)
    if top is not None:
        return _finish(top, model, dtype)
    replaced = []

    def walk(module, prefix):
        for name, child in list(module.named_children()):
            new = _convert_one(child)
            if new is not None:
                setattr(module, name, _finish(new, child, dtype))
                replaced.append(prefix + name)
            else:
                walk(child, prefix + name + ".")

    walk(model, "")
    if verbose:
        print("b200mm.convert: replaced", replaced)
    model._b200mm_converted = replaced
    return model


def _finish(new, old, dtype):
    p = next(old.parameters(), None)
    if p is not None:
        new = new.to(p.device)
    if dtype is not None:
        new = new.to(dtype)
    new.train(old.training)
    return new
