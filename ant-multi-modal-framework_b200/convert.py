"""b200mm.convert(model) — swap the reference's ViT / BERT / CN-CLIP modules of an ALREADY BUILT model for their B200 equivalents
(SURVEY.md §8b "injection precedent": antmmf/utils/optim_utils.py:24-33,59-93 replaces nn.LayerNorm by FastLayerNorm with exactly this
recursive `named_children()` walk + `setattr(module, name, new_module)` + `load_state_dict(child.state_dict())`).

Matched by class name and structure (the reference classes are not importable from here):
  VisionTransformer  antmmf/modules/vision/backbone/clip/model.py:275-335    -> b200mm.modules.VisionTransformer
  BertModel          antmmf/modules/vision/backbone/clip/modeling_bert.py:421 -> b200mm.modules.BertModel
  CNCLIP             antmmf/modules/vision/backbone/clip/cn_model.py:126      -> b200mm.modules.CNCLIP (adds the fused contrastive_loss)
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


def _is_ref(module, name):
    return type(module).__name__ == name and not type(module).__module__.startswith("b200mm")


def _convert_one(child):
    if _is_ref(child, "VisionTransformer") and hasattr(child, "conv1") and hasattr(child, "transformer") and hasattr(child, "class_embedding"):
        return _vit_from(child)
    if _is_ref(child, "BertModel") and hasattr(child, "embeddings") and hasattr(child, "encoder") and hasattr(child, "config"):
        return _bert_from(child)
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
