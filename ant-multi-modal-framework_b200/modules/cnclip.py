"""B200-native CN-CLIP dual encoder — drop-in for antmmf/modules/vision/backbone/clip/cn_model.py:124-226.

Same constructor kwargs (CONFIGS table), methods (encode_image, encode_text, forward) and parameter names
(visual.*, bert.*, text_projection, logit_scale), so checkpoints copy by key (cn_model.py:308-316).
`contrastive_loss` is the fused path (similarity + symmetric InfoNCE without materialising the logits).
"""
import numpy as np
import torch
from torch import nn

from .. import functional as Fn
from ..contrastive import clip_contrastive_loss
from .bert import BertConfig, BertModel
from .vit import VisionTransformer, _bf16

_TEXT_BASE = dict(vocab_size=21128, text_attention_probs_dropout_prob=0.1, text_hidden_act="gelu", text_hidden_dropout_prob=0.1,
                  text_hidden_size=768, text_initializer_range=0.02, text_intermediate_size=3072, text_max_position_embeddings=512,
                  text_num_attention_heads=12, text_num_hidden_layers=12, text_type_vocab_size=2)
_TEXT_LARGE = dict(_TEXT_BASE, text_hidden_size=1024, text_intermediate_size=4096, text_num_attention_heads=16, text_num_hidden_layers=24)

# size table of cn_model.py:20-114 (the ViT entries; the RN50 tower is not on this path)
CONFIGS = {
    "ViT-B-16": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16, **_TEXT_BASE),
    "ViT-L-14": dict(embed_dim=768, image_resolution=224, vision_layers=24, vision_width=1024, vision_head_width=64, vision_patch_size=14, **_TEXT_BASE),
    "ViT-L-14-336": dict(embed_dim=768, image_resolution=336, vision_layers=24, vision_width=1024, vision_head_width=64, vision_patch_size=14, **_TEXT_BASE),
    "ViT-H-14": dict(embed_dim=1024, image_resolution=224, vision_layers=32, vision_width=1280, vision_head_width=80, vision_patch_size=14, **_TEXT_LARGE),
}

PAD_ID = 0  # FullTokenizer().vocab["[PAD]"] in the bundled Chinese vocab (cn_model.py:205)


class CNCLIP(nn.Module):
    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, vocab_size,
                 text_attention_probs_dropout_prob, text_hidden_act, text_hidden_dropout_prob, text_hidden_size, text_initializer_range,
                 text_intermediate_size, text_max_position_embeddings, text_num_attention_heads, text_num_hidden_layers,
                 text_type_vocab_size, vision_head_width=64, model_type="all"):
        super().__init__()
        if model_type in ["vision", "all"]:
            if isinstance(vision_layers, (tuple, list)):
                raise NotImplementedError("b200mm CNCLIP: the ModifiedResNet tower is outside the ViT+BERT hot path")
            self.visual = VisionTransformer(input_resolution=image_resolution, patch_size=vision_patch_size, width=vision_width,
                                            layers=vision_layers, heads=vision_width // vision_head_width, output_dim=embed_dim)
        if model_type in ["language", "all"]:
            self.bert_config = BertConfig(vocab_size_or_config_json_file=vocab_size, hidden_size=text_hidden_size,
                                          num_hidden_layers=text_num_hidden_layers, num_attention_heads=text_num_attention_heads,
                                          intermediate_size=text_intermediate_size, hidden_act=text_hidden_act,
                                          hidden_dropout_prob=text_hidden_dropout_prob,
                                          attention_probs_dropout_prob=text_attention_probs_dropout_prob,
                                          max_position_embeddings=text_max_position_embeddings, type_vocab_size=text_type_vocab_size,
                                          initializer_range=text_initializer_range, layer_norm_eps=1e-12)
            self.bert = BertModel(self.bert_config)
            # torch.empty in the reference (cn_model.py:190-192, filled by the checkpoint); initialised here like CLIP's
            self.text_projection = nn.Parameter(torch.randn(text_hidden_size, embed_dim) * text_hidden_size ** -0.5)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def set_grad_checkpointing(self, enable=True):
        self.visual.set_grad_checkpointing(enable)
        self.bert.set_grad_checkpointing(enable)

    def encode_image(self, image):
        return self.visual(image)

    def encode_text(self, text):
        attn_mask = text.ne(PAD_ID)
        seq = self.bert(text, attention_mask=attn_mask)[0]  # [B, L, H]
        B, L, Hd = seq.shape
        return Fn.ClsHeadFn.apply(seq.reshape(B * L, Hd), None, None, _bf16(self.text_projection), B, L, 0.0)

    def encode_normalized(self, image, text):
        return Fn.RowNormFn.apply(self.encode_image(image)), Fn.RowNormFn.apply(self.encode_text(text))

    def forward(self, image, text):
        """(image_features, text_features, logits_per_image, logits_per_text) as cn_model.py:212-226.
        The logits are materialised here for API compatibility; training should call `contrastive_loss`."""
        img, txt = self.encode_normalized(image, text)
        logits = _ScaledSimFn.apply(img, txt, self.logit_scale)
        return img, txt, logits, logits.t()

    def contrastive_loss(self, image, text, group=None):
        """Symmetric InfoNCE of the (global) batch through the fused similarity/log-softmax kernels."""
        img, txt = self.encode_normalized(image, text)
        return clip_contrastive_loss(img, txt, self.logit_scale, group)


class _ScaledSimFn(torch.autograd.Function):
    """logits = exp(logit_scale) * I · T^T  (f32 [B, B]); small-batch convenience for CNCLIP.forward."""

    @staticmethod
    def forward(ctx, img, txt, log_scale):
        from .. import ops

        Bt = txt.shape[0]
        pad = (-Bt) % 8
        txt_p = torch.cat([txt, txt.new_zeros(pad, txt.shape[1])]) if pad else txt
        alpha = float(torch.exp(log_scale.detach().float()))
        logits = ops.gemm(img, txt_p, alpha=alpha, out_f32=True)[:, :Bt]
        ctx.save_for_backward(img, txt, logits)
        ctx.meta = (alpha, log_scale.dtype)
        return logits

    @staticmethod
    def backward(ctx, g):
        from .. import ops

        img, txt, logits = ctx.saved_tensors
        alpha, sdt = ctx.meta
        Bi, Bt = logits.shape
        pad = (-Bt) % 8
        gp = torch.zeros((Bi, Bt + pad), device=g.device, dtype=torch.bfloat16)
        gp[:, :Bt] = g
        txt_p = torch.cat([txt, txt.new_zeros(pad, txt.shape[1])]) if pad else txt
        d_img = ops.gemm(gp, txt_p, b_mn=True, alpha=alpha)
        d_txt = ops.gemm(gp, img, a_mn=True, b_mn=True, alpha=alpha)[:Bt]
        d_ls = (g.float() * logits).sum().to(sdt)
        return d_img, d_txt.contiguous(), d_ls


# ----------------------------------------------------------------------------------------------------------------------
# module-level loader API of cn_model.py:229-420 (build_model / load / available_models, the two single-tower wrappers)
# ----------------------------------------------------------------------------------------------------------------------
def available_models():
    """Names of the CN-CLIP sizes this path builds (cn_model.py:352-354; the RN50 tower is not on the path)."""
    return list(CONFIGS.keys())


def build_model(config: dict, state_dict: dict = None):
    """cn_model.py:308-316: CNCLIP(**config), parameters copied by key (a leading `module.` of DDP checkpoints is dropped), eval mode."""
    model = CNCLIP(**config)
    if state_dict is not None:
        own = model.state_dict()
        clean = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        missing = [k for k in own if k not in clean]
        if missing:
            raise KeyError(f"b200mm build_model: checkpoint lacks {len(missing)} parameters, e.g. {missing[:3]}")
        for k in own:
            own[k].copy_(clean[k])
    return model.eval()


def load(name: str, config: dict = None, pretrained: bool = True, device="cpu", download_root: str = None):
    """cn_model.py:357-409 without the download step (there is no network on this path): `name` is a size from CONFIGS (with
    pretrained=False) or the path of a CN-CLIP checkpoint file ({"state_dict": ...}); `config` defaults to CONFIGS[name]."""
    import os

    if config is None:
        if name not in CONFIGS:
            raise RuntimeError(f"Model {name} not found; available models = {available_models()} (or pass `config` with a checkpoint path)")
        config = CONFIGS[name]
    if not pretrained:
        return build_model(config).to(device)
    if not os.path.isfile(name):
        raise RuntimeError(f"b200mm load: pretrained=True needs a local checkpoint path, got {name!r} (downloads are not supported); "
                           f"available sizes = {available_models()}")
    ckpt = torch.load(name, map_location="cpu")
    state_dict = ckpt["state_dict"] if isinstance(ckpt, dict) and "state_dict" in ckpt else ckpt
    return build_model(config, state_dict).to(device)


class CNCLIPImageEncoder(nn.Module):
    """cn_model.py:229-250: `forward(x) = model.encode_image(x)`."""

    def __init__(self, model_name, config=None, pretrained=True):
        super().__init__()
        self.model = load(model_name, config, pretrained)

    def forward(self, x):
        return self.model.encode_image(x)


class CNCLIPLanguageEncoder(nn.Module):
    """cn_model.py:253-273: `forward(text) = model.encode_text(text)`."""

    def __init__(self, model_name, config=None, pretrained=True):
        super().__init__()
        self.model = load(model_name, config, pretrained)

    def forward(self, text):
        return self.model.encode_text(text)
