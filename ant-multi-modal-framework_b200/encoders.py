"""Registry-compatible B200 encoders: same constructor kwargs, forward signatures, return structures, attributes and
state-dict keys as the reference classes they replace.

  B200VitImageEncoder  <- VitImageEncoder      prj/base_vtp/roi_univl/univl/model/clip_visual_encoder.py:15-94
  B200RobertBertEncoder <- RobertBertEncoder   prj/base_vtp/roi_univl/univl/model/clip_text_encoder.py:131-263

`install_as_reference_names()` additionally registers them under the reference class names so that existing YAML
(`image_encoder: {type: VitImageEncoder, ...}`) selects the B200 path with no config change.
"""
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import functional as Fn
from .modules.bert import BertConfig, BertModel
from .modules.vit import VisionTransformer, _bf16
from .registry import ModuleRegistry, TextEncoder, VisualEncoder


def _load_cnclip_checkpoint(path):
    if not os.path.isfile(path):
        raise RuntimeError(f"b200mm: checkpoint {path!r} not found (no download is attempted; pass a local file or pretrained=False)")
    sd = torch.load(path, map_location="cpu")
    return sd["state_dict"] if "state_dict" in sd else sd


@VisualEncoder.register()
class B200VitImageEncoder(nn.Module):
    def __init__(self, model_name: str, input_resolution: int, patch_size: int, width: int, layers: int, out_dim: int, head_width=64,
                 pretrained=True, is_proj=True):
        super().__init__()
        if not is_proj:
            raise NotImplementedError("b200mm VitImageEncoder: is_proj=False (no projection) is not used on this path")
        self.visual = VisionTransformer(input_resolution=input_resolution, patch_size=patch_size, width=width, layers=layers,
                                        heads=width // head_width, output_dim=out_dim)
        self.out_dim = out_dim
        if pretrained:
            self.load_pretrained(model_name)

    def load_pretrained(self, path):
        """Copy `visual.*` entries of a CN-CLIP checkpoint by key (clip_visual_encoder.py:46-71)."""
        sd = _load_cnclip_checkpoint(path)
        own = self.visual.state_dict()
        for k, v in sd.items():
            k = k[len("module."):] if k.startswith("module.") else k
            if k.startswith("visual.") and k[len("visual."):] in own:
                own[k[len("visual."):]].copy_(v)

    def forward(self, image, image_mask):
        """image [b, N, C, H, W]; image_mask [b, N, H, W] bool (True = padding) -> dict(grid_feature [b, N, out_dim, 1, 1],
        grid_mask [b, N, 1, 1], grid_feature_with_pos=None)   (clip_visual_encoder.py:73-94)"""
        _B, _T, _C, _H, _W = image.shape
        feat = self.visual(image.reshape(_B * _T, _C, _H, _W))
        feat = feat.view(_B, _T, self.out_dim).unsqueeze(-1).unsqueeze(-1)
        mask = F.interpolate(image_mask.float(), size=(1, 1)).to(torch.bool)
        return dict(grid_feature=feat, grid_mask=mask, grid_feature_with_pos=None)


class _BertModel2(BertModel):
    """BertModel2.forward of clip_text_encoder.py:68-128: returns (sequence_output, sequence_output[:, 0])."""

    def forward(self, input_ids, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None):
        seq = super().forward(input_ids, attention_mask, token_type_ids, position_ids, head_mask)[0]
        return seq, seq[:, 0, :]


@TextEncoder.register()
class B200RobertBertEncoder(nn.Module):
    def __init__(self, model_name: str = "ViT-B-16", pretrained: bool = True, num_segments: int = None, model_type: str = "bert",
                 bert_model_name: str = "roberta_chinese_base", hidden_size: int = 768, intermediate_size: int = 3072,
                 num_hidden_layers: int = 12, start_hidden_layer: int = 0, num_attention_heads: int = 12, output_attentions: bool = False,
                 output_hidden_states: bool = False, vocab_size: int = 30522, gradient_checkpointing: bool = False, type_vocab_size: int = 2,
                 max_position_embeddings: int = 512, hidden_act: str = "gelu", hidden_dropout_prob: float = 0.1,
                 attention_probs_dropout_prob: float = 0.1, initializer_range: float = 0.02, layer_norm_eps: float = 1e-6,
                 is_proj: bool = True, out_dim: int = 768):
        super().__init__()
        self.bert_config = BertConfig(vocab_size_or_config_json_file=vocab_size, hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
                                      num_attention_heads=num_attention_heads, intermediate_size=intermediate_size, hidden_act=hidden_act,
                                      hidden_dropout_prob=hidden_dropout_prob, attention_probs_dropout_prob=attention_probs_dropout_prob,
                                      max_position_embeddings=max_position_embeddings, type_vocab_size=type_vocab_size,
                                      initializer_range=initializer_range, layer_norm_eps=1e-12,  # the reference pins 1e-12 (:176)
                                      output_attentions=output_attentions, output_hidden_states=output_hidden_states)
        module = _BertModel2(self.bert_config)
        self.encoder = module.encoder
        self.embeddings = module.embeddings
        self.module = module
        self.out_dim = out_dim
        self.num_segments = num_segments
        self._init_segment_embeddings()
        if gradient_checkpointing:
            module.set_grad_checkpointing(True)
        self.text_projection = nn.Parameter(torch.randn(hidden_size, out_dim) * hidden_size ** -0.5) if is_proj else None
        if pretrained:
            self.load_pretrained(model_name)

    def _init_segment_embeddings(self):
        """clip_text_encoder.py:229-246: widen token_type_embeddings to num_segments rows."""
        if self.num_segments is None or self.num_segments == self.embeddings.token_type_embeddings.num_embeddings:
            return
        old = self.embeddings.token_type_embeddings
        new = nn.Embedding(self.num_segments, self.bert_config.hidden_size)
        new.weight.data[:2].copy_(old.weight.data)
        for idx in range(2, self.num_segments - 1):
            new.weight.data[idx].copy_(old.weight.data.mean(dim=0))
        self.embeddings.token_type_embeddings = new

    def load_pretrained(self, path):
        """Copy `bert.*` and `text_projection` of a CN-CLIP checkpoint by key (clip_text_encoder.py:194-227)."""
        sd = _load_cnclip_checkpoint(path)
        own = self.module.state_dict()
        for k, v in sd.items():
            k = k[len("module."):] if k.startswith("module.") else k
            if k.startswith("bert.") and k[len("bert."):] in own:
                own[k[len("bert."):]].copy_(v)
            elif k == "text_projection" and self.text_projection is not None and v.shape == self.text_projection.shape:
                self.text_projection.data.copy_(v)

    def forward(self, input_ids, attention_mask, token_type_ids=None, position_ids=None, head_mask=None, output_attentions=False):
        if output_attentions:
            raise NotImplementedError("b200mm RobertBertEncoder: attention probabilities are never materialised")
        seq, cls = self.module(input_ids, attention_mask, token_type_ids, position_ids, head_mask)
        if self.text_projection is None:
            return seq, cls
        B, L, Hd = seq.shape
        pooled = Fn.ClsHeadFn.apply(seq.reshape(B * L, Hd), None, None, _bf16(self.text_projection), B, L, 0.0)
        return seq, pooled


def install_as_reference_names():
    """Make `type: VitImageEncoder` / `type: RobertBertEncoder` in existing AntMMF YAML resolve to the B200 classes."""
    ModuleRegistry.__register_module__["VitImageEncoder"] = B200VitImageEncoder
    ModuleRegistry.__register_module__["RobertBertEncoder"] = B200RobertBertEncoder
