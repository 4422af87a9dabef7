"""Registered losses (antmmf `registry.register_loss` plug-in API, antmmf/common/registry.py:246-271; invoked as
loss(sample_list, model_output) -> tensor like antmmf/modules/losses/losses.py:152-164)."""
import torch
from torch import nn

from .contrastive import clip_contrastive_loss, mil_nce_loss
from .functional import RowNormFn
from .registry import registry

BF16 = torch.bfloat16


def _feat(x):
    return (x if x.dtype == BF16 else x.to(BF16)).contiguous()


@registry.register_loss("b200_clip_nce")
class B200ClipNCELoss(nn.Module):
    """Symmetric InfoNCE over the global batch. model_output needs `image_features`, `text_features` ([B, E]) and
    `logit_scale` (log-temperature parameter); features are L2-normalised here unless `normalized=True`."""

    def __init__(self, normalized=False, **params):
        super().__init__()
        self.normalized = normalized

    def forward(self, sample_list, model_output, *args, **kwargs):
        img, txt = _feat(model_output["image_features"]), _feat(model_output["text_features"])
        if not self.normalized:
            img, txt = RowNormFn.apply(img), RowNormFn.apply(txt)
        return clip_contrastive_loss(img, txt, model_output["logit_scale"])


@registry.register_loss("b200_mil_nce")
class B200MilNCELoss(nn.Module):
    """get_mil_nce_loss (prj/base_vtp/roi_univl/univl/model/univl_video_ret.py:146-197, n_clips = 1) on
    `video_features` / `text_features` [B, E]; the all-gather of both modalities is fused in."""

    def forward(self, sample_list, model_output, *args, **kwargs):
        return mil_nce_loss(_feat(model_output["video_features"]), _feat(model_output["text_features"]))
