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
    """get_mil_nce_loss (prj/base_vtp/roi_univl/univl/model/univl_video_ret.py:146-197) on `video_features` [B * n_clips, E]
    (clips of a video adjacent) / `text_features` [B, E]; `n_clips` from the constructor or model_output["n_clips"]; the all-gather of
    both modalities is fused in."""

    def __init__(self, n_clips=1, **params):
        super().__init__()
        self.n_clips = n_clips

    def forward(self, sample_list, model_output, *args, **kwargs):
        n = int(model_output.get("n_clips", self.n_clips)) if hasattr(model_output, "get") else self.n_clips
        return mil_nce_loss(_feat(model_output["video_features"]), _feat(model_output["text_features"]), n_clips=n)
