"""Video/text embedding heads of base_vtp on the b200mm kernels — SURVEY.md §8 row a11.

Reference: UnivlVideoBase.forward_img_encoder / forward_text_encoder (prj/base_vtp/roi_univl/univl/model/univl_video_base.py:56-166):
the image encoder is run on every frame, frames of a clip are mean-pooled under the padding mask, an optional `img_proj` / `img_fc`
is applied, and clip / sentence embeddings are L2-normalised. Same arguments, same return dictionaries (`visual_embed`, `visual_mask`,
`visual_grid_shape`, `clip_feature`; `sequence_output`, `pooled_output`, `input_mask`, `words_importance`).
"""
import torch
from torch.autograd import Function

from . import functional as Fn
from . import ops
from .modules.vit import _bf16

BF16 = torch.bfloat16


class FramePoolFn(Function):
    """[R, P, E] bf16, pad [R, P] bool (True = padded) -> [R, E]: mean over the non-padded positions (univl_video_base.py:91-95)."""

    @staticmethod
    def forward(ctx, x, pad):
        y, inv, pad8 = ops.masked_mean_fwd(x, pad)
        ctx.save_for_backward(inv, pad8 if pad8 is not None else torch.empty(0, device=x.device, dtype=torch.uint8))
        ctx.meta = (x.shape[1], pad8 is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        inv, pad8 = ctx.saved_tensors
        P, has_pad = ctx.meta
        return ops.masked_mean_bwd(dy.to(BF16), pad8 if has_pad else None, inv, P), None


class _ProjFn(Function):
    """y = x @ W for the `img_proj` parameter ([out_dim, hidden], univl_video_base.py:31-35, :70-73) on the tcgen05 GEMM."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return ops.gemm(x, w, b_mn=True)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.to(BF16).contiguous()
        return ops.gemm(dy, w), ops.gemm(x, dy, a_mn=True, b_mn=True, out_f32=True).to(w.dtype)


def pool_clip_features(grid_feature, grid_mask, n_clips, n_frames):
    """grid_feature [b, n_clips*n_frames, c, h, w], grid_mask [b, n_clips*n_frames, h, w] (True = padded) -> clip_feature [b*n_clips, c].
    The reference flattens (frames, h, w) per channel and averages the unmasked positions; here the channel axis is innermost."""
    b, _, c, h, w = grid_feature.shape
    if h * w == 1:
        x = grid_feature.reshape(b * n_clips, n_frames, c)
    else:
        x = grid_feature.reshape(b * n_clips, n_frames, c, h * w).permute(0, 1, 3, 2).reshape(b * n_clips, n_frames * h * w, c)
    pad = grid_mask.reshape(b * n_clips, n_frames * h * w)
    return FramePoolFn.apply(_bf16(x).contiguous(), pad)


def forward_img_encoder(img_encoder, image_data, image_pad_mask, image_n_clips, image_num_frames, img_proj=None):
    """UnivlVideoBase.forward_img_encoder (:56-122)."""
    out = img_encoder(image_data, image_mask=image_pad_mask)
    grid_feature, grid_mask = out["grid_feature"], out["grid_mask"]
    if img_proj is not None:  # einsum("bnchw, cj -> bnjhw") of :70-73 as one GEMM over the channel axis
        b, n, c, h, w = grid_feature.shape
        flat = _bf16(grid_feature).permute(0, 1, 3, 4, 2).reshape(-1, c).contiguous()
        grid_feature = _ProjFn.apply(flat, _bf16(img_proj).contiguous()).view(b, n, h, w, -1).permute(0, 1, 4, 2, 3)
    grid_shape = grid_feature.shape[-2:]
    n_clips, n_frames = int(image_n_clips[0]), int(image_num_frames[0])
    bsz, c = grid_feature.size(0), grid_feature.size(2)
    clip_feature = pool_clip_features(grid_feature, grid_mask, n_clips, n_frames)
    clip_tokens = clip_feature.view(bsz, n_clips, c)
    clip_mask = torch.zeros((bsz, n_clips), device=clip_tokens.device, dtype=torch.bool)
    if "img_fc" in getattr(img_encoder, "_modules", {}):
        clip_feature = img_encoder.img_fc(clip_feature)
    clip_feature = Fn.RowNormFn.apply(_bf16(clip_feature))
    return dict(visual_embed=clip_tokens, visual_mask=clip_mask, visual_grid_shape=grid_shape, clip_feature=clip_feature)


def forward_text_encoder(text_encoder, input_ids, input_mask, arch_type="clip"):
    """UnivlVideoBase.forward_text_encoder (:124-166), arch_type "clip" (the "univl" branch needs attention probabilities,
    which this path never materialises)."""
    if arch_type != "clip":
        raise NotImplementedError("b200mm forward_text_encoder: only arch_type='clip' (no output_attentions)")
    sequence_output, pooled_output = text_encoder(input_ids=input_ids, attention_mask=input_mask)
    pooled_output = Fn.RowNormFn.apply(_bf16(pooled_output))
    return dict(sequence_output=sequence_output, pooled_output=pooled_output, input_mask=input_mask, words_importance=None)
